// Brute-force mutual-nearest-neighbour matching, float64 evaluation (algo 0).
// Reference: utils/matcher.py:227-234 -> skimage.feature.match_descriptors ->
// scipy.spatial.distance.cdist in float64 (sum of squared differences), np.argmin (first of ties),
// cross-check, strict `< max_distance`; pairs sorted by the first index.
//
// The [n,m] distance matrix is never written: 64x64 tiles are evaluated on chip (sum (a-b)^2 in
// float64 from the float32 inputs, exactly the arithmetic cdist performs) and reduced to per-tile
// row / column minima; a per-pair finalize CTA merges them in ascending index order (first of
// ties), applies the mutual check and the distance gate and compacts the survivors in order.
#include <math_constants.h>
#include "kb_common.cuh"

int kb_match_tc_run(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max, int m_max,
                    int D, double max_distance, int cross_check, int* pairs, double* dist, int* count, void* ws,
                    size_t ws_bytes, int phases, cudaStream_t st);
size_t kb_match_tc_workspace_bytes(int B, int n_max, int m_max, int D);

namespace {

constexpr int TM = 64, TN = 64, TK = 16, NT = 256;

struct Best {
    double d2;
    int idx;
    int pad;
};

struct MatchParams {
    const float* d0;     // [B,n_max,D]
    const float* d1;     // [B,m_max,D]
    const int* n0;
    const int* n1;
    Best* rowpart;       // [B,n_max,tiles_j]
    Best* colpart;       // [B,m_max,tiles_i]
    int B, n_max, m_max, D, tiles_i, tiles_j;
};

__global__ void __launch_bounds__(NT) dist_tile_kernel(MatchParams p) {
    __shared__ float As[TK][TM + 1];
    __shared__ float Bs[TK][TN + 1];
    __shared__ double red_v[TM][17];
    __shared__ int red_i[TM][17];
    const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    const int i0 = ti * TM, j0 = tj * TN;
    if (i0 >= n || j0 >= m) return;
    const float* A = p.d0 + (size_t)b * p.n_max * p.D;
    const float* Bm = p.d1 + (size_t)b * p.m_max * p.D;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;

    for (int k0 = 0; k0 < p.D; k0 += TK) {
        // stage TK columns of 64 rows of each operand (coalesced along k)
        for (int e = threadIdx.x; e < TM * TK; e += NT) {
            const int row = e / TK, k = e - row * TK;
            const int gi = i0 + row, gj = j0 + row, gk = k0 + k;
            As[k][row] = (gi < n && gk < p.D) ? A[(size_t)gi * p.D + gk] : 0.0f;
            Bs[k][row] = (gj < m && gk < p.D) ? Bm[(size_t)gj * p.D + gk] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            double a[4], bb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = (double)As[k][ty * 4 + r];
#pragma unroll
            for (int c = 0; c < 4; ++c) bb[c] = (double)Bs[k][tx * 4 + c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double d = a[r] - bb[c];
                    acc[r][c] = fma(d, d, acc[r][c]);
                }
        }
        __syncthreads();
    }
    // mask out-of-range entries
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (i0 + ty * 4 + r >= n || j0 + tx * 4 + c >= m) acc[r][c] = CUDART_INF;

    // ---- row minima over this tile's 64 columns (first of ties = smallest j) ------------------
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double best = acc[r][0];
        int bj = 0;
#pragma unroll
        for (int c = 1; c < 4; ++c)
            if (acc[r][c] < best) { best = acc[r][c]; bj = c; }
        red_v[ty * 4 + r][tx] = best;
        red_i[ty * 4 + r][tx] = j0 + tx * 4 + bj;
    }
    __syncthreads();
    if (threadIdx.x < TM) {
        const int row = threadIdx.x;
        double best = red_v[row][0];
        int bj = red_i[row][0];
        for (int g = 1; g < 16; ++g)
            if (red_v[row][g] < best) { best = red_v[row][g]; bj = red_i[row][g]; }
        if (i0 + row < n) {
            Best o; o.d2 = best; o.idx = bj; o.pad = 0;
            p.rowpart[((size_t)b * p.n_max + i0 + row) * p.tiles_j + tj] = o;
        }
    }
    __syncthreads();
    // ---- column minima over this tile's 64 rows (first of ties = smallest i) ------------------
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double best = acc[0][c];
        int bi = 0;
#pragma unroll
        for (int r = 1; r < 4; ++r)
            if (acc[r][c] < best) { best = acc[r][c]; bi = r; }
        red_v[tx * 4 + c][ty] = best;
        red_i[tx * 4 + c][ty] = i0 + ty * 4 + bi;
    }
    __syncthreads();
    if (threadIdx.x < TN) {
        const int col = threadIdx.x;
        double best = red_v[col][0];
        int bi = red_i[col][0];
        for (int g = 1; g < 16; ++g)
            if (red_v[col][g] < best) { best = red_v[col][g]; bi = red_i[col][g]; }
        if (j0 + col < m) {
            Best o; o.d2 = best; o.idx = bi; o.pad = 0;
            p.colpart[((size_t)b * p.m_max + j0 + col) * p.tiles_i + ti] = o;
        }
    }
}

struct FinalParams {
    const Best* rowpart;
    const Best* colpart;
    const int* n0;
    const int* n1;
    int* colbest;        // [B,m_max] argmin row of every column
    int* pairs;          // [B,n_max,2]
    double* dist;        // [B,n_max] or null
    int* count;          // [B]
    int B, n_max, m_max, tiles_i, tiles_j, cross_check;
    double max_distance;
};

__global__ void __launch_bounds__(1024) mnn_finalize_kernel(FinalParams p) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    if (n <= 0 || m <= 0) {
        if (threadIdx.x == 0) p.count[b] = 0;
        return;
    }
    const int used_ti = (n + TM - 1) / TM, used_tj = (m + TN - 1) / TN;
    int* colbest = p.colbest + (size_t)b * p.m_max;
    if (p.cross_check) {
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const Best* cp = p.colpart + ((size_t)b * p.m_max + j) * p.tiles_i;
            double best = cp[0].d2;
            int bi = cp[0].idx;
            for (int t = 1; t < used_ti; ++t)
                if (cp[t].d2 < best) { best = cp[t].d2; bi = cp[t].idx; }
            colbest[j] = bi;
        }
    }
    __syncthreads();
    int n_out = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool keep = false;
        int bj = 0;
        double d = 0.0;
        if (i < n) {
            const Best* rp = p.rowpart + ((size_t)b * p.n_max + i) * p.tiles_j;
            double best = rp[0].d2;
            bj = rp[0].idx;
            for (int t = 1; t < used_tj; ++t)
                if (rp[t].d2 < best) { best = rp[t].d2; bj = rp[t].idx; }
            d = sqrt(best);
            keep = (!p.cross_check || colbest[bj] == i) && (d < p.max_distance);
        }
        int tot;
        const int off = n_out + kb::block_exclusive_scan(keep ? 1 : 0, s_scan, &tot);
        if (keep) {
            p.pairs[((size_t)b * p.n_max + off) * 2 + 0] = i;
            p.pairs[((size_t)b * p.n_max + off) * 2 + 1] = bj;
            if (p.dist) p.dist[(size_t)b * p.n_max + off] = d;
        }
        n_out += tot;
    }
    if (threadIdx.x == 0) p.count[b] = n_out;
}

}  // namespace

extern "C" size_t kb_match_workspace_bytes(int B, int n_max, int m_max, int D, int algo) {
    if (B <= 0 || n_max <= 0 || m_max <= 0 || D <= 0) return 0;
    const size_t ti = (n_max + TM - 1) / TM, tj = (m_max + TN - 1) / TN;
    size_t f64 = kb_align_up((size_t)B * n_max * tj * sizeof(Best), 256) +
                 kb_align_up((size_t)B * m_max * ti * sizeof(Best), 256) +
                 kb_align_up((size_t)B * m_max * sizeof(int), 256) + 1024;
    const size_t tc = kb_match_tc_workspace_bytes(B, n_max, m_max, D);
    if (algo == 1) return tc;
    if (algo < 0) return tc > f64 ? tc : f64;
    return f64;
}

extern "C" int kb_match_mnn_phases(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max,
                                   int m_max, int D, double max_distance, int cross_check, int algo, int* pairs,
                                   double* dist, int* count, void* ws, size_t ws_bytes, int phases, kb_stream_t stream);

extern "C" int kb_match_mnn(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max,
                            int m_max, int D, double max_distance, int cross_check, int algo, int* pairs,
                            double* dist, int* count, void* ws, size_t ws_bytes, kb_stream_t stream) {
    return kb_match_mnn_phases(d0, d1, n0, n1, B, n_max, m_max, D, max_distance, cross_check, algo, pairs, dist, count, ws,
                               ws_bytes, 7, stream);
}

extern "C" int kb_match_mnn_phases(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max,
                                   int m_max, int D, double max_distance, int cross_check, int algo, int* pairs,
                                   double* dist, int* count, void* ws, size_t ws_bytes, int phases, kb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (phases <= 0 || phases > 15 || (phases & 9) == 9) return KB_ERR_BAD_ARG;      // bit 3: operands prepared by the sampler
    if (!d0 || !d1 || !pairs || !count || B <= 0 || n_max <= 0 || m_max <= 0 || D <= 0) return KB_ERR_BAD_ARG;
    if (algo < 0) algo = (D <= 256) ? 1 : 0;       // auto: tensor cores whenever the query tile fits on chip
    if (algo == 1)
        return kb_match_tc_run(d0, d1, n0, n1, B, n_max, m_max, D, max_distance, cross_check, pairs, dist, count,
                               ws, ws_bytes, phases, st);
    if (algo != 0) return KB_ERR_BAD_ARG;
    if (phases != 7) return KB_ERR_UNSUPPORTED;      // the float64 path has no separately timed / external parts
    if (B > 65535) return KB_ERR_UNSUPPORTED;
    MatchParams p;
    p.tiles_i = (n_max + TM - 1) / TM;
    p.tiles_j = (m_max + TN - 1) / TN;
    KbArena arena(ws, ws_bytes);
    p.rowpart = arena.take<Best>((size_t)B * n_max * p.tiles_j);
    p.colpart = arena.take<Best>((size_t)B * m_max * p.tiles_i);
    int* colbest = arena.take<int>((size_t)B * m_max);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    p.d0 = d0; p.d1 = d1; p.n0 = n0; p.n1 = n1; p.B = B; p.n_max = n_max; p.m_max = m_max; p.D = D;
    dim3 grid(p.tiles_j, p.tiles_i, B);
    dist_tile_kernel<<<grid, NT, 0, st>>>(p);
    KB_LAUNCH_CHECK();
    FinalParams f;
    f.rowpart = p.rowpart; f.colpart = p.colpart; f.n0 = n0; f.n1 = n1; f.colbest = colbest;
    f.pairs = pairs; f.dist = dist; f.count = count; f.B = B; f.n_max = n_max; f.m_max = m_max;
    f.tiles_i = p.tiles_i; f.tiles_j = p.tiles_j; f.cross_check = cross_check; f.max_distance = max_distance;
    mnn_finalize_kernel<<<B, 1024, 0, st>>>(f);
    KB_LAUNCH_CHECK();
    return KB_OK;
}
