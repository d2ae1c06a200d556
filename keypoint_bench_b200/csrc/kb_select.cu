// Border removal + threshold + raster-order compaction + top-k + min_score
// (reference: utils/extracter.py:164-190, 129-161, 217-220) and the batched detection() entry.
//
// One 1024-thread CTA per map.  Qualifying pixels are compacted in raster order as 64-bit
// priority keys (order-preserving score bits << 32 | ~raster), so "score desc, raster asc" is a
// plain descending integer order: a CTA-wide MSB radix select finds the top_k-th key, the
// survivors are bitonic-sorted in shared memory.  K <= top_k keeps raster order
// (extracter.py:217).
#include "kb_common.cuh"

int kb_nms_rounds_inplace(float* v, const float* src, int B, int H, int W, int nms_dist, int max_iter,
                          float min_value, int per_map, const int* active, const int* any_active, int* rounds,
                          void* ws, size_t ws_bytes, cudaStream_t st);
size_t kb_sparse_workspace_bytes(int B, int H, int W, int nms_dist, int top_k);
int kb_sparse_detect(const float* score, int B, int H, int W, int nms_dist, int border, float threshold,
                     float min_score, int top_k, float* xyp, int* raster, int* count, int* path,
                     int** need_fallback_out, int** any_fallback_out, void* ws, size_t ws_bytes, int phases,
                     cudaStream_t st);

namespace {

constexpr int NT = 1024;
constexpr int SORT_CAP = 8192;     // max top_k; also the "sort everything" shortcut size

struct SelectParams {
    const float* map;        // [B,H,W]
    uint64_t* cand;          // [B,cand_cap] raster-ordered priority keys
    int cand_cap;
    int B, H, W, border, top_k, cap;
    float threshold, min_score;
    float* xyp;              // [B,cap,3]
    int* raster;             // [B,cap]
    int* count;              // [B]
    int* total;              // [B] or null
    const int* skip;         // [B] or null: maps with skip[b] != want are left untouched
    int want;
    int* path;               // [B] or null: receives path_code for every map this launch handles
    int path_code;
};

__device__ __forceinline__ void emit_row(const SelectParams& p, int b, int slot, uint64_t key) {
    const uint32_t ras = kb::key_raster(key);
    const int row = ras / p.W, col = ras - row * p.W;
    float* o = p.xyp + ((size_t)b * p.cap + slot) * 3;
    o[0] = ((float)col + 0.5f) / (float)p.W;      // extracter.py:149,158 (one add, one IEEE divide)
    o[1] = ((float)row + 0.5f) / (float)p.H;
    o[2] = kb::key_score(key);
    p.raster[(size_t)b * p.cap + slot] = (int)ras;
}

__device__ void bitonic_sort_desc(uint64_t* a, int n /*pow2*/) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int l = i ^ j;
                if (l > i) {
                    uint64_t x = a[i], y = a[l];
                    bool desc = ((i & k) == 0);
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(NT) select_kernel(SelectParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw);     // SORT_CAP keys
    __shared__ int s_scan[33];
    __shared__ int s_hist[256];
    __shared__ unsigned long long s_prefix, s_mask;
    __shared__ int s_need, s_n;

    const int b = blockIdx.x;
    if (p.skip && p.skip[b] != p.want) return;      // e.g. maps the sparse path already certified
    const int H = p.H, W = p.W;
    const float* img = p.map + (size_t)b * H * W;
    uint64_t* cand = p.cand + (size_t)b * p.cand_cap;

    // ---- 1. raster-order compaction of (inside border) && (v > threshold) ---------------------
    // remove_border_points ZEROES the border (extracter.py:177-188) and the comparison is `> threshold`
    // (extracter.py:149): with a negative threshold the zeroed border pixels qualify too, as (x, y, 0) rows.
    const int y_lo = p.border, y_hi = H - p.border, x_lo = p.border, x_hi = W - p.border;
    const bool border_counts = p.threshold < 0.0f && p.border > 0;
    int K = 0;
    if ((y_hi > y_lo && x_hi > x_lo) || border_counts) {
        const int lo = border_counts ? 0 : y_lo * W, hi = border_counts ? H * W : y_hi * W;
        // 16 consecutive pixels per thread and step (four 16-byte loads when the row base allows it): one block-wide
        // scan per 16 K pixels instead of one per 4 K
        constexpr int PX = 16;
        const bool vec_ok = (reinterpret_cast<uintptr_t>(img) & 15u) == 0;
        for (int base = lo; base < hi; base += NT * PX) {
            const int i0 = base + threadIdx.x * PX;
            float v[PX];
            if (vec_ok && (i0 & 3) == 0 && i0 + PX <= hi) {
#pragma unroll
                for (int e = 0; e < PX / 4; ++e) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(img + i0) + e);
                    v[4 * e] = t.x; v[4 * e + 1] = t.y; v[4 * e + 2] = t.z; v[4 * e + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int e = 0; e < PX; ++e) v[e] = i0 + e < hi ? img[i0 + e] : 0.0f;
            }
            unsigned q = 0u;
            int row = i0 < hi ? i0 / W : 0;
            int col = i0 < hi ? i0 - row * W : 0;
#pragma unroll
            for (int e = 0; e < PX; ++e) {
                // extracter.py:149 (strict); columns inside the border only (rows too when the border counts)
                const bool inside = col >= x_lo && col < x_hi && (!border_counts || (row >= y_lo && row < y_hi));
                if (border_counts && !inside) v[e] = 0.0f;
                if (i0 + e < hi && (inside || border_counts) && v[e] > p.threshold) q |= 1u << e;
                if (++col == W) { col = 0; ++row; }
            }
            int tot;
            int off = K + kb::block_exclusive_scan(__popc(q), s_scan, &tot);
#pragma unroll
            for (int e = 0; e < PX; ++e) {
                if ((q >> e) & 1u) {
                    if (off < p.cand_cap) cand[off] = kb::priority_key(v[e], (uint32_t)(i0 + e));
                    ++off;
                }
            }
            K += tot;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && p.total) p.total[b] = K;
    if (threadIdx.x == 0 && p.path) p.path[b] = p.path_code;
    const int Kc = K < p.cand_cap ? K : p.cand_cap;     // candidates actually stored

    // ---- 2. raster mode (extracter.py:217: no sort unless K > top_k) --------------------------
    if (p.top_k <= 0 || K <= p.top_k) {
        int n_out = 0;
        for (int base = 0; base < Kc; base += NT) {
            const int i = base + threadIdx.x;
            uint64_t key = 0;
            bool keep = false;
            if (i < Kc) {
                key = cand[i];
                keep = !(p.min_score > 0.0f) || kb::key_score(key) > p.min_score;   // extracter.py:219-220
            }
            int tot;
            int off = n_out + kb::block_exclusive_scan(keep ? 1 : 0, s_scan, &tot);
            if (keep && off < p.cap) emit_row(p, b, off, key);
            n_out += tot;
        }
        if (threadIdx.x == 0) p.count[b] = n_out < p.cap ? n_out : p.cap;
        return;
    }

    // ---- 3. sorted mode: top_k largest keys, descending ---------------------------------------
    int n_sel;
    if (Kc <= SORT_CAP) {
        for (int i = threadIdx.x; i < Kc; i += NT) skeys[i] = cand[i];
        n_sel = Kc;
    } else {
        if (threadIdx.x == 0) { s_prefix = 0ull; s_mask = 0ull; s_need = p.top_k; s_n = 0; }
        __syncthreads();
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int i = threadIdx.x; i < 256; i += NT) s_hist[i] = 0;
            __syncthreads();
            const unsigned long long pre = s_prefix, msk = s_mask;
            for (int i = threadIdx.x; i < Kc; i += NT) {
                const uint64_t key = cand[i];
                if ((key & msk) == pre) atomicAdd(&s_hist[(int)((key >> shift) & 255ull)], 1);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int need = s_need, cum = 0, d = 255;
                for (; d > 0; --d) {
                    if (cum + s_hist[d] >= need) break;
                    cum += s_hist[d];
                }
                s_need = need - cum;
                s_prefix = pre | ((unsigned long long)d << shift);
                s_mask = msk | (255ull << shift);
            }
            __syncthreads();
        }
        const unsigned long long kth = s_prefix;         // the top_k-th largest key (keys are unique)
        for (int i = threadIdx.x; i < Kc; i += NT) {
            const uint64_t key = cand[i];
            if (key >= kth) {
                int slot = atomicAdd(&s_n, 1);
                if (slot < SORT_CAP) skeys[slot] = key;
            }
        }
        __syncthreads();
        n_sel = s_n < SORT_CAP ? s_n : SORT_CAP;
    }
    int n2 = 1;
    while (n2 < n_sel) n2 <<= 1;
    for (int i = n_sel + threadIdx.x; i < n2; i += NT) skeys[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(skeys, n2);
    const int n_top = n_sel < p.top_k ? n_sel : p.top_k;
    int n_out = 0;
    for (int i = threadIdx.x; i < n_top; i += NT) {
        const uint64_t key = skeys[i];
        const bool keep = !(p.min_score > 0.0f) || kb::key_score(key) > p.min_score;
        if (keep && i < p.cap) { emit_row(p, b, i, key); ++n_out; }   // dropped rows form a suffix
    }
    for (int d = 16; d > 0; d >>= 1) n_out += __shfl_xor_sync(0xffffffffu, n_out, d);
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && n_out) atomicAdd(&s_n, n_out);
    __syncthreads();
    if (threadIdx.x == 0) p.count[b] = s_n;
}

int launch_select(const SelectParams& p, cudaStream_t st) {
    const size_t smem = (size_t)SORT_CAP * sizeof(uint64_t);
    KB_CUDA_TRY(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    select_kernel<<<p.B, NT, smem, st>>>(p);
    KB_LAUNCH_CHECK();
    return KB_OK;
}

size_t nms_keep_bound(int H, int W, int r) {
    // kept pixels are pairwise more than r apart (Chebyshev) => at most ceil(H/(r+1))*ceil(W/(r+1))
    if (r <= 0) return (size_t)H * W;
    return (size_t)((H + r) / (r + 1)) * (size_t)((W + r) / (r + 1));
}

}  // namespace

extern "C" size_t kb_select_workspace_bytes(int B, int H, int W, int top_k) {
    (void)top_k;
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return kb_align_up((size_t)B * H * W * sizeof(uint64_t), 256) + 256;
}

extern "C" int kb_select(const float* nms_map, int B, int H, int W, int border_dist, float threshold,
                         float min_score, int top_k, int cap, float* xyp, int* raster, int* count, int* total,
                         void* ws, size_t ws_bytes, kb_stream_t stream) {
    if (!nms_map || !xyp || !raster || !count || B <= 0 || H <= 0 || W <= 0 || cap <= 0 || border_dist < 0)
        return KB_ERR_BAD_ARG;
    if (top_k > SORT_CAP) return KB_ERR_UNSUPPORTED;
    KbArena arena(ws, ws_bytes);
    SelectParams p;
    p.cand_cap = H * W;
    p.cand = arena.take<uint64_t>((size_t)B * p.cand_cap);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    p.map = nms_map; p.B = B; p.H = H; p.W = W; p.border = border_dist; p.top_k = top_k; p.cap = cap;
    p.threshold = threshold; p.min_score = min_score; p.xyp = xyp; p.raster = raster; p.count = count;
    p.total = total; p.skip = nullptr; p.want = 0; p.path = nullptr; p.path_code = 0;
    return launch_select(p, (cudaStream_t)stream);
}

// threshold < 0 lets suppressed (zeroed) pixels qualify, so the keep bound no longer applies
static size_t detect_cand_cap(int H, int W, int nms_dist, float threshold) {
    return (threshold >= 0.0f) ? nms_keep_bound(H, W, nms_dist) : (size_t)H * W;
}

static bool detect_uses_sparse(int H, int W, int nms_dist, float threshold, int top_k) {
    return nms_dist > 0 && threshold >= 0.0f && kb_sparse_workspace_bytes(1, H, W, nms_dist, top_k) > 0;
}

extern "C" size_t kb_detect_workspace_bytes(int B, int H, int W, int nms_dist, int top_k, float threshold) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    size_t n = 0;
    n += kb_align_up((size_t)B * H * W * sizeof(float), 256);                    // working map (fallback maps)
    n += kb_align_up(kb_fast_nms_workspace_bytes(B, H, W), 256);                 // NMS scratch
    n += kb_align_up((size_t)B * detect_cand_cap(H, W, nms_dist, threshold) * sizeof(uint64_t), 256);
    if (detect_uses_sparse(H, W, nms_dist, threshold, top_k))
        n += kb_align_up(kb_sparse_workspace_bytes(B, H, W, nms_dist, top_k), 256);
    return n + 1024;
}

extern "C" int kb_detect_phases(const float* score, int B, int H, int W, int nms_dist, int border_dist, float threshold,
                         float min_score, int top_k, float* xyp, int* raster, int* count, int* path, void* ws,
                         size_t ws_bytes, int phases, kb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!score || !xyp || !raster || !count || B <= 0 || H <= 0 || W <= 0 || nms_dist < 0 || border_dist < 0 ||
        top_k <= 0 || phases <= 0 || phases > 63 || ((phases >> 3) & ((phases >> 3) - 1)) != 0)
        return KB_ERR_BAD_ARG;
    if (top_k > SORT_CAP) return KB_ERR_UNSUPPORTED;
    if (B > 2048) return KB_ERR_UNSUPPORTED;         // callers split larger batches
    KbArena arena(ws, ws_bytes);
    float* work = arena.take<float>((size_t)B * H * W);
    const size_t nms_ws_bytes = kb_fast_nms_workspace_bytes(B, H, W);
    char* nms_ws = arena.take<char>(nms_ws_bytes);
    SelectParams p;
    p.cand_cap = (int)detect_cand_cap(H, W, nms_dist, threshold);
    p.cand = arena.take<uint64_t>((size_t)B * p.cand_cap);
    const bool sparse = detect_uses_sparse(H, W, nms_dist, threshold, top_k);
    const size_t sp_bytes = sparse ? kb_sparse_workspace_bytes(B, H, W, nms_dist, top_k) : 0;
    char* sp_ws = sparse ? arena.take<char>(sp_bytes) : nullptr;
    if (!arena.ok()) return KB_ERR_WORKSPACE;

    int* need_fallback = nullptr;
    int* any_fallback = nullptr;
    if (sparse) {
        // 1. sparse exact path for every map; maps it cannot certify are flagged on the device
        int rc = kb_sparse_detect(score, B, H, W, nms_dist, border_dist, threshold, min_score, top_k, xyp, raster,
                                  count, path, &need_fallback, &any_fallback, sp_ws, sp_bytes, phases, st);
        if (rc != KB_OK) return rc;
    }
    if (!(phases & 4)) return KB_OK;
    // 2. round-faithful NMS (each map on its own stopping rule) for the flagged maps -- all maps when
    //    the sparse path does not apply.  With nothing flagged both launches exit immediately.
    const float* sel_map = score;
    if (nms_dist > 0) {
        int rc = kb_nms_rounds_inplace(work, score, B, H, W, nms_dist, -1, 0.0f, /*per_map=*/1, need_fallback,
                                       any_fallback, nullptr, nms_ws, nms_ws_bytes, st);
        if (rc != KB_OK) return rc;
        sel_map = work;
    }
    p.map = sel_map; p.B = B; p.H = H; p.W = W; p.border = border_dist; p.top_k = top_k; p.cap = top_k;
    p.threshold = threshold; p.min_score = min_score; p.xyp = xyp; p.raster = raster; p.count = count;
    p.total = nullptr; p.skip = need_fallback; p.want = 1;
    p.path = path; p.path_code = 2;      // round-faithful path
    return launch_select(p, st);
}

extern "C" int kb_detect(const float* score, int B, int H, int W, int nms_dist, int border_dist, float threshold,
                         float min_score, int top_k, float* xyp, int* raster, int* count, int* path, void* ws,
                         size_t ws_bytes, kb_stream_t stream) {
    return kb_detect_phases(score, B, H, W, nms_dist, border_dist, threshold, min_score, top_k, xyp, raster, count,
                            path, ws, ws_bytes, 7, stream);
}
