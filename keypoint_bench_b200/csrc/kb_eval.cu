// Evaluation kernels: homography projection (utils/projection.py:137-167), the pairwise-distance /
// mutual-argmin / counting core of val_key_points (tasks/repeatability.py:39-51, 9-36, 69-85) and the
// MHA corner error (tasks/MHA.py:51-72).  Inputs are a few KB per pair, so these kernels are
// latency-bound; they exist to keep a whole batch of pairs on the device with no host round trip.
#include <math_constants.h>
#include "kb_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// warp_homography: one CTA per map, ordered compaction of valid / invalid points
// ------------------------------------------------------------------------------------------------
struct WarpParams {
    const float* pts;
    const int* count;
    const float* H33;
    const float* wh;
    float* kp_valid;
    float* kp_warp;
    int* ids;
    int* ids_out;
    int* n_valid;
    int stride, B, n_max;
};

__global__ void __launch_bounds__(1024) warp_homography_kernel(WarpParams p) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    const int n = p.count ? p.count[b] : p.n_max;
    const float* Hm = p.H33 + (size_t)b * 9;
    const float h00 = Hm[0], h01 = Hm[1], h02 = Hm[2], h10 = Hm[3], h11 = Hm[4], h12 = Hm[5], h20 = Hm[6],
                h21 = Hm[7], h22 = Hm[8];
    const float sx = p.wh[b * 2 + 0] - 1.0f, sy = p.wh[b * 2 + 1] - 1.0f;   // (w-1, h-1)
    int n_in = 0, n_out = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool live = i < n, valid = false;
        float px = 0, py = 0, u = 0, v = 0;
        if (live) {
            const float* pt = p.pts + ((size_t)b * p.n_max + i) * p.stride;
            px = __fmul_rn(pt[0], sx);                      // projection.py:147
            py = __fmul_rn(pt[1], sy);
            // einsum('ij,kj->ki'): q_i = H_i0*x + H_i1*y + H_i2*1    (projection.py:148-149)
            const float qx = __fadd_rn(__fadd_rn(__fmul_rn(h00, px), __fmul_rn(h01, py)), h02);
            const float qy = __fadd_rn(__fadd_rn(__fmul_rn(h10, px), __fmul_rn(h11, py)), h12);
            const float qz = __fadd_rn(__fadd_rn(__fmul_rn(h20, px), __fmul_rn(h21, py)), h22);
            u = qx / qz;                                    // projection.py:150
            v = qy / qz;
            valid = (u >= 0.0f) && (u <= sx) && (v >= 0.0f) && (v <= sy);   // projection.py:156
        }
        int tot_in, tot_out;
        const int off_in = n_in + kb::block_exclusive_scan((live && valid) ? 1 : 0, s_scan, &tot_in);
        const int off_out = n_out + kb::block_exclusive_scan((live && !valid) ? 1 : 0, s_scan, &tot_out);
        if (live && valid) {
            const size_t o = ((size_t)b * p.n_max + off_in) * 2;
            p.kp_valid[o + 0] = px / sx;                    // projection.py:165-166
            p.kp_valid[o + 1] = py / sy;
            p.kp_warp[o + 0] = u / sx;
            p.kp_warp[o + 1] = v / sy;
            p.ids[(size_t)b * p.n_max + off_in] = i;
        } else if (live) {
            p.ids_out[(size_t)b * p.n_max + off_out] = i;
        }
        n_in += tot_in;
        n_out += tot_out;
    }
    if (threadIdx.x == 0) p.n_valid[b] = n_in;
}

// ------------------------------------------------------------------------------------------------
// val_key_points core
// ------------------------------------------------------------------------------------------------
struct RepParams {
    const float* k0c;    // [B,a_max,2]
    const float* k01c;   // [B,a_max,2]
    const float* k1c;    // [B,b_max,2]
    const float* k10c;   // [B,b_max,2]
    const int* na;
    const int* nb;
    unsigned int* rowmin;   // [B,a_max] float bits (distances are >= 0: uint order == float order)
    unsigned int* colmin;   // [B,b_max]
    unsigned int* gmax;     // [B]
    double* stats;          // [B,4]
    float* errors;          // [B,a_max] or null
    int* pairs;             // [B,pair_cap,2] or null
    int B, a_max, b_max, pair_cap;
    float scale01, scale10, th;
};

__device__ __forceinline__ float dist_mutual(const RepParams& p, const float2 a, const float2 a1, const float2 bq,
                                             const float2 b0, int i, int j, int nd) {
    // dist01[i,j] = |k0c_i - k10c_j| ; dist10[j,i] = |k1c_j - k01c_i|   (repeatability.py:69-70)
    const float dx = __fsub_rn(a.x, b0.x), dy = __fsub_rn(a.y, b0.y);
    const float d01 = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float ex = __fsub_rn(bq.x, a1.x), ey = __fsub_rn(bq.y, a1.y);
    const float d10 = __fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
    float d = __fmul_rn(__fadd_rn(d01, d10), 0.5f);         // repeatability.py:71 (x/2 == x*0.5 exactly in IEEE fp32)
    if (i == j && i < nd) d = 99999.0f;                     // repeatability.py:72-73
    return d;
}

constexpr int RT = 256;     // threads; each block handles RB rows of dist_mutual
constexpr int RB = 8;

__global__ void __launch_bounds__(RT) rep_minima_kernel(RepParams p) {
    const int b = blockIdx.y;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    const int i0 = blockIdx.x * RB;
    if (i0 >= A || Bn <= 0) return;
    const int nd = A < Bn ? A : Bn;
    const float2* k0c = reinterpret_cast<const float2*>(p.k0c) + (size_t)b * p.a_max;
    const float2* k01c = reinterpret_cast<const float2*>(p.k01c) + (size_t)b * p.a_max;
    const float2* k1c = reinterpret_cast<const float2*>(p.k1c) + (size_t)b * p.b_max;
    const float2* k10c = reinterpret_cast<const float2*>(p.k10c) + (size_t)b * p.b_max;
    float2 a[RB], a1[RB];
    float rmin[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const int i = i0 + r < A ? i0 + r : A - 1;
        a[r] = k0c[i];
        a1[r] = k01c[i];
        rmin[r] = CUDART_INF_F;
    }
    float gmx = 0.0f;
    for (int j = threadIdx.x; j < Bn; j += RT) {
        const float2 bq = k1c[j], b0 = k10c[j];
        float cmin = CUDART_INF_F;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            if (i0 + r < A) {
                const float d = dist_mutual(p, a[r], a1[r], bq, b0, i0 + r, j, nd);
                rmin[r] = fminf(rmin[r], d);
                cmin = fminf(cmin, d);
                gmx = fmaxf(gmx, d);
            }
        }
        atomicMin(&p.colmin[(size_t)b * p.b_max + j], __float_as_uint(cmin));
    }
    __shared__ float s_red[RB][RT / 32];
    __shared__ float s_max[RT / 32];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        float v = rmin[r];
        for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
        if ((threadIdx.x & 31) == 0) s_red[r][threadIdx.x >> 5] = v;
    }
    for (int d = 16; d > 0; d >>= 1) gmx = fmaxf(gmx, __shfl_xor_sync(0xffffffffu, gmx, d));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = gmx;
    __syncthreads();
    if (threadIdx.x < RB && i0 + threadIdx.x < A) {
        float v = s_red[threadIdx.x][0];
        for (int w = 1; w < RT / 32; ++w) v = fminf(v, s_red[threadIdx.x][w]);
        p.rowmin[(size_t)b * p.a_max + i0 + threadIdx.x] = __float_as_uint(v);
        if (p.errors) p.errors[(size_t)b * p.a_max + i0 + threadIdx.x] = __fmul_rn(v, p.scale10);  // :78,85
    }
    if (threadIdx.x == 0) {
        float v = s_max[0];
        for (int w = 1; w < RT / 32; ++w) v = fmaxf(v, s_max[w]);
        atomicMax(&p.gmax[b], __float_as_uint(v));
    }
}

__global__ void __launch_bounds__(RT) rep_mutual_kernel(RepParams p) {
    const int b = blockIdx.y;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    const int i0 = blockIdx.x * RB;
    if (i0 >= A || Bn <= 0) return;
    const int nd = A < Bn ? A : Bn;
    const float2* k0c = reinterpret_cast<const float2*>(p.k0c) + (size_t)b * p.a_max;
    const float2* k01c = reinterpret_cast<const float2*>(p.k01c) + (size_t)b * p.a_max;
    const float2* k1c = reinterpret_cast<const float2*>(p.k1c) + (size_t)b * p.b_max;
    const float2* k10c = reinterpret_cast<const float2*>(p.k10c) + (size_t)b * p.b_max;
    // value = -dist_mutual; v = value - value.min() = (-d) - (-dmax)      (repeatability.py:18, 36)
    const float vmin = -__uint_as_float(p.gmax[b]);
    float2 a[RB], a1[RB];
    float rq[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const int i = i0 + r < A ? i0 + r : A - 1;
        a[r] = k0c[i];
        a1[r] = k01c[i];
        rq[r] = __fsub_rn(-__uint_as_float(p.rowmin[(size_t)b * p.a_max + i]), vmin);   // row max of v
    }
    int gt = 0, np = 0;
    double sum = 0.0;
    for (int j = threadIdx.x; j < Bn; j += RT) {
        const float2 bq = k1c[j], b0 = k10c[j];
        const float cq = __fsub_rn(-__uint_as_float(p.colmin[(size_t)b * p.b_max + j]), vmin);   // col max of v
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            if (i0 + r < A) {
                const float d = dist_mutual(p, a[r], a1[r], bq, b0, i0 + r, j, nd);
                const float v = __fsub_rn(-d, vmin);
                if (v == rq[r] && v == cq) {                // repeatability.py:25-28
                    const float ds = __fmul_rn(d, p.scale01);   // repeatability.py:76-80
                    ++np;
                    if (ds <= p.th) { ++gt; sum += (double)ds; }    // repeatability.py:82-83
                }
            }
        }
    }
    // block reduction then one atomic per block
    __shared__ int s_gt[RT / 32], s_np[RT / 32];
    __shared__ double s_sum[RT / 32];
    for (int d = 16; d > 0; d >>= 1) {
        gt += __shfl_xor_sync(0xffffffffu, gt, d);
        np += __shfl_xor_sync(0xffffffffu, np, d);
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
    }
    if ((threadIdx.x & 31) == 0) { s_gt[threadIdx.x >> 5] = gt; s_np[threadIdx.x >> 5] = np; s_sum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0, q = 0;
        double s = 0.0;
        for (int w = 0; w < RT / 32; ++w) { g += s_gt[w]; q += s_np[w]; s += s_sum[w]; }
        if (q) {
            atomicAdd(&p.stats[b * 4 + 0], (double)g);
            atomicAdd(&p.stats[b * 4 + 1], s);
            atomicAdd(&p.stats[b * 4 + 2], (double)q);
        }
    }
}

__global__ void rep_init_kernel(RepParams p) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t na = (size_t)p.B * p.a_max, nb = (size_t)p.B * p.b_max;
    if (i < na) p.rowmin[i] = 0x7f800000u;
    if (i < nb) p.colmin[i] = 0x7f800000u;
    if (i < (size_t)p.B) p.gmax[i] = 0u;
    if (i < (size_t)p.B * 4) p.stats[i] = 0.0;
    if (p.errors && i < na) p.errors[i] = 0.0f;
}

// second variant of the mutual sweep that also lists the pairs (unordered)
__global__ void __launch_bounds__(RT) rep_pairs_kernel(RepParams p, int* pair_count) {
    const int b = blockIdx.y;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    const int i0 = blockIdx.x * RB;
    if (i0 >= A || Bn <= 0) return;
    const int nd = A < Bn ? A : Bn;
    const float2* k0c = reinterpret_cast<const float2*>(p.k0c) + (size_t)b * p.a_max;
    const float2* k01c = reinterpret_cast<const float2*>(p.k01c) + (size_t)b * p.a_max;
    const float2* k1c = reinterpret_cast<const float2*>(p.k1c) + (size_t)b * p.b_max;
    const float2* k10c = reinterpret_cast<const float2*>(p.k10c) + (size_t)b * p.b_max;
    const float vmin = -__uint_as_float(p.gmax[b]);
    for (int r = 0; r < RB; ++r) {
        const int i = i0 + r;
        if (i >= A) break;
        const float2 a = k0c[i], a1 = k01c[i];
        const float rq = __fsub_rn(-__uint_as_float(p.rowmin[(size_t)b * p.a_max + i]), vmin);
        for (int j = threadIdx.x; j < Bn; j += RT) {
            const float d = dist_mutual(p, a, a1, k1c[j], k10c[j], i, j, nd);
            const float v = __fsub_rn(-d, vmin);
            const float cq = __fsub_rn(-__uint_as_float(p.colmin[(size_t)b * p.b_max + j]), vmin);
            if (v == rq && v == cq) {
                const int slot = atomicAdd(&pair_count[b], 1);
                if (slot < p.pair_cap) {
                    p.pairs[((size_t)b * p.pair_cap + slot) * 2 + 0] = i;
                    p.pairs[((size_t)b * p.pair_cap + slot) * 2 + 1] = j;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// MHA corner error (float64, 4 points per pair)
// ------------------------------------------------------------------------------------------------
__global__ void corner_error_kernel(const double* h_est, const double* h_real, const int* valid, int B, int w, int h,
                                    int resize_h, int resize_w, const double* th, int n_th, double* mean_dist,
                                    double* flags) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const bool ok = valid ? valid[b] != 0 : true;
    // corners as the reference writes them (x/y swapped on purpose, MHA.py:52-55)
    const double cx[4] = {0.0, (double)(h - 1), 0.0, (double)(h - 1)};
    const double cy[4] = {0.0, 0.0, (double)(w - 1), (double)(w - 1)};
    const double sx = (double)resize_h / (double)h, sy = (double)resize_w / (double)w;   // MHA.py:63-64
    const double* He = h_est + (size_t)b * 9;
    const double* Hr = h_real + (size_t)b * 9;
    double acc = 0.0;
    for (int c = 0; c < 4; ++c) {
        const double rz = Hr[6] * cx[c] + Hr[7] * cy[c] + Hr[8];
        const double rx = (Hr[0] * cx[c] + Hr[1] * cy[c] + Hr[2]) / rz;
        const double ry = (Hr[3] * cx[c] + Hr[4] * cy[c] + Hr[5]) / rz;
        const double ez = He[6] * cx[c] + He[7] * cy[c] + He[8];
        const double ex = (He[0] * cx[c] + He[1] * cy[c] + He[2]) / ez;
        const double ey = (He[3] * cx[c] + He[4] * cy[c] + He[5]) / ez;
        const double dx = rx * sx - ex * sx, dy = ry * sy - ey * sy;
        acc += sqrt(dx * dx + dy * dy);
    }
    const double md = acc / 4.0;                                    // MHA.py:66
    mean_dist[b] = ok ? md : CUDART_NAN;
    for (int t = 0; t < n_th; ++t) flags[(size_t)b * n_th + t] = (ok && md <= th[t]) ? 1.0 : 0.0;   // MHA.py:68-72
}

}  // namespace

extern "C" int kb_warp_homography(const float* pts, int pts_stride, const int* count, int B, int n_max,
                                  const float* H33, const float* wh, float* kp_valid, float* kp_warp, int* ids,
                                  int* ids_out, int* n_valid, kb_stream_t stream) {
    if (!pts || !H33 || !wh || !kp_valid || !kp_warp || !ids || !ids_out || !n_valid || B <= 0 || n_max <= 0 ||
        pts_stride < 2)
        return KB_ERR_BAD_ARG;
    WarpParams p;
    p.pts = pts; p.count = count; p.H33 = H33; p.wh = wh; p.kp_valid = kp_valid; p.kp_warp = kp_warp;
    p.ids = ids; p.ids_out = ids_out; p.n_valid = n_valid; p.stride = pts_stride; p.B = B; p.n_max = n_max;
    warp_homography_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(p);
    KB_LAUNCH_CHECK();
    return KB_OK;
}

extern "C" size_t kb_repeat_workspace_bytes(int B, int a_max, int b_max) {
    if (B <= 0 || a_max <= 0 || b_max <= 0) return 0;
    return kb_align_up((size_t)B * a_max * 4, 256) + kb_align_up((size_t)B * b_max * 4, 256) +
           kb_align_up((size_t)B * 4, 256) * 2 + 1024;
}

extern "C" int kb_repeat_counts(const float* k0c, const float* k01c, const int* na, const float* k1c,
                                const float* k10c, const int* nb, int B, int a_max, int b_max, float scale01,
                                float scale10, float th, double* stats, float* errors, int* pairs, int pair_cap,
                                void* ws, size_t ws_bytes, kb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!k0c || !k01c || !k1c || !k10c || !stats || B <= 0 || a_max <= 0 || b_max <= 0) return KB_ERR_BAD_ARG;
    if (B > 65535) return KB_ERR_UNSUPPORTED;
    if (pairs && pair_cap <= 0) return KB_ERR_BAD_ARG;
    KbArena arena(ws, ws_bytes);
    RepParams p;
    p.rowmin = arena.take<unsigned int>((size_t)B * a_max);
    p.colmin = arena.take<unsigned int>((size_t)B * b_max);
    p.gmax = arena.take<unsigned int>(B);
    int* pair_count = arena.take<int>(B);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    p.k0c = k0c; p.k01c = k01c; p.k1c = k1c; p.k10c = k10c; p.na = na; p.nb = nb; p.stats = stats;
    p.errors = errors; p.pairs = pairs; p.B = B; p.a_max = a_max; p.b_max = b_max; p.pair_cap = pair_cap;
    p.scale01 = scale01; p.scale10 = scale10; p.th = th;
    const size_t n_init = (size_t)B * (a_max > b_max ? a_max : b_max);
    const size_t n_init2 = n_init > (size_t)B * 4 ? n_init : (size_t)B * 4;
    rep_init_kernel<<<(unsigned)((n_init2 + 255) / 256), 256, 0, st>>>(p);
    KB_LAUNCH_CHECK();
    dim3 grid((a_max + RB - 1) / RB, B);
    rep_minima_kernel<<<grid, RT, 0, st>>>(p);
    KB_LAUNCH_CHECK();
    RepParams q = p;
    q.pairs = nullptr;
    rep_mutual_kernel<<<grid, RT, 0, st>>>(q);
    KB_LAUNCH_CHECK();
    if (pairs) {
        KB_CUDA_TRY(cudaMemsetAsync(pair_count, 0, (size_t)B * sizeof(int), st));
        rep_pairs_kernel<<<grid, RT, 0, st>>>(p, pair_count);
        KB_LAUNCH_CHECK();
    }
    return KB_OK;
}

extern "C" int kb_corner_error(const double* h_est, const double* h_real, const int* valid, int B, int w, int h,
                               int resize_h, int resize_w, const double* th, int n_th, double* mean_dist,
                               double* flags, kb_stream_t stream) {
    if (!h_est || !h_real || !th || !mean_dist || !flags || B <= 0 || n_th <= 0 || w <= 1 || h <= 1)
        return KB_ERR_BAD_ARG;
    corner_error_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h_est, h_real, valid, B, w, h, resize_h,
                                                                            resize_w, th, n_th, mean_dist, flags);
    KB_LAUNCH_CHECK();
    return KB_OK;
}
