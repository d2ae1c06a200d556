// Evaluation kernels: homography projection (utils/projection.py:137-167), the pairwise-distance /
// mutual-argmin / counting core of val_key_points (tasks/repeatability.py:39-51, 9-36, 69-85) and the
// MHA corner error (tasks/MHA.py:51-72).  Inputs are a few KB per pair, so these kernels are
// latency-bound; they exist to keep a whole batch of pairs on the device with no host round trip.
#include <math_constants.h>
#include <cstdlib>
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
    const int* only_flagged; // [B] or null: the exhaustive kernels skip maps with only_flagged[b] == 0
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
    if (p.only_flagged && !p.only_flagged[b]) return;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    if (Bn <= 0) return;
    const int nd = A < Bn ? A : Bn;
    const float2* k0c = reinterpret_cast<const float2*>(p.k0c) + (size_t)b * p.a_max;
    const float2* k01c = reinterpret_cast<const float2*>(p.k01c) + (size_t)b * p.a_max;
    const float2* k1c = reinterpret_cast<const float2*>(p.k1c) + (size_t)b * p.b_max;
    const float2* k10c = reinterpret_cast<const float2*>(p.k10c) + (size_t)b * p.b_max;
    __shared__ float s_red[RB][RT / 32];
    __shared__ float s_max[RT / 32];
    // the grid is small (these kernels only run for maps the pruned path rejected): every block walks row blocks
    for (int i0 = blockIdx.x * RB; i0 < A; i0 += gridDim.x * RB) {
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
    __syncthreads();
    }
}

__global__ void __launch_bounds__(RT) rep_mutual_kernel(RepParams p) {
    const int b = blockIdx.y;
    if (p.only_flagged && !p.only_flagged[b]) return;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    if (Bn <= 0) return;
    const int nd = A < Bn ? A : Bn;
    const float2* k0c = reinterpret_cast<const float2*>(p.k0c) + (size_t)b * p.a_max;
    const float2* k01c = reinterpret_cast<const float2*>(p.k01c) + (size_t)b * p.a_max;
    const float2* k1c = reinterpret_cast<const float2*>(p.k1c) + (size_t)b * p.b_max;
    const float2* k10c = reinterpret_cast<const float2*>(p.k10c) + (size_t)b * p.b_max;
    // value = -dist_mutual; v = value - value.min() = (-d) - (-dmax)      (repeatability.py:18, 36)
    const float vmin = -__uint_as_float(p.gmax[b]);
    int gt = 0, np = 0;
    double sum = 0.0;
    for (int i0 = blockIdx.x * RB; i0 < A; i0 += gridDim.x * RB) {       // small grid: every block walks row blocks
    float2 a[RB], a1[RB];
    float rq[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const int i = i0 + r < A ? i0 + r : A - 1;
        a[r] = k0c[i];
        a1[r] = k01c[i];
        rq[r] = __fsub_rn(-__uint_as_float(p.rowmin[(size_t)b * p.a_max + i]), vmin);   // row max of v
    }
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


// ------------------------------------------------------------------------------------------------
// Pruned variant of the two sweeps above (the default): eight lanes own one row (or column) of dist_mutual
// and walk over the other side's points staged in shared memory.  Since
//     dist_mutual = (|a - b0| + |bq - a1|) / 2  >=  max(|a - b0|, |bq - a1|) / 2,
// an entry whose squared lower bound exceeds the (squared) running minimum cannot lower it and is skipped
// before the two IEEE square roots -- almost every entry is.  The values that ARE evaluated use exactly the
// arithmetic of dist_mutual(), so minima, counts and sums are identical to the exhaustive sweeps.
// Precondition (rep_bound_kernel): every distance is < 99999, so min(-dist_mutual) = -99999 comes from the
// masked diagonal (repeatability.py:72-73); maps that violate it take the exhaustive kernels.
// ------------------------------------------------------------------------------------------------
constexpr int PT = 128;      // threads per block
constexpr int PL = 8;        // lanes that share one row (column): they split the other side's points
constexpr int PR = PT / PL;  // rows (columns) per block
constexpr int PTILE = 512;   // points of the other side per shared-memory tile

__global__ void __launch_bounds__(256) rep_bound_kernel(RepParams p, int* need_bf) {
    __shared__ float s_m[8];
    const int b = blockIdx.x;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    float m = 0.0f;
    bool bad = false;
    for (int i = threadIdx.x; i < 2 * A; i += 256) {
        const float v0 = fabsf(p.k0c[(size_t)b * p.a_max * 2 + i]), v1 = fabsf(p.k01c[(size_t)b * p.a_max * 2 + i]);
        bad |= !(v0 <= 1e30f) || !(v1 <= 1e30f);
        m = fmaxf(m, fmaxf(v0, v1));
    }
    for (int i = threadIdx.x; i < 2 * Bn; i += 256) {
        const float v0 = fabsf(p.k1c[(size_t)b * p.b_max * 2 + i]), v1 = fabsf(p.k10c[(size_t)b * p.b_max * 2 + i]);
        bad |= !(v0 <= 1e30f) || !(v1 <= 1e30f);
        m = fmaxf(m, fmaxf(v0, v1));
    }
    if (bad) m = CUDART_INF_F;
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_m[w]);
        // |p - q| <= 2*sqrt(2)*m per term, so every distance is <= 3*m
        const bool ok = A > 0 && Bn > 0 && (3.0f * m < 99999.0f);
        need_bf[b] = ok ? 0 : 1;
        if (ok) p.gmax[b] = __float_as_uint(99999.0f);
    }
}

// SIDE 0: one thread per row i (minimum over the columns); SIDE 1: one thread per column j.
template <int SIDE>
__global__ void __launch_bounds__(PT) rep_min_pruned_kernel(RepParams p, const int* need_bf) {
    __shared__ float4 s_pts[PTILE];
    const int b = blockIdx.y;
    if (need_bf[b]) return;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    const int n_own = SIDE ? Bn : A, n_oth = SIDE ? A : Bn;
    if (blockIdx.x * PR >= n_own) return;
    const int nd = A < Bn ? A : Bn;
    const int own = blockIdx.x * PR + threadIdx.x / PL, sub = threadIdx.x % PL;
    const bool live = own < n_own;
    // the two point sets of this map: "first" pairs with image-0 coordinates, "second" with image-1 coordinates
    const float2* own1 = reinterpret_cast<const float2*>(SIDE ? p.k10c : p.k0c) + (size_t)b * (SIDE ? p.b_max : p.a_max);
    const float2* own2 = reinterpret_cast<const float2*>(SIDE ? p.k1c : p.k01c) + (size_t)b * (SIDE ? p.b_max : p.a_max);
    const float2* oth1 = reinterpret_cast<const float2*>(SIDE ? p.k0c : p.k10c) + (size_t)b * (SIDE ? p.a_max : p.b_max);
    const float2* oth2 = reinterpret_cast<const float2*>(SIDE ? p.k01c : p.k1c) + (size_t)b * (SIDE ? p.a_max : p.b_max);
    const float2 P = live ? own1[own] : make_float2(0.f, 0.f), Q = live ? own2[own] : make_float2(0.f, 0.f);
    float best = CUDART_INF_F, thr = CUDART_INF_F;
    for (int t0 = 0; t0 < n_oth; t0 += PTILE) {
        __syncthreads();
        for (int t = threadIdx.x; t < PTILE && t0 + t < n_oth; t += PT) {
            const float2 o1 = oth1[t0 + t], o2 = oth2[t0 + t];
            s_pts[t] = make_float4(o1.x, o1.y, o2.x, o2.y);
        }
        __syncthreads();
        const int nt = min(PTILE, n_oth - t0);
        for (int t = sub; live && t < nt; t += PL) {
            const float4 o = s_pts[t];
            // (x - y)^2 == (y - x)^2 exactly, so the operand order of dist_mutual() does not matter here
            const float dx = __fsub_rn(P.x, o.x), dy = __fsub_rn(P.y, o.y);
            const float s1 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            if (s1 * 0.25f > thr) continue;
            const float ex = __fsub_rn(o.z, Q.x), ey = __fsub_rn(o.w, Q.y);
            const float s2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            if (s2 * 0.25f > thr) continue;
            float d = __fmul_rn(__fadd_rn(__fsqrt_rn(s1), __fsqrt_rn(s2)), 0.5f);         // repeatability.py:69-71
            if (t0 + t == own && own < nd) d = 99999.0f;                                  // repeatability.py:72-73
            if (d < best) { best = d; thr = d * d * 1.0001f; }
        }
        // the PL lanes of a row share their running minimum once per tile (tighter pruning)
#pragma unroll
        for (int m = PL / 2; m > 0; m >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, m));
        thr = best * best * 1.0001f;
    }
    if (live && sub == 0) {
        if (SIDE == 0) {
            p.rowmin[(size_t)b * p.a_max + own] = __float_as_uint(best);
            if (p.errors) p.errors[(size_t)b * p.a_max + own] = __fmul_rn(best, p.scale10);      // :78,85
        } else {
            p.colmin[(size_t)b * p.b_max + own] = __float_as_uint(best);
        }
    }
}

__global__ void __launch_bounds__(PT) rep_mutual_pruned_kernel(RepParams p, const int* need_bf) {
    __shared__ float4 s_pts[PTILE];
    __shared__ unsigned int s_cmin[PTILE];
    __shared__ int s_gt[PT / 32], s_np[PT / 32];
    __shared__ double s_sum[PT / 32];
    const int b = blockIdx.y;
    if (need_bf[b]) return;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    if (blockIdx.x * PR >= A) return;
    const int nd = A < Bn ? A : Bn;
    const int i = blockIdx.x * PR + threadIdx.x / PL, sub = threadIdx.x % PL;
    const bool live = i < A;
    const float2* k0c = reinterpret_cast<const float2*>(p.k0c) + (size_t)b * p.a_max;
    const float2* k01c = reinterpret_cast<const float2*>(p.k01c) + (size_t)b * p.a_max;
    const float2* k1c = reinterpret_cast<const float2*>(p.k1c) + (size_t)b * p.b_max;
    const float2* k10c = reinterpret_cast<const float2*>(p.k10c) + (size_t)b * p.b_max;
    const float vmin = -__uint_as_float(p.gmax[b]);          // = -99999: value.min() of repeatability.py:18
    const float2 a = live ? k0c[i] : make_float2(0.f, 0.f), a1 = live ? k01c[i] : make_float2(0.f, 0.f);
    const float rmin = live ? __uint_as_float(p.rowmin[(size_t)b * p.a_max + i]) : 0.0f;
    const float rq = __fsub_rn(-rmin, vmin);                 // row maximum of v = (-d) - min(-d)
    // v has a resolution of one ulp of 99999 (2^-7): every d that rounds onto rq lies within 2 ulps of the
    // row minimum; everything farther is skipped before the square roots
    const float lim = rmin + 0.0172f;
    const float thr = lim * lim * 1.0001f;
    int gt = 0, np = 0;
    double sum = 0.0;
    for (int t0 = 0; t0 < Bn; t0 += PTILE) {
        __syncthreads();
        for (int t = threadIdx.x; t < PTILE && t0 + t < Bn; t += PT) {
            const float2 o1 = k10c[t0 + t], o2 = k1c[t0 + t];
            s_pts[t] = make_float4(o1.x, o1.y, o2.x, o2.y);
            s_cmin[t] = p.colmin[(size_t)b * p.b_max + t0 + t];
        }
        __syncthreads();
        const int nt = min(PTILE, Bn - t0);
        for (int t = sub; live && t < nt; t += PL) {
            const float4 o = s_pts[t];
            const float dx = __fsub_rn(a.x, o.x), dy = __fsub_rn(a.y, o.y);
            const float s1 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const int j = t0 + t;
            const bool diag = (j == i && i < nd);
            if (!diag && s1 * 0.25f > thr) continue;
            const float ex = __fsub_rn(o.z, a1.x), ey = __fsub_rn(o.w, a1.y);
            const float s2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            if (!diag && s2 * 0.25f > thr) continue;
            float d = __fmul_rn(__fadd_rn(__fsqrt_rn(s1), __fsqrt_rn(s2)), 0.5f);
            if (diag) d = 99999.0f;
            const float v = __fsub_rn(-d, vmin);
            if (v == rq) {
                const float cq = __fsub_rn(-__uint_as_float(s_cmin[t]), vmin);           // column maximum of v
                if (v == cq) {                                                           // repeatability.py:25-28
                    const float ds = __fmul_rn(d, p.scale01);                            // :76-80
                    ++np;
                    if (ds <= p.th) { ++gt; sum += (double)ds; }                         // :82-83
                }
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        gt += __shfl_xor_sync(0xffffffffu, gt, d);
        np += __shfl_xor_sync(0xffffffffu, np, d);
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
    }
    if ((threadIdx.x & 31) == 0) { s_gt[threadIdx.x >> 5] = gt; s_np[threadIdx.x >> 5] = np; s_sum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0, q = 0;
        double sm = 0.0;
        for (int w = 0; w < PT / 32; ++w) { g += s_gt[w]; q += s_np[w]; sm += s_sum[w]; }
        if (q) {
            atomicAdd(&p.stats[b * 4 + 0], (double)g);
            atomicAdd(&p.stats[b * 4 + 1], sm);
            atomicAdd(&p.stats[b * 4 + 2], (double)q);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Sorted sweeps (the default when both sides have <= SRT_MAX points): the other side's points are sorted by the
// y coordinate of their FIRST pair of coordinates once per map, and a row (column) only visits the entries around
// its own y, outwards in both directions, until |dy| alone excludes them:
//     dist_mutual >= |a - b0| / 2 >= |dy| / 2,
// so with the other side sorted by y the walk stops as soon as dy^2 / 4 exceeds the squared running minimum --
// everything farther is farther in y.  A repeatable point finds its partner within the first few entries and stops
// after two or three steps; the old sweep above evaluates the bound for every entry.  Evaluated entries use exactly
// the arithmetic of dist_mutual(); skipped entries are provably larger than the minimum (same 1.0001 guard as
// above), so minima, counts and sums are identical.  The masked diagonal entry (d = 99999 whatever its geometry)
// is taken out of the walk and added back explicitly.
// ------------------------------------------------------------------------------------------------
constexpr int SRT_MAX = 4096;     // points per side the sorted path handles (shared-memory staging)
constexpr int SW_NT = 256;        // threads per block of the sweeps: 32 rows x 8 lanes
constexpr int SW_ROWS = SW_NT / PL;

struct SortedSide {
    float4* pts;      // [B, n_max] (o1.x, o1.y, o2.x, o2.y) sorted by o1.y ascending
    int* idx;         // [B, n_max] original index of every sorted entry
};

// side 0: the B points (k10c, k1c) sorted by k10c.y; side 1: the A points (k0c, k01c) sorted by k0c.y
__global__ void __launch_bounds__(1024) rep_sort_kernel(RepParams p, const int* need_bf, SortedSide s0, SortedSide s1) {
    extern __shared__ unsigned long long srt_keys[];
    const int b = blockIdx.x, side = blockIdx.y;
    if (need_bf[b]) return;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    const int n = side ? A : Bn, n_max = side ? p.a_max : p.b_max;
    const float2* o1 = reinterpret_cast<const float2*>(side ? p.k0c : p.k10c) + (size_t)b * n_max;
    const float2* o2 = reinterpret_cast<const float2*>(side ? p.k01c : p.k1c) + (size_t)b * n_max;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += 1024)
        srt_keys[i] = i < n ? (((unsigned long long)kb::float_order_key(o1[i].y) << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += 1024) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long x = srt_keys[i], y = srt_keys[l];
                    const bool asc = ((i & k) == 0);
                    if (asc ? (x > y) : (x < y)) { srt_keys[i] = y; srt_keys[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    float4* op = (side ? s1.pts : s0.pts) + (size_t)b * n_max;
    int* oi = (side ? s1.idx : s0.idx) + (size_t)b * n_max;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const int src = (int)(srt_keys[i] & 0xffffffffu);
        const float2 a = o1[src], c = o2[src];
        op[i] = make_float4(a.x, a.y, c.x, c.y);
        oi[i] = src;
    }
}

// the walk shared by the three sorted kernels: lanes 0-3 of a row go up from its y, lanes 4-7 go down
struct Walk {
    int k, step, n;
    __device__ __forceinline__ void init(int pos, int sub, int n_) {
        n = n_;
        if (sub < PL / 2) { k = pos + sub; step = PL / 2; } else { k = pos - 1 - (sub - PL / 2); step = -(PL / 2); }
    }
    __device__ __forceinline__ bool in_range() const { return k >= 0 && k < n; }
    __device__ __forceinline__ void stop() { k = -1; step = -1; }
};

__device__ __forceinline__ int lower_bound_y(const float4* pts, int n, float y) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pts[mid].y < y) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int SIDE>
__global__ void __launch_bounds__(SW_NT) rep_min_sorted_kernel(RepParams p, const int* need_bf, SortedSide so) {
    extern __shared__ __align__(16) unsigned char sw_smem[];
    const int b = blockIdx.y;
    if (need_bf[b]) return;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    const int n_own = SIDE ? Bn : A, n_oth = SIDE ? A : Bn, oth_max = SIDE ? p.a_max : p.b_max;
    if (blockIdx.x * SW_ROWS >= n_own) return;
    const int nd = A < Bn ? A : Bn;
    float4* s_pts = reinterpret_cast<float4*>(sw_smem);
    int* s_idx = reinterpret_cast<int*>(s_pts + n_oth);
    for (int t = threadIdx.x; t < n_oth; t += SW_NT) {
        s_pts[t] = so.pts[(size_t)b * oth_max + t];
        s_idx[t] = so.idx[(size_t)b * oth_max + t];
    }
    __syncthreads();
    const int own = blockIdx.x * SW_ROWS + threadIdx.x / PL, sub = threadIdx.x % PL;
    const bool live = own < n_own;
    const float2* own1 = reinterpret_cast<const float2*>(SIDE ? p.k10c : p.k0c) + (size_t)b * (SIDE ? p.b_max : p.a_max);
    const float2* own2 = reinterpret_cast<const float2*>(SIDE ? p.k1c : p.k01c) + (size_t)b * (SIDE ? p.b_max : p.a_max);
    const float2 P = live ? own1[own] : make_float2(0.f, 0.f), Q = live ? own2[own] : make_float2(0.f, 0.f);
    // the masked diagonal entry counts as 99999 wherever it lies; it is skipped in the walk
    const bool has_diag = live && own < nd;
    float best = has_diag ? 99999.0f : CUDART_INF_F;
    float thr = has_diag ? 99999.0f * 99999.0f * 1.0001f : CUDART_INF_F;
    Walk w;
    w.init(live ? lower_bound_y(s_pts, n_oth, P.y) : 0, sub, live ? n_oth : 0);
    while (true) {
        bool active = w.in_range();
        if (active) {
            const float4 o = s_pts[w.k];
            const float dy = __fsub_rn(P.y, o.y);
            if (__fmul_rn(dy, dy) * 0.25f > thr) {
                w.stop();                                   // everything farther in this direction is farther in y
                active = false;
            } else {
                const float dx = __fsub_rn(P.x, o.x);
                const float s1 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                if (!(s1 * 0.25f > thr) && !(has_diag && s_idx[w.k] == own)) {
                    const float ex = __fsub_rn(o.z, Q.x), ey = __fsub_rn(o.w, Q.y);
                    const float s2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    if (!(s2 * 0.25f > thr)) {
                        const float d = __fmul_rn(__fadd_rn(__fsqrt_rn(s1), __fsqrt_rn(s2)), 0.5f);     // repeatability.py:69-71
                        best = fminf(best, d);
                    }
                }
                w.k += w.step;
            }
        }
        // the PL lanes of a row share their running minimum after every step
#pragma unroll
        for (int m = PL / 2; m > 0; m >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, m));
        thr = best * best * 1.0001f;
        if (!__any_sync(0xffffffffu, active)) break;
    }
    if (live && sub == 0) {
        if (SIDE == 0) {
            p.rowmin[(size_t)b * p.a_max + own] = __float_as_uint(best);
            if (p.errors) p.errors[(size_t)b * p.a_max + own] = __fmul_rn(best, p.scale10);      // :78,85
        } else {
            p.colmin[(size_t)b * p.b_max + own] = __float_as_uint(best);
        }
    }
}

__global__ void __launch_bounds__(SW_NT) rep_mutual_sorted_kernel(RepParams p, const int* need_bf, SortedSide so) {
    extern __shared__ __align__(16) unsigned char sw_smem[];
    __shared__ int s_gt[SW_NT / 32], s_np[SW_NT / 32];
    __shared__ double s_sum[SW_NT / 32];
    const int b = blockIdx.y;
    if (need_bf[b]) return;
    const int A = p.na ? p.na[b] : p.a_max, Bn = p.nb ? p.nb[b] : p.b_max;
    if (blockIdx.x * SW_ROWS >= A) return;
    const int nd = A < Bn ? A : Bn;
    float4* s_pts = reinterpret_cast<float4*>(sw_smem);
    int* s_idx = reinterpret_cast<int*>(s_pts + Bn);
    for (int t = threadIdx.x; t < Bn; t += SW_NT) {
        s_pts[t] = so.pts[(size_t)b * p.b_max + t];
        s_idx[t] = so.idx[(size_t)b * p.b_max + t];
    }
    __syncthreads();
    const int i = blockIdx.x * SW_ROWS + threadIdx.x / PL, sub = threadIdx.x % PL;
    const bool live = i < A;
    const float2* k0c = reinterpret_cast<const float2*>(p.k0c) + (size_t)b * p.a_max;
    const float2* k01c = reinterpret_cast<const float2*>(p.k01c) + (size_t)b * p.a_max;
    const unsigned int* colmin = p.colmin + (size_t)b * p.b_max;
    const float vmin = -__uint_as_float(p.gmax[b]);          // = -99999: value.min() of repeatability.py:18
    const float2 a = live ? k0c[i] : make_float2(0.f, 0.f), a1 = live ? k01c[i] : make_float2(0.f, 0.f);
    const float rmin = live ? __uint_as_float(p.rowmin[(size_t)b * p.a_max + i]) : 0.0f;
    const float rq = __fsub_rn(-rmin, vmin);                 // row maximum of v = (-d) - min(-d)
    // v has a resolution of one ulp of 99999 (2^-7): every d that rounds onto rq lies within 2 ulps of the
    // row minimum; everything farther is skipped
    const float lim = rmin + 0.0172f;
    const float thr = lim * lim * 1.0001f;
    const bool has_diag = live && i < nd;
    int gt = 0, np = 0;
    double sum = 0.0;
    auto tally = [&](float d, int j) {
        const float v = __fsub_rn(-d, vmin);
        if (v == rq) {
            const float cq = __fsub_rn(-__uint_as_float(colmin[j]), vmin);               // column maximum of v
            if (v == cq) {                                                               // repeatability.py:25-28
                const float ds = __fmul_rn(d, p.scale01);                                // :76-80
                ++np;
                if (ds <= p.th) { ++gt; sum += (double)ds; }                             // :82-83
            }
        }
    };
    if (has_diag && sub == 0) tally(99999.0f, i);            // the masked diagonal entry (repeatability.py:72-73)
    Walk w;
    w.init(live ? lower_bound_y(s_pts, Bn, a.y) : 0, sub, live ? Bn : 0);
    while (w.in_range()) {
        const float4 o = s_pts[w.k];
        const float dy = __fsub_rn(a.y, o.y);
        if (__fmul_rn(dy, dy) * 0.25f > thr) break;
        const float dx = __fsub_rn(a.x, o.x);
        const float s1 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const int j = s_idx[w.k];
        if (!(s1 * 0.25f > thr) && !(has_diag && j == i)) {
            const float ex = __fsub_rn(o.z, a1.x), ey = __fsub_rn(o.w, a1.y);
            const float s2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            if (!(s2 * 0.25f > thr)) tally(__fmul_rn(__fadd_rn(__fsqrt_rn(s1), __fsqrt_rn(s2)), 0.5f), j);
        }
        w.k += w.step;
    }
    for (int d = 16; d > 0; d >>= 1) {
        gt += __shfl_xor_sync(0xffffffffu, gt, d);
        np += __shfl_xor_sync(0xffffffffu, np, d);
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
    }
    if ((threadIdx.x & 31) == 0) { s_gt[threadIdx.x >> 5] = gt; s_np[threadIdx.x >> 5] = np; s_sum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0, q = 0;
        double sm = 0.0;
        for (int wq = 0; wq < SW_NT / 32; ++wq) { g += s_gt[wq]; q += s_np[wq]; sm += s_sum[wq]; }
        if (q) {
            atomicAdd(&p.stats[b * 4 + 0], (double)g);
            atomicAdd(&p.stats[b * 4 + 1], sm);
            atomicAdd(&p.stats[b * 4 + 2], (double)q);
        }
    }
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

// ------------------------------------------------------------------------------------------------
// warp_se3 (utils/projection.py:194-267 with interpolate_depth :270-372): depth-based covisibility.
// One CTA per map; every keypoint is classified (no depth in view 0 / projected outside view 1's
// valid-corner area / no depth in view 1 / occluded / valid) and the lists are compacted in input order.
// ------------------------------------------------------------------------------------------------
struct Se3Params {
    const float* pts;
    const int* count;
    const float* depth0;     // [B,h0,w0]
    const float* depth1;     // [B,h1,w1]
    const float* kinv0;      // [B,9]  inverse intrinsics of view 0
    const float* k1;         // [B,9]
    const float* pose;       // [B,16] row-major 4x4 (pose01)
    const float* bbox0;      // [B,2] (row, col)
    const float* bbox1;
    float* kp_valid;
    float* kp_warp;
    int* ids;
    int* ids_out;
    int* n_valid;
    int* n_out;
    int stride, B, n_max, h0, w0, h1, w1;
};

// interpolate_depth for one point (x, y) in pixels: corners floor/ceil inside a 10-pixel border
// (projection.py:289-301), all four corner depths > 0 (:321-324), bilinear weights from the floor corner
// (:346-357).  Returns 0 = corners invalid, 1 = corners valid but a corner depth is missing, 2 = ok.
__device__ __forceinline__ int interp_depth(const float* depth, int h, int w, float x, float y, float* z) {
    const float border = 10.0f;
    const float i0 = floorf(y), j0 = floorf(x), i1 = ceilf(y), j1 = ceilf(x);
    // written so that NaN coordinates fail (torch's NaN -> long conversion never yields a valid corner)
    const bool corners = (i0 >= border) && (j0 >= border) && (j1 < (float)w - border) && (i1 < (float)h - border);
    if (!corners) return 0;
    const int ii0 = (int)i0, jj0 = (int)j0, ii1 = (int)i1, jj1 = (int)j1;
    const float d_tl = __ldg(depth + (size_t)ii0 * w + jj0), d_tr = __ldg(depth + (size_t)ii0 * w + jj1);
    const float d_bl = __ldg(depth + (size_t)ii1 * w + jj0), d_br = __ldg(depth + (size_t)ii1 * w + jj1);
    if (!(d_tl > 0.0f && d_tr > 0.0f && d_bl > 0.0f && d_br > 0.0f)) return 1;
    const float di = __fsub_rn(y, i0), dj = __fsub_rn(x, j0);
    const float omi = __fsub_rn(1.0f, di), omj = __fsub_rn(1.0f, dj);
    const float w_tl = __fmul_rn(omi, omj), w_tr = __fmul_rn(omi, dj), w_bl = __fmul_rn(di, omj), w_br = __fmul_rn(di, dj);
    *z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w_tl, d_tl), __fmul_rn(w_tr, d_tr)), __fmul_rn(w_bl, d_bl)),
                   __fmul_rn(w_br, d_br));
    return 2;
}

constexpr int SE3_NT = 1024;

__global__ void __launch_bounds__(SE3_NT) warp_se3_kernel(Se3Params p) {
    extern __shared__ int s_occ[];            // ids of occluded points (appended after the outside ones)
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    const int n = p.count ? p.count[b] : p.n_max;
    const float* d0 = p.depth0 + (size_t)b * p.h0 * p.w0;
    const float* d1 = p.depth1 + (size_t)b * p.h1 * p.w1;
    const float* Ki = p.kinv0 + (size_t)b * 9;
    const float* K1 = p.k1 + (size_t)b * 9;
    const float* T = p.pose + (size_t)b * 16;
    const float b0r = p.bbox0[b * 2 + 0], b0c = p.bbox0[b * 2 + 1], b1r = p.bbox1[b * 2 + 0], b1c = p.bbox1[b * 2 + 1];
    const float fw0 = (float)p.w0, fh0 = (float)p.h0, fw1 = (float)p.w1, fh1 = (float)p.h1;
    int n_in = 0, n_outside = 0, n_occ = 0;
    for (int base = 0; base < n; base += SE3_NT) {
        const int i = base + threadIdx.x;
        int cls = 0;                          // 0 dropped, 1 valid, 2 outside, 3 occluded
        float x = 0, y = 0, ux = 0, vy = 0;
        if (i < n) {
            const float* pt = p.pts + ((size_t)b * p.n_max + i) * p.stride;
            x = __fmul_rn(pt[0], fw0);                                        // projection.py:202
            y = __fmul_rn(pt[1], fh0);
            float z0;
            if (interp_depth(d0, p.h0, p.w0, x, y, &z0) == 2) {               // projection.py:209
                const float bx = __fadd_rn(__fadd_rn(x, b0c), 0.5f);          // COLMAP convention (:212)
                const float by = __fadd_rn(__fadd_rn(y, b0r), 0.5f);
                const float du = __fmul_rn(bx, z0), dv = __fmul_rn(by, z0);   // unproject (:44-48)
                const float X = fmaf(Ki[2], z0, fmaf(Ki[1], dv, __fmul_rn(Ki[0], du)));
                const float Y = fmaf(Ki[5], z0, fmaf(Ki[4], dv, __fmul_rn(Ki[3], du)));
                const float Z = fmaf(Ki[8], z0, fmaf(Ki[7], dv, __fmul_rn(Ki[6], du)));
                const float X1 = __fadd_rn(fmaf(T[2], Z, fmaf(T[1], Y, __fmul_rn(T[0], X))), T[3]);     // pose01 (:219)
                const float Y1 = __fadd_rn(fmaf(T[6], Z, fmaf(T[5], Y, __fmul_rn(T[4], X))), T[7]);
                const float Z1 = __fadd_rn(fmaf(T[10], Z, fmaf(T[9], Y, __fmul_rn(T[8], X))), T[11]);
                const float qx = fmaf(K1[2], Z1, fmaf(K1[1], Y1, __fmul_rn(K1[0], X1)));                // project (:71-77)
                const float qy = fmaf(K1[5], Z1, fmaf(K1[4], Y1, __fmul_rn(K1[3], X1)));
                const float qz = fmaf(K1[8], Z1, fmaf(K1[7], Y1, __fmul_rn(K1[6], X1)));
                ux = __fsub_rn(__fsub_rn(qx / qz, b1c), 0.5f);                // projection.py:225
                vy = __fsub_rn(__fsub_rn(qy / qz, b1r), 0.5f);
                float z1;
                const int r1 = interp_depth(d1, p.h1, p.w1, ux, vy, &z1);     // projection.py:232
                if (r1 == 0) cls = 2;                                         // :234-237 projected outside
                else if (r1 == 2) cls = (fabsf(__fsub_rn(qz, z1)) < 0.05f) ? 1 : 3;     // :244-247
            }
        }
        int t_in, t_outside, t_occ;
        const int o_in = n_in + kb::block_exclusive_scan(cls == 1 ? 1 : 0, s_scan, &t_in);
        const int o_outside = n_outside + kb::block_exclusive_scan(cls == 2 ? 1 : 0, s_scan, &t_outside);
        const int o_occ = n_occ + kb::block_exclusive_scan(cls == 3 ? 1 : 0, s_scan, &t_occ);
        if (cls == 1) {
            const size_t o = ((size_t)b * p.n_max + o_in) * 2;
            p.kp_valid[o + 0] = x / fw0;                                      // projection.py:265-266
            p.kp_valid[o + 1] = y / fh0;
            p.kp_warp[o + 0] = ux / fw1;
            p.kp_warp[o + 1] = vy / fh1;
            p.ids[(size_t)b * p.n_max + o_in] = i;
        } else if (cls == 2) {
            p.ids_out[(size_t)b * p.n_max + o_outside] = i;
        } else if (cls == 3) {
            s_occ[o_occ] = i;
        }
        n_in += t_in; n_outside += t_outside; n_occ += t_occ;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_occ; k += SE3_NT) p.ids_out[(size_t)b * p.n_max + n_outside + k] = s_occ[k];   // :259
    if (threadIdx.x == 0) { p.n_valid[b] = n_in; p.n_out[b] = n_outside + n_occ; }
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
           kb_align_up((size_t)B * 4, 256) * 3 + 1024 +
           kb_align_up((size_t)B * a_max * 16, 256) + kb_align_up((size_t)B * b_max * 16, 256) +     // sorted points
           kb_align_up((size_t)B * a_max * 4, 256) + kb_align_up((size_t)B * b_max * 4, 256);        // their indices
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
    int* need_bf = arena.take<int>(B);
    SortedSide s_b, s_a;                    // the B points sorted by k10c.y, the A points sorted by k0c.y
    s_b.pts = arena.take<float4>((size_t)B * b_max); s_b.idx = arena.take<int>((size_t)B * b_max);
    s_a.pts = arena.take<float4>((size_t)B * a_max); s_a.idx = arena.take<int>((size_t)B * a_max);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    p.only_flagged = nullptr;
    p.k0c = k0c; p.k01c = k01c; p.k1c = k1c; p.k10c = k10c; p.na = na; p.nb = nb; p.stats = stats;
    p.errors = errors; p.pairs = pairs; p.B = B; p.a_max = a_max; p.b_max = b_max; p.pair_cap = pair_cap;
    p.scale01 = scale01; p.scale10 = scale10; p.th = th;
    const size_t n_init = (size_t)B * (a_max > b_max ? a_max : b_max);
    const size_t n_init2 = n_init > (size_t)B * 4 ? n_init : (size_t)B * 4;
    rep_init_kernel<<<(unsigned)((n_init2 + 255) / 256), 256, 0, st>>>(p);
    KB_LAUNCH_CHECK();
    // pruned sweeps for every map whose distances are provably below the 99999 diagonal mask ...
    rep_bound_kernel<<<B, 256, 0, st>>>(p, need_bf);
    KB_LAUNCH_CHECK();
    const bool no_sort = kb_knobs[KB_KNOB_REP_NO_SORT] != 0;           // A/B timing and tests: the tile-walking kernels
    if (a_max <= SRT_MAX && b_max <= SRT_MAX && !no_sort) {
        int np2 = 1;
        while (np2 < (a_max > b_max ? a_max : b_max)) np2 <<= 1;
        rep_sort_kernel<<<dim3(B, 2), 1024, (size_t)np2 * 8, st>>>(p, need_bf, s_b, s_a);
        KB_LAUNCH_CHECK();
        const size_t sm_b = (size_t)b_max * 20, sm_a = (size_t)a_max * 20;
        KB_CUDA_TRY(cudaFuncSetAttribute(rep_min_sorted_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
        KB_CUDA_TRY(cudaFuncSetAttribute(rep_min_sorted_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a));
        KB_CUDA_TRY(cudaFuncSetAttribute(rep_mutual_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
        rep_min_sorted_kernel<0><<<dim3((a_max + SW_ROWS - 1) / SW_ROWS, B), SW_NT, sm_b, st>>>(p, need_bf, s_b);
        KB_LAUNCH_CHECK();
        rep_min_sorted_kernel<1><<<dim3((b_max + SW_ROWS - 1) / SW_ROWS, B), SW_NT, sm_a, st>>>(p, need_bf, s_a);
        KB_LAUNCH_CHECK();
        rep_mutual_sorted_kernel<<<dim3((a_max + SW_ROWS - 1) / SW_ROWS, B), SW_NT, sm_b, st>>>(p, need_bf, s_b);
        KB_LAUNCH_CHECK();
    } else {
        rep_min_pruned_kernel<0><<<dim3((a_max + PR - 1) / PR, B), PT, 0, st>>>(p, need_bf);
        KB_LAUNCH_CHECK();
        rep_min_pruned_kernel<1><<<dim3((b_max + PR - 1) / PR, B), PT, 0, st>>>(p, need_bf);
        KB_LAUNCH_CHECK();
        rep_mutual_pruned_kernel<<<dim3((a_max + PR - 1) / PR, B), PT, 0, st>>>(p, need_bf);
        KB_LAUNCH_CHECK();
    }
    // ... the exhaustive sweeps for the rest (they return at once for all other maps)
    dim3 grid((a_max + RB - 1) / RB, B);
    // exhaustive kernels: a handful of blocks per map (they loop over the row blocks); in the common case every
    // block returns at once, and 16 k empty blocks cost more than the sorted sweeps
    dim3 grid_x(grid.x < 8 ? grid.x : 8, B);
    p.only_flagged = need_bf;
    rep_minima_kernel<<<grid_x, RT, 0, st>>>(p);
    KB_LAUNCH_CHECK();
    RepParams q = p;
    q.pairs = nullptr;
    rep_mutual_kernel<<<grid_x, RT, 0, st>>>(q);
    KB_LAUNCH_CHECK();
    p.only_flagged = nullptr;
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

extern "C" int kb_warp_se3(const float* pts, int pts_stride, const int* count, int B, int n_max, const float* depth0,
                           int h0, int w0, const float* depth1, int h1, int w1, const float* kinv0, const float* k1,
                           const float* pose01, const float* bbox0, const float* bbox1, float* kp_valid,
                           float* kp_warp, int* ids, int* ids_out, int* n_valid, int* n_out, kb_stream_t stream) {
    if (!pts || !depth0 || !depth1 || !kinv0 || !k1 || !pose01 || !bbox0 || !bbox1 || !kp_valid || !kp_warp || !ids ||
        !ids_out || !n_valid || !n_out || B <= 0 || n_max <= 0 || pts_stride < 2 || h0 <= 0 || w0 <= 0 || h1 <= 0 || w1 <= 0)
        return KB_ERR_BAD_ARG;
    if ((size_t)n_max * 4 > 200 * 1024) return KB_ERR_UNSUPPORTED;
    Se3Params p;
    p.pts = pts; p.count = count; p.depth0 = depth0; p.depth1 = depth1; p.kinv0 = kinv0; p.k1 = k1; p.pose = pose01;
    p.bbox0 = bbox0; p.bbox1 = bbox1; p.kp_valid = kp_valid; p.kp_warp = kp_warp; p.ids = ids; p.ids_out = ids_out;
    p.n_valid = n_valid; p.n_out = n_out; p.stride = pts_stride; p.B = B; p.n_max = n_max;
    p.h0 = h0; p.w0 = w0; p.h1 = h1; p.w1 = w1;
    const size_t smem = (size_t)n_max * 4;
    KB_CUDA_TRY(cudaFuncSetAttribute(warp_se3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    warp_se3_kernel<<<B, SE3_NT, smem, (cudaStream_t)stream>>>(p);
    KB_LAUNCH_CHECK();
    return KB_OK;
}
