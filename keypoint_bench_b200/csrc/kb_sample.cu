// Bilinear descriptor sampling at keypoints, optional fused L2 normalisation.
// Reference: utils/matcher.py:221-226 (grid_sample, align_corners=True, no normalisation) and
// models/lightglue.py:24-41 (pixel keypoints, F.normalize).
//
// Descriptor maps arrive NCHW, so one keypoint's C taps are C strided 4-byte reads; a warp owns
// one keypoint, lanes stride over channels (each lane reads the 2x2 taps of its channel, the two
// x-neighbours share a 32-byte sector), and the [n,C] output row is written coalesced.  The L2
// norm is a warp-shuffle reduction over the lanes' partial sums.
#include "kb_common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

constexpr int NT = 256;   // 8 warps = 8 keypoints per block

struct SampleParams {
    const float* desc;   // [B,C,h,w]
    const float* pts;    // [B,n_max,stride]
    const int* count;    // [B] or null
    float* out;          // [B,n_max,C]
    int B, C, h, w, n_max, stride, normalize, coord_mode, s;
};

__global__ void __launch_bounds__(NT) sample_kernel(SampleParams p) {
    const int warp = (blockIdx.x * NT + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (warp >= p.n_max) return;
    const int n = p.count ? p.count[b] : p.n_max;
    if (warp >= n) return;
    const float* pt = p.pts + ((size_t)b * p.n_max + warp) * p.stride;
    const float kx = pt[0], ky = pt[1];
    float gx, gy;
    if (p.coord_mode == 0) {                        // matcher.py:221-222
        gx = (kx - 0.5f) * 2.0f;
        gy = (ky - 0.5f) * 2.0f;
    } else {                                        // lightglue.py:27-33
        const float s = (float)p.s;
        const float ax = kx - s / 2.0f + 0.5f, ay = ky - s / 2.0f + 0.5f;
        gx = ax / ((float)p.w * s - s / 2.0f - 0.5f) * 2.0f - 1.0f;
        gy = ay / ((float)p.h * s - s / 2.0f - 0.5f) * 2.0f - 1.0f;
    }
    // grid_sample un-normalisation with align_corners=True: ((g+1)/2)*(size-1)
    const float ix = ((gx + 1.0f) / 2.0f) * (float)(p.w - 1);
    const float iy = ((gy + 1.0f) / 2.0f) * (float)(p.h - 1);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
    const float w_nw = (fx1 - ix) * (fy1 - iy);
    const float w_ne = (ix - fx0) * (fy1 - iy);
    const float w_sw = (fx1 - ix) * (iy - fy0);
    const float w_se = (ix - fx0) * (iy - fy0);
    // out-of-range coordinates (incl. NaN/inf) sample zeros
    const bool finite = (ix > -2.0f) && (ix < (float)p.w + 1.0f) && (iy > -2.0f) && (iy < (float)p.h + 1.0f);
    const int x0 = finite ? (int)fx0 : -8, y0 = finite ? (int)fy0 : -8;
    const int x1 = x0 + 1, y1 = y0 + 1;
    const bool in_x0 = x0 >= 0 && x0 < p.w, in_x1 = x1 >= 0 && x1 < p.w;
    const bool in_y0 = y0 >= 0 && y0 < p.h, in_y1 = y1 >= 0 && y1 < p.h;
    const size_t plane = (size_t)p.h * p.w;
    const float* base = p.desc + (size_t)b * p.C * plane;
    float* o = p.out + ((size_t)b * p.n_max + warp) * p.C;
    float ss = 0.0f;
    for (int c0 = 0; c0 < p.C; c0 += 32) {
        const int c = c0 + lane;
        float val = 0.0f;
        if (c < p.C) {
            const float* m = base + (size_t)c * plane;
            const float nw = (in_x0 && in_y0) ? __ldg(m + (size_t)y0 * p.w + x0) : 0.0f;
            const float ne = (in_x1 && in_y0) ? __ldg(m + (size_t)y0 * p.w + x1) : 0.0f;
            const float sw = (in_x0 && in_y1) ? __ldg(m + (size_t)y1 * p.w + x0) : 0.0f;
            const float se = (in_x1 && in_y1) ? __ldg(m + (size_t)y1 * p.w + x1) : 0.0f;
            val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(nw, w_nw), __fmul_rn(ne, w_ne)), __fmul_rn(sw, w_sw)),
                            __fmul_rn(se, w_se));
            if (!p.normalize) o[c] = val;
        }
        ss += val * val;
    }
    if (p.normalize) {
        for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
        const float denom = fmaxf(sqrtf(ss), 1e-12f);          // F.normalize eps (lightglue.py:38-40)
        // second pass re-reads the taps from L1/L2 (C <= 256: a few lines per lane)
        for (int c0 = 0; c0 < p.C; c0 += 32) {
            const int c = c0 + lane;
            if (c < p.C) {
                const float* m = base + (size_t)c * plane;
                const float nw = (in_x0 && in_y0) ? __ldg(m + (size_t)y0 * p.w + x0) : 0.0f;
                const float ne = (in_x1 && in_y0) ? __ldg(m + (size_t)y0 * p.w + x1) : 0.0f;
                const float sw = (in_x0 && in_y1) ? __ldg(m + (size_t)y1 * p.w + x0) : 0.0f;
                const float se = (in_x1 && in_y1) ? __ldg(m + (size_t)y1 * p.w + x1) : 0.0f;
                const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(nw, w_nw), __fmul_rn(ne, w_ne)),
                                                      __fmul_rn(sw, w_sw)), __fmul_rn(se, w_se));
                o[c] = val / denom;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Plane-staged variant for low-resolution maps (SuperPoint / XFeat: 60x80 planes, ~1000+ keypoints):
// nearly every pixel of the map is touched by some keypoint, so instead of gathering C strided
// 4-byte taps per keypoint (4x sector amplification, L2-bound), a CTA pulls FOUR whole channel planes
// into shared memory with one bulk async copy (contiguous in NCHW, each map byte leaves HBM exactly
// once), then every thread interpolates its keypoints out of shared memory and writes the four
// channels as one 16-byte piece of the [n,C] row.  Two CTAs per SM overlap one CTA's copy with the
// other's arithmetic.  Same arithmetic (and rounding order) as sample_kernel.
constexpr int PL_C = 4;          // planes per CTA
constexpr int PL_NT = 256;

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(PL_NT, 2) sample_planes_kernel(SampleParams p) {
    extern __shared__ __align__(128) unsigned char pl_smem[];
    __shared__ __align__(8) unsigned long long pl_bar;
    float* planes = reinterpret_cast<float*>(pl_smem);
    const int b = blockIdx.y, c0 = blockIdx.x * PL_C;
    const int nc = min(PL_C, p.C - c0);
    const int hw = p.h * p.w;
    const int n = p.count ? p.count[b] : p.n_max;
    if (n <= 0) return;
    const float* src = p.desc + ((size_t)b * p.C + c0) * hw;
    const uint32_t bytes = (uint32_t)nc * hw * 4u;
    const bool bulk = ((reinterpret_cast<uintptr_t>(src) | bytes) & 15u) == 0;
    const uint32_t bar = smem_addr_u32(&pl_bar);
    if (bulk) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            // <= 128 KB per copy; split so that each piece stays well inside the instruction's size field
            uint32_t done = 0;
            while (done < bytes) {
                const uint32_t piece = min(bytes - done, 32768u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_addr_u32(pl_smem + done)), "l"(reinterpret_cast<const char*>(src) + done),
                               "r"(piece), "r"(bar) : "memory");
                done += piece;
            }
        }
    } else {
        for (int i = threadIdx.x; i < nc * hw; i += PL_NT) planes[i] = __ldg(src + i);
    }
    // keypoint coordinates go to shared memory while the plane copy is in flight (their global-load latency
    // would otherwise sit on every thread's critical path once per keypoint)
    float2* kxy = reinterpret_cast<float2*>(pl_smem + (size_t)PL_C * hw * 4);
    for (int k = threadIdx.x; k < n; k += PL_NT) {
        const float* pt = p.pts + ((size_t)b * p.n_max + k) * p.stride;
        kxy[k] = make_float2(__ldg(pt), __ldg(pt + 1));
    }
    __syncthreads();
    if (bulk) {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
        }
    }
    for (int k = threadIdx.x; k < n; k += PL_NT) {
        const float2 kp = kxy[k];
        const float kx = kp.x, ky = kp.y;
        float gx, gy;
        if (p.coord_mode == 0) {                        // matcher.py:221-222
            gx = (kx - 0.5f) * 2.0f;
            gy = (ky - 0.5f) * 2.0f;
        } else {                                        // lightglue.py:27-33
            const float s = (float)p.s;
            const float ax = kx - s / 2.0f + 0.5f, ay = ky - s / 2.0f + 0.5f;
            gx = ax / ((float)p.w * s - s / 2.0f - 0.5f) * 2.0f - 1.0f;
            gy = ay / ((float)p.h * s - s / 2.0f - 0.5f) * 2.0f - 1.0f;
        }
        const float ix = ((gx + 1.0f) / 2.0f) * (float)(p.w - 1);
        const float iy = ((gy + 1.0f) / 2.0f) * (float)(p.h - 1);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
        const float w_nw = (fx1 - ix) * (fy1 - iy);
        const float w_ne = (ix - fx0) * (fy1 - iy);
        const float w_sw = (fx1 - ix) * (iy - fy0);
        const float w_se = (ix - fx0) * (iy - fy0);
        const bool finite = (ix > -2.0f) && (ix < (float)p.w + 1.0f) && (iy > -2.0f) && (iy < (float)p.h + 1.0f);
        const int x0 = finite ? (int)fx0 : -8, y0 = finite ? (int)fy0 : -8;
        const int x1 = x0 + 1, y1 = y0 + 1;
        const bool in_x0 = x0 >= 0 && x0 < p.w, in_x1 = x1 >= 0 && x1 < p.w;
        const bool in_y0 = y0 >= 0 && y0 < p.h, in_y1 = y1 >= 0 && y1 < p.h;
        const bool t_nw = in_x0 && in_y0, t_ne = in_x1 && in_y0, t_sw = in_x0 && in_y1, t_se = in_x1 && in_y1;
        const int o_nw = y0 * p.w + x0, o_ne = o_nw + 1, o_sw = o_nw + p.w, o_se = o_sw + 1;
        KB_ASSERT(!t_nw || (o_nw >= 0 && o_nw < hw));
        KB_ASSERT(!t_se || (o_se >= 0 && o_se < hw));
        float val[PL_C];
#pragma unroll
        for (int c = 0; c < PL_C; ++c) {
            val[c] = 0.0f;
            if (c < nc) {
                const float* m = planes + c * hw;
                const float nw = t_nw ? m[o_nw] : 0.0f;
                const float ne = t_ne ? m[o_ne] : 0.0f;
                const float sw = t_sw ? m[o_sw] : 0.0f;
                const float se = t_se ? m[o_se] : 0.0f;
                val[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(nw, w_nw), __fmul_rn(ne, w_ne)), __fmul_rn(sw, w_sw)),
                                   __fmul_rn(se, w_se));
            }
        }
        float* o = p.out + ((size_t)b * p.n_max + k) * p.C + c0;
        if (PL_C == 4 && nc == PL_C && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0)) {
            *reinterpret_cast<float4*>(o) = make_float4(val[0], val[1], val[2], val[PL_C - 1]);
        } else if (PL_C == 2 && nc == PL_C && ((reinterpret_cast<uintptr_t>(o) & 7u) == 0)) {
            *reinterpret_cast<float2*>(o) = make_float2(val[0], val[1]);
        } else {
            for (int c = 0; c < nc; ++c) o[c] = val[c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Plane-staged sampling, eight channels per CTA -- the default for low-resolution maps (OPERANDS = false) -- and the same
// FUSED with the matcher's operand preparation (OPERANDS = true; kb_match_tc.cu: prep_kernel): a CTA owns
// EIGHT channels of one map -- four stages of two planes through two shared-memory buffers -- and every thread keeps
// the eight samples of its (up to four) keypoints in registers, so that it can write, per keypoint, the 32-byte piece
// of the float32 [n,C] row (a whole sector; the certification stages of the matcher read it), the 16-byte pieces of the
// hi and lo halves of the matcher's operand row, and the partial |row|^2 of the eight components.  The float32 rows no
// longer travel HBM -> SM -> HBM a second time just to be split.  Same sampling arithmetic as sample_kernel.
constexpr int OP_C = 8;           // channels per CTA
// planes per stage (SC) and buffers (NBUF): OP_C / SC stages through NBUF buffers of SC planes, the copy of stage s+NBUF is
// issued as soon as stage s has been read and overlaps the interpolation of the stages in between
constexpr int OP_KP = 4;          // keypoints per thread: n_max <= OP_KP * PL_NT

struct OperandParams {
    SampleParams s;
    unsigned short* S[2];         // operand rows of side 0 (maps [0, pairs)) and side 1 (maps [pairs, 2 pairs))
    float* part[2];
    int pairs, Dp, fp16;
};

__device__ __forceinline__ void split16_pair(float f, int fp16, unsigned short& h, unsigned short& l) {
    if (fp16) {
        const __half hh = __float2half_rn(f);
        h = __half_as_ushort(hh);
        l = __half_as_ushort(__float2half_rn(f - __half2float(hh)));
    } else {
        const __nv_bfloat16 hh = __float2bfloat16_rn(f);
        h = __bfloat16_as_ushort(hh);
        l = __bfloat16_as_ushort(__float2bfloat16_rn(f - __bfloat162float(hh)));
    }
}

template <bool OPERANDS, int OP_SC, int NBUF>
__global__ void __launch_bounds__(PL_NT, 2) sample_planes_operands_kernel(OperandParams q) {
    constexpr int OP_STAGES = OP_C / OP_SC;
    const SampleParams& p = q.s;
    extern __shared__ __align__(128) unsigned char pl_smem[];
    __shared__ __align__(8) unsigned long long pl_bar[NBUF];
    float* planes = reinterpret_cast<float*>(pl_smem);
    const int b = blockIdx.y, c0 = blockIdx.x * OP_C;
    const int hw = p.h * p.w;
    const int n = p.count ? p.count[b] : p.n_max;
    const int side = b >= q.pairs ? 1 : 0, bs = b - side * q.pairs;
    const float* src = p.desc + ((size_t)b * p.C + c0) * hw;
    const uint32_t bytes = (uint32_t)OP_SC * hw * 4u;                // one stage (host: 16-byte aligned, a multiple of 16)
    const uint32_t bar0 = smem_addr_u32(&pl_bar[0]);
    auto issue = [&](int stage) {                                   // thread 0: planes c0 + OP_SC*stage .. into buffer stage % NBUF
        const uint32_t bar = bar0 + 8u * (stage % NBUF);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        uint32_t done = 0;
        while (done < bytes) {
            const uint32_t piece = min(bytes - done, 32768u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr_u32(pl_smem + (size_t)(stage % NBUF) * bytes + done)),
                           "l"(reinterpret_cast<const char*>(src) + (size_t)stage * bytes + done), "r"(piece), "r"(bar) : "memory");
            done += piece;
        }
    };
    if (n > 0 && threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NBUF; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * i) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int i = 0; i < NBUF && i < OP_STAGES; ++i) issue(i);
    }
    __syncthreads();                                                // nobody polls a barrier before it is initialised
    // this thread's keypoints (k = threadIdx.x + i * PL_NT): taps and weights, while the first copy is in flight
    float w_nw[OP_KP], w_ne[OP_KP], w_sw[OP_KP], w_se[OP_KP];
    int o_nw[OP_KP];
    unsigned taps[OP_KP];
#pragma unroll
    for (int i = 0; i < OP_KP; ++i) {
        const int k = threadIdx.x + i * PL_NT;
        taps[i] = 0u; o_nw[i] = 0;
        w_nw[i] = w_ne[i] = w_sw[i] = w_se[i] = 0.0f;
        if (k < n) {
            const float* pt = p.pts + ((size_t)b * p.n_max + k) * p.stride;
            const float kx = __ldg(pt), ky = __ldg(pt + 1);
            float gx, gy;
            if (p.coord_mode == 0) {                        // matcher.py:221-222
                gx = (kx - 0.5f) * 2.0f;
                gy = (ky - 0.5f) * 2.0f;
            } else {                                        // lightglue.py:27-33
                const float s = (float)p.s;
                const float ax = kx - s / 2.0f + 0.5f, ay = ky - s / 2.0f + 0.5f;
                gx = ax / ((float)p.w * s - s / 2.0f - 0.5f) * 2.0f - 1.0f;
                gy = ay / ((float)p.h * s - s / 2.0f - 0.5f) * 2.0f - 1.0f;
            }
            const float ix = ((gx + 1.0f) / 2.0f) * (float)(p.w - 1);
            const float iy = ((gy + 1.0f) / 2.0f) * (float)(p.h - 1);
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
            w_nw[i] = (fx1 - ix) * (fy1 - iy);
            w_ne[i] = (ix - fx0) * (fy1 - iy);
            w_sw[i] = (fx1 - ix) * (iy - fy0);
            w_se[i] = (ix - fx0) * (iy - fy0);
            const bool finite = (ix > -2.0f) && (ix < (float)p.w + 1.0f) && (iy > -2.0f) && (iy < (float)p.h + 1.0f);
            const int x0 = finite ? (int)fx0 : -8, y0 = finite ? (int)fy0 : -8;
            const int x1 = x0 + 1, y1 = y0 + 1;
            const bool in_x0 = x0 >= 0 && x0 < p.w, in_x1 = x1 >= 0 && x1 < p.w;
            const bool in_y0 = y0 >= 0 && y0 < p.h, in_y1 = y1 >= 0 && y1 < p.h;
            taps[i] = (in_x0 && in_y0 ? 1u : 0u) | (in_x1 && in_y0 ? 2u : 0u) | (in_x0 && in_y1 ? 4u : 0u) | (in_x1 && in_y1 ? 8u : 0u);
            o_nw[i] = y0 * p.w + x0;
            KB_ASSERT(!(taps[i] & 1u) || (o_nw[i] >= 0 && o_nw[i] < hw));
            KB_ASSERT(!(taps[i] & 8u) || (o_nw[i] + p.w + 1 >= 0 && o_nw[i] + p.w + 1 < hw));
        }
    }
    float val[OP_KP][OP_C];
#pragma unroll
    for (int stage = 0; stage < OP_STAGES; ++stage) {
        if (n > 0) {
            uint32_t ok = 0;
            while (!ok) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok) : "r"(bar0 + 8u * (stage % NBUF)), "r"((uint32_t)((stage / NBUF) & 1)) : "memory");
            }
        }
        const float* buf = planes + (size_t)(stage % NBUF) * OP_SC * hw;
#pragma unroll
        for (int i = 0; i < OP_KP; ++i) {
#pragma unroll
            for (int c = 0; c < OP_SC; ++c) {
                const float* m = buf + c * hw;
                const float nw = (taps[i] & 1u) ? m[o_nw[i]] : 0.0f;
                const float ne = (taps[i] & 2u) ? m[o_nw[i] + 1] : 0.0f;
                const float sw = (taps[i] & 4u) ? m[o_nw[i] + p.w] : 0.0f;
                const float se = (taps[i] & 8u) ? m[o_nw[i] + p.w + 1] : 0.0f;
                val[i][stage * OP_SC + c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(nw, w_nw[i]), __fmul_rn(ne, w_ne[i])),
                                                                 __fmul_rn(sw, w_sw[i])), __fmul_rn(se, w_se[i]));
            }
        }
        if (stage + NBUF < OP_STAGES && n > 0) {
            __syncthreads();                                        // everyone is done with this buffer
            if (threadIdx.x == 0) issue(stage + NBUF);
        }
    }
    // rows: float32 piece, operand halves, partial |row|^2; rows beyond the count get zero operands (as prep_kernel writes them)
    const int nparts = q.Dp / OP_C;
#pragma unroll
    for (int i = 0; i < OP_KP; ++i) {
        const int k = threadIdx.x + i * PL_NT;
        if (k >= p.n_max) continue;
        const size_t row = (size_t)bs * p.n_max + k;
        if (k < n) {
            float* o = p.out + ((size_t)b * p.n_max + k) * p.C + c0;
            *reinterpret_cast<float4*>(o) = make_float4(val[i][0], val[i][1], val[i][2], val[i][3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(val[i][4], val[i][5], val[i][6], val[i][7]);
        }
        if (!OPERANDS) continue;
        unsigned short* srow = (side ? q.S[1] : q.S[0]) + row * (2 * (size_t)q.Dp) + c0;
        if (k < n) {
            uint32_t hw2[OP_C / 2], lw2[OP_C / 2];                  // two 16-bit halves per word
            float ss = 0.0f;
#pragma unroll
            for (int e = 0; e < OP_C; e += 2) {
                unsigned short h0, l0, h1, l1;
                split16_pair(val[i][e], q.fp16, h0, l0);
                split16_pair(val[i][e + 1], q.fp16, h1, l1);
                hw2[e / 2] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                lw2[e / 2] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                ss = fmaf(val[i][e], val[i][e], ss);
                ss = fmaf(val[i][e + 1], val[i][e + 1], ss);
            }
            *reinterpret_cast<uint4*>(srow) = make_uint4(hw2[0], hw2[1], hw2[2], hw2[3]);
            *reinterpret_cast<uint4*>(srow + q.Dp) = make_uint4(lw2[0], lw2[1], lw2[2], lw2[3]);
            (side ? q.part[1] : q.part[0])[row * nparts + blockIdx.x] = ss;
        } else {
            *reinterpret_cast<uint4*>(srow) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(srow + q.Dp) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// L2 normalisation of the sampled rows in place (lightglue.py:38-40: x / max(||x||_2, 1e-12)), one warp per row, with
// the summation order of sample_kernel's fused norm (lanes stride over the channels, xor-shuffle tree), so that the
// plane-staged sampler + this kernel give the same bits as the gather kernel with normalize = 1.
__global__ void __launch_bounds__(256) normalize_rows_kernel(float* out, const int* count, int n_max, int C) {
    const int b = blockIdx.y;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int n = count ? count[b] : n_max;
    if (row >= n) return;
    float* o = out + ((size_t)b * n_max + row) * C;
    float ss = 0.0f;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const float v = c < C ? o[c] : 0.0f;
        ss += v * v;
    }
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < C; c += 32) o[c] = o[c] / denom;
}

}  // namespace

extern "C" int kb_sample_desc(const float* desc, int B, int C, int h, int w, const float* pts, int pts_stride,
                              const int* count, int n_max, int normalize, int coord_mode, int s, float* out,
                              kb_stream_t stream) {
    if (!desc || !pts || !out || B <= 0 || C <= 0 || h <= 0 || w <= 0 || n_max <= 0 || pts_stride < 2)
        return KB_ERR_BAD_ARG;
    if (coord_mode != 0 && coord_mode != 1) return KB_ERR_BAD_ARG;
    if (B > 65535) return KB_ERR_UNSUPPORTED;
    SampleParams p;
    p.desc = desc; p.pts = pts; p.count = count; p.out = out;
    p.B = B; p.C = C; p.h = h; p.w = w; p.n_max = n_max; p.stride = pts_stride;
    p.normalize = normalize; p.coord_mode = coord_mode; p.s = s;
    // low-resolution, densely sampled maps: stage whole planes (see sample_planes_kernel)
    const size_t plane_bytes = (size_t)h * w * 4;
    // ... eight channels per CTA through double-buffered two-plane stages with 32-byte output pieces where the shapes
    // allow it (138 us against 158 us for the four-channel kernel at cfg2; kb_debug_knob(KB_KNOB_SAMPLE_4CH, 1) forces
    // the four-channel kernel; identical bits)
    if (!kb_knobs[KB_KNOB_SAMPLE_4CH] && C % OP_C == 0 && n_max <= OP_KP * PL_NT && (plane_bytes % 16) == 0 &&
        plane_bytes * PL_C <= 110 * 1024 && (size_t)n_max * 16 > (size_t)h * w &&
        ((reinterpret_cast<uintptr_t>(desc) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
        OperandParams q;
        q.s = p; q.s.normalize = 0;           // the staged kernels write raw samples; rows are normalised afterwards
        q.S[0] = q.S[1] = nullptr; q.part[0] = q.part[1] = nullptr; q.pairs = B; q.Dp = 0; q.fp16 = 0;
        const size_t smem = plane_bytes * PL_C;
        KB_CUDA_TRY(cudaFuncSetAttribute(sample_planes_operands_kernel<false, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sample_planes_operands_kernel<false, 2, 2><<<dim3(C / OP_C, B), PL_NT, smem, (cudaStream_t)stream>>>(q);
        KB_LAUNCH_CHECK();
        if (normalize) {
            normalize_rows_kernel<<<dim3((n_max + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(out, count, n_max, C);
            KB_LAUNCH_CHECK();
        }
        return KB_OK;
    }
    if (plane_bytes * PL_C + (size_t)n_max * 8 <= 110 * 1024 && (size_t)n_max * 16 > (size_t)h * w) {
        const size_t smem = plane_bytes * PL_C + (size_t)n_max * 8;
        KB_CUDA_TRY(cudaFuncSetAttribute(sample_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((C + PL_C - 1) / PL_C, B);
        p.normalize = 0;                      // the staged kernel writes raw samples; rows are normalised afterwards
        sample_planes_kernel<<<grid, PL_NT, smem, (cudaStream_t)stream>>>(p);
        KB_LAUNCH_CHECK();
        if (normalize) {
            normalize_rows_kernel<<<dim3((n_max + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(out, count, n_max, C);
            KB_LAUNCH_CHECK();
        }
        return KB_OK;
    }
    dim3 grid((n_max * 32 + NT - 1) / NT, B);
    sample_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(p);
    KB_LAUNCH_CHECK();
    return KB_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused entry: kb_sample_desc(normalize = 0) for the 2 * pairs maps of a batch of pairs (maps [0, pairs) are image 0,
// maps [pairs, 2 pairs) image 1) that ALSO writes the operand rows of the kb_match_mnn(algo 1) call that follows on
// (out[0 : pairs], out[pairs : 2 pairs]) into that call's workspace; the match call then passes phases 6 | 8.
extern "C" int kb_sample_desc_operands_supported(int C, int h, int w, int n_max) {
    const size_t plane_bytes = (size_t)h * w * 4;
    return C > 0 && C % 64 == 0 && C <= 256 && n_max > 0 && n_max <= OP_KP * PL_NT && (plane_bytes % 16) == 0 &&
           plane_bytes * PL_C <= 110 * 1024 && (size_t)n_max * 16 > (size_t)h * w;
}

extern "C" int kb_sample_desc_operands(const float* desc, int pairs, int C, int h, int w, const float* pts, int pts_stride,
                                       const int* count, int n_max, int coord_mode, int s, float* out, void* match_ws,
                                       size_t match_ws_bytes, kb_stream_t stream) {
    if (!desc || !pts || !out || !match_ws || pairs <= 0 || C <= 0 || h <= 0 || w <= 0 || n_max <= 0 || pts_stride < 2)
        return KB_ERR_BAD_ARG;
    if (coord_mode != 0 && coord_mode != 1) return KB_ERR_BAD_ARG;
    if (!kb_sample_desc_operands_supported(C, h, w, n_max) || 2 * pairs > 65535) return KB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(desc) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return KB_ERR_UNSUPPORTED;
    KbOperandSinks sinks;
    const int rc = kb_match_tc_operand_sinks(match_ws, match_ws_bytes, pairs, n_max, n_max, C, &sinks);
    if (rc != KB_OK) return rc;
    if (sinks.Dp != C) return KB_ERR_UNSUPPORTED;
    OperandParams q;
    q.s.desc = desc; q.s.pts = pts; q.s.count = count; q.s.out = out;
    q.s.B = 2 * pairs; q.s.C = C; q.s.h = h; q.s.w = w; q.s.n_max = n_max; q.s.stride = pts_stride;
    q.s.normalize = 0; q.s.coord_mode = coord_mode; q.s.s = s;
    q.S[0] = sinks.S[0]; q.S[1] = sinks.S[1]; q.part[0] = sinks.part[0]; q.part[1] = sinks.part[1];
    q.pairs = pairs; q.Dp = sinks.Dp; q.fp16 = sinks.fp16;
    const size_t smem = (size_t)h * w * 4 * PL_C;
    KB_CUDA_TRY(cudaFuncSetAttribute(sample_planes_operands_kernel<true, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_planes_operands_kernel<true, 2, 2><<<dim3(C / OP_C, 2 * pairs), PL_NT, smem, (cudaStream_t)stream>>>(q);
    KB_LAUNCH_CHECK();
    return KB_OK;
}
