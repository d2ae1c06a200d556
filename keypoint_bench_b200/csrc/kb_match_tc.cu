// Tensor-core mutual-NN matcher (algo 1): tcgen05 Gram tiles in TMEM with a fused top-2 epilogue,
// float64 certification of the winners.
// Reference semantics: utils/matcher.py:227-234 -> skimage match_descriptors over float64 cdist.
//
// For a query row x and database rows y_j:  argmin_j |x-y_j|^2  ==  argmax_j  t_j = x.y_j - |y_j|^2/2.
//   prep_kernel     float32 descriptors -> split operand rows [hi(D) | lo(D)] of 16-bit floats, c_j = -|y_j|^2/2, row
//                   norms.  Default for D > 64: fp16 halves (x = hi + lo up to 2^-22 |x|) and TWO products per Gram,
//                   (hi_x + lo_x).hi_y -- the database side is represented by its fp16 rounding alone, so x.y is
//                   reproduced to 2^-12 |x||y| (a-priori bound, part of the certification), at two thirds of the MMA
//                   work and half the database bytes of the alternative (the default for D <= 64, where the epilogue is
//                   the bound; KB_KNOB_TC_BF16X3 = 1 / 2 forces either): bf16 halves and THREE
//                   products hi.hi + lo.hi + hi.lo, ~2^-16 relative.  Either way the pairs are certified exactly; the
//                   wider bound only sends ~3 % instead of ~1 % of the unmatched rows to the two-candidate float64 check.
//   nn_top2_kernel  persistent, warp-specialised, 640 threads: warp 0 TMA producer, warp 1 MMA issuer (both run
//                   their loops warp-uniformly and predicate only the issue on lane 0), warp 2 TMEM allocation,
//                   16 epilogue warps (4 TMEM lane quadrants x 4 column slices).  tcgen05.mma kind::f16,
//                   M=128 N=256 K=16, two 256-column accumulators in TMEM (all 512 columns), so the epilogue of
//                   one column tile overlaps the MMAs of the next.  The query tile's operand rows stay resident
//                   in shared memory while the database blocks (256 rows x 64 k, 128B-swizzled K-major) stream
//                   through a TMA ring sized from the free shared memory.  The epilogue folds every 128x256 tile
//                   into per-row running (best, second, third, argbest, argsecond) with software-pipelined
//                   tcgen05.ld.32x32b.x16; the [n,m] matrix never leaves the SM.  Both directions (rows->cols,
//                   cols->rows for the cross-check) are work items of the same launch.  KB_KNOB_TC_ONE_PASS selects
//                   the one-pass cross-check instead: while a warp holds 32 rows x 16 columns of a tile it also
//                   reduces every COLUMN over its 32 rows (redux.sync.max on the order-preserving bits of
//                   u = x.y - |x|^2/2 + offset) to the (largest, second largest) of that 32-row group, stored per
//                   (group, column); whether row i is column j's best is then decided from those group maxima
//                   (resolve / gate below), exactly or with a float64 rescan of the column when it is too close to
//                   call.  Same pairs; measured SLOWER than the second Gram (two dependent redux.sync per column cost
//                   more than folding the transposed tile in-lane), so it is not the default.  nn_top2_kernel<2, .> is the
//                   CTA-pair variant (cluster of 2, cta_group::2, M=256), selectable with kb_debug_knob(KB_KNOB_TC_CLUSTER, 2).
//   resolve_kernel  merges the column slices; best-second above twice the a-priori error bound of the split
//                   product certifies the argmax; best-third above it leaves two candidates that are compared
//                   exactly in float64; anything else is queued for rescan_kernel, an exact float64 scan of the
//                   row (first of ties, as np.argmin).
//                   With the one-pass cross-check its second grid plane reduces the group maxima of every column to
//                   (largest, largest of the rest, group of the largest).
//   gate_kernel     mutual check + strict < max_distance.  Row i (best column j, column-side score u(i,j)) is mutual
//                   when u(i,j) is the largest of its group and beats everything else of column j by more than twice
//                   the error bound, not mutual when something beats it by more than that, and otherwise queued for an
//                   exact float64 rescan of column j (second rescan launch, second gate launch for those rows).  The
//                   distance gate is decided from the certified tensor-core score when the caller wants pairs only,
//                   float64 distance inside the error band or when distances are returned.
//   pairs_kernel    ordered compaction (per pair).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include "kb_common.cuh"

namespace kbtc {

constexpr int BM = 128, BN = 256, BK = 64;       // BN = 256: one tcgen05.mma covers 128x256x16 (the issue rate of
                                                 // a single thread, ~70 cycles, cannot feed N = 128 instructions)
constexpr int TILE_BYTES = BM * BK * 2;          // 16 KB, one 128B-swizzled K-major tile
constexpr int MAX_SLOTS = 8;                     // ring depth is chosen on the host from the free shared memory
constexpr int EPI_SLICES = 4;                    // column slices of a tile, one epilogue warp per (lane quadrant, slice)
constexpr int EPI_WARPS = 4 * EPI_SLICES;        // 16 epilogue warps: four per SM sub-partition hide each other's latencies
constexpr int NT = 128 + 32 * EPI_WARPS;         // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4-19 epilogue
constexpr int TMEM_COLS = 512;                   // two 256-column fp32 accumulators (all of TMEM)
constexpr int SLOT_BYTES = 2 * TILE_BYTES;       // one database block: 256 rows x 64 k (two TMA boxes)

struct Top2 {                 // per query row: three best scores, indices of the best two (32 bytes)
    float best, second, third;
    float u1;                 // column-side score u = x.y - |x|^2/2 + offset of (row, idx) -- one-pass cross-check
    int idx, idx2;
    float u2;                 // ... of (row, idx2)
    int pad;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        // a protocol bug must fail loudly, not hang the GPU: give up after ~2 s of waiting
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) __trap();
    }
}
// mbar_wait that adds the cycles spent waiting to `acc` when profiling is on
__device__ __forceinline__ void mbar_wait_prof(uint32_t bar, uint32_t parity, bool on, long long& acc) {
    if (!on) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// CTA-pair (cta_group::2) variants.  The leader CTA (cluster rank 0) issues the MMAs for both CTAs, so the loads of
// the partner complete their bytes on the LEADER's barrier (address mapped into the cluster window with mapa).
__device__ __forceinline__ uint32_t leader_addr(uint32_t local) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(0));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// the same arrival delivered to the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3) : "memory");
}
// M = 256 across the pair: rows 0-127 from the leader's query tile, 128-255 from the partner's; each CTA supplies
// 128 of the 256 database rows; each CTA's TMEM receives the accumulators of its own 128 query rows.
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO=64), LBO=1,
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10; cleared = fp16), both K-major, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC_BF16_BITS = (1u << 7) | (1u << 10);
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC_PAIR = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
// x -> (hi, lo) as raw 16-bit patterns: fp16 halves (two-product Gram) or bf16 halves (three-product Gram)
__device__ __forceinline__ void split16(float f, int fp16, unsigned short& h, unsigned short& l) {
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

// A-priori bound on |t_computed - t_exact| for a query of squared norm xq2 against a database whose largest squared
// norm is ymax2 (t = x.y - |y|^2/2), Dp = padded descriptor length.
//   bf16 x 3: dropped lo.lo / residual terms (3 * 2^-18 |x||y|), fp32 accumulation in the tensor core (K/16 roundings)
//             and the fp32 -|y|^2/2 term, generous factor on top;
//   fp16 x 2: the database enters by its fp16 rounding alone: |x.(y - hi_y)| <= 2^-12 |x||y| (2.44e-4), plus the same
//             accumulation and -|y|^2/2 terms, plus the absolute rounding of fp16 subnormals (2^-25 per component).
__device__ __forceinline__ float tc_err_bound(float xq2, float ymax2, int fp16, int Dp) {
    const float xy = sqrtf(xq2) * sqrtf(ymax2);
    if (!fp16) return 6.2e-5f * xy + 3.1e-5f * ymax2;
    return 2.8e-4f * xy + 3.1e-5f * ymax2 + 3.0e-8f * sqrtf((float)Dp) * (sqrtf(xq2) + sqrtf(ymax2));
}
constexpr float FP16_NORM2_LIMIT = 4.0e9f;      // |row|^2 below this keeps every component inside fp16's range (65504^2 = 4.29e9)

struct PrepParams {
    const float* d;          // [B,n_max,D]
    const int* cnt;          // [B] or null
    unsigned short* S;       // [B*n_max, 2*Dp] 16-bit float patterns
    int fp16;                // 1: fp16 halves, 0: bf16 halves
    float* c;                // [B, cs]   -|y|^2/2, -inf beyond the count
    float* norm2;            // [B*n_max]
    unsigned int* maxn;      // [B] max |row|^2 (float bits)
    int B, n_max, D, Dp, cs;
};

constexpr int PREP_ROWS = 2;          // rows per warp in flight (measured at cfg2: 8 rows 72 us, 4 rows 61 us, 2 rows 53 us, 1 row 51 us)

// returns |row|^2 (all lanes)
__device__ __forceinline__ float prep_finish_row(const PrepParams& p, int b, int row, float ss, int lane) {
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if (lane == 0) {
        p.c[(size_t)b * p.cs + row] = -0.5f * ss;
        p.norm2[(size_t)b * p.n_max + row] = ss;
    }
    return ss;
}

struct PrepPair {
    PrepParams side[2];       // d0 and d1 of the matcher, handled by one launch (blockIdx.z)
};

__global__ void __launch_bounds__(256) prep_kernel(PrepPair pp) {
    const PrepParams& p = pp.side[blockIdx.z];
    const int warp0 = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * 8;
    const int b = blockIdx.y;
    __shared__ float s_max[8];
    float wmax = 0.0f;                      // max |row|^2 over this warp's rows (one atomic per CTA, not per row)
    const int n = p.cnt ? p.cnt[b] : p.n_max;
    const bool vec = (p.D & 3) == 0 && (reinterpret_cast<uintptr_t>(p.d) & 15u) == 0;
    for (int row0 = warp0 * PREP_ROWS; row0 < p.cs; row0 += n_warps * PREP_ROWS)
    if (vec && p.D == 256 && row0 + PREP_ROWS <= n) {
        // D = 256: eight components per lane -- the whole row in one step, 16-byte stores of the hi and lo halves
        const float* x = p.d + ((size_t)b * p.n_max + row0) * 256 + 8 * lane;
        unsigned short* out = p.S + ((size_t)b * p.n_max + row0) * 512 + 8 * lane;
        float4 v[PREP_ROWS][2];
#pragma unroll
        for (int r = 0; r < PREP_ROWS; ++r) {
            v[r][0] = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * 256));
            v[r][1] = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * 256) + 1);
        }
#pragma unroll
        for (int r = 0; r < PREP_ROWS; ++r) {
            const float f[8] = {v[r][0].x, v[r][0].y, v[r][0].z, v[r][0].w, v[r][1].x, v[r][1].y, v[r][1].z, v[r][1].w};
            __align__(16) unsigned short h[8], l[8];
            // the summation order of the 4-component path (k = 4*lane, then k + 128): this lane holds components
            // 8*lane .. 8*lane+7, i.e. the partial sums of two "4-component lanes" -- any order is a valid |row|^2, the
            // error bound of the resolver covers the fp32 rounding of the sum
            float ss = 0.0f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                split16(f[e], p.fp16, h[e], l[e]);
                ss = fmaf(f[e], f[e], ss);
            }
            *reinterpret_cast<uint4*>(out + (size_t)r * 512) = *reinterpret_cast<const uint4*>(h);
            *reinterpret_cast<uint4*>(out + (size_t)r * 512 + 256) = *reinterpret_cast<const uint4*>(l);
            wmax = fmaxf(wmax, prep_finish_row(p, b, row0 + r, ss, lane));
        }
    } else if (vec && row0 + PREP_ROWS <= n) {
        // fast path: four valid rows, 4 components per lane and step
        const float* x = p.d + ((size_t)b * p.n_max + row0) * p.D;
        unsigned short* out = p.S + ((size_t)b * p.n_max + row0) * (2 * p.Dp);
        float ss[PREP_ROWS];
#pragma unroll
        for (int r = 0; r < PREP_ROWS; ++r) ss[r] = 0.f;
        for (int k = 4 * lane; k < p.Dp; k += 128) {
            float4 v[PREP_ROWS];
#pragma unroll
            for (int r = 0; r < PREP_ROWS; ++r)
                v[r] = k < p.D ? __ldg(reinterpret_cast<const float4*>(x + (size_t)r * p.D + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < PREP_ROWS; ++r) {
                const float f[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
                __align__(8) unsigned short h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    split16(f[e], p.fp16, h[e], l[e]);
                    ss[r] = fmaf(f[e], f[e], ss[r]);
                }
                unsigned short* o = out + (size_t)r * (2 * p.Dp);
                *reinterpret_cast<uint2*>(o + k) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(o + p.Dp + k) = *reinterpret_cast<const uint2*>(l);
            }
        }
#pragma unroll
        for (int r = 0; r < PREP_ROWS; ++r) wmax = fmaxf(wmax, prep_finish_row(p, b, row0 + r, ss[r], lane));
    } else for (int row = row0; row < row0 + PREP_ROWS && row < p.cs; ++row) {
        if (row >= p.n_max) {                       // padding of the c array up to a whole column tile
            if (lane == 0) p.c[(size_t)b * p.cs + row] = -CUDART_INF_F;
            continue;
        }
        unsigned short* out = p.S + ((size_t)b * p.n_max + row) * (2 * p.Dp);
        if (row >= n) {
            for (int k = lane; k < 2 * p.Dp; k += 32) out[k] = 0;
            if (lane == 0) { p.c[(size_t)b * p.cs + row] = -CUDART_INF_F; p.norm2[(size_t)b * p.n_max + row] = 0.0f; }
            continue;
        }
        const float* x = p.d + ((size_t)b * p.n_max + row) * p.D;
        float ss = 0.0f;
        for (int k = lane; k < p.Dp; k += 32) {
            const float v = k < p.D ? x[k] : 0.0f;
            unsigned short h, l;
            split16(v, p.fp16, h, l);
            out[k] = h;
            out[p.Dp + k] = l;
            ss = fmaf(v, v, ss);
        }
        wmax = fmaxf(wmax, prep_finish_row(p, b, row, ss, lane));
    }
    if (lane == 0) s_max[threadIdx.x >> 5] = wmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = s_max[0];
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_max[w]);
        if (mx > 0.0f) atomicMax(&p.maxn[b], __float_as_uint(mx));
    }
}

// Operand rows written by the fused sampler (kb_sample.cu: 16-bit halves and, per row and group of 8 components, the
// partial |row|^2): what is left of the preparation is one thread per row -- |row|^2 summed in a fixed order, c, the
// -inf padding and the largest |row|^2 of the batch item.
struct PrepFinish {
    const float* part;       // [B*n_max, Dp/8]
    const int* cnt;
    float* c;
    float* norm2;
    unsigned int* maxn;
    int n_max, cs, nparts;
};
struct PrepFinishPair {
    PrepFinish side[2];
};

__global__ void __launch_bounds__(256) prep_finish_kernel(PrepFinishPair pp) {
    const PrepFinish& p = pp.side[blockIdx.z];
    const int b = blockIdx.y, row = blockIdx.x * 256 + threadIdx.x;
    __shared__ float s_max[8];
    const int n = p.cnt ? p.cnt[b] : p.n_max;
    float ss = 0.0f;
    if (row < n) {
        const float4* q = reinterpret_cast<const float4*>(p.part + ((size_t)b * p.n_max + row) * p.nparts);
        for (int k = 0; k < p.nparts / 4; ++k) {
            const float4 v = __ldg(q + k);
            ss += v.x; ss += v.y; ss += v.z; ss += v.w;
        }
        for (int k = p.nparts / 4 * 4; k < p.nparts; ++k) ss += p.part[((size_t)b * p.n_max + row) * p.nparts + k];
        p.c[(size_t)b * p.cs + row] = -0.5f * ss;
        p.norm2[(size_t)b * p.n_max + row] = ss;
    } else if (row < p.cs) {
        p.c[(size_t)b * p.cs + row] = -CUDART_INF_F;
        if (row < p.n_max) p.norm2[(size_t)b * p.n_max + row] = 0.0f;
    }
    float mx = ss;
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_max[w]);
        if (mx > 0.0f) atomicMax(&p.maxn[b], __float_as_uint(mx));
    }
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
struct MainParams {
    const int* n0;           // [B] or null
    const int* n1;
    const float* c0;         // [B,cs0]
    const float* c1;         // [B,cs1]
    Top2* res0;              // [B*n_max]
    Top2* res1;              // [B*m_max]
    int B, n_max, m_max, cs0, cs1, KB, tiles0, tiles1, n_dirs, n_slots;
    int cl;                  // CTAs per cluster (1 or 2): a pair works on two adjacent query tiles and shares the database stream
    int align_slack;         // bytes the kernel may spend on aligning its tiles to 1024
    float2* gm;              // [B, gm_groups, cs1] (largest, second largest) column-side score of every 32-row group
    const unsigned int* maxn0;   // [B] max |x|^2, max |y|^2 (float bits) from prep_kernel: the offset of the column-side scores
    const unsigned int* maxn1;
    int gm_groups;           // 4 * tiles0
    int fp16;                // 1 = fp16 halves, two products (the database's lo half is never loaded); 0 = bf16, three
    int colside;             // 1 = one-pass cross-check (direction-0 items only, group maxima per column)
    long long* prof;         // timing experiments only (KB_KNOB_TC_DEBUG & 4): per CTA 8 cycle counters, see scripts/tc_pipeline_profile.py
    int dbg;                 // timing experiments only (KB_KNOB_TC_DEBUG): 1 = epilogue skips the fold, 2 = no MMAs issued
};

struct Item {
    int dir, b, q_row0, n_q, n_db, q_base, db_base;
};

// Work items are (direction, batch entry, group of p.cl adjacent query tiles); CTA `crank` of the cluster takes tile
// group * cl + crank.  The return value is uniform over the cluster: a CTA whose own tile lies beyond the query count
// still runs the item (its rows are never stored) because its partner needs its half of every database block.
template <int CL>
__device__ __forceinline__ bool decode_item(const MainParams& p, int item, int crank, Item& it) {
    const int t0 = (p.tiles0 + CL - 1) / CL, t1 = (p.tiles1 + CL - 1) / CL;
    const int per_dir0 = p.B * t0;
    int dir = 0, rem = item;
    if (item >= per_dir0) { dir = 1; rem = item - per_dir0; }
    const int tiles = dir ? t1 : t0;
    const int b = rem / tiles, group = rem - b * tiles;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    it.dir = dir; it.b = b;
    it.q_row0 = (group * CL + crank) * BM;
    it.n_q = dir ? m : n;
    it.n_db = dir ? n : m;
    it.q_base = dir ? b * p.m_max : b * p.n_max;
    it.db_base = dir ? b * p.n_max : b * p.m_max;
    return group * CL * BM < it.n_q && it.n_db > 0;
}

template <int CL, bool COLSIDE>
__global__ void __launch_bounds__(NT, 1) nn_top2_kernel(const __grid_constant__ CUtensorMap map0,
                                                        const __grid_constant__ CUtensorMap map1, MainParams p) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;                   // 1024B alignment for SWIZZLE_128B
    const int KB = p.KB;
    const uint32_t a_tiles = base;                                  // 2*KB tiles: hi blocks then lo blocks
    const uint32_t b_slots = base + (uint32_t)(2 * KB) * TILE_BYTES;
    const int NSLOT = p.n_slots;
    // CTA pair: every CTA keeps only ITS 128 of the 256 database rows of a block, so a slot is one 16 KB box
    constexpr uint32_t SLOTB = CL == 2 ? TILE_BYTES : SLOT_BYTES;
    const uint32_t bars = b_slots + NSLOT * SLOTB;
    if (base - raw > (uint32_t)p.align_slack) __trap();             // the host sized the allocation for this slack
    // barrier map (8 bytes each)
    // one (full, free) barrier pair per 64-wide K block of the query tile, so that the next item's query
    // blocks are reloaded while the last column tile of the current item is still being multiplied
    const uint32_t bar_a_full = bars, bar_a_free = bars + 32;
    const uint32_t bar_b_full = bars + 64, bar_b_empty = bars + 64 + 8 * MAX_SLOTS;
    const uint32_t bar_t_full = bars + 64 + 16 * MAX_SLOTS, bar_t_empty = bar_t_full + 16;
    const uint32_t tmem_slot = bar_t_empty + 16;
    float* cbuf = reinterpret_cast<float*>(smem_dyn + (tmem_slot + 16 - raw));     // [2][BN], see the epilogue
    // thr[row]: a lower bound on the row's third-best score over ALL column slices (the largest third-best any
    // slice has reported so far), shared by the four epilogue warps that own the row -- see the epilogue
    volatile float* thr = cbuf + 2 * BN;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    if (threadIdx.x == 0) {
        for (int kb = 0; kb < 4; ++kb) { mbar_init(bar_a_full + 8 * kb, 1); mbar_init(bar_a_free + 8 * kb, 1); }
        for (int s = 0; s < NSLOT; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
        // the leader's accumulator-empty barrier collects the epilogue warps of both CTAs of a pair
        for (int s = 0; s < 2; ++s) { mbar_init(bar_t_full + 8 * s, 1); mbar_init(bar_t_empty + 8 * s, CL * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CL == 2) {                    // one warp of each CTA of the pair
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                 // the partner's barriers exist before anything is signalled on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int g0 = (p.tiles0 + CL - 1) / CL, g1 = (p.tiles1 + CL - 1) / CL;
    const int n_items = p.n_dirs == 2 ? p.B * (g0 + g1) : p.B * g0;
    const int item0 = blockIdx.x / CL, item_step = gridDim.x / CL;

    // The producer and the MMA issuer run their loops with the WHOLE warp (uniform control flow keeps
    // addresses, descriptors and phases in uniform registers); only the issuing instructions themselves
    // are predicated on one lane.  A divergent single-lane loop costs ~150 cycles per tcgen05.mma in
    // address arithmetic alone, more than twice the 64 cycles the instruction occupies the tensor core.
    if (warp == 0) {
        // ================================ TMA producer ==========================================
        const bool issuer = lane == 0;
        const bool prof = (p.dbg & 4) != 0;
        long long w_afree = 0, w_bempty = 0;
        const long long t_start = clock64();
        uint32_t a_phase = 0, slot = 0, b_phase = 0;
        for (int item = item0; item < n_items; item += item_step) {
            Item it;
            if (!decode_item<CL>(p, item, crank, it)) continue;
            const CUtensorMap* qmap = it.dir ? &map1 : &map0;
            const CUtensorMap* dmap = it.dir ? &map0 : &map1;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait_prof(bar_a_free + 8 * kb, a_phase ^ 1, prof, w_afree);   // the previous item is done with this K block
                if (issuer) {
                    if constexpr (CL == 2) {
                        // both CTAs' query tiles are operands of the leader's MMAs: all four boxes land on its barrier
                        const uint32_t lb = leader_addr(bar_a_full + 8 * kb);
                        if (crank == 0) mbar_expect_tx(bar_a_full + 8 * kb, 4 * TILE_BYTES);
                        tma_load_2d_pair(a_tiles + kb * TILE_BYTES, qmap, kb * BK, it.q_base + it.q_row0, lb);
                        tma_load_2d_pair(a_tiles + (KB + kb) * TILE_BYTES, qmap, (KB + kb) * BK, it.q_base + it.q_row0, lb);
                    } else {
                        mbar_expect_tx(bar_a_full + 8 * kb, 2 * TILE_BYTES);
                        tma_load_2d(a_tiles + kb * TILE_BYTES, qmap, kb * BK, it.q_base + it.q_row0, bar_a_full + 8 * kb);
                        tma_load_2d(a_tiles + (KB + kb) * TILE_BYTES, qmap, (KB + kb) * BK, it.q_base + it.q_row0,
                                    bar_a_full + 8 * kb);
                    }
                }
            }
            a_phase ^= 1;
            const int n_ct = (it.n_db + BN - 1) / BN;
            for (int ct = 0; ct < n_ct; ++ct) {
                const int row = it.db_base + ct * BN;
                for (int kb = 0; kb < KB; ++kb) {
                    for (int part = 0; part < (p.fp16 ? 1 : 2); ++part) {          // hi block (then lo block: three-product Gram)
                        mbar_wait_prof(bar_b_empty + 8 * slot, b_phase ^ 1, prof, w_bempty);
                        if (issuer) {
                            if constexpr (CL == 2) {
                                // this CTA fetches its 128-row half of the block; both halves complete on the leader
                                if (crank == 0) mbar_expect_tx(bar_b_full + 8 * slot, 2 * TILE_BYTES);
                                tma_load_2d_pair(b_slots + slot * SLOTB, dmap, (part * KB + kb) * BK, row + crank * BM,
                                                 leader_addr(bar_b_full + 8 * slot));
                            } else {
                                mbar_expect_tx(bar_b_full + 8 * slot, SLOT_BYTES);
                                tma_load_2d(b_slots + slot * SLOT_BYTES, dmap, (part * KB + kb) * BK, row,
                                            bar_b_full + 8 * slot);
                                tma_load_2d(b_slots + slot * SLOT_BYTES + TILE_BYTES, dmap, (part * KB + kb) * BK, row + BM,
                                            bar_b_full + 8 * slot);
                            }
                        }
                        if (++slot == (uint32_t)NSLOT) { slot = 0; b_phase ^= 1; }
                    }
                }
            }
            __syncwarp();
        }
        if (prof && issuer) {
            long long* o = p.prof + (size_t)blockIdx.x * 8;
            o[0] = clock64() - t_start; o[1] = w_afree; o[2] = w_bempty;
        }
    } else if (warp == 1 && crank == 0) {
        // ================================ MMA issuer (the leader CTA of a pair issues for both) ==
        const bool issuer = lane == 0;
        const bool two = p.fp16 != 0;                               // fp16 halves: (hi_x + lo_x).hi_y only
        const uint32_t ID = (CL == 2 ? IDESC_PAIR : IDESC) & ~(two ? IDESC_BF16_BITS : 0u);
        auto mma = [ID](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
            if constexpr (CL == 2) tc_mma_pair(d, a, b, ID, acc); else tc_mma(d, a, b, ID, acc);
        };
        auto commit = [](uint32_t bar) {
            if constexpr (CL == 2) tc_commit_pair(bar); else tc_commit(bar);
        };
        const bool prof = (p.dbg & 4) != 0;
        long long w_tempty = 0, w_afull = 0, w_bfull = 0;
        const long long t_start = clock64();
        uint32_t a_phase = 0, slot = 0, b_phase = 0, acc_buf = 0, t_phase = 0;      // t_phase: one bit per buffer
        const uint64_t desc0 = umma_desc(0);                        // everything but the start address
        for (int item = item0; item < n_items; item += item_step) {
            Item it;
            if (!decode_item<CL>(p, item, crank, it)) continue;
            const int n_ct = (it.n_db + BN - 1) / BN;
            for (int ct = 0; ct < n_ct; ++ct) {
                mbar_wait_prof(bar_t_empty + 8 * acc_buf, ((t_phase >> acc_buf) & 1u) ^ 1u, prof, w_tempty);   // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc_buf * BN;
                for (int kb = 0; kb < KB; ++kb) {
                    if (ct == 0) {                                  // this item's query block has landed
                        mbar_wait_prof(bar_a_full + 8 * kb, a_phase, prof, w_afull);
                        tc_fence_after();
                    }
                    const uint64_t a_hi = desc0 | (uint64_t)((a_tiles + kb * TILE_BYTES) >> 4);
                    const uint64_t a_lo = desc0 | (uint64_t)((a_tiles + (KB + kb) * TILE_BYTES) >> 4);
                    // ---- database hi block: hi.hi and lo.hi
                    mbar_wait_prof(bar_b_full + 8 * slot, b_phase, prof, w_bfull);
                    tc_fence_after();
                    uint64_t bt = desc0 | (uint64_t)((b_slots + slot * SLOTB) >> 4);
                    if (issuer && !(p.dbg & 2)) {
                        mma(d_tmem, a_hi, bt, kb > 0 ? 1u : 0u);                  // 32 bytes (16 bf16) per K step: +2
                        mma(d_tmem, a_hi + 2, bt + 2, 1u);
                        mma(d_tmem, a_hi + 4, bt + 4, 1u);
                        mma(d_tmem, a_hi + 6, bt + 6, 1u);
                        mma(d_tmem, a_lo, bt, 1u);
                        mma(d_tmem, a_lo + 2, bt + 2, 1u);
                        mma(d_tmem, a_lo + 4, bt + 4, 1u);
                        mma(d_tmem, a_lo + 6, bt + 6, 1u);
                    }
                    if (issuer) {
                        commit(bar_b_empty + 8 * slot);
                        if (two && ct == n_ct - 1) commit(bar_a_free + 8 * kb);  // query block kb may be overwritten
                    }
                    __syncwarp();
                    if (++slot == (uint32_t)NSLOT) { slot = 0; b_phase ^= 1; }
                    if (!two) {
                        // ---- database lo block: hi.lo (three-product Gram only)
                        mbar_wait_prof(bar_b_full + 8 * slot, b_phase, prof, w_bfull);
                        tc_fence_after();
                        bt = desc0 | (uint64_t)((b_slots + slot * SLOTB) >> 4);
                        if (issuer && !(p.dbg & 2)) {
                            mma(d_tmem, a_hi, bt, 1u);
                            mma(d_tmem, a_hi + 2, bt + 2, 1u);
                            mma(d_tmem, a_hi + 4, bt + 4, 1u);
                            mma(d_tmem, a_hi + 6, bt + 6, 1u);
                        }
                        if (issuer) {
                            commit(bar_b_empty + 8 * slot);
                            if (ct == n_ct - 1) commit(bar_a_free + 8 * kb);     // query block kb may be overwritten
                        }
                        __syncwarp();
                        if (++slot == (uint32_t)NSLOT) { slot = 0; b_phase ^= 1; }
                    }
                }
                if (issuer) commit(bar_t_full + 8 * acc_buf);       // accumulator ready for the epilogue
                __syncwarp();
                t_phase ^= 1u << acc_buf;
                acc_buf ^= 1;
            }
            a_phase ^= 1;
        }
        if (prof && issuer) {
            long long* o = p.prof + (size_t)blockIdx.x * 8;
            o[3] = clock64() - t_start; o[4] = w_tempty; o[5] = w_afull; o[6] = w_bfull;
        }
    } else if (warp >= 4) {
        // ================================ epilogue ==============================================
        const int ew = warp - 4;
        const int quad = ew & 3;                                    // == warp % 4: TMEM lanes 32*quad .. 32*quad+31
        const int slice = ew >> 2;                                  // columns 64*slice .. 64*slice+63 of every tile
        const int row_in_tile = quad * 32 + lane;
        const int cidx = ew * 32 + lane;                            // c entry this thread stages (first 8 warps)
        constexpr int SW = BN / EPI_SLICES;                         // 64 columns per warp and tile
        uint32_t acc_buf = 0, t_phase = 0;
        const bool prof = (p.dbg & 4) != 0;
        long long w_tfull = 0;
        // -|y|^2/2 of a column tile is staged in one half of a 2 x 256-float shared buffer: the values of the NEXT
        // tile (of this item or of the next one) are fetched while the current tile is folded and published with one
        // named barrier per tile, so neither the load latency nor a second barrier sits on the per-tile path.
        auto c_of = [&](const Item& q) { return (q.dir ? p.c0 : p.c1) + (size_t)q.b * (q.dir ? p.cs0 : p.cs1); };
        auto next_valid = [&](int from, Item& q) {
            for (int k = from; k < n_items; k += item_step)
                if (decode_item<CL>(p, k, crank, q)) return true;
            return false;
        };
        int cpar = 0;
        {
            Item first;
            if (next_valid(item0, first) && cidx < BN) cbuf[cidx] = __ldg(c_of(first) + cidx);
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        }
        for (int item = item0; item < n_items; item += item_step) {
            Item it;
            if (!decode_item<CL>(p, item, crank, it)) continue;
            const float* cvec = c_of(it);
            // running top-3 scores and top-2 indices of this row over this warp's column slice; the slices
            // are merged by the resolver.  (Which of two exactly tied scores ranks first is irrelevant
            // here: near-ties are settled exactly in the resolver.)
            float b1 = -CUDART_INF_F, b2 = -CUDART_INF_F, b3 = -CUDART_INF_F;
            int i1 = 0, i2 = 0;
            float u1 = 0.0f, u2 = 0.0f;
            // One-pass cross-check: u(i,j) = x_i.y_j - |x_i|^2/2 + offset ranks the ROWS of column j like their distances
            // do.  The offset makes every valid score positive, so its float bits order like signed integers and
            // redux.sync.max.s32 reduces a column over the warp's 32 rows; rows beyond the count carry -inf (negative as
            // an integer, never a maximum).
            constexpr bool colside = COLSIDE;
            float kx = 0.0f;
            float2* gmrow = nullptr;
            if constexpr (colside) {
                const float mx0 = __uint_as_float(p.maxn0[it.b]), mx1 = __uint_as_float(p.maxn1[it.b]);
                const float off = (0.5f * mx0 + sqrtf(mx0 * mx1)) * 1.01f + 1e-6f;
                kx = off + __ldg(p.c0 + (size_t)it.b * p.cs0 + it.q_row0 + row_in_tile);      // -inf beyond the count
                gmrow = p.gm + ((size_t)it.b * p.gm_groups + (it.q_row0 >> 5) + quad) * p.cs1;
            }
            const int n_ct = (it.n_db + BN - 1) / BN;
            // Each warp sees only a quarter of the columns, so on its own its third-best -- the threshold below
            // which columns are skipped -- rises four times more slowly than the row's.  After every tile a warp
            // publishes its third-best if it beats the shared bound and adopts the bound before the next tile: a
            // score below ANY slice's third-best cannot be among the row's best three, and the min/max/select
            // sequence of an insertion (ALU pipe, half rate) is what this epilogue spends its time on.  The bound is
            // reset by slice 0 while the others are in the item's first tile (which they fold without it); plain
            // racing stores are fine, every value ever stored is a valid bound.
            if (slice == 0) thr[row_in_tile] = -CUDART_INF_F;
            float th = -CUDART_INF_F;
            // TMEM -> register bandwidth (~56 B/clk/SM measured) is what bounds this epilogue, so every warp
            // keeps a tcgen05.ld in flight while it folds the previous 16 columns (two 16-register buffers;
            // the first chunk of the next tile is requested before the last chunk of this one is folded).
            auto fold = [&](const uint32_t (&v)[16], int ct, int col0) {
                if (p.dbg & 1) { b1 = fmaxf(b1, __uint_as_float(v[0] ^ v[7] ^ v[15])); return; }   // timing experiment: no fold
                const float4* c4 = reinterpret_cast<const float4*>(cbuf + cpar * BN + col0);
                const int j0 = ct * BN + col0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t4 = c4[q];
                    float t[4];
                    t[0] = __uint_as_float(v[4 * q + 0]) + t4.x;              // x.y - |y|^2/2 (-inf beyond the count)
                    t[1] = __uint_as_float(v[4 * q + 1]) + t4.y;
                    t[2] = __uint_as_float(v[4 * q + 2]) + t4.z;
                    t[3] = __uint_as_float(v[4 * q + 3]) + t4.w;
                    float uu[4] = {0.f, 0.f, 0.f, 0.f};
                    if constexpr (colside) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) uu[e] = __uint_as_float(v[4 * q + e]) + kx;
                    }
                    const float m4 = fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3]));
                    // groups of four columns that cannot enter any lane's top-3 are skipped with one warp vote
                    if (__any_sync(0xffffffffu, m4 > b3)) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float te = t[e];
                            const int j = j0 + 4 * q + e;
                            const bool g1 = te > b1, g2 = te > b2;
                            b3 = fmaxf(b3, fminf(te, b2));
                            b2 = fmaxf(b2, fminf(te, b1));
                            b1 = fmaxf(b1, te);
                            i2 = g1 ? i1 : (g2 ? j : i2);
                            i1 = g1 ? j : i1;
                            if constexpr (colside) {
                                u2 = g1 ? u1 : (g2 ? uu[e] : u2);
                                u1 = g1 ? uu[e] : u1;
                            }
                        }
                    }
                    if constexpr (colside) {
                        // every column: (largest, second largest) score over this warp's 32 rows
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int key = __float_as_int(uu[e]);
                            const int m1 = __reduce_max_sync(0xffffffffu, key);
                            const bool top = key == m1;
                            const unsigned eq = __ballot_sync(0xffffffffu, top);
                            int m2 = __reduce_max_sync(0xffffffffu, top ? (int)0x80000000 : key);
                            if (eq & (eq - 1u)) m2 = m1;                          // the largest occurs twice
                            KB_ASSERT(j0 + 4 * q + e < p.cs1 && (it.q_row0 >> 5) + quad < p.gm_groups);
                            if (lane == 0) gmrow[j0 + 4 * q + e] = make_float2(__int_as_float(m1), __int_as_float(m2));
                        }
                    }
                }
            };
            const int col_base = slice * SW;
            uint32_t ra[16] = {}, rb[16] = {};
            const bool noload = (p.dbg & 8) != 0;      // timing experiment: the epilogue leaves TMEM alone
            // prologue: first tile
            mbar_wait_prof(bar_t_full + 8 * acc_buf, (t_phase >> acc_buf) & 1u, prof, w_tfull);
            t_phase ^= 1u << acc_buf;
            tc_fence_after();
            uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_buf * BN + col_base;
            if (!noload) tmem_ld16(taddr, ra);
            tmem_ld_wait();
            for (int ct = 0; ct < n_ct; ++ct) {
                // the tile after this one (uniform over the 16 warps): its c values travel while this tile is folded
                const float* nc = nullptr;
                if (ct + 1 < n_ct) nc = cvec + (ct + 1) * BN;
                else { Item nx; if (next_valid(item + item_step, nx)) nc = c_of(nx); }
                const float cvn = (nc != nullptr && cidx < BN) ? __ldg(nc + cidx) : 0.0f;
                if (ct > 0) { th = thr[row_in_tile]; b3 = fmaxf(b3, th); }
                if (!noload) tmem_ld16(taddr + 16, rb);
                fold(ra, ct, col_base);
                tmem_ld_wait();
                if (!noload) tmem_ld16(taddr + 32, ra);
                fold(rb, ct, col_base + 16);
                tmem_ld_wait();
                if (!noload) tmem_ld16(taddr + 48, rb);
                fold(ra, ct, col_base + 32);
                tmem_ld_wait();
                // the slice has left TMEM: hand the accumulator back to the MMA issuer
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CL == 2) mbar_arrive_cluster(leader_addr(bar_t_empty + 8 * acc_buf));
                    else mbar_arrive(bar_t_empty + 8 * acc_buf);
                }
                acc_buf ^= 1;
                fold(rb, ct, col_base + 48);
                const bool more = ct + 1 < n_ct;
                if (more) {
                    mbar_wait_prof(bar_t_full + 8 * acc_buf, (t_phase >> acc_buf) & 1u, prof, w_tfull);
                    t_phase ^= 1u << acc_buf;
                    tc_fence_after();
                    taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_buf * BN + col_base;
                    if (!noload) tmem_ld16(taddr, ra);
                }
                if (more && b3 > th) thr[row_in_tile] = b3;
                if (nc != nullptr) {
                    // every warp has passed the previous barrier, i.e. finished reading the other half one tile ago
                    if (cidx < BN) cbuf[(cpar ^ 1) * BN + cidx] = cvn;
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                    cpar ^= 1;
                }
                if (more) tmem_ld_wait();
            }
            const int qi = it.q_row0 + row_in_tile;
            if (qi < it.n_q) {
                KB_ASSERT(i1 >= 0 && i1 < it.n_db + BN && it.q_base + qi < (it.dir ? p.B * p.m_max : p.B * p.n_max));
                Top2 o;
                o.best = b1; o.second = b2; o.third = fminf(b3, b2);                    // b3 may hold the adopted bound
                o.u1 = u1; o.idx = i1; o.idx2 = i2; o.u2 = u2; o.pad = 0;
                (it.dir ? p.res1 : p.res0)[((size_t)it.q_base + qi) * EPI_SLICES + slice] = o;
            }
        }
        if (prof && ew == 0 && lane == 0) p.prof[(size_t)blockIdx.x * 8 + 7] = w_tfull;
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                 // no CTA leaves while its partner can still write to it
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CL == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// certification / exact resolution (one warp per query)
// ------------------------------------------------------------------------------------------------
struct RescanPart;
struct ResolveParams {
    const float* d0;         // [B,n_max,D]
    const float* d1;         // [B,m_max,D]
    const int* n0;
    const int* n1;
    const Top2* res0;
    const Top2* res1;
    const float* norm2_0;
    const float* norm2_1;
    const unsigned int* maxn0;
    const unsigned int* maxn1;
    int* nn0;                // [B*n_max]
    int* nn1;                // [B*m_max]
    int* n_exact;            // [1] number of rows queued for the exact rescan
    int* n_pair;             // [1] rows settled by the two-candidate exact check (statistics)
    float* tsel;             // [B*n_max] tensor-core score t = x.y - |y|^2/2 of the direction-0 winner (NaN: unknown)
    float* usel;             // [B*n_max] its column-side score u (one-pass cross-check; NaN: unknown)
    const float2* gm;        // [B, gm_groups, cs1] group maxima of the column-side scores (search kernel)
    int4* colinfo;           // [B*m_max] per column: (largest u, largest of everything else, group of the largest, 0) as int bits
    int gm_groups, cs1, colside;
    int fp16, Dp;            // operand split of the search (error bound), padded descriptor length
    int2* list;              // [list_cap] queued rows: (dir | b << 1, query row)
    int list_cap;
    struct RescanPart* parts;   // [list_cap, RESCAN_SPLIT] partial minima of the split rescan
    int* tickets;            // [list_cap] arrival counters of the splits (zeroed per call)
    int B, n_max, m_max, D, n_dirs;
};

__device__ __forceinline__ double warp_dist2(const float* x, const float* y, int D, int lane) {
    double acc = 0.0;
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0) {
        for (int k = 4 * lane; k < D; k += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x + k)), b = __ldg(reinterpret_cast<const float4*>(y + k));
            double d = (double)a.x - (double)b.x; acc = fma(d, d, acc);
            d = (double)a.y - (double)b.y; acc = fma(d, d, acc);
            d = (double)a.z - (double)b.z; acc = fma(d, d, acc);
            d = (double)a.w - (double)b.w; acc = fma(d, d, acc);
        }
    } else {
        for (int k = lane; k < D; k += 32) {
            const double d = (double)x[k] - (double)y[k];
            acc = fma(d, d, acc);
        }
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    return acc;
}

// One thread per query row.  With e the a-priori error bound of the split product:
//   best - second > 2e : the tensor-core argmax is certified;
//   best - third  > 2e : the exact winner is one of the two best -- their float64 distances are
//                        evaluated on the spot by the whole warp (smaller wins, ties to the smaller index);
//   otherwise          : the row is queued for the exact rescan (three or more near-ties: duplicates).
__global__ void __launch_bounds__(256) resolve_kernel(ResolveParams p) {
    const int q = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
    const int b = blockIdx.y, dir = blockIdx.z;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    if (p.colside && dir == 1) {
        // One-pass cross-check: column q of the database.  Over the 32-row groups of the query rows (every group of a
        // row tile that the search kernel ran is written; rows beyond the count score -inf): the largest group maximum,
        // its group, and the largest of EVERYTHING ELSE = max(second largest group maximum, second largest inside the
        // winning group).  Scores are positive floats (or -inf), so their bits compare as signed integers.
        if (q >= m || n <= 0) return;
        const int n_groups = 4 * ((n + BM - 1) / BM);
        const float2* g = p.gm + (size_t)b * p.gm_groups * p.cs1 + q;
        int v1 = (int)0x80000000, v2 = (int)0x80000000, s1 = (int)0x80000000, g1 = 0;
        for (int k = 0; k < n_groups; ++k) {
            const float2 e = __ldg(g + (size_t)k * p.cs1);
            const int a = __float_as_int(e.x), a2 = __float_as_int(e.y);
            if (a > v1) { v2 = v1; v1 = a; g1 = k; s1 = a2; }
            else if (a > v2) v2 = a;                               // a == v1 lands here: a tie between groups
        }
        p.colinfo[(size_t)b * p.m_max + q] = make_int4(v1, v2 > s1 ? v2 : s1, g1, 0);
        return;
    }
    const int nq = dir ? m : n, ndb = dir ? n : m;
    if (ndb <= 0) return;
    const int q_stride = dir ? p.m_max : p.n_max, db_stride = dir ? p.n_max : p.m_max;
    int mode = 0;                        // 0 nothing to do, 1 certified, 2 two candidates, 3 rescan
    int j1 = 0, j2 = 0;
    float tbest = 0.0f, tsecond = 0.0f, ubest = 0.0f, usecond = 0.0f;
    if (q < nq) {
        Top2 r = (dir ? p.res1 : p.res0)[((size_t)b * q_stride + q) * EPI_SLICES];
#pragma unroll
        for (int sl = 1; sl < EPI_SLICES; ++sl) {                     // merge the column slices of the epilogue
            const Top2 o = (dir ? p.res1 : p.res0)[((size_t)b * q_stride + q) * EPI_SLICES + sl];
            const float cand[3] = {o.best, o.second, o.third};
            const int cix[3] = {o.idx, o.idx2, 0};
            const float cu[3] = {o.u1, o.u2, 0.0f};
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const float te = cand[e];
                const bool g1 = te > r.best, g2 = te > r.second;
                r.third = fmaxf(r.third, fminf(te, r.second));
                r.second = fmaxf(r.second, fminf(te, r.best));
                r.best = fmaxf(r.best, te);
                r.idx2 = g1 ? r.idx : (g2 ? cix[e] : r.idx2);
                r.idx = g1 ? cix[e] : r.idx;
                r.u2 = g1 ? r.u1 : (g2 ? cu[e] : r.u2);
                r.u1 = g1 ? cu[e] : r.u1;
            }
        }
        const float nq2 = (dir ? p.norm2_1 : p.norm2_0)[(size_t)b * q_stride + q];
        const float dbmax2 = __uint_as_float((dir ? p.maxn0 : p.maxn1)[b]);
        const float e = tc_err_bound(nq2, dbmax2, p.fp16, p.Dp);      // bound on |t_computed - t_exact|
        j1 = r.idx; j2 = r.idx2;
        tbest = r.best; tsecond = r.second;
        ubest = r.u1; usecond = r.u2;
        const bool ok1 = j1 >= 0 && j1 < ndb, ok2 = j2 >= 0 && j2 < ndb && j2 != j1;
        // fp16 operands need every component inside fp16's range: guaranteed below a squared-norm limit, otherwise
        // (absurdly large descriptors) every row of the pair is resolved by the exact rescan
        const bool in_range = !p.fp16 || (__uint_as_float(p.maxn0[b]) < FP16_NORM2_LIMIT && __uint_as_float(p.maxn1[b]) < FP16_NORM2_LIMIT);
        if (in_range && ok1 && (r.best - r.second) > 2.0f * e) mode = 1;
        else if (in_range && ok1 && ok2 && (r.best - r.third) > 2.0f * e) mode = 2;
        else mode = 3;
    }
    int* nn = dir ? p.nn1 : p.nn0;
    if (mode == 1) nn[(size_t)b * q_stride + q] = j1;
    if (dir == 0 && mode != 0) {
        p.tsel[(size_t)b * q_stride + q] = mode == 1 ? tbest : CUDART_NAN_F;
        if (p.colside) p.usel[(size_t)b * q_stride + q] = mode == 1 ? ubest : CUDART_NAN_F;
    }
    if (mode == 3) {
        const int slot = atomicAdd(p.n_exact, 1);
        if (slot < p.list_cap) p.list[slot] = make_int2(dir | (b << 1), q);
    }
    __syncwarp();                       // the NaN placeholders above are ordered before lane 0's final values
    unsigned pend = __ballot_sync(0xffffffffu, mode == 2);
    const float* Qb = (dir ? p.d1 : p.d0) + (size_t)b * q_stride * p.D;
    const float* DBs = (dir ? p.d0 : p.d1) + (size_t)b * db_stride * p.D;
    while (pend) {
        const int src = __ffs(pend) - 1;
        pend &= pend - 1;
        const int qq = __shfl_sync(0xffffffffu, q, src);
        const int a1 = __shfl_sync(0xffffffffu, j1, src), a2 = __shfl_sync(0xffffffffu, j2, src);
        const double d1 = warp_dist2(Qb + (size_t)qq * p.D, DBs + (size_t)a1 * p.D, p.D, lane);
        const double d2 = warp_dist2(Qb + (size_t)qq * p.D, DBs + (size_t)a2 * p.D, p.D, lane);
        const float t1 = __shfl_sync(0xffffffffu, tbest, src), t2 = __shfl_sync(0xffffffffu, tsecond, src);
        const float w1 = __shfl_sync(0xffffffffu, ubest, src), w2 = __shfl_sync(0xffffffffu, usecond, src);
        if (lane == 0) {
            const bool first = d1 < d2 || (d1 == d2 && a1 < a2);
            nn[(size_t)b * q_stride + qq] = first ? a1 : a2;
            if (dir == 0) {
                p.tsel[(size_t)b * q_stride + qq] = first ? t1 : t2;
                if (p.colside) p.usel[(size_t)b * q_stride + qq] = first ? w1 : w2;
            }
            if (p.n_pair) atomicAdd(p.n_pair, 1);
        }
    }
}

// Exact float64 resolution of the queued rows (best/second closer than the error bound of the split
// product, e.g. duplicated descriptors): one CTA per row, lanes over components, eight candidates in
// flight per warp (16 independent 16-byte loads per lane at D=256).  First of ties wins, as np.argmin.
constexpr int RESCAN_WARPS = 8;
constexpr int RESCAN_SPLIT = 8;       // a queued row is rare but must not become a long tail: one SM reads the
                                      // database at ~60 GB/s, so every row is scanned by 8 CTAs (last one to finish merges)

template <bool VEC>
__device__ __forceinline__ void rescan_row(const float* xs, const float* DBs, int D, int c_begin, int ndb, int warp, int lane,
                                           double& bd, int& bj) {
    constexpr int G = 8;
    for (int c0 = c_begin + warp * G; c0 < ndb; c0 += RESCAN_WARPS * G) {
        double acc[G];
#pragma unroll
        for (int u = 0; u < G; ++u) acc[u] = 0.0;
        if (VEC) {
            for (int k = 4 * lane; k < D; k += 128) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + k);
                float4 yv[G];
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const int c = min(c0 + u, ndb - 1);
                    yv[u] = __ldg(reinterpret_cast<const float4*>(DBs + (size_t)c * D + k));
                }
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    double d = (double)xv.x - (double)yv[u].x; acc[u] = fma(d, d, acc[u]);
                    d = (double)xv.y - (double)yv[u].y; acc[u] = fma(d, d, acc[u]);
                    d = (double)xv.z - (double)yv[u].z; acc[u] = fma(d, d, acc[u]);
                    d = (double)xv.w - (double)yv[u].w; acc[u] = fma(d, d, acc[u]);
                }
            }
        } else {
            for (int k = lane; k < D; k += 32) {
                const double xv = (double)xs[k];
#pragma unroll
                for (int u = 0; u < G; ++u) {
                    const int c = min(c0 + u, ndb - 1);
                    const double d = xv - (double)__ldg(DBs + (size_t)c * D + k);
                    acc[u] = fma(d, d, acc[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
            double a = acc[u];
            for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
            if (c0 + u < ndb && a < bd) { bd = a; bj = c0 + u; }      // ascending c within the warp: strict <
        }
    }
}

struct RescanPart {
    double d2;
    int j, pad;
};

__global__ void __launch_bounds__(RESCAN_WARPS * 32) rescan_kernel(ResolveParams p) {
    extern __shared__ __align__(16) float xs[];
    __shared__ double s_d[RESCAN_WARPS];
    __shared__ int s_j[RESCAN_WARPS];
    __shared__ int s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    int total = *p.n_exact;
    if (total > p.list_cap) total = p.list_cap;
    const bool vec = (p.D & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.d0) | reinterpret_cast<uintptr_t>(p.d1)) & 15u) == 0;
    for (int e = blockIdx.y; e < total; e += gridDim.y) {
        const int2 ent = p.list[e];
        const int dir = ent.x & 1, b = ent.x >> 1, q = ent.y;
        const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
        const int ndb = dir ? n : m;
        const int q_stride = dir ? p.m_max : p.n_max, db_stride = dir ? p.n_max : p.m_max;
        const float* Q = (dir ? p.d1 : p.d0) + ((size_t)b * q_stride + q) * p.D;
        const float* DBs = (dir ? p.d0 : p.d1) + (size_t)b * db_stride * p.D;
        const int chunk = ((ndb + RESCAN_SPLIT - 1) / RESCAN_SPLIT + 7) & ~7;
        const int c_begin = split * chunk, c_end = min(ndb, c_begin + chunk);
        __syncthreads();
        for (int k = threadIdx.x; k < p.D; k += RESCAN_WARPS * 32) xs[k] = Q[k];
        __syncthreads();
        double bd = CUDART_INF;
        int bj = 0x7fffffff;
        if (vec) rescan_row<true>(xs, DBs, p.D, c_begin, c_end, warp, lane, bd, bj);
        else rescan_row<false>(xs, DBs, p.D, c_begin, c_end, warp, lane, bd, bj);
        if (lane == 0) { s_d[warp] = bd; s_j[warp] = bj; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double d2 = s_d[0];
            int j = s_j[0];
            for (int w = 1; w < RESCAN_WARPS; ++w)
                if (s_d[w] < d2 || (s_d[w] == d2 && s_j[w] < j)) { d2 = s_d[w]; j = s_j[w]; }
            RescanPart part;
            part.d2 = d2; part.j = j; part.pad = 0;
            p.parts[(size_t)e * RESCAN_SPLIT + split] = part;
            __threadfence();
            s_last = atomicAdd(&p.tickets[e], 1) == RESCAN_SPLIT - 1;
            if (s_last) {                                              // every split of this row has been published
                __threadfence();
                const volatile RescanPart* pp = p.parts + (size_t)e * RESCAN_SPLIT;
                double best = pp[0].d2;
                int bjj = pp[0].j;
                for (int t = 1; t < RESCAN_SPLIT; ++t) {
                    const double dt = pp[t].d2;
                    const int jt = pp[t].j;
                    if (dt < best || (dt == best && jt < bjj)) { best = dt; bjj = jt; }
                }
                (dir ? p.nn1 : p.nn0)[(size_t)b * q_stride + q] = bjj;
                p.tickets[e] = 0;                                          // ready for the next rescan launch of this call
            }
        }
    }
}

// One warp per four rows of d0: mutual check, float64 distance of the surviving pair, strict < max_distance.
struct GateParams {
    const float* d0;
    const float* d1;
    const int* n0;
    const int* n1;
    const int* nn0;
    const int* nn1;
    int* keep_j;             // [B*n_max] matched column, -1 (no match) or -2 (waiting for the exact rescan of its column)
    double* dist_i;          // [B*n_max]
    const float* tsel;       // [B*n_max] tensor-core score of the winner (NaN: unknown)
    const float* usel;       // [B*n_max] its column-side score (NaN: unknown)
    const int4* colinfo;     // [B*m_max] see resolve_kernel
    const float* norm2_0;    // [B*n_max] |x|^2
    const unsigned int* maxn0;   // [B] max |x|^2 (float bits)
    const unsigned int* maxn1;   // [B] max |y|^2 (float bits)
    int2* list2;             // columns queued for the exact rescan: (1 | b << 1, column)
    int* n_list2;
    int list2_cap;
    int n_max, m_max, D, cross_check, want_dist, colside, fp16, Dp;
    double max_distance;
};

// pass 0: every row.  pass 1 (after the rescan of the queued columns): the rows left at -2.
// (3 CTAs per SM: the kernel is a chain of dependent loads, its time is the number of waves)
__global__ void __launch_bounds__(256, 3) gate_kernel(GateParams p, int pass) {
    constexpr int G = 4;                  // rows per warp, their loads issued together
    const int i0 = ((blockIdx.x * 256 + threadIdx.x) >> 5) * G, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    if (i0 >= n || m <= 0) return;
    if (pass == 1 && *p.n_list2 == 0) return;
    int jj[G];
    bool live[G];
    const float xmax2 = __uint_as_float(p.maxn0[b]), ymax2 = __uint_as_float(p.maxn1[b]);
    // (the loads of the four rows are issued together: first every row's column, then what depends on it)
    const bool use_nn1 = pass == 1 || (p.cross_check && !p.colside);
    int back[G], kept[G];
#pragma unroll
    for (int g = 0; g < G; ++g) jj[g] = i0 + g < n ? p.nn0[(size_t)b * p.n_max + i0 + g] : 0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        back[g] = (use_nn1 && i0 + g < n) ? p.nn1[(size_t)b * p.m_max + jj[g]] : -1;
        kept[g] = (pass == 1 && i0 + g < n) ? p.keep_j[(size_t)b * p.n_max + i0 + g] : 0;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int i = i0 + g;
        const size_t row = (size_t)b * p.n_max + i;
        live[g] = false;
        if (i >= n) continue;
        int st;                             // 0 no match, 1 mutual, 2 too close to call, 3 leave the row alone
        if (pass == 1) {
            st = kept[g] == -2 ? (back[g] == i ? 1 : 0) : 3;
        } else if (!p.cross_check) {
            st = 1;
        } else if (!p.colside) {
            st = back[g] == i ? 1 : 0;
        } else {
            // Is row i the best row of column j?  u(i,j) against the group maxima of column j.  |u - exact| <= e for
            // every row, so a lead of more than 2 e settles it either way; anything closer is rescanned exactly.
            const float us = p.usel[row];
            const int4 ci = p.colinfo[(size_t)b * p.m_max + jj[g]];
            const float off = (0.5f * xmax2 + sqrtf(xmax2 * ymax2)) * 1.01f + 1e-6f;
            const float e = tc_err_bound(ymax2, xmax2, p.fp16, p.Dp) + 6.0e-7f * off;   // (the rows are the "database" of a column)
            st = 2;
            if (us == us) {
                const float v1 = __int_as_float(ci.x), rest = __int_as_float(ci.y);
                if (ci.z == (i >> 5) && __float_as_int(us) == ci.x) { if (us - rest > 2.0f * e) st = 1; }
                else if (v1 - us > 2.0f * e) st = 0;
            }
        }
        if (st == 2 && lane == 0) {
            p.keep_j[row] = -2;
            const int slot = atomicAdd(p.n_list2, 1);
            if (slot < p.list2_cap) p.list2[slot] = make_int2(1 | (b << 1), jj[g]);
        }
        if (st == 0 && lane == 0) p.keep_j[row] = -1;
        live[g] = st == 1;
    }
    if (!p.want_dist) {
        // The caller only wants the pairs: |x - y|^2 = |x|^2 - 2 t with t = x.y - |y|^2/2 from the tensor cores,
        // known to within `err`; the float64 distance is evaluated only where that cannot decide d < max_distance.
        const double md2 = p.max_distance * p.max_distance;
        bool any_exact = false;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (!live[g]) continue;
            const size_t row = (size_t)b * p.n_max + i0 + g;
            const float t = p.tsel[row], x2 = p.norm2_0[row];
            const float e = tc_err_bound(x2, ymax2, p.fp16, p.Dp);
            const double d2 = (double)x2 - 2.0 * (double)t, err = 2.0 * (double)e + 4e-6 * (double)x2;
            if (d2 + err < md2) {                       // certainly inside the gate
                if (lane == 0) p.keep_j[row] = jj[g];
                live[g] = false;
            } else if (d2 - err >= md2) {               // certainly outside
                live[g] = false;
                if (lane == 0) p.keep_j[row] = -1;
            } else {
                any_exact = true;                       // NaN or too close to call: exact evaluation below
            }
        }
        if (!any_exact) return;
    }
    double acc[G] = {0.0, 0.0, 0.0, 0.0};
    const bool vec = (p.D & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.d0) | reinterpret_cast<uintptr_t>(p.d1)) & 15u) == 0;
    if (vec) {
        for (int k = 4 * lane; k < p.D; k += 128) {
            float4 xv[G], yv[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                xv[g] = yv[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live[g]) {
                    xv[g] = __ldg(reinterpret_cast<const float4*>(p.d0 + ((size_t)b * p.n_max + i0 + g) * p.D + k));
                    yv[g] = __ldg(reinterpret_cast<const float4*>(p.d1 + ((size_t)b * p.m_max + jj[g]) * p.D + k));
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                double d = (double)xv[g].x - (double)yv[g].x; acc[g] = fma(d, d, acc[g]);
                d = (double)xv[g].y - (double)yv[g].y; acc[g] = fma(d, d, acc[g]);
                d = (double)xv[g].z - (double)yv[g].z; acc[g] = fma(d, d, acc[g]);
                d = (double)xv[g].w - (double)yv[g].w; acc[g] = fma(d, d, acc[g]);
            }
        }
    } else for (int k = lane; k < p.D; k += 32) {
        float xv[G], yv[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            xv[g] = live[g] ? __ldg(p.d0 + ((size_t)b * p.n_max + i0 + g) * p.D + k) : 0.0f;
            yv[g] = live[g] ? __ldg(p.d1 + ((size_t)b * p.m_max + jj[g]) * p.D + k) : 0.0f;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const double d = (double)xv[g] - (double)yv[g];
            acc[g] = fma(d, d, acc[g]);
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        double a = acc[g];
        for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
        if (lane == 0 && live[g]) {
            const double d = sqrt(a);
            p.keep_j[(size_t)b * p.n_max + i0 + g] = d < p.max_distance ? jj[g] : -1;
            p.dist_i[(size_t)b * p.n_max + i0 + g] = d;
        }
    }
}

struct PairsParams {
    const int* n0;
    const int* n1;
    const int* keep_j;
    const double* dist_i;
    int* pairs;
    double* dist;
    int* count;
    int n_max, m_max;
};

__global__ void __launch_bounds__(1024) pairs_kernel(PairsParams p) {
    __shared__ int s_scan[33];
    const int b = blockIdx.x;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    if (n <= 0 || m <= 0) {
        if (threadIdx.x == 0) p.count[b] = 0;
        return;
    }
    int n_out = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int j = i < n ? p.keep_j[(size_t)b * p.n_max + i] : -1;
        const bool keep = j >= 0;
        int tot;
        const int off = n_out + kb::block_exclusive_scan(keep ? 1 : 0, s_scan, &tot);
        if (keep) {
            p.pairs[((size_t)b * p.n_max + off) * 2 + 0] = i;
            p.pairs[((size_t)b * p.n_max + off) * 2 + 1] = j;
            if (p.dist) p.dist[(size_t)b * p.n_max + off] = p.dist_i[(size_t)b * p.n_max + i];
        }
        n_out += tot;
    }
    if (threadIdx.x == 0) p.count[b] = n_out;
}

}  // namespace kbtc

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)ptr;
    }
    return fn;
}

static int make_map(CUtensorMap* map, void* base, uint64_t rows, uint64_t ks, int fp16) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return KB_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {ks, rows};
    cuuint64_t strides[1] = {ks * 2};
    cuuint32_t box[2] = {(cuuint32_t)kbtc::BK, (cuuint32_t)kbtc::BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? KB_OK : KB_ERR_UNSUPPORTED;
}

struct TcLayout {
    int Dp, KB, cs0, cs1, tiles0, tiles1;
    size_t bytes;
};

static TcLayout tc_layout(int B, int n_max, int m_max, int D) {
    TcLayout L;
    L.Dp = (D + 63) / 64 * 64;
    L.KB = L.Dp / 64;
    L.tiles0 = (n_max + kbtc::BM - 1) / kbtc::BM;
    L.tiles1 = (m_max + kbtc::BM - 1) / kbtc::BM;
    L.cs0 = (n_max + kbtc::BN - 1) / kbtc::BN * kbtc::BN;     // -|y|^2/2 padded to whole column tiles
    L.cs1 = (m_max + kbtc::BN - 1) / kbtc::BN * kbtc::BN;
    size_t n = 0;
    auto add = [&](size_t bytes) { n += kb_align_up(bytes, 256); };
    add((size_t)B * n_max * 2 * L.Dp * 2);      // S0
    add((size_t)B * m_max * 2 * L.Dp * 2);      // S1
    add((size_t)B * L.cs0 * 4);                 // c0
    add((size_t)B * L.cs1 * 4);                 // c1
    add((size_t)B * n_max * 4);                 // norm2_0
    add((size_t)B * m_max * 4);                 // norm2_1
    add((size_t)B * 4);                         // maxn0
    add((size_t)B * 4);                         // maxn1
    add((size_t)B * n_max * sizeof(kbtc::Top2) * kbtc::EPI_SLICES);
    add((size_t)B * m_max * sizeof(kbtc::Top2) * kbtc::EPI_SLICES);
    add((size_t)B * n_max * 4);                 // nn0
    add((size_t)B * m_max * 4);                 // nn1
    add((size_t)B * n_max * 8);                 // d2_0
    add(256);                                   // n_exact
    add((size_t)B * (n_max + m_max) * 8);       // rescan list
    add((size_t)B * n_max * 4);                 // keep_j
    add((size_t)B * (n_max + m_max) * 8 * 16);  // rescan partial minima
    add((size_t)B * (n_max + m_max) * 4);       // rescan tickets
    add((size_t)B * n_max * 4);                 // tensor-core score of the direction-0 winners
    add((size_t)B * n_max * 4);                 // ... and their column-side score
    if (kb_knobs[KB_KNOB_TC_ONE_PASS])
        add((size_t)B * 4 * L.tiles0 * L.cs1 * 8);  // group maxima of the column-side scores (one-pass cross-check only)
    add((size_t)B * m_max * 16);                // per-column summary of the group maxima
    add((size_t)B * n_max * 8);                 // columns queued for the exact rescan
    add(8 * 512 * 8);                           // pipeline wait counters (KB_KNOB_TC_DEBUG & 4)
    add((size_t)B * n_max * (L.Dp / 8) * 4);    // partial |row|^2 per 8 components (operands written by the sampler)
    add((size_t)B * m_max * (L.Dp / 8) * 4);
    L.bytes = n + 1024;
    return L;
}

struct TcBuffers {
    unsigned short *S0, *S1;
    float *c0, *c1, *norm2_0, *norm2_1;
    unsigned int *maxn0, *maxn1;
    kbtc::Top2 *res0, *res1;
    int *nn0, *nn1;
    double* d2_0;
    int* n_exact;
    size_t zero_bytes;          // from maxn0 to the end of the tickets
    int2* list;
    int* keep_j;
    void* parts;
    int* tickets;
    float* tsel;
    float* usel;
    float2* gm;
    int4* colinfo;
    int2* list2;
    long long* prof;
    float *part0, *part1;
    bool ok;
};

static TcBuffers tc_carve(void* ws, size_t ws_bytes, int B, int n_max, int m_max, const TcLayout& L) {
    KbArena arena(ws, ws_bytes);
    TcBuffers t;
    t.S0 = arena.take<unsigned short>((size_t)B * n_max * 2 * L.Dp);
    t.S1 = arena.take<unsigned short>((size_t)B * m_max * 2 * L.Dp);
    t.c0 = arena.take<float>((size_t)B * L.cs0);
    t.c1 = arena.take<float>((size_t)B * L.cs1);
    t.norm2_0 = arena.take<float>((size_t)B * n_max);
    t.norm2_1 = arena.take<float>((size_t)B * m_max);
    t.maxn0 = arena.take<unsigned int>(B);
    t.maxn1 = arena.take<unsigned int>(B);
    // maxn0, maxn1, n_exact and the rescan tickets are zeroed per call: kept adjacent for one memset
    t.n_exact = arena.take<int>(4);             // n_exact, n_pair, n_list2, -
    t.tickets = arena.take<int>((size_t)B * (n_max + m_max));
    t.zero_bytes = (size_t)((char*)(t.tickets + (size_t)B * (n_max + m_max)) - (char*)t.maxn0);
    t.res0 = arena.take<kbtc::Top2>((size_t)B * n_max * kbtc::EPI_SLICES);
    t.res1 = arena.take<kbtc::Top2>((size_t)B * m_max * kbtc::EPI_SLICES);
    t.nn0 = arena.take<int>((size_t)B * n_max);
    t.nn1 = arena.take<int>((size_t)B * m_max);
    t.d2_0 = arena.take<double>((size_t)B * n_max);
    t.list = arena.take<int2>((size_t)B * (n_max + m_max));
    t.keep_j = arena.take<int>((size_t)B * n_max);
    t.parts = arena.take<char>((size_t)B * (n_max + m_max) * 8 * 16);
    t.tsel = arena.take<float>((size_t)B * n_max);
    t.usel = arena.take<float>((size_t)B * n_max);
    t.gm = kb_knobs[KB_KNOB_TC_ONE_PASS] ? arena.take<float2>((size_t)B * 4 * L.tiles0 * L.cs1) : nullptr;
    t.colinfo = arena.take<int4>((size_t)B * m_max);
    t.list2 = arena.take<int2>((size_t)B * n_max);
    t.prof = arena.take<long long>(8 * 512);
    t.part0 = arena.take<float>((size_t)B * n_max * (L.Dp / 8));
    t.part1 = arena.take<float>((size_t)B * m_max * (L.Dp / 8));
    t.ok = arena.ok();
    return t;
}

// Diagnostics for the tests: byte offsets inside a kb_match_mnn(algo=1) workspace of
// [0] res0 (Top2[B*n_max]: best, second, idx, pad), [1] res1, [2] n_exact (int), [3] norm2_0, [4] norm2_1.
extern "C" KB_API int kb_match_tc_debug_offsets(int B, int n_max, int m_max, int D, size_t* off) {
    if (!off || B <= 0 || n_max <= 0 || m_max <= 0 || D <= 0) return KB_ERR_BAD_ARG;
    const TcLayout L = tc_layout(B, n_max, m_max, D);
    char* base = (char*)4096;                       // any non-null, 256-aligned base
    TcBuffers t = tc_carve(base, (size_t)-1, B, n_max, m_max, L);
    off[0] = (char*)t.res0 - base; off[1] = (char*)t.res1 - base; off[2] = (char*)t.n_exact - base;
    off[3] = (char*)t.norm2_0 - base; off[4] = (char*)t.norm2_1 - base; off[5] = (char*)t.prof - base;
    return KB_OK;
}

static int tc_operand_fp16(int Dp) {
    // operand split: fp16 halves x 2 products where the search is MMA-bound (D > 64), bf16 halves x 3 where its
    // epilogue is the bound anyway (D <= 64: the wider error bound would only add two-candidate checks and rescans);
    // KB_KNOB_TC_BF16X3 = 1 / 2 forces bf16 x 3 / fp16 x 2
    return kb_knobs[KB_KNOB_TC_BF16X3] == 1 ? 0 : (kb_knobs[KB_KNOB_TC_BF16X3] == 2 ? 1 : (Dp > 64 ? 1 : 0));
}

// Where a producer other than prep_kernel (the fused sampler, kb_sample.cu) writes the operand rows of a
// kb_match_mnn(algo = 1) workspace; the match call that follows passes phases bit 3 instead of bit 0.
int kb_match_tc_operand_sinks(void* ws, size_t ws_bytes, int B, int n_max, int m_max, int D, KbOperandSinks* out) {
    if (B > 65535 || D % 8 != 0) return KB_ERR_UNSUPPORTED;
    const TcLayout L = tc_layout(B, n_max, m_max, D);
    if (L.KB > 4) return KB_ERR_UNSUPPORTED;
    TcBuffers tb = tc_carve(ws, ws_bytes, B, n_max, m_max, L);
    if (!tb.ok) return KB_ERR_WORKSPACE;
    out->S[0] = tb.S0; out->S[1] = tb.S1; out->part[0] = tb.part0; out->part[1] = tb.part1;
    out->Dp = L.Dp; out->fp16 = tc_operand_fp16(L.Dp);
    return KB_OK;
}

size_t kb_match_tc_workspace_bytes(int B, int n_max, int m_max, int D) {
    return tc_layout(B, n_max, m_max, D).bytes;
}

// `phases` (bit 0 operand preparation, bit 1 tensor-core search, bit 2 certification / rescans / gate / pairs)
// lets the benchmark time one part alone on a workspace that a full call has filled; product calls pass 7.
int kb_match_tc_run(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max, int m_max,
                    int D, double max_distance, int cross_check, int* pairs, double* dist, int* count, void* ws,
                    size_t ws_bytes, int phases, cudaStream_t st) {
    using namespace kbtc;
    if (B > 65535) return KB_ERR_UNSUPPORTED;
    const TcLayout L = tc_layout(B, n_max, m_max, D);
    if (L.KB > 4) return KB_ERR_UNSUPPORTED;                // query tile would not stay resident in shared memory
    TcBuffers tb = tc_carve(ws, ws_bytes, B, n_max, m_max, L);
    if (!tb.ok) return KB_ERR_WORKSPACE;
    unsigned short *S0 = tb.S0, *S1 = tb.S1;
    const int fp16 = tc_operand_fp16(L.Dp);
    float *c0 = tb.c0, *c1 = tb.c1, *norm2_0 = tb.norm2_0, *norm2_1 = tb.norm2_1;
    unsigned int *maxn0 = tb.maxn0, *maxn1 = tb.maxn1;
    Top2 *res0 = tb.res0, *res1 = tb.res1;
    int *nn0 = tb.nn0, *nn1 = tb.nn1, *n_exact = tb.n_exact;
    double* d2_0 = tb.d2_0;

    if ((phases & 7) == 7 || (phases & 15) == 14) {
        KB_CUDA_TRY(cudaMemsetAsync(maxn0, 0, tb.zero_bytes, st));      // maxn0, maxn1, n_exact, tickets in one node
    } else {
        if (phases & 9) {
            KB_CUDA_TRY(cudaMemsetAsync(maxn0, 0, (size_t)B * 4, st));
            KB_CUDA_TRY(cudaMemsetAsync(maxn1, 0, (size_t)B * 4, st));
        }
        if (phases & 4) {
            KB_CUDA_TRY(cudaMemsetAsync(n_exact, 0, 16, st));
            KB_CUDA_TRY(cudaMemsetAsync(tb.tickets, 0, (size_t)B * (n_max + m_max) * 4, st));
        }
    }
    int sms = 0;
    { const int rc_sm = kb_sm_count(&sms); if (rc_sm != KB_OK) return rc_sm; }
    {
        PrepPair pp;
        PrepParams& q0 = pp.side[0];
        q0.d = d0; q0.cnt = n0; q0.S = S0; q0.c = c0; q0.norm2 = norm2_0; q0.maxn = maxn0;
        q0.B = B; q0.n_max = n_max; q0.D = D; q0.Dp = L.Dp; q0.cs = L.cs0; q0.fp16 = fp16;
        PrepParams& q1 = pp.side[1];
        q1 = q0;
        q1.d = d1; q1.cnt = n1; q1.S = S1; q1.c = c1; q1.norm2 = norm2_1; q1.maxn = maxn1;
        q1.n_max = m_max; q1.cs = L.cs1;
        // a few waves of CTAs: every warp walks over groups of PREP_ROWS rows
        const int cs_max = L.cs0 > L.cs1 ? L.cs0 : L.cs1;
        int ctas = (cs_max / PREP_ROWS + 7) / 8;                     // CTAs that give every warp one group
        const int wave = 4 * ((sms * 8 + 2 * B - 1) / (2 * B));         // CTAs per (batch, side): four resident waves (49.9 us against 52.8 us with one)
        if (ctas > wave) ctas = wave;
        if (ctas < 1) ctas = 1;
        if (phases & 1) {
            prep_kernel<<<dim3(ctas, B, 2), 256, 0, st>>>(pp);
            KB_LAUNCH_CHECK();
        }
        if (phases & 8) {                   // the operand rows and partial norms are in place (kb_sample_desc_operands)
            PrepFinishPair fp;
            fp.side[0] = PrepFinish{tb.part0, n0, c0, norm2_0, maxn0, n_max, L.cs0, L.Dp / 8};
            fp.side[1] = PrepFinish{tb.part1, n1, c1, norm2_1, maxn1, m_max, L.cs1, L.Dp / 8};
            prep_finish_kernel<<<dim3((cs_max + 255) / 256, B, 2), 256, 0, st>>>(fp);
            KB_LAUNCH_CHECK();
        }
    }
    // the two operand maps depend on (base, rows, row length) only: steady-state callers reuse their workspace, so the
    // driver call that encodes a map is made once per distinct operand array and thread
    struct MapCache { void* base; uint64_t rows, ks; int fp16; CUtensorMap map; };
    static thread_local MapCache cache[2] = {};
    auto cached_map = [&](int slot, void* base, uint64_t rows, uint64_t ks, CUtensorMap* out) -> int {
        MapCache& c = cache[slot];
        if (c.base != base || c.rows != rows || c.ks != ks || c.fp16 != fp16) {
            const int r = make_map(&c.map, base, rows, ks, fp16);
            if (r != KB_OK) { c.base = nullptr; return r; }
            c.base = base; c.rows = rows; c.ks = ks; c.fp16 = fp16;
        }
        *out = c.map;
        return KB_OK;
    };
    CUtensorMap map0, map1;
    int rc = cached_map(0, S0, (uint64_t)B * n_max, (uint64_t)2 * L.Dp, &map0);
    if (rc != KB_OK) return rc;
    rc = cached_map(1, S1, (uint64_t)B * m_max, (uint64_t)2 * L.Dp, &map1);
    if (rc != KB_OK) return rc;

    MainParams mp;
    mp.n0 = n0; mp.n1 = n1; mp.c0 = c0; mp.c1 = c1; mp.res0 = res0; mp.res1 = res1;
    mp.B = B; mp.n_max = n_max; mp.m_max = m_max; mp.cs0 = L.cs0; mp.cs1 = L.cs1; mp.KB = L.KB;
    // cross-check: the two-direction search (default), or one Gram pass + column-group maxima (KB_KNOB_TC_ONE_PASS)
    const int colside = (cross_check && kb_knobs[KB_KNOB_TC_ONE_PASS]) ? 1 : 0;
    mp.tiles0 = L.tiles0; mp.tiles1 = L.tiles1; mp.n_dirs = (cross_check && !colside) ? 2 : 1;
    mp.fp16 = fp16;
    mp.gm = tb.gm; mp.maxn0 = maxn0; mp.maxn1 = maxn1; mp.gm_groups = 4 * L.tiles0; mp.colside = colside;
    // after the tiles: barriers (256 bytes), the 2 x 256-float c buffer and the 128-float shared bound; the tiles must start on a
    // 1024-byte boundary (128B swizzle) -- dynamic shared memory normally does, `slack` covers the rest
    const size_t fixed = 256 + 2 * BN * 4 + BM * 4;
    const size_t budget = 227 * 1024;
    // CTA pairs (kb_debug_knob(KB_KNOB_TC_CLUSTER, 2): thread-block clusters of 2, tcgen05 cta_group::2 MMAs with M = 256): the pair
    // multiplies two adjacent query tiles against one database stream of which every CTA holds half, so the operand
    // bytes per CTA halve and the ring is twice as deep.  Measured on B200 it is neither faster nor slower than the
    // 1-CTA kernel (172 vs 170-178 us on cfg2): the kernel's time goes to the tensor pipe sharing TMEM with the
    // epilogue's tcgen05.ld and to the epilogue itself, not to operand delivery (scripts/tc_pipeline_profile.py), so
    // the 1-CTA kernel stays the default and the pair kernel is kept selectable and tested.
    int cl = 1;
    if (kb_knobs[KB_KNOB_TC_CLUSTER] == 2 && (L.tiles0 > 1 || L.tiles1 > 1)) cl = 2;
    mp.cl = cl;
    const size_t slot_bytes = cl == 2 ? TILE_BYTES : SLOT_BYTES;
    int n_slots = (int)((budget - fixed - (size_t)2 * L.KB * TILE_BYTES) / slot_bytes);
    if (n_slots > MAX_SLOTS) n_slots = MAX_SLOTS;
    if (n_slots < 2) return KB_ERR_UNSUPPORTED;
    const size_t used = (size_t)2 * L.KB * TILE_BYTES + (size_t)n_slots * slot_bytes + fixed;
    size_t slack = budget - used;
    if (slack > 1024) slack = 1024;
    mp.n_slots = n_slots;
    mp.align_slack = (int)slack;
    const size_t smem = used + slack;
    mp.dbg = kb_knobs[KB_KNOB_TC_DEBUG];
    mp.prof = tb.prof;
    static bool attr_set = false;                 // the opt-in to > 48 KB of dynamic shared memory is per process
    if (!attr_set) {
        KB_CUDA_TRY(cudaFuncSetAttribute(nn_top2_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        KB_CUDA_TRY(cudaFuncSetAttribute(nn_top2_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        KB_CUDA_TRY(cudaFuncSetAttribute(nn_top2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        KB_CUDA_TRY(cudaFuncSetAttribute(nn_top2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        attr_set = true;
    }
    const int g0 = (L.tiles0 + cl - 1) / cl, g1 = (L.tiles1 + cl - 1) / cl;
    const int n_items = mp.n_dirs == 2 ? B * (g0 + g1) : B * g0;
    const int max_groups = sms / cl;
    const int grid = (n_items < max_groups ? n_items : max_groups) * cl;
    if ((phases & 2) && cl == 1) {
        if (colside) nn_top2_kernel<1, true><<<grid, NT, smem, st>>>(map0, map1, mp);
        else nn_top2_kernel<1, false><<<grid, NT, smem, st>>>(map0, map1, mp);
        KB_LAUNCH_CHECK();
    } else if (phases & 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cl; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (colside) KB_CUDA_TRY(cudaLaunchKernelEx(&cfg, nn_top2_kernel<2, true>, map0, map1, mp));
        else KB_CUDA_TRY(cudaLaunchKernelEx(&cfg, nn_top2_kernel<2, false>, map0, map1, mp));
        KB_LAUNCH_CHECK();
    }
    if (!(phases & 4)) return KB_OK;

    ResolveParams rp;
    rp.d0 = d0; rp.d1 = d1; rp.n0 = n0; rp.n1 = n1; rp.res0 = res0; rp.res1 = res1;
    rp.norm2_0 = norm2_0; rp.norm2_1 = norm2_1; rp.maxn0 = maxn0; rp.maxn1 = maxn1;
    rp.nn0 = nn0; rp.nn1 = nn1; rp.n_exact = n_exact; rp.n_pair = n_exact + 1;
    rp.list = tb.list; rp.list_cap = B * (n_max + m_max);
    rp.parts = (RescanPart*)tb.parts; rp.tickets = tb.tickets; rp.tsel = tb.tsel; rp.usel = tb.usel;
    rp.gm = tb.gm; rp.colinfo = tb.colinfo; rp.gm_groups = mp.gm_groups; rp.cs1 = L.cs1; rp.colside = colside;
    rp.B = B; rp.n_max = n_max; rp.m_max = m_max; rp.D = D; rp.n_dirs = mp.n_dirs; rp.fp16 = fp16; rp.Dp = L.Dp;
    const int qmax = n_max > m_max ? n_max : m_max;
    // grid plane 1: the other direction's rows (two-pass) or the per-column summary of the group maxima (one-pass)
    resolve_kernel<<<dim3((qmax + 255) / 256, B, cross_check ? 2 : 1), 256, 0, st>>>(rp);
    KB_LAUNCH_CHECK();
    if ((size_t)D * 4 > 48 * 1024) return KB_ERR_UNSUPPORTED;
    // (queued rows are rare: a quarter of the SMs' worth of row slots keeps the usual, empty launch short; the kernel
    // loops over the queue, so any number of rows is handled)
    const int rescan_rows = sms / 4 > 1 ? sms / 4 : 1;
    rescan_kernel<<<dim3(RESCAN_SPLIT, rescan_rows), RESCAN_WARPS * 32, (size_t)D * 4, st>>>(rp);
    KB_LAUNCH_CHECK();

    GateParams gp;
    gp.d0 = d0; gp.d1 = d1; gp.n0 = n0; gp.n1 = n1; gp.nn0 = nn0; gp.nn1 = nn1; gp.keep_j = tb.keep_j; gp.dist_i = d2_0;
    gp.n_max = n_max; gp.m_max = m_max; gp.D = D; gp.cross_check = cross_check; gp.max_distance = max_distance;
    gp.tsel = tb.tsel; gp.usel = tb.usel; gp.colinfo = tb.colinfo; gp.norm2_0 = norm2_0; gp.maxn0 = maxn0; gp.maxn1 = maxn1;
    gp.list2 = tb.list2; gp.n_list2 = n_exact + 2; gp.list2_cap = B * n_max; gp.colside = colside;
    gp.want_dist = dist != nullptr; gp.fp16 = fp16; gp.Dp = L.Dp;
    const dim3 gate_grid(((n_max + 3) / 4 * 32 + 255) / 256, B);
    gate_kernel<<<gate_grid, 256, 0, st>>>(gp, 0);
    KB_LAUNCH_CHECK();
    if (colside) {
        // the columns that were too close to call: exact float64 rescan, then the gate for the rows waiting on them
        ResolveParams rp2 = rp;
        rp2.list = tb.list2; rp2.n_exact = n_exact + 2; rp2.n_pair = nullptr; rp2.list_cap = B * n_max;
        rescan_kernel<<<dim3(RESCAN_SPLIT, rescan_rows), RESCAN_WARPS * 32, (size_t)D * 4, st>>>(rp2);
        KB_LAUNCH_CHECK();
        gate_kernel<<<gate_grid, 256, 0, st>>>(gp, 1);
        KB_LAUNCH_CHECK();
    }

    PairsParams pp;
    pp.n0 = n0; pp.n1 = n1; pp.keep_j = tb.keep_j; pp.dist_i = d2_0; pp.pairs = pairs; pp.dist = dist;
    pp.count = count; pp.n_max = n_max; pp.m_max = m_max;
    pairs_kernel<<<B, 1024, 0, st>>>(pp);
    KB_LAUNCH_CHECK();
    return KB_OK;
}
