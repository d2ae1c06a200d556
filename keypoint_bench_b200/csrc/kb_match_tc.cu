// Tensor-core mutual-NN matcher (algo 1): tcgen05 Gram tiles in TMEM with a fused top-2 epilogue,
// float64 certification of the winners.
// Reference semantics: utils/matcher.py:227-234 -> skimage match_descriptors over float64 cdist.
//
// For a query row x and database rows y_j:  argmin_j |x-y_j|^2  ==  argmax_j  t_j = x.y_j - |y_j|^2/2.
//   prep_kernel     float32 descriptors -> split-bf16 operand rows [hi(D) | lo(D)] (x = hi + lo up to
//                   2^-18 |x|), c_j = -|y_j|^2/2, row norms.  Three bf16 MMAs (hi.hi + hi.lo + lo.hi)
//                   reproduce x.y to ~2^-16 relative: float32-grade, at bf16 tensor throughput.
//   nn_top2_kernel  persistent, warp-specialised: one TMA producer lane, one MMA-issuing lane
//                   (tcgen05.mma cta_group::1 kind::f16, M=128 N=128 K=16, accumulators double-buffered
//                   in TMEM), four epilogue warps (tcgen05.ld 32x32b.x32) that fold every 128x128 tile
//                   into a per-row running (best, second best, argbest).  The [n,m] matrix never
//                   leaves the SM.  The query tile's operand rows stay resident in shared memory while
//                   the database tiles stream through a 6-slot TMA ring (128B-swizzled K-major tiles).
//                   Both directions (rows->cols, cols->rows for the cross-check) are work items of
//                   the same launch.
//   resolve_kernel  one warp per query: if best-second exceeds the a-priori error bound of the split
//                   product the argmax is certified and only its float64 distance is evaluated;
//                   otherwise the row is rescanned exactly in float64 (first of ties, as np.argmin).
//   pairs_kernel    mutual check, strict < max_distance gate, ordered compaction (per pair).
#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>
#include "kb_common.cuh"

namespace kbtc {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int TILE_BYTES = BM * BK * 2;          // 16 KB, one 128B-swizzled K-major tile
constexpr int NSLOT = 6;
constexpr int NT = 256;                          // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-7 epilogue
constexpr int TMEM_COLS = 256;                   // two 128-column fp32 accumulators

struct Top2 {
    float best, second;
    int idx, pad;
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO=64), LBO=1,
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
struct PrepParams {
    const float* d;          // [B,n_max,D]
    const int* cnt;          // [B] or null
    __nv_bfloat16* S;        // [B*n_max, 2*Dp]
    float* c;                // [B, cs]   -|y|^2/2, -inf beyond the count
    float* norm2;            // [B*n_max]
    unsigned int* maxn;      // [B] max |row|^2 (float bits)
    int B, n_max, D, Dp, cs;
};

__global__ void __launch_bounds__(256) prep_kernel(PrepParams p) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (warp >= p.cs) return;
    const int n = p.cnt ? p.cnt[b] : p.n_max;
    const int row = warp;
    if (row >= p.n_max) {                       // padding of the c array up to a whole column tile
        if (lane == 0) p.c[(size_t)b * p.cs + row] = -CUDART_INF_F;
        return;
    }
    __nv_bfloat16* out = p.S + ((size_t)b * p.n_max + row) * (2 * p.Dp);
    if (row >= n) {
        for (int k = lane; k < 2 * p.Dp; k += 32) out[k] = __float2bfloat16(0.0f);
        if (lane == 0) { p.c[(size_t)b * p.cs + row] = -CUDART_INF_F; p.norm2[(size_t)b * p.n_max + row] = 0.0f; }
        return;
    }
    const float* x = p.d + ((size_t)b * p.n_max + row) * p.D;
    float ss = 0.0f;
    if ((p.D & 3) == 0 && (reinterpret_cast<uintptr_t>(p.d) & 15u) == 0) {
        // 4 components per lane and step: one 16-byte load, two 8-byte stores
        for (int k = 4 * lane; k < p.Dp; k += 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < p.D) v = __ldg(reinterpret_cast<const float4*>(x + k));
            const float f[4] = {v.x, v.y, v.z, v.w};
            __nv_bfloat16 h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                h[e] = __float2bfloat16_rn(f[e]);
                l[e] = __float2bfloat16_rn(f[e] - __bfloat162float(h[e]));
                ss = fmaf(f[e], f[e], ss);
            }
            *reinterpret_cast<uint2*>(out + k) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(out + p.Dp + k) = *reinterpret_cast<const uint2*>(l);
        }
    } else {
        for (int k = lane; k < p.Dp; k += 32) {
            const float v = k < p.D ? x[k] : 0.0f;
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
            out[k] = h;
            out[p.Dp + k] = l;
            ss = fmaf(v, v, ss);
        }
    }
    for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    if (lane == 0) {
        p.c[(size_t)b * p.cs + row] = -0.5f * ss;
        p.norm2[(size_t)b * p.n_max + row] = ss;
        atomicMax(&p.maxn[b], __float_as_uint(ss));
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
    int B, n_max, m_max, cs0, cs1, KB, tiles0, tiles1, n_dirs;
};

struct Item {
    int dir, b, q_row0, n_q, n_db, q_base, db_base;
};

__device__ __forceinline__ bool decode_item(const MainParams& p, int item, Item& it) {
    const int per_dir0 = p.B * p.tiles0;
    int dir = 0, rem = item;
    if (item >= per_dir0) { dir = 1; rem = item - per_dir0; }
    const int tiles = dir ? p.tiles1 : p.tiles0;
    const int b = rem / tiles, tile = rem - b * tiles;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    it.dir = dir; it.b = b;
    it.q_row0 = tile * BM;
    it.n_q = dir ? m : n;
    it.n_db = dir ? n : m;
    it.q_base = dir ? b * p.m_max : b * p.n_max;
    it.db_base = dir ? b * p.n_max : b * p.m_max;
    return it.q_row0 < it.n_q && it.n_db > 0;
}

__global__ void __launch_bounds__(NT, 1) nn_top2_kernel(const __grid_constant__ CUtensorMap map0,
                                                        const __grid_constant__ CUtensorMap map1, MainParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;                   // 1024B alignment for SWIZZLE_128B
    const int KB = p.KB;
    const uint32_t a_tiles = base;                                  // 2*KB tiles: hi blocks then lo blocks
    const uint32_t b_slots = base + (uint32_t)(2 * KB) * TILE_BYTES;
    const uint32_t bars = b_slots + NSLOT * TILE_BYTES;
    // barrier map (8 bytes each)
    const uint32_t bar_a_full = bars, bar_a_free = bars + 8;
    const uint32_t bar_b_full = bars + 16, bar_b_empty = bars + 16 + 8 * NSLOT;
    const uint32_t bar_t_full = bars + 16 + 16 * NSLOT, bar_t_empty = bar_t_full + 16;
    const uint32_t tmem_slot = bar_t_empty + 16;
    float* cbuf = reinterpret_cast<float*>(smem_dyn + (tmem_slot + 16 - raw));     // [2][BN]
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_free, 1);
        for (int s = 0; s < NSLOT; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_t_full + 8 * s, 1); mbar_init(bar_t_empty + 8 * s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int n_items = p.n_dirs == 2 ? p.B * (p.tiles0 + p.tiles1) : p.B * p.tiles0;

    if (warp == 0) {
        // ================================ TMA producer ==========================================
        if (lane == 0) {
            uint32_t a_phase = 0, slot = 0, b_phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                Item it;
                if (!decode_item(p, item, it)) continue;
                const CUtensorMap* qmap = it.dir ? &map1 : &map0;
                const CUtensorMap* dmap = it.dir ? &map0 : &map1;
                mbar_wait(bar_a_free, a_phase ^ 1);                 // previous item's MMAs are done with A
                mbar_expect_tx(bar_a_full, (uint32_t)(2 * KB) * TILE_BYTES);
                for (int t = 0; t < 2 * KB; ++t)
                    tma_load_2d(a_tiles + t * TILE_BYTES, qmap, t * BK, it.q_base + it.q_row0, bar_a_full);
                a_phase ^= 1;
                const int n_ct = (it.n_db + BN - 1) / BN;
                for (int ct = 0; ct < n_ct; ++ct) {
                    for (int kb = 0; kb < KB; ++kb) {
                        for (int part = 0; part < 2; ++part) {      // hi block then lo block
                            mbar_wait(bar_b_empty + 8 * slot, b_phase ^ 1);
                            mbar_expect_tx(bar_b_full + 8 * slot, TILE_BYTES);
                            tma_load_2d(b_slots + slot * TILE_BYTES, dmap, (part * KB + kb) * BK,
                                        it.db_base + ct * BN, bar_b_full + 8 * slot);
                            if (++slot == NSLOT) { slot = 0; b_phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ============================================
        if (lane == 0) {
            uint32_t a_phase = 0, slot = 0, b_phase = 0, acc_buf = 0, t_phase[2] = {0, 0};
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                Item it;
                if (!decode_item(p, item, it)) continue;
                mbar_wait(bar_a_full, a_phase);
                a_phase ^= 1;
                tc_fence_after();
                const int n_ct = (it.n_db + BN - 1) / BN;
                for (int ct = 0; ct < n_ct; ++ct) {
                    mbar_wait(bar_t_empty + 8 * acc_buf, t_phase[acc_buf] ^ 1);   // epilogue drained this buffer
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc_buf * BN;
                    uint32_t accum = 0;
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint32_t a_hi = a_tiles + kb * TILE_BYTES, a_lo = a_tiles + (KB + kb) * TILE_BYTES;
                        // ---- database hi block: hi.hi and lo.hi
                        mbar_wait(bar_b_full + 8 * slot, b_phase);
                        tc_fence_after();
                        uint32_t bt = b_slots + slot * TILE_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            tc_mma(d_tmem, umma_desc(a_hi + k * 32), umma_desc(bt + k * 32), IDESC, accum);
                            accum = 1;
                        }
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc_mma(d_tmem, umma_desc(a_lo + k * 32), umma_desc(bt + k * 32), IDESC, 1);
                        tc_commit(bar_b_empty + 8 * slot);
                        if (++slot == NSLOT) { slot = 0; b_phase ^= 1; }
                        // ---- database lo block: hi.lo
                        mbar_wait(bar_b_full + 8 * slot, b_phase);
                        tc_fence_after();
                        bt = b_slots + slot * TILE_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc_mma(d_tmem, umma_desc(a_hi + k * 32), umma_desc(bt + k * 32), IDESC, 1);
                        tc_commit(bar_b_empty + 8 * slot);
                        if (++slot == NSLOT) { slot = 0; b_phase ^= 1; }
                    }
                    tc_commit(bar_t_full + 8 * acc_buf);            // accumulator ready for the epilogue
                    t_phase[acc_buf] ^= 1;
                    acc_buf ^= 1;
                }
                tc_commit(bar_a_free);                              // A tiles may be overwritten
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ==============================================
        const int ew = warp - 4;                                    // TMEM lanes 32*ew .. 32*ew+31
        const int row_in_tile = ew * 32 + lane;
        uint32_t acc_buf = 0, t_phase[2] = {0, 0};
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            Item it;
            if (!decode_item(p, item, it)) continue;
            const float* cvec = (it.dir ? p.c0 : p.c1) + (size_t)it.b * (it.dir ? p.cs0 : p.cs1);
            float best = -CUDART_INF_F, second = -CUDART_INF_F;
            int bj = 0;
            const int n_ct = (it.n_db + BN - 1) / BN;
            // -|y|^2/2 of the next column tile travels through a register while the current tile is
            // reduced, then through a 2 x 128-float shared buffer (one named barrier per tile)
            float c_next = __ldg(cvec + ew * 32 + lane);
            for (int ct = 0; ct < n_ct; ++ct) {
                float* cb = cbuf + (ct & 1) * BN;
                cb[ew * 32 + lane] = c_next;
                if (ct + 1 < n_ct) c_next = __ldg(cvec + (ct + 1) * BN + ew * 32 + lane);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                mbar_wait(bar_t_full + 8 * acc_buf, t_phase[acc_buf]);
                t_phase[acc_buf] ^= 1;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc_buf * BN;
                uint32_t va[32], vb[32];
                auto fold = [&](const uint32_t (&v)[32], int cc) {
                    const float4* c4 = reinterpret_cast<const float4*>(cb + cc * 32);
                    const int j0 = ct * BN + cc * 32;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 t4 = c4[q];
                        const float cv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float t = __uint_as_float(v[4 * q + e]) + cv[e];   // x.y - |y|^2/2 (-inf beyond the count)
                            second = fmaxf(second, fminf(t, best));
                            bj = (t > best) ? (j0 + 4 * q + e) : bj;                 // strict: first of ties
                            best = fmaxf(best, t);
                        }
                    }
                };
                tmem_ld32(taddr, va);
                tmem_ld_wait();
                tmem_ld32(taddr + 32, vb);
                fold(va, 0);
                tmem_ld_wait();
                tmem_ld32(taddr + 64, va);
                fold(vb, 1);
                tmem_ld_wait();
                tmem_ld32(taddr + 96, vb);
                fold(va, 2);
                tmem_ld_wait();
                // the accumulator is in registers: hand the TMEM buffer back before the last fold
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_t_empty + 8 * acc_buf);
                fold(vb, 3);
                acc_buf ^= 1;
            }
            const int qi = it.q_row0 + row_in_tile;
            if (qi < it.n_q) {
                Top2 o; o.best = best; o.second = second; o.idx = bj; o.pad = 0;
                (it.dir ? p.res1 : p.res0)[(size_t)it.q_base + qi] = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// certification / exact resolution (one warp per query)
// ------------------------------------------------------------------------------------------------
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
    int2* list;              // [list_cap] queued rows: (dir | b << 1, query row)
    int list_cap;
    int B, n_max, m_max, D, n_dirs;
};

__device__ __forceinline__ double warp_dist2(const float* x, const float* y, int D, int lane) {
    double acc = 0.0;
    for (int k = lane; k < D; k += 32) {
        const double d = (double)x[k] - (double)y[k];
        acc = fma(d, d, acc);
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    return acc;
}

// One thread per query row: certify the tensor-core argmax or queue the row for the exact rescan.
__global__ void __launch_bounds__(256) resolve_kernel(ResolveParams p) {
    const int q = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y, dir = blockIdx.z;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    const int nq = dir ? m : n, ndb = dir ? n : m;
    if (q >= nq || ndb <= 0) return;
    const int q_stride = dir ? p.m_max : p.n_max;
    const Top2 r = (dir ? p.res1 : p.res0)[(size_t)b * q_stride + q];
    const float nq2 = (dir ? p.norm2_1 : p.norm2_0)[(size_t)b * q_stride + q];
    const float dbmax2 = __uint_as_float((dir ? p.maxn0 : p.maxn1)[b]);
    // a-priori bound on |t_computed - t_exact|: dropped lo.lo / residual terms (3*2^-18 |x||y|), fp32
    // accumulation in the tensor core (K/16 roundings) and the fp32 -|y|^2/2 term; generous factor on top
    const float e = 6.2e-5f * sqrtf(nq2) * sqrtf(dbmax2) + 3.1e-5f * dbmax2;
    const int j = r.idx;
    const bool certain = (r.best - r.second) > 2.0f * e && j >= 0 && j < ndb;
    if (certain) {
        (dir ? p.nn1 : p.nn0)[(size_t)b * q_stride + q] = j;
    } else {
        const int slot = atomicAdd(p.n_exact, 1);
        if (slot < p.list_cap) p.list[slot] = make_int2(dir | (b << 1), q);
    }
}

// Exact float64 resolution of the queued rows (best/second closer than the error bound of the split
// product, e.g. duplicated descriptors): one CTA per row, lanes over components, eight candidates in
// flight per warp (16 independent 16-byte loads per lane at D=256).  First of ties wins, as np.argmin.
template <bool VEC>
__device__ __forceinline__ void rescan_row(const float* xs, const float* DBs, int D, int ndb, int warp, int lane,
                                           double& bd, int& bj) {
    constexpr int G = 8;
    for (int c0 = warp * G; c0 < ndb; c0 += 8 * G) {
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

__global__ void __launch_bounds__(256) rescan_kernel(ResolveParams p) {
    extern __shared__ __align__(16) float xs[];
    __shared__ double s_d[8];
    __shared__ int s_j[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int total = *p.n_exact;
    if (total > p.list_cap) total = p.list_cap;
    const bool vec = (p.D & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.d0) | reinterpret_cast<uintptr_t>(p.d1)) & 15u) == 0;
    for (int e = blockIdx.x; e < total; e += gridDim.x) {
        const int2 ent = p.list[e];
        const int dir = ent.x & 1, b = ent.x >> 1, q = ent.y;
        const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
        const int ndb = dir ? n : m;
        const int q_stride = dir ? p.m_max : p.n_max, db_stride = dir ? p.n_max : p.m_max;
        const float* Q = (dir ? p.d1 : p.d0) + ((size_t)b * q_stride + q) * p.D;
        const float* DBs = (dir ? p.d0 : p.d1) + (size_t)b * db_stride * p.D;
        __syncthreads();
        for (int k = threadIdx.x; k < p.D; k += 256) xs[k] = Q[k];
        __syncthreads();
        double bd = CUDART_INF;
        int bj = 0x7fffffff;
        if (vec) rescan_row<true>(xs, DBs, p.D, ndb, warp, lane, bd, bj);
        else rescan_row<false>(xs, DBs, p.D, ndb, warp, lane, bd, bj);
        if (lane == 0) { s_d[warp] = bd; s_j[warp] = bj; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double d2 = s_d[0];
            int j = s_j[0];
            for (int w = 1; w < 8; ++w)
                if (s_d[w] < d2 || (s_d[w] == d2 && s_j[w] < j)) { d2 = s_d[w]; j = s_j[w]; }
            (dir ? p.nn1 : p.nn0)[(size_t)b * q_stride + q] = j;
        }
    }
}

// One warp per row of d0: mutual check, float64 distance of the surviving pair, strict < max_distance.
struct GateParams {
    const float* d0;
    const float* d1;
    const int* n0;
    const int* n1;
    const int* nn0;
    const int* nn1;
    int* keep_j;             // [B*n_max] matched column or -1
    double* dist_i;          // [B*n_max]
    int n_max, m_max, D, cross_check;
    double max_distance;
};

__global__ void __launch_bounds__(256) gate_kernel(GateParams p) {
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int n = p.n0 ? p.n0[b] : p.n_max, m = p.n1 ? p.n1[b] : p.m_max;
    if (i >= n || m <= 0) return;
    const int j = p.nn0[(size_t)b * p.n_max + i];
    int keep = -1;
    double d = 0.0;
    if (!p.cross_check || p.nn1[(size_t)b * p.m_max + j] == i) {
        d = sqrt(warp_dist2(p.d0 + ((size_t)b * p.n_max + i) * p.D, p.d1 + ((size_t)b * p.m_max + j) * p.D, p.D, lane));
        if (d < p.max_distance) keep = j;
    }
    if (lane == 0) { p.keep_j[(size_t)b * p.n_max + i] = keep; p.dist_i[(size_t)b * p.n_max + i] = d; }
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

static int make_map(CUtensorMap* map, void* base, uint64_t rows, uint64_t ks) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return KB_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {ks, rows};
    cuuint64_t strides[1] = {ks * 2};
    cuuint32_t box[2] = {(cuuint32_t)kbtc::BK, (cuuint32_t)kbtc::BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
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
    L.cs0 = L.tiles0 * kbtc::BM;
    L.cs1 = L.tiles1 * kbtc::BM;
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
    add((size_t)B * n_max * sizeof(kbtc::Top2));
    add((size_t)B * m_max * sizeof(kbtc::Top2));
    add((size_t)B * n_max * 4);                 // nn0
    add((size_t)B * m_max * 4);                 // nn1
    add((size_t)B * n_max * 8);                 // d2_0
    add(256);                                   // n_exact
    add((size_t)B * (n_max + m_max) * 8);       // rescan list
    add((size_t)B * n_max * 4);                 // keep_j
    L.bytes = n + 1024;
    return L;
}

struct TcBuffers {
    __nv_bfloat16 *S0, *S1;
    float *c0, *c1, *norm2_0, *norm2_1;
    unsigned int *maxn0, *maxn1;
    kbtc::Top2 *res0, *res1;
    int *nn0, *nn1;
    double* d2_0;
    int* n_exact;
    int2* list;
    int* keep_j;
    bool ok;
};

static TcBuffers tc_carve(void* ws, size_t ws_bytes, int B, int n_max, int m_max, const TcLayout& L) {
    KbArena arena(ws, ws_bytes);
    TcBuffers t;
    t.S0 = arena.take<__nv_bfloat16>((size_t)B * n_max * 2 * L.Dp);
    t.S1 = arena.take<__nv_bfloat16>((size_t)B * m_max * 2 * L.Dp);
    t.c0 = arena.take<float>((size_t)B * L.cs0);
    t.c1 = arena.take<float>((size_t)B * L.cs1);
    t.norm2_0 = arena.take<float>((size_t)B * n_max);
    t.norm2_1 = arena.take<float>((size_t)B * m_max);
    t.maxn0 = arena.take<unsigned int>(B);
    t.maxn1 = arena.take<unsigned int>(B);
    t.res0 = arena.take<kbtc::Top2>((size_t)B * n_max);
    t.res1 = arena.take<kbtc::Top2>((size_t)B * m_max);
    t.nn0 = arena.take<int>((size_t)B * n_max);
    t.nn1 = arena.take<int>((size_t)B * m_max);
    t.d2_0 = arena.take<double>((size_t)B * n_max);
    t.n_exact = arena.take<int>(1);
    t.list = arena.take<int2>((size_t)B * (n_max + m_max));
    t.keep_j = arena.take<int>((size_t)B * n_max);
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
    off[3] = (char*)t.norm2_0 - base; off[4] = (char*)t.norm2_1 - base;
    return KB_OK;
}

size_t kb_match_tc_workspace_bytes(int B, int n_max, int m_max, int D) {
    return tc_layout(B, n_max, m_max, D).bytes;
}

int kb_match_tc_run(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max, int m_max,
                    int D, double max_distance, int cross_check, int* pairs, double* dist, int* count, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
    using namespace kbtc;
    if (B > 65535) return KB_ERR_UNSUPPORTED;
    const TcLayout L = tc_layout(B, n_max, m_max, D);
    if (L.KB > 4) return KB_ERR_UNSUPPORTED;                // query tile would not stay resident in shared memory
    TcBuffers tb = tc_carve(ws, ws_bytes, B, n_max, m_max, L);
    if (!tb.ok) return KB_ERR_WORKSPACE;
    __nv_bfloat16 *S0 = tb.S0, *S1 = tb.S1;
    float *c0 = tb.c0, *c1 = tb.c1, *norm2_0 = tb.norm2_0, *norm2_1 = tb.norm2_1;
    unsigned int *maxn0 = tb.maxn0, *maxn1 = tb.maxn1;
    Top2 *res0 = tb.res0, *res1 = tb.res1;
    int *nn0 = tb.nn0, *nn1 = tb.nn1, *n_exact = tb.n_exact;
    double* d2_0 = tb.d2_0;

    KB_CUDA_TRY(cudaMemsetAsync(maxn0, 0, (size_t)B * 4, st));
    KB_CUDA_TRY(cudaMemsetAsync(maxn1, 0, (size_t)B * 4, st));
    KB_CUDA_TRY(cudaMemsetAsync(n_exact, 0, 4, st));
    {
        PrepParams q;
        q.d = d0; q.cnt = n0; q.S = S0; q.c = c0; q.norm2 = norm2_0; q.maxn = maxn0;
        q.B = B; q.n_max = n_max; q.D = D; q.Dp = L.Dp; q.cs = L.cs0;
        prep_kernel<<<dim3((L.cs0 * 32 + 255) / 256, B), 256, 0, st>>>(q);
        KB_LAUNCH_CHECK();
        q.d = d1; q.cnt = n1; q.S = S1; q.c = c1; q.norm2 = norm2_1; q.maxn = maxn1;
        q.n_max = m_max; q.cs = L.cs1;
        prep_kernel<<<dim3((L.cs1 * 32 + 255) / 256, B), 256, 0, st>>>(q);
        KB_LAUNCH_CHECK();
    }
    CUtensorMap map0, map1;
    int rc = make_map(&map0, S0, (uint64_t)B * n_max, (uint64_t)2 * L.Dp);
    if (rc != KB_OK) return rc;
    rc = make_map(&map1, S1, (uint64_t)B * m_max, (uint64_t)2 * L.Dp);
    if (rc != KB_OK) return rc;

    MainParams mp;
    mp.n0 = n0; mp.n1 = n1; mp.c0 = c0; mp.c1 = c1; mp.res0 = res0; mp.res1 = res1;
    mp.B = B; mp.n_max = n_max; mp.m_max = m_max; mp.cs0 = L.cs0; mp.cs1 = L.cs1; mp.KB = L.KB;
    mp.tiles0 = L.tiles0; mp.tiles1 = L.tiles1; mp.n_dirs = cross_check ? 2 : 1;
    const size_t smem = (size_t)(2 * L.KB + NSLOT) * TILE_BYTES + 1024 + 256 + 2 * BN * 4;
    KB_CUDA_TRY(cudaFuncSetAttribute(nn_top2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    KB_CUDA_TRY(cudaGetDevice(&dev));
    KB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n_items = mp.n_dirs == 2 ? B * (L.tiles0 + L.tiles1) : B * L.tiles0;
    const int grid = n_items < sms ? n_items : sms;
    nn_top2_kernel<<<grid, NT, smem, st>>>(map0, map1, mp);
    KB_LAUNCH_CHECK();

    ResolveParams rp;
    rp.d0 = d0; rp.d1 = d1; rp.n0 = n0; rp.n1 = n1; rp.res0 = res0; rp.res1 = res1;
    rp.norm2_0 = norm2_0; rp.norm2_1 = norm2_1; rp.maxn0 = maxn0; rp.maxn1 = maxn1;
    rp.nn0 = nn0; rp.nn1 = nn1; rp.n_exact = n_exact;
    rp.list = tb.list; rp.list_cap = B * (n_max + m_max);
    rp.B = B; rp.n_max = n_max; rp.m_max = m_max; rp.D = D; rp.n_dirs = mp.n_dirs;
    const int qmax = n_max > m_max ? n_max : m_max;
    resolve_kernel<<<dim3((qmax + 255) / 256, B, mp.n_dirs), 256, 0, st>>>(rp);
    KB_LAUNCH_CHECK();
    if ((size_t)D * 4 > 48 * 1024) return KB_ERR_UNSUPPORTED;
    rescan_kernel<<<sms * 4, 256, (size_t)D * 4, st>>>(rp);
    KB_LAUNCH_CHECK();

    GateParams gp;
    gp.d0 = d0; gp.d1 = d1; gp.n0 = n0; gp.n1 = n1; gp.nn0 = nn0; gp.nn1 = nn1; gp.keep_j = tb.keep_j; gp.dist_i = d2_0;
    gp.n_max = n_max; gp.m_max = m_max; gp.D = D; gp.cross_check = cross_check; gp.max_distance = max_distance;
    gate_kernel<<<dim3((n_max * 32 + 255) / 256, B), 256, 0, st>>>(gp);
    KB_LAUNCH_CHECK();

    PairsParams pp;
    pp.n0 = n0; pp.n1 = n1; pp.keep_j = tb.keep_j; pp.dist_i = d2_0; pp.pairs = pairs; pp.dist = dist;
    pp.count = count; pp.n_max = n_max; pp.m_max = m_max;
    pairs_kernel<<<B, 1024, 0, st>>>(pp);
    KB_LAUNCH_CHECK();
    return KB_OK;
}
