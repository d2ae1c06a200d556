// Streaming form of round 1 of the sparse exact detection path (contract of the two lists: kb_sparse_nms.cu).
//
// The tiled round1_kernel stages a 2r halo around every 32x128 tile (2.08x the pixels at r = 6) and finds the
// round-1 maxima of the reference's first fast_nms round (utils/extracter.py:54-70) hierarchically, ~100
// thread-instructions per pixel.  This kernel reads every pixel once with no horizontal halo and a vertical halo
// of a few 8-row chunks per band:
//
//   * a CTA owns FULL-WIDTH bands of consecutive 8-row chunks of the stacked batch (total chunks / grid each, so
//     all CTAs are resident at once and finish together); thread t owns columns 4t..4t+3;
//   * rows travel global -> a 32-row ring in shared memory with cp.async (16 bytes per thread and row, zero-filled
//     outside the map = the reference's zero padding, extracter.py:58), issued one iteration ahead;
//   * per chunk: vertical (2r+1)-window maximum of 8 output rows from 8+2r ring rows of the thread's own columns
//     (shared core + running prefix / suffix maxima, ~4 FMNMX per pixel), written to a row buffer; the horizontal
//     (2r+1)-window maximum of a thread's 4 columns from its neighbours' 16-byte pieces of that buffer (shared core
//     again); a pixel is a round-1 candidate iff it equals its window maximum and exceeds tau;
//   * torch.argmax returns the FIRST maximum of the window (extracter.py:69-70): a candidate is a maximum iff no
//     element EARLIER in raster order equals it.  Candidates are rare (1 pixel in ~(2r+1)^2) and handled by the whole
//     warp, one at a time: one ballot tells which columns of the window hold an equal value at all (from the row
//     buffer) and whether the candidate's own column does above it; only a tie in another column takes the slow
//     scan.  A maximum ORs its (2r+1)^2 coverage into a bit ring, one window row per lane;
//   * every thread keeps the `score > tau` and the maximum bits of its own 32 pixels of a chunk in two registers; one
//     chunk later, when the coverage of those rows is complete, it turns them into the two lists (maxima > tau;
//     uncovered pixels > tau) as 64-bit priority keys;
//   * list space is reserved per WARP in blocks of 16 / 32 entries with the next block always reserved ahead, so no
//     thread ever waits for an atomic it has just issued; unused entries of a warp's last blocks are written as
//     null keys (0), which sparse_kernel skips.
//
// Scores on this path are >= 0 (maps with a negative score are flagged for the round-faithful kernel), so the zero
// padding never beats a candidate.
#include "kb_sparse.cuh"

namespace kbsparse {
namespace {

constexpr int S = 8;                      // rows per chunk
constexpr int RING = 32;                  // map rows resident per CTA (4 chunks)
constexpr int MAX_NT = 320;               // threads per CTA = columns / 4 rounded up to a warp: W <= 1280
constexpr int BLK_M = 16, BLK_O = 32;     // list entries reserved per atomic
constexpr int MIN_BAND = 16;              // chunks per band below which the tiled kernel is the better choice (see launch_t)
constexpr unsigned FULL = 0xffffffffu;

// 16 / 4 bytes global -> shared, asynchronously; !valid writes zeros (the source is not read)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// rows S*chunk .. S*chunk+S-1 of the thread's four columns into the ring (zeros outside the map)
template <bool VEC>
__device__ __forceinline__ void issue_chunk(const float* __restrict__ img, float* raw, int chunk, int P, int x4, int H, int W) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
        const int row = S * chunk + i;
        const bool row_ok = row >= 0 && row < H;
        float* dst = raw + (row & (RING - 1)) * P + x4;
        const float* src = img + (size_t)(row_ok ? row : 0) * W + x4;
        if (VEC) {
            const bool ok = row_ok && x4 < W;
            cp_async16(dst, ok ? src : img, ok);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = row_ok && x4 + j < W;
                cp_async4(dst + j, ok ? src + j : img, ok);
            }
        }
    }
    cp_async_commit();
}

// Slow path of the first-of-ties rule: does an element of the window that comes EARLIER in raster order equal v?
// (Only reached when v occurs in another column of the window.)  The whole warp scans the R full rows above the
// centre and the R elements left of it.
template <int R>
__device__ __noinline__ bool earlier_equal_warp(const float* raw, int P, int x, int row, float v, int lane) {
    constexpr int W = 2 * R + 1, N_BEFORE = R * W + R;
    bool hit = false;
    for (int e = lane; e < N_BEFORE; e += 32) {
        const int dy = e / W - R, dx = e - (e / W) * W - R;
        const int xx = x + dx;
        if (xx >= 0 && xx < P) hit |= raw[((row + dy) & (RING - 1)) * P + xx] == v;
    }
    return __any_sync(FULL, hit);
}

// Per-warp writer of one list: entries go to `cur + fill`; `spare` (lane 0 only) is the next block, reserved when
// the warp moved into `cur`, so its atomic has long returned when it is needed.  fill == BLK: no current block.
struct ListStream {
    int cur, fill, spare;
};

// Appends the pixels selected by `bits` (bit 4*o + j = row o of the chunk, column j of the thread) of every lane:
// `pre` = entries of the lower lanes, `n` = entries of the whole warp (uniform).
template <int BLK>
__device__ __forceinline__ void append_bits(ListStream& s, uint64_t* list, int* cnt, uint32_t bits, int pre, int n,
                                            int lane, const float* raw, int P, int x4, int row0, int Wd) {
    int done = 0;
    while (done < n) {                              // uniform: n, fill and done are the same in every lane
        if (s.fill == BLK) {
            s.cur = __shfl_sync(FULL, s.spare, 0);
            s.fill = 0;
            if (lane == 0) s.spare = atomicAdd(cnt, BLK);
        }
        const int take = min(BLK - s.fill, n - done);
        int i = pre;
        uint32_t bb = bits;
        while (bb) {
            const int bit = __ffs(bb) - 1;
            bb &= bb - 1;
            if (i >= done && i < done + take) {
                const int slot = s.cur + s.fill + (i - done);
                const int row = row0 + (bit >> 2), x = x4 + (bit & 3);
                KB_ASSERT(slot >= 0 && row >= 0 && x < Wd);
                if (slot < LIST_CAP)
                    list[slot] = kb::priority_key(raw[(row & (RING - 1)) * P + x], (uint32_t)(row * Wd + x));
            }
            ++i;
        }
        s.fill += take;
        done += take;
    }
}

template <int BLK>
__device__ __forceinline__ void pad_stream(const ListStream& s, uint64_t* list, int lane) {
    if (s.fill < BLK)
        for (int i = s.fill + lane; i < BLK; i += 32)
            if (s.cur + i < LIST_CAP) list[s.cur + i] = 0ull;
    const int sp = __shfl_sync(FULL, s.spare, 0);
    for (int i = lane; i < BLK; i += 32)
        if (sp + i < LIST_CAP) list[sp + i] = 0ull;
}

template <int R, bool VEC>
__global__ void __launch_bounds__(MAX_NT, 1) round1_stream_kernel(SparseParams p, int cpm, int total_chunks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x, P = 4 * NT, WW = NT / 8, VP = P + 16;
    float* raw = reinterpret_cast<float*>(smem_raw);              // [RING][P]   map rows, thread t owns columns 4t..4t+3
    float* VM = raw + RING * P;                                   // [S][VP]     vertical window maxima of one chunk, 8 zero columns each side
    uint32_t* Cr = reinterpret_cast<uint32_t*>(VM + S * VP);      // [RING][WW]  coverage bits, 32 columns per word
    const int t = threadIdx.x, lane = t & 31, x4 = 4 * t;
    const int H = p.H, Wd = p.W;

    for (int i = t; i < S * 16; i += NT) {
        const int o = i >> 4, c = i & 15;
        VM[o * VP + (c < 8 ? c : P + c)] = 0.0f;
    }

    int g = (int)((long long)total_chunks * blockIdx.x / gridDim.x);
    const int g_end = (int)((long long)total_chunks * (blockIdx.x + 1) / gridDim.x);
    while (g < g_end) {
        // ---- one band: chunks c0 .. c1-1 of map b ------------------------------------------------------
        const int b = g / cpm, c0 = g - b * cpm, c1 = min(cpm, c0 + (g_end - g));
        g += c1 - c0;
        const float* __restrict__ img = p.score + (size_t)b * H * Wd;
        const float tau = p.tau[b];
        uint64_t* LM = p.listM + (size_t)b * LIST_CAP;
        uint64_t* LO = p.listO + (size_t)b * LIST_CAP;
        __syncthreads();                                          // the previous band is done with the rings
        issue_chunk<VEC>(img, raw, c0 - 2, P, x4, H, Wd);
        issue_chunk<VEC>(img, raw, c0 - 1, P, x4, H, Wd);
        issue_chunk<VEC>(img, raw, c0, P, x4, H, Wd);
        for (int i = t; i < RING * WW; i += NT) Cr[i] = 0u;
        ListStream sM{0, BLK_M, 0}, sO{0, BLK_O, 0};
        if (lane == 0) {
            sM.spare = atomicAdd(&p.cntM[b], BLK_M);
            sO.spare = atomicAdd(&p.cntO[b], BLK_O);
        }
        uint32_t negbits = 0u;
        uint32_t hot_prev = 0u, max_prev = 0u;                    // this thread's 8 x 4 pixels of the previous chunk
        // iteration k: maxima of chunk k (the band's chunks and one chunk either side), lists of chunk k-1
        for (int k = c0 - 1; k <= c1; ++k) {
            const bool active = (S * k + S > 0) && (S * k < H);
            float4 out[S];
            cp_async_wait_all();                                  // chunk k+1 (issued one iteration ago) has landed: own columns
            if (active) {
                const int row0 = S * k - R;                       // out[o] = max over ring rows row0+o .. row0+o+2R
                window_max_rows<R, S>([&](int i) {
                    return *reinterpret_cast<const float4*>(raw + ((row0 + i) & (RING - 1)) * P + x4);
                }, out);
#pragma unroll
                for (int o = 0; o < S; ++o) *reinterpret_cast<float4*>(VM + o * VP + 8 + x4) = out[o];
            }
            __syncthreads();                                      // row buffer and chunk k+1 complete; iteration k-1 finished everywhere
            if (k < c1) issue_chunk<VEC>(img, raw, k + 2, P, x4, H, Wd);      // over chunk k-2 (the last iteration needs none)
            Cr[((S * (k + 2)) & (RING - 1)) * WW + t] = 0u;       // the 8 coverage rows of chunk k+2: NT words
            uint32_t hot = 0u, mx = 0u;
            if (active) {
                // dense part, branch-free over the 8 rows: window maxima, `score > tau`, candidates (bit 4*o + j)
                uint32_t cand = 0u;
#pragma unroll
                for (int o = 0; o < S; ++o) {
                    const int slot = (S * k + o) & (RING - 1);
                    const float4 wm = window_max_cols<R>(VM + o * VP + 8, x4, out[o]);
                    const float4 v = *reinterpret_cast<const float4*>(raw + slot * P + x4);
                    negbits |= __float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w);
                    const uint32_t hn = (v.x > tau ? 1u : 0u) | (v.y > tau ? 2u : 0u) | (v.z > tau ? 4u : 0u) | (v.w > tau ? 8u : 0u);
                    const uint32_t cn = ((v.x == wm.x ? 1u : 0u) | (v.y == wm.y ? 2u : 0u) | (v.z == wm.z ? 4u : 0u) | (v.w == wm.w ? 8u : 0u)) & hn;
                    hot |= hn << (4 * o);
                    cand |= cn << (4 * o);
                }
                // candidates: rare (~1 pixel in (2R+1)^2); the whole warp takes one at a time
                unsigned pend = __ballot_sync(FULL, cand != 0u);
                while (pend) {
                    const int src = __ffs(pend) - 1;
                    pend &= pend - 1;
                    unsigned cb = __shfl_sync(FULL, cand, src);
                    const int xb = 4 * ((t & ~31) + src);
                    while (cb) {
                        const int bit = __ffs(cb) - 1;
                        cb &= cb - 1;
                        const int o = bit >> 2, x = xb + (bit & 3), row = S * k + o;
                        const float vv = raw[(row & (RING - 1)) * P + x];                 // (one address for the warp)
                        // lanes 0..2R: does column x-R+lane of the window hold vv (row buffer)?  lanes 2R+1..3R: does the
                        // own column hold it 1..R rows above?
                        bool e = false;
                        if (lane <= 2 * R) e = VM[o * VP + 8 + x + lane - R] == vv;
                        else if (lane <= 3 * R) e = raw[((row - (lane - 2 * R)) & (RING - 1)) * P + x] == vv;
                        const unsigned eq = __ballot_sync(FULL, e);
                        bool earlier;
                        if ((eq & ((2u << (2 * R)) - 1u)) == (1u << R)) earlier = (eq >> (2 * R + 1)) != 0u;
                        else earlier = earlier_equal_warp<R>(raw, P, x, row, vv, lane);
                        if (!earlier) {                           // a round-1 maximum: its bit, its coverage (one window row per lane)
                            if (lane == src) mx |= 1u << bit;
                            if (lane <= 2 * R) {
                                const int lo = max(x - R, 0), hi = min(x + R, P - 1);
                                const int w0 = lo >> 5, w1 = hi >> 5;
                                const uint32_t m0 = FULL << (lo & 31), m1 = FULL >> (31 - (hi & 31));
                                KB_ASSERT(w0 >= 0 && w1 < WW && x >= 0 && x < P);
                                uint32_t* crow = Cr + ((row + lane - R) & (RING - 1)) * WW;
                                if (w0 == w1) {
                                    atomicOr(&crow[w0], m0 & m1);
                                } else {
                                    atomicOr(&crow[w0], m0);
                                    atomicOr(&crow[w1], m1);
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();                                      // the coverage of chunk k-1 is complete
            const int j = k - 1;
            if (j >= c0 && j < c1) {
                uint32_t cov = 0u;
#pragma unroll
                for (int o = 0; o < S; ++o)
                    cov |= ((Cr[((S * j + o) & (RING - 1)) * WW + (t >> 3)] >> (x4 & 31)) & 0xfu) << (4 * o);
                const uint32_t emM = hot_prev & max_prev, emO = hot_prev & ~cov;
                const int mine = __popc(emM) | (__popc(emO) << 16);
                int inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += u;
                }
                const int tot = __shfl_sync(FULL, inc, 31);
                const int pre = inc - mine;
                if (tot & 0xffff) append_bits<BLK_M>(sM, LM, &p.cntM[b], emM, pre & 0xffff, tot & 0xffff, lane, raw, P, x4, S * j, Wd);
                if (tot >> 16) append_bits<BLK_O>(sO, LO, &p.cntO[b], emO, pre >> 16, tot >> 16, lane, raw, P, x4, S * j, Wd);
            }
            hot_prev = hot;
            max_prev = mx;
        }
        cp_async_wait_all();                                      // the last chunk issued is never used
        pad_stream<BLK_M>(sM, LM, lane);
        pad_stream<BLK_O>(sO, LO, lane);
        if (__any_sync(FULL, (negbits >> 31) != 0u) && lane == 0) atomicOr(&p.flags[b], 1);
    }
}

size_t stream_smem_bytes(int nt) {
    const size_t P = 4 * (size_t)nt, WW = nt / 8;
    return (RING * P + S * (P + 16)) * 4 + RING * WW * 4;
}

template <int R, bool VEC>
int launch_t(const SparseParams& p, int nt, int cpm, int total, bool force, cudaStream_t st) {
    // one device per process (torchrun: one rank per GPU): attributes and occupancy are looked up once per shape
    static int cached_nt = 0, cached_grid = 0;
    static bool configured = false;
    const size_t smem = stream_smem_bytes(nt);
    if (!configured) {
        KB_CUDA_TRY(cudaFuncSetAttribute(round1_stream_kernel<R, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        KB_CUDA_TRY(cudaFuncSetAttribute(round1_stream_kernel<R, VEC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    if (cached_nt != nt) {
        int dev = 0, sms = 0, occ = 0;
        KB_CUDA_TRY(cudaGetDevice(&dev));
        KB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        KB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, round1_stream_kernel<R, VEC>, nt, smem));
        if (occ < 1) return KB_ERR_UNSUPPORTED;
        cached_grid = occ * sms;
        cached_nt = nt;
    }
    // A band costs a vertical halo of up to four chunks and leaves up to 1.5 reserved blocks per warp and list
    // unused (null keys in the lists): bands are kept at MIN_BAND chunks or more, and when that leaves fewer
    // than 64 CTAs (a handful of maps) the tiled kernel, whose grid is tiles x maps, is the better choice.
    int grid = total / MIN_BAND;
    if (grid > cached_grid) grid = cached_grid;
    if (force && grid < 1) grid = 1;
    if (!force && grid < 64) return KB_ERR_UNSUPPORTED;
    round1_stream_kernel<R, VEC><<<grid, nt, smem, st>>>(p, cpm, total);
    KB_LAUNCH_CHECK();
    return KB_OK;
}

template <int R>
int launch_r(const SparseParams& p, int nt, int cpm, int total, bool vec, bool force, cudaStream_t st) {
    return vec ? launch_t<R, true>(p, nt, cpm, total, force, st) : launch_t<R, false>(p, nt, cpm, total, force, st);
}

}  // namespace

int launch_round1_stream(const SparseParams& p, bool force, cudaStream_t st) {
    if (p.W > 4 * MAX_NT || p.r < 1 || p.r > 8) return KB_ERR_UNSUPPORTED;
    const int nt = 32 * ((p.W + 127) / 128);
    const int cpm = (p.H + S - 1) / S;
    const long long total = (long long)p.B * cpm;
    if (total > 0x3fffffff) return KB_ERR_UNSUPPORTED;
    const bool vec = (p.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.score) & 15u) == 0);
    switch (p.r) {
        case 1: return launch_r<1>(p, nt, cpm, (int)total, vec, force, st);
        case 2: return launch_r<2>(p, nt, cpm, (int)total, vec, force, st);
        case 3: return launch_r<3>(p, nt, cpm, (int)total, vec, force, st);
        case 4: return launch_r<4>(p, nt, cpm, (int)total, vec, force, st);
        case 5: return launch_r<5>(p, nt, cpm, (int)total, vec, force, st);
        case 6: return launch_r<6>(p, nt, cpm, (int)total, vec, force, st);
        case 7: return launch_r<7>(p, nt, cpm, (int)total, vec, force, st);
        case 8: return launch_r<8>(p, nt, cpm, (int)total, vec, force, st);
        default: return KB_ERR_UNSUPPORTED;
    }
}

}  // namespace kbsparse
