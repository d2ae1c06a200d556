// Packed streaming form of round 1 of the sparse exact detection path (contract of the two lists: kb_sparse_nms.cu).
//
// The fp32 round-1 kernels (tiled: kb_sparse_nms.cu, streaming: kb_round1_stream.cu) spend ~50-100 thread-instructions
// per pixel finding the pixels that win their (2r+1)^2 window (the reference's first fast_nms round,
// utils/extracter.py:54-70).  Almost all of that work only has to establish "some pixel of the window is larger", for
// which a MONOTONE 16-bit image of the score is enough: q = fp16(max((score - tau) * qscale, 0)), one FFMA and one
// round-to-nearest conversion, so u >= v implies q(u) >= q(v); the subtraction spends fp16's 11 bits on the range the
// candidates live in and qscale (a power of two from tau_kernel: the largest sampled score lands at 2^15) keeps the
// image inside fp16's range whatever the scale of the scores.  This kernel therefore
//
//   * processes TWO maps at once: a CTA owns full-width bands of 8-row chunks of a PAIR of maps; thread t owns columns
//     4t..4t+3 of both; every value it keeps on chip is a half2 (low half = map 2i, high half = map 2i+1), so each
//     maximum instruction (HMNMX2) and each shared-memory word serves two pixels with no cross-lane traffic between
//     the halves;
//   * reads the fp32 rows once, straight into registers (16-byte loads issued half a chunk ahead), takes `score > tau`
//     in fp32 there (exact, one bit per pixel kept in registers), packs, and stores the packed rows into a 32-row ring;
//   * per chunk: separable (2r+1)-window maximum of the packed rows (vertical from the ring rows of the thread's own
//     columns, horizontal from the neighbours' 16-byte pieces of a row buffer); a pixel is a CANDIDATE iff it is above
//     tau and its packed value equals the packed window maximum.  A true round-1 maximum always is one; a candidate is
//     a true maximum unless another pixel of the window has the same packed value, which the owning thread finds out
//     by scanning the 2r other columns of the row buffer and the 2r other rows of its own ring column (16-bit reads);
//     only when such a 16-bit tie exists are fp32 values fetched (from L2) and compared with the first-of-ties rule of
//     torch.argmax (extracter.py:69-70: strictly larger than everything earlier in raster order, >= everything later);
//   * keeps all per-pixel bits in "patch order": a thread's 8 rows x 4 columns of a chunk are one 32-bit word per map
//     (bit 4*row + column).  The coverage of the maxima (every pixel within r of one) is a dilation of those words:
//     vertically by funnel shifts over the thread's own words of three consecutive chunks, horizontally by nibble-wise
//     prefix / suffix ORs of the five neighbouring threads' words (exchanged through shared memory);
//   * emits the two lists (maxima > tau; uncovered pixels > tau) two chunks behind the maxima search: counts are
//     scanned per warp, list space is reserved with one atomic per warp, list and chunk whose result is not needed
//     until the packed rows of the next chunk have been processed, and the fp32 scores of the listed pixels (3 % of
//     the map) are fetched from L2 the same way -- issued first, consumed after the vertical pass.
//
// Scores on this path are >= 0 (maps with a negative score are flagged for the round-faithful kernel), so the zero
// padding never beats a candidate.
#include "kb_sparse.cuh"
#include <type_traits>

namespace kbsparse {
namespace {

constexpr int S = 8;                      // rows per chunk
constexpr int RING = 32;                  // packed map rows resident per CTA (4 chunks)
constexpr int PADC = 8;                   // zero columns either side of the ring / row buffer rows
constexpr int MAX_NT = 320;               // threads per CTA = columns / 4 rounded up to a warp: W <= 1280
constexpr int NPRE = 4;                   // listed pixels per thread, map and chunk whose score is fetched ahead
constexpr int MIN_BAND = 8;               // chunks per band below which the tiled kernel is the better choice
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t hmax2u(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint4 qmax(uint4 a, uint4 b) {
    return make_uint4(hmax2u(a.x, b.x), hmax2u(a.y, b.y), hmax2u(a.z, b.z), hmax2u(a.w, b.w));
}
// {lo, hi} -> half2 bits of max(x, 0), round to nearest even (monotone)
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// 0xffff in every half where a == b
__device__ __forceinline__ uint32_t heq2_mask(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// Split-phase CTA barriers (mbarrier): a thread ARRIVES when its writes are done and WAITS only where it needs the other
// threads' data, doing list / store work that touches no shared data of the others in between, so a warp that is late
// (a few candidates more) does not idle the rest of the CTA.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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

template <bool VEC>
__device__ __forceinline__ float4 load_row4(const float* __restrict__ img, int row, int x4, int H, int W) {   // any row
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= 0 && row < H) {
        const float* src = img + (size_t)row * W + x4;
        if (VEC) {
            if (x4 < W) v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
            if (x4 < W) v.x = __ldg(src);
            if (x4 + 1 < W) v.y = __ldg(src + 1);
            if (x4 + 2 < W) v.z = __ldg(src + 2);
            if (x4 + 3 < W) v.w = __ldg(src + 3);
        }
    }
    return v;
}

// word |= bit where a > b: one FSETP and one predicated LOP3 per pixel (`set.gt.u32.f32` is lowered to FSETP + SEL, and
// the selects of a C++ conditional cost a third instruction)
#define KB_HOT_BIT(word, a, b, bit) \
    asm("{ .reg .pred p; setp.gt.f32 p, %1, %2; @p or.b32 %0, %0, %3; }" : "+r"(word) : "f"(a), "f"(b), "n"(bit))
// bits 4*O .. 4*O+3 of a patch word: score > tau of row O
template <int O>
__device__ __forceinline__ void hot4(uint32_t& word, float4 v, float tau) {
    KB_HOT_BIT(word, v.x, tau, 1u << (4 * O));
    KB_HOT_BIT(word, v.y, tau, 2u << (4 * O));
    KB_HOT_BIT(word, v.z, tau, 4u << (4 * O));
    KB_HOT_BIT(word, v.w, tau, 8u << (4 * O));
}

// (2R+1)-window maximum along the row for a thread's 4 columns (packed pairs); vmrow points at column 0 of a row buffer
// row that is readable (and zero) for 8 columns either side of the data; `own` = the row's values at x4..x4+3.
template <int R>
__device__ __forceinline__ uint4 window_max_cols_q(const uint32_t* vmrow, int x4, uint4 own) {
    constexpr int NBR = (R + 3) / 4, C = 4 * NBR;
    uint32_t a[4 * (2 * NBR + 1)];
#pragma unroll
    for (int nb = -NBR; nb <= NBR; ++nb) {
        const uint4 q = nb == 0 ? own : *reinterpret_cast<const uint4*>(vmrow + x4 + 4 * nb);
        a[C + 4 * nb + 0] = q.x; a[C + 4 * nb + 1] = q.y; a[C + 4 * nb + 2] = q.z; a[C + 4 * nb + 3] = q.w;
    }
    uint4 r;
    if constexpr (R >= 2) {
        uint32_t core = a[C + 3 - R];
#pragma unroll
        for (int i = C + 4 - R; i <= C + R; ++i) core = hmax2u(core, a[i]);
        const uint32_t l2 = a[C + 2 - R], l1 = hmax2u(l2, a[C + 1 - R]), l0 = hmax2u(l1, a[C - R]);
        const uint32_t r1 = a[C + R + 1], r2 = hmax2u(r1, a[C + R + 2]), r3 = hmax2u(r2, a[C + R + 3]);
        r.x = hmax2u(core, l0);
        r.y = hmax2u(hmax2u(core, l1), r1);
        r.z = hmax2u(hmax2u(core, l2), r2);
        r.w = hmax2u(core, r3);
    } else {
        r.x = hmax2u(hmax2u(a[C - 1], a[C]), a[C + 1]);
        r.y = hmax2u(hmax2u(a[C], a[C + 1]), a[C + 2]);
        r.z = hmax2u(hmax2u(a[C + 1], a[C + 2]), a[C + 3]);
        r.w = hmax2u(hmax2u(a[C + 2], a[C + 3]), a[C + 4]);
    }
    return r;
}

// ---- patch-order bit words: bit 4*row + column of a thread's 8 x 4 pixels of one chunk and map ---------------------
// rows: out row o = OR of rows o-R .. o+R over the words of the previous / this / the next chunk
template <int R>
__device__ __forceinline__ uint32_t vdilate(uint32_t prev, uint32_t cur, uint32_t next) {
    uint32_t acc = cur;
#pragma unroll
    for (int d = 1; d <= R; ++d) {
        acc |= (4 * d < 32) ? __funnelshift_r(prev, cur, 32 - 4 * d) : prev;
        acc |= (4 * d < 32) ? __funnelshift_r(cur, next, 4 * d) : next;
    }
    return acc;
}
__device__ __forceinline__ uint32_t prefix_or4(uint32_t x) {      // bit s of a nibble = OR of its bits 0..s
    x |= (x << 1) & 0xEEEEEEEEu;
    x |= (x << 2) & 0xCCCCCCCCu;
    return x;
}
__device__ __forceinline__ uint32_t suffix_or4(uint32_t x) {      // bit s of a nibble = OR of its bits s..3
    x |= (x >> 1) & 0x77777777u;
    x |= (x >> 2) & 0x33333333u;
    return x;
}
// Source columns s of the neighbour n threads away that lie within R of own column j: |4n + s - j| <= R, a range that
// touches column 0 (n > 0: prefix OR, bit hi) or column 3 (n < 0: suffix OR, bit lo).  mask(n, sh, ...) = the nibble bits
// b of that prefix / suffix word that serve an own column j = b + sh.
template <int R>
__host__ __device__ constexpr uint32_t hd_mask(int n, int sh, bool prefix) {
    uint32_t m = 0;
    for (int j = 0; j < 4; ++j) {
        int lo = j - 4 * n - R, hi = j - 4 * n + R;
        lo = lo < 0 ? 0 : lo;
        hi = hi > 3 ? 3 : hi;
        if (lo > hi) continue;
        const bool use_prefix = (n > 0) || (n == 0 && lo == 0);
        if (use_prefix != prefix) continue;
        const int b = use_prefix ? hi : lo;
        if (j - b == sh) m |= 1u << b;
    }
    return m * 0x11111111u;
}
// columns: out column j of thread t = OR of columns within R, N[n + 2] = word of thread t + n
template <int R>
__device__ __forceinline__ uint32_t hdilate(const uint32_t (&N)[5]) {
    uint32_t cov = 0u;
#pragma unroll
    for (int n = -2; n <= 2; ++n) {
        const uint32_t PX = prefix_or4(N[n + 2]), SX = suffix_or4(N[n + 2]);      // the unused one is dead code
        if (n == 0 && R >= 3) {                                   // the thread's own columns all see each other
            cov |= ((PX >> 3) & 0x11111111u) * 15u;
            continue;
        }
#pragma unroll
        for (int sh = -3; sh <= 3; ++sh) {
            const uint32_t mp = hd_mask<R>(n, sh, true), ms = hd_mask<R>(n, sh, false);
            if (mp) cov |= sh >= 0 ? (PX & mp) << (sh >= 0 ? sh : 0) : (PX & mp) >> (sh < 0 ? -sh : 0);
            if (ms) cov |= sh >= 0 ? (SX & ms) << (sh >= 0 ? sh : 0) : (SX & ms) >> (sh < 0 ? -sh : 0);
        }
    }
    return cov;
}

// ---- exact verdict on a candidate -------------------------------------------------------------------------------
// (row, x) of map half h (0 / 1) is above tau and its packed value equals the packed window maximum.  It is a round-1
// maximum unless another pixel of the window shares its packed value AND beats it in fp32 (larger, or equal and
// earlier in raster order).
template <int R>
__device__ __forceinline__ bool tie_verdict(const uint32_t* raw, int VP, uint32_t tie_cols, uint32_t tie_rows, uint32_t qv, float v,
                                         int row, int x, int h, const float* __restrict__ img, int H, int W) {
    const uint16_t* rw16 = reinterpret_cast<const uint16_t*>(raw) + h;
    // own column
    while (tie_rows) {
        const int dy = __ffs(tie_rows) - 1 - R;
        tie_rows &= tie_rows - 1;
        const int yy = row + dy;
        if (yy < 0 || yy >= H) continue;                      // zero padding: 0 < v
        const float u = __ldg(img + (size_t)yy * W + x);
        if (u > v || (u == v && dy < 0)) return false;
    }
    while (tie_cols) {
        const int dx = __ffs(tie_cols) - 1 - R;
        tie_cols &= tie_cols - 1;
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        for (int dy = -R; dy <= R; ++dy) {
            const int yy = row + dy;
            if (rw16[2 * ((yy & (RING - 1)) * VP + PADC + xx)] != qv) continue;
            if (yy < 0 || yy >= H) continue;
            const float u = __ldg(img + (size_t)yy * W + xx);
            if (u > v || (u == v && (dy < 0 || (dy == 0 && dx < 0)))) return false;
        }
    }
    return true;
}

// Fast part of the verdict: does any OTHER pixel of the candidate's row-buffer row (2R columns) or of its own ring
// column (2R rows) share its packed value?  own = the candidate's word in the ring (chunk k, row o of the chunk), vrow =
// its word in the row buffer; fix_prev / fix_next = what to add to own + d * VP when row o + d lies in chunk k-1 / k+1 and
// that chunk's ring slot is not adjacent (the ring wraps every four chunks).  One load, one LOP3 (xor + half mask) and one
// unsigned minimum per element: the minimum is 0 iff some element ties.
template <int R>
__device__ __forceinline__ bool is_round1_max(const uint32_t* raw, const uint32_t* own, const uint32_t* vrow, int VP, int fix_prev,
                                              int fix_next, int o, int row, int x, int h, const float* __restrict__ img, int H, int W) {
    const uint32_t hmask = h ? 0xffff0000u : 0x0000ffffu;
    const uint32_t qv2 = __byte_perm(*own, 0u, h ? 0x3232 : 0x1010);            // the candidate's value in both halves
    // the candidate's fp32 score, needed only on a 16-bit tie: fetched now (volatile: not sunk into the rare branch) so
    // that its L2 latency runs under the scan instead of in front of the verdict
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(img + (size_t)row * W + x));
    uint32_t mn4[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};    // four independent chains
#pragma unroll
    for (int d = -R; d <= R; ++d) {
        if (d == 0) continue;
        mn4[(d + R) & 1] = min(mn4[(d + R) & 1], (vrow[d] ^ qv2) & hmask);
        const uint32_t* a = own + d * VP;
        if (d < 0 && o + d < 0) a += fix_prev;
        if (d > 0 && o + d >= S) a += fix_next;
        mn4[2 + ((d + R) & 1)] = min(mn4[2 + ((d + R) & 1)], (*a ^ qv2) & hmask);
    }
    const uint32_t mn_cols = min(mn4[0], mn4[1]), mn_rows = min(mn4[2], mn4[3]);
    if (min(mn_cols, mn_rows) != 0u) return true;
    // a 16-bit tie (rare): which columns of the row buffer / which rows of the own column, then fp32
    const uint32_t qv = qv2 & 0xffffu;
    uint32_t tie_cols = 0u, tie_rows = 0u;
    if (mn_cols == 0u) {
#pragma unroll
        for (int d = -R; d <= R; ++d)
            if (d != 0 && ((vrow[d] ^ qv2) & hmask) == 0u) tie_cols |= 1u << (d + R);
    }
    if (mn_rows == 0u) {
#pragma unroll
        for (int d = -R; d <= R; ++d) {
            if (d == 0) continue;
            const uint32_t* a = own + d * VP;
            if (d < 0 && o + d < 0) a += fix_prev;
            if (d > 0 && o + d >= S) a += fix_next;
            if (((*a ^ qv2) & hmask) == 0u) tie_rows |= 1u << (d + R);
        }
    }
    return tie_verdict<R>(raw, VP, tie_cols, tie_rows, qv, v, row, x, h, img, H, W);
}

// NTC: threads per CTA as a compile-time constant (0 = blockDim.x): with it the row pitch of the shared arrays is an
// immediate and almost all shared-memory addresses are register + constant.
// WC: the map width as a compile-time constant (0 = p.W): global row offsets become immediates as well.
template <int R, bool VEC, int NTC, int WC>
__global__ void __launch_bounds__(NTC ? NTC : MAX_NT, (NTC && NTC <= 160) ? 2 : 1) round1_packed_kernel(SparseParams p, int cpm, int total_chunks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = NTC ? NTC : (int)blockDim.x, P = 4 * NT, VP = P + 2 * PADC, XP = NT + 4;
    uint32_t* raw = reinterpret_cast<uint32_t*>(smem_raw);        // [RING][VP] packed map rows, data at column PADC
    uint32_t* VM = raw + RING * VP;                               // [S][VP]    packed vertical window maxima of one chunk
    uint32_t* XB = VM + S * VP;                                   // [2][XP]    vertically dilated maxima words, 2 pad entries each side
    const int t = threadIdx.x, lane = t & 31, x4 = 4 * t;
    const int H = p.H, Wd = WC ? WC : p.W;
    uint32_t* const rawc = raw + PADC + x4;                       // the thread's own columns: ring ...
    uint32_t* const VMc = VM + PADC + x4;                         // ... and row buffer
    auto slot = [&](int chunk) { return rawc + ((chunk & 3) * S) * VP; };      // ring row 0 of a chunk
    const bool col_ok = (WC && NTC && 4 * NTC <= WC) ? true : x4 < Wd;
    __shared__ __align__(8) uint64_t s_bar[2];                    // [0] row buffer complete, [1] chunk done everywhere
    const uint32_t bar_rows = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
    const uint32_t bar_done = (uint32_t)__cvta_generic_to_shared(&s_bar[1]);
    uint32_t par_rows = 0u, par_done = 0u;
    if (t == 0) {
        mbar_init(bar_rows, NT);
        mbar_init(bar_done, NT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    for (int i = t; i < (RING + S) * 2 * PADC; i += NT) {         // (raw and VM are contiguous)
        const int r = i / (2 * PADC), c = i - r * (2 * PADC);
        raw[r * VP + (c < PADC ? c : P + c)] = 0u;
    }
    if (t < 8) XB[(t >> 2) * XP + ((t & 3) < 2 ? (t & 3) : NT + (t & 3))] = 0u;

    int g = (int)((long long)total_chunks * blockIdx.x / gridDim.x);
    const int g_end = (int)((long long)total_chunks * (blockIdx.x + 1) / gridDim.x);
    while (g < g_end) {
        // ---- one band: chunks c0 .. c1-1 of the map pair pr ---------------------------------------------
        const int pr = g / cpm, c0 = g - pr * cpm, c1 = min(cpm, c0 + (g_end - g));
        g += c1 - c0;
        const int bm[2] = {2 * pr, min(2 * pr + 1, p.B - 1)};
        const bool has1 = 2 * pr + 1 < p.B;
        const int Hm[2] = {H, has1 ? H : 0};                       // an odd batch: the last pair's second half is empty
        const size_t map_stride4 = ((size_t)H * Wd) >> 2;          // (VEC: Wd % 4 == 0) float4s between the maps of a pair
        const float* __restrict__ img[2] = {p.score + (size_t)bm[0] * H * Wd, p.score + (size_t)bm[1] * H * Wd};
        const float tau[2] = {p.tau[bm[0]], p.tau[bm[1]]};
        // 16-bit image: fp16(max((score - tau) * qs, 0)) as one FFMA per pixel (monotone in the score: one rounding)
        const float qs[2] = {p.qscale[bm[0]], p.qscale[bm[1]]};
        const float qc[2] = {-tau[0] * qs[0], -tau[1] * qs[1]};
        uint64_t* LM[2] = {p.listM + (size_t)bm[0] * LIST_CAP, p.listM + (size_t)bm[1] * LIST_CAP};
        uint64_t* LO[2] = {p.listO + (size_t)bm[0] * LIST_CAP, p.listO + (size_t)bm[1] * LIST_CAP};

        // per-pixel bits of this thread's patches, one word per map and chunk:
        // hot[m][i]: score > tau of chunk k-2+i (i = 4: the chunk being loaded); mx[m][i]: round-1 maxima of chunk k-2+i
        uint32_t hot[2][5] = {{0u, 0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u, 0u}};
        uint32_t mx[2][3] = {{0u, 0u, 0u}, {0u, 0u, 0u}};
        uint32_t neg0 = 0u, neg1 = 0u;                              // sign bits of everything read

        auto store_row = [&](auto O, uint32_t* dst, const float4& a, const float4& b, uint32_t& h0, uint32_t& h1) {
            constexpr int o = decltype(O)::value;
            hot4<o>(h0, a, tau[0]);
            hot4<o>(h1, b, tau[1]);
            // the 16-bit image is taken of (score - tau) * qscale (monotone; pixels at or below tau become 0): fp16 spends
            // its 11 bits on the range the candidates live in, so two pixels of a window rarely share the largest packed
            // value, and qscale (tau_kernel: the largest sample at 2^15) keeps it inside fp16's range whatever the scores' scale
            const uint4 q = make_uint4(pack2(fmaf(a.x, qs[0], qc[0]), fmaf(b.x, qs[1], qc[1])), pack2(fmaf(a.y, qs[0], qc[0]), fmaf(b.y, qs[1], qc[1])),
                                       pack2(fmaf(a.z, qs[0], qc[0]), fmaf(b.z, qs[1], qc[1])), pack2(fmaf(a.w, qs[0], qc[0]), fmaf(b.w, qs[1], qc[1])));
            neg0 |= __float_as_uint(a.x) | __float_as_uint(a.y) | __float_as_uint(a.z) | __float_as_uint(a.w);
            neg1 |= __float_as_uint(b.x) | __float_as_uint(b.y) | __float_as_uint(b.z) | __float_as_uint(b.w);
            *reinterpret_cast<uint4*>(dst + o * VP) = q;
        };
        // rows o0 .. o0+3 of a chunk (o0 = 0 or 4): `score > tau` bits, packed pairs into the ring
        auto store_rows = [&](int chunk, auto O0, const float4 (&f0)[S / 2], const float4 (&f1)[S / 2], uint32_t& h0, uint32_t& h1) {
            constexpr int o0 = decltype(O0)::value;
            uint32_t* dst = slot(chunk);
            store_row(std::integral_constant<int, o0>{}, dst, f0[0], f1[0], h0, h1);
            store_row(std::integral_constant<int, o0 + 1>{}, dst, f0[1], f1[1], h0, h1);
            store_row(std::integral_constant<int, o0 + 2>{}, dst, f0[2], f1[2], h0, h1);
            store_row(std::integral_constant<int, o0 + 3>{}, dst, f0[3], f1[3], h0, h1);
        };
        auto load_rows = [&](int chunk, int o0, float4 (&f0)[S / 2], float4 (&f1)[S / 2]) {
            const int r0 = S * chunk + o0;
            if (VEC && r0 >= 0 && r0 + S / 2 <= H && has1) {       // (uniform) all four rows inside both maps
                const float4* src0 = reinterpret_cast<const float4*>(img[0] + (size_t)r0 * Wd + x4);
                const float4* src1 = src0 + map_stride4;
                const int w4 = Wd >> 2;
#pragma unroll
                for (int i = 0; i < S / 2; ++i) {
                    f0[i] = f1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (col_ok) {
                        f0[i] = __ldg(src0 + i * w4);
                        f1[i] = __ldg(src1 + i * w4);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < S / 2; ++i) {
                    f0[i] = load_row4<VEC>(img[0], r0 + i, x4, Hm[0], Wd);
                    f1[i] = load_row4<VEC>(img[1], r0 + i, x4, Hm[1], Wd);
                }
            }
        };

        __syncthreads();                                          // the previous band is done with the ring
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {                          // chunks c0-2, c0-1, c0
            float4 fa0[S / 2], fa1[S / 2], fb0[S / 2], fb1[S / 2];
            load_rows(c0 - 2 + cc, 0, fa0, fa1);
            load_rows(c0 - 2 + cc, S / 2, fb0, fb1);
            store_rows(c0 - 2 + cc, std::integral_constant<int, 0>{}, fa0, fa1, hot[0][1 + cc], hot[1][1 + cc]);
            store_rows(c0 - 2 + cc, std::integral_constant<int, S / 2>{}, fb0, fb1, hot[0][1 + cc], hot[1][1 + cc]);
        }
        XB[2 + t] = 0u;
        XB[XP + 2 + t] = 0u;
        __syncthreads();
        mbar_arrive(bar_done);                                    // (the first iteration's wait)

        // lists, part 2: the state part 1 left one iteration ago
        bool q_valid = false;
        int q_j = 0, q_base = 0, q_pre[2] = {0, 0}, q_tot[2] = {0, 0};
        uint32_t q_both[2] = {0u, 0u}, q_emM[2] = {0u, 0u};
        float q_sc[2][NPRE];
        auto write_lists = [&]() {
            if (!q_valid) return;
            q_valid = false;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (q_tot[m] == 0) continue;                      // (warp-uniform)
                int offM = __shfl_sync(FULL, q_base, 2 * m) + (q_pre[m] & 0xffff);
                int offO = __shfl_sync(FULL, q_base, 2 * m + 1) + (q_pre[m] >> 16);
                uint32_t bb = q_both[m];
                auto put = [&](int bit, float score) {
                    const int row = S * q_j + (bit >> 2), x = x4 + (bit & 3);
                    KB_ASSERT(row >= 0 && row < H && x < Wd);
                    // (listed scores are > tau >= 0: their order key is the bit pattern with the top bit set)
                    const uint64_t key = ((uint64_t)(__float_as_uint(score) | 0x80000000u) << 32) |
                                         (uint64_t)(0xffffffffu - (uint32_t)(row * Wd + x));
                    if ((q_emM[m] >> bit) & 1u) { if (offM < LIST_CAP) LM[m][offM] = key; ++offM; }
                    else { if (offO < LIST_CAP) LO[m][offO] = key; ++offO; }
                };
#pragma unroll
                for (int i = 0; i < NPRE; ++i) {
                    if (bb) {
                        const int bit = __ffs(bb) - 1;
                        bb &= bb - 1;
                        put(bit, q_sc[m][i]);
                    }
                }
                while (bb) {                                      // more than NPRE listed pixels in one patch: rare
                    const int bit = __ffs(bb) - 1;
                    bb &= bb - 1;
                    put(bit, __ldg(img[m] + (unsigned)((S * q_j + (bit >> 2)) * Wd + x4 + (bit & 3))));
                }
            }
        };

        // iteration k: maxima of chunk k, coverage words of chunk k-1, lists of chunk k-2, rows of chunk k+2
        for (int k = c0 - 1; k <= c1 + 1; ++k) {
            // ---- rows of chunk k+2, first half, on their way -------------------------------------------------
            const bool pf = k + 2 <= c1 + 1;
            float4 f0[S / 2], f1[S / 2];
            if (pf) load_rows(k + 2, 0, f0, f1);
            mbar_wait(bar_done, par_done);                        // chunk k-1 is done everywhere: row buffer, ring slot and
            par_done ^= 1u;                                       // the coverage words may be reused / read

            // ---- lists of chunk j = k-2, part 1: what to list, where, and the scores on their way ---------------------
            const int j = k - 2;
            const bool listing = j >= c0;
            uint32_t both[2] = {0u, 0u}, emM[2] = {0u, 0u};
            int pre[2] = {0, 0}, tot[2] = {0, 0}, base = 0;
            float sc[2][NPRE];
            if (listing) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    uint32_t N[5];
#pragma unroll
                    for (int n = 0; n < 5; ++n) N[n] = XB[m * XP + t + n];
                    const uint32_t cov = hdilate<R>(N);
                    emM[m] = hot[m][0] & mx[m][0];
                    both[m] = emM[m] | (hot[m][0] & ~cov);
                    const int mine = __popc(emM[m]) | (__popc(both[m] & ~emM[m]) << 16);
                    int inc = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int u = __shfl_up_sync(FULL, inc, d);
                        if (lane >= d) inc += u;
                    }
                    tot[m] = __shfl_sync(FULL, inc, 31);
                    pre[m] = inc - mine;
                }
                if (lane < 4) {                                   // lane 2m: maxima list of map m, lane 2m+1: the other list
                    const int tt = (lane & 2) ? tot[1] : tot[0], bb = (lane & 2) ? bm[1] : bm[0];
                    const int n = (lane & 1) ? (tt >> 16) : (tt & 0xffff);
                    if (n) base = atomicAdd((lane & 1) ? &p.cntO[bb] : &p.cntM[bb], n);
                }
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    uint32_t bb = both[m];
#pragma unroll
                    for (int i = 0; i < NPRE; ++i) {
                        sc[m][i] = 0.0f;
                        if (bb) {
                            const int bit = __ffs(bb) - 1;
                            bb &= bb - 1;
                            sc[m][i] = __ldg(img[m] + (unsigned)((S * j + (bit >> 2)) * Wd + x4 + (bit & 3)));
                        }
                    }
                }
            }

            // ---- vertical window maxima of chunk k ----------------------------------------------------------
            const bool active = k <= c1 && (S * k + S > 0) && (S * k < H);
            uint4 out[S];
            if (active) {
                // out[o] = max over map rows S*k-R+o .. S*k+R+o: input row i lies in chunk k-1 + (i+S-R)/S
                const uint32_t* cb[3] = {slot(k - 1), slot(k), slot(k + 1)};
                window_max_rows_t<R, S, uint4>([&](int i) {
                    return *reinterpret_cast<const uint4*>(cb[(i + S - R) / S] + ((i + S - R) % S) * VP);
                }, [](uint4 a, uint4 b) { return qmax(a, b); }, out);
#pragma unroll
                for (int o = 0; o < S; ++o) *reinterpret_cast<uint4*>(VMc + o * VP) = out[o];
            }
            if (pf) {
                store_rows(k + 2, std::integral_constant<int, 0>{}, f0, f1, hot[0][4], hot[1][4]);
                load_rows(k + 2, S / 2, f0, f1);
            }
            mbar_arrive(bar_rows);                                // my part of the row buffer is written, the list words are read

            // ---- lists, part 2, of the chunk part 1 prepared ONE ITERATION AGO (global memory only: in the shadow of the
            //      barrier; the atomics and score loads of that part 1 have had a whole iteration to land) ------------
            write_lists();
            if (listing) {
                q_valid = true; q_j = j; q_base = base;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    q_both[m] = both[m]; q_emM[m] = emM[m]; q_pre[m] = pre[m]; q_tot[m] = tot[m];
#pragma unroll
                    for (int i = 0; i < NPRE; ++i) q_sc[m][i] = sc[m][i];
                }
            }

            mbar_wait(bar_rows, par_rows);                        // row buffer complete
            par_rows ^= 1u;

            // ---- horizontal window maxima, candidates of chunk k ----------------------------------------------
            uint32_t cand[2] = {0u, 0u};
            if (active) {
                uint32_t lo4 = 0u, hi4 = 0u;                      // rows 0..3 / 4..7: low half = map 0, high half = map 1
                const uint32_t* cur = slot(k);
#pragma unroll
                for (int o = 0; o < S; ++o) {
                    const uint4 wm = window_max_cols_q<R>(VM + o * VP + PADC, x4, out[o]);
                    const uint4 v = *reinterpret_cast<const uint4*>(cur + o * VP);
                    const uint32_t c1b = 0x00010001u << (4 * (o & 3));
                    uint32_t e = heq2_mask(v.x, wm.x) & c1b;
                    e |= heq2_mask(v.y, wm.y) & (c1b << 1);
                    e |= heq2_mask(v.z, wm.z) & (c1b << 2);
                    e |= heq2_mask(v.w, wm.w) & (c1b << 3);
                    if (o < 4) lo4 |= e; else hi4 |= e;
                }
                cand[0] = __byte_perm(lo4, hi4, 0x5410) & hot[0][2];
                cand[1] = __byte_perm(lo4, hi4, 0x7632) & hot[1][2];
            }
            // ---- candidates -> round-1 maxima (rare: ~1 pixel in (2R+1)^2), both maps in one loop ---------------
            {
                uint32_t ca = cand[0], cb2 = cand[1];
                const uint32_t* cur = slot(k);
                const int fix_prev = (k & 3) == 0 ? RING * VP : 0, fix_next = (k & 3) == 3 ? -RING * VP : 0;
                while (ca | cb2) {
                    const int m = ca ? 0 : 1;
                    uint32_t& cc = ca ? ca : cb2;
                    const int bit = __ffs(cc) - 1;
                    cc &= cc - 1;
                    const int o = bit >> 2, jj = bit & 3;
                    if (is_round1_max<R>(raw, cur + o * VP + jj, VMc + o * VP + jj, VP, fix_prev, fix_next, o, S * k + o, x4 + jj, m,
                                         m ? img[1] : img[0], H, Wd)) {
                        if (m) mx[1][2] |= 1u << bit; else mx[0][2] |= 1u << bit;
                    }
                }
            }
            // ---- coverage words of chunk k-1 (rows), for the neighbours ----------------------------------------
            XB[2 + t] = vdilate<R>(mx[0][0], mx[0][1], mx[0][2]);
            XB[XP + 2 + t] = vdilate<R>(mx[1][0], mx[1][1], mx[1][2]);
            mbar_arrive(bar_done);                                // chunk k done here
            // (rows 4..7 of chunk k+2 go to a ring slot nobody reads before the next row-buffer barrier)
            if (pf) store_rows(k + 2, std::integral_constant<int, S / 2>{}, f0, f1, hot[0][4], hot[1][4]);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                hot[m][0] = hot[m][1]; hot[m][1] = hot[m][2]; hot[m][2] = hot[m][3]; hot[m][3] = hot[m][4]; hot[m][4] = 0u;
                mx[m][0] = mx[m][1]; mx[m][1] = mx[m][2]; mx[m][2] = 0u;
            }
        }
        write_lists();                                            // the last chunk of the band
        mbar_wait(bar_done, par_done);                            // (every arrival is waited for: the parity stays in step)
        par_done ^= 1u;
        if (__any_sync(FULL, (neg0 >> 31) != 0u) && lane == 0) atomicOr(&p.flags[bm[0]], 1);
        if (__any_sync(FULL, (neg1 >> 31) != 0u) && lane == 0) atomicOr(&p.flags[bm[1]], 1);
    }
}

size_t packed_smem_bytes(int nt) {
    const size_t P = 4 * (size_t)nt, VP = P + 2 * PADC;
    return ((RING + S) * VP + 2 * ((size_t)nt + 4)) * 4;
}

template <int R, bool VEC, int NTC, int WC>
int launch_t(const SparseParams& p, int nt, int cpm, int total, bool force, cudaStream_t st) {
    // one device per process (torchrun: one rank per GPU): attributes and occupancy are looked up once per shape
    static int cached_nt = 0, cached_grid = 0;
    static bool configured = false;
    const size_t smem = packed_smem_bytes(nt);
    if (!configured) {
        KB_CUDA_TRY(cudaFuncSetAttribute(round1_packed_kernel<R, VEC, NTC, WC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));   // (16 bytes of static shared memory: the barriers)
        KB_CUDA_TRY(cudaFuncSetAttribute(round1_packed_kernel<R, VEC, NTC, WC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    if (cached_nt != nt) {
        int sms = 0, occ = 0;
        int rc = kb_sm_count(&sms);
        if (rc != KB_OK) return rc;
        KB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, round1_packed_kernel<R, VEC, NTC, WC>, nt, smem));
        if (occ < 1) return KB_ERR_UNSUPPORTED;
        cached_grid = occ * sms;
        cached_nt = nt;
    }
    // A band costs three extra iterations and the synchronous load of three chunks: bands are kept at MIN_BAND chunks
    // or more, and when that leaves fewer than 96 CTAs (a handful of maps; measured: 8 maps of 1024 x 1024 take 59 us here
    // and 42 us in the tiled kernel) the tiled kernel, whose grid is tiles x maps, is the better choice.
    int grid = total / MIN_BAND;
    if (grid > cached_grid) grid = cached_grid;
    if (force && grid < 1) grid = 1;
    if (!force && grid < 96) return KB_ERR_UNSUPPORTED;
    round1_packed_kernel<R, VEC, NTC, WC><<<grid, nt, smem, st>>>(p, cpm, total);
    KB_LAUNCH_CHECK();
    return KB_OK;
}

// 640 columns (the 480x640 maps of the benchmark configurations) get compile-time pitches
template <int R>
int launch_r(const SparseParams& p, int nt, int cpm, int total, bool vec, bool force, cudaStream_t st) {
    if (vec && p.W == 640) return launch_t<R, true, 160, 640>(p, nt, cpm, total, force, st);
    return vec ? launch_t<R, true, 0, 0>(p, nt, cpm, total, force, st) : launch_t<R, false, 0, 0>(p, nt, cpm, total, force, st);
}

}  // namespace

int launch_round1_packed(const SparseParams& p, bool force, cudaStream_t st) {
    if (p.W > 4 * MAX_NT || p.r < 1 || p.r > 8) return KB_ERR_UNSUPPORTED;
    const int nt = 32 * ((p.W + 127) / 128);
    const int cpm = (p.H + S - 1) / S;
    const long long total = (long long)((p.B + 1) / 2) * cpm;
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
