// Sparse exact detection path: NMS + border + threshold + top-k for non-negative score maps.
//
// The fixed point of the reference's fast_nms rounds (utils/extracter.py:49-98) on a non-negative map
// is greedy NMS over the total order (score desc, raster asc) with Chebyshev radius r: a pixel is kept
// iff no KEPT pixel of higher priority lies within r.  Two consequences drive this path:
//   * the pixels that win their whole (2r+1)^2 window ("round-1 maxima", exactly what the reference
//     finds in its first round, extracter.py:54-70) are certainly kept, and every other pixel within r
//     of one of them is certainly dead and never suppresses anything;
//   * whether a pixel is kept depends only on pixels of HIGHER priority, so detection()'s top_k
//     (extracter.py:217-218) only needs the pixels above the score of the (top_k+1)-th kept one.
//
//   tau_kernel     per map: 4096 sampled pixels (128 runs of 32 floats) give the score tau above which
//                  ~1/12 of the pixels lie; lower pixels are not listed at all.
//   round1_kernel  THE streaming kernel (one read of every pixel, 32x128 tiles + 2r halo staged in shared
//                  memory with 16-byte loads): round-1 maxima found hierarchically (4x4 block maxima, a
//                  prefilter against whole neighbouring blocks, an exact warp-wide window check of the few
//                  survivors with the first-of-ties rule), 1-bit maxima mask dilated by r on 32-bit words;
//                  emits two candidate lists per map as 64-bit priority keys: round-1 maxima > tau and
//                  pixels > tau not covered by any round-1 maximum.  Everything else is decided.
//   sparse_kernel  one CTA per map: takes the ~1.25*top_k best round-1 maxima and every uncovered
//                  candidate at or above the weakest of them, bins them on a coarse cell grid in shared
//                  memory and resolves the remaining keep/suppress decisions by priority (a candidate is
//                  kept once every higher-priority neighbour within r is dead, dead once one is kept),
//                  then sorts the kept interior pixels and writes the top_k rows.
//
// A map is certified when it produced more than top_k interior survivors (K > top_k: rows sorted by
// priority) or when its candidate lists were complete (K <= top_k: raster order, extracter.py:217).
// Anything else (negative scores, list overflow, nms_dist > 8) is flagged on the device for the
// round-faithful kernel (kb_nms.cu), so results are exact in all cases.
#include "kb_sparse.cuh"

namespace kbsparse {

// ------------------------------------------------------------------------------------------------
// tau: the rank-th largest of 4096 sampled pixels (bitwise counting select, no atomics)
__global__ void __launch_bounds__(TAU_NT) tau_kernel(SparseParams p) {
    __shared__ int s_part[TAU_NT / 32];
    __shared__ int s_part2[2][TAU_NT / 32];
    const int b = blockIdx.x;
    const long long npx = (long long)p.H * p.W;
    const float theta = fmaxf(p.threshold, 0.0f);
    if (threadIdx.x == 0) { p.cntM[b] = 0; p.cntO[b] = 0; p.flags[b] = 0; p.need_fallback[b] = 0; }
    if (b == 0 && threadIdx.x == 0) *p.any_fallback = 0;
    const long long rank = ((long long)p.c_pix * SAMPLES + npx - 1) / npx;
    if (npx <= 4 * SAMPLES || rank >= SAMPLES / 2) {        // small map / dense request: list everything
        if (threadIdx.x == 0) { p.tau[b] = theta; p.qscale[b] = 1.0f; }
        return;
    }
    const float* img = p.score + (size_t)b * npx;
    constexpr int PER = SAMPLES / TAU_NT;                   // 16 keys per thread
    constexpr int RUN = 32;                                 // floats per contiguous run
    const long long n_runs = SAMPLES / RUN, stride = npx / n_runs;
    uint32_t keys[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int k = i * TAU_NT + threadIdx.x;             // sample index: run = k / 32, element = k % 32
        const int run = k / RUN, el = k % RUN;
        uint32_t h = (uint32_t)run * 2654435761u;
        h ^= h >> 15;
        long long start = (long long)run * stride + (long long)(h % (uint32_t)(stride - RUN + 1));
        keys[i] = kb::float_order_key(__ldg(img + start + el));
    }
    uint32_t v = 0u;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t t = v | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) c += keys[i] >= t ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        int* part = s_part2[bit & 1];                       // two buffers: one barrier per bit instead of two
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < TAU_NT / 32; ++w) tot += part[w];
        if (tot >= (int)rank) v = t;
    }
    // the largest sample (for the packed round-1 kernel's 16-bit image, which must neither overflow nor underflow on
    // maps whose scores live far from 1: it takes fp16((score - tau) * qscale) with the largest SAMPLE at 2^15, so only
    // pixels more than twice the largest sample above tau saturate)
    uint32_t kmax = 0u;
#pragma unroll
    for (int i = 0; i < PER; ++i) kmax = max(kmax, keys[i]);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = (int)kmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float t = kb::float_from_order_key(v);
        const float tau = (t > theta) ? t : theta;          // NaN-safe: falls back to theta
        p.tau[b] = tau;
        uint32_t km = 0u;
#pragma unroll
        for (int w = 0; w < TAU_NT / 32; ++w) km = max(km, (uint32_t)s_part[w]);
        const float span = kb::float_from_order_key(km) - tau;
        // (a power of two: the scaling itself is then exact; NaN / inf / non-positive spans fall back to 1)
        float sc = 1.0f;
        if (span > 1e-30f && span < 1e30f) sc = exp2f(15.0f - ceilf(log2f(span)));
        p.qscale[b] = sc;
    }
}

// ------------------------------------------------------------------------------------------------
// round-1 maxima of a tile, hierarchically.  A pixel that wins its (2R+1)^2 window also wins the
// BSxBS block it lies in (BS - 1 <= R), so:
//   1. every thread reduces one BSxBS block of the staged tile to (maximum, first position of it);
//   2. the block winner is compared with the maxima of the neighbouring blocks that lie ENTIRELY inside its
//      window -- almost all winners die here with a handful of shared-memory reads;
//   3. the few survivors (~1 per 9 blocks) are checked exactly against their whole window, one warp per
//      survivor (lanes over the (2R+1)^2 entries): strictly greater than everything earlier in raster order,
//      >= everything later (torch.argmax returns the first maximum, extracter.py:69-70).
// The maxima are kept as one bit per pixel; their dilation by R (coverage), the `score > tau` mask and the
// emission all work on 32-pixel words.
template <int R>
struct Tile {
    static constexpr int BS = (R >= 3) ? 4 : 2;                   // block edge (BS - 1 <= R)
    static constexpr int SH = DTH + 4 * R, SW = DTW + 4 * R;      // staged scores (halo 2R); multiples of BS
    static constexpr int SP = SW;                                 // row pitch (multiple of 4: float4 rows)
    static constexpr int MH = DTH + 2 * R, MW = DTW + 2 * R;      // region whose round-1 maxima matter
    static constexpr int MWW = (MW + 31) / 32;                    // mask words per region row
    static constexpr int NBY = SH / BS, NBX = SW / BS;            // blocks of the staged tile
    static constexpr int OW = DTW / 32;                           // mask words per output row
    static constexpr int SWW = (SW + 31) / 32;                    // mask words per staged row
    static constexpr size_t smem_bytes() {
        return (size_t)SH * SP * 4 + (size_t)NBY * NBX * 4 + (size_t)NBY * NBX + 16 +
               (size_t)(2 * MH + DTH) * MWW * 4 + (size_t)DTH * OW * 4 + 64;
    }
};

template <int R>
__global__ void __launch_bounds__(DNT) round1_kernel(SparseParams p) {
    using T = Tile<R>;
    constexpr int W = 2 * R + 1, BS = T::BS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* S = reinterpret_cast<float*>(smem_raw);                // [SH][SP]
    float* BV = S + T::SH * T::SP;                                // [NBY][NBX] block maxima
    uint32_t* MB = reinterpret_cast<uint32_t*>(BV + T::NBY * T::NBX);      // [MH][MWW] round-1 maxima bits
    uint32_t* DB = MB + T::MH * T::MWW;                           // [MH][MWW] horizontally dilated
    uint32_t* CB = DB + T::MH * T::MWW;                           // [DTH][MWW] coverage
    uint32_t* TB = CB + DTH * T::MWW;                             // [DTH][OW] score > tau (output columns)
    uint8_t* BP = reinterpret_cast<uint8_t*>(TB + DTH * T::OW);   // [NBY][NBX] block-local position of the maximum
    __shared__ int s_scan[33];
    __shared__ int s_base[2];

    const int b = blockIdx.z;
    const int y0 = blockIdx.y * DTH, x0 = blockIdx.x * DTW;
    const int H = p.H, Wd = p.W;
    const float* img = p.score + (size_t)b * H * Wd;
    const float tau = p.tau[b];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- stage the tile with a 2R halo (zero padding, extracter.py:58) ---------------------------
    // 16-byte chunks when rows are 16-byte aligned (chunks then never straddle the image border),
    // otherwise one warp per row with scalar loads.
    bool neg = false;
    const bool vec = (R % 2 == 0) && (Wd % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15u) == 0);
    if (vec) {
        constexpr int C4 = T::SW / 4;                               // chunks per staged row
        constexpr int N4 = T::SH * C4;
        constexpr int PER = (N4 + DNT - 1) / DNT;
        float4 val[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int idx = threadIdx.x + k * DNT;
            const int sy = idx / C4, c4 = idx - sy * C4;
            const int gy = y0 + sy - 2 * R, gx = x0 - 2 * R + 4 * c4;
            val[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < N4 && gy >= 0 && gy < H && gx >= 0 && gx < Wd)
                val[k] = __ldg(reinterpret_cast<const float4*>(img + (size_t)gy * Wd + gx));
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int idx = threadIdx.x + k * DNT;
            if (idx < N4) {
                neg |= (val[k].x < 0.0f) | (val[k].y < 0.0f) | (val[k].z < 0.0f) | (val[k].w < 0.0f);
                *reinterpret_cast<float4*>(S + 4 * idx) = val[k];   // SP == SW: the staged tile is dense
            }
        }
    } else {
        constexpr int CH = (T::SW + 31) / 32;
        const int gx_base = x0 - 2 * R + lane;
        for (int sy = warp; sy < T::SH; sy += DNT / 32) {
            const int gy = y0 + sy - 2 * R;
            const bool row_ok = gy >= 0 && gy < H;
            const float* grow = img + (size_t)(row_ok ? gy : 0) * Wd;
            float val[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int gx = gx_base + 32 * k;
                val[k] = (row_ok && gx >= 0 && gx < Wd) ? __ldg(grow + gx) : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                neg |= val[k] < 0.0f;
                if (lane + 32 * k < T::SW) S[sy * T::SP + lane + 32 * k] = val[k];
            }
        }
    }
    for (int i = threadIdx.x; i < T::MH * T::MWW; i += DNT) MB[i] = 0u;
    __syncthreads();
    if (__any_sync(0xffffffffu, neg) && lane == 0) atomicOr(&p.flags[b], 1);

    // ---- 1. block maxima (value, first position in raster order inside the block) ---------------
    for (int blk = threadIdx.x; blk < T::NBY * T::NBX; blk += DNT) {
        const int by = blk / T::NBX, bx = blk - by * T::NBX;
        const float* src = S + (by * BS) * T::SP + bx * BS;
        float best = -1.0f;                                        // scores are >= 0 on this path
        int pos = 0;
#pragma unroll
        for (int yy = 0; yy < BS; ++yy) {
            float v[BS];
            if (BS == 4) {
                const float4 q = *reinterpret_cast<const float4*>(src + yy * T::SP);
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[BS - 1] = q.w;
            } else {
                const float2 q = *reinterpret_cast<const float2*>(src + yy * T::SP);
                v[0] = q.x; v[BS - 1] = q.y;
            }
#pragma unroll
            for (int xx = 0; xx < BS; ++xx)
                if (v[xx] > best) { best = v[xx]; pos = yy * BS + xx; }          // strict: first maximum
        }
        BV[blk] = best;
        BP[blk] = (uint8_t)pos;
    }
    __syncthreads();

    // ---- 2. + 3. prefilter against whole neighbouring blocks, exact check of the survivors -------
    // Only maxima above tau matter: a maximum <= tau neither gets listed nor covers a pixel > tau, because
    // everything within R of it scores lower.  The exact check is one warp per survivor, lanes over the
    // (2R+1)^2 window entries.
    {
        constexpr int NB = T::NBY * T::NBX;
        constexpr int NK = (W * W + 31) / 32;
        // lane-constant part of the exact check: offsets of this lane's window entries and whether they come
        // earlier in raster order (bit k of `early`); entries past the window are parked on the centre
        int woff[NK];                                               // in floats, relative to the centre
        constexpr int CENTRE = R * W + R;                            // window entries before the centre come earlier
        constexpr int KMIX = CENTRE / 32;                            // the only k whose 32 entries straddle the centre
        const bool mix_early = lane + 32 * KMIX < CENTRE;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int idx = lane + 32 * k;
            const int dy = idx / W - R, dx = idx - (idx / W) * W - R;
            woff[k] = idx < W * W ? dy * T::SP + dx : 0;
        }
        for (int base = 0; base < NB; base += DNT) {
            const int blk = base + threadIdx.x;
            bool surv = false;
            int sy = 0, sx = 0;
            float c = 0.0f;
            if (blk < NB) {
                const int by = blk / T::NBX, bx = blk - by * T::NBX;
                const int pos = BP[blk];
                const int oy = pos / BS, ox = pos - oy * BS;
                sy = by * BS + oy; sx = bx * BS + ox;
                c = BV[blk];
                // only pixels of the region (R inside the staged tile) can matter
                surv = c > tau && sy >= R && sy < T::SH - R && sx >= R && sx < T::SW - R;
                if (surv) {
                    // blocks by+dy with BS*(by+dy) >= sy-R and BS*(by+dy)+BS-1 <= sy+R are inside the window
                    const int dy0 = -((R - oy) / BS), dy1 = (oy + R - BS + 1) / BS;
                    const int dx0 = -((R - ox) / BS), dx1 = (ox + R - BS + 1) / BS;
                    for (int dy = dy0; dy <= dy1 && surv; ++dy)
                        for (int dx = dx0; dx <= dx1; ++dx)
                            if ((dy | dx) != 0 && BV[(by + dy) * T::NBX + bx + dx] > c) { surv = false; break; }
                }
            }
            unsigned pend = __ballot_sync(0xffffffffu, surv);
            const int cofs = sy * T::SP + sx;
            while (pend) {
                const int src = __ffs(pend) - 1;
                pend &= pend - 1;
                const int co = __shfl_sync(0xffffffffu, cofs, src);
                const float cc = __shfl_sync(0xffffffffu, c, src);
                // entries earlier in raster order must be < cc, later ones <= cc.  cc > tau >= 0, so "v >= cc" is
                // "v > the float just below cc": one strict comparison per entry against a per-k threshold
                const float cc_lo = __uint_as_float(__float_as_uint(cc) - 1u);
                bool bad = false;
                const float* centre = S + co;
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const float v = centre[woff[k]];
                    const float thr = k < KMIX ? cc_lo : (k > KMIX ? cc : (mix_early ? cc_lo : cc));
                    bad |= v > thr;                                             // the centre itself: v > cc is false
                }
                if (!__any_sync(0xffffffffu, bad) && lane == 0) {
                    const int cy = co / T::SP, cx = co - cy * T::SP;
                    const int my = cy - R, mx = cx - R;                         // region coordinates
                    atomicOr(&MB[my * T::MWW + (mx >> 5)], 1u << (mx & 31));
                }
            }
        }
    }
    __syncthreads();

    // ---- dilation of the maxima mask by R: rows, then columns -----------------------------------
    for (int i = threadIdx.x; i < T::MH * T::MWW; i += DNT) {
        const int my = i / T::MWW, w = i - my * T::MWW;
        const uint32_t cur = MB[i];
        const uint32_t prev = w > 0 ? MB[i - 1] : 0u;
        const uint32_t next = w + 1 < T::MWW ? MB[i + 1] : 0u;
        uint32_t acc = cur;
#pragma unroll
        for (int d = 1; d <= R; ++d)
            acc |= (cur << d) | (prev >> (32 - d)) | (cur >> d) | (next << (32 - d));
        DB[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DTH * T::MWW; i += DNT) {
        const int y = i / T::MWW, w = i - y * T::MWW;
        uint32_t acc = 0u;
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) acc |= DB[(y + d) * T::MWW + w];
        CB[i] = acc;
    }
    __syncthreads();

    // ---- emission: round-1 maxima and uncovered pixels above tau, one 32-pixel word per thread ---
    uint32_t emM = 0u, emO = 0u;
    int ey = 0, ew = 0;
    if (threadIdx.x < DTH * T::OW) {
        ey = threadIdx.x / T::OW;
        ew = threadIdx.x - ey * T::OW;
        // output column x is region column x + R: realign the region words
        const uint32_t* mrow = MB + (ey + R) * T::MWW;
        const uint32_t* crow = CB + ey * T::MWW;
        const uint32_t m_lo = mrow[ew], m_hi = ew + 1 < T::MWW ? mrow[ew + 1] : 0u;
        const uint32_t c_lo = crow[ew], c_hi = ew + 1 < T::MWW ? crow[ew + 1] : 0u;
        const uint32_t m1w = __funnelshift_r(m_lo, m_hi, R), covw = __funnelshift_r(c_lo, c_hi, R);
        // score > tau of this thread's 32 output pixels (pixels outside the image hold 0 <= tau)
        const float* srow = S + (ey + 2 * R) * T::SP + 2 * R + 32 * ew;
        uint32_t hitw = 0u;
        if ((2 * R) % 4 == 0) {                                     // 16-byte aligned rows
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = reinterpret_cast<const float4*>(srow)[q];
                hitw |= ((v.x > tau ? 1u : 0u) | (v.y > tau ? 2u : 0u) | (v.z > tau ? 4u : 0u) | (v.w > tau ? 8u : 0u)) << (4 * q);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) hitw |= (srow[q] > tau ? 1u : 0u) << q;
        }
        emM = hitw & m1w;
        emO = hitw & ~m1w & ~covw;
    }
    int tot;
    const int packed = kb::block_exclusive_scan(__popc(emM) | (__popc(emO) << 16), s_scan, &tot);
    if (threadIdx.x == 0) {
        s_base[0] = (tot & 0xffff) ? atomicAdd(&p.cntM[b], tot & 0xffff) : 0;
        s_base[1] = (tot >> 16) ? atomicAdd(&p.cntO[b], tot >> 16) : 0;
    }
    __syncthreads();
    int offM = s_base[0] + (packed & 0xffff), offO = s_base[1] + (packed >> 16);
    uint64_t* outM = p.listM + (size_t)b * LIST_CAP;
    uint64_t* outO = p.listO + (size_t)b * LIST_CAP;
    uint32_t both = emM | emO;
    while (both) {
        const int bit = __ffs(both) - 1;
        both &= both - 1;
        const int x = 32 * ew + bit;                                // tile column
        const float sc = S[(ey + 2 * R) * T::SP + x + 2 * R];
        const uint64_t key = kb::priority_key(sc, (uint32_t)((y0 + ey) * Wd + x0 + x));
        if (emM & (1u << bit)) { if (offM < LIST_CAP) outM[offM] = key; ++offM; }
        else { if (offO < LIST_CAP) outO[offO] = key; ++offO; }
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit(const SparseParams& p, int b, int slot, float score, uint32_t ras) {
    const int row = ras / p.W, col = ras - row * p.W;
    KB_ASSERT(slot >= 0 && slot < p.top_k && ras < (uint32_t)(p.H * p.W));
    float* o = p.xyp + ((size_t)b * p.top_k + slot) * 3;
    o[0] = ((float)col + 0.5f) / (float)p.W;          // extracter.py:149,158
    o[1] = ((float)row + 0.5f) / (float)p.H;
    o[2] = score;
    p.raster[(size_t)b * p.top_k + slot] = (int)ras;
}

// Descending bitonic sort of a[0..n) (n a power of two) by the whole CTA.  Element i is always touched by
// thread i % SP_NT or by its partner i ^ j, which lives in the same warp whenever j < 32: only the steps
// with j >= 32 (and the first step of every merge) need a block-wide barrier, the rest a warp barrier.
__device__ void bitonic_desc(uint64_t* a, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        if (k > 32) __syncthreads();
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += SP_NT) {
                const int l = i ^ j;
                if (l > i) {
                    const uint64_t x = a[i], y = a[l];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
                }
            }
            if (j >= 32) __syncthreads(); else __syncwarp();
        }
    }
    __syncthreads();
}

// Same sort by a TEAM of the CTA's first n / E threads with E keys per thread held in registers (blocked layout:
// thread t owns a[t*E .. t*E+E-1]): compare-exchange distances below E stay inside a thread, distances below 32*E
// are 64-bit warp shuffles, and only the remaining ones go through shared memory behind a named barrier of the team.
// The sort is a chain of ~log^2(n)/2 dependent steps, so its time is (steps that synchronise) x (cost of the
// synchronisation), not instruction issue: with 8 keys per thread 2048 keys need 6 team barriers among 256 threads
// where one or two keys per thread over all 1024 threads needed 15-27 block-wide ones (34 k cycles -> measured
// in profiles/r02_sparse_phases.txt).  The rest of the CTA waits at the closing __syncthreads.
// `spare` (may be null): n more keys of scratch, which saves the second barrier of every shared-memory step.
__device__ __forceinline__ void team_barrier(int nts) { asm volatile("bar.sync 1, %0;" ::"r"(nts) : "memory"); }

template <int E>
__device__ void bitonic_desc_team(uint64_t* a, uint64_t* spare, int n) {
    const int t = threadIdx.x, nts = n / E;               // nts is a multiple of 32
    if (t < nts) {
        uint64_t v[E];
#pragma unroll
        for (int s = 0; s < E; ++s) v[s] = a[t * E + s];
        int flip = 0;
        for (int k = 2; k <= n; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j >= 32 * E) {
                    uint64_t* buf = (spare != nullptr && (flip & 1)) ? spare : a;
                    if (spare == nullptr && nts > 32) team_barrier(nts);      // everybody has read the previous step's values
#pragma unroll
                    for (int s = 0; s < E; ++s) buf[t * E + s] = v[s];
                    team_barrier(nts);
#pragma unroll
                    for (int s = 0; s < E; ++s) {
                        const int e = t * E + s, q = e ^ j;
                        const uint64_t o = buf[q];
                        const bool up = (e & k) == 0, first = e < q;
                        const bool take_max = up == first;            // descending block: the lower index keeps the larger key
                        v[s] = take_max ? (v[s] > o ? v[s] : o) : (v[s] < o ? v[s] : o);
                    }
                    ++flip;
                } else if (j >= E) {
                    const int lane_mask = j / E;
#pragma unroll
                    for (int s = 0; s < E; ++s) {
                        const int e = t * E + s;
                        const uint64_t o = __shfl_xor_sync(0xffffffffu, v[s], lane_mask);
                        const bool up = (e & k) == 0, first = (e & j) == 0;
                        const bool take_max = up == first;
                        v[s] = take_max ? (v[s] > o ? v[s] : o) : (v[s] < o ? v[s] : o);
                    }
                } else {
                    // j < E: both keys live in this thread; j must be a compile-time constant for v[] to stay in registers
#pragma unroll
                    for (int J = E / 2; J > 0; J >>= 1) {
                        if (J == j) {
#pragma unroll
                            for (int s = 0; s < E; ++s) {
                                if ((s & J) == 0) {
                                    const int e = t * E + s;
                                    const bool up = (e & k) == 0;
                                    const uint64_t x = v[s], y = v[s | J];
                                    const bool swap = up ? (x < y) : (x > y);
                                    v[s] = swap ? y : x;
                                    v[s | J] = swap ? x : y;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (nts > 32) team_barrier(nts);                              // the last readers of `a` are done
#pragma unroll
        for (int s = 0; s < E; ++s) a[t * E + s] = v[s];
    }
    __syncthreads();
}

// dispatch on the padded size (a power of two); `cap` = keys that fit behind a[0..n) for the spare buffer
__device__ void sort_desc(uint64_t* a, int n, int cap) {
    uint64_t* spare = (2 * n <= cap) ? a + n : nullptr;
    if (n < 64) { bitonic_desc(a, n); return; }
    if (n >= 256) bitonic_desc_team<8>(a, spare, n);                  // 32 .. 1024 threads (16 keys per thread would spill
                                                                      // under the 64-register cap of a 1024-thread CTA)
    else if (n == 128) bitonic_desc_team<4>(a, spare, n);             // one warp: no barrier at all
    else bitonic_desc_team<2>(a, spare, n);
}

// The first `want` keys of the DESCENDING order of src[0..n) into dst (dst must hold n keys and not overlap src).
// Bucket sort: the keys' upper halves (the score's order-preserving bits) are spread linearly over NBINS buckets
// between their minimum and maximum, a counting sort puts every key into its bucket, and the buckets that reach
// into the first `want` places -- a handful of keys each -- are put in order by one thread each.  About ten
// block-wide barriers and a few passes over the keys, where the bitonic network is a chain of log^2(n)/2 dependent
// 64-bit compare-exchange steps (2048 keys: 40 k cycles; this: measured in profiles/r02_sparse_phases.txt).
// Returns false (dst unspecified, src intact) when some bucket is too full for that -- heavily tied scores -- and the
// caller falls back to the bitonic sort.  `hist` = NBINS + 1 words of scratch.
constexpr int SORT_BINS = 4096;
constexpr int SORT_BIN_MAX = 48;

__device__ bool bucket_sort_desc(const uint64_t* src, uint64_t* dst, int n, int want, uint32_t* hist, int* s_scan, int* s_part) {
    __shared__ uint32_t s_lo, s_hi;
    __shared__ int s_bad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // 1. range of the score bits
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int i = threadIdx.x; i < n; i += SP_NT) {
        const uint32_t k = (uint32_t)(src[i] >> 32);
        lo = min(lo, k); hi = max(hi, k);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    uint32_t* s_w = reinterpret_cast<uint32_t*>(s_part);            // SP_NT / 32 words: minima, then maxima
    if (lane == 0) s_w[warp] = lo;
    for (int i = threadIdx.x; i <= SORT_BINS; i += SP_NT) hist[i] = 0u;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_w[lane];
        v = __reduce_min_sync(0xffffffffu, v);
        if (lane == 0) s_lo = v;
    }
    __syncthreads();
    if (lane == 0) s_w[warp] = hi;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_w[lane];
        v = __reduce_max_sync(0xffffffffu, v);
        if (lane == 0) { s_hi = v; s_bad = 0; }
    }
    __syncthreads();
    lo = s_lo; hi = s_hi;
    const uint32_t range = hi - lo;
    int shift = 0;
    while (shift < 32 && (range >> shift) >= (uint32_t)SORT_BINS) ++shift;
    // 2. bucket sizes (bucket 0 = the largest scores)
    for (int i = threadIdx.x; i < n; i += SP_NT) {
        const uint32_t k = (uint32_t)(src[i] >> 32);
        atomicAdd(&hist[SORT_BINS - 1 - (int)((k - lo) >> shift)], 1u);
    }
    __syncthreads();
    // 3. exclusive scan of the bucket sizes -> first place of every bucket; too full a bucket among the wanted ones?
    {
        constexpr int PERB = SORT_BINS / SP_NT;                       // 4
        uint32_t loc[PERB];
        int sum = 0;
#pragma unroll
        for (int i = 0; i < PERB; ++i) { loc[i] = hist[threadIdx.x * PERB + i]; sum += (int)loc[i]; }
        int tot;
        int run = kb::block_exclusive_scan(sum, s_scan, &tot);
        bool bad = false;
#pragma unroll
        for (int i = 0; i < PERB; ++i) {
            hist[threadIdx.x * PERB + i] = (uint32_t)run;
            bad |= run < want && (int)loc[i] > SORT_BIN_MAX;
            run += (int)loc[i];
        }
        if (threadIdx.x == SP_NT - 1) hist[SORT_BINS] = (uint32_t)run;
        if (bad) s_bad = 1;
    }
    __syncthreads();
    if (s_bad) return false;
    // 4. scatter: hist[b] walks from the first place of bucket b to the first place of bucket b + 1
    for (int i = threadIdx.x; i < n; i += SP_NT) {
        const uint64_t key = src[i];
        const int bin = SORT_BINS - 1 - (int)(((uint32_t)(key >> 32) - lo) >> shift);
        KB_ASSERT(bin >= 0 && bin < SORT_BINS);
        const uint32_t place = atomicAdd(&hist[bin], 1u);
        KB_ASSERT(place < (uint32_t)n);
        dst[place] = key;
    }
    __syncthreads();
    // 5. order inside the buckets that matter (bucket b now ends at hist[b] and starts where bucket b-1 ends)
    for (int bin = threadIdx.x; bin < SORT_BINS; bin += SP_NT) {
        const int start = bin ? (int)hist[bin - 1] : 0, end = (int)hist[bin];
        if (start >= want || end - start < 2) continue;
        for (int i = start + 1; i < end; ++i) {                      // insertion sort, descending
            const uint64_t key = dst[i];
            int j = i - 1;
            while (j >= start && dst[j] < key) { dst[j + 1] = dst[j]; --j; }
            dst[j + 1] = key;
        }
    }
    __syncthreads();
    return true;
}

__device__ __forceinline__ int block_sum(int v, int* s_part) {
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int w = 0; w < SP_NT / 32; ++w) tot += s_part[w];
    __syncthreads();
    return tot;
}

constexpr uint8_t ST_DEAD = 0, ST_UNDEC = 1, ST_KEPT = 2;

// phase timestamps of map 0's CTA (clock64), written when KB_KNOB_SPARSE_PROF is set: kb_debug_sparse_prof reads them
__device__ long long g_sparse_prof[16];
#define KB_SP_PROF(k) do { if (p.prof && b == 0 && threadIdx.x == 0) g_sparse_prof[k] = clock64(); } while (0)

__global__ void __launch_bounds__(SP_NT, 1) sparse_kernel(SparseParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);                   // [SMEM_CAP]
    uint32_t* pos = reinterpret_cast<uint32_t*>(keys + SMEM_CAP);             // [SMEM_CAP] y << 16 | x
    uint32_t* cstart = pos + SMEM_CAP;                                        // [MAX_CELLS + 1]
    volatile uint8_t* state = reinterpret_cast<volatile uint8_t*>(cstart + MAX_CELLS + 1);   // [SMEM_CAP]
    uint16_t* und = reinterpret_cast<uint16_t*>(const_cast<uint8_t*>(state) + SMEM_CAP);      // [SMEM_CAP] undecided slots, by band
    __shared__ int s_scan[33];
    __shared__ int s_part[SP_NT / 32];
    __shared__ int s_cut[3];
    __shared__ int s_und[3];
    __shared__ int s_band[33], s_fill[32];

    const int b = blockIdx.x;
    const int H = p.H, W = p.W, r = p.r;
    KB_SP_PROF(0);
    if (p.prof && b == 0 && threadIdx.x == 0) { g_sparse_prof[13] = 0; g_sparse_prof[15] = 0; }
    auto fallback = [&]() {
        if (threadIdx.x == 0) { p.need_fallback[b] = 1; atomicExch(p.any_fallback, 1); }
    };
    const int nM = p.cntM[b], nO = p.cntO[b];
    if ((p.flags[b] & 1) || nM > LIST_CAP || nO > LIST_CAP) { fallback(); return; }
    const uint64_t* LM = p.listM + (size_t)b * LIST_CAP;
    const uint64_t* LO = p.listO + (size_t)b * LIST_CAP;
    const float theta = fmaxf(p.threshold, 0.0f);
    const bool lists_complete = (p.tau[b] == theta);

    // The listing is exact for every pixel with score >= T for ANY cut T; try the cheapest cut first
    // (the ~1.25*top_k best round-1 maxima), widen if it does not certify.
    for (int attempt = 0; attempt < 3; ++attempt) {
        long long ksel = attempt == 0 ? (long long)p.top_k + p.top_k / 4 + 32
                       : attempt == 1 ? 3LL * p.top_k + 64 : (long long)LIST_CAP + 1;
        // ---- cut: score key of the ksel-th largest round-1 maximum (0 = take everything) -------
        // Radix select on the score key's bits 31..12 (any cut is exact, so the 12 low bits are not resolved): a
        // 2048-bin histogram of bits 31..21, then a 512-bin histogram of bits 20..12 inside the bin that holds the
        // ksel-th largest -- a dozen block-wide barriers where one counting vote per bit took 20 to 40.
        uint32_t tkey = 0u;
        if ((long long)nM > ksel) {
            uint32_t* hist = cstart;                               // scratch until the cell grid is built
            for (int i = threadIdx.x; i < 2048 + 512; i += SP_NT) hist[i] = 0u;
            if (threadIdx.x == 0) { s_cut[0] = -1; s_cut[1] = 0; s_cut[2] = -1; }
            __syncthreads();
            for (int i = threadIdx.x; i < nM; i += SP_NT) {
                const uint64_t key = LM[i];
                KB_ASSERT((uint32_t)(key >> 53) < 2048u);
                if (key != 0ull) atomicAdd(&hist[(uint32_t)(key >> 53)], 1u);       // bits 31..21 of the score key
            }
            __syncthreads();
            {   // thread t owns bins 2047-2t and 2046-2t: S(d) = number of keys with digit >= d
                const int d0 = 2047 - 2 * (int)threadIdx.x, d1 = d0 - 1;
                const int h0 = (int)hist[d0], h1 = (int)hist[d1];
                int tot;
                const int above = kb::block_exclusive_scan(h0 + h1, s_scan, &tot);
                if ((long long)above < ksel && (long long)(above + h0) >= ksel) { s_cut[0] = d0; s_cut[1] = above; }
                else if ((long long)(above + h0) < ksel && (long long)(above + h0 + h1) >= ksel) { s_cut[0] = d1; s_cut[1] = above + h0; }
            }
            __syncthreads();
            const int dA = s_cut[0], aboveA = s_cut[1];
            if (dA >= 0) {                                          // (fewer than ksel real keys: take everything)
                uint32_t* hist2 = hist + 2048;
                for (int i = threadIdx.x; i < nM; i += SP_NT) {
                    const uint32_t k32 = (uint32_t)(LM[i] >> 32);
                    if ((int)(k32 >> 21) == dA && LM[i] != 0ull) atomicAdd(&hist2[(k32 >> 12) & 0x1ffu], 1u);
                }
                __syncthreads();
                {
                    const int d = 511 - (int)threadIdx.x;
                    const int h = d >= 0 ? (int)hist2[d] : 0;
                    int tot;
                    const int above = aboveA + kb::block_exclusive_scan(h, s_scan, &tot);
                    if (d >= 0 && (long long)above < ksel && (long long)(above + h) >= ksel) s_cut[2] = d;
                }
                __syncthreads();
                const int dB = s_cut[2] >= 0 ? s_cut[2] : 0;
                tkey = ((uint32_t)dA << 21) | ((uint32_t)dB << 12);
            }
            __syncthreads();                                        // s_cut and the histograms are free again
        }
        const bool cut_complete = lists_complete && tkey == 0u;
        KB_SP_PROF(1);
        // ---- the candidates at or above the cut, stored in CELL ORDER -------------------------------
        // A counting sort by coarse cell straight from the two lists (read twice: count, then place), so that the
        // neighbourhood of a candidate is three contiguous runs of (position, key, state) entries -- no index
        // indirection between a cell and its candidates.
        const int n_cells = p.gw * p.gh;
        for (int i = threadIdx.x; i <= n_cells; i += SP_NT) cstart[i] = 0u;
        __syncthreads();
        for (int pass = 0; pass < 2; ++pass) {
            const uint64_t* L = pass ? LO : LM;
            const int n = pass ? nO : nM;
            for (int idx = threadIdx.x; idx < n; idx += SP_NT) {
                const uint64_t key = L[idx];
                // (the streaming round-1 kernel pads the blocks a warp reserved with null keys)
                if (key == 0ull || (uint32_t)(key >> 32) < tkey) continue;
                const uint32_t ras = kb::key_raster(key);
                const uint32_t y = ras / (uint32_t)W, x = ras - y * (uint32_t)W;
                atomicAdd(&cstart[(int)(y >> p.cell_shift) * p.gw + (int)(x >> p.cell_shift) + 1], 1u);
            }
        }
        __syncthreads();
        int c = 0;
        {   // inclusive scan over cstart[1..n_cells] -> cstart[k] = first slot of cell k
            constexpr int PERC = MAX_CELLS / SP_NT;               // 8
            uint32_t loc[PERC];
            int sum = 0;
#pragma unroll
            for (int i = 0; i < PERC; ++i) {
                const int cell = threadIdx.x * PERC + i;
                loc[i] = cell < n_cells ? cstart[cell + 1] : 0u;
                sum += (int)loc[i];
            }
            int run = kb::block_exclusive_scan(sum, s_scan, &c);
#pragma unroll
            for (int i = 0; i < PERC; ++i) {
                const int cell = threadIdx.x * PERC + i;
                if (cell < n_cells) cstart[cell + 1] = (uint32_t)run;     // start of this cell, shifted by one
                run += (int)loc[i];
            }
        }
        __syncthreads();
        if (c > SMEM_CAP) { fallback(); return; }
        KB_SP_PROF(2);
        // cstart[cell+1] holds the start of `cell`; bumping it while placing leaves the start of cell+1
        if (threadIdx.x == 0) { s_und[0] = 0; s_und[1] = (int)0xffffffffu; s_und[2] = 0; }
        __syncthreads();
        {
            int my_und = 0;
            uint32_t my_lo = 0xffffffffu, my_hi = 0u;
            for (int pass = 0; pass < 2; ++pass) {
                const uint64_t* L = pass ? LO : LM;
                const int n = pass ? nO : nM;
                for (int idx = threadIdx.x; idx < n; idx += SP_NT) {
                    const uint64_t key = L[idx];
                    if (key == 0ull || (uint32_t)(key >> 32) < tkey) continue;
                    const uint32_t ras = kb::key_raster(key);
                    const uint32_t y = ras / (uint32_t)W, x = ras - y * (uint32_t)W;
                    KB_ASSERT((int)(y >> p.cell_shift) * p.gw + (int)(x >> p.cell_shift) < n_cells && y < (uint32_t)H);
                    const uint32_t slot = atomicAdd(&cstart[(int)(y >> p.cell_shift) * p.gw + (int)(x >> p.cell_shift) + 1], 1u);
                    KB_ASSERT(slot < (uint32_t)c && c <= SMEM_CAP);
                    keys[slot] = key;
                    pos[slot] = (y << 16) | x;
                    state[slot] = pass ? ST_UNDEC : ST_KEPT;            // a round-1 maximum is kept for certain
                    if (pass) {
                        ++my_und;
                        my_lo = min(my_lo, (uint32_t)(key >> 32));
                        my_hi = max(my_hi, (uint32_t)(key >> 32));
                    }
                }
            }
            // number of undecided candidates and the range of their score bits (for the priority bands below)
            my_und = __reduce_add_sync(0xffffffffu, my_und);
            my_lo = __reduce_min_sync(0xffffffffu, my_lo);
            my_hi = __reduce_max_sync(0xffffffffu, my_hi);
            if ((threadIdx.x & 31) == 0 && my_und) {
                atomicAdd(&s_und[0], my_und);
                atomicMin(reinterpret_cast<unsigned int*>(&s_und[1]), my_lo);
                atomicMax(reinterpret_cast<unsigned int*>(&s_und[2]), my_hi);
            }
        }
        __syncthreads();
        // now cell k occupies slots cstart[k] .. cstart[k+1]  (cstart[0] == 0)
        KB_SP_PROF(3);

        // ---- keep / suppress decisions by priority -------------------------------------------------
        // A candidate is dead once a kept neighbour of higher priority exists and kept once every such neighbour is
        // dead; it waits while one is undecided.  No block-wide barrier per round: every thread re-examines ITS
        // undecided candidates until none is left -- the states are volatile shared memory, a decision made by any warp
        // is seen by the others on their next look, and the undecided candidate of highest priority can always be
        // decided, so the loops terminate.  (The round count is capped all the same: a stuck CTA must not hang the GPU.)
        // The undecided candidates are first gathered into a dense list (a warp that walks the slots pays a full look
        // whenever ANY of its lanes holds an undecided one, so sparse work costs as much as dense work), grouped in BANDS
        // of descending score (linear in the score bits) when there are many: a candidate only depends on higher
        // priorities, i.e. on earlier bands -- decided, a block-wide barrier separates the bands -- and on its own band,
        // so a round only touches candidates that can actually be decided (r = 4, top_k = 4096: ~4 000 undecided
        // candidates with chains of dependent decisions; 224 k -> measured in profiles/r02_sparse_phases.txt).
        bool stuck = false;
        {
            const int n_und = s_und[0];
            const uint32_t k_lo = (uint32_t)s_und[1], k_hi = (uint32_t)s_und[2];
            int n_bands = n_und / 512;
            n_bands = n_bands < 1 ? 1 : (n_bands > 32 ? 32 : n_bands);
            int bshift = 0;
            while (n_bands > 1 && ((k_hi - k_lo) >> bshift) >= (uint32_t)n_bands) ++bshift;
            auto band_of = [&](uint64_t key) { return n_bands > 1 ? (int)(((uint32_t)(key >> 32) - k_lo) >> bshift) : 0; };
            if (threadIdx.x < 33) s_band[threadIdx.x] = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < c; i += SP_NT)
                if (state[i] == ST_UNDEC) atomicAdd(&s_band[band_of(keys[i]) + 1], 1);
            __syncthreads();
            if (threadIdx.x == 0) {                                   // s_band[k] = first list entry of band k (ascending bands)
                for (int k = 1; k <= 32; ++k) s_band[k] += s_band[k - 1];
                for (int k = 0; k < 32; ++k) s_fill[k] = s_band[k];
            }
            __syncthreads();
            for (int i = threadIdx.x; i < c; i += SP_NT)
                if (state[i] == ST_UNDEC) und[atomicAdd(&s_fill[band_of(keys[i])], 1)] = (uint16_t)i;
            __syncthreads();
            if (p.prof && b == 0 && threadIdx.x == 0) g_sparse_prof[13] = clock64() - g_sparse_prof[3];     // list building
            // One look = the ~5-30 candidates of the 3x3 cells around a candidate, a chain of dependent shared-memory reads
            // and compares per entry (~200 cycles each): EIGHT lanes share one candidate and split its neighbours, so a
            // look takes two or three such steps instead of fifteen (ncu: the serial scan of the busiest warp was the
            // whole decision phase at r = 4, top_k = 4096).
            const int lane = threadIdx.x & 31, grp = lane >> 3, gl = lane & 7;
            const int team = (threadIdx.x >> 5) * 4 + grp;            // 128 teams of eight lanes
            for (int band = n_bands - 1; band >= 0; --band) {
                const int e0 = s_band[band], e1 = s_band[band + 1];
                bool pending = true;
                int rounds = 0;
                while (pending) {
                    bool mine_waits = false;
                    for (int base = e0; base < e1; base += 4 * (SP_NT / 32)) {      // uniform trip count over the CTA
                        const int e = base + team;
                        int i = -1;
                        if (e < e1) { i = und[e]; if (state[i] != ST_UNDEC) i = -1; }
                        bool blocked = false, wait = false;
                        if (i >= 0) {
                            const uint64_t key = keys[i];
                            const uint32_t q = pos[i];
                            const int x = (int)(q & 0xffffu), y = (int)(q >> 16);
                            const int cx = x >> p.cell_shift, cy = y >> p.cell_shift;
                            const int cx0 = max(cx - 1, 0), cx1 = min(cx + 1, p.gw - 1);
                            for (int yy = max(cy - 1, 0); yy <= min(cy + 1, p.gh - 1); ++yy) {
                                const int lo = (int)cstart[yy * p.gw + cx0], hi = (int)cstart[yy * p.gw + cx1 + 1];
                                KB_ASSERT(lo >= 0 && lo <= hi && hi <= c);
                                for (int t = lo + gl; t < hi; t += 8) {
                                    const uint32_t qj = pos[t];
                                    const uint64_t kj = keys[t];
                                    const uint8_t sj = state[t];
                                    const int dx = (int)(qj & 0xffffu) - x, dy = (int)(qj >> 16) - y;
                                    if (dx > r || dx < -r || dy > r || dy < -r || kj <= key) continue;   // far, lower priority or itself
                                    blocked |= (sj == ST_KEPT);
                                    wait |= (sj == ST_UNDEC);
                                }
                            }
                        }
                        // the team's verdict (all 32 lanes are here: the ballots are warp-wide, the teams read their byte)
                        const unsigned bb = (__ballot_sync(0xffffffffu, blocked) >> (8 * grp)) & 0xffu;
                        const unsigned ww = (__ballot_sync(0xffffffffu, wait) >> (8 * grp)) & 0xffu;
                        if (i >= 0) {
                            if (bb) { if (gl == 0) state[i] = ST_DEAD; }
                            else if (!ww) { if (gl == 0) state[i] = ST_KEPT; }
                            else mine_waits = true;
                        }
                    }
                    pending = __any_sync(0xffffffffu, mine_waits);
                    if (++rounds > 100000) { stuck = true; break; }
                }
                if (p.prof && b == 0) { int mr = __reduce_max_sync(0xffffffffu, rounds); if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(&g_sparse_prof[15]), (unsigned long long)mr); if (threadIdx.x == 0 && band == n_bands - 1) g_sparse_prof[14] = n_bands; }
                if (n_bands > 1) __syncthreads();                   // the band is decided before the next one starts
            }
        }
        if (__syncthreads_or(stuck)) { fallback(); return; }

        KB_SP_PROF(4);
        // ---- kept interior candidates, compacted to the front of keys[] ---------------------------
        int n_ki = 0;
        for (int base = 0; base < c; base += SP_NT) {
            const int i = base + threadIdx.x;
            bool interior = false;
            uint64_t key = 0ull;
            if (i < c && state[i] == ST_KEPT) {
                const uint32_t q = pos[i];
                const int x = (int)(q & 0xffffu), y = (int)(q >> 16);
                interior = x >= p.border && x < W - p.border && y >= p.border && y < H - p.border;
                key = keys[i];
            }
            int tot;
            const int off = n_ki + kb::block_exclusive_scan(interior ? 1 : 0, s_scan, &tot);
            if (interior) keys[off] = key;          // off <= i: never clobbers an unread candidate
            n_ki += tot;
        }
        __syncthreads();

        KB_SP_PROF(5);
        if (n_ki > p.top_k) {
            // K > top_k: rows sorted by score descending (extracter.py:217-218), canonical tie order
            // pos[] is dead now: 48 KB behind keys[] that take the bucket-sorted keys
            uint64_t* sorted = reinterpret_cast<uint64_t*>(pos);
            constexpr int SORTED_CAP = SMEM_CAP / 2;               // pos[] alone: 6144 keys (cstart, the scratch, follows it)
            bool done = false;
            if (n_ki <= SORTED_CAP) done = bucket_sort_desc(keys, sorted, n_ki, p.top_k, cstart, s_scan, s_part);
            if (!done) {
                int n2 = 1;
                while (n2 < n_ki) n2 <<= 1;
                for (int i = n_ki + threadIdx.x; i < n2; i += SP_NT) keys[i] = 0ull;
                __syncthreads();
                sort_desc(keys, n2, SMEM_CAP);
                sorted = keys;
            }
            KB_SP_PROF(6);
            int cnt = 0;
            for (int i = threadIdx.x; i < p.top_k; i += SP_NT) {
                const uint64_t key = sorted[i];
                const float sc = kb::key_score(key);
                // rows with score <= min_score form a suffix of the sorted list (extracter.py:219-220)
                if (!(p.min_score > 0.0f) || sc > p.min_score) { emit(p, b, i, sc, kb::key_raster(key)); ++cnt; }
            }
            const int tot = block_sum(cnt, s_part);
            if (threadIdx.x == 0) { p.count[b] = tot; if (p.path) p.path[b] = 1; }
            KB_SP_PROF(7);
            if (p.prof && b == 0 && threadIdx.x == 0) { g_sparse_prof[8] = c; g_sparse_prof[9] = n_ki; g_sparse_prof[10] = nM; g_sparse_prof[11] = nO; g_sparse_prof[12] = attempt; }
            return;
        }
        if (cut_complete) {
            // K <= top_k: raster order (extracter.py:217), then the min_score filter (extracter.py:219-220)
            int n2b = 1;
            while (n2b < n_ki) n2b <<= 1;
            for (int i = threadIdx.x; i < n2b; i += SP_NT) {
                uint64_t k2 = 0ull;
                if (i < n_ki) {
                    const uint64_t key = keys[i];
                    k2 = ((uint64_t)(0xffffffffu - kb::key_raster(key)) << 32) | (key >> 32);
                }
                keys[i] = k2;
            }
            __syncthreads();
            sort_desc(keys, n2b, SMEM_CAP);
            int n_out = 0;
            for (int base = 0; base < n_ki; base += SP_NT) {
                const int i = base + threadIdx.x;
                bool keep = false;
                float sc = 0.f;
                uint32_t ras = 0;
                if (i < n_ki) {
                    const uint64_t k2 = keys[i];
                    ras = 0xffffffffu - (uint32_t)(k2 >> 32);
                    sc = kb::float_from_order_key((uint32_t)(k2 & 0xffffffffu));
                    keep = !(p.min_score > 0.0f) || sc > p.min_score;
                }
                int tot;
                const int off = n_out + kb::block_exclusive_scan(keep ? 1 : 0, s_scan, &tot);
                if (keep) emit(p, b, off, sc, ras);
                n_out += tot;
            }
            if (threadIdx.x == 0) { p.count[b] = n_out; if (p.path) p.path[b] = 1; }
            return;
        }
        if (tkey == 0u) break;          // everything listed was used and it is still not enough
        __syncthreads();
    }
    fallback();
}

template <int R>
static int launch_round1(const SparseParams& p, cudaStream_t st) {
    const size_t smem = Tile<R>::smem_bytes();
    KB_CUDA_TRY(cudaFuncSetAttribute(round1_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.W + DTW - 1) / DTW, (p.H + DTH - 1) / DTH, p.B);
    round1_kernel<R><<<grid, DNT, smem, st>>>(p);
    KB_LAUNCH_CHECK();
    return KB_OK;
}

constexpr size_t sparse_smem_bytes() {
    return (size_t)SMEM_CAP * 8 + (size_t)SMEM_CAP * 4 + (size_t)(MAX_CELLS + 1) * 4 + SMEM_CAP + (size_t)SMEM_CAP * 2 + 64;
}

}  // namespace kbsparse

// ------------------------------------------------------------------------------------------------
// host side (called from kb_detect in kb_select.cu)
// ------------------------------------------------------------------------------------------------
struct KbSparsePlan {
    bool ok;
    int cell_shift, gw, gh, c_pix;
};

static KbSparsePlan kb_sparse_plan(int H, int W, int nms_dist, int top_k) {
    KbSparsePlan pl{false, 0, 0, 0, 0};
    if (nms_dist < 1 || nms_dist > 8) return pl;
    if (H >= 32768 || W >= 32768 || H < 1 || W < 1) return pl;
    if (top_k < 1 || top_k + 1 > kbsparse::SMEM_CAP / 2) return pl;
    int sh = 2;                                         // cell edge >= max(r, 4), a power of two
    while ((1 << sh) < nms_dist) ++sh;
    while ((long long)((W >> sh) + 1) * ((H >> sh) + 1) > kbsparse::MAX_CELLS) ++sh;
    pl.cell_shift = sh;
    pl.gw = (W >> sh) + 1;
    pl.gh = (H >> sh) + 1;
    long long cp = (long long)H * W / 12;               // pixels listed per map: ~1/12 of the map ...
    const long long lo = 6LL * top_k, hi = 40000;       // ... at least 6*top_k, at most 40000
    if (cp < lo) cp = lo;
    if (cp > hi) cp = hi;
    pl.c_pix = (int)cp;
    pl.ok = true;
    return pl;
}

bool kb_sparse_supported(int H, int W, int nms_dist, int top_k) { return kb_sparse_plan(H, W, nms_dist, top_k).ok; }

size_t kb_sparse_workspace_bytes(int B, int H, int W, int nms_dist, int top_k) {
    if (!kb_sparse_plan(H, W, nms_dist, top_k).ok) return 0;
    return 2 * kb_align_up((size_t)B * kbsparse::LIST_CAP * sizeof(uint64_t), 256) + 8 * kb_align_up((size_t)B * 4, 256) + 1024;
}

// Runs the sparse path for all B maps.  need_fallback[B] / any_fallback[1] (device) report what is left.
int kb_sparse_detect(const float* score, int B, int H, int W, int nms_dist, int border, float threshold,
                     float min_score, int top_k, float* xyp, int* raster, int* count, int* path,
                     int** need_fallback_out, int** any_fallback_out, void* ws, size_t ws_bytes, int phases,
                     cudaStream_t st) {
    using namespace kbsparse;
    const KbSparsePlan pl = kb_sparse_plan(H, W, nms_dist, top_k);
    if (!pl.ok) return KB_ERR_UNSUPPORTED;
    if (B > 65535) return KB_ERR_UNSUPPORTED;
    KbArena arena(ws, ws_bytes);
    SparseParams p;
    p.listM = arena.take<uint64_t>((size_t)B * LIST_CAP);
    p.listO = arena.take<uint64_t>((size_t)B * LIST_CAP);
    p.tau = arena.take<float>(B);
    p.qscale = arena.take<float>(B);
    p.cntM = arena.take<int>(B);
    p.cntO = arena.take<int>(B);
    p.flags = arena.take<int>(B);
    p.need_fallback = arena.take<int>(B);
    p.any_fallback = arena.take<int>(1);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    p.score = score; p.xyp = xyp; p.raster = raster; p.count = count; p.path = path;
    p.B = B; p.H = H; p.W = W; p.r = nms_dist; p.border = border; p.top_k = top_k; p.c_pix = pl.c_pix;
    p.cell_shift = pl.cell_shift; p.gw = pl.gw; p.gh = pl.gh;
    p.threshold = threshold; p.min_score = min_score;
    p.prof = kb_knobs[KB_KNOB_SPARSE_PROF];
    *need_fallback_out = p.need_fallback;
    *any_fallback_out = p.any_fallback;
    // `phases` (bit 0 tau, bit 1 round-1, bit 2 sparse resolve) lets the benchmark time one kernel alone on a
    // workspace that an earlier full call has filled; every product call passes 7
    if (phases & 1) {
        tau_kernel<<<B, TAU_NT, 0, st>>>(p);
        KB_LAUNCH_CHECK();
    }
    // Round 1 has three kernels with identical lists: the packed streaming kernel of kb_round1_packed.cu (pairs of maps
    // as half2; the default whenever the batch is large enough to give every CTA a long band; phases bit 5 forces it),
    // the tiled round1_kernel (small batches, very wide maps; bit 3) and the fp32 streaming kernel of
    // kb_round1_stream.cu (bit 4; measured equal to the tiled one at 480x640 r = 6, slower elsewhere -- DESIGN.md).
    int rc = (phases & 2) ? KB_ERR_UNSUPPORTED : KB_OK;
    if ((phases & 2) && (phases & 16)) rc = launch_round1_stream(p, true, st);
    else if ((phases & 2) && !(phases & 8)) rc = launch_round1_packed(p, (phases & 32) != 0, st);
    if ((phases & 2) && rc == KB_ERR_UNSUPPORTED) switch (nms_dist) {
        case 1: rc = launch_round1<1>(p, st); break;
        case 2: rc = launch_round1<2>(p, st); break;
        case 3: rc = launch_round1<3>(p, st); break;
        case 4: rc = launch_round1<4>(p, st); break;
        case 5: rc = launch_round1<5>(p, st); break;
        case 6: rc = launch_round1<6>(p, st); break;
        case 7: rc = launch_round1<7>(p, st); break;
        case 8: rc = launch_round1<8>(p, st); break;
        default: break;
    }
    if (rc != KB_OK) return rc;
    if (!(phases & 4)) return KB_OK;
    const size_t smem = sparse_smem_bytes();
    KB_CUDA_TRY(cudaFuncSetAttribute(sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sparse_kernel<<<B, SP_NT, smem, st>>>(p);
    KB_LAUNCH_CHECK();
    return KB_OK;
}

extern "C" int kb_debug_sparse_prof(long long* host_out) {
    if (!host_out) return KB_ERR_BAD_ARG;
    KB_CUDA_TRY(cudaMemcpyFromSymbol(host_out, kbsparse::g_sparse_prof, 16 * sizeof(long long)));
    return KB_OK;
}
