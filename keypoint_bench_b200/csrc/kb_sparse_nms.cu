// Sparse exact detection path: NMS + border + threshold + top-k for non-negative score maps.
//
// The fixed point of the reference's fast_nms rounds (utils/extracter.py:49-98) on a non-negative map
// is greedy NMS over the total order (score desc, raster asc) with Chebyshev radius r, and whether a
// pixel is kept depends only on pixels of HIGHER priority.  detection() (extracter.py:193-221) keeps
// the top_k interior survivors, so only the highest-scoring pixels can matter:
//
//   tau_kernel      per map: a 4096-pixel sample picks tau so that ~C_target pixels exceed it;
//   extract_kernel  ONE streaming pass over the maps (the HBM-bound kernel: float4 loads, 1 read of
//                   every pixel): pixels > tau are appended to the map's candidate list as 64-bit
//                   priority keys; also flags maps containing negative scores (those need the
//                   round-faithful kernel because the reference's stop rule is then not monotone);
//   greedy_kernel   one CTA per map: bitonic-sort the candidates by priority in shared memory, then
//                   greedy NMS in that order against a 1-bit-per-pixel "kept" bitmap in shared memory,
//                   1024 candidates per step (in-step conflicts resolved by a short fixed-point loop),
//                   stopping as soon as top_k+1 interior survivors exist.
//
// A map is certified when it produced more than top_k interior survivors (K > top_k: output sorted by
// priority) or when its candidate list was complete (tau == threshold: output in raster order,
// extracter.py:217).  Anything else is flagged for the round-faithful kernel (kb_nms.cu).
#include "kb_common.cuh"

namespace kbsparse {

constexpr int SAMPLES = 4096;
constexpr int TAU_NT = 256;
constexpr int EX_NT = 256;
constexpr int GR_NT = 1024;

struct SparseParams {
    const float* score;       // [B,H,W]
    float* tau;               // [B]
    int* cand_count;          // [B]
    uint64_t* cand;           // [B,cap]
    int* flags;               // [B] bit0: has negative score
    int* need_fallback;       // [B] (out) 1 = run the round-faithful path for this map
    int* any_fallback;        // [1]
    float* xyp;               // [B,top_k,3]
    int* raster;              // [B,top_k]
    int* count;               // [B]
    int* path;                // [B] or null
    int B, H, W, r, border, top_k, cap, c_target;
    float threshold, min_score;
};

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TAU_NT) tau_kernel(SparseParams p) {
    __shared__ uint32_t s_keys[SAMPLES];
    __shared__ int s_hist[256];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ int s_need;
    const int b = blockIdx.x;
    const long long npx = (long long)p.H * p.W;
    const float theta = fmaxf(p.threshold, 0.0f);
    if (threadIdx.x == 0) { p.cand_count[b] = 0; p.flags[b] = 0; p.need_fallback[b] = 0; }
    if (b == 0 && threadIdx.x == 0) *p.any_fallback = 0;
    // rank of the sample that estimates the C_target-th largest pixel
    const long long rank = ((long long)p.c_target * SAMPLES + npx - 1) / npx;
    if (npx <= SAMPLES || rank >= SAMPLES / 2) {        // small map / dense request: take everything
        if (threadIdx.x == 0) p.tau[b] = theta;
        return;
    }
    const float* img = p.score + (size_t)b * npx;
    const long long stride = npx / SAMPLES;
    for (int k = threadIdx.x; k < SAMPLES; k += TAU_NT) {
        uint32_t h = (uint32_t)k * 2654435761u;
        h ^= h >> 15;
        const long long idx = (long long)k * stride + (long long)(h % (uint32_t)stride);
        s_keys[k] = kb::float_order_key(img[idx]);
    }
    if (threadIdx.x == 0) { s_prefix = 0u; s_mask = 0u; s_need = (int)rank; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        s_hist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t pre = s_prefix, msk = s_mask;
        for (int k = threadIdx.x; k < SAMPLES; k += TAU_NT) {
            const uint32_t key = s_keys[k];
            if ((key & msk) == pre) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int need = s_need, cum = 0, d = 255;
            for (; d > 0; --d) {
                if (cum + s_hist[d] >= need) break;
                cum += s_hist[d];
            }
            s_need = need - cum;
            s_prefix = pre | ((uint32_t)d << shift);
            s_mask = msk | (255u << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float t = kb::float_from_order_key(s_prefix);
        p.tau[b] = (t > theta) ? t : theta;             // NaN-safe: falls back to theta
    }
}

// ------------------------------------------------------------------------------------------------
// Streaming candidate extraction.  grid = (chunks per map, B); every thread owns 4 consecutive floats
// per iteration (float4 when the map's byte offset allows it).
constexpr int EX_ITEMS = 4;
constexpr int EX_CHUNK = EX_NT * EX_ITEMS * 8;      // floats per CTA

__global__ void __launch_bounds__(EX_NT) extract_kernel(SparseParams p) {
    const int b = blockIdx.y;
    const long long npx = (long long)p.H * p.W;
    const float* img = p.score + (size_t)b * npx;
    const float tau = p.tau[b];
    uint64_t* out = p.cand + (size_t)b * p.cap;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(img) & 15u) == 0);
    const long long base0 = (long long)blockIdx.x * EX_CHUNK;
    bool neg = false;
    const int lane = threadIdx.x & 31;
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
        const long long i0 = base0 + ((long long)it * EX_NT + threadIdx.x) * EX_ITEMS;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        int nvalid = 0;
        if (i0 + 3 < npx && vec_ok) {
            const float4 q = __ldcs(reinterpret_cast<const float4*>(img + i0));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            nvalid = 4;
        } else if (i0 < npx) {
            nvalid = (int)((npx - i0) < 4 ? (npx - i0) : 4);
            for (int e = 0; e < nvalid; ++e) v[e] = __ldcs(img + i0 + e);
        }
        int c = 0;
        unsigned hit = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool ok = e < nvalid;
            neg |= ok && (v[e] < 0.0f);
            if (ok && v[e] > tau) { hit |= 1u << e; ++c; }
        }
        // warp-aggregated append (order inside the list is irrelevant: it is sorted later)
        int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const int wtot = __shfl_sync(0xffffffffu, inc, 31);
        if (wtot) {
            int wbase = 0;
            if (lane == 31) wbase = atomicAdd(&p.cand_count[b], wtot);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            int off = wbase + inc - c;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (hit & (1u << e)) {
                    if (off < p.cap) out[off] = kb::priority_key(v[e], (uint32_t)(i0 + e));
                    ++off;
                }
            }
        }
    }
    if (__any_sync(0xffffffffu, neg) && lane == 0) atomicOr(&p.flags[b], 1);
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit(const SparseParams& p, int b, int slot, float score, uint32_t ras) {
    const int row = ras / p.W, col = ras - row * p.W;
    float* o = p.xyp + ((size_t)b * p.top_k + slot) * 3;
    o[0] = ((float)col + 0.5f) / (float)p.W;          // extracter.py:149,158
    o[1] = ((float)row + 0.5f) / (float)p.H;
    o[2] = score;
    p.raster[(size_t)b * p.top_k + slot] = (int)ras;
}

__device__ void bitonic_desc(uint64_t* a, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += GR_NT) {
                const int l = i ^ j;
                if (l > i) {
                    const uint64_t x = a[i], y = a[l];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
    }
}

constexpr uint32_t ST_DEAD = 0u, ST_UNDEC = 1u, ST_KEPT = 2u;

__global__ void __launch_bounds__(GR_NT, 1) greedy_kernel(SparseParams p, int maxc /*pow2*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int H = p.H, W = p.W, r = p.r;
    const int Ww = (W + 31) >> 5;
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(keys + maxc);
    uint32_t* chunk = bitmap + (size_t)H * Ww;             // GR_NT packed (state<<30 | y<<15 | x)
    __shared__ int s_scan[33];

    const int total = p.cand_count[b];
    const bool neg = (p.flags[b] & 1) != 0;
    if (neg || total > maxc || total > p.cap) {
        if (threadIdx.x == 0) { p.need_fallback[b] = 1; atomicExch(p.any_fallback, 1); }
        return;
    }
    const int c = total;
    const uint64_t* src = p.cand + (size_t)b * p.cap;
    int n2 = 1;
    while (n2 < c) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += GR_NT) keys[i] = (i < c) ? src[i] : 0ull;
    for (int i = threadIdx.x; i < H * Ww; i += GR_NT) bitmap[i] = 0u;
    __syncthreads();
    bitonic_desc(keys, n2);

    int n_emit = 0;     // interior survivors so far, compacted in priority order into keys[0..n_emit)
    for (int base = 0; base < c && n_emit <= p.top_k; base += GR_NT) {
        const int i = base + threadIdx.x;
        const bool valid = i < c;
        uint64_t key = 0ull;
        int x = 0, y = 0;
        bool alive = false;
        if (valid) {
            key = keys[i];
            const uint32_t ras = kb::key_raster(key);
            y = ras / W;
            x = ras - y * W;
            // any survivor of an earlier step within Chebyshev distance r?
            const int xl = max(x - r, 0), xh = min(x + r, W - 1);
            const int wl = xl >> 5, wh = xh >> 5;
            const uint32_t ml = 0xffffffffu << (xl & 31);
            const uint32_t mh = 0xffffffffu >> (31 - (xh & 31));
            uint32_t any = 0u;
            const int yl = max(y - r, 0), yh = min(y + r, H - 1);
            for (int yy = yl; yy <= yh; ++yy) {
                const uint32_t* row = bitmap + yy * Ww;
                if (wl == wh) {
                    any |= row[wl] & ml & mh;
                } else {
                    any |= (row[wl] & ml) | (row[wh] & mh);
                    for (int w = wl + 1; w < wh; ++w) any |= row[w];
                }
            }
            alive = (any == 0u);
        }
        chunk[threadIdx.x] = ((alive ? ST_UNDEC : ST_DEAD) << 30) | ((uint32_t)y << 15) | (uint32_t)x;
        __syncthreads();
        // conflicts with higher-priority candidates of the same step (rare: record up to 4, rescan if more)
        int conf[4] = {-1, -1, -1, -1};
        int nconf = 0;
        if (alive) {
            for (int j = 0; j < (int)threadIdx.x; ++j) {
                const uint32_t o = chunk[j];
                if ((o >> 30) == ST_DEAD) continue;
                const int ox = (int)(o & 0x7fffu), oy = (int)((o >> 15) & 0x7fffu);
                if (abs(ox - x) <= r && abs(oy - y) <= r) {
                    if (nconf < 4) conf[nconf] = j;
                    ++nconf;
                }
            }
        }
        uint32_t st = alive ? ST_UNDEC : ST_DEAD;
        // fixed-point loop: a candidate is kept once every conflicting earlier candidate is dead
        while (true) {
            uint32_t nst = st;
            if (st == ST_UNDEC) {
                bool blocked = false, wait = false;
                if (nconf <= 4) {
                    for (int q = 0; q < nconf; ++q) {
                        const uint32_t s = chunk[conf[q]] >> 30;
                        blocked |= (s == ST_KEPT);
                        wait |= (s == ST_UNDEC);
                    }
                } else {
                    for (int j = 0; j < (int)threadIdx.x; ++j) {
                        const uint32_t o = chunk[j];
                        const uint32_t s = o >> 30;
                        if (s == ST_DEAD) continue;
                        const int ox = (int)(o & 0x7fffu), oy = (int)((o >> 15) & 0x7fffu);
                        if (abs(ox - x) <= r && abs(oy - y) <= r) {
                            blocked |= (s == ST_KEPT);
                            wait |= (s == ST_UNDEC);
                        }
                    }
                }
                nst = blocked ? ST_DEAD : (wait ? ST_UNDEC : ST_KEPT);
            }
            __syncthreads();
            if (nst != st) chunk[threadIdx.x] = (nst << 30) | ((uint32_t)y << 15) | (uint32_t)x;
            st = nst;
            if (!__syncthreads_or(st == ST_UNDEC)) break;
        }
        const bool kept = (st == ST_KEPT);
        if (kept) atomicOr(&bitmap[y * Ww + (x >> 5)], 1u << (x & 31));
        const bool interior = kept && x >= p.border && x < W - p.border && y >= p.border && y < H - p.border;
        int tot;
        const int off = n_emit + kb::block_exclusive_scan(interior ? 1 : 0, s_scan, &tot);
        if (interior) keys[off] = key;          // off <= i: never clobbers an unread candidate
        n_emit += tot;
        __syncthreads();
    }

    const float theta = fmaxf(p.threshold, 0.0f);
    const bool complete = (p.tau[b] == theta) && (p.threshold >= 0.0f);
    if (n_emit > p.top_k) {
        // K > top_k: rows sorted by score descending (extracter.py:217-218), canonical tie order
        int cnt = 0;
        for (int i = threadIdx.x; i < p.top_k; i += GR_NT) {
            const uint64_t key = keys[i];
            const float sc = kb::key_score(key);
            // rows with score <= min_score form a suffix of the sorted list (extracter.py:219-220)
            if (!(p.min_score > 0.0f) || sc > p.min_score) { emit(p, b, i, sc, kb::key_raster(key)); ++cnt; }
        }
        int tot;
        kb::block_exclusive_scan(cnt, s_scan, &tot);
        if (threadIdx.x == 0) { p.count[b] = tot; if (p.path) p.path[b] = 1; }
    } else if (complete) {
        // K <= top_k: raster order (extracter.py:217), then the min_score filter (extracter.py:219-220)
        int n2b = 1;
        while (n2b < n_emit) n2b <<= 1;
        for (int i = threadIdx.x; i < n2b; i += GR_NT) {
            uint64_t k2 = 0ull;
            if (i < n_emit) {
                const uint64_t key = keys[i];
                k2 = ((uint64_t)(0xffffffffu - kb::key_raster(key)) << 32) | (key >> 32);
            }
            keys[i] = k2;
        }
        __syncthreads();
        bitonic_desc(keys, n2b);
        int n_out = 0;
        for (int base = 0; base < n_emit; base += GR_NT) {
            const int i = base + threadIdx.x;
            bool keep = false;
            float sc = 0.f;
            uint32_t ras = 0;
            if (i < n_emit) {
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
    } else {
        if (threadIdx.x == 0) { p.need_fallback[b] = 1; atomicExch(p.any_fallback, 1); }
    }
}

}  // namespace kbsparse

// ------------------------------------------------------------------------------------------------
// host side (called from kb_detect in kb_select.cu)
// ------------------------------------------------------------------------------------------------
struct KbSparsePlan {
    int maxc;          // candidate capacity per map (power of two), 0 = sparse path not applicable
    int c_target;
    size_t smem;
};

KbSparsePlan kb_sparse_plan(int H, int W, int top_k) {
    KbSparsePlan pl{0, 0, 0};
    if (H >= 32768 || W >= 32768) return pl;
    const size_t bitmap = (size_t)H * ((W + 31) / 32) * 4;
    const size_t fixed = bitmap + kbsparse::GR_NT * 4 + 1024;
    const size_t budget = 220 * 1024;
    if (fixed + 2048 * 8 > budget) return pl;
    int maxc = 2048;
    while ((size_t)maxc * 2 * 8 + fixed <= budget && maxc < 16384 && maxc < 8 * top_k) maxc <<= 1;
    if (maxc < top_k + 2) return pl;
    pl.maxc = maxc;
    long long ct = 4LL * top_k;
    if (ct > (long long)maxc * 6 / 10) ct = (long long)maxc * 6 / 10;
    if (ct < top_k + 2) ct = top_k + 2;
    pl.c_target = (int)ct;
    pl.smem = (size_t)maxc * 8 + fixed;
    return pl;
}

size_t kb_sparse_workspace_bytes(int B, int H, int W, int top_k) {
    KbSparsePlan pl = kb_sparse_plan(H, W, top_k);
    if (!pl.maxc) return 0;
    return kb_align_up((size_t)B * pl.maxc * sizeof(uint64_t), 256) + 5 * kb_align_up((size_t)B * 4, 256) + 1024;
}

// Runs the sparse path for all B maps.  need_fallback[B] / any_fallback[1] (device) report what is left.
int kb_sparse_detect(const float* score, int B, int H, int W, int nms_dist, int border, float threshold,
                     float min_score, int top_k, float* xyp, int* raster, int* count, int* path,
                     int** need_fallback_out, int** any_fallback_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    using namespace kbsparse;
    KbSparsePlan pl = kb_sparse_plan(H, W, top_k);
    if (!pl.maxc) return KB_ERR_UNSUPPORTED;
    KbArena arena(ws, ws_bytes);
    SparseParams p;
    p.cand = arena.take<uint64_t>((size_t)B * pl.maxc);
    p.tau = arena.take<float>(B);
    p.cand_count = arena.take<int>(B);
    p.flags = arena.take<int>(B);
    p.need_fallback = arena.take<int>(B);
    p.any_fallback = arena.take<int>(1);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    p.score = score; p.xyp = xyp; p.raster = raster; p.count = count; p.path = path;
    p.B = B; p.H = H; p.W = W; p.r = nms_dist; p.border = border; p.top_k = top_k; p.cap = pl.maxc;
    p.c_target = pl.c_target; p.threshold = threshold; p.min_score = min_score;
    *need_fallback_out = p.need_fallback;
    *any_fallback_out = p.any_fallback;
    if (B > 65535) return KB_ERR_UNSUPPORTED;
    tau_kernel<<<B, TAU_NT, 0, st>>>(p);
    KB_LAUNCH_CHECK();
    const long long npx = (long long)H * W;
    dim3 grid((unsigned)((npx + EX_CHUNK - 1) / EX_CHUNK), B);
    extract_kernel<<<grid, EX_NT, 0, st>>>(p);
    KB_LAUNCH_CHECK();
    KB_CUDA_TRY(cudaFuncSetAttribute(greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    greedy_kernel<<<B, GR_NT, pl.smem, st>>>(p, pl.maxc);
    KB_LAUNCH_CHECK();
    return KB_OK;
}
