// Round-faithful score-map NMS (reference: utils/extracter.py:6-100), one cooperative launch.
//
// Each round is two tile sweeps separated by grid-wide barriers:
//   A. find the pixels that are the first maximal entry of their zero-padded (2r+1)^2 window
//      (separable: full-row window maxima of the rows above / below, half-row maxima left / right;
//      strictly greater than everything earlier in raster order, >= everything later);
//   B. overwrite every pixel that has another maximum within Chebyshev distance r.
// The loop ends on the device when the batch-wide count of maxima repeats (extracter.py:73-78),
// so there is no host round trip per round (the reference syncs the host every round).
#include <cooperative_groups.h>
#include "kb_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int TH = 32;        // tile rows
constexpr int TW = 64;        // tile cols
constexpr int NT = 256;       // threads per block
constexpr int MAX_R = 24;     // largest supported nms_dist

constexpr int MAX_GROUPS = 2048;   // maps per launch when every map stops on its own count

struct NmsParams {
    float* v;                 // [B,H,W] working map (in place)
    const float* src;         // optional: copied into v (active maps only) before the first round
    uint8_t* mask;            // [B,H,W] maxima of the current round
    unsigned long long* cnt;  // [2,G] maxima counters (ping-pong by round parity)
    const int* active;        // optional [B]: maps with active[b]==0 are left untouched
    const int* any_active;    // optional [1]: 0 = nothing to do (uniform early exit)
    int* rounds_out;          // may be null
    int B, H, W, r, max_iter;
    int per_map;              // 0: one joint count over the batch (fast_nms); 1: every map stops alone
    float min_value;
};

__device__ __forceinline__ void tile_coords(int t, int tiles_x, int tiles_y, int& b, int& ty, int& tx) {
    tx = t % tiles_x;
    int q = t / tiles_x;
    ty = q % tiles_y;
    b = q / tiles_y;
}

// Sweep A for one tile.  Returns the number of maxima found by this thread's pixels.
__device__ int sweep_find_maxima(const NmsParams& p, int b, int y0, int x0, float* S, float* HF) {
    const int r = p.r, H = p.H, W = p.W;
    const int SW = TW + 2 * r;          // staged width
    const int SH = TH + 2 * r;          // staged height
    const int SP = SW + 1;              // row pitch (odd: fewer bank conflicts on column walks)
    const float* img = p.v + (size_t)b * H * W;
    for (int i = threadIdx.x; i < SH * SW; i += NT) {
        int sy = i / SW, sx = i - sy * SW;
        int gy = y0 + sy - r, gx = x0 + sx - r;
        float val = 0.0f;                                   // zero padding (extracter.py:58)
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) val = img[(size_t)gy * W + gx];
        S[sy * SP + sx] = val;
    }
    __syncthreads();
    // full-width horizontal window max for every staged row at the tile's columns
    for (int i = threadIdx.x; i < SH * TW; i += NT) {
        int sy = i / TW, x = i - sy * TW;
        const float* row = S + sy * SP + x;                 // window = row[0 .. 2r]
        float m = row[0];
        for (int d = 1; d <= 2 * r; ++d) m = fmaxf(m, row[d]);
        HF[sy * TW + x] = m;
    }
    __syncthreads();
    int found = 0;
    uint8_t* mk = p.mask + (size_t)b * H * W;
    for (int i = threadIdx.x; i < TH * TW; i += NT) {
        int y = i / TW, x = i - y * TW;
        int gy = y0 + y, gx = x0 + x;
        if (gy >= H || gx >= W) continue;
        const float c = S[(y + r) * SP + (x + r)];
        // entries earlier in raster order: rows above (full width) + same row, left part
        float before = HF[y * TW + x];
        for (int d = 1; d < r; ++d) before = fmaxf(before, HF[(y + d) * TW + x]);
        const float* row = S + (y + r) * SP + x;
        for (int d = 0; d < r; ++d) before = fmaxf(before, row[d]);
        // entries later in raster order: same row, right part + rows below
        float after = HF[(y + r + 1) * TW + x];
        for (int d = 2; d <= r; ++d) after = fmaxf(after, HF[(y + r + d) * TW + x]);
        for (int d = 1; d <= r; ++d) after = fmaxf(after, row[r + d]);
        const bool is_max = (c > before) && (c >= after);
        mk[(size_t)gy * W + gx] = is_max ? 1 : 0;
        found += is_max ? 1 : 0;
    }
    __syncthreads();
    return found;
}

// Sweep B for one tile: suppress pixels that see another maximum within Chebyshev distance r.
__device__ void sweep_suppress(const NmsParams& p, int b, int y0, int x0, uint8_t* M, uint8_t* RS) {
    const int r = p.r, H = p.H, W = p.W;
    const int SW = TW + 2 * r, SH = TH + 2 * r;
    const uint8_t* mk = p.mask + (size_t)b * H * W;
    for (int i = threadIdx.x; i < SH * SW; i += NT) {
        int sy = i / SW, sx = i - sy * SW;
        int gy = y0 + sy - r, gx = x0 + sx - r;
        uint8_t val = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) val = mk[(size_t)gy * W + gx];
        M[sy * SW + sx] = val;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SH * TW; i += NT) {
        int sy = i / TW, x = i - sy * TW;
        const uint8_t* row = M + sy * SW + x;
        uint8_t any = 0;
        for (int d = 0; d <= 2 * r; ++d) any |= row[d];
        RS[sy * TW + x] = any;
    }
    __syncthreads();
    float* img = p.v + (size_t)b * H * W;
    for (int i = threadIdx.x; i < TH * TW; i += NT) {
        int y = i / TW, x = i - y * TW;
        int gy = y0 + y, gx = x0 + x;
        if (gy >= H || gx >= W) continue;
        uint8_t any = 0;
        for (int d = 0; d <= 2 * r; ++d) any |= RS[(y + d) * TW + x];
        // two maxima are never within r of each other, so "another maximum in the window"
        // == "some maximum in the window and I am not one"
        if (any && !M[(y + r) * SW + (x + r)]) img[(size_t)gy * W + gx] = p.min_value;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT) nms_rounds_kernel(NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    if (p.any_active && *p.any_active == 0) return;          // uniform: nothing was flagged
    const int r = p.r;
    const int SW = TW + 2 * r, SH = TH + 2 * r;
    const int G = p.per_map ? p.B : 1;
    // per-group bookkeeping lives at the front of shared memory (every block keeps its own copy and
    // reaches the same decisions, so no extra grid barrier is needed to publish them)
    unsigned long long* g_seen = reinterpret_cast<unsigned long long*>(smem_raw);
    int* g_found = reinterpret_cast<int*>(g_seen + G);
    uint8_t* g_done = reinterpret_cast<uint8_t*>(g_found + G);
    unsigned char* tile_mem = smem_raw + (((size_t)G * 13 + 15) / 16) * 16;
    float* S = reinterpret_cast<float*>(tile_mem);
    float* HF = S + SH * (SW + 1);
    uint8_t* M = tile_mem;                      // sweep B reuses the same bytes
    uint8_t* RS = M + SH * SW;

    const int tiles_x = (p.W + TW - 1) / TW, tiles_y = (p.H + TH - 1) / TH;
    const int tiles_per_map = tiles_x * tiles_y;
    const int n_tiles = p.B * tiles_per_map;
    for (int g = threadIdx.x; g < G; g += NT) {
        g_seen[g] = ~0ull;                      // "count = None" (extracter.py:45)
        bool act = true;
        if (p.active) {
            act = false;
            if (p.per_map) act = p.active[g] != 0;
            else for (int b = 0; b < p.B; ++b) act |= p.active[b] != 0;
        }
        g_done[g] = act ? 0 : 1;
    }
    __syncthreads();
    if (p.src) {                                // working copy of the active maps
        const size_t npx = (size_t)p.H * p.W;
        for (int b = 0; b < p.B; ++b) {
            if (p.active && !p.active[b]) continue;
            const float* s = p.src + (size_t)b * npx;
            float* d = p.v + (size_t)b * npx;
            for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < npx; i += (size_t)gridDim.x * NT) d[i] = s[i];
        }
        grid.sync();
    }
    int round = 0;
    while (round != p.max_iter) {
        const int par = round & 1;
        for (int g = threadIdx.x; g < G; g += NT) g_found[g] = 0;
        __syncthreads();
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int b, ty, tx;
            tile_coords(t, tiles_x, tiles_y, b, ty, tx);
            const int g = p.per_map ? b : 0;
            if (g_done[g] || (p.active && !p.active[b])) continue;
            int found = sweep_find_maxima(p, b, ty * TH, tx * TW, S, HF);
            for (int d = 16; d > 0; d >>= 1) found += __shfl_xor_sync(0xffffffffu, found, d);
            if ((threadIdx.x & 31) == 0 && found) atomicAdd(&g_found[g], found);
        }
        __syncthreads();
        for (int g = threadIdx.x; g < G; g += NT)
            if (g_found[g]) atomicAdd(&p.cnt[(size_t)par * G + g], (unsigned long long)g_found[g]);
        grid.sync();
        int still = 0;
        for (int g = threadIdx.x; g < G; g += NT) {
            if (!g_done[g]) {
                const unsigned long long now = *((volatile unsigned long long*)&p.cnt[(size_t)par * G + g]);
                if (now == g_seen[g]) g_done[g] = 1;    // extracter.py:76-77 (tested before suppressing)
                else { g_seen[g] = now; still = 1; }
            }
            if (blockIdx.x == 0) p.cnt[(size_t)(par ^ 1) * G + g] = 0ull;   // next round's counter
        }
        if (!__syncthreads_or(still)) break;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int b, ty, tx;
            tile_coords(t, tiles_x, tiles_y, b, ty, tx);
            const int g = p.per_map ? b : 0;
            if (g_done[g] || (p.active && !p.active[b])) continue;
            sweep_suppress(p, b, ty * TH, tx * TW, M, RS);
        }
        ++round;
        grid.sync();
    }
    if (p.rounds_out && blockIdx.x == 0 && threadIdx.x == 0) *p.rounds_out = round;
}

size_t nms_smem_bytes(int r, int groups) {
    const size_t SW = TW + 2 * r, SH = TH + 2 * r;
    size_t a = (SH * (SW + 1) + SH * TW) * sizeof(float);
    size_t b = SH * SW + SH * TW;
    return (a > b ? a : b) + (((size_t)groups * 13 + 15) / 16) * 16;
}

}  // namespace

extern "C" size_t kb_fast_nms_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return kb_align_up((size_t)B * H * W, 256) + kb_align_up((size_t)2 * (B > 1 ? B : 1) * 8, 256) + 256;
}

// Shared with kb_detect (kb_select.cu): run the rounds in place on `v` (optionally seeded from `src`).
// per_map = 0 reproduces fast_nms on a batch (one joint count); per_map = 1 treats every map as its own
// fast_nms call.  `active` / `any_active` restrict the work to flagged maps (device-side decision).
int kb_nms_rounds_inplace(float* v, const float* src, int B, int H, int W, int nms_dist, int max_iter,
                          float min_value, int per_map, const int* active, const int* any_active, int* rounds,
                          void* ws, size_t ws_bytes, cudaStream_t st) {
    if (nms_dist > MAX_R) return KB_ERR_UNSUPPORTED;
    const int G = per_map ? B : 1;
    if (G > MAX_GROUPS) return KB_ERR_UNSUPPORTED;
    KbArena arena(ws, ws_bytes);
    NmsParams p;
    p.mask = arena.take<uint8_t>((size_t)B * H * W);
    p.cnt = arena.take<unsigned long long>((size_t)2 * G);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    p.v = v; p.src = src; p.active = active; p.any_active = any_active; p.per_map = per_map;
    p.rounds_out = rounds;
    p.B = B; p.H = H; p.W = W; p.r = nms_dist; p.max_iter = max_iter; p.min_value = min_value;
    KB_CUDA_TRY(cudaMemsetAsync(p.cnt, 0, (size_t)2 * G * sizeof(unsigned long long), st));
    const size_t smem = nms_smem_bytes(nms_dist, G);
    KB_CUDA_TRY(cudaFuncSetAttribute(nms_rounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, occ = 0;
    KB_CUDA_TRY(cudaGetDevice(&dev));
    KB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    KB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nms_rounds_kernel, NT, smem));
    if (occ < 1) return KB_ERR_UNSUPPORTED;
    const int tiles = B * ((W + TW - 1) / TW) * ((H + TH - 1) / TH);
    int grid = sms * occ;
    if (grid > tiles) grid = tiles;
    void* args[] = {&p};
    KB_CUDA_TRY(cudaLaunchCooperativeKernel((void*)nms_rounds_kernel, dim3(grid), dim3(NT), args, smem, st));
    return KB_OK;
}

extern "C" int kb_fast_nms(const float* score, float* out, int B, int H, int W, int nms_dist, int max_iter,
                           float min_value, int* rounds, void* ws, size_t ws_bytes, kb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!score || !out || B <= 0 || H <= 0 || W <= 0 || nms_dist < 0) return KB_ERR_BAD_ARG;
    if (score == out) return KB_ERR_BAD_ARG;
    if (nms_dist == 0 || max_iter == 0) {       // extracter.py:40-41 / :50-51
        KB_CUDA_TRY(cudaMemcpyAsync(out, score, (size_t)B * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (rounds) KB_CUDA_TRY(cudaMemsetAsync(rounds, 0, sizeof(int), st));
        return KB_OK;
    }
    return kb_nms_rounds_inplace(out, score, B, H, W, nms_dist, max_iter, min_value, /*per_map=*/0, nullptr,
                                 nullptr, rounds, ws, ws_bytes, st);
}
