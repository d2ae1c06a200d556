// Tensor Lucas-Kanade matcher            utils/matcher.py:7-142 (OpticalFlow), 188-203
//
// The reference unfolds win x win patches of both images and of the Sobel gradients of the second
// one into C*win^2-channel maps (1.6 GB per map at 480x640, win 21) and bilinearly samples them at
// the track positions, 40 Gauss-Newton iterations per pyramid level.  Here nothing is unfolded: one
// CTA owns one keypoint for the whole coarse-to-fine schedule, keeps the template patch in shared
// memory and evaluates entry (c,u,v) of the unfolded maps as the 4-corner blend of shifted pixels
// (a corner outside the map contributes nothing, like grid_sample's zero padding on the unfolded
// map).  Pyramid levels (avg_pool2d of the original image, kernel 2j) and the per-channel Sobel
// maps of the second image are built once per call by two small kernels; they stay L2-resident.
#include "kb_common.cuh"
#include <math.h>

namespace {

constexpr int LK_NT = 128;
constexpr int LK_MAX_LEVELS = 8;

struct LkLevel {
    const float* i0;     // [B,C,Hl,Wl] first image at this level
    const float* i1;     // second image
    const float* dx;     // Sobel of the second image
    const float* dy;
    int H, W;
};

struct LkParams {
    LkLevel lv[LK_MAX_LEVELS];
    const float* pts0;   // [B,n_max,2] pixels (full resolution)
    const float* init;   // [B,n_max,2] pixels
    const int* count;    // [B] or null
    float* out;          // [B,n_max,2]
    int B, C, n_max, win, levels, iterations;
};

// avg_pool2d(img, kernel=k, stride=k): running sum in row-major order, divided by k*k (matcher.py:44-46).
__global__ void lk_avgpool_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int H, int W, int k) {
    const int ho = H / k, wo = W / k;
    const size_t total = (size_t)planes * ho * wo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % wo);
        const int y = (int)((i / wo) % ho);
        const size_t pl = i / ((size_t)wo * ho);
        const float* src = in + pl * H * W + (size_t)y * k * W + (size_t)x * k;
        float acc = 0.0f;
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) acc += src[(size_t)a * W + b];
        out[i] = acc / (float)(k * k);
    }
}

// conv2d(img, dx/dy, padding=1) with the per-channel Sobel kernels (matcher.py:22-35, 104-109).
__global__ void lk_sobel_kernel(const float* __restrict__ in, float* __restrict__ dx, float* __restrict__ dy, int planes,
                                int H, int W) {
    const size_t total = (size_t)planes * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const float* pl = in + (i / ((size_t)W * H)) * H * W;
        auto px = [&](int yy, int xx) -> float {
            return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? pl[(size_t)yy * W + xx] : 0.0f;
        };
        const float tl = px(y - 1, x - 1), tc = px(y - 1, x), tr = px(y - 1, x + 1);
        const float ml = px(y, x - 1), mr = px(y, x + 1);
        const float bl = px(y + 1, x - 1), bc = px(y + 1, x), br = px(y + 1, x + 1);
        dx[i] = (tl - tr) + 2.0f * (ml - mr) + (bl - br);
        dy[i] = (tl + 2.0f * tc + tr) - (bl + 2.0f * bc + br);
    }
}

struct Taps {
    int x0, y0;
    float w[4];      // nw, ne, sw, se; 0 for corners outside the map
    bool ok[4];
};

// grid_sample's un-normalisation of pts / (W-1, H-1) * 2 - 1 with align_corners=True (matcher.py:126-127, 132).
__device__ __forceinline__ Taps make_taps(float px, float py, int H, int W) {
    Taps t;
    const float gx = px / (float)(W - 1) * 2.0f - 1.0f, gy = py / (float)(H - 1) * 2.0f - 1.0f;
    const float ix = ((gx + 1.0f) / 2.0f) * (float)(W - 1), iy = ((gy + 1.0f) / 2.0f) * (float)(H - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const float x1 = fx + 1.0f, y1 = fy + 1.0f;
    t.w[0] = (x1 - ix) * (y1 - iy);
    t.w[1] = (ix - fx) * (y1 - iy);
    t.w[2] = (x1 - ix) * (iy - fy);
    t.w[3] = (ix - fx) * (iy - fy);
    const bool fin = isfinite(ix) && isfinite(iy) && fabsf(ix) < 1e9f && fabsf(iy) < 1e9f;
    t.x0 = fin ? (int)fx : -(1 << 20);
    t.y0 = fin ? (int)fy : -(1 << 20);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int xc = t.x0 + (c & 1), yc = t.y0 + (c >> 1);
        t.ok[c] = fin && xc >= 0 && xc < W && yc >= 0 && yc < H;
    }
    return t;
}

// Entry (u,v) of the unfolded map of `plane`, bilinearly sampled: sum over the in-map corners of
// w * plane[yc + u - pad, xc + v - pad] (zero outside the image: unfold's padding).
__device__ __forceinline__ float tap_patch(const float* __restrict__ plane, const Taps& t, int du, int dv, int H, int W) {
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int xx = t.x0 + (c & 1) + dv, yy = t.y0 + (c >> 1) + du;
        if (t.ok[c] && xx >= 0 && xx < W && yy >= 0 && yy < H) acc += __ldg(plane + (size_t)yy * W + xx) * t.w[c];
    }
    return acc;
}

__global__ void __launch_bounds__(LK_NT) lk_track_kernel(LkParams p) {
    extern __shared__ float tmpl[];                  // [C*win*win] template patch of the current level
    __shared__ double red[2][LK_NT / 32][5];
    const int b = blockIdx.y, i = blockIdx.x;
    const int n = p.count ? p.count[b] : p.n_max;
    if (i >= n) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int win = p.win, pad = win / 2, K = p.C * win * win;
    const size_t row = ((size_t)b * p.n_max + i) * 2;
    const float p0x = p.pts0[row], p0y = p.pts0[row + 1];
    float cx = p.init[row], cy = p.init[row + 1];
    int par = 0;
    for (int li = 0; li < p.levels; ++li) {
        const int lvl = p.levels - 1 - li;
        const float sc = (float)(1 << lvl);
        const LkLevel L = p.lv[lvl];
        const size_t plane = (size_t)L.H * L.W, img = (size_t)b * p.C * plane;
        cx /= sc; cy /= sc;
        {
            const Taps t = make_taps(p0x / sc, p0y / sc, L.H, L.W);
            for (int e = tid; e < K; e += LK_NT) {
                const int c = e / (win * win), r = e - c * win * win, u = r / win, v = r - u * win;
                tmpl[e] = tap_patch(L.i0 + img + c * plane, t, u - pad, v - pad, L.H, L.W);
            }
        }
        __syncthreads();
        for (int it = 0; it < p.iterations; ++it) {
            const Taps t = make_taps(cx, cy, L.H, L.W);
            double gxx = 0.0, gxy = 0.0, gyy = 0.0, bx = 0.0, by = 0.0;
            for (int e = tid; e < K; e += LK_NT) {
                const int c = e / (win * win), r = e - c * win * win, u = r / win, v = r - u * win;
                const size_t off = img + c * plane;
                const float v1 = tap_patch(L.i1 + off, t, u - pad, v - pad, L.H, L.W);
                const float jx = tap_patch(L.dx + off, t, u - pad, v - pad, L.H, L.W);
                const float jy = tap_patch(L.dy + off, t, u - pad, v - pad, L.H, L.W);
                const float di = tmpl[e] - v1;
                gxx += (double)(jx * jx); gxy += (double)(jx * jy); gyy += (double)(jy * jy);
                bx += (double)(di * jx); by += (double)(di * jy);
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                gxx += __shfl_xor_sync(0xffffffffu, gxx, s);
                gxy += __shfl_xor_sync(0xffffffffu, gxy, s);
                gyy += __shfl_xor_sync(0xffffffffu, gyy, s);
                bx += __shfl_xor_sync(0xffffffffu, bx, s);
                by += __shfl_xor_sync(0xffffffffu, by, s);
            }
            if (lane == 0) {
                red[par][warp][0] = gxx; red[par][warp][1] = gxy; red[par][warp][2] = gyy;
                red[par][warp][3] = bx; red[par][warp][4] = by;
            }
            __syncthreads();
            double tot[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                tot[q] = 0.0;
#pragma unroll
                for (int wq = 0; wq < LK_NT / 32; ++wq) tot[q] += red[par][wq][q];
            }
            par ^= 1;
            // G and b as the reference holds them (float32), det / inverse from those
            const float Gxx = (float)tot[0], Gxy = (float)tot[1], Gyy = (float)tot[2];
            const float Bx = (float)tot[3], By = (float)tot[4];
            const double det = (double)Gxx * (double)Gyy - (double)Gxy * (double)Gxy;
            if ((float)det > 1e-6f) {                                   // matcher.py:137
                const float i00 = (float)((double)Gyy / det), i01 = (float)(-(double)Gxy / det);
                const float i11 = (float)((double)Gxx / det);
                // einsum('bik,bk->bk') of the reference (matcher.py:139): row sum of the inverse times b
                cx -= i00 * Bx + i01 * Bx;
                cy -= i01 * By + i11 * By;
            }
        }
        cx *= sc; cy *= sc;
        __syncthreads();                                                // tmpl is rewritten by the next level
    }
    if (tid == 0) {
        p.out[row] = cx;
        p.out[row + 1] = cy;
    }
}

int level_dims(int H, int W, int j, int* h, int* w) {
    if (j == 0) { *h = H; *w = W; return 1; }
    *h = H / (2 * j); *w = W / (2 * j);
    return *h > 1 && *w > 1;
}

}  // namespace

extern "C" KB_API size_t kb_lk_workspace_bytes(int B, int C, int H, int W, int levels) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || levels <= 0 || levels > LK_MAX_LEVELS) return 0;
    size_t total = 256;
    for (int j = 0; j < levels; ++j) {
        int h, w;
        if (!level_dims(H, W, j, &h, &w)) return 0;
        const size_t sz = kb_align_up((size_t)B * C * h * w * sizeof(float), 256);
        total += (j == 0 ? 2 : 4) * sz;            // level 0 reads the caller's images; dx, dy everywhere
    }
    return total;
}

extern "C" KB_API int kb_lk_track(const float* img0, const float* img1, int B, int C, int H, int W, const float* pts0_px,
                                  const float* init_px, const int* count, int n_max, int win_size, int levels,
                                  int iterations, float* out_px, void* ws, size_t ws_bytes, kb_stream_t stream) {
    if (B < 0 || C <= 0 || H <= 1 || W <= 1 || n_max < 0 || win_size <= 0 || (win_size & 1) == 0 || levels <= 0 ||
        levels > LK_MAX_LEVELS || iterations < 0)
        return KB_ERR_BAD_ARG;
    if (B == 0 || n_max == 0) return KB_OK;
    if (!img0 || !img1 || !pts0_px || !init_px || !out_px) return KB_ERR_BAD_ARG;
    const size_t smem = (size_t)C * win_size * win_size * sizeof(float);
    if (smem > 200 * 1024) return KB_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    KbArena arena(ws, ws_bytes);
    LkParams p{};
    for (int j = 0; j < levels; ++j) {
        int h, w;
        if (!level_dims(H, W, j, &h, &w)) return KB_ERR_BAD_ARG;
        const size_t elems = (size_t)B * C * h * w;
        float* a0 = nullptr; float* a1 = nullptr;
        if (j > 0) { a0 = arena.take<float>(elems); a1 = arena.take<float>(elems); }
        float* gx = arena.take<float>(elems);
        float* gy = arena.take<float>(elems);
        if (!arena.ok()) return KB_ERR_WORKSPACE;
        p.lv[j] = LkLevel{j ? a0 : img0, j ? a1 : img1, gx, gy, h, w};
        const int blocks = (int)((elems + 255) / 256 < 148 * 16 ? (elems + 255) / 256 : 148 * 16);
        if (j > 0) {
            lk_avgpool_kernel<<<blocks, 256, 0, st>>>(img0, a0, B * C, H, W, 2 * j);
            lk_avgpool_kernel<<<blocks, 256, 0, st>>>(img1, a1, B * C, H, W, 2 * j);
        }
        lk_sobel_kernel<<<blocks, 256, 0, st>>>(p.lv[j].i1, gx, gy, B * C, h, w);
    }
    p.pts0 = pts0_px; p.init = init_px; p.count = count; p.out = out_px;
    p.B = B; p.C = C; p.n_max = n_max; p.win = win_size; p.levels = levels; p.iterations = iterations;
    KB_CUDA_TRY(cudaFuncSetAttribute(lk_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lk_track_kernel<<<dim3(n_max, B), LK_NT, smem, st>>>(p);
    KB_LAUNCH_CHECK();
    return KB_OK;
}
