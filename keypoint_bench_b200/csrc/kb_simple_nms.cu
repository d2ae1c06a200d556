// simple_nms(scores, nms_radius)                       models/lightglue.py:904-920
//
// The max-pool NMS of the LightGlue-style extractor: M = (s == pool(s)); twice: supp = pool(M) > 0,
// ss = supp ? 0 : s, M |= (ss == pool(ss)) & ~supp; out = M ? s : 0, where pool is the (2r+1)^2
// maximum with -inf outside the image (max_pool2d's implicit padding).
//
// Three launches of one kernel (the first computes M from s alone, the last also writes `out`).
// A CTA owns a 32x64 output tile: it stages the mask with a 2r halo, dilates it separably to get
// `supp` on the tile + r halo, forms ss there, and max-pools ss separably -- so the only intermediate
// that travels between launches is the byte mask (ping-pong; neighbours read the previous one).
// HBM per launch: 4 B/px read (+1 B mask read, 1 B mask write; +4 B on the last) -- the maps of a
// batch are L2-resident between launches up to ~25 M pixels.
#include "kb_common.cuh"
#include <math.h>

namespace {

constexpr int TH = 32, TW = 64, NT = 256;
constexpr int MAX_R = 16;

struct SnmsParams {
    const float* s;
    const unsigned char* m_in;   // null on the first launch
    unsigned char* m_out;
    float* out;                  // non-null on the last launch
    int H, W, r;
};

__global__ void __launch_bounds__(NT) simple_nms_kernel(SnmsParams p) {
    extern __shared__ unsigned char smem_raw[];
    const int r = p.r;
    const int MW = TW + 4 * r, MH = TH + 4 * r;      // mask region
    const int SW = TW + 2 * r, SH = TH + 2 * r;      // ss / supp region
    float* ss = reinterpret_cast<float*>(smem_raw);               // [SH][SW]
    float* rm = ss + SH * SW;                                     // [SH][TW] row maxima
    unsigned char* mk = reinterpret_cast<unsigned char*>(rm + SH * TW);   // [MH][MW]
    unsigned char* hm = mk + MH * MW;                             // [MH][SW] row-dilated mask
    unsigned char* sp = hm + MH * SW;                             // [SH][SW] supp

    const size_t map = (size_t)blockIdx.z * p.H * p.W;
    const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const int tid = threadIdx.x;
    const bool first = p.m_in == nullptr;

    if (!first) {
        for (int i = tid; i < MH * MW; i += NT) {
            const int yy = i / MW, xx = i - yy * MW;
            const int y = y0 - 2 * r + yy, x = x0 - 2 * r + xx;
            mk[i] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? p.m_in[map + (size_t)y * p.W + x] : 0;
        }
        __syncthreads();
        for (int i = tid; i < MH * SW; i += NT) {
            const int yy = i / SW, xx = i - yy * SW;
            unsigned char v = 0;
            for (int d = 0; d <= 2 * r; ++d) v |= mk[yy * MW + xx + d];
            hm[i] = v;
        }
        __syncthreads();
    }
    for (int i = tid; i < SH * SW; i += NT) {
        const int yy = i / SW, xx = i - yy * SW;
        const int y = y0 - r + yy, x = x0 - r + xx;
        unsigned char v = 0;
        if (!first)
            for (int d = 0; d <= 2 * r; ++d) v |= hm[(yy + d) * SW + xx];
        sp[i] = v;
        float s = -INFINITY;                                      // max_pool2d pads with -inf
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
            s = p.s[map + (size_t)y * p.W + x];
            if (v) s = 0.0f;                                      // torch.where(supp_mask, zeros, scores)
        }
        ss[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < SH * TW; i += NT) {
        const int yy = i / TW, xx = i - yy * TW;
        float v = ss[yy * SW + xx];
        for (int d = 1; d <= 2 * r; ++d) v = fmaxf(v, ss[yy * SW + xx + d]);
        rm[i] = v;
    }
    __syncthreads();
    for (int i = tid; i < TH * TW; i += NT) {
        const int yy = i / TW, xx = i - yy * TW;
        const int y = y0 + yy, x = x0 + xx;
        if (y >= p.H || x >= p.W) continue;
        float v = rm[yy * TW + xx];
        for (int d = 1; d <= 2 * r; ++d) v = fmaxf(v, rm[(yy + d) * TW + xx]);
        const int c = (yy + r) * SW + xx + r;
        const bool newmax = ss[c] == v && !sp[c];
        const unsigned char old = first ? 0 : mk[(yy + 2 * r) * MW + xx + 2 * r];
        const unsigned char m = old | (newmax ? 1 : 0);
        const size_t o = map + (size_t)y * p.W + x;
        p.m_out[o] = m;
        if (p.out) p.out[o] = m ? p.s[o] : 0.0f;
    }
}

size_t smem_bytes(int r) {
    const size_t MW = TW + 4 * r, MH = TH + 4 * r, SW = TW + 2 * r, SH = TH + 2 * r;
    return (SH * SW + SH * TW) * sizeof(float) + MH * MW + MH * SW + SH * SW;
}

}  // namespace

extern "C" KB_API size_t kb_simple_nms_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return 2 * kb_align_up((size_t)B * H * W, 256) + 256;
}

extern "C" KB_API int kb_simple_nms(const float* score, float* out, int B, int H, int W, int nms_radius, void* ws,
                                    size_t ws_bytes, kb_stream_t stream) {
    if (B < 0 || H < 0 || W < 0 || nms_radius < 0 || nms_radius > MAX_R) return KB_ERR_BAD_ARG;
    if (B == 0 || H == 0 || W == 0) return KB_OK;
    if (!score || !out || score == out) return KB_ERR_BAD_ARG;
    KbArena arena(ws, ws_bytes);
    unsigned char* m0 = arena.take<unsigned char>((size_t)B * H * W);
    unsigned char* m1 = arena.take<unsigned char>((size_t)B * H * W);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = smem_bytes(nms_radius);
    KB_CUDA_TRY(cudaFuncSetAttribute(simple_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B);
    SnmsParams p{score, nullptr, m0, nullptr, H, W, nms_radius};
    simple_nms_kernel<<<grid, NT, smem, st>>>(p);                   // max_mask = scores == max_pool(scores)
    p.m_in = m0; p.m_out = m1;
    simple_nms_kernel<<<grid, NT, smem, st>>>(p);                   // first suppression pass
    p.m_in = m1; p.m_out = m0; p.out = out;
    simple_nms_kernel<<<grid, NT, smem, st>>>(p);                   // second pass + torch.where(max_mask, scores, 0)
    KB_LAUNCH_CHECK();
    return KB_OK;
}
