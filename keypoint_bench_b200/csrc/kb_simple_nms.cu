// simple_nms(scores, nms_radius)                       models/lightglue.py:904-920
//
// The max-pool NMS of the LightGlue-style extractor: M = (s == pool(s)); twice: supp = pool(M) > 0,
// ss = supp ? 0 : s, M |= (ss == pool(ss)) & ~supp; out = M ? s : 0, where pool is the (2r+1)^2
// maximum with -inf outside the image (max_pool2d's implicit padding).
//
// Three launches of one kernel (the first computes M from s alone, the last also writes `out`).  The
// mask travels between launches bit-packed (one 32-pixel word per 128 bytes of scores).  A CTA owns
// a 32x64 output tile: it loads the mask words of the tile + 2r rows (and one word left / right),
// dilates them by r with shifts and ORs (rows) and a (2r+1)-row OR (columns) to get `supp` on the
// tile + r halo, forms ss there, and max-pools ss separably with every thread producing a run of
// four outputs from 4 + 2r inputs held in registers (the window core is shared by the four).
// HBM per launch: 4 B/px scores (+ 1 bit mask in and out; + 4 B/px on the last launch).
#include "kb_common.cuh"
#include <math.h>

namespace {

constexpr int TH = 32, TW = 64, NT = 256;
constexpr int MAX_R = 16;

struct SnmsParams {
    const float* s;
    const uint32_t* m_in;        // [B,H,WW] bit mask of the previous launch, null on the first one
    uint32_t* m_out;
    float* out;                  // non-null on the last launch
    int H, W, WW;
};

// out[i] = max(in[i .. i + 2R]) for i = 0..3
template <int R>
__device__ __forceinline__ void run_max4(const float (&in)[4 + 2 * R + 3], float (&o)[4]) {
    if constexpr (R == 0) {
        o[0] = in[0]; o[1] = in[1]; o[2] = in[2]; o[3] = in[3];
    } else if constexpr (R == 1) {
        o[0] = fmaxf(fmaxf(in[0], in[1]), in[2]);
        o[1] = fmaxf(fmaxf(in[1], in[2]), in[3]);
        o[2] = fmaxf(fmaxf(in[2], in[3]), in[4]);
        o[3] = fmaxf(fmaxf(in[3], in[4]), in[5]);
    } else {
        float core = in[3];                                   // in[3 .. 2R] belongs to all four windows
#pragma unroll
        for (int k = 4; k <= 2 * R; ++k) core = fmaxf(core, in[k]);
        const float pre2 = in[2], pre1 = fmaxf(in[1], pre2), pre0 = fmaxf(in[0], pre1);
        const float post1 = in[2 * R + 1], post2 = fmaxf(post1, in[2 * R + 2]), post3 = fmaxf(post2, in[2 * R + 3]);
        o[0] = fmaxf(core, pre0);
        o[1] = fmaxf(fmaxf(core, pre1), post1);
        o[2] = fmaxf(fmaxf(core, pre2), post2);
        o[3] = fmaxf(core, post3);
    }
}

template <int R>
__global__ void __launch_bounds__(NT) simple_nms_kernel(SnmsParams p) {
    constexpr int AH = TH + 4 * R;                    // mask rows needed (2R above and below)
    constexpr int BH = TH + 2 * R, BW = TW + 2 * R;   // region of ss / supp
    constexpr int NIN = 4 + 2 * R;                    // inputs of a run of four outputs
    constexpr int NV4 = (NIN + 3) / 4;                // float4 loads of a horizontal run
    constexpr int SSP = (TW - 4) + 4 * NV4 + 4;       // row pitch of ss: multiple of 4, covers the last run's loads
    __shared__ uint32_t MB[AH][4];                    // mask bits, columns x0-32 .. x0+TW+31
    __shared__ uint32_t DB[AH][4];                    // dilated along rows
    __shared__ uint32_t SB[BH][4];                    // supp bits (rows y0-R ..)
    __shared__ __align__(16) float ss[BH * SSP];
    __shared__ __align__(16) float RM[BH * TW];       // row maxima

    const int b = blockIdx.z;
    const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const int tid = threadIdx.x;
    const bool first = p.m_in == nullptr;
    const size_t map = (size_t)b * p.H * p.W;
    const size_t mmap = (size_t)b * p.H * p.WW;
    const int w0 = (x0 >> 5) - 1;                     // global word index of MB[.][0]

    if (!first) {
        for (int i = tid; i < AH * 4; i += NT) {
            const int ya = i >> 2, w = i & 3;
            const int gy = y0 - 2 * R + ya, gw = w0 + w;
            MB[ya][w] = (gy >= 0 && gy < p.H && gw >= 0 && gw < p.WW) ? p.m_in[mmap + (size_t)gy * p.WW + gw] : 0u;
        }
        __syncthreads();
        for (int i = tid; i < AH * 4; i += NT) {
            const int ya = i >> 2, w = i & 3;
            const uint32_t cur = MB[ya][w], prev = w > 0 ? MB[ya][w - 1] : 0u, next = w < 3 ? MB[ya][w + 1] : 0u;
            uint32_t acc = cur;
#pragma unroll
            for (int d = 1; d <= R; ++d) acc |= (cur << d) | (prev >> (32 - d)) | (cur >> d) | (next << (32 - d));
            DB[ya][w] = acc;
        }
        __syncthreads();
        for (int i = tid; i < BH * 4; i += NT) {
            const int yb = i >> 2, w = i & 3;
            uint32_t acc = 0u;
#pragma unroll
            for (int d = 0; d <= 2 * R; ++d) acc |= DB[yb + d][w];
            SB[yb][w] = acc;
        }
        __syncthreads();
    }
    // ss on the tile + R halo; -inf outside the image (max_pool2d's padding) and in the pitch padding
    for (int i = tid; i < BH * SSP; i += NT) {
        const int yb = i / SSP, xb = i - yb * SSP;
        const int gy = y0 - R + yb, gx = x0 - R + xb;
        float v = -INFINITY;
        if (xb < BW && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
            v = p.s[map + (size_t)gy * p.W + gx];
            if (!first) {
                const int bit = xb - R + 32;                  // column relative to x0 - 32
                if ((SB[yb][bit >> 5] >> (bit & 31)) & 1u) v = 0.0f;      // torch.where(supp_mask, zeros, scores)
            }
        }
        ss[i] = v;
    }
    __syncthreads();
    // horizontal maxima: runs of four outputs
    for (int i = tid; i < BH * (TW / 4); i += NT) {
        const int yb = i / (TW / 4), xr = (i - yb * (TW / 4)) * 4;
        float in[NIN + 3];
        const float4* src = reinterpret_cast<const float4*>(ss + yb * SSP + xr);
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
            const float4 q = src[k];
            in[4 * k] = q.x; in[4 * k + 1] = q.y; in[4 * k + 2] = q.z; in[4 * k + 3] = q.w;
        }
        float o[4];
        run_max4<R>(in, o);
        *reinterpret_cast<float4*>(RM + yb * TW + xr) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // vertical maxima (runs of four rows), the new mask and, on the last launch, the output
    for (int i = tid; i < (TH / 4) * TW; i += NT) {
        const int yr = (i / TW) * 4, x = i - (i / TW) * TW;
        float in[NIN + 3];
#pragma unroll
        for (int k = 0; k < NIN; ++k) in[k] = RM[(yr + k) * TW + x];
        float o[4];
        run_max4<R>(in, o);
        const int gx = x0 + x;
        const int wsel = 1 + (x >> 5), bsel = x & 31;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int gy = y0 + yr + r;
            const bool inside = gy < p.H && gx < p.W;
            bool m = false;
            if (inside) {
                const bool supp = !first && ((SB[yr + r + R][wsel] >> bsel) & 1u);
                const bool old = !first && ((MB[yr + r + 2 * R][wsel] >> bsel) & 1u);
                m = old || (!supp && ss[(yr + r + R) * SSP + x + R] == o[r]);
            }
            const uint32_t word = __ballot_sync(0xffffffffu, m);
            const int gw = (x0 >> 5) + (x >> 5);
            if ((tid & 31) == 0 && gy < p.H && gw < p.WW) p.m_out[mmap + (size_t)gy * p.WW + gw] = word;
            if (p.out && inside) {
                const size_t o_idx = map + (size_t)gy * p.W + gx;
                p.out[o_idx] = m ? p.s[o_idx] : 0.0f;
            }
        }
    }
}

template <int R>
void launch3(const float* score, float* out, int B, int H, int W, uint32_t* m0, uint32_t* m1, cudaStream_t st) {
    const dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B);
    SnmsParams p{score, nullptr, m0, nullptr, H, W, (W + 31) / 32};
    simple_nms_kernel<R><<<grid, NT, 0, st>>>(p);                   // max_mask = scores == max_pool(scores)
    p.m_in = m0; p.m_out = m1;
    simple_nms_kernel<R><<<grid, NT, 0, st>>>(p);                   // first suppression pass
    p.m_in = m1; p.m_out = m0; p.out = out;
    simple_nms_kernel<R><<<grid, NT, 0, st>>>(p);                   // second pass + torch.where(max_mask, scores, 0)
}

}  // namespace

extern "C" KB_API size_t kb_simple_nms_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return 2 * kb_align_up((size_t)B * H * ((W + 31) / 32) * 4, 256) + 256;
}

extern "C" KB_API int kb_simple_nms(const float* score, float* out, int B, int H, int W, int nms_radius, void* ws,
                                    size_t ws_bytes, kb_stream_t stream) {
    if (B < 0 || H < 0 || W < 0 || nms_radius < 0 || nms_radius > MAX_R) return KB_ERR_BAD_ARG;
    if (B == 0 || H == 0 || W == 0) return KB_OK;
    if (!score || !out || score == out || B > 65535) return KB_ERR_BAD_ARG;
    KbArena arena(ws, ws_bytes);
    const size_t words = (size_t)B * H * ((W + 31) / 32);
    uint32_t* m0 = arena.take<uint32_t>(words);
    uint32_t* m1 = arena.take<uint32_t>(words);
    if (!arena.ok()) return KB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    switch (nms_radius) {
#define KB_SNMS_CASE(RR) case RR: launch3<RR>(score, out, B, H, W, m0, m1, st); break;
        KB_SNMS_CASE(0) KB_SNMS_CASE(1) KB_SNMS_CASE(2) KB_SNMS_CASE(3) KB_SNMS_CASE(4) KB_SNMS_CASE(5) KB_SNMS_CASE(6)
        KB_SNMS_CASE(7) KB_SNMS_CASE(8) KB_SNMS_CASE(9) KB_SNMS_CASE(10) KB_SNMS_CASE(11) KB_SNMS_CASE(12)
        KB_SNMS_CASE(13) KB_SNMS_CASE(14) KB_SNMS_CASE(15) KB_SNMS_CASE(16)
#undef KB_SNMS_CASE
        default: return KB_ERR_BAD_ARG;
    }
    KB_LAUNCH_CHECK();
    return KB_OK;
}
