// Shared declarations of the sparse exact detection path (kb_sparse_nms.cu, kb_round1_stream.cu).
#pragma once
#include "kb_common.cuh"

namespace kbsparse {

constexpr int SAMPLES = 4096;
constexpr int TAU_NT = 256;
constexpr int DTH = 32, DTW = 128, DNT = 256;     // round-1 tile
constexpr int SP_NT = 1024;                       // sparse kernel threads
constexpr int MAX_CELLS = 8192;
constexpr int LIST_CAP = 16384;                   // entries per list and map
constexpr int SMEM_CAP = 12288;                   // candidates resolved in shared memory per map

struct SparseParams {
    const float* score;       // [B,H,W]
    float* tau;               // [B]
    float* qscale;            // [B] scale of the packed kernel's 16-bit score image: fp16((score - tau) * qscale)
    uint64_t* listM;          // [B,LIST_CAP] round-1 maxima above tau
    uint64_t* listO;          // [B,LIST_CAP] uncovered pixels above tau
    int* cntM;                // [B]
    int* cntO;                // [B]
    int* flags;               // [B] bit0: has negative score
    int* need_fallback;       // [B] (out) 1 = run the round-faithful path for this map
    int* any_fallback;        // [1]
    float* xyp;               // [B,top_k,3]
    int* raster;              // [B,top_k]
    int* count;               // [B]
    int* path;                // [B] or null
    int B, H, W, r, border, top_k, c_pix;
    int cell_shift, gw, gh;   // coarse grid of the sparse stage
    float threshold, min_score;
    int prof;                 // KB_KNOB_SPARSE_PROF: map 0's CTA records clock64 at its phase boundaries
};


#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// (2R+1)-window maxima shared by the round-1 kernels
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// out[o] = max over input rows o .. o+2R of a thread's 4 columns, o = 0..NS-1; ld(i) returns input row i (0..NS+2R-1).
// All windows share rows NS-1 .. 2R (when there are any); the rest are suffixes of rows 0 .. NS-2 and prefixes of
// rows 2R+1 .. : ~4 maximum operations per element at NS = 8, R = 6.  V / mx: the vector type and its maximum (float4 +
// max4 for the fp32 kernels, four packed half2 pairs for the packed kernel).
template <int R, int NS, typename V, typename Max, typename Load>
__device__ __forceinline__ void window_max_rows_t(Load ld, Max mx, V (&out)[NS]) {
    if constexpr (2 * R >= NS - 1 && NS >= 2) {
        V run = ld(NS - 2);
        out[NS - 2] = run;
#pragma unroll
        for (int j = NS - 3; j >= 0; --j) { run = mx(run, ld(j)); out[j] = run; }
        V core = ld(NS - 1);
#pragma unroll
        for (int i = NS; i <= 2 * R; ++i) core = mx(core, ld(i));
#pragma unroll
        for (int o = 0; o < NS - 1; ++o) out[o] = mx(out[o], core);
        out[NS - 1] = core;
        run = ld(2 * R + 1);
        out[1] = mx(out[1], run);
#pragma unroll
        for (int o = 2; o < NS; ++o) { run = mx(run, ld(2 * R + o)); out[o] = mx(out[o], run); }
    } else {
        V in[NS + 2 * R];
#pragma unroll
        for (int i = 0; i < NS + 2 * R; ++i) in[i] = ld(i);
#pragma unroll
        for (int o = 0; o < NS; ++o) {
            V m = in[o];
#pragma unroll
            for (int d = 1; d <= 2 * R; ++d) m = mx(m, in[o + d]);
            out[o] = m;
        }
    }
}

template <int R, int NS, typename Load>
__device__ __forceinline__ void window_max_rows(Load ld, float4 (&out)[NS]) {
    window_max_rows_t<R, NS, float4>(ld, [](float4 a, float4 b) { return max4(a, b); }, out);
}

// (2R+1)-window maximum along the row for a thread's 4 columns x4..x4+3; vmrow points at column 0 of a row that is
// readable (and zero) for 8 columns either side of the data; `own` = the row's values at x4..x4+3.
template <int R>
__device__ __forceinline__ float4 window_max_cols(const float* vmrow, int x4, float4 own) {
    constexpr int NBR = (R + 3) / 4, C = 4 * NBR;
    float a[4 * (2 * NBR + 1)];
#pragma unroll
    for (int nb = -NBR; nb <= NBR; ++nb) {
        const float4 q = nb == 0 ? own : *reinterpret_cast<const float4*>(vmrow + x4 + 4 * nb);
        a[C + 4 * nb + 0] = q.x; a[C + 4 * nb + 1] = q.y; a[C + 4 * nb + 2] = q.z; a[C + 4 * nb + 3] = q.w;
    }
    float4 r;
    if constexpr (R >= 2) {
        float core = a[C + 3 - R];
#pragma unroll
        for (int i = C + 4 - R; i <= C + R; ++i) core = fmaxf(core, a[i]);
        const float l2 = a[C + 2 - R], l1 = fmaxf(l2, a[C + 1 - R]), l0 = fmaxf(l1, a[C - R]);
        const float r1 = a[C + R + 1], r2 = fmaxf(r1, a[C + R + 2]), r3 = fmaxf(r2, a[C + R + 3]);
        r.x = fmaxf(core, l0);
        r.y = fmaxf(fmaxf(core, l1), r1);
        r.z = fmaxf(fmaxf(core, l2), r2);
        r.w = fmaxf(core, r3);
    } else {
        r.x = fmaxf(fmaxf(a[C - 1], a[C]), a[C + 1]);
        r.y = fmaxf(fmaxf(a[C], a[C + 1]), a[C + 2]);
        r.z = fmaxf(fmaxf(a[C + 1], a[C + 2]), a[C + 3]);
        r.w = fmaxf(fmaxf(a[C + 2], a[C + 3]), a[C + 4]);
    }
    return r;
}
#endif

// kb_round1_stream.cu: the streaming form of round 1 (full-width bands), an alternative to the tiled round1_kernel
// with identical lists.  Returns KB_ERR_UNSUPPORTED when the map is too wide for it or (without `force`) the batch
// too small to give every CTA a long band.
int launch_round1_stream(const SparseParams& p, bool force, cudaStream_t st);

// kb_round1_packed.cu: streaming round 1 over PAIRS of maps held as packed half2 (a 16-bit monotone image of the scores
// decides almost everything; fp32 only for the `score > tau` test and for ties at 16 bits), identical lists.  Returns
// KB_ERR_UNSUPPORTED when the map is too wide or (without `force`) the batch too small.
int launch_round1_packed(const SparseParams& p, bool force, cudaStream_t st);

}  // namespace kbsparse
