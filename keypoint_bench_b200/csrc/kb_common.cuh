// Shared device/host helpers for the keypoint_bench hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/kb_b200.h"

#define KB_CUDA_TRY(expr)                         \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) return (int)_e;    \
    } while (0)

#define KB_LAUNCH_CHECK()                         \
    do {                                          \
        cudaError_t _e = cudaGetLastError();      \
        if (_e != cudaSuccess) return (int)_e;    \
    } while (0)

// Bounds checks of the debug build (`make debug` -> libkb_b200_dbg.so, -DKB_BOUNDS): compute-sanitizer is closed on the
// GPU pool, so the kernels carry their own index asserts; a failed one traps (the launch then returns an error instead
// of silently corrupting memory).  Compiled out of the release library.
#ifdef KB_BOUNDS
#define KB_ASSERT(cond) do { if (!(cond)) { printf("KB_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define KB_ASSERT(cond) do { } while (0)
#endif

// experiment knobs (kb_debug_knob, kb_api.cu); index = KB_KNOB_*
extern int kb_knobs[8];
// multiprocessor count of the current device (looked up once per device and process)
int kb_sm_count(int* sms);

// kb_match_tc.cu: where the fused sampler (kb_sample.cu) writes the matcher's operand rows
struct KbOperandSinks {
    unsigned short* S[2];     // [B*n_max, 2*Dp] 16-bit halves [hi | lo] of side 0 / side 1
    float* part[2];           // [B*n_max, Dp/8] partial |row|^2 per group of 8 components
    int Dp, fp16;
};
int kb_match_tc_operand_sinks(void* ws, size_t ws_bytes, int B, int n_max, int m_max, int D, KbOperandSinks* out);

static inline size_t kb_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace (the library never allocates).
struct KbArena {
    char* base;
    size_t cap;
    size_t off;
    __host__ KbArena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
    template <typename T>
    __host__ T* take(size_t n) {
        off = kb_align_up(off, 256);
        T* r = (T*)(base + off);
        off += n * sizeof(T);
        return r;
    }
    __host__ bool ok() const { return off <= cap && (base != nullptr || off == 0); }
};

#ifdef __CUDACC__
namespace kb {

constexpr int WARP = 32;

// Monotone map float -> uint32 (larger float => larger key), -0 < +0, NaN sorts high.
__device__ __forceinline__ uint32_t float_order_key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// Priority key of a pixel: higher key = earlier in (score desc, raster asc).
__device__ __forceinline__ uint64_t priority_key(float score, uint32_t raster) {
    return ((uint64_t)float_order_key(score) << 32) | (uint64_t)(0xffffffffu - raster);
}
__device__ __forceinline__ uint32_t key_raster(uint64_t k) { return 0xffffffffu - (uint32_t)(k & 0xffffffffu); }
__device__ __forceinline__ float key_score(uint64_t k) { return float_from_order_key((uint32_t)(k >> 32)); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}

// Exclusive block scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// `smem` needs 33 ints.  Returns the exclusive prefix; *total receives the block sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int* total) {
    const int lane = lane_id(), wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_inclusive_scan(v);
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? smem[lane] : 0;
        int winc = warp_inclusive_scan(w);
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    int res = inc - v + smem[wid];
    *total = smem[32];
    __syncthreads();
    return res;
}

}  // namespace kb
#endif
