// Microbenchmark (not product code): TMEM -> register bandwidth of tcgen05.ld.32x32b.x32 with W warps per CTA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(1024, 1) ld_kernel(int iters, int pipelined, long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) * 64) % 512;
    uint32_t acc = 0;
    uint32_t va[32], vb[32];
    __syncthreads();
    long long t0 = clock64();
    if (pipelined) {
        for (int i = 0; i < iters; ++i) {
            tmem_ld32(taddr, va);
            tmem_ld32(taddr + 32, vb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= va[i & 31] ^ vb[(i + 7) & 31];
        }
    } else {
        for (int i = 0; i < iters; ++i) {
            tmem_ld32(taddr, va);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= va[i & 31];
            tmem_ld32(taddr + 32, vb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= vb[i & 31];
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* d; uint32_t* sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 148 * 1024 * 4);
    for (int warps : {4, 8, 16, 32}) for (int pip : {0, 1}) {
        const int iters = 2000;
        ld_kernel<<<148, warps * 32>>>(iters, pip, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        double bytes = (double)warps * iters * 2 * 4096;
        printf("warps=%d pipelined=%d: %s  %.1f B/clk/SM  (%.0f cycles per x32 load per warp)\n", warps, pip, cudaGetErrorString(e),
               bytes / h, (double)h / (iters * 2));
    }
    return 0;
}
