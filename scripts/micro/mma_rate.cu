// Microbenchmark (not product code): issue rate of tcgen05.mma kind::f16 with both operands in shared
// memory, M=128, N in {64,128,256}, K=16, cta_group::1, back-to-back on one accumulator.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
    return (uint64_t)((a & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int b_tiles, long long* out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    // zero the operand area (A: 16 KB at base, B tiles after it)
    for (int i = threadIdx.x; i < (16384 + b_tiles * 32768) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem_dyn + (base - smem_u32(smem_dyn)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x < 32) {
        const uint64_t a = umma_desc(base);
        long long t0 = clock64();
        if (threadIdx.x == 0) {
            for (int i = 0; i < iters; ++i) {
                const uint64_t b = umma_desc(base + 16384 + (i % b_tiles) * 32768);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma(tmem, a + 2 * k, b + 2 * k, IDESC, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

template <int N>
void run(int grid, int b_tiles) {
    long long* d;
    cudaMalloc(&d, 8);
    const int iters = 4096;
    size_t smem = 16384 + (size_t)b_tiles * 32768 + 2048;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rate_kernel<N><<<grid, 128, smem>>>(iters, b_tiles, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%d grid=%d b_tiles=%d: %s  %.1f cycles per MMA (M128 x N%d x K16)\n", N, grid, b_tiles, cudaGetErrorString(e),
           (double)h / (iters * 4.0), N);
    cudaFree(d);
}

int main() {
    run<64>(148, 2);
    run<128>(148, 2);
    run<256>(148, 2);
    run<128>(1, 2);
    run<256>(1, 2);
    run<128>(148, 5);
    return 0;
}
