// Calibration: issue->retire rate of back-to-back tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M = 128, K = 16,
// cta_group::1) from one thread, as a function of N, of the A operand source (shared memory descriptor vs tensor
// memory) and of whether consecutive instructions accumulate into the same or into alternating TMEM regions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench7 tools/microbench7.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = (uint64_t)((addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

__global__ void __launch_bounds__(128, 1) k_mma_rate(int n, int a_from_tmem, int alternate, int iters, long long *out, int commit_every, int fence_every) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long bar;
    __shared__ __align__(8) unsigned long long bar2;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint64_t ad = smem_desc(base), bd = smem_desc(base + 16384);
        const uint32_t id = idesc(n);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + ((alternate && (i & 1)) ? 256u : 0u);
            const uint64_t ko = (uint64_t)(2 * (i & 3));
            if (a_from_tmem)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                             "r"(tmem + 256u + 16u * (uint32_t)(i & 3)), "l"(bd + ko), "r"(id), "r"(1)
                             : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                             "l"(ad + ko), "l"(bd + ko), "r"(id), "r"(1)
                             : "memory");
            if (commit_every > 0 && (i & (commit_every - 1)) == commit_every - 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            if (fence_every > 0 && (i & (fence_every - 1)) == fence_every - 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok)
                         : "r"(smem_u32(&bar)), "r"(0)
                         : "memory");
        } while (!ok);
        const long long t1 = clock64();
        if (blockIdx.x == 0) *out = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long *out, h;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    for (int n : {128, 256})
        for (int a_tm : {0, 1})
            for (int ce : {0, 4, 1})
                for (int fe : {0, 4}) {
                    k_mma_rate<<<148, 128, 64 * 1024>>>(n, a_tm, 0, iters, out, ce, fe);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
                    printf("N %3d  A from %s  commit every %d  fence every %d: %.1f cycles per MMA (%s)\n", n, a_tm ? "tmem" : "smem", ce, fe,
                           (double)h / iters, cudaGetErrorString(e));
                }
    return 0;
}
