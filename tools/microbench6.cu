// Calibration: how fast can 148 CTAs stream HBM-resident data into shared memory with cp.async.bulk (TMA) rings,
// as a function of the copy size and of the number of slots in flight -- the plane ring of k_lift_planes holds
// 8 slots of 18.9 KB per SM.  Reference: the same bytes read with LDG.128 by 1024 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench6 tools/microbench6.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_tma_stream(const char *src, size_t total, uint32_t chunk, int slots, int *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem);
    const uint32_t bar0 = smem_u32(bars), buf0 = smem_u32(smem + 256);
    const uint32_t pitch = (chunk + 127u) & ~127u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < slots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t n_chunks = total / chunk;
    size_t k = blockIdx.x;
    int issued = 0, s = 0;
    uint32_t ph = 0;
    // fill the ring
    for (; issued < slots && k < n_chunks; ++issued, k += gridDim.x) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * issued), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf0 + issued * pitch),
                     "l"(src + k * (size_t)chunk), "r"(chunk), "r"(bar0 + 8 * issued)
                     : "memory");
    }
    int done = 0;
    while (done < issued) {
        uint32_t ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok)
                         : "r"(bar0 + 8 * s), "r"(ph)
                         : "memory");
        } while (!ok);
        ++done;
        if (k < n_chunks) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf0 + s * pitch),
                         "l"(src + k * (size_t)chunk), "r"(chunk), "r"(bar0 + 8 * s)
                         : "memory");
            k += gridDim.x;
            ++issued;
        }
        if (++s == slots) { s = 0; ph ^= 1u; }
    }
    if (sink != nullptr && done == -1) *sink = done;
}

__global__ void __launch_bounds__(1024, 1) k_ldg_stream(const uint4 *src, size_t n_vec, uint32_t *sink) {
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(src + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const size_t total = (size_t)1 << 30;      // 1 GiB source, 8x the L2
    char *src;
    int *sink;
    cudaMalloc(&src, total);
    cudaMalloc(&sink, 4);
    cudaMemset(src, 1, total);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    {
        for (int it = 0; it < 2; ++it) {
            cudaEventRecord(e0);
            k_ldg_stream<<<sms, 1024>>>(reinterpret_cast<const uint4 *>(src), total / 16, reinterpret_cast<uint32_t *>(sink));
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it) printf("ldg.128 stream, %d CTAs x 1024 threads: %.0f GB/s\n", sms, total / ms / 1e6);
        }
    }
    const uint32_t chunks[] = {4720, 18880, 37760, 75520};
    const int slot_list[] = {2, 4, 8, 11};
    cudaFuncSetAttribute(k_tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (uint32_t chunk : chunks)
        for (int slots : slot_list) {
            const size_t pitch = (chunk + 127u) & ~127u;
            const size_t smem = 256 + pitch * slots;
            if (smem > 227 * 1024) continue;
            const size_t use = total / chunk * chunk;
            float ms = 0;
            for (int it = 0; it < 2; ++it) {
                cudaEventRecord(e0);
                k_tma_stream<<<sms, 128, smem>>>(src, use, chunk, slots, sink);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            cudaError_t err = cudaGetLastError();
            printf("tma ring: chunk %6u B x %2d slots (%6.1f KB in flight per SM): %.0f GB/s %s\n", chunk, slots,
                   chunk * slots / 1024.0, use / ms / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
    return 0;
}
