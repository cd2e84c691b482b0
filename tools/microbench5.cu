// Calibration: cost of W warps signalling "done with this stage" once per stage.
//   mode 0: lane 0 of every warp: mbarrier.arrive on ONE barrier (count W), one waiter warp polls try_wait
//   mode 1: same, but 4 barriers (count W/4-ish groups) -- spreads the arrivals
//   mode 2: lane 0 of every warp: atomicAdd on a shared counter, waiter polls the counter
//   mode 3: bar.sync among the W warps + the waiter (hardware barrier)
//   mode 4: nothing (loop overhead)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__global__ void __launch_bounds__(832, 1) k(int mode, int iters, long long *cyc) {
    __shared__ unsigned long long bars[8];
    __shared__ unsigned long long back[1];
    __shared__ int counter;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = (blockDim.x >> 5) - 1;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) {
            const int cnt = mode == 1 ? (W + 3 - i) / 4 : W;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(cnt));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&back[0])), "r"(1));
        counter = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    uint32_t par = 0;
    for (int it = 0; it < iters; ++it) {
        if (warp < W) {            // workers: signal, then wait for the waiter's go (so that rounds do not overlap)
            if (mode == 0) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[0])) : "memory"); }
            else if (mode == 1) { if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[warp & 3])) : "memory"); }
            else if (mode == 2) { if (lane == 0) atomicAdd(&counter, 1); }
            else if (mode == 3) { asm volatile("bar.sync 1, %0;" ::"r"((W + 1) * 32) : "memory"); }
            if (mode != 3) mbar_wait(smem_u32(&back[0]), par);
        } else {                   // waiter
            if (mode == 0) mbar_wait(smem_u32(&bars[0]), par);
            else if (mode == 1) { for (int i = 0; i < 4; ++i) mbar_wait(smem_u32(&bars[i]), par); }
            else if (mode == 2) { while (*(volatile int *)&counter < (it + 1) * W) {} }
            else if (mode == 3) { asm volatile("bar.sync 1, %0;" ::"r"((W + 1) * 32) : "memory"); }
            if (mode != 3 && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&back[0])) : "memory");
        }
        par ^= 1u;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    long long *cyc; CK(cudaMalloc(&cyc, 8));
    for (int W : {25, 12, 4}) for (int mode = 0; mode < 5; ++mode) {
        const int iters = 2000;
        k<<<148, (W + 1) * 32>>>(mode, iters, cyc);
        CK(cudaDeviceSynchronize());
        long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("W=%2d mode %d: %.1f cycles per round\n", W, mode, (double)h / iters); fflush(stdout);
    }
    return 0;
}
