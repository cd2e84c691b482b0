// tools/microbench8: throughput of shared-memory atomics on sm_100a: fp32 add (ATOMS.CAST.SPIN compare-and-swap loop) vs native int32 add, random addresses
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(const uint16_t *idx, int n, int reps, float *out) {
    extern __shared__ float s[];
    const int words = 37760;
    for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = 0.f;
    __syncthreads();
    for (int r = 0; r < reps; ++r)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t o = idx[i];
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(s + o);
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + k * 18880), "f"(1.0f + k) : "memory");
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a + k * 18880), "r"(1u + k) : "memory");
            }
        }
    __syncthreads();
    float t = 0;
    for (int i = threadIdx.x; i < words; i += blockDim.x) t += s[i];
    if (t == 12345.f) out[0] = t;
}
int main() {
    const int n = 5900;   // valid voxels of one view
    uint16_t *h = new uint16_t[n];
    uint32_t x = 1;
    for (int i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; h[i] = (x >> 8) % 4720; }
    uint16_t *d; float *o;
    cudaMalloc(&d, n * 2); cudaMalloc(&o, 4);
    cudaMemcpy(d, h, n * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 151040);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 151040);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int it = 0; it < 2; ++it) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 512, 151040>>>(d, n, 20, o); else k<1><<<148, 512, 151040>>>(d, n, 20, o);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double atomics = 148.0 * 20 * n * 8;
        printf("%s: %.1f us, %.2f G atomics/s on 148 SMs, %.2f cycles per warp-atomic per SM at 1.9 GHz\n", mode == 0 ? "red.shared.add.f32 (CAS loop)" : "red.shared.add.u32 (native)",
               ms * 1e3, atomics / ms / 1e6, ms * 1e-3 * 1.9e9 / (20.0 * n * 8 / 32));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
