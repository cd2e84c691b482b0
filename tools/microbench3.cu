// Third calibration pass: is an L2-resident scratch really absorbed by L2 when it is written?
// copy (cold HBM source -> fixed scratch), write-only, and read-only, to be run under
// ncu --cache-control none for dram__bytes_{read,write}.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MODE>  // 0 default store, 1 __stcg, 2 __stwt
__global__ void k_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        if (MODE == 0) { dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d; }
        else { __stcg(dst + i, a); __stcg(dst + i + stride, b); __stcg(dst + i + 2 * stride, c); __stcg(dst + i + 3 * stride, d); }
    }
    for (; i < n; i += stride) dst[i] = __ldcs(src + i);
}
__global__ void k_write(float4 *dst, size_t n, float v) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = make_float4(v, v, v, v);
}
__global__ void k_read(const float4 *__restrict__ p, size_t n, float *sink) {
    float acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { float4 a = __ldcs(p + i); acc += a.x + a.y; }
    if (acc == 123.456f) *sink = acc;
}

int main(int argc, char **argv) {
    float *sink; CK(cudaMalloc(&sink, 4));
    const size_t big = (size_t)2 << 30;
    float4 *src; CK(cudaMalloc(&src, big));
    float4 *dst; CK(cudaMalloc(&dst, (size_t)256 << 20));
    CK(cudaMemset(src, 0, big));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int reps = 16;
    for (int mode : {0, 1}) {
        for (size_t mb : {30, 60, 120}) {
            for (int cps : {4, 8}) {
                const size_t n = (mb << 20) / 16;
                CK(cudaDeviceSynchronize());
                cudaEventRecord(a);
                for (int r = 0; r < reps; ++r) {
                    const float4 *s = src + ((size_t)r * (mb << 20) % (big - (mb << 20))) / 16;
                    if (mode == 0) k_copy<0><<<148 * cps, 256>>>(s, dst, n); else k_copy<1><<<148 * cps, 256>>>(s, dst, n);
                }
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
                printf("copy mode %d %3zu MB cold src -> fixed scratch, %d CTA/SM: %7.1f GB/s read (%.1f us)\n", mode, mb, cps,
                       (double)(mb << 20) / ms / 1e6, ms * 1e3);
            }
        }
    }
    for (size_t mb : {30, 60, 120}) {
        const size_t n = (mb << 20) / 16;
        CK(cudaDeviceSynchronize());
        cudaEventRecord(a);
        for (int r = 0; r < reps; ++r) k_write<<<148 * 8, 256>>>(dst, n, (float)r);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
        printf("write-only %3zu MB fixed scratch: %7.1f GB/s (%.1f us)\n", mb, (double)(mb << 20) / ms / 1e6, ms * 1e3);
    }
    for (size_t mb : {60, 240}) {
        const size_t n = (mb << 20) / 16;
        CK(cudaDeviceSynchronize());
        cudaEventRecord(a);
        for (int r = 0; r < reps; ++r) {
            const float4 *s = src + ((size_t)r * (mb << 20) % (big - (mb << 20))) / 16;
            k_read<<<148 * 8, 256>>>(s, n, sink);
        }
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
        printf("read-only cold %3zu MB: %7.1f GB/s (%.1f us)\n", mb, (double)(mb << 20) / ms / 1e6, ms * 1e3);
    }
    return 0;
}
