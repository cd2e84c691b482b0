// Microbenchmarks that calibrate the lift design on the B200 at hand:
// HBM stream read, L2-resident read, L2-resident write+read, 1 KB-row gather from L2.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_read(const float4 *__restrict__ p, size_t n, float *sink) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(p + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}
__global__ void k_write(float4 *p, size_t n, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_float4(v, v, v, v);
}
// warp reads random 1 KB rows (row = 256 floats), like the lift gather
__global__ void k_gather(const float4 *__restrict__ p, uint32_t n_rows, int rows_per_warp, int row_f4, float *sink) {
    const int lane = threadIdx.x & 31;
    uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t s = w * 2654435761u + 12345u;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < rows_per_warp; i += 4) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = (s >> 8) % n_rows;
            for (int k = 0; k < row_f4 / 32; ++k) v[u * 2 + k] = __ldg(p + (size_t)r * row_f4 + k * 32 + lane);
        }
#pragma unroll
        for (int u = 0; u < 4 * (row_f4 / 32 > 1 ? 2 : 1); ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

template <typename F> float timeit(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    float *sink; CK(cudaMalloc(&sink, 4));
    const size_t big = (size_t)1 << 30;
    float4 *buf; CK(cudaMalloc(&buf, big));
    CK(cudaMemset(buf, 0, big));
    const int grid = 148 * 8, block = 256;
    for (size_t mb : {8, 16, 32, 48, 64, 96, 128, 256, 1024}) {
        size_t bytes = mb << 20, n = bytes / 16;
        float ms = timeit([&] { k_read<<<grid, block>>>(buf, n, sink); }, mb >= 256 ? 10 : 50);
        printf("read  %5zu MB : %8.1f GB/s  (%.1f us)\n", mb, bytes / ms / 1e6, ms * 1e3);
    }
    for (size_t mb : {16, 32, 64, 256}) {
        size_t bytes = mb << 20, n = bytes / 16;
        float ms = timeit([&] { k_write<<<grid, block>>>(buf, n, 1.0f); }, mb >= 256 ? 10 : 50);
        printf("write %5zu MB : %8.1f GB/s  (%.1f us)\n", mb, bytes / ms / 1e6, ms * 1e3);
    }
    for (size_t mb : {32, 60}) {
        size_t bytes = mb << 20, n = bytes / 16;
        float ms = timeit([&] { k_write<<<grid, block>>>(buf, n, 1.0f); k_read<<<grid, block>>>(buf, n, sink); }, 50);
        printf("write+read %3zu MB (2 launches): %8.1f GB/s aggregate (%.1f us)\n", mb, 2 * bytes / ms / 1e6, ms * 1e3);
    }
    // gather: rows of 1 KB / 256 B from a 60 MB region, total traffic ~ 400 MB
    for (int row_f4 : {64, 32}) {
        for (size_t mb : {5, 60, 240}) {
            const uint32_t n_rows = (uint32_t)((mb << 20) / (row_f4 * 16));
            const int warps = grid * block / 32, rpw = (int)((size_t)400e6 / (row_f4 * 16) / warps / 4 * 4);
            float ms = timeit([&] { k_gather<<<grid, block>>>(buf, n_rows, rpw, row_f4, sink); }, 20);
            printf("gather rows of %4d B from %3zu MB: %8.1f GB/s (%.1f us)\n", row_f4 * 16, mb,
                   (double)warps * rpw * row_f4 * 16 / ms / 1e6, ms * 1e3);
        }
    }
    // empty kernel launch + back-to-back launch gap
    float ms = timeit([&] { k_write<<<1, 32>>>(buf, 32, 0.f); }, 1000);
    printf("tiny kernel back-to-back: %.2f us per launch\n", ms * 1e3);
    return 0;
}
