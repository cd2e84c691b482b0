// Second calibration pass: sustained L2-hit bandwidth (streaming and row gathers of several
// widths), copy (HBM read + L2-resident write), cooperative grid-barrier cost.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_stream(const float4 *__restrict__ p, size_t n, int iters, float *sink) {
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n; i += 4 * stride) {
            float4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
            acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y;
        }
        for (; i < n; i += stride) { float4 a = __ldg(p + i); acc.x += a.x; }
    }
    if (acc.x + acc.y == 123.456f) *sink = acc.x;
}

template <int VEC>  // floats per lane per row: 1, 2, 4, 8
__global__ void k_rows(const float *__restrict__ p, uint32_t n_rows, int rows_per_warp, float *sink) {
    const int lane = threadIdx.x & 31;
    uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t s = w * 2654435761u + 12345u;
    float acc = 0.f;
    constexpr int U = 8;
    constexpr int ROWF = 32 * VEC;
    for (int i = 0; i < rows_per_warp; i += U) {
        float v[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = (uint32_t)(((uint64_t)(s >> 4) * n_rows) >> 28);
            const float *row = p + (size_t)r * ROWF;
            if constexpr (VEC == 1) v[u][0] = __ldg(row + lane);
            else if constexpr (VEC == 2) { float2 t = __ldg((const float2 *)row + lane); v[u][0] = t.x; v[u][1] = t.y; }
            else if constexpr (VEC == 4) { float4 t = __ldg((const float4 *)row + lane); v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w; }
            else { float4 t = __ldg((const float4 *)row + lane), q = __ldg((const float4 *)row + 32 + lane);
                   v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
                   v[u][4] = q.x; v[u][5] = q.y; v[u][6] = q.z; v[u][7] = q.w; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc = fmaf(v[u][j], v[u][j], acc);
    }
    if (acc == 123.456f) *sink = acc;
}

__global__ void k_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = __ldcs(src + i);
}

__global__ void k_gridsync(int n, int *out) {
    cg::grid_group g = cg::this_grid();
    int acc = 0;
    for (int i = 0; i < n; ++i) { g.sync(); acc += i; }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = acc;
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

template <int VEC> void rows_test(const float *buf, float *sink, size_t region_mb, int ctas_per_sm, int threads) {
    const int grid = 148 * ctas_per_sm;
    const uint32_t n_rows = (uint32_t)((region_mb << 20) / (32 * VEC * 4));
    const int warps = grid * threads / 32;
    int rpw = (int)((size_t)800e6 / (32 * VEC * 4) / warps / 8 * 8);
    if (rpw < 8) rpw = 8;
    float ms = timeit([&] { k_rows<VEC><<<grid, threads>>>(buf, n_rows, rpw, sink); }, 5);
    printf("rows %4d B, region %4zu MB, %d x %4d thr/SM: %8.1f GB/s (%.1f us)\n", 32 * VEC * 4, region_mb, ctas_per_sm,
           threads, (double)warps * rpw * 32 * VEC * 4 / ms / 1e6, ms * 1e3);
}

int main() {
    float *sink; CK(cudaMalloc(&sink, 4));
    const size_t big = (size_t)1 << 30;
    float4 *buf; CK(cudaMalloc(&buf, big));
    float4 *buf2; CK(cudaMalloc(&buf2, big / 2));
    CK(cudaMemset(buf, 0, big));
    for (size_t mb : {16, 32, 64, 96}) {
        const int iters = (int)(2048 / mb);
        for (int cps : {2, 8}) {
            float ms = timeit([&] { k_stream<<<148 * cps, 256>>>(buf, (mb << 20) / 16, iters, sink); }, 3);
            printf("L2 stream %3zu MB x %3d iters, %d CTA/SM: %8.1f GB/s (%.1f us)\n", mb, iters, cps,
                   (double)(mb << 20) * iters / ms / 1e6, ms * 1e3);
        }
    }
    {
        float ms = timeit([&] { k_stream<<<148 * 8, 256>>>(buf, big / 16, 1, sink); }, 5);
        printf("HBM stream 1024 MB: %8.1f GB/s\n", (double)big / ms / 1e6);
    }
    rows_test<8>((float *)buf, sink, 30, 2, 1024);
    rows_test<8>((float *)buf, sink, 60, 2, 1024);
    rows_test<8>((float *)buf, sink, 60, 8, 256);
    rows_test<4>((float *)buf, sink, 60, 2, 1024);
    rows_test<2>((float *)buf, sink, 60, 2, 1024);
    rows_test<1>((float *)buf, sink, 30, 2, 1024);
    rows_test<1>((float *)buf, sink, 60, 2, 1024);
    rows_test<8>((float *)buf, sink, 240, 2, 1024);
    rows_test<8>((float *)buf, sink, 1000, 2, 1024);
    for (size_t mb : {30, 60, 240}) {
        size_t n = (mb << 20) / 16;
        float ms = timeit([&] { k_copy<<<148 * 8, 256>>>(buf + (size_t)(512 << 20) / 16, buf2, n); }, 10);
        printf("copy %3zu MB HBM->scratch: %8.1f GB/s read (+same written) (%.1f us)\n", mb, (double)(mb << 20) / ms / 1e6, ms * 1e3);
    }
    for (int cps : {1, 2}) {
        int *out; CK(cudaMalloc(&out, 4));
        int n = 100;
        void *args[] = {&n, &out};
        dim3 grid(148 * cps), block(512);
        float ms = timeit([&] { CK(cudaLaunchCooperativeKernel((void *)k_gridsync, grid, block, args, 0, 0)); }, 10);
        printf("cooperative grid.sync, %d CTA/SM x 512 thr: %.2f us per sync\n", cps, ms * 1e3 / n);
    }
    return 0;
}
