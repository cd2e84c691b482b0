// Calibration: how fast can one SM pull small rows (offset tables) out of L2 into shared memory / registers?
//   mode 0: cp.async.cg 16 B per lane (LDGSTS), one 512 B row per warp instruction, 25 warps per CTA
//   mode 1: ld.global.nc.v4 (LDG.128) into registers, same rows
//   mode 2: cp.async.bulk (TMA) of `row_bytes` per copy, issued by lane 0 of each of the 25 warps, own mbarrier
//   mode 3: cp.async.bulk issued by ONE thread of the CTA
// All 148 CTAs (1 per SM) read from the same L2-resident table of `table_mb`.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(800, 1)
k_rows(const char *__restrict__ table, size_t table_bytes, int mode, int row_bytes, int iters, float *sink, long long *cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bars[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = blockDim.x >> 5;
    if (threadIdx.x < 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[threadIdx.x])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const size_t n_rows = table_bytes / row_bytes;
    const uint32_t dst = smem_u32(smem) + warp * 4 * 2048;   // 4 slots of up to 2 KB per warp
    float acc = 0.f;
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
            const size_t r = ((size_t)blockIdx.x * 7919 + (size_t)warp * 131 + (size_t)i * 977) % n_rows;
            const char *src = table + r * row_bytes + lane * 16;
            for (int b = 0; b < row_bytes; b += 512)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (i & 3) * 2048 + b + lane * 16), "l"(src + b) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 3;" ::: "memory");
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (mode == 1) {
        float4 v[4];
        for (int k = 0; k < 4; ++k) v[k] = make_float4(0, 0, 0, 0);
        for (int i = 0; i < iters; ++i) {
            const size_t r = ((size_t)blockIdx.x * 7919 + (size_t)warp * 131 + (size_t)i * 977) % n_rows;
            const char *src = table + r * row_bytes + lane * 16;
            acc += v[i & 3].x;
            for (int b = 0; b < row_bytes; b += 512)
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[i & 3].x), "=f"(v[i & 3].y), "=f"(v[i & 3].z), "=f"(v[i & 3].w) : "l"(src + b));
        }
        for (int k = 0; k < 4; ++k) acc += v[k].x;
    } else if (mode == 2 || mode == 3) {
        const bool issuer = mode == 2 ? lane == 0 : threadIdx.x == 0;
        const int n_it = mode == 2 ? iters : iters * W;
        if (issuer) {
            const uint32_t bar = smem_u32(&bars[warp]);
            uint32_t parity = 0;
            for (int i = 0; i < n_it; i += 4) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4 * row_bytes) : "memory");
                for (int k = 0; k < 4; ++k) {
                    const size_t r = ((size_t)blockIdx.x * 7919 + (size_t)warp * 131 + (size_t)(i + k) * 977) % n_rows;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + k * 2048),
                                 "l"(table + r * row_bytes), "r"(row_bytes), "r"(bar) : "memory");
                }
                uint32_t done = 0;
                while (!done) {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                }
                parity ^= 1u;
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const size_t table_bytes = (size_t)3 << 20;
    char *table; CK(cudaMalloc(&table, table_bytes)); CK(cudaMemset(table, 1, table_bytes));
    float *sink; CK(cudaMalloc(&sink, 4));
    long long *cyc; CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int W = 25, iters = 512;
    for (int mode = 0; mode < 4; ++mode) {
        for (int row : {256, 512, 1024, 2048}) {
            if (mode == 1 && row > 512) continue;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(a));
                k_rows<<<148, W * 32, 200 * 1024>>>(table, table_bytes, mode, row, iters, sink, cyc);
                CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
                CK(cudaGetLastError());
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            const double rows_per_sm = (double)W * iters;
            printf("mode %d row %4d B: %.1f us, %.0f cycles per SM, %.1f cycles per row per SM, %.1f B/clk/SM, %.2f TB/s chip\n", mode, row,
                   ms * 1e3, avg, avg / rows_per_sm, rows_per_sm * row / avg, 148.0 * rows_per_sm * row / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
