// Shared helpers for the sm_100a kernels behind include/nerfdet_lift.h.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nerfdet_lift.h"

namespace nd {

void set_error(const char *fmt, ...);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

#define ND_REQUIRE(cond, status, ...)                 \
    do {                                              \
        if (!(cond)) {                                \
            nd::set_error(__VA_ARGS__);               \
            return (status);                          \
        }                                             \
    } while (0)

#define ND_CUDA_LAUNCH_CHECK(name)                                                   \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            nd::set_error("%s: CUDA error %s", (name), cudaGetErrorString(e__));     \
            return ND_ERR_CUDA;                                                      \
        }                                                                            \
    } while (0)

// ---- the bit-exact projection contract ---------------------
// q_r = P[r][3] + fma(P[r][2], Z, fma(P[r][1], Y, rn(P[r][0] * X)))   -- the K=4 FMA
// chain torch.bmm evaluates (reference nerfdet.py:398); intrinsics keep nvcc from
// re-associating or contracting differently.
__device__ __forceinline__ float chain4(const float *__restrict__ p, float x, float y, float z) {
    float t = __fmul_rn(p[0], x);
    t = __fmaf_rn(p[1], y, t);
    t = __fmaf_rn(p[2], z, t);
    return __fmaf_rn(p[3], 1.0f, t);
}

// Nearest pixel of a voxel in one view (nerfdet.py:400-403).  Returns false when the
// voxel-view is invalid; xi/yi are only meaningful when true.
__device__ __forceinline__ bool project_nearest(const float *__restrict__ p, float X, float Y, float Z,
                                                int height, int width, float &xr, float &yr, float &q2) {
    const float q0 = chain4(p, X, Y, Z);
    const float q1 = chain4(p + 4, X, Y, Z);
    q2 = chain4(p + 8, X, Y, Z);
    xr = rintf(__fdiv_rn(q0, q2));   // round-half-to-even, IEEE divide
    yr = rintf(__fdiv_rn(q1, q2));
    return (xr >= 0.0f) && (yr >= 0.0f) && (xr < (float)width) && (yr < (float)height) && (q2 > 0.0f);
}

// ---- cache-hinted accesses ----------------------------------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

}  // namespace nd
