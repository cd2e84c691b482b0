// Row N1 of SURVEY.md section 8f: the backward of the fused lift (nerfdet.py:164-181 under autograd).
//
// Forward, per channel c and voxel n with count = number of views that see n, denom = count + 1e-8:
//   mean = sum_v x_v / denom                      x_v = features[v][c][pixel(v, n)] for the valid views, 0 otherwise
//   var  = sum over ALL views (x_v - mean)^2 / denom,  cov = exp(-var)        (unobserved voxels: constants, no gradient)
// With g_mean, g_cov the incoming gradients and g_var = -cov * g_cov, the gradient that reaches a valid x_v is
//   g_x = a + b * x_v,   b = 2 g_var / denom,
//                        a = (g_mean - 2 g_var * mean * (denom - nv) / denom) / denom - b * mean
// (the second term of a: d var / d mean = -2 (S1 - nv * mean) / denom with S1 = mean * denom).  Every voxel that lands
// on pixel p of view v reads the SAME x = features[v][c][p], so the scatter back into the feature map is
//   g_features[v][c][p] = A[p] + features[v][c][p] * B[p],    A[p] = sum_{n -> p} a_n,   B[p] = sum_{n -> p} b_n,
// i.e. two scatter-adds of per-voxel coefficients -- the exact transpose of the forward gather -- and no gather at all.
//
//   k_bwd_index    pixel offset (uint16, 0xffff = invalid) of every voxel-view: project_nearest() + the depth gate, the
//                  same device function the forward kernels use, so the masks are bit-identical to the forward's
//   k_bwd_coef     a, b [C][N] from mean, cov, count and the incoming gradients
//   k_bwd_scatter  one CTA per (view, group of 4 channels): the A and B planes of the group live in shared memory
//                  (8 x 18.9 KB at 59 x 80), the CTA walks the voxels of the view with red.shared.add.f32 and writes the
//                  finished planes with coalesced stores.  A CTA owns its planes: no global atomics, no memset of the
//                  241.7 MB gradient.
#include "nd_common.cuh"

namespace nd {

constexpr int kBwdGroup = 4;            // channels per scatter CTA
constexpr int kBwdThreads = 512;

__global__ void k_bwd_index(const float *__restrict__ points, const float *__restrict__ proj, int64_t n_vox, int height,
                            int width, const float *__restrict__ depth, float voxel_z, uint16_t *__restrict__ idx) {
    const int v = blockIdx.y;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float p[12];
    if (threadIdx.x < 12) p[threadIdx.x] = proj[v * 12 + threadIdx.x];
    __syncthreads();
    if (n >= n_vox) return;
    float xr, yr, q2;
    bool ok = project_nearest(p, points[n], points[n_vox + n], points[2 * n_vox + n], height, width, xr, yr, q2);
    int off = 0;
    if (ok) {
        const int xi = (int)xr, yi = (int)yr;
        off = yi * width + xi;
        if (depth != nullptr) {                            // B4, nerfdet.py:405-411
            const float d = depth[((int64_t)v * height + yi) * width + xi];
            ok = (q2 > __fsub_rn(d, voxel_z)) && (q2 < __fadd_rn(d, voxel_z));
        }
    }
    idx[(int64_t)v * n_vox + n] = ok ? (uint16_t)off : (uint16_t)0xffffu;
}

__global__ void k_bwd_coef(const float *__restrict__ mean, const float *__restrict__ cov, const int64_t *__restrict__ count,
                           const float *__restrict__ g_mean, const float *__restrict__ g_cov, int channels, int64_t n_vox,
                           int n_views_total, float *__restrict__ a, float *__restrict__ b) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (n >= n_vox) return;
    const int64_t i = (int64_t)c * n_vox + n;
    const int64_t cnt = count[n];
    float ca = 0.0f, cb = 0.0f;
    if (cnt > 0) {
        const float denom = __fadd_rn((float)cnt, 1e-8f);
        const float m = mean[i];
        const float gm = g_mean != nullptr ? g_mean[i] : 0.0f;
        const float gv = (g_cov != nullptr && cov != nullptr) ? -cov[i] * g_cov[i] : 0.0f;
        cb = 2.0f * gv / denom;
        const float gm_total = gm - 2.0f * gv * m * (denom - (float)n_views_total) / denom;
        ca = gm_total / denom - cb * m;
    }
    a[i] = ca;
    b[i] = cb;
}

template <typename T> __device__ __forceinline__ T from_f32_t(float v);
template <> __device__ __forceinline__ float from_f32_t<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32_t<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ void red_shared_add(float *p, float v) {
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v) : "memory");
}

// grid (ceil(C / kBwdGroup), nv).  Shared memory: A[kBwdGroup][plane], B[kBwdGroup][plane] fp32.
template <typename T>
__global__ void __launch_bounds__(kBwdThreads)
k_bwd_scatter(const T *__restrict__ feat, int64_t sv, int64_t sc, int64_t sy, int64_t sx, int channels, int height, int width,
              const uint16_t *__restrict__ idx, const float *__restrict__ a, const float *__restrict__ b, int64_t n_vox,
              T *__restrict__ g_feat) {
    extern __shared__ __align__(16) float s_planes[];
    const int v = blockIdx.y, c0 = blockIdx.x * kBwdGroup;
    const int plane = height * width;
    const int nc = min(kBwdGroup, channels - c0);
    float *sA = s_planes, *sB = s_planes + (size_t)kBwdGroup * plane;
    for (int i = threadIdx.x; i < 2 * kBwdGroup * plane; i += blockDim.x) s_planes[i] = 0.0f;
    __syncthreads();
    const uint16_t *row = idx + (int64_t)v * n_vox;
    const float *a0 = a + (int64_t)c0 * n_vox, *b0 = b + (int64_t)c0 * n_vox;
    // two consecutive voxels per thread and pass (one 32-bit load of the offsets); a warp covers 64 voxels of a z-run pair,
    // which a view sees or misses as a whole most of the time, so the coefficient loads of unseen runs are skipped
    const bool pairs = (n_vox & 1) == 0 && (reinterpret_cast<uintptr_t>(row) & 3) == 0;
    if (pairs) {
        for (int64_t n = 2 * (int64_t)threadIdx.x; n < n_vox; n += 2 * (int64_t)blockDim.x) {
            const uint32_t two = __ldg(reinterpret_cast<const uint32_t *>(row + n));
            const uint32_t o0 = two & 0xffffu, o1 = two >> 16;
            if ((o0 & o1) == 0xffffu) continue;
#pragma unroll
            for (int k = 0; k < kBwdGroup; ++k) {
                if (k < nc) {
                    const float2 ca = __ldg(reinterpret_cast<const float2 *>(a0 + (int64_t)k * n_vox + n));
                    const float2 cb = __ldg(reinterpret_cast<const float2 *>(b0 + (int64_t)k * n_vox + n));
                    if (o0 != 0xffffu) { red_shared_add(sA + k * plane + o0, ca.x); red_shared_add(sB + k * plane + o0, cb.x); }
                    if (o1 != 0xffffu) { red_shared_add(sA + k * plane + o1, ca.y); red_shared_add(sB + k * plane + o1, cb.y); }
                }
            }
        }
    } else {
        for (int64_t n = threadIdx.x; n < n_vox; n += blockDim.x) {
            const uint32_t o = row[n];
            if (o == 0xffffu) continue;
            for (int k = 0; k < nc; ++k) {
                red_shared_add(sA + k * plane + o, a0[(int64_t)k * n_vox + n]);
                red_shared_add(sB + k * plane + o, b0[(int64_t)k * n_vox + n]);
            }
        }
    }
    __syncthreads();
    // g = A + x * B, written in the layout of a contiguous [nv][C][height][width] gradient
    for (int i = threadIdx.x; i < nc * plane; i += blockDim.x) {
        const int k = i / plane, p = i - k * plane, y = p / width, x = p - y * width;
        const float f = to_f32<T>(feat[v * sv + (int64_t)(c0 + k) * sc + (int64_t)y * sy + (int64_t)x * sx]);
        g_feat[((int64_t)v * channels + c0 + k) * plane + p] = from_f32_t<T>(fmaf(f, sB[k * plane + p], sA[k * plane + p]));
    }
}

}  // namespace nd

using namespace nd;

extern "C" {

size_t nd_lift_backward_workspace_bytes(const nd_maps *features, int64_t n_voxels) {
    if (features == nullptr || n_voxels <= 0) return 0;
    const size_t idx = ((size_t)features->n_views * n_voxels * sizeof(uint16_t) + 255) & ~(size_t)255;
    const size_t coef = ((size_t)features->channels * n_voxels * sizeof(float) + 255) & ~(size_t)255;
    return idx + 2 * coef;
}

int nd_lift_backward(const nd_maps *features, const float *points, const float *projection, int64_t n_voxels,
                     const float *depth_resized, float voxel_z, int n_views_total, const float *mean, const float *cov,
                     const int64_t *count, const float *grad_mean, const float *grad_cov, void *grad_features,
                     void *workspace, size_t workspace_bytes, void *stream) {
    ND_REQUIRE(features && features->data && points && projection && mean && count && grad_features && workspace,
               ND_ERR_BAD_ARG, "nd_lift_backward: null pointer");
    ND_REQUIRE(grad_mean != nullptr || grad_cov != nullptr, ND_ERR_BAD_ARG, "nd_lift_backward: no incoming gradient");
    ND_REQUIRE(grad_cov == nullptr || cov != nullptr, ND_ERR_BAD_ARG, "nd_lift_backward: grad_cov needs the forward's cov");
    const int nv = features->n_views, ch = features->channels, h = features->height, w = features->width;
    ND_REQUIRE(nv > 0 && ch > 0 && h > 0 && w > 0 && n_voxels > 0 && nv <= 65535, ND_ERR_BAD_SHAPE, "nd_lift_backward: bad shape");
    ND_REQUIRE(features->dtype == ND_F32 || features->dtype == ND_BF16, ND_ERR_BAD_ARG, "nd_lift_backward: dtype");
    const int64_t plane = (int64_t)h * w;
    ND_REQUIRE(plane < 0xffff, ND_ERR_BAD_SHAPE, "nd_lift_backward: planes of %lld pixels (limit 65534)", (long long)plane);
    const size_t smem = (size_t)2 * kBwdGroup * plane * sizeof(float);
    ND_REQUIRE(smem <= 220 * 1024, ND_ERR_BAD_SHAPE, "nd_lift_backward: planes of %lld pixels do not fit shared memory",
               (long long)plane);
    ND_REQUIRE(workspace_bytes >= nd_lift_backward_workspace_bytes(features, n_voxels) &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               ND_ERR_WORKSPACE, "nd_lift_backward: workspace too small or not 256-byte aligned");
    if (n_views_total <= 0) n_views_total = nv;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    uint16_t *idx = reinterpret_cast<uint16_t *>(ws);
    const size_t idx_bytes = ((size_t)nv * n_voxels * sizeof(uint16_t) + 255) & ~(size_t)255;
    const size_t coef_bytes = ((size_t)ch * n_voxels * sizeof(float) + 255) & ~(size_t)255;
    float *a = reinterpret_cast<float *>(ws + idx_bytes), *b = reinterpret_cast<float *>(ws + idx_bytes + coef_bytes);
    k_bwd_index<<<dim3((unsigned)ceil_div(n_voxels, 256), nv), 256, 0, st>>>(points, projection, n_voxels, h, w, depth_resized,
                                                                             voxel_z, idx);
    ND_CUDA_LAUNCH_CHECK("k_bwd_index");
    k_bwd_coef<<<dim3((unsigned)ceil_div(n_voxels, 256), ch), 256, 0, st>>>(mean, cov, count, grad_mean, grad_cov, ch, n_voxels,
                                                                            n_views_total, a, b);
    ND_CUDA_LAUNCH_CHECK("k_bwd_coef");
    const dim3 grid((unsigned)ceil_div(ch, kBwdGroup), nv);
    cudaError_t e;
    if (features->dtype == ND_F32) {
        auto kern = k_bwd_scatter<float>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            kern<<<grid, kBwdThreads, smem, st>>>((const float *)features->data, features->stride_v, features->stride_c,
                                                  features->stride_y, features->stride_x, ch, h, w, idx, a, b, n_voxels,
                                                  (float *)grad_features);
    } else {
        auto kern = k_bwd_scatter<__nv_bfloat16>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            kern<<<grid, kBwdThreads, smem, st>>>((const __nv_bfloat16 *)features->data, features->stride_v, features->stride_c,
                                                  features->stride_y, features->stride_x, ch, h, w, idx, a, b, n_voxels,
                                                  (__nv_bfloat16 *)grad_features);
    }
    if (e != cudaSuccess) {
        set_error("nd_lift_backward: %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    ND_CUDA_LAUNCH_CHECK("k_bwd_scatter");
    return ND_OK;
}

}  // extern "C"
