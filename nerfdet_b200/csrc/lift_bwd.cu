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
// i.e. two sums of per-voxel coefficients over the voxels of a pixel -- the exact transpose of the forward gather -- and no
// feature gather at all.
//
//   k_bwd_index    pixel offset (uint16, 0xffff = invalid) of every voxel-view: project_nearest() + the depth gate, the
//                  same device function the forward kernels use, so the masks are bit-identical to the forward's
// The sums are GATHERS, the mirror image of the forward kernel -- there the planes sit in shared memory and the voxels gather
// from them, here the coefficient row of a channel sits in shared memory and the pixels gather from it.
//   k_bwd_csr      per (view, part of <= 25 600 voxels): the voxels of every pixel as a CSR (counting sort in shared memory,
//                  every pixel's short list sorted by voxel index, so the sums have ONE order: the gradient is deterministic)
//   k_bwd_gather   one CTA per (channel, range of views): a, b of the part's voxels computed from mean / cov / count and the
//                  incoming gradients straight into shared memory (float2 x 25 600 = 205 KB); then, per view, one thread per
//                  pixel reads its 8-byte record {count, first three voxels}, adds the coefficients of its voxels, forms
//                  A + x * B and stores it (coalesced).  No atomics, no memset of the 241.7 MB gradient; parts beyond the
//                  first add to what the previous launch wrote.
// The first version scattered with red.shared.add.f32 into planes in shared memory: a compare-and-swap loop on sm_100a
// (ATOMS.CAST.SPIN; 0.73 T atomics/s on the whole GPU against 2.1 T/s for native int32 adds, tools/microbench8.cu), 1.32 ms
// at the bench shape.
#include "nd_common.cuh"

namespace nd {

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

template <typename T> __device__ __forceinline__ T from_f32_t(float v);
template <> __device__ __forceinline__ float from_f32_t<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32_t<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------------------------------
// gather form
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBwdPartMax = 25600;      // voxels whose (a, b) fit shared memory as float2
constexpr int kBwdCsrPixels = 20480;    // planes the CSR builder takes (start + cursor words and the part's list in shared memory)
constexpr int kBwdCsrThreads = 1024;

// the coefficients of one (channel, voxel); count == 0: no gradient (the forward wrote constants there)
__device__ __forceinline__ float2 bwd_coef(float m, float cv, int64_t cnt, float gm, float gc, int n_views_total) {
    if (cnt <= 0) return make_float2(0.0f, 0.0f);
    const float denom = __fadd_rn((float)cnt, 1e-8f);
    const float gv = -cv * gc;
    const float cb = 2.0f * gv / denom;
    const float gm_total = gm - 2.0f * gv * m * (denom - (float)n_views_total) / denom;
    return make_float2(gm_total / denom - cb * m, cb);
}

// grid (parts, nv).  rec: uint2 [nv][parts][plane]; start: uint16 [nv][parts][plane + 1]; list: uint16 [nv][parts][part_len] (voxel index inside the part)
__global__ void __launch_bounds__(kBwdCsrThreads)
k_bwd_csr(const uint16_t *__restrict__ idx, int64_t n_vox, int part_len, int plane, uint16_t *__restrict__ start,
          uint16_t *__restrict__ list, uint2 *__restrict__ rec) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t *s_start = reinterpret_cast<uint32_t *>(s_raw);               // [plane + 1]
    uint32_t *s_cur = s_start + plane + 1;                                 // [plane]
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_cur + plane);        // [part_len]
    __shared__ uint32_t s_warp[32];
    const int part = blockIdx.x, v = blockIdx.y, parts = gridDim.x;
    const int64_t n0 = (int64_t)part * part_len;
    const int len = (int)min((int64_t)part_len, n_vox - n0);
    const uint16_t *row = idx + (int64_t)v * n_vox + n0;
    for (int i = threadIdx.x; i <= plane; i += blockDim.x) s_start[i] = 0u;
    __syncthreads();
    for (int n = threadIdx.x; n < len; n += blockDim.x) {
        const uint32_t o = row[n];
        if (o != 0xffffu) atomicAdd(&s_start[o], 1u);
    }
    __syncthreads();
    // exclusive scan of the pixel counts: a run of consecutive pixels per thread, then a block scan of the run sums
    const int per = (plane + blockDim.x - 1) / blockDim.x;
    const int p0 = min(threadIdx.x * per, plane), p1 = min(p0 + per, plane);
    uint32_t sum = 0;
    for (int p = p0; p < p1; ++p) sum += s_start[p];
    uint32_t inc = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    uint32_t run = inc - sum + (warp > 0 ? s_warp[warp - 1] : 0u);        // exclusive prefix of this thread's run
    for (int p = p0; p < p1; ++p) {
        const uint32_t c = s_start[p];
        s_start[p] = run;
        s_cur[p] = run;
        run += c;
    }
    if (threadIdx.x == 0) s_start[plane] = s_warp[31];
    __syncthreads();
    for (int n = threadIdx.x; n < len; n += blockDim.x) {
        const uint32_t o = row[n];
        if (o != 0xffffu) s_list[atomicAdd(&s_cur[o], 1u)] = (uint16_t)n;
    }
    __syncthreads();
    // one order per pixel: ascending voxel index (the lists are short: 1.8 voxels per touched pixel at the bench shape)
    for (int p = threadIdx.x; p < plane; p += blockDim.x) {
        const int a = (int)s_start[p], b = (int)s_start[p + 1];
        for (int i = a + 1; i < b; ++i) {
            const uint16_t key = s_list[i];
            int j = i - 1;
            while (j >= a && s_list[j] > key) { s_list[j + 1] = s_list[j]; --j; }
            s_list[j + 1] = key;
        }
    }
    __syncthreads();
    uint16_t *g_start = start + ((int64_t)v * parts + part) * (plane + 1);
    uint16_t *g_list = list + ((int64_t)v * parts + part) * part_len;
    uint2 *g_rec = rec + ((int64_t)v * parts + part) * plane;
    for (int i = threadIdx.x; i <= plane; i += blockDim.x) g_start[i] = (uint16_t)s_start[i];
    const int total = (int)s_start[plane];
    for (int i = threadIdx.x; i < total; i += blockDim.x) g_list[i] = s_list[i];
    // what the gather kernel reads per pixel: one 8-byte record {count, first three voxels}; the 5 % of the pixels with more
    // than three voxels continue in the list at start[p] + 3
    for (int p = threadIdx.x; p < plane; p += blockDim.x) {
        const int a = (int)s_start[p], cnt = (int)s_start[p + 1] - a;
        const uint32_t e0 = cnt > 0 ? s_list[a] : 0u, e1 = cnt > 1 ? s_list[a + 1] : 0u, e2 = cnt > 2 ? s_list[a + 2] : 0u;
        g_rec[p] = make_uint2((uint32_t)cnt | (e0 << 16), e1 | (e2 << 16));
    }
}

// grid (view ranges, C), one launch per part.  Shared memory: float2 [part_len] coefficients.  A thread owns kBwdPpt pixel
// positions (the same in every view: their offsets inside a feature plane are computed once) and per view reads ONE 8-byte
// record per pixel -- count and the first three voxels -- with coalesced loads.
constexpr int kBwdPpt = 5;
constexpr int kBwdGatherThreads = 1024;

template <typename T>
__global__ void __launch_bounds__(kBwdGatherThreads, 1)
k_bwd_gather(const T *__restrict__ feat, int64_t sv, int64_t sc, int64_t sy, int64_t sx, int n_views, int channels, int height,
             int width, const float *__restrict__ mean, const float *__restrict__ cov, const int64_t *__restrict__ count,
             const float *__restrict__ g_mean, const float *__restrict__ g_cov, int64_t n_vox, int n_views_total, int part,
             int parts, int part_len, const uint2 *__restrict__ rec, const uint16_t *__restrict__ start,
             const uint16_t *__restrict__ list, T *__restrict__ g_feat) {
    extern __shared__ __align__(16) float2 s_coef[];
    const int c = blockIdx.y;
    const int plane = height * width;
    const int64_t n0 = (int64_t)part * part_len;
    const int len = (int)min((int64_t)part_len, n_vox - n0);
    const int v_per = (n_views + gridDim.x - 1) / gridDim.x;
    const int v0 = blockIdx.x * v_per, v1 = min(n_views, v0 + v_per);
    const int64_t row = (int64_t)c * n_vox + n0;
    for (int n = threadIdx.x; n < len; n += blockDim.x)
        s_coef[n] = bwd_coef(mean[row + n], g_cov != nullptr ? cov[row + n] : 0.0f, count[n0 + n],
                             g_mean != nullptr ? g_mean[row + n] : 0.0f, g_cov != nullptr ? g_cov[row + n] : 0.0f, n_views_total);
    __syncthreads();
    for (int pb = 0; pb < plane; pb += kBwdPpt * kBwdGatherThreads) {
        int64_t foff[kBwdPpt];
#pragma unroll
        for (int j = 0; j < kBwdPpt; ++j) {
            const int p = min(pb + j * kBwdGatherThreads + (int)threadIdx.x, plane - 1);
            const int y = p / width, x = p - y * width;
            foff[j] = (int64_t)y * sy + (int64_t)x * sx;
        }
        for (int v = v0; v < v1; ++v) {
            const int64_t vp = (int64_t)v * parts + part;
            const uint2 *rc = rec + vp * plane;
            const T *fv = feat + v * sv + (int64_t)c * sc;
            T *gv = g_feat + ((int64_t)v * channels + c) * plane;
            uint2 r[kBwdPpt];
            float x[kBwdPpt], prev[kBwdPpt];
#pragma unroll
            for (int j = 0; j < kBwdPpt; ++j) {                            // all loads of the view up front
                const int p = pb + j * kBwdGatherThreads + threadIdx.x;
                r[j] = p < plane ? __ldg(rc + p) : make_uint2(0u, 0u);
                x[j] = (r[j].x & 0xffffu) ? to_f32<T>(fv[foff[j]]) : 0.0f;
                prev[j] = (part > 0 && p < plane) ? to_f32<T>(gv[p]) : 0.0f;   // what the previous part's launch left
            }
#pragma unroll
            for (int j = 0; j < kBwdPpt; ++j) {
                const int p = pb + j * kBwdGatherThreads + threadIdx.x;
                if (p >= plane) continue;
                const int cnt = (int)(r[j].x & 0xffffu);
                float A = 0.0f, B = 0.0f;                                  // ascending voxel index: one order, deterministic
                if (cnt > 0) { const float2 ab = s_coef[r[j].x >> 16]; A += ab.x; B += ab.y; }
                if (cnt > 1) { const float2 ab = s_coef[r[j].y & 0xffffu]; A += ab.x; B += ab.y; }
                if (cnt > 2) { const float2 ab = s_coef[r[j].y >> 16]; A += ab.x; B += ab.y; }
                if (cnt > 3) {
                    const uint16_t *ls = list + vp * part_len + start[vp * (plane + 1) + p];
                    for (int k = 3; k < cnt; ++k) {
                        const float2 ab = s_coef[__ldg(ls + k)];
                        A += ab.x;
                        B += ab.y;
                    }
                }
                gv[p] = from_f32_t<T>(prev[j] + (cnt > 0 ? fmaf(x[j], B, A) : 0.0f));
            }
        }
    }
}

}  // namespace nd

using namespace nd;

extern "C" {

static int bwd_parts(int64_t n_voxels) { return (int)ceil_div(n_voxels, (int64_t)kBwdPartMax); }
static int bwd_part_len(int64_t n_voxels) { return (int)ceil_div(n_voxels, (int64_t)bwd_parts(n_voxels)); }

size_t nd_lift_backward_workspace_bytes(const nd_maps *features, int64_t n_voxels) {
    if (features == nullptr || n_voxels <= 0) return 0;
    const size_t idx = ((size_t)features->n_views * n_voxels * sizeof(uint16_t) + 255) & ~(size_t)255;
    const size_t plane = (size_t)features->height * features->width;
    const size_t parts = (size_t)bwd_parts(n_voxels);
    const size_t st = ((size_t)features->n_views * parts * (plane + 1) * sizeof(uint16_t) + 255) & ~(size_t)255;
    const size_t rc = ((size_t)features->n_views * parts * plane * sizeof(uint2) + 255) & ~(size_t)255;
    const size_t ls = ((size_t)features->n_views * parts * (size_t)bwd_part_len(n_voxels) * sizeof(uint16_t) + 255) & ~(size_t)255;
    return idx + st + ls + rc;
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
    ND_REQUIRE(plane <= kBwdCsrPixels, ND_ERR_BAD_SHAPE, "nd_lift_backward: planes of %lld pixels (limit %d)", (long long)plane,
               kBwdCsrPixels);
    ND_REQUIRE(workspace_bytes >= nd_lift_backward_workspace_bytes(features, n_voxels) &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               ND_ERR_WORKSPACE, "nd_lift_backward: workspace too small or not 256-byte aligned");
    if (n_views_total <= 0) n_views_total = nv;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    uint16_t *idx = reinterpret_cast<uint16_t *>(ws);
    const size_t idx_bytes = ((size_t)nv * n_voxels * sizeof(uint16_t) + 255) & ~(size_t)255;
    k_bwd_index<<<dim3((unsigned)ceil_div(n_voxels, 256), nv), 256, 0, st>>>(points, projection, n_voxels, h, w, depth_resized,
                                                                             voxel_z, idx);
    ND_CUDA_LAUNCH_CHECK("k_bwd_index");
    {
        const int parts = bwd_parts(n_voxels), part_len = bwd_part_len(n_voxels);
        const size_t st_bytes = ((size_t)nv * parts * (size_t)(plane + 1) * sizeof(uint16_t) + 255) & ~(size_t)255;
        const size_t ls_bytes = ((size_t)nv * parts * (size_t)part_len * sizeof(uint16_t) + 255) & ~(size_t)255;
        uint16_t *start = reinterpret_cast<uint16_t *>(ws + idx_bytes), *list = reinterpret_cast<uint16_t *>(ws + idx_bytes + st_bytes);
        uint2 *rec = reinterpret_cast<uint2 *>(ws + idx_bytes + st_bytes + ls_bytes);
        const size_t csr_smem = (size_t)(2 * plane + 1) * sizeof(uint32_t) + (size_t)part_len * sizeof(uint16_t);
        cudaError_t e = cudaFuncSetAttribute(k_bwd_csr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csr_smem);
        if (e == cudaSuccess)
            k_bwd_csr<<<dim3(parts, nv), kBwdCsrThreads, csr_smem, st>>>(idx, n_voxels, part_len, (int)plane, start, list, rec);
        const size_t g_smem = (size_t)part_len * sizeof(float2);
        // enough CTAs for ~3 rounds on the SMs (one CTA per SM: the coefficient row takes the shared memory)
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int vsplit = (int)ceil_div((int64_t)3 * sms, (int64_t)ch);
        vsplit = vsplit < 1 ? 1 : vsplit > nv ? nv : vsplit;
        for (int part = 0; part < parts && e == cudaSuccess; ++part) {
            if (features->dtype == ND_F32) {
                auto kern = k_bwd_gather<float>;
                e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem);
                if (e == cudaSuccess)
                    kern<<<dim3(vsplit, ch), kBwdGatherThreads, g_smem, st>>>((const float *)features->data, features->stride_v, features->stride_c,
                                                                features->stride_y, features->stride_x, nv, ch, h, w, mean, cov, count,
                                                                grad_mean, grad_cov, n_voxels, n_views_total, part, parts, part_len,
                                                                rec, start, list, (float *)grad_features);
            } else {
                auto kern = k_bwd_gather<__nv_bfloat16>;
                e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem);
                if (e == cudaSuccess)
                    kern<<<dim3(vsplit, ch), kBwdGatherThreads, g_smem, st>>>((const __nv_bfloat16 *)features->data, features->stride_v,
                                                                features->stride_c, features->stride_y, features->stride_x, nv, ch, h, w,
                                                                mean, cov, count, grad_mean, grad_cov, n_voxels, n_views_total, part,
                                                                parts, part_len, rec, start, list, (__nv_bfloat16 *)grad_features);
            }
        }
        if (e != cudaSuccess) {
            set_error("nd_lift_backward: %s", cudaGetErrorString(e));
            return ND_ERR_CUDA;
        }
        ND_CUDA_LAUNCH_CHECK("k_bwd_gather");
    }
    return ND_OK;
}

}  // extern "C"
