// Row N4 of SURVEY.md section 8f: the image metrics of the render_testing evaluator as one batched GPU pass
// (reference mmdet3d/models/model_utils/save_rendered_img.py:10-78).
//
//   PSNR  compute_psnr (:13-20): -10 log10(mean((pred - target)^2)) over one view's H x W x 3 values, in float64 (the
//         reference's gt_rgb is float64, so torch promotes the difference).
//   SSIM  compute_ssim (:22-38) -> skimage.metrics.structural_similarity of scikit-image 0.18.1 (requirements/runtime.txt:7)
//         as the reference ends up calling it (multichannel=True after the ValueError of the first attempt): per channel, 7 x 7
//         uniform window, sample covariance (x 49/48), data_range = 2 (the float dtype range), K1 = 0.01, K2 = 0.03, all in
//         float64, the SSIM map cropped by 3 pixels on every side and averaged; then the mean over the channels.
//   RMSE  the reference's "rsme" (:51, :78): the per-pixel squared depth error averaged over the views (a map, not a scalar).
//
// One thread per (view, channel, interior pixel) gathers its 49 pixel pairs; block sums go to a partials array that one block
// per view adds in a fixed order, so the numbers are deterministic.
#include <algorithm>

#include "nd_common.cuh"

namespace nd {

constexpr int kMetricThreads = 128;
constexpr int kSsimWin = 7, kSsimPad = 3;

__device__ __forceinline__ double block_sum(double v, double *s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)blockDim.x / 32; ++w) t += s_warp[w];
    __syncthreads();
    return t;                                                              // valid on thread 0
}

// grid (blocks per view, nv).  partial[v][block] = {sum of SSIM values of the block's interior pixels, sum of squared errors of
// the block's pixels}.  pred / target: [nv][H][W][3].
template <typename TP, typename TT>
__global__ void __launch_bounds__(kMetricThreads)
k_image_metrics(const TP *__restrict__ pred, const TT *__restrict__ target, int height, int width, double c1, double c2,
                double *__restrict__ partial) {
    __shared__ double s_warp[kMetricThreads / 32];
    const int v = blockIdx.y;
    const int64_t plane = (int64_t)height * width;
    const TP *p = pred + (int64_t)v * plane * 3;
    const TT *t = target + (int64_t)v * plane * 3;
    const int ih = height - 2 * kSsimPad, iw = width - 2 * kSsimPad;       // interior the cropped SSIM map covers
    const int64_t n_int = ih > 0 && iw > 0 ? (int64_t)ih * iw * 3 : 0;
    double ssim = 0.0, se = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane * 3; i += (int64_t)gridDim.x * blockDim.x) {
        const double d = (double)p[i] - (double)t[i];
        se += d * d;
        if (i < n_int) {
            const int c = (int)(i % 3);
            const int64_t q = i / 3;
            const int y = (int)(q / iw) + kSsimPad, x = (int)(q - (int64_t)(q / iw) * iw) + kSsimPad;
            double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
            for (int dy = -kSsimPad; dy <= kSsimPad; ++dy)
                for (int dx = -kSsimPad; dx <= kSsimPad; ++dx) {
                    const int64_t o = ((int64_t)(y + dy) * width + (x + dx)) * 3 + c;
                    const double a = (double)p[o], b = (double)t[o];
                    sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
                }
            const double np_ = (double)(kSsimWin * kSsimWin), cov_norm = np_ / (np_ - 1.0);
            const double ux = sx / np_, uy = sy / np_, uxx = sxx / np_, uyy = syy / np_, uxy = sxy / np_;
            const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
            ssim += ((2.0 * ux * uy + c1) * (2.0 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
        }
    }
    const double a = block_sum(ssim, s_warp);
    const double b = block_sum(se, s_warp);
    if (threadIdx.x == 0) {
        partial[((int64_t)v * gridDim.x + blockIdx.x) * 2] = a;
        partial[((int64_t)v * gridDim.x + blockIdx.x) * 2 + 1] = b;
    }
}

// one block per view: partials added in block order; out[v] = {psnr, ssim}
__global__ void k_metrics_finish(const double *__restrict__ partial, int blocks, int height, int width, double *__restrict__ out) {
    const int v = blockIdx.x;
    if (threadIdx.x != 0) return;
    double ssim = 0.0, se = 0.0;
    for (int b = 0; b < blocks; ++b) {
        ssim += partial[((int64_t)v * blocks + b) * 2];
        se += partial[((int64_t)v * blocks + b) * 2 + 1];
    }
    const double n = (double)height * width * 3.0;
    const int ih = height - 2 * kSsimPad, iw = width - 2 * kSsimPad;
    out[v * 2] = -10.0 * log(se / n) / log(10.0);
    out[v * 2 + 1] = ih > 0 && iw > 0 ? ssim / ((double)ih * iw * 3.0) : nan("");
}

template <typename TP, typename TT>
__global__ void k_depth_sqerr(const TP *__restrict__ depth, const TT *__restrict__ gt, int nv, int64_t plane,
                              double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
    double s = 0.0;
    for (int v = 0; v < nv; ++v) {                                         // views in order, like the reference's loop
        const double d = (double)depth[(int64_t)v * plane + i] - (double)gt[(int64_t)v * plane + i];
        s += d * d;
    }
    out[i] = s / (double)nv;
}

}  // namespace nd

using namespace nd;

extern "C" {

size_t nd_image_metrics_workspace_bytes(int n_views, int height, int width) {
    if (n_views <= 0 || height <= 0 || width <= 0) return 0;
    const int64_t blocks = std::min<int64_t>(ceil_div((int64_t)height * width * 3, kMetricThreads), 256);
    return (size_t)n_views * blocks * 2 * sizeof(double);
}

int nd_image_metrics(const float *pred, const void *target, int target_is_f64, int n_views, int height, int width,
                     double data_range, double *psnr_ssim, void *workspace, size_t workspace_bytes, void *stream) {
    ND_REQUIRE(pred && target && psnr_ssim && workspace, ND_ERR_BAD_ARG, "nd_image_metrics: null pointer");
    ND_REQUIRE(n_views > 0 && height > 0 && width > 0 && data_range > 0, ND_ERR_BAD_SHAPE, "nd_image_metrics: bad shape");
    ND_REQUIRE(workspace_bytes >= nd_image_metrics_workspace_bytes(n_views, height, width), ND_ERR_WORKSPACE,
               "nd_image_metrics: workspace too small");
    const int blocks = (int)std::min<int64_t>(ceil_div((int64_t)height * width * 3, kMetricThreads), 256);
    const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = static_cast<double *>(workspace);
    if (target_is_f64)
        k_image_metrics<float, double><<<dim3(blocks, n_views), kMetricThreads, 0, st>>>(pred, (const double *)target, height, width,
                                                                                          c1, c2, partial);
    else
        k_image_metrics<float, float><<<dim3(blocks, n_views), kMetricThreads, 0, st>>>(pred, (const float *)target, height, width,
                                                                                         c1, c2, partial);
    ND_CUDA_LAUNCH_CHECK("k_image_metrics");
    k_metrics_finish<<<n_views, 32, 0, st>>>(partial, blocks, height, width, psnr_ssim);
    ND_CUDA_LAUNCH_CHECK("k_metrics_finish");
    return ND_OK;
}

int nd_depth_sqerr(const float *depth, const void *gt_depth, int gt_is_f64, int n_views, int64_t n_pixels, double *out,
                   void *stream) {
    ND_REQUIRE(depth && gt_depth && out, ND_ERR_BAD_ARG, "nd_depth_sqerr: null pointer");
    ND_REQUIRE(n_views > 0 && n_pixels >= 0, ND_ERR_BAD_SHAPE, "nd_depth_sqerr: bad shape");
    if (n_pixels == 0) return ND_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (gt_is_f64)
        k_depth_sqerr<float, double><<<(unsigned)ceil_div(n_pixels, 256), 256, 0, st>>>(depth, (const double *)gt_depth, n_views,
                                                                                        n_pixels, out);
    else
        k_depth_sqerr<float, float><<<(unsigned)ceil_div(n_pixels, 256), 256, 0, st>>>(depth, (const float *)gt_depth, n_views,
                                                                                       n_pixels, out);
    ND_CUDA_LAUNCH_CHECK("k_depth_sqerr");
    return ND_OK;
}

}  // extern "C"
