// Row N3 of SURVEY.md section 8f: the two host-side producers of the render branch's inputs, moved onto the GPU.
//
//   k_generate_rays   datasets/pipelines/multi_view.py:124-132 + data_augment_utils.py:410-424 (get_dtu_raydir) +
//                     formating.py:70-75: for every target view the un-normalised ray directions of the pixel grid
//                     [margin, W - margin) x [margin, H - margin) and the camera centre repeated per ray.  The reference
//                     builds 660 000 rays x 2 tensors per scene on the host and ships them; here they are written where
//                     render_rays reads them.  Arithmetic as numpy evaluates it: (px + 0.5 - cx) / fx in float32, the
//                     3 x 3 rotation product in float64 (camrotc2w is float64), rounded to float32.
//   k_denorm_images   multi_view.py:107-110 (mmcv.imdenormalize(img, mean, std, to_bgr=True).astype(uint8) / 255.0) +
//                     formating.py:87-91: the [0, 1] source images the NeRF branch samples colours from, derived from the
//                     normalised network input already on the device instead of a second host copy (46 MB at 50 views).
//                     Arithmetic as OpenCV evaluates it (checked against cv2 4.13): the product in float64 rounded to
//                     float32, the sum in float32, truncation to uint8, division by 255 in float64 rounded to float32.
#include "nd_common.cuh"

namespace nd {

struct RayCam {
    float fx, fy, cx, cy;
};

__global__ void k_generate_rays(RayCam k, const double *__restrict__ rot, const float *__restrict__ lightpos, int nt, int gw,
                                int gh, int margin, float *__restrict__ ray_d, float *__restrict__ ray_o) {
    const int64_t npix = (int64_t)gw * gh;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix * nt) return;
    const int t = (int)(i / npix);
    const int64_t p = i - (int64_t)t * npix;
    const int yy = (int)(p / gw), xx = (int)(p - (int64_t)yy * gw);        // meshgrid order: rows = y
    const float px = (float)(margin + xx), py = (float)(margin + yy);
    const float x = __fdiv_rn(__fsub_rn(__fadd_rn(px, 0.5f), k.cx), k.fx);
    const float y = __fdiv_rn(__fsub_rn(__fadd_rn(py, 0.5f), k.cy), k.fy);
    const double *r = rot + (int64_t)t * 9;
#pragma unroll
    for (int j = 0; j < 3; ++j) {                                          // dirs @ rot.T: row j of rot
        const double d = __dadd_rn(__dadd_rn(__dmul_rn((double)x, r[j * 3]), __dmul_rn((double)y, r[j * 3 + 1])), r[j * 3 + 2]);
        ray_d[i * 3 + j] = __double2float_rn(d);
        ray_o[i * 3 + j] = lightpos[t * 3 + j];
    }
}

struct Denorm {
    double std[3];
    float mean[3];
};

__global__ void k_denorm_images(const float *__restrict__ img, Denorm dn, int to_bgr, int64_t n_img, int64_t plane,
                                float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_img * 3 * plane) return;
    const int64_t pl = i / plane;                                          // (image, output channel)
    const int c_out = (int)(pl % 3);
    const int c_in = to_bgr ? 2 - c_out : c_out;                           // cvtColor(RGB2BGR) after the arithmetic
    const float v = img[(pl - c_out + c_in) * plane + (i - pl * plane)];
    const float m = __double2float_rn(__dmul_rn((double)v, dn.std[c_in]));
    const float s = __fadd_rn(m, dn.mean[c_in]);
    const unsigned char u = (unsigned char)(int)s;                         // numpy astype(uint8): truncate, low byte
    out[i] = __double2float_rn((double)u / 255.0);
}

}  // namespace nd

using namespace nd;

extern "C" {

int nd_generate_rays(const float *intrinsic3x3_host, const double *rot, const float *lightpos, int n_target_views, int height,
                     int width, int margin, float *ray_d, float *ray_o, void *stream) {
    ND_REQUIRE(intrinsic3x3_host && rot && lightpos && ray_d && ray_o, ND_ERR_BAD_ARG, "nd_generate_rays: null pointer");
    ND_REQUIRE(n_target_views >= 0 && margin >= 0 && width > 2 * margin && height > 2 * margin, ND_ERR_BAD_SHAPE,
               "nd_generate_rays: bad shape");
    if (n_target_views == 0) return ND_OK;
    const RayCam k{intrinsic3x3_host[0], intrinsic3x3_host[4], intrinsic3x3_host[2], intrinsic3x3_host[5]};
    const int gw = width - 2 * margin, gh = height - 2 * margin;
    const int64_t total = (int64_t)gw * gh * n_target_views;
    k_generate_rays<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(k, rot, lightpos, n_target_views, gw, gh,
                                                                                      margin, ray_d, ray_o);
    ND_CUDA_LAUNCH_CHECK("k_generate_rays");
    return ND_OK;
}

int nd_denorm_images(const float *img, const double *mean3_host, const double *std3_host, int to_bgr, int64_t n_images,
                     int height, int width, float *out, void *stream) {
    ND_REQUIRE(img && mean3_host && std3_host && out, ND_ERR_BAD_ARG, "nd_denorm_images: null pointer");
    ND_REQUIRE(n_images >= 0 && height > 0 && width > 0, ND_ERR_BAD_SHAPE, "nd_denorm_images: bad shape");
    if (n_images == 0) return ND_OK;
    Denorm dn;
    for (int c = 0; c < 3; ++c) {
        dn.std[c] = std3_host[c];
        dn.mean[c] = (float)mean3_host[c];
    }
    const int64_t plane = (int64_t)height * width;
    k_denorm_images<<<(unsigned)ceil_div(n_images * 3 * plane, 256), 256, 0, (cudaStream_t)stream>>>(img, dn, to_bgr, n_images,
                                                                                                     plane, out);
    ND_CUDA_LAUNCH_CHECK("k_denorm_images");
    return ND_OK;
}

}  // extern "C"
