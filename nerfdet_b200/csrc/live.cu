// Live 35-channel voxel statistics (B8 + B9 of SURVEY.md section 8a; reference nerfdet.py:200-210, 232-253).
//
// For every voxel n and view v the reference gathers
//   * 3 RGB channels from the stride-1 projection onto denorm_images[:, :, :h, :w]  (0 where that projection is invalid),
//   * Cm = 32 mapped channels: mapping(volume) where `volume` is 0 for invalid views, i.e. the Linear bias
//     (SURVEY.md section 0.6); gather-of-mapped-2D == map-of-gathered-3D bit-identically (section 8a B9),
//     so the mapped 2-D maps (B7) are gathered directly,
// and reduces them over views with the FEATURE-level count:
//   mean = sum / (count + 1e-8)            (NOT zeroed where count == 0)
//   var  = sum over ALL views (h - mean)^2 / (count + 1e-8), 1e6 where count == 0;  cov = exp(-var)
// The MLP input row of voxel n is channel-INTERLEAVED [m0, c0, m1, c1, ...].
//
// One thread per voxel, 2 * 35 accumulators in registers; the sources (76 MB at nv = 50) are L2-resident
// and adjacent voxels (Z fastest) hit neighbouring pixels.
#include "nd_common.cuh"

namespace nd {

constexpr int kLiveMaxCm = 32;

// B4 (nerfdet.py:405-411) for a voxel-view that passed project_nearest: |z - depth[v][y][x]| < voxel_z with the depth map
// already resized to this level's resolution; no depth map = no gate.  Same arithmetic as k_backproject / k_q_index.
__device__ __forceinline__ bool depth_keeps(const float *__restrict__ depth, int v, int height, int width, float xr, float yr,
                                            float z, float voxel_z) {
    if (depth == nullptr) return true;
    const float d = __ldg(depth + ((int64_t)v * height + (int)yr) * width + (int)xr);
    return (z > __fsub_rn(d, voxel_z)) && (z < __fadd_rn(d, voxel_z));
}

template <typename T>
__global__ void __launch_bounds__(128)
k_live_stats(const T *__restrict__ mapped, int64_t m_sv, int64_t m_sc, int64_t m_sy, int64_t m_sx, int cm, int hf, int wf,
             const float *__restrict__ rgb, int64_t r_sv, int64_t r_sc, int64_t r_sy, int64_t r_sx, int hr, int wr,
             const float *__restrict__ points, const float *__restrict__ proj_f, const float *__restrict__ proj_r, int nv,
             int64_t n_vox, const float *__restrict__ bias, const float *__restrict__ depth_f,
             const float *__restrict__ depth_r, float voxel_z, float *__restrict__ glob, float *__restrict__ mean_out,
             float *__restrict__ cov_out, int64_t *__restrict__ count_out) {
    extern __shared__ float sp[];                       // [nv][12] feature-level, [nv][12] rgb-level, [cm] bias
    float *spr = sp + nv * 12;
    float *sb = spr + nv * 12;
    for (int i = threadIdx.x; i < nv * 12; i += blockDim.x) {
        sp[i] = proj_f[i];
        spr[i] = proj_r[i];
    }
    for (int i = threadIdx.x; i < cm; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_vox) return;
    const float X = points[n], Y = points[n_vox + n], Z = points[2 * n_vox + n];

    float s1[3 + kLiveMaxCm], s2[3 + kLiveMaxCm];
#pragma unroll
    for (int k = 0; k < 3 + kLiveMaxCm; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    int cnt = 0;
    for (int v = 0; v < nv; ++v) {
        float xr, yr, q2;
        const bool ok_f = project_nearest(sp + v * 12, X, Y, Z, hf, wf, xr, yr, q2) &&
                          depth_keeps(depth_f, v, hf, wf, xr, yr, q2, voxel_z);
        const int64_t off_f = (int64_t)v * m_sv + (int64_t)(int)yr * m_sy + (int64_t)(int)xr * m_sx;
        float xr2, yr2, q22;
        const bool ok_r = project_nearest(spr + v * 12, X, Y, Z, hr, wr, xr2, yr2, q22) &&
                          depth_keeps(depth_r, v, hr, wr, xr2, yr2, q22, voxel_z);
        const int64_t off_r = (int64_t)v * r_sv + (int64_t)(int)yr2 * r_sy + (int64_t)(int)xr2 * r_sx;
        cnt += ok_f ? 1 : 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float h = ok_r ? __ldg(rgb + off_r + k * r_sc) : 0.0f;
            s1[k] += h;
            s2[k] = fmaf(h, h, s2[k]);
        }
#pragma unroll
        for (int k = 0; k < kLiveMaxCm; ++k) {
            if (k < cm) {
                const float h = ok_f ? to_f32<T>(mapped[off_f + k * m_sc]) : sb[k];
                s1[3 + k] += h;
                s2[3 + k] = fmaf(h, h, s2[3 + k]);
            }
        }
    }
    const int ct = 3 + cm;
    const float denom = __fadd_rn((float)cnt, 1e-8f);   // count + 1e-8 in fp32 (== count for count >= 1)
    float *row = glob + n * (int64_t)(2 * ct);
#pragma unroll
    for (int k = 0; k < 3 + kLiveMaxCm; ++k) {
        if (k < ct) {
            const float m = s1[k] / denom;
            float cv = 0.0f;                             // exp(-1e6) == 0 where count == 0 (nerfdet.py:249-250)
            if (cnt > 0) {
                // sum over all views of (h - m)^2 = S2 - 2 m S1 + nv m^2
                float ssd = fmaf(-2.0f * m, s1[k], s2[k]);
                ssd = fmaxf(fmaf((float)nv * m, m, ssd), 0.0f);
                cv = expf(-(ssd / denom));
            }
            row[2 * k] = m;
            row[2 * k + 1] = cv;
            if (mean_out != nullptr) mean_out[(int64_t)k * n_vox + n] = m;
            if (cov_out != nullptr) cov_out[(int64_t)k * n_vox + n] = cv;
        }
    }
    if (count_out != nullptr) count_out[n] = cnt;
}

// Channels-last build (the product path; `mapped` as [nv][h][w][Cm], which is what the per-pixel Linear of B7
// produces): one WARP per voxel, lane = mapped channel (lanes 0-2 also carry RGB).  The two projections run with
// lane = view; only the valid views are visited, each mapped gather is one coalesced row of Cm elements.  Invalid
// views contribute the bias (mapped) or 0 (RGB); their share is added in closed form after the loop.
constexpr int kLcWarps = 4, kLcPerWarp = 1;     // small blocks: per-voxel cost varies 10x across the room, the
                                                // hardware block scheduler does the balancing

template <typename T>
__global__ void __launch_bounds__(kLcWarps * 32)
k_live_stats_cl(const T *__restrict__ mapped, int64_t m_sv, int64_t m_sy, int64_t m_sx, int cm, int hf, int wf,
                const float *__restrict__ rgb, int64_t r_sv, int64_t r_sc, int64_t r_sy, int64_t r_sx, int hr, int wr,
                const float *__restrict__ points, const float *__restrict__ proj_f, const float *__restrict__ proj_r, int nv,
                int64_t n_vox, const float *__restrict__ bias, const float *__restrict__ depth_f,
                const float *__restrict__ depth_r, float voxel_z, float *__restrict__ glob, float *__restrict__ mean_out,
                float *__restrict__ cov_out, int64_t *__restrict__ count_out) {
    extern __shared__ float sp[];                       // [nv][12] feature-level, [nv][12] rgb-level
    float *spr = sp + nv * 12;
    for (int i = threadIdx.x; i < nv * 12; i += blockDim.x) {
        sp[i] = proj_f[i];
        spr[i] = proj_r[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const float b = lane < cm ? bias[lane] : 0.0f;
    const int ct = 3 + cm;
    for (int it = 0; it < kLcPerWarp; ++it) {
        const int64_t n = ((int64_t)blockIdx.x * kLcWarps + warp) * kLcPerWarp + it;
        if (n >= n_vox) break;                                             // warp-uniform
        const float X = __ldg(points + n), Y = __ldg(points + n_vox + n), Z = __ldg(points + 2 * n_vox + n);
        float s1 = 0.f, s2 = 0.f, r1 = 0.f, r2 = 0.f;
        int cnt = 0;
        for (int v0 = 0; v0 < nv; v0 += 32) {
            const int v = v0 + lane;
            bool ok_f = false, ok_r = false;
            int off_f = 0, off_r = 0;
            if (v < nv) {
                float xr, yr, q2;
                ok_f = project_nearest(sp + v * 12, X, Y, Z, hf, wf, xr, yr, q2) &&
                       depth_keeps(depth_f, v, hf, wf, xr, yr, q2, voxel_z);
                if (ok_f) off_f = (int)((int)yr * m_sy + (int)xr * m_sx);
                ok_r = project_nearest(spr + v * 12, X, Y, Z, hr, wr, xr, yr, q2) &&
                       depth_keeps(depth_r, v, hr, wr, xr, yr, q2, voxel_z);
                if (ok_r) off_r = (int)((int)yr * r_sy + (int)xr * r_sx);
            }
            const unsigned bf = __ballot_sync(full, ok_f), br = __ballot_sync(full, ok_r);
            cnt += __popc(bf);
            unsigned act = bf | br;
            while (act) {
                const int src = __ffs(act) - 1;
                act &= act - 1;
                const int64_t vv = v0 + src;
                if ((bf >> src) & 1u) {
                    const int of = __shfl_sync(full, off_f, src);
                    if (lane < cm) {
                        const float hv = to_f32<T>(mapped[vv * m_sv + of + lane]);
                        s1 += hv;
                        s2 = fmaf(hv, hv, s2);
                    }
                }
                if ((br >> src) & 1u) {
                    const int orr = __shfl_sync(full, off_r, src);
                    if (lane < 3) {
                        const float hv = __ldg(rgb + vv * r_sv + lane * r_sc + orr);
                        r1 += hv;
                        r2 = fmaf(hv, hv, r2);
                    }
                }
            }
        }
        // invalid views enter the mapped statistics as the Linear bias
        const float n_inv = (float)(nv - cnt);
        s1 = fmaf(n_inv, b, s1);
        s2 = fmaf(n_inv * b, b, s2);
        const float denom = __fadd_rn((float)cnt, 1e-8f);
        float *row = glob + n * (int64_t)(2 * ct);
        auto finish = [&](int k, float a1, float a2) {
            const float m = a1 / denom;
            float cv = 0.0f;
            if (cnt > 0) {
                float ssd = fmaf(-2.0f * m, a1, a2);
                ssd = fmaxf(fmaf((float)nv * m, m, ssd), 0.0f);
                cv = expf(-(ssd / denom));
            }
            *reinterpret_cast<float2 *>(row + 2 * k) = make_float2(m, cv);
            if (mean_out != nullptr) mean_out[(int64_t)k * n_vox + n] = m;
            if (cov_out != nullptr) cov_out[(int64_t)k * n_vox + n] = cv;
        };
        if (lane < 3) finish(lane, r1, r2);
        if (lane < cm) finish(3 + lane, s1, s2);
        if (lane == 0 && count_out != nullptr) count_out[n] = cnt;
    }
}

// Backward of the mapped channels of the same statistics (row N1): with x_u the value of view u (the mapped feature for a
// valid view, the Linear bias for an invalid one), S1 = sum_u x_u, m = S1 / denom, var = sum_u (x_u - m)^2 / denom,
//   g_x = g_m / denom + g_var * (2 (x_u - m) / denom - 2 (S1 - nv m) / denom^2),   g_var = -exp(-var) * g_cov (0 if count = 0)
// for EVERY view: valid views scatter it to their pixel of the channels-last gradient (one coalesced 128-byte reduction per
// warp), the nv - count invalid views add theirs to the gradient of the bias.  Same walk as the forward kernel, twice: once
// for S1, once to scatter.  The RGB channels carry no gradient (input images).
__global__ void __launch_bounds__(kLcWarps * 32)
k_live_stats_bwd(const float *__restrict__ mapped, int64_t m_sv, int64_t m_sy, int64_t m_sx, int cm, int hf, int wf,
                 const float *__restrict__ points, const float *__restrict__ proj_f, int nv, int64_t n_vox,
                 const float *__restrict__ bias, const float *__restrict__ depth_f, float voxel_z,
                 const float *__restrict__ glob, const float *__restrict__ g_glob,
                 float *__restrict__ g_mapped, float *__restrict__ g_bias) {
    extern __shared__ float sp[];                       // [nv][12] feature-level
    for (int i = threadIdx.x; i < nv * 12; i += blockDim.x) sp[i] = proj_f[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const float b = lane < cm ? bias[lane] : 0.0f;
    const int ct = 3 + cm;
    const int64_t n = (int64_t)blockIdx.x * kLcWarps + warp;
    if (n >= n_vox) return;                                                // warp-uniform
    const float X = __ldg(points + n), Y = __ldg(points + n_vox + n), Z = __ldg(points + 2 * n_vox + n);
    float cm_ = 0.f, cb = 0.f, mean = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
        float s1 = 0.f;
        int cnt = 0;
        for (int v0 = 0; v0 < nv; v0 += 32) {
            const int v = v0 + lane;
            bool ok_f = false;
            int off_f = 0;
            if (v < nv) {
                float xr, yr, q2;
                ok_f = project_nearest(sp + v * 12, X, Y, Z, hf, wf, xr, yr, q2) &&
                       depth_keeps(depth_f, v, hf, wf, xr, yr, q2, voxel_z);
                if (ok_f) off_f = (int)((int)yr * m_sy + (int)xr * m_sx);
            }
            unsigned act = __ballot_sync(full, ok_f);
            cnt += __popc(act);
            while (act) {
                const int src = __ffs(act) - 1;
                act &= act - 1;
                const int of = __shfl_sync(full, off_f, src);
                if (lane < cm) {
                    const int64_t o = (int64_t)(v0 + src) * m_sv + of + lane;
                    const float hv = mapped[o];
                    if (pass == 0) s1 += hv;
                    else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g_mapped + o), "f"(fmaf(cb, hv - mean, cm_)) : "memory");
                }
            }
        }
        if (pass == 0) {
            const float n_inv = (float)(nv - cnt);
            s1 = fmaf(n_inv, b, s1);
            const float denom = __fadd_rn((float)cnt, 1e-8f);
            const float *row = glob + n * (int64_t)(2 * ct), *grow = g_glob + n * (int64_t)(2 * ct);
            if (lane < cm) {
                mean = row[2 * (3 + lane)];
                const float gvar = cnt > 0 ? -row[2 * (3 + lane) + 1] * grow[2 * (3 + lane) + 1] : 0.0f;
                cb = 2.0f * gvar / denom;
                cm_ = grow[2 * (3 + lane)] / denom - cb * (s1 - (float)nv * mean) / denom;
                // the invalid views of this voxel: x_u = bias
                if (n_inv > 0.0f) atomicAdd(g_bias + lane, n_inv * fmaf(cb, b - mean, cm_));
            }
        }
    }
}

}  // namespace nd

using namespace nd;

extern "C" int nd_live_stats_gated(const nd_maps *mapped, const nd_maps *rgb, const float *points, const float *projection,
                                   const float *rgb_projection, int64_t n_voxels, const float *map_bias,
                                   const float *depth_mapped, const float *depth_rgb, float voxel_z, float *global_volume,
                                   float *mean35, float *cov35, int64_t *count, void *stream) {
    ND_REQUIRE((depth_mapped == nullptr) == (depth_rgb == nullptr), ND_ERR_BAD_ARG,
               "nd_live_stats_gated: the depth gate needs the depth map at both resolutions (or neither)");
    ND_REQUIRE(mapped && rgb && mapped->data && rgb->data && points && projection && rgb_projection && map_bias &&
                   global_volume,
               ND_ERR_BAD_ARG, "nd_live_stats: null pointer");
    ND_REQUIRE(mapped->n_views == rgb->n_views && mapped->n_views > 0, ND_ERR_BAD_SHAPE,
               "nd_live_stats: %d mapped views vs %d rgb views", mapped->n_views, rgb->n_views);
    ND_REQUIRE(rgb->channels == 3 && rgb->dtype == ND_F32, ND_ERR_BAD_SHAPE, "nd_live_stats: rgb must be f32 with 3 channels");
    ND_REQUIRE(mapped->channels >= 1 && mapped->channels <= kLiveMaxCm, ND_ERR_BAD_SHAPE,
               "nd_live_stats: %d mapped channels (max %d)", mapped->channels, kLiveMaxCm);
    ND_REQUIRE(n_voxels >= 0, ND_ERR_BAD_SHAPE, "nd_live_stats: negative voxel count");
    if (n_voxels == 0) return ND_OK;
    const int nv = mapped->n_views;
    const size_t smem = ((size_t)nv * 24 + mapped->channels) * sizeof(float);
    ND_REQUIRE(smem <= 48 * 1024, ND_ERR_BAD_SHAPE, "nd_live_stats: too many views (%d)", nv);
    cudaStream_t st = (cudaStream_t)stream;
    if (mapped->stride_c == 1 && (reinterpret_cast<uintptr_t>(global_volume) & 7) == 0) {      // channels-last: warp per voxel
        const size_t sm = (size_t)nv * 24 * sizeof(float);
        const unsigned g = (unsigned)ceil_div(n_voxels, (int64_t)kLcWarps * kLcPerWarp);
        if (mapped->dtype == ND_F32)
            k_live_stats_cl<float><<<g, kLcWarps * 32, sm, st>>>(
                (const float *)mapped->data, mapped->stride_v, mapped->stride_y, mapped->stride_x, mapped->channels,
                mapped->height, mapped->width, (const float *)rgb->data, rgb->stride_v, rgb->stride_c, rgb->stride_y,
                rgb->stride_x, rgb->height, rgb->width, points, projection, rgb_projection, nv, n_voxels, map_bias,
                depth_mapped, depth_rgb, voxel_z, global_volume, mean35, cov35, count);
        else
            k_live_stats_cl<__nv_bfloat16><<<g, kLcWarps * 32, sm, st>>>(
                (const __nv_bfloat16 *)mapped->data, mapped->stride_v, mapped->stride_y, mapped->stride_x, mapped->channels,
                mapped->height, mapped->width, (const float *)rgb->data, rgb->stride_v, rgb->stride_c, rgb->stride_y,
                rgb->stride_x, rgb->height, rgb->width, points, projection, rgb_projection, nv, n_voxels, map_bias,
                depth_mapped, depth_rgb, voxel_z, global_volume, mean35, cov35, count);
        ND_CUDA_LAUNCH_CHECK("k_live_stats_cl");
        return ND_OK;
    }
    const unsigned grid = (unsigned)ceil_div(n_voxels, 128);
    if (mapped->dtype == ND_F32)
        k_live_stats<float><<<grid, 128, smem, st>>>(
            (const float *)mapped->data, mapped->stride_v, mapped->stride_c, mapped->stride_y, mapped->stride_x,
            mapped->channels, mapped->height, mapped->width, (const float *)rgb->data, rgb->stride_v, rgb->stride_c,
            rgb->stride_y, rgb->stride_x, rgb->height, rgb->width, points, projection, rgb_projection, nv, n_voxels,
            map_bias, depth_mapped, depth_rgb, voxel_z, global_volume, mean35, cov35, count);
    else
        k_live_stats<__nv_bfloat16><<<grid, 128, smem, st>>>(
            (const __nv_bfloat16 *)mapped->data, mapped->stride_v, mapped->stride_c, mapped->stride_y, mapped->stride_x,
            mapped->channels, mapped->height, mapped->width, (const float *)rgb->data, rgb->stride_v, rgb->stride_c,
            rgb->stride_y, rgb->stride_x, rgb->height, rgb->width, points, projection, rgb_projection, nv, n_voxels,
            map_bias, depth_mapped, depth_rgb, voxel_z, global_volume, mean35, cov35, count);
    ND_CUDA_LAUNCH_CHECK("k_live_stats");
    return ND_OK;
}

extern "C" int nd_live_stats(const nd_maps *mapped, const nd_maps *rgb, const float *points, const float *projection,
                             const float *rgb_projection, int64_t n_voxels, const float *map_bias, float *global_volume,
                             float *mean35, float *cov35, int64_t *count, void *stream) {
    return nd_live_stats_gated(mapped, rgb, points, projection, rgb_projection, n_voxels, map_bias, nullptr, nullptr, 0.0f,
                               global_volume, mean35, cov35, count, stream);
}

extern "C" int nd_live_stats_bwd_gated(const nd_maps *mapped, const float *points, const float *projection, int64_t n_voxels,
                                       const float *map_bias, const float *depth_mapped, float voxel_z,
                                       const float *global_volume, const float *grad_global_volume, float *grad_mapped,
                                       float *grad_bias, void *stream) {
    ND_REQUIRE(mapped && mapped->data && points && projection && map_bias && global_volume && grad_global_volume && grad_mapped &&
                   grad_bias,
               ND_ERR_BAD_ARG, "nd_live_stats_bwd: null pointer");
    ND_REQUIRE(mapped->dtype == ND_F32 && mapped->stride_c == 1 && mapped->channels >= 1 && mapped->channels <= kLiveMaxCm &&
                   mapped->n_views > 0 && n_voxels >= 0,
               ND_ERR_BAD_SHAPE, "nd_live_stats_bwd: channels-last f32 maps with at most %d channels", kLiveMaxCm);
    if (n_voxels == 0) return ND_OK;
    const int nv = mapped->n_views;
    const size_t sm = (size_t)nv * 12 * sizeof(float);
    ND_REQUIRE(sm <= 48 * 1024, ND_ERR_BAD_SHAPE, "nd_live_stats_bwd: too many views (%d)", nv);
    k_live_stats_bwd<<<(unsigned)ceil_div(n_voxels, (int64_t)kLcWarps), kLcWarps * 32, sm, (cudaStream_t)stream>>>(
        (const float *)mapped->data, mapped->stride_v, mapped->stride_y, mapped->stride_x, mapped->channels, mapped->height,
        mapped->width, points, projection, nv, n_voxels, map_bias, depth_mapped, voxel_z, global_volume, grad_global_volume,
        grad_mapped, grad_bias);
    ND_CUDA_LAUNCH_CHECK("k_live_stats_bwd");
    return ND_OK;
}

extern "C" int nd_live_stats_bwd(const nd_maps *mapped, const float *points, const float *projection, int64_t n_voxels,
                                 const float *map_bias, const float *global_volume, const float *grad_global_volume,
                                 float *grad_mapped, float *grad_bias, void *stream) {
    return nd_live_stats_bwd_gated(mapped, points, projection, n_voxels, map_bias, nullptr, 0.0f, global_volume,
                                   grad_global_volume, grad_mapped, grad_bias, stream);
}
