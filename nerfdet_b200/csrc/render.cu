// NeRF-branch kernels around the shared MLP (R2, R4-R8 of SURVEY.md section 8a).
//
//   k_sample_rays          render_ray.py:145-189   z_vals and sample points (jitter numbers supplied by the caller)
//   k_render_gather_stats  projection.py:24-151 + render_ray.py:71-93, 301-303
//                          per ray sample: projection into every source view, bilinear gather of the 3 image
//                          channels and the D mapped-feature channels, masked mean / all-view variance,
//                          globalfeat [P][2*(3+D)] -- the reference's [rays, samples, views, 35] tensor
//                          (917 MB at nv = 50) is never materialised
//   k_composite            render_ray.py:196-247   alpha compositing per ray
//   k_volume_sample        render_ray.py:26-46     trilinear lookup (dead branch in nerfdet, standalone-callable)
#include <algorithm>

#include "nd_common.cuh"

namespace nd {

// -------------------------------------------------------------------------------------------------
// R2
// -------------------------------------------------------------------------------------------------
__global__ void k_sample_rays(const float *__restrict__ ray_o, const float *__restrict__ ray_d, int64_t n_rays, int ns,
                              float near, float far, const float *__restrict__ t_rand, float *__restrict__ pts,
                              float *__restrict__ z_vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays * ns) return;
    const int64_t r = i / ns;
    const int s = (int)(i - r * ns);
    // near_depth * ones, (far - near) / (N - 1), near + i * step: every op rounded like the torch expression
    const float step = __fdiv_rn(__fsub_rn(far, near), (float)(ns - 1));
    auto zi = [&](int k) { return __fadd_rn(near, __fmul_rn((float)k, step)); };
    float z = zi(s);
    if (t_rand != nullptr) {
        const float lower = s == 0 ? z : __fmul_rn(0.5f, __fadd_rn(z, zi(s - 1)));
        const float upper = s == ns - 1 ? z : __fmul_rn(0.5f, __fadd_rn(zi(s + 1), z));
        z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[i]));
    }
    z_vals[i] = z;
#pragma unroll
    for (int k = 0; k < 3; ++k) pts[i * 3 + k] = __fadd_rn(__fmul_rn(z, ray_d[r * 3 + k]), ray_o[r * 3 + k]);
}

// -------------------------------------------------------------------------------------------------
// R4 + R5 + R6
// -------------------------------------------------------------------------------------------------
struct ViewSample {          // what one (point, view) needs for the bilinear gathers of one source
    int32_t base;            // element offset of the north-west corner inside a plane (may be out of range)
    float fx, fy;            // ix - floor(ix), iy - floor(iy)
    uint32_t inb;            // bit0 nw, bit1 ne, bit2 sw, bit3 se in bounds; bit 8: view mask (in-bound & in-front)
};

// ATen grid_sampler_2d, bilinear, zeros padding, align_corners = True
__device__ __forceinline__ ViewSample make_sample(float gx, float gy, int hs, int ws) {
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(ws - 1));
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(hs - 1));
    const float x0f = floorf(ix), y0f = floorf(iy);
    ViewSample s;
    s.fx = ix - x0f;
    s.fy = iy - y0f;
    // clamp before the int conversion: coordinates are clamped to +-1e6 pixels upstream, far outside any plane
    const int x0 = (int)fminf(fmaxf(x0f, -4.0f), (float)(ws + 2));
    const int y0 = (int)fminf(fmaxf(y0f, -4.0f), (float)(hs + 2));
    const bool xin0 = x0 >= 0 && x0 < ws, xin1 = x0 + 1 >= 0 && x0 + 1 < ws;
    const bool yin0 = y0 >= 0 && y0 < hs, yin1 = y0 + 1 >= 0 && y0 + 1 < hs;
    s.inb = (xin0 && yin0 ? 1u : 0u) | (xin1 && yin0 ? 2u : 0u) | (xin0 && yin1 ? 4u : 0u) | (xin1 && yin1 ? 8u : 0u);
    if (!(ix == ix) || !(iy == iy)) s.inb = 0;          // NaN coordinates sample nothing
    s.base = y0 * ws + x0;
    return s;
}

template <typename T>
__device__ __forceinline__ float bilinear(const T *__restrict__ plane, const ViewSample &s, int ws) {
    const float nw = (1.0f - s.fx) * (1.0f - s.fy), ne = s.fx * (1.0f - s.fy);
    const float sw = (1.0f - s.fx) * s.fy, se = s.fx * s.fy;
    float acc = 0.0f;
    if (s.inb & 1u) acc = fmaf(to_f32<T>(plane[s.base]), nw, acc);
    if (s.inb & 2u) acc = fmaf(to_f32<T>(plane[s.base + 1]), ne, acc);
    if (s.inb & 4u) acc = fmaf(to_f32<T>(plane[s.base + ws]), sw, acc);
    if (s.inb & 8u) acc = fmaf(to_f32<T>(plane[s.base + ws + 1]), se, acc);
    return acc;
}

constexpr int kRgThreads = 128;
constexpr int kRgChunk = 8;          // feature channels per pass (3 accumulators each)

// One thread per ray sample.  Pass 0 projects the sample into every view and parks the two
// ViewSamples (image, feature map) in shared memory; the channel passes then walk the views.
template <typename T>
__global__ void __launch_bounds__(kRgThreads)
k_render_gather_stats(const float *__restrict__ pts, int64_t n_pts, const float *__restrict__ cams, int nv,
                      const float *__restrict__ img, int64_t i_sv, int64_t i_sc, int hi, int wi,
                      const T *__restrict__ feat, int64_t f_sv, int64_t f_sc, int d, int hf, int wf,
                      float *__restrict__ glob, uint8_t *__restrict__ view_mask, uint8_t *__restrict__ pixel_mask,
                      float *__restrict__ pix_out, uint8_t *__restrict__ front_out, float *__restrict__ view_feat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sP = reinterpret_cast<float *>(smem_raw);                       // [nv][12] rows 0-2 of K @ E
    ViewSample *sI = reinterpret_cast<ViewSample *>(sP + nv * 12);        // [nv][threads] image samples
    ViewSample *sF = sI + (size_t)nv * kRgThreads;                        // [nv][threads] feature-map samples
    // P = K4 @ E (projection.py:57, a 4x4 bmm), rows 0-2
    for (int i = threadIdx.x; i < nv * 12; i += blockDim.x) {
        const int v = i / 12, rc = i - v * 12, r = rc >> 2, c = rc & 3;
        const float *K = cams + v * 34 + 2, *E = cams + v * 34 + 18;
        // torch's 4x4 @ 4x4 bmm rounds every product and every sum (no FMA contraction), k ascending -- checked
        // bit for bit against the reference run on the CPU; the 4x4 @ 4xN product below IS an FMA chain
        float t = __fmul_rn(K[r * 4 + 0], E[0 * 4 + c]);
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 1], E[1 * 4 + c]));
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 2], E[2 * 4 + c]));
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 3], E[3 * 4 + c]));
        sP[i] = t;
    }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * kRgThreads + threadIdx.x;
    const bool live = p < n_pts;
    const float h = cams[0], w = cams[1];                                  // img_shape of the source views
    float X = 0.f, Y = 0.f, Z = 0.f;
    if (live) { X = pts[p * 3]; Y = pts[p * 3 + 1]; Z = pts[p * 3 + 2]; }
    const float wm1 = __fsub_rn(w, 1.0f), hm1 = __fsub_rn(h, 1.0f);
    int cnt = 0;
    for (int v = 0; v < nv; ++v) {
        const float *P = sP + v * 12;
        const float q0 = chain4(P, X, Y, Z), q1 = chain4(P + 4, X, Y, Z), q2 = chain4(P + 8, X, Y, Z);
        const float zc = fmaxf(q2, 1e-8f);
        float px = __fdiv_rn(q0, zc), py = __fdiv_rn(q1, zc);
        px = fminf(fmaxf(px, -1e6f), 1e6f);
        py = fminf(fmaxf(py, -1e6f), 1e6f);
        const bool front = q2 > 0.0f;
        const bool inb = (px <= wm1) && (px >= 0.0f) && (py <= hm1) && (py >= 0.0f);
        const bool m = inb && front;
        cnt += m ? 1 : 0;
        // normalize (projection.py:37-40): 2 * pix / [w - 1, h - 1] - 1
        const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, px), wm1), 1.0f);
        const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, py), hm1), 1.0f);
        ViewSample si = make_sample(gx, gy, hi, wi);
        ViewSample sf = make_sample(gx, gy, hf, wf);
        si.inb |= m ? 0x100u : 0u;
        sI[(size_t)v * kRgThreads + threadIdx.x] = si;
        sF[(size_t)v * kRgThreads + threadIdx.x] = sf;
        if (live && view_mask != nullptr) view_mask[p * nv + v] = m ? 1 : 0;
        if (live && front_out != nullptr) front_out[(int64_t)v * n_pts + p] = front ? 1 : 0;
        if (live && pix_out != nullptr) {
            pix_out[((int64_t)v * n_pts + p) * 2] = px;
            pix_out[((int64_t)v * n_pts + p) * 2 + 1] = py;
        }
    }
    if (!live) return;
    const float denom = __fadd_rn((float)cnt, 1e-8f);
    const int ct = 3 + d;
    float *row = glob + p * (int64_t)(2 * ct);
    if (pixel_mask != nullptr) pixel_mask[p] = cnt > 1 ? 1 : 0;

    auto finish = [&](int ch, float sm, float sa1, float sa2) {
        const float mean = sm / denom;                                     // sum f * (mask / (c + 1e-8))
        float ssd = fmaf(-2.0f * mean, sa1, sa2);                          // sum over ALL views of (f - mean)^2
        ssd = fmaxf(fmaf((float)nv * mean, mean, ssd), 0.0f);
        row[ch] = mean;
        row[ct + ch] = expf(-(ssd / denom));
    };
    {   // image channels
        float sm[3] = {0.f, 0.f, 0.f}, sa1[3] = {0.f, 0.f, 0.f}, sa2[3] = {0.f, 0.f, 0.f};
        for (int v = 0; v < nv; ++v) {
            const ViewSample s = sI[(size_t)v * kRgThreads + threadIdx.x];
            const bool m = (s.inb & 0x100u) != 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float f = bilinear<float>(img + v * i_sv + k * i_sc, s, wi);
                if (view_feat != nullptr) view_feat[(p * nv + v) * (int64_t)ct + k] = f;
                sm[k] += m ? f : 0.0f;
                sa1[k] += f;
                sa2[k] = fmaf(f, f, sa2[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) finish(k, sm[k], sa1[k], sa2[k]);
    }
    for (int c0 = 0; c0 < d; c0 += kRgChunk) {
        float sm[kRgChunk], sa1[kRgChunk], sa2[kRgChunk];
#pragma unroll
        for (int k = 0; k < kRgChunk; ++k) { sm[k] = 0.f; sa1[k] = 0.f; sa2[k] = 0.f; }
        for (int v = 0; v < nv; ++v) {
            const ViewSample s = sF[(size_t)v * kRgThreads + threadIdx.x];
            const bool m = (sI[(size_t)v * kRgThreads + threadIdx.x].inb & 0x100u) != 0;
            if (s.inb == 0 && view_feat == nullptr) continue;              // all four corners outside: f = 0
            const T *plane0 = feat + v * f_sv + (int64_t)c0 * f_sc;
#pragma unroll
            for (int k = 0; k < kRgChunk; ++k) {
                if (c0 + k < d) {
                    const float f = bilinear<T>(plane0 + k * f_sc, s, wf);
                    if (view_feat != nullptr) view_feat[(p * nv + v) * (int64_t)ct + 3 + c0 + k] = f;
                    sm[k] += m ? f : 0.0f;
                    sa1[k] += f;
                    sa2[k] = fmaf(f, f, sa2[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kRgChunk; ++k)
            if (c0 + k < d) finish(3 + c0 + k, sm[k], sa1[k], sa2[k]);
    }
}

// Channels-last build of the same statistics (the product path): one WARP per ray sample.
//
//   phase 1, lane = view: project the sample (IEEE divisions: the masks are bit-exact), build the two bilinear
//            samples (image, feature map) and their four weights, and append the views with at least one corner in
//            bounds (~35 % of them) to the warp's list in shared memory -- ascending view order, ballot + popc.
//   phase 2, quarter-warp = view: the four quarters walk the list four entries at a time.  The 8 lanes of a quarter
//            fetch the 32 mapped channels of a corner as one coalesced 128-byte row [pixel][D] (4 channels per
//            lane) and one image channel each; corner addresses are clamped into the map, so all eight loads of an
//            entry are unconditional and in flight together, and a corner outside the map is left out by the
//            predicate on its FMA (zeros padding of grid_sample, corner order nw, ne, sw, se like ATen).
//   phase 3: the quarters are combined with shuffles and every lane finishes ONE mapped channel (mean, variance
//            over ALL views, exp(-var)); three lanes finish the image channels.
//
// A warp walks the samples with a grid stride, so the eight warps of a CTA sit on eight consecutive samples of a
// ray (their corners share L1 lines) and the uneven number of visible views per sample averages out.
//
// Measured on B200 at 2048 rays x 64 samples x 50 views (profiles/r02_render_gather.md): 256 us, against 343 us for
// the round-1 loop (predicated loads issued two at a time, a find-first-set selection per step).  Tried and dropped:
// a 64-register cap for a fourth CTA per SM (362 us: the eight loads no longer stay in flight together), splitting
// phase 1 into a cheap candidate pass and one compacted round of sample building (20 % fewer instructions, 267 us:
// the kernel is bound by the latency of a warp's serial phases, not by issue slots), prefetch.global.L1 of the
// corner lines from phase 1 (311 us).
constexpr int kRcWarps = 8, kRcList = 64;

struct ViewSampleG {         // like ViewSample, for arbitrary strides
    int32_t x0, y0;
    float fx, fy;
    uint32_t inb;
};

// ATen grid_sampler_2d unnormalisation ((g + 1) / 2 * (size - 1); the halving is exact) + corner bookkeeping
__device__ __forceinline__ ViewSampleG make_sample_g(float gx, float gy, int hs, int ws) {
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(ws - 1));
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(hs - 1));
    const float x0f = floorf(ix), y0f = floorf(iy);
    ViewSampleG s;
    s.fx = ix - x0f;
    s.fy = iy - y0f;
    s.x0 = (int)fminf(fmaxf(x0f, -4.0f), (float)(ws + 2));
    s.y0 = (int)fminf(fmaxf(y0f, -4.0f), (float)(hs + 2));
    const bool xin0 = s.x0 >= 0 && s.x0 < ws, xin1 = s.x0 + 1 >= 0 && s.x0 + 1 < ws;
    const bool yin0 = s.y0 >= 0 && s.y0 < hs, yin1 = s.y0 + 1 >= 0 && s.y0 + 1 < hs;
    s.inb = (xin0 && yin0 ? 1u : 0u) | (xin1 && yin0 ? 2u : 0u) | (xin0 && yin1 ? 4u : 0u) | (xin1 && yin1 ? 8u : 0u);
    if (!(ix == ix) || !(iy == iy)) s.inb = 0;
    return s;
}

// One listed (sample, view): north-west corners clamped into the maps (so every corner address is loadable),
// whether the east / south neighbours are one pixel further (else they alias the clamped corner and their
// predicate is off), and the bilinear weights.  48 bytes = three broadcast LDS.128 per quarter-warp.
struct __align__(16) ViewEntry {
    int32_t offI, offF;      // element offsets of the clamped north-west corners (view base included)
    uint32_t bits;           // image corners | feature corners << 4 | view mask << 8 | dxI, dyI, dxF, dyF << 9
    float maskf;             // 1.0f when the view sees the sample (in bounds and in front), else 0.0f
    float wI[4], wF[4];      // nw, ne, sw, se
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <typename T> __device__ __forceinline__ float4 load4(const T *p);
template <> __device__ __forceinline__ float4 load4<float>(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
}

template <typename T, bool kFeat>
__global__ void __launch_bounds__(kRcWarps * 32, 3)
k_render_gather_stats_cl(const float *__restrict__ pts, int64_t n_pts, const float *__restrict__ cams, int nv,
                         const float *__restrict__ img, int i_sv, int i_sc, int i_sy, int i_sx, int hi, int wi,
                         const T *__restrict__ feat, int f_sv, int f_sy, int f_sx, int d, int hf, int wf,
                         float *__restrict__ glob, uint8_t *__restrict__ view_mask, uint8_t *__restrict__ pixel_mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ViewEntry *sL = reinterpret_cast<ViewEntry *>(smem_raw) + (threadIdx.x >> 5) * kRcList;  // [warps][kRcList]
    float *sP = reinterpret_cast<float *>(smem_raw + sizeof(ViewEntry) * kRcList * kRcWarps); // [nv][12] rows 0-2 of K @ E
    for (int i = threadIdx.x; i < nv * 12; i += blockDim.x) {
        const int v = i / 12, rc = i - v * 12, r = rc >> 2, c = rc & 3;
        const float *K = cams + v * 34 + 2, *E = cams + v * 34 + 18;
        float t = __fmul_rn(K[r * 4 + 0], E[0 * 4 + c]);                   // see k_render_gather_stats
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 1], E[1 * 4 + c]));
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 2], E[2 * 4 + c]));
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 3], E[3 * 4 + c]));
        sP[i] = t;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, quarter = lane >> 3, l8 = lane & 7;
    const float h = cams[0], w = cams[1];
    const float wm1 = __fsub_rn(w, 1.0f), hm1 = __fsub_rn(h, 1.0f);
    const int ct = 3 + d;
    const unsigned full = 0xffffffffu, below = (1u << lane) - 1u;
    const bool feat_lane = 4 * l8 < d, img_lane = l8 < 3;
    const float *img_l = img + (img_lane ? l8 * i_sc : 0);                 // lanes 3-7 of a quarter repeat channel 0
    const T *feat_l = feat + (feat_lane ? 4 * l8 : 0);
    const int64_t stride = (int64_t)gridDim.x * kRcWarps;
    for (int64_t p = (int64_t)blockIdx.x * kRcWarps + warp; p < n_pts; p += stride) {
        const float X = __ldg(pts + p * 3), Y = __ldg(pts + p * 3 + 1), Z = __ldg(pts + p * 3 + 2);
        int cnt = 0, n_list = 0;
        float smF[4] = {0.f, 0.f, 0.f, 0.f}, s1F[4] = {0.f, 0.f, 0.f, 0.f}, s2F[4] = {0.f, 0.f, 0.f, 0.f};
        float smI = 0.f, s1I = 0.f, s2I = 0.f;
        for (int v0 = 0; v0 < nv; v0 += 32) {
            // ---- phase 1: lane = view
            const int v = v0 + lane;
            bool m = false, listed = false;
            ViewEntry e;
            if (v < nv) {
                const float *P = sP + v * 12;
                const float q0 = chain4(P, X, Y, Z), q1 = chain4(P + 4, X, Y, Z), q2 = chain4(P + 8, X, Y, Z);
                const float zc = fmaxf(q2, 1e-8f);
                float px = __fdiv_rn(q0, zc), py = __fdiv_rn(q1, zc);      // IEEE: the masks are bit-exact
                px = fminf(fmaxf(px, -1e6f), 1e6f);
                py = fminf(fmaxf(py, -1e6f), 1e6f);
                const bool front = q2 > 0.0f;
                const bool inb = (px <= wm1) && (px >= 0.0f) && (py <= hm1) && (py >= 0.0f);
                m = inb && front;
                // normalize (projection.py:37-40): 2 * pix / [w - 1, h - 1] - 1
                const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, px), wm1), 1.0f);
                const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, py), hm1), 1.0f);
                const ViewSampleG si = make_sample_g(gx, gy, hi, wi);
                ViewSampleG sf = make_sample_g(gx, gy, hf, wf);
                if (!kFeat) sf.inb = 0u;
                listed = (si.inb | sf.inb) != 0u;
                const int xi = clampi(si.x0, 0, wi - 1), yi = clampi(si.y0, 0, hi - 1);
                const int xf = clampi(sf.x0, 0, wf - 1), yf = clampi(sf.y0, 0, hf - 1);
                e.offI = v * i_sv + yi * i_sy + xi * i_sx;
                e.offF = v * f_sv + yf * f_sy + xf * f_sx;
                e.bits = si.inb | (sf.inb << 4) | (m ? 0x100u : 0u) |
                         ((si.x0 >= 0 && si.x0 + 1 < wi) ? 0x200u : 0u) | ((si.y0 >= 0 && si.y0 + 1 < hi) ? 0x400u : 0u) |
                         ((sf.x0 >= 0 && sf.x0 + 1 < wf) ? 0x800u : 0u) | ((sf.y0 >= 0 && sf.y0 + 1 < hf) ? 0x1000u : 0u);
                e.maskf = m ? 1.0f : 0.0f;
                e.wI[0] = (1.0f - si.fx) * (1.0f - si.fy); e.wI[1] = si.fx * (1.0f - si.fy);
                e.wI[2] = (1.0f - si.fx) * si.fy;          e.wI[3] = si.fx * si.fy;
                e.wF[0] = (1.0f - sf.fx) * (1.0f - sf.fy); e.wF[1] = sf.fx * (1.0f - sf.fy);
                e.wF[2] = (1.0f - sf.fx) * sf.fy;          e.wF[3] = sf.fx * sf.fy;
                if (view_mask != nullptr) view_mask[p * nv + v] = m ? 1 : 0;
            }
            cnt += __popc(__ballot_sync(full, m));
            const unsigned act = __ballot_sync(full, listed);
            if (listed) sL[n_list + __popc(act & below)] = e;
            n_list += __popc(act);
            if (n_list <= kRcList - 32 && v0 + 32 < nv) continue;          // room for another round of views
            __syncwarp();
            // ---- phase 2: quarter-warp = listed view
            for (int i = quarter; i < n_list; i += 4) {
                const int4 a4 = *reinterpret_cast<const int4 *>(&sL[i]);               // offI, offF, bits, maskf
                const float4 wi4 = *reinterpret_cast<const float4 *>(sL[i].wI);
                const uint32_t b = (uint32_t)a4.z;
                const float mvf = __int_as_float(a4.w);
                const float *qi = img_l + a4.x;
                const int dxi = (b & 0x200u) ? i_sx : 0, dyi = (b & 0x400u) ? i_sy : 0;
                const float i0 = __ldg(qi), i1 = __ldg(qi + dxi), i2 = __ldg(qi + dyi), i3 = __ldg(qi + dyi + dxi);
                float4 c0, c1, c2, c3, wf4;
                if (kFeat) {
                    const T *qf = feat_l + a4.y;
                    const int dxf = (b & 0x800u) ? f_sx : 0, dyf = (b & 0x1000u) ? f_sy : 0;
                    c0 = load4<T>(qf); c1 = load4<T>(qf + dxf); c2 = load4<T>(qf + dyf); c3 = load4<T>(qf + dyf + dxf);
                    wf4 = *reinterpret_cast<const float4 *>(sL[i].wF);
                }
                {
                    float f = 0.0f;
                    if (b & 1u) f = fmaf(i0, wi4.x, f);
                    if (b & 2u) f = fmaf(i1, wi4.y, f);
                    if (b & 4u) f = fmaf(i2, wi4.z, f);
                    if (b & 8u) f = fmaf(i3, wi4.w, f);
                    smI = fmaf(f, mvf, smI);                                // sum f * mask, like the reference's product
                    s1I += f;
                    s2I = fmaf(f, f, s2I);
                }
                if (kFeat) {
                    const float a0[4] = {c0.x, c0.y, c0.z, c0.w}, a1[4] = {c1.x, c1.y, c1.z, c1.w};
                    const float a2[4] = {c2.x, c2.y, c2.z, c2.w}, a3[4] = {c3.x, c3.y, c3.z, c3.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float f = 0.0f;
                        if (b & 0x10u) f = fmaf(a0[k], wf4.x, f);
                        if (b & 0x20u) f = fmaf(a1[k], wf4.y, f);
                        if (b & 0x40u) f = fmaf(a2[k], wf4.z, f);
                        if (b & 0x80u) f = fmaf(a3[k], wf4.w, f);
                        smF[k] = fmaf(f, mvf, smF[k]);
                        s1F[k] += f;
                        s2F[k] = fmaf(f, f, s2F[k]);
                    }
                }
            }
            n_list = 0;
            __syncwarp();                                                  // the list is free for the next round
        }
        // ---- phase 3: combine the quarters (lanes l8, l8 + 8, l8 + 16, l8 + 24 hold the same channels)
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            if (kFeat) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    smF[k] += __shfl_xor_sync(full, smF[k], o);
                    s1F[k] += __shfl_xor_sync(full, s1F[k], o);
                    s2F[k] += __shfl_xor_sync(full, s2F[k], o);
                }
            }
            smI += __shfl_xor_sync(full, smI, o);
            s1I += __shfl_xor_sync(full, s1I, o);
            s2I += __shfl_xor_sync(full, s2I, o);
        }
        const float denom = __fadd_rn((float)cnt, 1e-8f);
        float *row = glob + p * (int64_t)(2 * ct);
        if (lane == 0 && pixel_mask != nullptr) pixel_mask[p] = cnt > 1 ? 1 : 0;
        auto finish = [&](int ch, float sm, float s1, float s2) {
            const float mean = sm / denom;
            float ssd = fmaf(-2.0f * mean, s1, s2);                        // sum over ALL views of (f - mean)^2
            ssd = fmaxf(fmaf((float)nv * mean, mean, ssd), 0.0f);
            row[ch] = mean;
            row[ct + ch] = expf(-(ssd / denom));
        };
        if (kFeat && feat_lane) {                                          // quarter q finishes channel 4 * l8 + q
            const float sm = quarter == 0 ? smF[0] : quarter == 1 ? smF[1] : quarter == 2 ? smF[2] : smF[3];
            const float s1 = quarter == 0 ? s1F[0] : quarter == 1 ? s1F[1] : quarter == 2 ? s1F[2] : s1F[3];
            const float s2 = quarter == 0 ? s2F[0] : quarter == 1 ? s2F[1] : quarter == 2 ? s2F[2] : s2F[3];
            finish(3 + 4 * l8 + quarter, sm, s1, s2);
        }
        if (quarter == 0 && img_lane) finish(l8, smI, s1I, s2I);
    }
}

// Backward of the same statistics with respect to the mapped feature maps (row N1; autograd of projection.py:91-151 +
// render_ray.py:71-93 for `featmaps`).  Per sample and channel, with denom = count + 1e-8, mean and e = exp(-var) from the
// forward's output row and g_mean, g_e the incoming gradients:  g_var = -e * g_e and every listed view receives
//   g_f = m_v * (g_mean / denom - 2 g_var (S1 - nv * mean) / denom^2) + 2 g_var (f_v - mean) / denom
// (m_v = the view mask, S1 = the sum of f over ALL views), which the bilinear weights spread over the four corners
// (zeros padding: corners outside the map take nothing).  Same phases as the forward kernel: list the visible views, walk
// the list once for S1, once more to recompute f_v and scatter -- 16-byte vector reductions (red.global.add.v4.f32) into
// the channels-last gradient.  The images get no gradient (they are input data).
__device__ __forceinline__ void red_add_f4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kRcWarps * 32, 3)
k_render_gather_stats_bwd(const float *__restrict__ pts, int64_t n_pts, const float *__restrict__ cams, int nv, int list_cap,
                          int hi, int wi, const float *__restrict__ feat, int f_sv, int f_sy, int f_sx, int d, int hf, int wf,
                          const float *__restrict__ glob, const float *__restrict__ g_glob, float *__restrict__ g_feat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ViewEntry *sL = reinterpret_cast<ViewEntry *>(smem_raw) + (size_t)(threadIdx.x >> 5) * list_cap;  // [warps][list_cap >= nv]
    float *sP = reinterpret_cast<float *>(smem_raw + sizeof(ViewEntry) * (size_t)list_cap * kRcWarps);
    for (int i = threadIdx.x; i < nv * 12; i += blockDim.x) {
        const int v = i / 12, rc = i - v * 12, r = rc >> 2, c = rc & 3;
        const float *K = cams + v * 34 + 2, *E = cams + v * 34 + 18;
        float t = __fmul_rn(K[r * 4 + 0], E[0 * 4 + c]);                   // see k_render_gather_stats
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 1], E[1 * 4 + c]));
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 2], E[2 * 4 + c]));
        t = __fadd_rn(t, __fmul_rn(K[r * 4 + 3], E[3 * 4 + c]));
        sP[i] = t;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, quarter = lane >> 3, l8 = lane & 7;
    const float h = cams[0], w = cams[1];
    const float wm1 = __fsub_rn(w, 1.0f), hm1 = __fsub_rn(h, 1.0f);
    const int ct = 3 + d;
    const unsigned full = 0xffffffffu, below = (1u << lane) - 1u;
    const bool feat_lane = 4 * l8 < d;
    const int c_off = feat_lane ? 4 * l8 : 0;
    const int64_t stride = (int64_t)gridDim.x * kRcWarps;
    for (int64_t p = (int64_t)blockIdx.x * kRcWarps + warp; p < n_pts; p += stride) {
        const float X = __ldg(pts + p * 3), Y = __ldg(pts + p * 3 + 1), Z = __ldg(pts + p * 3 + 2);
        int cnt = 0, n_list = 0;
        for (int v0 = 0; v0 < nv; v0 += 32) {                              // phase 1: the forward's, feature map only
            const int v = v0 + lane;
            bool m = false, listed = false;
            ViewEntry e;
            if (v < nv) {
                const float *P = sP + v * 12;
                const float q0 = chain4(P, X, Y, Z), q1 = chain4(P + 4, X, Y, Z), q2 = chain4(P + 8, X, Y, Z);
                const float zc = fmaxf(q2, 1e-8f);
                float px = __fdiv_rn(q0, zc), py = __fdiv_rn(q1, zc);
                px = fminf(fmaxf(px, -1e6f), 1e6f);
                py = fminf(fmaxf(py, -1e6f), 1e6f);
                const bool front = q2 > 0.0f;
                const bool inb = (px <= wm1) && (px >= 0.0f) && (py <= hm1) && (py >= 0.0f);
                m = inb && front;
                const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, px), wm1), 1.0f);
                const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, py), hm1), 1.0f);
                const ViewSampleG sf = make_sample_g(gx, gy, hf, wf);
                listed = sf.inb != 0u;
                const int xf = clampi(sf.x0, 0, wf - 1), yf = clampi(sf.y0, 0, hf - 1);
                e.offI = 0;
                e.offF = v * f_sv + yf * f_sy + xf * f_sx;
                e.bits = (sf.inb << 4) | ((sf.x0 >= 0 && sf.x0 + 1 < wf) ? 0x800u : 0u) | ((sf.y0 >= 0 && sf.y0 + 1 < hf) ? 0x1000u : 0u);
                e.maskf = m ? 1.0f : 0.0f;
                e.wI[0] = e.wI[1] = e.wI[2] = e.wI[3] = 0.0f;
                e.wF[0] = (1.0f - sf.fx) * (1.0f - sf.fy); e.wF[1] = sf.fx * (1.0f - sf.fy);
                e.wF[2] = (1.0f - sf.fx) * sf.fy;          e.wF[3] = sf.fx * sf.fy;
            }
            cnt += __popc(__ballot_sync(full, m));
            const unsigned act = __ballot_sync(full, listed);
            if (listed) sL[n_list + __popc(act & below)] = e;
            n_list += __popc(act);
        }
        __syncwarp();
        // the bilinear value of a listed view for this lane's four channels
        auto value = [&](int i, float (&f)[4], uint32_t &b, float4 &w4, const float *&q, int &dxf, int &dyf, float &mvf) {
            const int4 a4 = *reinterpret_cast<const int4 *>(&sL[i]);
            w4 = *reinterpret_cast<const float4 *>(sL[i].wF);
            b = (uint32_t)a4.z;
            mvf = __int_as_float(a4.w);
            q = feat + c_off + a4.y;
            dxf = (b & 0x800u) ? f_sx : 0;
            dyf = (b & 0x1000u) ? f_sy : 0;
            const float4 c0 = load4<float>(q), c1 = load4<float>(q + dxf), c2 = load4<float>(q + dyf), c3 = load4<float>(q + dyf + dxf);
            const float a0[4] = {c0.x, c0.y, c0.z, c0.w}, a1[4] = {c1.x, c1.y, c1.z, c1.w};
            const float a2[4] = {c2.x, c2.y, c2.z, c2.w}, a3[4] = {c3.x, c3.y, c3.z, c3.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t = 0.0f;
                if (b & 0x10u) t = fmaf(a0[k], w4.x, t);
                if (b & 0x20u) t = fmaf(a1[k], w4.y, t);
                if (b & 0x40u) t = fmaf(a2[k], w4.z, t);
                if (b & 0x80u) t = fmaf(a3[k], w4.w, t);
                f[k] = t;
            }
        };
        float s1[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = quarter; i < n_list; i += 4) {                        // pass A: sum over all views
            float f[4], mvf;
            uint32_t b;
            float4 w4;
            const float *q;
            int dxf, dyf;
            value(i, f, b, w4, q, dxf, dyf, mvf);
#pragma unroll
            for (int k = 0; k < 4; ++k) s1[k] += f[k];
        }
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1)
#pragma unroll
            for (int k = 0; k < 4; ++k) s1[k] += __shfl_xor_sync(full, s1[k], o);
        const float denom = __fadd_rn((float)cnt, 1e-8f);
        const float *row = glob + p * (int64_t)(2 * ct), *grow = g_glob + p * (int64_t)(2 * ct);
        float cm[4], cb[4], mean[4];                                       // g_f = m_v * cm + cb * (f - mean)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ch = 3 + c_off + k;
            mean[k] = feat_lane ? row[ch] : 0.0f;
            const float gvar = feat_lane ? -row[ct + ch] * grow[ct + ch] : 0.0f;
            const float gm = feat_lane ? grow[ch] : 0.0f;
            cb[k] = 2.0f * gvar / denom;
            cm[k] = gm / denom - cb[k] * (s1[k] - (float)nv * mean[k]) / denom;
        }
        if (feat_lane) {
            for (int i = quarter; i < n_list; i += 4) {                    // pass B: recompute, scatter
                float f[4], mvf;
                uint32_t b;
                float4 w4;
                const float *q;
                int dxf, dyf;
                value(i, f, b, w4, q, dxf, dyf, mvf);
                float g[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) g[k] = fmaf(mvf, cm[k], cb[k] * (f[k] - mean[k]));
                float *gq = g_feat + (q - feat);
                if (b & 0x10u) red_add_f4(gq, g[0] * w4.x, g[1] * w4.x, g[2] * w4.x, g[3] * w4.x);
                if (b & 0x20u) red_add_f4(gq + dxf, g[0] * w4.y, g[1] * w4.y, g[2] * w4.y, g[3] * w4.y);
                if (b & 0x40u) red_add_f4(gq + dyf, g[0] * w4.z, g[1] * w4.z, g[2] * w4.z, g[3] * w4.z);
                if (b & 0x80u) red_add_f4(gq + dyf + dxf, g[0] * w4.w, g[1] * w4.w, g[2] * w4.w, g[3] * w4.w);
            }
        }
        __syncwarp();                                                      // the list is free for the next sample
    }
}

// -------------------------------------------------------------------------------------------------
// R7
// -------------------------------------------------------------------------------------------------
// One WARP per ray: lanes over samples (coalesced loads), the transmittance cumprod(1 - alpha + 1e-10) as a shuffle
// scan per 32 samples with a running carry, the weighted sums as warp reductions.  (One thread per ray walking its
// 64 samples left 2048 threads on the whole GPU: 49 us of pure latency.)
__global__ void __launch_bounds__(128)
k_composite(const float *__restrict__ rgb, const float *__restrict__ sigma, const float *__restrict__ z,
            const uint8_t *__restrict__ pmask, int64_t n_rays, int ns, const float *__restrict__ z_bounds,
            int white_bkgd, float *__restrict__ out_rgb, float *__restrict__ depth,
            float *__restrict__ weights, float *__restrict__ alpha_o, float *__restrict__ trans_o,
            uint8_t *__restrict__ ray_mask) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rays) return;                                               // warp-uniform
    const unsigned full = 0xffffffffu;
    float carry = 1.0f, c0 = 0.f, c1 = 0.f, c2 = 0.f, wsum = 0.f, dsum = 0.f;
    int msum = 0;
    for (int s0 = 0; s0 < ns; s0 += 32) {
        const int s = s0 + lane;
        const bool ok = s < ns;
        const int64_t i = r * ns + (ok ? s : 0);
        const float a = ok ? __fsub_rn(1.0f, expf(-sigma[i])) : 0.0f;      // sigma2alpha, no interval term
        const float f = ok ? __fadd_rn(__fsub_rn(1.0f, a), 1e-10f) : 1.0f; // factor of cumprod(1 - alpha + 1e-10)
        float incl = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl *= t;
        }
        float excl = __shfl_up_sync(full, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float T = carry * excl;                                      // transmittance before this sample
        carry *= __shfl_sync(full, incl, 31);
        if (ok) {
            const float wgt = __fmul_rn(a, T);
            if (alpha_o != nullptr) alpha_o[i] = a;
            if (trans_o != nullptr) trans_o[i] = T;
            if (weights != nullptr) weights[i] = wgt;
            c0 = fmaf(wgt, rgb[i * 3], c0);
            c1 = fmaf(wgt, rgb[i * 3 + 1], c1);
            c2 = fmaf(wgt, rgb[i * 3 + 2], c2);
            wsum += wgt;
            dsum = fmaf(wgt, z[i], dsum);
            if (pmask != nullptr) msum += pmask[i] ? 1 : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(full, c0, o);
        c1 += __shfl_xor_sync(full, c1, o);
        c2 += __shfl_xor_sync(full, c2, o);
        wsum += __shfl_xor_sync(full, wsum, o);
        dsum += __shfl_xor_sync(full, dsum, o);
        msum += __shfl_xor_sync(full, msum, o);
    }
    if (lane != 0) return;
    if (white_bkgd) {
        const float bg = 1.0f - wsum;
        c0 += bg; c1 += bg; c2 += bg;
    }
    out_rgb[r * 3] = c0;
    out_rgb[r * 3 + 1] = c1;
    out_rgb[r * 3 + 2] = c2;
    float dd = dsum / __fadd_rn(wsum, 1e-8f);
    depth[r] = fminf(fmaxf(dd, __ldg(z_bounds)), __ldg(z_bounds + 1));
    if (ray_mask != nullptr) ray_mask[r] = msum > 8 ? 1 : 0;
}

// -------------------------------------------------------------------------------------------------
// R8: ATen grid_sampler_3d, bilinear, border padding, align_corners = True.  The normalised x
// coordinate indexes the LAST volume axis (D2), y -> D1, z -> D0, applied literally like the reference.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float unnorm_border(float g, int size) {
    float c = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), (float)(size - 1));
    return fminf(fmaxf(c, 0.0f), (float)(size - 1));                       // clip_coordinates
}

__global__ void k_volume_sample(const float *__restrict__ vol, int c, int d0, int d1, int d2,
                                const float *__restrict__ pts, int64_t n_pts, float3 lo, float3 inv,
                                float *__restrict__ out, uint8_t *__restrict__ inside) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const float nx = __fsub_rn(__fmul_rn(__fsub_rn(pts[p * 3], lo.x), inv.x), 1.0f);
    const float ny = __fsub_rn(__fmul_rn(__fsub_rn(pts[p * 3 + 1], lo.y), inv.y), 1.0f);
    const float nz = __fsub_rn(__fmul_rn(__fsub_rn(pts[p * 3 + 2], lo.z), inv.z), 1.0f);
    if (inside != nullptr)
        inside[p] = (nx < 1.f && nx > -1.f && ny < 1.f && ny > -1.f && nz < 1.f && nz > -1.f) ? 1 : 0;
    const float ix = unnorm_border(nx, d2), iy = unnorm_border(ny, d1), iz = unnorm_border(nz, d0);
    const float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
    const int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
    const float tx = ix - x0f, ty = iy - y0f, tz = iz - z0f;
    const int64_t plane = (int64_t)d0 * d1 * d2;
    for (int ch = 0; ch < c; ++ch) {
        const float *v = vol + ch * plane;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
            const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
            const float wgt = (dx ? tx : 1.0f - tx) * (dy ? ty : 1.0f - ty) * (dz ? tz : 1.0f - tz);
            if (xx < d2 && yy < d1 && zz < d0) acc = fmaf(v[((int64_t)zz * d1 + yy) * d2 + xx], wgt, acc);
        }
        out[p * c + ch] = acc;
    }
}

}  // namespace nd

using namespace nd;

extern "C" {

int nd_sample_rays(const float *ray_o, const float *ray_d, int64_t n_rays, int n_samples, float near_depth,
                   float far_depth, const float *t_rand, float *pts, float *z_vals, void *stream) {
    ND_REQUIRE(ray_o && ray_d && pts && z_vals, ND_ERR_BAD_ARG, "nd_sample_rays: null pointer");
    ND_REQUIRE(n_rays >= 0 && n_samples >= 2, ND_ERR_BAD_SHAPE, "nd_sample_rays: need n_samples >= 2");
    ND_REQUIRE(near_depth > 0.f && far_depth > near_depth, ND_ERR_BAD_ARG, "nd_sample_rays: need 0 < near < far");
    if (n_rays == 0) return ND_OK;
    const int64_t total = n_rays * n_samples;
    k_sample_rays<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(ray_o, ray_d, n_rays, n_samples,
                                                                                    near_depth, far_depth, t_rand, pts,
                                                                                    z_vals);
    ND_CUDA_LAUNCH_CHECK("k_sample_rays");
    return ND_OK;
}

int nd_render_gather_stats(const float *pts, int64_t n_points, const float *cameras, int n_views,
                           const nd_maps *images, const nd_maps *featmaps, float *globalfeat, uint8_t *view_mask,
                           uint8_t *pixel_mask, float *pixel_locations, uint8_t *in_front, float *view_features, void *stream) {
    ND_REQUIRE(pts && cameras && images && featmaps && images->data && (featmaps->data || featmaps->channels == 0) &&
                   globalfeat,
               ND_ERR_BAD_ARG, "nd_render_gather_stats: null pointer");
    ND_REQUIRE(n_views > 0 && images->n_views == n_views && featmaps->n_views == n_views, ND_ERR_BAD_SHAPE,
               "nd_render_gather_stats: view counts differ");
    ND_REQUIRE(images->dtype == ND_F32 && images->channels == 3, ND_ERR_BAD_SHAPE,
               "nd_render_gather_stats: images must be f32 [nv,3,H,W]");
    ND_REQUIRE(n_points >= 0, ND_ERR_BAD_SHAPE, "nd_render_gather_stats: negative point count");
    if (n_points == 0) return ND_OK;
    // product path: channels-last feature maps ([nv][h][w][D], D <= 32) and nothing materialised per view
    if ((featmaps->channels == 0 || (featmaps->stride_c == 1 && featmaps->channels <= 32 && featmaps->channels % 4 == 0 &&
                                     featmaps->stride_x % 4 == 0 && featmaps->stride_y % 4 == 0 && featmaps->stride_v % 4 == 0 &&
                                     (reinterpret_cast<uintptr_t>(featmaps->data) & 15) == 0)) &&
        pixel_locations == nullptr &&
        in_front == nullptr && view_features == nullptr) {
        const size_t sm = (size_t)n_views * 12 * sizeof(float) + sizeof(ViewEntry) * kRcList * kRcWarps;
        ND_REQUIRE(sm <= 48 * 1024, ND_ERR_BAD_SHAPE, "nd_render_gather_stats: too many views (%d)", n_views);
        const int64_t span_i = (int64_t)n_views * images->stride_v, span_f = (int64_t)n_views * featmaps->stride_v;
        ND_REQUIRE(span_i < (1ll << 31) && span_f < (1ll << 31) && images->stride_v >= 0 && featmaps->stride_v >= 0,
                   ND_ERR_BAD_SHAPE, "nd_render_gather_stats: source stacks beyond 2^31 elements");
        // grid-stride over the samples: as many CTAs as stay resident (3 per SM at 80 registers), no more than needed
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const unsigned g = (unsigned)std::min<int64_t>(ceil_div(n_points, (int64_t)kRcWarps), (int64_t)sms * 3);
        cudaStream_t s0 = (cudaStream_t)stream;
#define ND_RC_LAUNCH(T, FEAT)                                                                                          \
    k_render_gather_stats_cl<T, FEAT><<<g, kRcWarps * 32, sm, s0>>>(                                                   \
        pts, n_points, cameras, n_views, (const float *)images->data, (int)images->stride_v, (int)images->stride_c,    \
        (int)images->stride_y, (int)images->stride_x, images->height, images->width, (const T *)featmaps->data,        \
        (int)featmaps->stride_v, (int)featmaps->stride_y, (int)featmaps->stride_x, featmaps->channels,                 \
        featmaps->channels ? featmaps->height : 1, featmaps->channels ? featmaps->width : 1, globalfeat, view_mask,    \
        pixel_mask)
        if (featmaps->channels == 0)
            ND_RC_LAUNCH(float, false);
        else if (featmaps->dtype == ND_F32)
            ND_RC_LAUNCH(float, true);
        else
            ND_RC_LAUNCH(__nv_bfloat16, true);
#undef ND_RC_LAUNCH
        ND_CUDA_LAUNCH_CHECK("k_render_gather_stats_cl");
        return ND_OK;
    }
    ND_REQUIRE(images->stride_x == 1 && images->stride_y == images->width && featmaps->stride_x == 1 &&
                   featmaps->stride_y == featmaps->width,
               ND_ERR_BAD_SHAPE, "nd_render_gather_stats: the materialising path needs contiguous NCHW planes");
    const size_t smem = (size_t)n_views * 12 * sizeof(float) + (size_t)2 * n_views * kRgThreads * sizeof(ViewSample);
    ND_REQUIRE(smem <= 220 * 1024, ND_ERR_BAD_SHAPE, "nd_render_gather_stats: too many views (%d)", n_views);
    const unsigned grid = (unsigned)ceil_div(n_points, kRgThreads);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if (featmaps->dtype == ND_F32) {
        auto kern = k_render_gather_stats<float>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            kern<<<grid, kRgThreads, smem, st>>>(pts, n_points, cameras, n_views, (const float *)images->data,
                                                 images->stride_v, images->stride_c, images->height, images->width,
                                                 (const float *)featmaps->data, featmaps->stride_v, featmaps->stride_c,
                                                 featmaps->channels, featmaps->height, featmaps->width, globalfeat,
                                                 view_mask, pixel_mask, pixel_locations, in_front, view_features);
    } else {
        auto kern = k_render_gather_stats<__nv_bfloat16>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            kern<<<grid, kRgThreads, smem, st>>>(pts, n_points, cameras, n_views, (const float *)images->data,
                                                 images->stride_v, images->stride_c, images->height, images->width,
                                                 (const __nv_bfloat16 *)featmaps->data, featmaps->stride_v,
                                                 featmaps->stride_c, featmaps->channels, featmaps->height,
                                                 featmaps->width, globalfeat, view_mask, pixel_mask, pixel_locations, in_front, view_features);
    }
    if (e != cudaSuccess) {
        set_error("nd_render_gather_stats: %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    ND_CUDA_LAUNCH_CHECK("k_render_gather_stats");
    return ND_OK;
}

int nd_render_gather_stats_bwd(const float *pts, int64_t n_points, const float *cameras, int n_views, int image_height,
                               int image_width, const nd_maps *featmaps, const float *globalfeat, const float *grad_globalfeat,
                               float *grad_featmaps, void *stream) {
    ND_REQUIRE(pts && cameras && featmaps && featmaps->data && globalfeat && grad_globalfeat && grad_featmaps, ND_ERR_BAD_ARG,
               "nd_render_gather_stats_bwd: null pointer");
    ND_REQUIRE(n_views > 0 && featmaps->n_views == n_views && n_points >= 0 && image_height > 0 && image_width > 0,
               ND_ERR_BAD_SHAPE, "nd_render_gather_stats_bwd: bad shape");
    ND_REQUIRE(featmaps->dtype == ND_F32 && featmaps->stride_c == 1 && featmaps->channels > 0 && featmaps->channels <= 32 &&
                   featmaps->channels % 4 == 0 && featmaps->stride_x == featmaps->channels &&
                   featmaps->stride_y == (int64_t)featmaps->width * featmaps->channels &&
                   featmaps->stride_v == (int64_t)featmaps->height * featmaps->width * featmaps->channels &&
                   (reinterpret_cast<uintptr_t>(featmaps->data) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad_featmaps) & 15) == 0,
               ND_ERR_BAD_SHAPE, "nd_render_gather_stats_bwd: contiguous channels-last f32 maps [nv][h][w][D], D <= 32, D %% 4 == 0");
    if (n_points == 0) return ND_OK;
    const int list_cap = (n_views + 31) & ~31;
    const size_t sm = (size_t)n_views * 12 * sizeof(float) + sizeof(ViewEntry) * (size_t)list_cap * kRcWarps;
    ND_REQUIRE(sm <= 200 * 1024, ND_ERR_BAD_SHAPE, "nd_render_gather_stats_bwd: too many views (%d)", n_views);
    ND_REQUIRE((int64_t)n_views * featmaps->stride_v < (1ll << 31), ND_ERR_BAD_SHAPE, "nd_render_gather_stats_bwd: maps beyond 2^31 elements");
    cudaError_t e = cudaFuncSetAttribute(k_render_gather_stats_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) {
        set_error("nd_render_gather_stats_bwd: %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned g = (unsigned)std::min<int64_t>(ceil_div(n_points, (int64_t)kRcWarps), (int64_t)sms * 3);
    k_render_gather_stats_bwd<<<g, kRcWarps * 32, sm, (cudaStream_t)stream>>>(
        pts, n_points, cameras, n_views, list_cap, image_height, image_width, (const float *)featmaps->data, (int)featmaps->stride_v,
        (int)featmaps->stride_y, (int)featmaps->stride_x, featmaps->channels, featmaps->height, featmaps->width, globalfeat,
        grad_globalfeat, grad_featmaps);
    ND_CUDA_LAUNCH_CHECK("k_render_gather_stats_bwd");
    return ND_OK;
}

int nd_composite(const float *rgb, const float *sigma, const float *z_vals, const uint8_t *pixel_mask, int64_t n_rays,
                 int n_samples, const float *z_bounds, int white_bkgd, float *out_rgb, float *out_depth,
                 float *weights, float *alpha, float *transparency, uint8_t *ray_mask, void *stream) {
    ND_REQUIRE(rgb && sigma && z_vals && z_bounds && out_rgb && out_depth, ND_ERR_BAD_ARG, "nd_composite: null pointer");
    ND_REQUIRE(n_rays >= 0 && n_samples >= 1, ND_ERR_BAD_SHAPE, "nd_composite: bad shape");
    if (n_rays == 0) return ND_OK;
    k_composite<<<(unsigned)ceil_div(n_rays, 4), 128, 0, (cudaStream_t)stream>>>(
        rgb, sigma, z_vals, pixel_mask, n_rays, n_samples, z_bounds, white_bkgd, out_rgb, out_depth, weights, alpha,
        transparency, ray_mask);
    ND_CUDA_LAUNCH_CHECK("k_composite");
    return ND_OK;
}

int nd_volume_sample_trilinear(const float *volume, int channels, int d0, int d1, int d2, const float *pts,
                               int64_t n_points, const float *aabb_min_host, const float *aabb_max_host, float *out,
                               uint8_t *inside, void *stream) {
    ND_REQUIRE(volume && pts && aabb_min_host && aabb_max_host && out, ND_ERR_BAD_ARG,
               "nd_volume_sample_trilinear: null pointer");
    ND_REQUIRE(channels > 0 && d0 > 0 && d1 > 0 && d2 > 0 && n_points >= 0, ND_ERR_BAD_SHAPE,
               "nd_volume_sample_trilinear: bad shape");
    if (n_points == 0) return ND_OK;
    // 1 / (aabb_max - aabb_min) * 2 as the reference evaluates it in fp32 (render_ray.py:34)
    float3 lo = make_float3(aabb_min_host[0], aabb_min_host[1], aabb_min_host[2]);
    float3 inv;
    inv.x = (1.0f / (aabb_max_host[0] - aabb_min_host[0])) * 2.0f;
    inv.y = (1.0f / (aabb_max_host[1] - aabb_min_host[1])) * 2.0f;
    inv.z = (1.0f / (aabb_max_host[2] - aabb_min_host[2])) * 2.0f;
    k_volume_sample<<<(unsigned)ceil_div(n_points, 128), 128, 0, (cudaStream_t)stream>>>(volume, channels, d0, d1, d2, pts,
                                                                                        n_points, lo, inv, out, inside);
    ND_CUDA_LAUNCH_CHECK("k_volume_sample");
    return ND_OK;
}

}  // extern "C"
