// Voxel lifting kernels (B3, B5, B6 of SURVEY.md section 8a) for sm_100a.
//
// Data flow of the fused path (replaces reference nerfdet.py:164-181):
//
//   k_pixel_index     points, projection            -> pix[N][nvp] int32   (nearest pixel or -1)
//   per channel chunk (chunk = 32/64/128/256 channels, sized so that the staging
//   buffer of ALL views stays resident in the 126 MB L2):
//     k_to_pixel_major  NCHW slice (streamed once from HBM) -> stage[nv][P][chunk]
//     k_lift_gather     one warp per voxel, lanes = channels: for every VALID view the
//                       warp reads one contiguous pixel row of `stage`, keeps sum /
//                       sum-of-squares in registers, and the CTA writes mean /
//                       exp(-var) back transposed through shared memory as full
//                       128-byte row segments of the [C][N] outputs.
//
// The per-view volume [nv][C][N] of the reference (1.3 GB at nv=50) never exists.
#include "nd_common.cuh"

namespace nd {

constexpr int kGatherThreads = 256;
constexpr int kTileVox = 32;           // voxels per gather CTA (one 128 B output segment per row)
constexpr int kUnroll = 4;             // pixel rows in flight per warp

// --------------------------------------------------------------------------------------
// B3 debug / parity op: x, y, valid exactly as the reference computes them.
// --------------------------------------------------------------------------------------
__global__ void k_project_voxels(const float *__restrict__ points, const float *__restrict__ proj,
                                 int64_t n_vox, int height, int width, int64_t *__restrict__ xo,
                                 int64_t *__restrict__ yo, uint8_t *__restrict__ valid) {
    const int v = blockIdx.y;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float p[12];
    if (threadIdx.x < 12) p[threadIdx.x] = proj[v * 12 + threadIdx.x];
    __syncthreads();
    if (n >= n_vox) return;
    float xr, yr, q2;
    const bool ok = project_nearest(p, points[n], points[n_vox + n], points[2 * n_vox + n], height, width,
                                    xr, yr, q2);
    const int64_t o = (int64_t)v * n_vox + n;
    // float -> int64 like .long(): values on invalid lanes are unspecified by contract
    xo[o] = (int64_t)xr;
    yo[o] = (int64_t)yr;
    valid[o] = ok ? 1 : 0;
}

// --------------------------------------------------------------------------------------
// Pre-pass: nearest-pixel table pix[n][v] (row pitch nvp, multiple of 32), -1 = invalid.
// Stored value = y * pix_sy + x * pix_sx  (pixel units of the buffer the gather reads).
// --------------------------------------------------------------------------------------
__global__ void k_pixel_index(const float *__restrict__ points, const float *__restrict__ proj, int nv,
                              int nvp, int64_t n_vox, int height, int width, int pix_sy, int pix_sx,
                              int32_t *__restrict__ pix) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = t / nvp;
    const int v = (int)(t % nvp);
    if (n >= n_vox) return;
    int32_t out = -1;
    if (v < nv) {
        float xr, yr, q2;
        if (project_nearest(proj + v * 12, points[n], points[n_vox + n], points[2 * n_vox + n], height,
                            width, xr, yr, q2))
            out = (int32_t)yr * pix_sy + (int32_t)xr * pix_sx;
    }
    pix[t] = out;
}

// --------------------------------------------------------------------------------------
// NCHW slice -> pixel-major staging  stage[v][p][chunk]  for channels [c0, c0 + chunk).
// 64 px x 64 ch tiles through shared memory; input is read with streaming (evict-first)
// loads, output is written with default policy so it stays in L2 for the gather.
// --------------------------------------------------------------------------------------
template <typename T> struct Pack4;
template <> struct Pack4<float> {
    using type = float4;
    static __device__ __forceinline__ type make(float a, float b, float c, float d) {
        return make_float4(a, b, c, d);
    }
};
template <> struct Pack4<__nv_bfloat16> {
    using type = uint2;
    static __device__ __forceinline__ type make(__nv_bfloat16 a, __nv_bfloat16 b, __nv_bfloat16 c,
                                                __nv_bfloat16 d) {
        __nv_bfloat162 lo = __halves2bfloat162(a, b), hi = __halves2bfloat162(c, d);
        uint2 r;
        r.x = *reinterpret_cast<uint32_t *>(&lo);
        r.y = *reinterpret_cast<uint32_t *>(&hi);
        return r;
    }
};

template <typename T, bool kPlaneContig>
__global__ void __launch_bounds__(256)
k_to_pixel_major(const T *__restrict__ in, int64_t sv, int64_t sc, int64_t sy, int64_t sx, int width,
                 int n_pix, int c0, int c_valid, int chunk, T *__restrict__ stage) {
    __shared__ T tile[64][65];
    const int v = blockIdx.y;
    const int p0 = blockIdx.x * 64;
    const int tid = threadIdx.x;
    const T zero = T(0.0f);
    for (int cb = 0; cb < chunk; cb += 64) {
        const int nch = min(64, chunk - cb);           // channels of this sub-tile (32 or 64)
        // ---- read: rows = channels, columns = pixels ----
        if constexpr (kPlaneContig && sizeof(T) == 4) {
            const int q = tid & 15, r = tid >> 4;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int ch = rr * 16 + r;
                if (ch >= nch) continue;
                const int p = p0 + 4 * q;
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cb + ch < c_valid) {
                    const float *src = reinterpret_cast<const float *>(in) + v * sv + (int64_t)(c0 + cb + ch) * sc + p;
                    if (p + 3 < n_pix && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
                        val = ld_stream_f4(reinterpret_cast<const float4 *>(src));
                    } else {
                        if (p + 0 < n_pix) val.x = ld_stream(src + 0);
                        if (p + 1 < n_pix) val.y = ld_stream(src + 1);
                        if (p + 2 < n_pix) val.z = ld_stream(src + 2);
                        if (p + 3 < n_pix) val.w = ld_stream(src + 3);
                    }
                }
                float *dst = reinterpret_cast<float *>(&tile[ch][4 * q]);
                dst[0] = val.x; dst[1] = val.y; dst[2] = val.z; dst[3] = val.w;
            }
        } else {
            const int px = tid & 63, r = tid >> 6;
            const int p = p0 + px;
            int64_t off = 0;
            if (p < n_pix) off = kPlaneContig ? (int64_t)p : (int64_t)(p / width) * sy + (int64_t)(p % width) * sx;
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
                const int ch = rr * 4 + r;
                if (ch >= nch) continue;
                T val = zero;
                if (p < n_pix && cb + ch < c_valid) val = __ldcs(in + v * sv + (int64_t)(c0 + cb + ch) * sc + off);
                tile[ch][px] = val;
            }
        }
        __syncthreads();
        // ---- write: rows = pixels, 4 channels per thread ----
        {
            const int cq = tid & 15, pr = tid >> 4;
            if (4 * cq < nch) {
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    const int px = pp * 16 + pr;
                    const int p = p0 + px;
                    if (p >= n_pix) continue;
                    typename Pack4<T>::type val = Pack4<T>::make(tile[4 * cq + 0][px], tile[4 * cq + 1][px],
                                                                 tile[4 * cq + 2][px], tile[4 * cq + 3][px]);
                    T *dst = stage + ((int64_t)v * n_pix + p) * chunk + cb + 4 * cq;
                    *reinterpret_cast<typename Pack4<T>::type *>(dst) = val;
                }
            }
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------------
// Row loader: one pixel row of `chunk` = 32*CPL channels, CPL channels per lane.
// --------------------------------------------------------------------------------------
template <typename T, int CPL> struct RowLoad;

template <int CPL> struct RowLoad<float, CPL> {
    // CPL <= 4: channels lane*CPL + j;  CPL == 8: {4*lane + j} and {128 + 4*lane + j}
    static __device__ __forceinline__ int channel(int lane, int j) {
        if constexpr (CPL == 8) return (j < 4) ? 4 * lane + j : 128 + 4 * lane + (j - 4);
        return lane * CPL + j;
    }
    static __device__ __forceinline__ void load(const float *__restrict__ row, int lane, float (&f)[CPL]) {
        if constexpr (CPL == 1) {
            f[0] = __ldg(row + lane);
        } else if constexpr (CPL == 2) {
            const float2 a = __ldg(reinterpret_cast<const float2 *>(row) + lane);
            f[0] = a.x; f[1] = a.y;
        } else if constexpr (CPL == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(row) + lane);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        } else {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(row) + lane);
            const float4 b = __ldg(reinterpret_cast<const float4 *>(row) + 32 + lane);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
            f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        }
    }
};

template <int CPL> struct RowLoad<__nv_bfloat16, CPL> {
    static __device__ __forceinline__ int channel(int lane, int j) { return lane * CPL + j; }
    static __device__ __forceinline__ void unpack(uint32_t w, float &a, float &b) {
        a = __uint_as_float(w << 16);
        b = __uint_as_float(w & 0xffff0000u);
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *__restrict__ row, int lane, float (&f)[CPL]) {
        if constexpr (CPL == 1) {
            f[0] = __bfloat162float(row[lane]);
        } else if constexpr (CPL == 2) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(row) + lane);
            unpack(w, f[0], f[1]);
        } else if constexpr (CPL == 4) {
            const uint2 w = __ldg(reinterpret_cast<const uint2 *>(row) + lane);
            unpack(w.x, f[0], f[1]);
            unpack(w.y, f[2], f[3]);
        } else {
            const uint4 w = __ldg(reinterpret_cast<const uint4 *>(row) + lane);
            unpack(w.x, f[0], f[1]);
            unpack(w.y, f[2], f[3]);
            unpack(w.z, f[4], f[5]);
            unpack(w.w, f[6], f[7]);
        }
    }
};

// --------------------------------------------------------------------------------------
// Gather + statistics.  kRaw = false: writes mean / exp(-var) (fused single-GPU path);
// kRaw = true: writes raw sum / sum-of-squares for the view-sharded all-reduce.
//
// Statistics are accumulated about a per-(voxel, channel) shift K = first valid sample,
// so  sum_valid (f - mean)^2 = S2d - S1d^2 / cnt  has no large-magnitude cancellation.
// The reference's variance runs over ALL views with invalid ones contributing 0
// (nerfdet.py:179):  var * cnt = ssd_valid + (n_views - cnt) * mean^2.
// --------------------------------------------------------------------------------------
struct GatherArgs {
    const void *src;          // pixel-major rows: element (v, pixel) at src + v*view_pitch + pixel*row_pitch
    int64_t view_pitch;       // elements
    int64_t row_pitch;        // elements
    const int32_t *pix;       // [N][nvp]
    int nv, nvp;
    int64_t n_vox;
    int c0, c_valid;          // this chunk covers channels [c0, c0 + c_valid)
    int n_views_total;
    const float *alpha;       // nullable [N]
    float *out_a;             // mean  (or S1)  [C][N]
    float *out_b;             // cov   (or S2)  [C][N], nullable when !kRaw
    int64_t *count_i64;       // nullable (written when c0 == 0 and !kRaw)
    float *count_f32;         // nullable (written when c0 == 0 and kRaw)
};

template <typename T, int CPL, bool kRaw>
__global__ void __launch_bounds__(kGatherThreads)
k_lift_gather(const GatherArgs a) {
    constexpr int kChunk = 32 * CPL;
    extern __shared__ float smem[];
    float *st_a = smem;                               // [kChunk][kTileVox + 1]
    float *st_b = smem + kChunk * (kTileVox + 1);
    __shared__ int s_cnt[kTileVox];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n0 = (int64_t)blockIdx.x * kTileVox;
    const T *__restrict__ src = reinterpret_cast<const T *>(a.src);
    constexpr int kVoxPerWarp = kTileVox / (kGatherThreads / 32);

#pragma unroll 1
    for (int i = 0; i < kVoxPerWarp; ++i) {
        const int vt = warp * kVoxPerWarp + i;
        const int64_t n = n0 + vt;
        float s1[CPL], s2[CPL], k0[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) { s1[j] = 0.f; s2[j] = 0.f; k0[j] = 0.f; }
        int cnt = 0;
        bool first = true;
        if (n < a.n_vox) {
            const int32_t *prow = a.pix + n * a.nvp;
            for (int vb = 0; vb < a.nv; vb += 32) {
                const int idx = (vb + lane < a.nv) ? __ldg(prow + vb + lane) : -1;
                unsigned m = __ballot_sync(0xffffffffu, idx >= 0);
                cnt += __popc(m);
                while (m) {
                    float f[kUnroll][CPL];
                    bool has[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        has[u] = (m != 0);
                        const int b = (__ffs(m) - 1) & 31;
                        m &= (m - 1);
                        const int p = __shfl_sync(0xffffffffu, idx, b);
                        if (has[u]) {
                            const T *row = src + (int64_t)(vb + b) * a.view_pitch + (int64_t)p * a.row_pitch;
                            RowLoad<T, CPL>::load(row, lane, f[u]);
                        }
                    }
                    if (first) {
#pragma unroll
                        for (int j = 0; j < CPL; ++j) k0[j] = f[0][j];
                        first = false;
                    }
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        if (has[u]) {
#pragma unroll
                            for (int j = 0; j < CPL; ++j) {
                                const float d = f[u][j] - k0[j];
                                s1[j] += d;
                                s2[j] = fmaf(d, d, s2[j]);
                            }
                        }
                    }
                }
            }
        }
        // ---- per-voxel finalisation into the staging tile ----
        const float cf = (float)cnt;                       // == cnt + 1e-8 in fp32 for cnt >= 1
        const float rest = (float)(a.n_views_total - cnt);
        float al = 1.0f;
        if (!kRaw && a.alpha != nullptr && n < a.n_vox) al = __ldg(a.alpha + n);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            float oa, ob;
            if (kRaw) {
                oa = fmaf(cf, k0[j], s1[j]);
                ob = fmaf(cf * k0[j], k0[j], fmaf(2.0f * k0[j], s1[j], s2[j]));
            } else if (cnt > 0) {
                const float md = s1[j] / cf;
                const float mean = k0[j] + md;
                float ssd = fmaxf(fmaf(-md, s1[j], s2[j]), 0.0f);
                ssd = fmaf(rest * mean, mean, ssd);
                oa = mean * al;
                ob = expf(-(ssd / cf));
            } else {
                oa = 0.0f;                                 // nerfdet.py:176
                ob = 0.0f;                                 // exp(-1e6) == 0 in fp32 (nerfdet.py:180-181)
            }
            const int ch = RowLoad<T, CPL>::channel(lane, j);
            st_a[ch * (kTileVox + 1) + vt] = oa;
            st_b[ch * (kTileVox + 1) + vt] = ob;
        }
        if (lane == 0) s_cnt[vt] = cnt;
    }
    __syncthreads();
    // ---- transposed write-back: each warp stores full 128 B segments of the [C][N] rows ----
    const int64_t n = n0 + lane;
    if (n < a.n_vox) {
        for (int ch = warp; ch < a.c_valid; ch += kGatherThreads / 32) {
            const int64_t o = (int64_t)(a.c0 + ch) * a.n_vox + n;
            st_stream(a.out_a + o, st_a[ch * (kTileVox + 1) + lane]);
            if (a.out_b != nullptr) st_stream(a.out_b + o, st_b[ch * (kTileVox + 1) + lane]);
        }
        if (a.c0 == 0 && warp == 0) {
            if (a.count_i64 != nullptr) a.count_i64[n] = (int64_t)s_cnt[lane];
            if (a.count_f32 != nullptr) a.count_f32[n] = (float)s_cnt[lane];
        }
    }
}

// --------------------------------------------------------------------------------------
// Finalise from all-reduced raw accumulators (view-sharded path).
// --------------------------------------------------------------------------------------
__global__ void k_lift_finalize(const float *__restrict__ s1, const float *__restrict__ s2,
                                const float *__restrict__ cnt, int n_views_total, int channels,
                                int64_t n_vox, const float *__restrict__ alpha, float *__restrict__ mean,
                                float *__restrict__ cov, int64_t *__restrict__ count) {
    const int64_t total = (int64_t)channels * n_vox;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i % n_vox;
        const float cf = __ldg(cnt + n);
        float m = 0.f, cv = 0.f;
        if (cf > 0.f) {
            const float a = s1[i], b = s2[i];
            m = a / cf;
            float ssd = fmaxf(fmaf(-m, a, b), 0.0f);
            ssd = fmaf(((float)n_views_total - cf) * m, m, ssd);
            cv = expf(-(ssd / cf));
            if (alpha != nullptr) m *= __ldg(alpha + n);
        }
        mean[i] = m;
        if (cov != nullptr) cov[i] = cv;
        if (count != nullptr && i < n_vox) count[i] = (int64_t)cf;
    }
}

// --------------------------------------------------------------------------------------
// Compatibility path: materialised per-view volume (reference backproject, nerfdet.py:393-420).
// --------------------------------------------------------------------------------------
template <typename T>
__global__ void k_backproject(const T *__restrict__ in, int64_t sv, int64_t sc, int64_t sy, int64_t sx,
                              int channels, int height, int width, const float *__restrict__ points,
                              const float *__restrict__ proj, int64_t n_vox,
                              const float *__restrict__ depth, float voxel_z, float *__restrict__ volume,
                              uint8_t *__restrict__ valid) {
    const int v = blockIdx.y;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float p[12];
    if (threadIdx.x < 12) p[threadIdx.x] = proj[v * 12 + threadIdx.x];
    __syncthreads();
    if (n >= n_vox) return;
    float xr, yr, q2;
    bool ok = project_nearest(p, points[n], points[n_vox + n], points[2 * n_vox + n], height, width, xr, yr, q2);
    int xi = 0, yi = 0;
    if (ok) {
        xi = (int)xr;
        yi = (int)yr;
        if (depth != nullptr) {                            // B4, nerfdet.py:405-411
            const float d = depth[((int64_t)v * height + yi) * width + xi];
            ok = (q2 > __fsub_rn(d, voxel_z)) && (q2 < __fadd_rn(d, voxel_z));
        }
    }
    valid[(int64_t)v * n_vox + n] = ok ? 1 : 0;
    const T *src = in + v * sv + (int64_t)yi * sy + (int64_t)xi * sx;
    float *dst = volume + (int64_t)v * channels * n_vox + n;
    for (int c = 0; c < channels; ++c) dst[(int64_t)c * n_vox] = ok ? to_f32<T>(src[(int64_t)c * sc]) : 0.0f;
}

// ======================================================================================
// Host side
// ======================================================================================
// plane-resident path (lift_quads.cu)
bool lift_quads_eligible(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt);
size_t lift_quads_plan_bytes(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt);
nd_status lift_quads_plan_build(const nd_maps *f, const float *points, const float *proj, int64_t n_vox, const float *depth,
                                float voxel_z, void *plan, size_t plan_bytes, const nd_lift_options *opt, cudaStream_t st);
template <typename T, bool kRaw>
nd_status lift_quads_run(const nd_maps *f, const void *plan, size_t plan_bytes, int64_t n_vox, uint32_t launch_index,
                         int n_views_total, const float *alpha, float *out_a, float *out_b, int64_t *count_i64,
                         float *count_f32, const nd_lift_options *opt, cudaStream_t st);

static size_t scratch_budget(const nd_lift_options *opt) {
    return (opt && opt->scratch_budget_bytes) ? opt->scratch_budget_bytes : ((size_t)64 << 20);
}

struct LiftPlan {
    bool quads;             // plane-resident path (lift_quads.cu): contiguous NCHW planes in shared memory
    int nv, nvp, c, h, w, n_pix;
    int elt;                // bytes per feature element
    bool direct;            // features already pixel-major (channels-last): no staging
    int chunk;              // channels per staged chunk (32/64/128/256)
    int n_chunks;
    size_t pix_bytes, stage_bytes, total_bytes;
    int64_t direct_view_pitch, direct_row_pitch;
    int pix_sy, pix_sx;
};

static nd_status validate_maps(const nd_maps *f, const char *who) {
    ND_REQUIRE(f != nullptr && f->data != nullptr, ND_ERR_BAD_ARG, "%s: null feature maps", who);
    ND_REQUIRE(f->dtype == ND_F32 || f->dtype == ND_BF16, ND_ERR_BAD_ARG, "%s: unsupported dtype %d", who, f->dtype);
    ND_REQUIRE(f->n_views > 0 && f->channels > 0 && f->height > 0 && f->width > 0, ND_ERR_BAD_SHAPE,
               "%s: empty feature maps (%d views, %d channels, %dx%d)", who, f->n_views, f->channels, f->height,
               f->width);
    ND_REQUIRE((int64_t)f->height * f->width < (1ll << 30), ND_ERR_BAD_SHAPE, "%s: feature map too large", who);
    return ND_OK;
}

static LiftPlan make_plan(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt) {
    LiftPlan p{};
    p.nv = f->n_views;
    p.nvp = (int)align_up((size_t)p.nv, 32);
    p.c = f->channels;
    p.h = f->height;
    p.w = f->width;
    p.n_pix = p.h * p.w;
    p.elt = f->dtype == ND_F32 ? 4 : 2;
    p.pix_bytes = align_up((size_t)n_vox * p.nvp * sizeof(int32_t), 256);
    const int vec_elems = 16 / p.elt;
    // channels-last input: rows of C contiguous channels per pixel, usable in place
    p.direct = f->stride_c == 1 && f->stride_x >= p.c && f->stride_y % f->stride_x == 0 &&
               f->stride_x % vec_elems == 0 && f->stride_v % vec_elems == 0 &&
               (reinterpret_cast<uintptr_t>(f->data) % 16) == 0 &&
               (f->stride_y / f->stride_x) * (int64_t)p.h < (1ll << 30) && p.c % 32 == 0 && p.c <= 256 &&
               (p.c == 32 || p.c == 64 || p.c == 128 || p.c == 256);
    const bool force_staged = opt != nullptr && opt->path == ND_LIFT_PATH_STAGED;
    p.quads = !p.direct && !force_staged && lift_quads_eligible(f, n_vox, opt);
    if (p.quads) {
        p.total_bytes = lift_quads_plan_bytes(f, n_vox, opt);
        return p;
    }
    if (p.direct) {
        p.chunk = p.c;
        p.n_chunks = 1;
        p.stage_bytes = 0;
        p.direct_view_pitch = f->stride_v;
        p.direct_row_pitch = f->stride_x;
        p.pix_sy = (int)(f->stride_y / f->stride_x);
        p.pix_sx = 1;
    } else {
        const size_t budget = scratch_budget(opt);
        int chunk = 256;
        while (chunk > 32 && (chunk / 2 >= p.c || (size_t)p.nv * p.n_pix * chunk * p.elt > budget)) chunk /= 2;
        p.chunk = chunk;
        p.n_chunks = (int)ceil_div(p.c, chunk);
        p.stage_bytes = align_up((size_t)p.nv * p.n_pix * chunk * p.elt, 256);
        p.pix_sy = p.w;
        p.pix_sx = 1;
    }
    p.total_bytes = p.pix_bytes + p.stage_bytes;
    return p;
}

template <typename T, int CPL, bool kRaw>
static nd_status launch_gather(const GatherArgs &ga, cudaStream_t st) {
    constexpr int kChunk = 32 * CPL;
    const size_t smem = (size_t)2 * kChunk * (kTileVox + 1) * sizeof(float);
    auto kern = k_lift_gather<T, CPL, kRaw>;
    if (smem > 48 * 1024)              // per device and idempotent: set on every call (no process-wide state)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = (unsigned)ceil_div(ga.n_vox, kTileVox);
    kern<<<grid, kGatherThreads, smem, st>>>(ga);
    ND_CUDA_LAUNCH_CHECK("k_lift_gather");
    return ND_OK;
}

template <typename T, bool kRaw>
static nd_status dispatch_gather(int chunk, const GatherArgs &ga, cudaStream_t st) {
    switch (chunk) {
        case 32: return launch_gather<T, 1, kRaw>(ga, st);
        case 64: return launch_gather<T, 2, kRaw>(ga, st);
        case 128: return launch_gather<T, 4, kRaw>(ga, st);
        case 256: return launch_gather<T, 8, kRaw>(ga, st);
    }
    set_error("lift: unsupported chunk width %d", chunk);
    return ND_ERR_BAD_ARG;
}

template <typename T>
static nd_status launch_stage(const nd_maps *f, const LiftPlan &p, int c0, int c_valid, T *stage,
                              cudaStream_t st) {
    const bool contig = f->stride_x == 1 && f->stride_y == f->width;
    dim3 grid((unsigned)ceil_div(p.n_pix, 64), (unsigned)p.nv);
    const T *in = reinterpret_cast<const T *>(f->data);
    if (contig)
        k_to_pixel_major<T, true><<<grid, 256, 0, st>>>(in, f->stride_v, f->stride_c, f->stride_y, f->stride_x,
                                                        p.w, p.n_pix, c0, c_valid, p.chunk, stage);
    else
        k_to_pixel_major<T, false><<<grid, 256, 0, st>>>(in, f->stride_v, f->stride_c, f->stride_y, f->stride_x,
                                                         p.w, p.n_pix, c0, c_valid, p.chunk, stage);
    ND_CUDA_LAUNCH_CHECK("k_to_pixel_major");
    return ND_OK;
}

template <typename T, bool kRaw>
static nd_status run_lift(const nd_maps *f, const float *points, const float *proj, int64_t n_vox,
                          const float *alpha, float *out_a, float *out_b, int64_t *count_i64, float *count_f32,
                          void *ws, size_t ws_bytes, const nd_lift_options *opt, cudaStream_t st) {
    const LiftPlan p = make_plan(f, n_vox, opt);
    if (p.quads) {
        // one-shot form: the caller's workspace holds the geometry plan of this call
        ND_REQUIRE(ws != nullptr && ws_bytes >= p.total_bytes, ND_ERR_WORKSPACE,
                   "lift: workspace too small (%zu < %zu bytes)", ws_bytes, p.total_bytes);
        nd_status s = lift_quads_plan_build(f, points, proj, n_vox, nullptr, 0.0f, ws, ws_bytes, opt, st);
        if (s != ND_OK) return s;
        return lift_quads_run<T, kRaw>(f, ws, ws_bytes, n_vox, 0u, f->n_views, alpha, out_a, out_b, count_i64, count_f32,
                                       opt, st);
    }
    ND_REQUIRE(ws != nullptr && ws_bytes >= p.total_bytes, ND_ERR_WORKSPACE,
               "lift: workspace too small (%zu < %zu bytes)", ws_bytes, p.total_bytes);
    ND_REQUIRE((reinterpret_cast<uintptr_t>(ws) % 256) == 0, ND_ERR_BAD_ALIGNMENT, "lift: workspace not 256-byte aligned");
    int32_t *pix = reinterpret_cast<int32_t *>(ws);
    T *stage = reinterpret_cast<T *>(reinterpret_cast<char *>(ws) + p.pix_bytes);
    {
        const int64_t total = n_vox * p.nvp;
        k_pixel_index<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(points, proj, p.nv, p.nvp, n_vox, p.h, p.w,
                                                                      p.pix_sy, p.pix_sx, pix);
        ND_CUDA_LAUNCH_CHECK("k_pixel_index");
    }
    GatherArgs ga{};
    ga.pix = pix;
    ga.nv = p.nv;
    ga.nvp = p.nvp;
    ga.n_vox = n_vox;
    ga.n_views_total = p.nv;
    ga.alpha = alpha;
    ga.out_a = out_a;
    ga.out_b = out_b;
    ga.count_i64 = count_i64;
    ga.count_f32 = count_f32;
    if (p.direct) {
        ga.src = f->data;
        ga.view_pitch = p.direct_view_pitch;
        ga.row_pitch = p.direct_row_pitch;
        ga.c0 = 0;
        ga.c_valid = p.c;
        return dispatch_gather<T, kRaw>(p.chunk, ga, st);
    }
    for (int k = 0; k < p.n_chunks; ++k) {
        const int c0 = k * p.chunk;
        const int c_valid = min(p.chunk, p.c - c0);
        nd_status s = launch_stage<T>(f, p, c0, c_valid, stage, st);
        if (s != ND_OK) return s;
        ga.src = stage;
        ga.view_pitch = (int64_t)p.n_pix * p.chunk;
        ga.row_pitch = p.chunk;
        ga.c0 = c0;
        ga.c_valid = c_valid;
        s = dispatch_gather<T, kRaw>(p.chunk, ga, st);
        if (s != ND_OK) return s;
    }
    return ND_OK;
}

}  // namespace nd

using namespace nd;

extern "C" {

int nd_project_voxels(const float *points, const float *projection, int n_views, int64_t n_voxels, int height,
                      int width, int64_t *x, int64_t *y, uint8_t *valid, void *stream) {
    ND_REQUIRE(points && projection && x && y && valid, ND_ERR_BAD_ARG, "nd_project_voxels: null pointer");
    ND_REQUIRE(n_views > 0 && n_voxels >= 0 && height > 0 && width > 0, ND_ERR_BAD_SHAPE,
               "nd_project_voxels: bad shape");
    if (n_voxels == 0) return ND_OK;
    dim3 grid((unsigned)ceil_div(n_voxels, 256), (unsigned)n_views);
    k_project_voxels<<<grid, 256, 0, (cudaStream_t)stream>>>(points, projection, n_voxels, height, width, x, y, valid);
    ND_CUDA_LAUNCH_CHECK("k_project_voxels");
    return ND_OK;
}

int nd_backproject(const nd_maps *f, const float *points, const float *projection, int64_t n_voxels,
                   const float *depth_resized, float voxel_z, float *volume, uint8_t *valid, void *stream) {
    nd_status s = validate_maps(f, "nd_backproject");
    if (s != ND_OK) return s;
    ND_REQUIRE(points && projection && volume && valid, ND_ERR_BAD_ARG, "nd_backproject: null pointer");
    ND_REQUIRE(n_voxels >= 0, ND_ERR_BAD_SHAPE, "nd_backproject: negative voxel count");
    if (n_voxels == 0) return ND_OK;
    dim3 grid((unsigned)ceil_div(n_voxels, 256), (unsigned)f->n_views);
    cudaStream_t st = (cudaStream_t)stream;
    if (f->dtype == ND_F32)
        k_backproject<float><<<grid, 256, 0, st>>>((const float *)f->data, f->stride_v, f->stride_c, f->stride_y,
                                                   f->stride_x, f->channels, f->height, f->width, points, projection,
                                                   n_voxels, depth_resized, voxel_z, volume, valid);
    else
        k_backproject<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)f->data, f->stride_v, f->stride_c,
                                                           f->stride_y, f->stride_x, f->channels, f->height, f->width,
                                                           points, projection, n_voxels, depth_resized, voxel_z,
                                                           volume, valid);
    ND_CUDA_LAUNCH_CHECK("k_backproject");
    return ND_OK;
}

size_t nd_lift_workspace_bytes(const nd_maps *f, int64_t n_voxels, const nd_lift_options *opt) {
    if (validate_maps(f, "nd_lift_workspace_bytes") != ND_OK || n_voxels < 0) return 0;
    return make_plan(f, n_voxels, opt).total_bytes;
}

int nd_lift_launch_count(const nd_maps *f, int64_t n_voxels, const nd_lift_options *opt) {
    if (validate_maps(f, "nd_lift_launch_count") != ND_OK || n_voxels < 0) return -1;
    const LiftPlan p = make_plan(f, n_voxels, opt);
    if (p.quads) return 4;   // k_q_index, k_q_rank, k_q_pack, k_lift_quads (1 when the geometry plan is reused: nd_lift_plan_*)
    if (p.direct) return 2;
    return 1 + 2 * p.n_chunks;
}

int nd_lift_mean_var(const nd_maps *f, const float *points, const float *projection, int64_t n_voxels,
                     const float *alpha, float *mean, float *cov, int64_t *count, void *workspace,
                     size_t workspace_bytes, const nd_lift_options *opt, void *stream) {
    nd_status s = validate_maps(f, "nd_lift_mean_var");
    if (s != ND_OK) return s;
    ND_REQUIRE(points && projection && mean && count, ND_ERR_BAD_ARG, "nd_lift_mean_var: null pointer");
    ND_REQUIRE(n_voxels >= 0, ND_ERR_BAD_SHAPE, "nd_lift_mean_var: negative voxel count");
    if (n_voxels == 0) return ND_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (f->dtype == ND_F32)
        return run_lift<float, false>(f, points, projection, n_voxels, alpha, mean, cov, count, nullptr, workspace,
                                      workspace_bytes, opt, st);
    return run_lift<__nv_bfloat16, false>(f, points, projection, n_voxels, alpha, mean, cov, count, nullptr,
                                          workspace, workspace_bytes, opt, st);
}

int nd_lift_accumulate(const nd_maps *f, const float *points, const float *projection, int64_t n_voxels, float *s1,
                       float *s2, float *cnt, void *workspace, size_t workspace_bytes, const nd_lift_options *opt,
                       void *stream) {
    nd_status s = validate_maps(f, "nd_lift_accumulate");
    if (s != ND_OK) return s;
    ND_REQUIRE(points && projection && s1 && cnt, ND_ERR_BAD_ARG, "nd_lift_accumulate: null pointer");
    ND_REQUIRE(n_voxels >= 0, ND_ERR_BAD_SHAPE, "nd_lift_accumulate: negative voxel count");
    if (n_voxels == 0) return ND_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (f->dtype == ND_F32)
        return run_lift<float, true>(f, points, projection, n_voxels, nullptr, s1, s2, nullptr, cnt, workspace,
                                     workspace_bytes, opt, st);
    return run_lift<__nv_bfloat16, true>(f, points, projection, n_voxels, nullptr, s1, s2, nullptr, cnt, workspace,
                                         workspace_bytes, opt, st);
}

size_t nd_lift_plan_bytes(const nd_maps *f, int64_t n_voxels, const nd_lift_options *opt) {
    if (f == nullptr || n_voxels <= 0 || f->n_views <= 0 || f->channels <= 0 || f->height <= 0 || f->width <= 0) return 0;
    if (f->dtype != ND_F32 && f->dtype != ND_BF16) return 0;
    if (opt != nullptr && opt->path == ND_LIFT_PATH_STAGED) return 0;
    return lift_quads_plan_bytes(f, n_voxels, opt);
}

int nd_lift_plan_build(const nd_maps *f, const float *points, const float *projection, int64_t n_voxels,
                       const float *depth_resized, float voxel_z, void *plan, size_t plan_bytes,
                       const nd_lift_options *opt, void *stream) {
    ND_REQUIRE(f != nullptr && points && projection && plan, ND_ERR_BAD_ARG, "nd_lift_plan_build: null pointer");
    ND_REQUIRE(f->dtype == ND_F32 || f->dtype == ND_BF16, ND_ERR_BAD_ARG, "nd_lift_plan_build: unsupported dtype %d", f->dtype);
    ND_REQUIRE(f->n_views > 0 && f->channels > 0 && f->height > 0 && f->width > 0 && n_voxels > 0, ND_ERR_BAD_SHAPE,
               "nd_lift_plan_build: empty input");
    return lift_quads_plan_build(f, points, projection, n_voxels, depth_resized, voxel_z, plan, plan_bytes, opt,
                                 (cudaStream_t)stream);
}

int nd_lift_plan_mean_var(const nd_maps *f, const void *plan, size_t plan_bytes, int64_t n_voxels, uint32_t launch_index,
                          int n_views_total, const float *alpha, float *mean, float *cov, int64_t *count,
                          const nd_lift_options *opt, void *stream) {
    nd_status s = validate_maps(f, "nd_lift_plan_mean_var");
    if (s != ND_OK) return s;
    ND_REQUIRE(plan && mean && count, ND_ERR_BAD_ARG, "nd_lift_plan_mean_var: null pointer");
    ND_REQUIRE(n_voxels > 0 && n_views_total >= 0, ND_ERR_BAD_SHAPE, "nd_lift_plan_mean_var: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (f->dtype == ND_F32)
        return lift_quads_run<float, false>(f, plan, plan_bytes, n_voxels, launch_index, n_views_total, alpha, mean, cov,
                                            count, nullptr, opt, st);
    return lift_quads_run<__nv_bfloat16, false>(f, plan, plan_bytes, n_voxels, launch_index, n_views_total, alpha, mean,
                                                cov, count, nullptr, opt, st);
}

int nd_lift_plan_accumulate(const nd_maps *f, const void *plan, size_t plan_bytes, int64_t n_voxels, uint32_t launch_index,
                            float *s1, float *s2, float *cnt, const nd_lift_options *opt, void *stream) {
    nd_status s = validate_maps(f, "nd_lift_plan_accumulate");
    if (s != ND_OK) return s;
    ND_REQUIRE(plan && s1 && cnt, ND_ERR_BAD_ARG, "nd_lift_plan_accumulate: null pointer");
    ND_REQUIRE(n_voxels > 0, ND_ERR_BAD_SHAPE, "nd_lift_plan_accumulate: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (f->dtype == ND_F32)
        return lift_quads_run<float, true>(f, plan, plan_bytes, n_voxels, launch_index, 0, nullptr, s1, s2, nullptr, cnt,
                                           opt, st);
    return lift_quads_run<__nv_bfloat16, true>(f, plan, plan_bytes, n_voxels, launch_index, 0, nullptr, s1, s2, nullptr,
                                               cnt, opt, st);
}

int nd_lift_finalize(const float *s1, const float *s2, const float *cnt, int n_views_total, int channels,
                     int64_t n_voxels, const float *alpha, float *mean, float *cov, int64_t *count, void *stream) {
    ND_REQUIRE(s1 && s2 && cnt && mean, ND_ERR_BAD_ARG, "nd_lift_finalize: null pointer");
    ND_REQUIRE(channels > 0 && n_voxels >= 0 && n_views_total > 0, ND_ERR_BAD_SHAPE, "nd_lift_finalize: bad shape");
    if (n_voxels == 0) return ND_OK;
    const int64_t total = (int64_t)channels * n_voxels;
    const unsigned grid = (unsigned)(ceil_div(total, 256) < 148 * 16 ? ceil_div(total, 256) : 148 * 16);
    k_lift_finalize<<<grid, 256, 0, (cudaStream_t)stream>>>(s1, s2, cnt, n_views_total, channels, n_voxels, alpha, mean,
                                                            cov, count);
    ND_CUDA_LAUNCH_CHECK("k_lift_finalize");
    return ND_OK;
}

}  // extern "C"
