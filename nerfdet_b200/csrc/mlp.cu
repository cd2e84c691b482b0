// The shared NeRF / geometry MLP (row M of SURVEY.md section 8a; reference nerf_mlp.py:11-234,
// instantiated at nerfdet.py:62-69 with net_depth 4, net_width 256, skip_layer 3, feature_dim 70).
//
//   in   = [posenc(x) (3 + 6 * pos_octaves), features (F)]                      nerf_mlp.py:181-197, 209-216
//   h    = relu(L_i h), i < depth; after layer i with i % skip == 0, i > 0:  h = [h, in]   nerf_mlp.py:80-90
//   sig  = relu(w_sigma . h + b)                                                nerf_mlp.py:138-144
//   bott = W_b h + b_b ;  rgb = sigmoid(W_o relu(W_r [bott, viewenc(d)] + b_r) + b_o)      nerf_mlp.py:146-161
//
// This file is the fp32 path (FFMA, fp32 accumulate): it is the one that meets the 1e-4 relative
// tolerance of BASELINE.json.  One CTA owns a tile of 64 points and walks the whole network with the
// activations resident in shared memory (they never touch HBM); the pre-transposed weights
// (nd_pack_mlp_weights, k-major so that 32 lanes read 32 x 32 B contiguous) stream from L2 through a
// cp.async double buffer of 16-row chunks.  Each thread owns an 8-point x NT-output register tile.
#include <math.h>
#include <stdlib.h>

#include "nd_common.cuh"

namespace nd {

constexpr int kMlpTile = 64;           // points per CTA
constexpr int kMlpThreads = 256;
constexpr int kMlpChunk = 16;          // weight rows per cp.async stage
constexpr int kMlpWidth = 256;         // net_width supported by the register tiling
constexpr int kMlpMaxDepth = 8;

struct MlpShape {                                 // what travels as a kernel parameter
    int depth, width, skip, feat, cond_width, pos_oct, view_oct;
};

struct MlpDims {
    int depth, width, skip, feat, cond_width, pos_oct, view_oct;
    int in_dim, in_pad, cond_dim, cond_pad;       // 63 + F (padded to 4), 3 + 6 * view_oct (padded to 4)
    bool cat_after[kMlpMaxDepth];                 // [h, in] concatenation after layer i
    int rows[kMlpMaxDepth];                       // packed K rows of hidden layer i
    int head_rows;                                // packed K rows of the sigma / bottleneck heads
    int rgbh_rows;                                // width + cond_pad
    // float offsets into the packed buffer
    size_t off_w[kMlpMaxDepth], off_b[kMlpMaxDepth], off_sig_w, off_sig_b, off_bot_w, off_bot_b, off_rh_w, off_rh_b,
        off_ro_w, off_ro_b, total;
};

__host__ __device__ static inline int pad4(int v) { return (v + 3) & ~3; }

// layout of the packed buffer; evaluated on the host and (from the 7 shape integers) again inside the kernel
__host__ __device__ static inline void mlp_layout(const MlpShape &sh, MlpDims &d) {
    d.depth = sh.depth; d.width = sh.width; d.skip = sh.skip; d.feat = sh.feat;
    d.cond_width = sh.cond_width; d.pos_oct = sh.pos_oct; d.view_oct = sh.view_oct;
    d.in_dim = 3 + 6 * d.pos_oct + d.feat;
    d.in_pad = pad4(d.in_dim);
    d.cond_dim = 3 + 6 * d.view_oct;
    d.cond_pad = pad4(d.cond_dim);
    size_t off = 0;
    bool cat = false;                              // is the current activation [h, in] ?
    for (int i = 0; i < d.depth; ++i) {
        d.rows[i] = i == 0 ? d.in_pad : d.width + (cat ? d.in_pad : 0);
        d.off_w[i] = off;
        off += (size_t)d.rows[i] * d.width;
        d.off_b[i] = off;
        off += d.width;
        cat = d.skip > 0 && i % d.skip == 0 && i > 0;
        d.cat_after[i] = cat;
    }
    d.head_rows = d.width + (cat ? d.in_pad : 0);
    d.off_sig_w = off; off += d.head_rows;
    d.off_sig_b = off; off += 4;
    d.off_bot_w = off; off += (size_t)d.head_rows * d.width;
    d.off_bot_b = off; off += d.width;
    d.rgbh_rows = d.width + d.cond_pad;
    d.off_rh_w = off; off += (size_t)d.rgbh_rows * d.cond_width;
    d.off_rh_b = off; off += d.cond_width;
    d.off_ro_w = off; off += 3 * (size_t)d.cond_width;
    d.off_ro_b = off; off += 4;
    d.total = off;
}

static bool mlp_dims(const nd_mlp_weights *w, MlpDims &d, bool report) {
    MlpShape sh{w->net_depth, w->net_width, w->skip_layer, w->feature_dim, w->cond_width, w->pos_octaves, w->view_octaves};
    if (sh.depth < 1 || sh.depth > kMlpMaxDepth || sh.width != kMlpWidth || sh.cond_width != 128 || sh.feat < 0 ||
        sh.pos_oct < 0 || sh.pos_oct > 16 || sh.view_oct < 0 || sh.view_oct > 16 || sh.skip < 0) {
        if (report)
            set_error("nerf_mlp: unsupported architecture (depth %d, width %d, cond width %d): this build supports "
                      "net_width 256, net_width_condition 128, 1 <= net_depth <= %d",
                      sh.depth, sh.width, sh.cond_width, kMlpMaxDepth);
        return false;
    }
    mlp_layout(sh, d);
    if (d.in_pad > 256) {
        if (report) set_error("nerf_mlp: input width %d too large", d.in_dim);
        return false;
    }
    return true;
}

// ---- packing: reference [out][in] -> k-major [rows][out], segment paddings zero-filled -------------------
// ref column r of the reference weight maps to packed row `r` when r < seg0, else to seg0_pad + (r - seg0).
__global__ void k_pack_linear(const float *__restrict__ w, const float *__restrict__ b, int n_out, int k_ref, int seg0,
                              int seg0_pad, int rows, float *__restrict__ dw, float *__restrict__ db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * n_out) {
        const int r = i / n_out, o = i - r * n_out;
        int col = -1;
        if (r < seg0) col = r;
        else if (r >= seg0_pad && (r - seg0_pad) + seg0 < k_ref) col = r - seg0_pad + seg0;
        dw[i] = col >= 0 ? w[(size_t)o * k_ref + col] : 0.0f;
    }
    if (i < n_out && db != nullptr) db[i] = b != nullptr ? b[i] : 0.0f;
}

__global__ void k_copy_f32(const float *__restrict__ src, float *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// ---- device helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ void cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

struct Seg {                  // one K segment of a layer input: activations in shared memory
    const float *act;         // [kMlpTile][pitch]
    int pitch, rows;          // rows % 4 == 0
};

// out[p][n] = act(bias[n] + sum_k in[p][k] * Wt[k][n]) for the CTA's 64 points, N = 32 * NT outputs.
// Thread (tx = tid % 32, ty = tid / 32) owns points ty*8 .. ty*8+7 and outputs tx*NT .. tx*NT+NT-1.
template <int NT>
__device__ __forceinline__ void dense_layer(const Seg *segs, int n_seg, const float *__restrict__ wt,
                                            const float *__restrict__ bias, float *wbuf, float *out, int out_pitch,
                                            bool relu) {
    constexpr int N = 32 * NT;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    float acc[8][NT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = 0.0f;

    int total_rows = 0;
    for (int s = 0; s < n_seg; ++s) total_rows += segs[s].rows;
    const int n_chunks = (total_rows + kMlpChunk - 1) / kMlpChunk;
    const uint32_t wb = (uint32_t)__cvta_generic_to_shared(wbuf);
    auto load_chunk = [&](int ch) {
        const int r0 = ch * kMlpChunk;
        const int nr = min(kMlpChunk, total_rows - r0);
        const int n_vec = nr * (N / 4);
        const float *src = wt + (size_t)r0 * N;
        const uint32_t dst = wb + (uint32_t)(ch & 1) * (kMlpChunk * N * 4);
        for (int v = tid; v < n_vec; v += kMlpThreads) cp16(dst + v * 16, src + v * 4);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_chunk(0);
    int seg = 0, seg_r = 0;                        // position of the current chunk inside the segment list
    for (int ch = 0; ch < n_chunks; ++ch) {
        if (ch + 1 < n_chunks) {
            load_chunk(ch + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float *wc = wbuf + (ch & 1) * (kMlpChunk * N);
        const int nr = min(kMlpChunk, total_rows - ch * kMlpChunk);
        for (int kk = 0; kk < nr; kk += 4) {
            while (seg_r >= segs[seg].rows) { seg_r -= segs[seg].rows; ++seg; }
            const float *ap = segs[seg].act + (size_t)(ty * 8) * segs[seg].pitch + seg_r;
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(ap + (size_t)i * segs[seg].pitch);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float w[NT];
#pragma unroll
                for (int j = 0; j < NT; j += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(wc + (kk + r) * N + tx * NT + j);
                    w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av = r == 0 ? a[i].x : r == 1 ? a[i].y : r == 2 ? a[i].z : a[i].w;
#pragma unroll
                    for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(av, w[j], acc[i][j]);
                }
            }
            seg_r += 4;
        }
        __syncthreads();                           // the buffer this chunk used is refilled two chunks later
    }
    float bv[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) bv[j] = __ldg(bias + tx * NT + j);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float *op = out + (size_t)(ty * 8 + i) * out_pitch + tx * NT;
#pragma unroll
        for (int j = 0; j < NT; j += 4) {
            float4 t;
            t.x = acc[i][j] + bv[j]; t.y = acc[i][j + 1] + bv[j + 1];
            t.z = acc[i][j + 2] + bv[j + 2]; t.w = acc[i][j + 3] + bv[j + 3];
            if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
            *reinterpret_cast<float4 *>(op + j) = t;
        }
    }
    __syncthreads();
}

// [x, sin(x * 2^k) (k outer, xyz inner), sin(x * 2^k + fp32(pi/2))] -- nerf_mlp.py:181-197.  The cosine half is
// the sine of the fp32-ROUNDED sum; accurate sinf, arguments reach ~4000 rad.
__device__ __forceinline__ void encode(const float *xyz, int n_oct, float *dst) {
    const float half_pi = 1.57079637050628662109375f;              // fp32(0.5 * math.pi)
    for (int c = 0; c < 3; ++c) dst[c] = xyz[c];
    for (int k = 0; k < n_oct; ++k) {
        const float sc = (float)(1 << k);
        for (int c = 0; c < 3; ++c) {
            const float xb = __fmul_rn(xyz[c], sc);
            dst[3 + k * 3 + c] = sinf(xb);
            dst[3 + n_oct * 3 + k * 3 + c] = sinf(__fadd_rn(xb, half_pi));
        }
    }
}

struct MlpArgs {
    MlpShape shape;
    const float *packed;
    const float *x;            // [P][3]
    const float *feat;         // [P][F]
    const float *cond;         // [P / samples_per_ray][3] or null (density only)
    int64_t n_points;
    int samples_per_ray;
    float *sigma, *alpha, *rgb;   // [P], [P], [P][3]; any may be null
};

__global__ void __launch_bounds__(kMlpThreads, 1)
k_nerf_mlp(const MlpArgs a) {
    extern __shared__ __align__(16) float sm[];
    MlpDims d;
    mlp_layout(a.shape, d);
    float *s_in = sm;                                   // [64][in_pad]
    float *s_h0 = s_in + kMlpTile * d.in_pad;           // [64][256]
    float *s_h1 = s_h0 + kMlpTile * kMlpWidth;          // [64][256]
    float *s_cond = s_h1 + kMlpTile * kMlpWidth;        // [64][cond_pad]
    float *wbuf = s_cond + kMlpTile * d.cond_pad;       // [2][16][256]
    const int tid = threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.x * kMlpTile;
    const bool want_rgb = a.rgb != nullptr && a.cond != nullptr;

    // ---- inputs: positional encoding + features (+ view encoding) ----
    const int n_enc = 3 + 6 * d.pos_oct;
    for (int p = tid; p < kMlpTile; p += kMlpThreads) {
        float *row = s_in + (size_t)p * d.in_pad;
        const int64_t gp = p0 + p;
        if (gp < a.n_points) {
            const float xyz[3] = {a.x[gp * 3], a.x[gp * 3 + 1], a.x[gp * 3 + 2]};
            encode(xyz, d.pos_oct, row);
        } else {
            for (int k = 0; k < n_enc; ++k) row[k] = 0.0f;
        }
        for (int k = d.in_dim; k < d.in_pad; ++k) row[k] = 0.0f;
        if (want_rgb) {
            float *cr = s_cond + (size_t)p * d.cond_pad;
            if (gp < a.n_points) {
                const int64_t r = gp / a.samples_per_ray;
                const float dir[3] = {a.cond[r * 3], a.cond[r * 3 + 1], a.cond[r * 3 + 2]};
                encode(dir, d.view_oct, cr);
            } else {
                for (int k = 0; k < d.cond_dim; ++k) cr[k] = 0.0f;
            }
            for (int k = d.cond_dim; k < d.cond_pad; ++k) cr[k] = 0.0f;
        }
    }
    for (int i = tid; i < kMlpTile * d.feat; i += kMlpThreads) {
        const int p = i / d.feat, k = i - p * d.feat;
        const int64_t gp = p0 + p;
        s_in[(size_t)p * d.in_pad + n_enc + k] = gp < a.n_points ? a.feat[gp * d.feat + k] : 0.0f;
    }
    __syncthreads();

    // ---- trunk ----
    float *cur = s_in, *nxt = s_h0;
    bool cat = false;
    for (int i = 0; i < d.depth; ++i) {
        Seg segs[2];
        int ns = 1;
        if (i == 0) {
            segs[0] = Seg{s_in, d.in_pad, d.in_pad};
        } else {
            segs[0] = Seg{cur, kMlpWidth, kMlpWidth};
            if (cat) { segs[1] = Seg{s_in, d.in_pad, d.in_pad}; ns = 2; }
        }
        dense_layer<8>(segs, ns, a.packed + d.off_w[i], a.packed + d.off_b[i], wbuf, nxt, kMlpWidth, true);
        cur = nxt;
        nxt = cur == s_h0 ? s_h1 : s_h0;
        cat = d.cat_after[i];
    }
    // `cur` holds h; the head input is [h] or [h, in]
    // ---- density head: 4 threads per point ----
    {
        const int p = tid >> 2, q = tid & 3;
        const float *ws = a.packed + d.off_sig_w;
        float s = 0.0f;
        const float *hr = cur + (size_t)p * kMlpWidth;
        for (int k = q; k < kMlpWidth; k += 4) s = fmaf(hr[k], __ldg(ws + k), s);
        if (cat) {
            const float *ir = s_in + (size_t)p * d.in_pad;
            for (int k = q; k < d.in_pad; k += 4) s = fmaf(ir[k], __ldg(ws + kMlpWidth + k), s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const int64_t gp = p0 + p;
        if (q == 0 && gp < a.n_points) {
            const float sg = fmaxf(s + __ldg(a.packed + d.off_sig_b), 0.0f);
            if (a.sigma != nullptr) a.sigma[gp] = sg;
            if (a.alpha != nullptr) a.alpha[gp] = 1.0f - expf(-sg);     // nerfdet.py:258
        }
    }
    if (!want_rgb) return;
    // ---- colour branch ----
    {
        Seg segs[2];
        segs[0] = Seg{cur, kMlpWidth, kMlpWidth};
        int ns = 1;
        if (cat) { segs[1] = Seg{s_in, d.in_pad, d.in_pad}; ns = 2; }
        dense_layer<8>(segs, ns, a.packed + d.off_bot_w, a.packed + d.off_bot_b, wbuf, nxt, kMlpWidth, false);
        float *bott = nxt;
        segs[0] = Seg{bott, kMlpWidth, kMlpWidth};
        segs[1] = Seg{s_cond, d.cond_pad, d.cond_pad};
        dense_layer<4>(segs, 2, a.packed + d.off_rh_w, a.packed + d.off_rh_b, wbuf, cur, kMlpWidth, true);
        // output layer 128 -> 3, sigmoid: 4 threads per point
        const int p = tid >> 2, q = tid & 3;
        const float *wo = a.packed + d.off_ro_w;
        const float *tr = cur + (size_t)p * kMlpWidth;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        for (int k = q; k < 128; k += 4) {
            const float t = tr[k];
            c0 = fmaf(t, __ldg(wo + k), c0);
            c1 = fmaf(t, __ldg(wo + 128 + k), c1);
            c2 = fmaf(t, __ldg(wo + 256 + k), c2);
        }
#pragma unroll
        for (int m = 1; m <= 2; m <<= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, m);
            c1 += __shfl_xor_sync(0xffffffffu, c1, m);
            c2 += __shfl_xor_sync(0xffffffffu, c2, m);
        }
        const int64_t gp = p0 + p;
        if (q == 0 && gp < a.n_points) {
            const float *bo = a.packed + d.off_ro_b;
            a.rgb[gp * 3] = 1.0f / (1.0f + expf(-(c0 + __ldg(bo))));
            a.rgb[gp * 3 + 1] = 1.0f / (1.0f + expf(-(c1 + __ldg(bo + 1))));
            a.rgb[gp * 3 + 2] = 1.0f / (1.0f + expf(-(c2 + __ldg(bo + 2))));
        }
    }
}

}  // namespace nd

using namespace nd;

extern "C" {

size_t nd_mlp_packed_bytes(const nd_mlp_weights *w) {
    MlpDims d;
    if (w == nullptr || !mlp_dims(w, d, false)) return 0;
    return d.total * sizeof(float);
}

int nd_pack_mlp_weights(const nd_mlp_weights *w, void *packed, size_t packed_bytes, void *stream) {
    ND_REQUIRE(w != nullptr && packed != nullptr, ND_ERR_BAD_ARG, "nd_pack_mlp_weights: null pointer");
    MlpDims d;
    if (!mlp_dims(w, d, true)) return ND_ERR_BAD_SHAPE;
    ND_REQUIRE(packed_bytes >= d.total * sizeof(float), ND_ERR_WORKSPACE, "nd_pack_mlp_weights: buffer too small (%zu < %zu)",
               packed_bytes, d.total * sizeof(float));
    ND_REQUIRE((reinterpret_cast<uintptr_t>(packed) % 16) == 0, ND_ERR_BAD_ALIGNMENT, "nd_pack_mlp_weights: buffer not 16-byte aligned");
    for (int i = 0; i < d.depth; ++i)
        ND_REQUIRE(w->base_w[i] != nullptr && w->base_b[i] != nullptr, ND_ERR_BAD_ARG, "nd_pack_mlp_weights: layer %d missing", i);
    ND_REQUIRE(w->sigma_w && w->sigma_b && w->bottleneck_w && w->bottleneck_b && w->rgb_hidden_w && w->rgb_hidden_b &&
                   w->rgb_out_w && w->rgb_out_b,
               ND_ERR_BAD_ARG, "nd_pack_mlp_weights: head weights missing");
    cudaStream_t st = (cudaStream_t)stream;
    float *p = reinterpret_cast<float *>(packed);
    auto pack = [&](const float *ww, const float *bb, int n_out, int k_ref, int seg0, int seg0_pad, int rows, size_t ow,
                    size_t ob) {
        const int total = rows * n_out > n_out ? rows * n_out : n_out;
        k_pack_linear<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(ww, bb, n_out, k_ref, seg0, seg0_pad, rows, p + ow,
                                                                      p + ob);
    };
    bool cat = false;
    for (int i = 0; i < d.depth; ++i) {
        const int k_ref = i == 0 ? d.in_dim : d.width + (cat ? d.in_dim : 0);
        // segment 0 = h (width, already a multiple of 4) or the input itself for layer 0
        const int seg0 = i == 0 ? d.in_dim : d.width, seg0_pad = i == 0 ? d.in_pad : d.width;
        pack(w->base_w[i], w->base_b[i], d.width, k_ref, seg0, seg0_pad, d.rows[i], d.off_w[i], d.off_b[i]);
        cat = d.cat_after[i];
    }
    const int head_ref = d.width + (cat ? d.in_dim : 0);
    pack(w->sigma_w, w->sigma_b, 1, head_ref, d.width, d.width, d.head_rows, d.off_sig_w, d.off_sig_b);
    pack(w->bottleneck_w, w->bottleneck_b, d.width, head_ref, d.width, d.width, d.head_rows, d.off_bot_w, d.off_bot_b);
    pack(w->rgb_hidden_w, w->rgb_hidden_b, d.cond_width, d.width + d.cond_dim, d.width, d.width, d.rgbh_rows, d.off_rh_w,
         d.off_rh_b);
    // output layer stays [3][cond_width] (three dot products per point)
    k_copy_f32<<<(unsigned)ceil_div(3 * d.cond_width, 256), 256, 0, st>>>(w->rgb_out_w, p + d.off_ro_w, 3 * d.cond_width);
    k_copy_f32<<<1, 32, 0, st>>>(w->rgb_out_b, p + d.off_ro_b, 3);
    ND_CUDA_LAUNCH_CHECK("k_pack_linear");
    return ND_OK;
}

int nd_nerf_mlp_fwd(const nd_mlp_weights *arch, const void *packed, const float *x, const float *features,
                    const float *cond, int64_t n_points, int samples_per_ray, float *sigma, float *alpha, float *rgb,
                    void *stream) {
    ND_REQUIRE(n_points >= 0, ND_ERR_BAD_SHAPE, "nd_nerf_mlp_fwd: negative point count");
    if (n_points == 0) return ND_OK;
    ND_REQUIRE(arch != nullptr && packed != nullptr && x != nullptr, ND_ERR_BAD_ARG, "nd_nerf_mlp_fwd: null pointer");
    MlpDims d;
    if (!mlp_dims(arch, d, true)) return ND_ERR_BAD_SHAPE;
    ND_REQUIRE(d.feat == 0 || features != nullptr, ND_ERR_BAD_ARG, "nd_nerf_mlp_fwd: features missing");
    ND_REQUIRE(n_points >= 0, ND_ERR_BAD_SHAPE, "nd_nerf_mlp_fwd: negative point count");
    ND_REQUIRE(rgb == nullptr || (cond != nullptr && samples_per_ray > 0 && n_points % samples_per_ray == 0),
               ND_ERR_BAD_SHAPE, "nd_nerf_mlp_fwd: rgb needs cond [P / samples_per_ray][3]");
    if (n_points == 0) return ND_OK;
    MlpArgs a{};
    a.shape = MlpShape{d.depth, d.width, d.skip, d.feat, d.cond_width, d.pos_oct, d.view_oct};
    a.packed = reinterpret_cast<const float *>(packed);
    a.x = x;
    a.feat = features;
    a.cond = cond;
    a.n_points = n_points;
    a.samples_per_ray = samples_per_ray > 0 ? samples_per_ray : 1;
    a.sigma = sigma;
    a.alpha = alpha;
    a.rgb = rgb;
    const size_t smem = ((size_t)kMlpTile * (d.in_pad + 2 * kMlpWidth + d.cond_pad) + 2 * kMlpChunk * kMlpWidth) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(k_nerf_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("nd_nerf_mlp_fwd: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    k_nerf_mlp<<<(unsigned)ceil_div(n_points, kMlpTile), kMlpThreads, smem, (cudaStream_t)stream>>>(a);
    ND_CUDA_LAUNCH_CHECK("k_nerf_mlp");
    return ND_OK;
}

}  // extern "C"
