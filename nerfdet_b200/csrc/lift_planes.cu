// Plane-resident fused lift for the reference's own layout (NCHW maps whose sliced planes
// are contiguous, fp32 or bf16).  Replaces reference nerfdet.py:164-181.
//
// Observation: both sides of the lift are CHANNEL-major -- input planes feat[v][c][pixel] and
// output rows mean/cov[c][voxel] -- so no transposition is needed, only a fast random-access
// memory for one (view, channel) plane (18.9 KB at 59x80 fp32).  That memory is shared memory:
//
//   k_plane_index   one pass over (voxel, view): the bit-exact nearest-pixel projection, stored
//                   as a uint16 BYTE OFFSET into a plane (invalid -> offset of a zero word behind
//                   the plane), the per-voxel view count, and per 512-voxel tile a bitmask of the
//                   views that see any voxel of the tile.               [nv][Np] u16 = 2.6 MB
//   k_lift_planes   work unit = (channel c, part p of the voxel tiles).  A producer warp streams
//                   the nv planes of channel c through an S-stage mbarrier ring with TMA bulk
//                   copies (each plane byte leaves HBM exactly once).  Each compute warp owns one
//                   512-voxel tile, 16 consecutive voxels per lane, and keeps sum / sum-of-squares
//                   for them in 32 registers across all views; per view it loads its 16 offsets
//                   (32 B per lane, 1 KB contiguous per warp, L2-resident), gathers from the plane
//                   in shared memory and accumulates.  Views whose bitmask bit is clear are
//                   skipped by the whole warp.  The epilogue turns the accumulators into
//                   mean / exp(-var) (or raw S1 / S2 for the view-sharded path) and stores 64 B
//                   per lane, contiguous per warp.
//
// Nothing of size [nv][C][N] (the reference's 1.3 GB volume) or [nv][pixel][C] (a pixel-major
// staging copy) is ever written.
#include "nd_common.cuh"

namespace nd {

constexpr int kPV = 16;                 // voxels per lane
constexpr int kPTile = 32 * kPV;        // voxels per warp tile
constexpr int kPMaxWarps = 25;          // compute warps per CTA (+1 producer warp)
constexpr int kPMaxStages = 4;
constexpr int kPMaskWords = 8;          // 64-view words per tile -> nv <= 512

struct PlaneArgs {
    // geometry tables (workspace)
    const uint16_t *off16;     // [nv][n_pad]
    const int32_t *cnt;        // [n_pad]
    const uint64_t *vmask;     // [n_tiles][nw64]
    int nv, nw64;
    int64_t n_vox, n_pad;
    int n_tiles, n_parts;
    // planes
    const void *feat;
    int64_t sv, sc;            // elements
    uint32_t plane_bytes, stage_bytes;
    // outputs
    int n_views_total;
    const float *alpha;
    float *out_a, *out_b;
    int64_t *count_i64;
    float *count_f32;
};

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_bf16(uint32_t addr) {
    uint16_t h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __uint_as_float((uint32_t)h << 16);
}
template <typename T> __device__ __forceinline__ float lds_elt(uint32_t addr);
template <> __device__ __forceinline__ float lds_elt<float>(uint32_t addr) { return lds_f32(addr); }
template <> __device__ __forceinline__ float lds_elt<__nv_bfloat16>(uint32_t addr) { return lds_bf16(addr); }

// ---------------------------------------------------------------------------------------------
// Geometry tables.  One CTA per 512-voxel tile, one thread per voxel, loop over views.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPTile)
k_plane_index(const float *__restrict__ points, const float *__restrict__ proj, int nv, int nw64, int64_t n_vox,
              int64_t n_pad, int height, int width, int elt, uint32_t zero_off, uint16_t *__restrict__ off16,
              int32_t *__restrict__ cnt, uint64_t *__restrict__ vmask) {
    extern __shared__ float sp[];                                   // [nv][12]
    __shared__ unsigned long long smask[kPMaskWords];
    for (int i = threadIdx.x; i < nv * 12; i += blockDim.x) sp[i] = proj[i];
    if (threadIdx.x < kPMaskWords) smask[threadIdx.x] = 0ull;
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * kPTile + threadIdx.x;
    const bool inside = n < n_vox;
    float X = 0.f, Y = 0.f, Z = 0.f;
    if (inside) {
        X = __ldg(points + n);
        Y = __ldg(points + n_vox + n);
        Z = __ldg(points + 2 * n_vox + n);
    }
    int count = 0;
    for (int w = 0; w < nw64; ++w) {
        unsigned long long m = 0ull;
        const int v_end = min(nv, (w + 1) * 64);
#pragma unroll 4
        for (int v = w * 64; v < v_end; ++v) {
            float xr, yr, q2;
            const bool ok = project_nearest(sp + v * 12, X, Y, Z, height, width, xr, yr, q2) && inside;
            const uint32_t off = ok ? (uint32_t)((int)yr * width + (int)xr) * (uint32_t)elt : zero_off;
            off16[(int64_t)v * n_pad + n] = (uint16_t)off;
            count += ok ? 1 : 0;
            if (__any_sync(0xffffffffu, ok)) m |= 1ull << (v & 63);
        }
        if ((threadIdx.x & 31) == 0 && m != 0ull) atomicOr(&smask[w], m);
    }
    cnt[n] = count;
    __syncthreads();
    if (threadIdx.x < nw64) vmask[(int64_t)blockIdx.x * nw64 + threadIdx.x] = smask[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// Gather + statistics with the planes of one channel streamed through shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void gather16(uint32_t sb, const uint4 &a, const uint4 &b, float (&s1)[kPV], float (&s2)[kPV]) {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float f[kPV];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        f[2 * j] = lds_elt<T>(sb + (w[j] & 0xffffu));
        f[2 * j + 1] = lds_elt<T>(sb + (w[j] >> 16));
    }
#pragma unroll
    for (int j = 0; j < kPV; ++j) {
        s1[j] += f[j];
        s2[j] = fmaf(f[j], f[j], s2[j]);
    }
}

template <typename T, int S, bool kRaw>
__global__ void __launch_bounds__((kPMaxWarps + 1) * 32, 1)
k_lift_planes(const PlaneArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = (blockDim.x >> 5) - 1;                           // compute warps
    const int c = blockIdx.x / a.n_parts;
    const int part = blockIdx.x - c * a.n_parts;

    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)S * a.stage_bytes);
    unsigned long long *s_mask = bars + 2 * S;                     // [W][nw64]
    const uint32_t sm_base = smem_u32(smem);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + S);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, (uint32_t)W);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < S) *reinterpret_cast<uint32_t *>(smem + (size_t)threadIdx.x * a.stage_bytes + a.plane_bytes) = 0u;
    for (int i = threadIdx.x; i < W * a.nw64; i += blockDim.x) {
        const int wq = i / a.nw64;
        const int t = part * W + wq;
        s_mask[i] = t < a.n_tiles ? a.vmask[(int64_t)t * a.nw64 + (i - wq * a.nw64)] : 0ull;
    }
    __syncthreads();

    if (warp == W) {
        // ---------------- producer: one elected lane streams the nv planes of channel c ----------------
        if (lane == 0) {
            const char *src = reinterpret_cast<const char *>(a.feat) + (int64_t)c * a.sc * (int64_t)sizeof(T);
            const int64_t view_bytes = a.sv * (int64_t)sizeof(T);
            for (int v = 0; v < a.nv; ++v) {
                const int s = v % S, k = v / S;
                if (k > 0) mbar_wait(bar_empty + 8 * s, (uint32_t)((k - 1) & 1));
                mbar_expect_tx(bar_full + 8 * s, a.plane_bytes);
                bulk_g2s(sm_base + s * a.stage_bytes, src + v * view_bytes, a.plane_bytes, bar_full + 8 * s);
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int tile = part * W + warp;
    const bool tile_ok = tile < a.n_tiles;
    const int64_t n0 = (int64_t)tile * kPTile + lane * kPV;
    const unsigned long long *my_mask = s_mask + warp * a.nw64;
    const uint4 *idx = reinterpret_cast<const uint4 *>(a.off16 + (tile_ok ? n0 : 0));
    const int64_t idx_pitch = a.n_pad / 8;                         // uint4 per view row

    float s1[kPV], s2[kPV];
#pragma unroll
    for (int j = 0; j < kPV; ++j) { s1[j] = 0.f; s2[j] = 0.f; }

    uint4 nx0 = make_uint4(0, 0, 0, 0), nx1 = nx0;
    if (my_mask[0] & 1ull) {
        nx0 = __ldg(idx);
        nx1 = __ldg(idx + 1);
    }
    uint32_t parity = 0;
    for (int v0 = 0; v0 < a.nv; v0 += S) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int v = v0 + s;
            if (v < a.nv) {
                const bool active = (my_mask[v >> 6] >> (v & 63)) & 1ull;
                const uint4 c0 = nx0, c1 = nx1;
                const int vn = v + 1;
                if (vn < a.nv && ((my_mask[vn >> 6] >> (vn & 63)) & 1ull)) {
                    nx0 = __ldg(idx + vn * idx_pitch);
                    nx1 = __ldg(idx + vn * idx_pitch + 1);
                }
                mbar_wait(bar_full + 8 * s, parity);
                if (active) gather16<T>(sm_base + s * a.stage_bytes, c0, c1, s1, s2);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * s);
            }
        }
        parity ^= 1u;
    }
    if (!tile_ok) return;

    // ---------------- epilogue ----------------
    const int64_t row = (int64_t)c * a.n_vox;
    const bool vec_ok = (n0 + kPV <= a.n_vox) && ((row + n0) % 4 == 0);
#pragma unroll
    for (int g = 0; g < kPV / 4; ++g) {
        const int4 cn = __ldg(reinterpret_cast<const int4 *>(a.cnt + n0) + g);
        const int cnv[4] = {cn.x, cn.y, cn.z, cn.w};
        float oa[4], ob[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = 4 * g + j;
            const float cf = (float)cnv[j];
            if (kRaw) {
                oa[j] = s1[jj];
                ob[j] = s2[jj];
            } else if (cnv[j] > 0) {
                const float m = s1[jj] / cf;                                  // == S1 / (count + 1e-8) in fp32
                float ssd = fmaxf(fmaf(-m, s1[jj], s2[jj]), 0.0f);            // sum over valid views of (f - m)^2
                ssd = fmaf((float)(a.n_views_total - cnv[j]) * m, m, ssd);    // invalid views contribute m^2 each
                float al = 1.0f;
                if (a.alpha != nullptr && n0 + jj < a.n_vox) al = __ldg(a.alpha + n0 + jj);
                oa[j] = m * al;
                ob[j] = expf(-(ssd / cf));
            } else {
                oa[j] = 0.0f;                                                 // nerfdet.py:176
                ob[j] = 0.0f;                                                 // exp(-1e6) == 0 (nerfdet.py:180-181)
            }
        }
        const int64_t o = row + n0 + 4 * g;
        if (vec_ok) {
            __stcs(reinterpret_cast<float4 *>(a.out_a + o), make_float4(oa[0], oa[1], oa[2], oa[3]));
            if (a.out_b != nullptr) __stcs(reinterpret_cast<float4 *>(a.out_b + o), make_float4(ob[0], ob[1], ob[2], ob[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n0 + 4 * g + j < a.n_vox) {
                    a.out_a[o + j] = oa[j];
                    if (a.out_b != nullptr) a.out_b[o + j] = ob[j];
                }
            }
        }
        if (c == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t n = n0 + 4 * g + j;
                if (n < a.n_vox) {
                    if (a.count_i64 != nullptr) a.count_i64[n] = (int64_t)cnv[j];
                    if (a.count_f32 != nullptr) a.count_f32[n] = (float)cnv[j];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
struct PlaneGeom {
    int elt, n_pix, nw64, n_tiles, n_parts, warps, stages;
    int64_t n_pad;
    uint32_t plane_bytes, stage_bytes;
    size_t off_bytes, cnt_bytes, mask_bytes, total_bytes, smem_bytes;
};

static bool plane_geom(const nd_maps *f, int64_t n_vox, PlaneGeom &g) {
    g.elt = f->dtype == ND_F32 ? 4 : 2;
    g.n_pix = f->height * f->width;
    if (f->stride_x != 1 || f->stride_y != f->width) return false;          // planes must be contiguous
    const int64_t pb = (int64_t)g.n_pix * g.elt;
    if (pb % 16 != 0 || pb + 4 > 65535) return false;                       // TMA granule; uint16 byte offsets
    if ((reinterpret_cast<uintptr_t>(f->data) & 15) != 0 || (f->stride_v * g.elt) % 16 != 0 ||
        (f->stride_c * g.elt) % 16 != 0)
        return false;
    if (f->n_views > 64 * kPMaskWords || (size_t)f->n_views * 12 * sizeof(float) > 40 * 1024) return false;
    g.plane_bytes = (uint32_t)pb;
    g.stage_bytes = (uint32_t)align_up((size_t)pb + 16, 128);
    const size_t budget = 200 * 1024;
    g.stages = (int)(budget / g.stage_bytes);
    if (g.stages > kPMaxStages) g.stages = kPMaxStages;
    if (g.stages > f->n_views) g.stages = f->n_views < 2 ? 2 : f->n_views;
    if (g.stages < 2) return false;
    g.nw64 = (f->n_views + 63) / 64;
    g.n_tiles = (int)ceil_div(n_vox, kPTile);
    g.n_pad = (int64_t)g.n_tiles * kPTile;
    g.n_parts = (int)ceil_div(g.n_tiles, kPMaxWarps);
    g.warps = (int)ceil_div(g.n_tiles, g.n_parts);
    g.off_bytes = align_up((size_t)f->n_views * g.n_pad * sizeof(uint16_t), 256);
    g.cnt_bytes = align_up((size_t)g.n_pad * sizeof(int32_t), 256);
    g.mask_bytes = align_up((size_t)g.n_tiles * g.nw64 * sizeof(uint64_t), 256);
    g.total_bytes = g.off_bytes + g.cnt_bytes + g.mask_bytes;
    g.smem_bytes = (size_t)g.stages * g.stage_bytes + 2 * g.stages * 8 + (size_t)g.warps * g.nw64 * 8;
    return true;
}

bool lift_planes_eligible(const nd_maps *f, int64_t n_vox) {
    PlaneGeom g;
    return n_vox > 0 && plane_geom(f, n_vox, g);
}

size_t lift_planes_workspace_bytes(const nd_maps *f, int64_t n_vox) {
    PlaneGeom g;
    if (!plane_geom(f, n_vox, g)) return 0;
    return g.total_bytes;
}

template <typename T, int S, bool kRaw>
static nd_status launch_planes(const PlaneArgs &a, const PlaneGeom &g, int channels, cudaStream_t st) {
    auto kern = k_lift_planes<T, S, kRaw>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    if (e != cudaSuccess) {
        set_error("k_lift_planes: cannot reserve %zu bytes of shared memory: %s", g.smem_bytes, cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    const unsigned grid = (unsigned)channels * (unsigned)g.n_parts;
    kern<<<grid, (g.warps + 1) * 32, g.smem_bytes, st>>>(a);
    ND_CUDA_LAUNCH_CHECK("k_lift_planes");
    return ND_OK;
}

template <typename T, bool kRaw>
nd_status run_lift_planes(const nd_maps *f, const float *points, const float *proj, int64_t n_vox, const float *alpha,
                          float *out_a, float *out_b, int64_t *count_i64, float *count_f32, void *ws, size_t ws_bytes,
                          cudaStream_t st) {
    PlaneGeom g;
    ND_REQUIRE(plane_geom(f, n_vox, g), ND_ERR_BAD_ARG, "lift: input not eligible for the plane-resident path");
    ND_REQUIRE(ws != nullptr && ws_bytes >= g.total_bytes, ND_ERR_WORKSPACE, "lift: workspace too small (%zu < %zu bytes)",
               ws_bytes, g.total_bytes);
    ND_REQUIRE((reinterpret_cast<uintptr_t>(ws) % 256) == 0, ND_ERR_BAD_ALIGNMENT, "lift: workspace not 256-byte aligned");
    char *wsb = reinterpret_cast<char *>(ws);
    uint16_t *off16 = reinterpret_cast<uint16_t *>(wsb);
    int32_t *cnt = reinterpret_cast<int32_t *>(wsb + g.off_bytes);
    uint64_t *vmask = reinterpret_cast<uint64_t *>(wsb + g.off_bytes + g.cnt_bytes);

    const size_t idx_smem = (size_t)f->n_views * 12 * sizeof(float);
    k_plane_index<<<(unsigned)g.n_tiles, kPTile, idx_smem, st>>>(points, proj, f->n_views, g.nw64, n_vox, g.n_pad,
                                                                 f->height, f->width, g.elt, g.plane_bytes, off16, cnt,
                                                                 vmask);
    ND_CUDA_LAUNCH_CHECK("k_plane_index");

    PlaneArgs a{};
    a.off16 = off16;
    a.cnt = cnt;
    a.vmask = vmask;
    a.nv = f->n_views;
    a.nw64 = g.nw64;
    a.n_vox = n_vox;
    a.n_pad = g.n_pad;
    a.n_tiles = g.n_tiles;
    a.n_parts = g.n_parts;
    a.feat = f->data;
    a.sv = f->stride_v;
    a.sc = f->stride_c;
    a.plane_bytes = g.plane_bytes;
    a.stage_bytes = g.stage_bytes;
    a.n_views_total = f->n_views;
    a.alpha = alpha;
    a.out_a = out_a;
    a.out_b = out_b;
    a.count_i64 = count_i64;
    a.count_f32 = count_f32;
    switch (g.stages) {
        case 2: return launch_planes<T, 2, kRaw>(a, g, f->channels, st);
        case 3: return launch_planes<T, 3, kRaw>(a, g, f->channels, st);
        default: return launch_planes<T, 4, kRaw>(a, g, f->channels, st);
    }
}

#define ND_INSTANTIATE_PLANES(T, R)                                                                                   \
    template nd_status run_lift_planes<T, R>(const nd_maps *, const float *, const float *, int64_t, const float *,  \
                                             float *, float *, int64_t *, float *, void *, size_t, cudaStream_t);
ND_INSTANTIATE_PLANES(float, false)
ND_INSTANTIATE_PLANES(float, true)
ND_INSTANTIATE_PLANES(__nv_bfloat16, false)
ND_INSTANTIATE_PLANES(__nv_bfloat16, true)

}  // namespace nd
