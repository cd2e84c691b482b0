// Plane-resident fused lift for the reference's own layout (NCHW maps whose sliced planes
// are contiguous, fp32 or bf16).  Replaces reference nerfdet.py:164-181.
//
// Observation: both sides of the lift are CHANNEL-major -- input planes feat[v][c][pixel] and
// output rows mean/cov[c][voxel] -- so no transposition is needed, only a fast random-access
// memory for one (view, channel) plane (18.9 KB at 59x80 fp32).  That memory is shared memory.
//
//   Tiling          a RUN is 16 consecutive voxels (one lane's work, 64 B of an output row); a TILE
//                   is 32 runs (one warp's work).  When the caller passes the grid shape and
//                   Z % 16 == 0, X % 4 == 0, Y % 8 == 0 a tile is a compact 4 x 8 block of Z-runs,
//                   so that whole tiles fall outside a camera frustum; otherwise tiles are 512
//                   consecutive voxels.  Tables below are stored in tile order ("positions").
//   k_plane_index   one pass over (position, view): the bit-exact nearest-pixel projection, stored
//                   as a uint16 BYTE OFFSET into a plane (invalid -> offset of a zero word behind
//                   the plane); per-voxel partial view counts (uint8 per 16-view group); and per
//                   (tile, view) a 4-bit mask saying which quarter of the runs (4 voxels of every
//                   lane) has any valid voxel.
//   k_lift_planes   work unit = (channel c, part p of the tiles).  A producer warp streams the nv
//                   planes of channel c through an S-stage mbarrier ring with TMA bulk copies (each
//                   plane byte leaves HBM once).  Each compute warp owns one tile and keeps sum /
//                   sum-of-squares of its 16 voxels per lane in 32 registers across all views; per
//                   view it loads its 16 offsets (32 B per lane, 1 KB contiguous per warp, L2
//                   resident), gathers the unmasked quarters from the plane in shared memory and
//                   accumulates with packed f32x2 adds / FMAs.  The epilogue turns the
//                   accumulators into mean / exp(-var) (or raw S1 / S2 for the view-sharded path).
//
// Nothing of size [nv][C][N] (the reference's 1.3 GB volume) or [nv][pixel][C] (a pixel-major
// staging copy) is ever written.
#include <stdlib.h>

#include "nd_common.cuh"

namespace nd {

constexpr int kPV = 16;                 // voxels per lane (one run)
constexpr int kPTile = 32 * kPV;        // voxels per warp tile
constexpr int kPMaxWarps = 25;          // compute warps per CTA (+1 producer warp)
constexpr int kPMaxViews = 512;
constexpr int kPMaxStages = 16;
constexpr int kPListPitch = 256;        // entries per warp in the active-view list (nv <= 255)
constexpr int kPRowDepth = 2;           // per-warp ring of offset rows (cp.async prefetch distance 2 views)
constexpr int kPBx = 4, kPBy = 8;       // compact tile = kPBx x kPBy runs in (x, y)

// run index (first voxel / 16) of lane `lane` of tile `t`
struct Tiling {
    int compact;            // 0: tile = 32 consecutive runs
    int gy, nzr, tiles_y;   // compact: grid Y, runs per Z column, tiles along Y
    __host__ __device__ __forceinline__ int64_t run(int t, int lane) const {
        if (!compact) return (int64_t)t * 32 + lane;
        const int zr = t % nzr, txy = t / nzr;
        const int ty = txy % tiles_y, tx = txy / tiles_y;
        const int ix = tx * kPBx + (lane >> 3), iy = ty * kPBy + (lane & 7);
        return ((int64_t)ix * gy + iy) * nzr + zr;
    }
};

struct PlaneArgs {
    Tiling tiling;
    // geometry tables (workspace), all in position order p = tile * 512 + lane * 16 + j
    const uint16_t *off16;     // [nv][n_pad]
    const uint8_t *cnt8;       // [nw16][n_pad] partial view counts
    const uint64_t *vmask;     // [n_tiles][nw16], 4 bits per view
    int nv, nw16;
    int64_t n_vox, n_pad;
    int n_tiles, n_parts, n_units;   // unit = (channel, part of the tiles); unit u -> c = u / n_parts
    // planes
    const void *feat;
    int64_t sv, sc;            // elements
    uint32_t plane_bytes, plane_pitch;   // smem slot = plane + zero word, padded to plane_pitch
    int l2_ahead;              // views the producer's L2 prefetch runs ahead of the shared-memory fill
    int stages, group;         // ring of `stages` stages, `group` consecutive views per stage
    int *trace;                // diagnostics: per-warp per-stage clock trace of CTA 0 (tools/lift_trace.py), or null
    int debug;                 // diagnostics (tools/lift_probe.py): 1 = no gather, 4 = no plane copies, 16 = no epilogue
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
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(2000u)   // suspend-time hint (ns): fewer wake-ups of idle warps
            : "memory");
    } while (!done);
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
template <typename T> __device__ __forceinline__ float lds_elt(uint32_t addr);
template <> __device__ __forceinline__ float lds_elt<float>(uint32_t addr) { return lds_f32(addr); }
template <> __device__ __forceinline__ float lds_elt<__nv_bfloat16>(uint32_t addr) { return lds_bf16(addr); }

// packed fp32 pair accumulate: s1 += f, s2 += f * f  (FADD2 / FFMA2 on sm_100)
__device__ __forceinline__ void acc2(unsigned long long &s1, unsigned long long &s2, float fa, float fb) {
    unsigned long long f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(f) : "f"(fa), "f"(fb));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s1) : "l"(f));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s2) : "l"(f));
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// ---------------------------------------------------------------------------------------------
// Geometry tables.  grid = (tiles, 16-view groups); one thread per position.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPTile)
k_plane_index(const Tiling tiling, const float *__restrict__ points, const float *__restrict__ proj, int nv, int nw16,
              int64_t n_vox, int64_t n_pad, int height, int width, int elt, uint32_t zero_off,
              uint16_t *__restrict__ off16, uint8_t *__restrict__ cnt8, uint64_t *__restrict__ vmask) {
    __shared__ float sp[16 * 12];
    __shared__ unsigned long long smask;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let k_lift_planes start streaming planes
    const int v0 = blockIdx.y * 16;
    const int nvg = min(16, nv - v0);
    if (threadIdx.x < nvg * 12) sp[threadIdx.x] = proj[v0 * 12 + threadIdx.x];
    if (threadIdx.x == 0) smask = 0ull;
    __syncthreads();
    const int lane_slot = threadIdx.x >> 4, j = threadIdx.x & 15;
    const int64_t n = tiling.run(blockIdx.x, lane_slot) * kPV + j;
    const int64_t p = (int64_t)blockIdx.x * kPTile + threadIdx.x;
    const bool inside = n < n_vox;
    float X = 0.f, Y = 0.f, Z = 0.f;
    if (inside) {
        X = __ldg(points + n);
        Y = __ldg(points + n_vox + n);
        Z = __ldg(points + 2 * n_vox + n);
    }
    int count = 0;
    unsigned long long m = 0ull;
#pragma unroll 4
    for (int i = 0; i < nvg; ++i) {
        float xr, yr, q2;
        const bool ok = project_nearest(sp + i * 12, X, Y, Z, height, width, xr, yr, q2) && inside;
        const uint32_t off = ok ? (uint32_t)((int)yr * width + (int)xr) * (uint32_t)elt : zero_off;
        off16[(int64_t)(v0 + i) * n_pad + p] = (uint16_t)off;
        count += ok ? 1 : 0;
        // warp = 2 lane slots x 16 voxels; quarter g of a run = voxels 4g..4g+3
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        unsigned nib = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) nib |= (b & (0x000f000fu << (4 * g))) ? (1u << g) : 0u;
        m |= (unsigned long long)nib << (4 * i);
    }
    if ((threadIdx.x & 31) == 0 && m != 0ull) atomicOr(&smask, m);
    cnt8[(int64_t)blockIdx.y * n_pad + p] = (uint8_t)count;
    __syncthreads();
    if (threadIdx.x == 0) vmask[(int64_t)blockIdx.x * nw16 + blockIdx.y] = smask;
}

// ---------------------------------------------------------------------------------------------
// Gather + statistics with the planes of one channel streamed through shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void gather_quarter(uint32_t sb, uint32_t w0, uint32_t w1, unsigned long long &s1a,
                                               unsigned long long &s2a, unsigned long long &s1b,
                                               unsigned long long &s2b) {
    const float f0 = lds_elt<T>(sb + (w0 & 0xffffu));
    const float f1 = lds_elt<T>(sb + (w0 >> 16));
    const float f2 = lds_elt<T>(sb + (w1 & 0xffffu));
    const float f3 = lds_elt<T>(sb + (w1 >> 16));
    acc2(s1a, s2a, f0, f1);
    acc2(s1b, s2b, f2, f3);
}

template <typename T, bool kRaw>
__global__ void __launch_bounds__((kPMaxWarps + 1) * 32, 1)
k_lift_planes(const PlaneArgs a) {
    const int S = a.stages, G = a.group;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = (blockDim.x >> 5) - 1;                           // compute warps
    // persistent CTA: units blockIdx.x, blockIdx.x + gridDim.x, ...; gridDim.x is a multiple of n_parts,
    // so the part (and with it every tile-dependent table) is the same for all units of a CTA
    const int part = blockIdx.x % a.n_parts;

    // ring: S stages x G plane slots; then the barriers, the offset rings and the active-view lists
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)S * G * a.plane_pitch);
    unsigned char *s_rings = smem + (((size_t)S * G * a.plane_pitch + 2 * S * 8 + 15) & ~(size_t)15);   // [W][depth][1 KB]
    unsigned char *s_lists = s_rings + (size_t)W * kPRowDepth * (kPTile * 2);                              // [W][kPListPitch] u16
    const uint32_t sm_base = smem_u32(smem);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + S);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, (uint32_t)W);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < S * G; i += blockDim.x)          // the zero word behind every plane slot
        *reinterpret_cast<uint32_t *>(smem + (size_t)i * a.plane_pitch + a.plane_bytes) = 0u;
    __syncthreads();

    const long long t_cta = clock64();
    if (warp == W) {
        // ---------------- producer: one elected lane streams the planes of this CTA's units ----------------
        // (the planes are kernel inputs, not products of k_plane_index: no dependency wait here)
        if (lane == 0) {
            const int64_t view_bytes = a.sv * (int64_t)sizeof(T);
            int s = 0, n_tr = 0;
            uint32_t parity = 1;                                   // first pass over the ring: slots are free
            bool first_pass = true;
            // L2 prefetch runs `a.l2_ahead` views ahead of the shared-memory fill (it costs no shared memory and
            // turns the HBM latency of the fill into an L2 hit); pu / pv walk the same (unit, view) sequence
            int pu = blockIdx.x, pv = 0;
            auto l2_step = [&]() {
                if (pu < a.n_units) {
                    bulk_prefetch_l2(reinterpret_cast<const char *>(a.feat) +
                                         ((int64_t)(pu / a.n_parts) * a.sc + (int64_t)pv * a.sv) * (int64_t)sizeof(T),
                                     a.plane_bytes);
                    if (++pv == a.nv) { pv = 0; pu += gridDim.x; }
                }
            };
            for (int i = 0; i < a.l2_ahead; ++i) l2_step();
            for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
                const int c = u / a.n_parts;
                const char *src = reinterpret_cast<const char *>(a.feat) + (int64_t)c * a.sc * (int64_t)sizeof(T);
                for (int v0 = 0; v0 < a.nv; v0 += G) {
                    const int g_n = min(G, a.nv - v0);
                    if (a.l2_ahead > 0)
                        for (int g = 0; g < g_n; ++g) l2_step();
                    const long long tp0 = clock64();
                    if (!first_pass) mbar_wait(bar_empty + 8 * s, parity);
                    if (a.trace != nullptr && blockIdx.x == 0 && n_tr < 256) {
                        int *t = a.trace + ((size_t)W * 256 + n_tr) * 4;
                        t[0] = (int)(tp0 - t_cta);
                        t[1] = (int)(clock64() - t_cta);
                        t[2] = 0;
                        t[3] = 0;
                        ++n_tr;
                    }
                    if (a.debug & 4) {
                        mbar_arrive(bar_full + 8 * s);
                    } else {
                        mbar_expect_tx(bar_full + 8 * s, (uint32_t)g_n * a.plane_bytes);
                        for (int g = 0; g < g_n; ++g)
                            bulk_g2s(sm_base + (uint32_t)(s * G + g) * a.plane_pitch, src + (v0 + g) * view_bytes,
                                     a.plane_bytes, bar_full + 8 * s);
                    }
                    if (++s == S) { s = 0; parity ^= 1u; first_pass = false; }
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    asm volatile("griddepcontrol.wait;" ::: "memory");             // tables come from k_plane_index (PDL)
    const int tile = part * W + warp;
    const bool tile_ok = tile < a.n_tiles;
    const int64_t p0 = (int64_t)(tile_ok ? tile : 0) * kPTile + lane * kPV;
    const char *idx = reinterpret_cast<const char *>(a.off16 + p0);
    const uint32_t idx_pitch = (uint32_t)(a.n_pad * (int64_t)sizeof(uint16_t));   // bytes per view row
    const int64_t n0 = a.tiling.run(tile_ok ? tile : 0, lane) * kPV;
    const bool lane_ok = tile_ok && n0 < a.n_vox;

    // view counts of this lane's 16 voxels (sum of the uint8 partials), packed 4 per word
    uint4 cw = make_uint4(0, 0, 0, 0);
    for (int g = 0; g < a.nw16; ++g) {
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(a.cnt8 + (int64_t)g * a.n_pad + p0));
        cw.x = __vadd4(cw.x, t.x); cw.y = __vadd4(cw.y, t.y); cw.z = __vadd4(cw.z, t.z); cw.w = __vadd4(cw.w, t.w);
    }

    // The views that see this warp's tile, as a list (view | quarter mask << 8): views the tile
    // does not see cost the warp nothing but the stage hand-shake.
    uint16_t *act = reinterpret_cast<uint16_t *>(s_lists) + warp * kPListPitch;
    int n_act = 0;
    for (int vb = 0; vb < a.nv; vb += 32) {
        const int v = vb + lane;
        uint32_t nib = 0;
        if (tile_ok && v < a.nv) nib = (uint32_t)(__ldg(a.vmask + (int64_t)tile * a.nw16 + (v >> 4)) >> ((v & 15) * 4)) & 0xfu;
        const unsigned b = __ballot_sync(0xffffffffu, nib != 0);
        if (nib) act[n_act + __popc(b & ((1u << lane) - 1u))] = (uint16_t)(v | (nib << 8));
        n_act += __popc(b);
    }
    if (lane == 0) act[n_act] = 0xffffu;                           // sentinel: view 255 is never reached
    __syncwarp();

    // this lane's 32 B of offsets per list entry travel through a warp-private ring in shared
    // memory (cp.async, kPRowDepth - 1 list entries ahead, across unit boundaries)
    const uint32_t ring = smem_u32(s_rings) + (uint32_t)(warp * kPRowDepth) * (kPTile * 2) + lane * (kPV * 2);
    int kf = 0;                                                    // next list entry to prefetch
    auto fetch = [&](int slot) {
        if (n_act > 0) {
            const uint32_t v = act[kf] & 0xffu;
            cp_async16(ring + slot * (kPTile * 2), idx + v * idx_pitch);
            cp_async16(ring + slot * (kPTile * 2) + 16, idx + v * idx_pitch + 16);
            if (++kf == n_act) kf = 0;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < kPRowDepth - 1; ++i) fetch(i);

    int s = 0, d = 0, n_tr = 0;
    uint32_t parity = 0;
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        const int c = u / a.n_parts;
        unsigned long long s1[kPV / 2], s2[kPV / 2];
#pragma unroll
        for (int j = 0; j < kPV / 2; ++j) { s1[j] = 0ull; s2[j] = 0ull; }
        int k = 0;
        uint32_t ent = act[0];

        for (int v0 = 0; v0 < a.nv; v0 += G) {
            const long long tc0 = a.trace != nullptr ? clock64() : 0;
            mbar_wait(bar_full + 8 * s, parity);
            const long long tc1 = a.trace != nullptr ? clock64() : 0;
            const int k_before = k;
            const uint32_t v_end = (uint32_t)min(v0 + G, a.nv);
            while ((ent & 0xffu) < v_end) {                        // warp-uniform: the list is per warp
                fetch(d == 0 ? kPRowDepth - 1 : d - 1);            // refill the slot consumed last
                cp_async_wait<kPRowDepth - 1>();                   // this entry's offsets have landed
                if (!(a.debug & 1)) {
                    const uint32_t sb = sm_base + (uint32_t)(s * G + (int)(ent & 0xffu) - v0) * a.plane_pitch;
                    const uint4 c0 = lds_u4(ring + d * (kPTile * 2)), c1 = lds_u4(ring + d * (kPTile * 2) + 16);
                    if (ent & 0x100u) gather_quarter<T>(sb, c0.x, c0.y, s1[0], s2[0], s1[1], s2[1]);
                    if (ent & 0x200u) gather_quarter<T>(sb, c0.z, c0.w, s1[2], s2[2], s1[3], s2[3]);
                    if (ent & 0x400u) gather_quarter<T>(sb, c1.x, c1.y, s1[4], s2[4], s1[5], s2[5]);
                    if (ent & 0x800u) gather_quarter<T>(sb, c1.z, c1.w, s1[6], s2[6], s1[7], s2[7]);
                }
                if (++d == kPRowDepth) d = 0;
                ent = act[++k];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);
            if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && n_tr < 256) {
                int *t = a.trace + ((size_t)warp * 256 + n_tr) * 4;
                t[0] = (int)(tc0 - t_cta);
                t[1] = (int)(tc1 - t_cta);
                t[2] = (int)(clock64() - t_cta);
                t[3] = k - k_before;
                ++n_tr;
            }
            if (++s == S) { s = 0; parity ^= 1u; }
        }
        if (!lane_ok || (a.debug & 16)) continue;

        // ---------------- epilogue of unit (c, part): the producer is already streaming the next unit ----------------
        const int64_t row = (int64_t)c * a.n_vox;
        const bool vec_ok = (n0 + kPV <= a.n_vox) && ((row + n0) % 4 == 0) &&
                            ((reinterpret_cast<uintptr_t>(a.out_a) | reinterpret_cast<uintptr_t>(a.out_b)) % 16 == 0);
        const uint32_t cws[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
        for (int g = 0; g < kPV / 4; ++g) {
            int cnv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) cnv[j] = (int)((cws[g] >> (8 * j)) & 0xffu);
            const float2 a0 = unpack2(s1[2 * g]), a1 = unpack2(s1[2 * g + 1]);
            const float2 b0 = unpack2(s2[2 * g]), b1 = unpack2(s2[2 * g + 1]);
            const float v1[4] = {a0.x, a0.y, a1.x, a1.y};
            const float v2[4] = {b0.x, b0.y, b1.x, b1.y};
            float oa[4], ob[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float cf = (float)cnv[j];
                if (kRaw) {
                    oa[j] = v1[j];
                    ob[j] = v2[j];
                } else if (cnv[j] > 0) {
                    const float m = v1[j] / cf;                                   // == S1 / (count + 1e-8) in fp32
                    float ssd = fmaxf(fmaf(-m, v1[j], v2[j]), 0.0f);              // sum over valid views of (f - m)^2
                    ssd = fmaf((float)(a.n_views_total - cnv[j]) * m, m, ssd);    // invalid views contribute m^2 each
                    float al = 1.0f;
                    if (a.alpha != nullptr && n0 + 4 * g + j < a.n_vox) al = __ldg(a.alpha + n0 + 4 * g + j);
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
                if (a.out_b != nullptr)
                    __stcs(reinterpret_cast<float4 *>(a.out_b + o), make_float4(ob[0], ob[1], ob[2], ob[3]));
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
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
static int *g_trace = nullptr;   // device buffer [(warps + 1)][256][4] int32 set by nd_debug_set_trace (tools only)
void set_lift_trace(int *buf) { g_trace = buf; }

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
struct PlaneGeom {
    Tiling tiling;
    int elt, n_pix, nw16, n_tiles, n_parts, warps, stages, group, grid;
    int64_t n_pad;
    uint32_t plane_bytes, plane_pitch;
    size_t off_bytes, cnt_bytes, mask_bytes, total_bytes, smem_bytes;
};

static bool plane_geom(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt, PlaneGeom &g) {
    g.elt = f->dtype == ND_F32 ? 4 : 2;
    g.n_pix = f->height * f->width;
    if (f->stride_x != 1 || f->stride_y != f->width) return false;          // planes must be contiguous
    const int64_t pb = (int64_t)g.n_pix * g.elt;
    if (pb % 16 != 0 || pb + 4 > 65535) return false;                       // TMA granule; uint16 byte offsets
    if ((reinterpret_cast<uintptr_t>(f->data) & 15) != 0 || (f->stride_v * g.elt) % 16 != 0 ||
        (f->stride_c * g.elt) % 16 != 0)
        return false;
    if (f->n_views > 255) return false;                                     // uint8 view counts
    g.plane_bytes = (uint32_t)pb;
    g.nw16 = (f->n_views + 15) / 16;
    g.n_tiles = (int)ceil_div(n_vox, kPTile);
    g.n_pad = (int64_t)g.n_tiles * kPTile;
    g.tiling = Tiling{0, 0, 0, 0};
    if (opt != nullptr && opt->grid_x > 0 && opt->grid_y > 0 && opt->grid_z > 0 &&
        (int64_t)opt->grid_x * opt->grid_y * opt->grid_z == n_vox && opt->grid_z % kPV == 0 &&
        opt->grid_x % kPBx == 0 && opt->grid_y % kPBy == 0) {
        g.tiling.compact = 1;
        g.tiling.gy = opt->grid_y;
        g.tiling.nzr = opt->grid_z / kPV;
        g.tiling.tiles_y = opt->grid_y / kPBy;
    }
    int max_warps = kPMaxWarps, stages = 2, group = 4;
    if (const char *e = getenv("ND_LIFT_STAGES")) stages = atoi(e);         // tuning knobs for tools/lift_probe.py
    if (const char *e = getenv("ND_LIFT_GROUP")) group = atoi(e);
    if (const char *e = getenv("ND_LIFT_WARPS")) {
        const int v = atoi(e);
        if (v >= 1 && v < max_warps) max_warps = v;
    }
    g.n_parts = (int)ceil_div(g.n_tiles, max_warps);
    g.warps = (int)ceil_div(g.n_tiles, g.n_parts);
    g.plane_pitch = (uint32_t)align_up((size_t)pb + 16, 128);
    const size_t fixed = 2 * kPMaxStages * 8 + 16 + (size_t)g.warps * (kPRowDepth * kPTile * 2 + kPListPitch * 2);
    const int slots = (int)(((size_t)(224 * 1024) - fixed) / g.plane_pitch);   // plane slots that fit one SM
    if (slots < 2) return false;
    if (group < 1) group = 1;
    if (stages < 2) stages = 2;
    if (stages > kPMaxStages) stages = kPMaxStages;
    while (stages * group > slots) {                                        // shrink the ring to what fits
        if (group > 1 && (group >= stages || stages == 2)) --group; else --stages;
    }
    g.stages = stages;
    g.group = group;
    // persistent grid: one CTA per SM, a multiple of n_parts (see k_lift_planes)
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int64_t n_units = (int64_t)f->channels * g.n_parts;
    int64_t grid = n_units < sms ? n_units : sms;
    grid -= grid % g.n_parts;
    if (grid < g.n_parts) grid = g.n_parts;
    g.grid = (int)grid;
    g.off_bytes = align_up((size_t)f->n_views * g.n_pad * sizeof(uint16_t), 256);
    g.cnt_bytes = align_up((size_t)g.nw16 * g.n_pad, 256);
    g.mask_bytes = align_up((size_t)g.n_tiles * g.nw16 * sizeof(uint64_t), 256);
    g.total_bytes = g.off_bytes + g.cnt_bytes + g.mask_bytes;
    g.smem_bytes = (size_t)g.stages * g.group * g.plane_pitch + 2 * g.stages * 8 + 16 +
                   (size_t)g.warps * (kPRowDepth * kPTile * 2 + kPListPitch * 2);
    return true;
}

bool lift_planes_eligible(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt) {
    PlaneGeom g;
    return n_vox > 0 && plane_geom(f, n_vox, opt, g);
}

size_t lift_planes_workspace_bytes(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt) {
    PlaneGeom g;
    if (!plane_geom(f, n_vox, opt, g)) return 0;
    return g.total_bytes;
}

template <typename T, bool kRaw>
static nd_status launch_planes(const PlaneArgs &a, const PlaneGeom &g, int channels, cudaStream_t st) {
    auto kern = k_lift_planes<T, kRaw>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    if (e != cudaSuccess) {
        set_error("k_lift_planes: cannot reserve %zu bytes of shared memory: %s", g.smem_bytes, cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    // programmatic dependent launch: prologue and plane streaming overlap the tail of k_plane_index;
    // the consumers execute griddepcontrol.wait before touching its tables
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)g.grid);
    cfg.blockDim = dim3((unsigned)(g.warps + 1) * 32);
    cfg.dynamicSmemBytes = g.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess) {
        set_error("k_lift_planes: CUDA error %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    return ND_OK;
}

template <typename T, bool kRaw>
nd_status run_lift_planes(const nd_maps *f, const float *points, const float *proj, int64_t n_vox, const float *alpha,
                          float *out_a, float *out_b, int64_t *count_i64, float *count_f32, void *ws, size_t ws_bytes,
                          const nd_lift_options *opt, cudaStream_t st) {
    PlaneGeom g;
    ND_REQUIRE(plane_geom(f, n_vox, opt, g), ND_ERR_BAD_ARG, "lift: input not eligible for the plane-resident path");
    ND_REQUIRE(ws != nullptr && ws_bytes >= g.total_bytes, ND_ERR_WORKSPACE, "lift: workspace too small (%zu < %zu bytes)",
               ws_bytes, g.total_bytes);
    ND_REQUIRE((reinterpret_cast<uintptr_t>(ws) % 256) == 0, ND_ERR_BAD_ALIGNMENT, "lift: workspace not 256-byte aligned");
    char *wsb = reinterpret_cast<char *>(ws);
    uint16_t *off16 = reinterpret_cast<uint16_t *>(wsb);
    uint8_t *cnt8 = reinterpret_cast<uint8_t *>(wsb + g.off_bytes);
    uint64_t *vmask = reinterpret_cast<uint64_t *>(wsb + g.off_bytes + g.cnt_bytes);

    k_plane_index<<<dim3((unsigned)g.n_tiles, (unsigned)g.nw16), kPTile, 0, st>>>(
        g.tiling, points, proj, f->n_views, g.nw16, n_vox, g.n_pad, f->height, f->width, g.elt, g.plane_bytes, off16,
        cnt8, vmask);
    ND_CUDA_LAUNCH_CHECK("k_plane_index");

    PlaneArgs a{};
    a.tiling = g.tiling;
    a.off16 = off16;
    a.cnt8 = cnt8;
    a.vmask = vmask;
    a.nv = f->n_views;
    a.nw16 = g.nw16;
    a.n_vox = n_vox;
    a.n_pad = g.n_pad;
    a.n_tiles = g.n_tiles;
    a.n_parts = g.n_parts;
    a.feat = f->data;
    a.sv = f->stride_v;
    a.sc = f->stride_c;
    a.plane_bytes = g.plane_bytes;
    a.plane_pitch = g.plane_pitch;
    a.n_units = f->channels * g.n_parts;
    a.group = g.group;
    a.l2_ahead = 0;
    a.trace = g_trace;
    if (const char *e = getenv("ND_LIFT_L2AHEAD")) a.l2_ahead = atoi(e);
    a.n_views_total = f->n_views;
    a.alpha = alpha;
    a.out_a = out_a;
    a.out_b = out_b;
    a.count_i64 = count_i64;
    a.count_f32 = count_f32;
    a.stages = g.stages;
    if (const char *e = getenv("ND_LIFT_DEBUG")) a.debug = atoi(e);
    return launch_planes<T, kRaw>(a, g, f->channels, st);
}

#define ND_INSTANTIATE_PLANES(T, R)                                                                                   \
    template nd_status run_lift_planes<T, R>(const nd_maps *, const float *, const float *, int64_t, const float *,  \
                                             float *, float *, int64_t *, float *, void *, size_t,                   \
                                             const nd_lift_options *, cudaStream_t);
ND_INSTANTIATE_PLANES(float, false)
ND_INSTANTIATE_PLANES(float, true)
ND_INSTANTIATE_PLANES(__nv_bfloat16, false)
ND_INSTANTIATE_PLANES(__nv_bfloat16, true)

}  // namespace nd
