// Plane-resident fused lift for the reference's own layout (NCHW maps whose sliced planes
// are contiguous, fp32 or bf16).  Replaces reference nerfdet.py:164-181.
//
// Observation: both sides of the lift are CHANNEL-major -- input planes feat[v][c][pixel] and
// output rows mean/cov[c][voxel] -- so no transposition is needed, only a fast random-access
// memory for one (view, channel) plane (18.9 KB at 59x80 fp32).  That memory is shared memory.
//
//   Octs            an OCT is 32 lanes x 8 consecutive voxels (256 voxels, 32 B of an output row per
//                   lane); its two QUADS (4 voxels per lane) are the unit of frustum culling.  When the
//                   caller passes the grid shape and Z % 8 == 0, X % 4 == 0, Y % 8 == 0 an oct is a
//                   compact 4 x 8 x 8 block of voxels, so that whole octs fall outside a camera frustum;
//                   otherwise octs are 256 consecutive voxels.
//   k_plane_index   one pass over (position, view): the bit-exact nearest-pixel projection, stored
//                   as a uint16 BYTE OFFSET into a plane (invalid -> offset of a zero word behind
//                   the plane), one 512 B row per (view, oct); per-voxel partial view counts (uint8 per
//                   16-view group); per (oct, view) a 2-bit mask saying which quad has any valid voxel,
//                   and the number of active quad-views of every oct (its cost).
//   k_plane_pack    (every block, redundantly) ranks the octs by cost and pairs the k-th most with the k-th least expensive one
//                   (one pair per compute warp, so that all warps of a CTA carry the same load); a PART
//                   is the set of pairs one CTA owns.  Per (part, view) the offset rows of the octs
//                   that see the view are compacted into one contiguous block, so that the lift kernel
//                   can fetch them with a single bulk copy, and every warp gets its entry
//                   (view | quad mask | row slots).
//   k_lift_planes   work unit = (channel c, part p).  A producer warp streams, per view, the plane of
//                   channel c AND the part's offset block through an S-stage mbarrier ring with TMA
//                   bulk copies (each plane byte leaves HBM once; the consumers never touch global
//                   memory inside the loop).  Each compute warp keeps sum / sum-of-squares of its
//                   16 voxels per lane in 32 registers across all views; per view that sees it, it
//                   reads its offsets (16 B per lane and oct) from the stage, gathers the active quads
//                   from the plane in shared memory and accumulates with packed f32x2 adds / FMAs.
//                   The epilogue turns the accumulators into mean / exp(-var) (or raw S1 / S2 for
//                   the view-sharded path).
//
// Nothing of size [nv][C][N] (the reference's 1.3 GB volume) or [nv][pixel][C] (a pixel-major
// staging copy) is ever written.
#include <stdlib.h>

#include <algorithm>

#include "nd_common.cuh"

namespace nd {

constexpr int kOV = 8;                  // voxels per lane and oct
constexpr int kOct = 32 * kOV;          // voxels per oct
constexpr int kPV = 2 * kOV;            // voxels per lane (two octs per warp)
constexpr int kRowBytes = kOct * 2;     // one offset row: uint16 per voxel of an oct
constexpr int kPMaxWarps = 25;          // compute warps per CTA (+1 producer warp)
constexpr int kPMaxStages = 16;
constexpr int kPBx = 4, kPBy = 8;       // compact oct = kPBx x kPBy columns in (x, y) x kOV in z
constexpr int kIdxViews = 4;            // views per block of k_plane_index (more, smaller blocks: the pass is latency bound)
constexpr int kPMaxOcts = 256;          // cost-balanced oct pairing up to this many octs (else index order)

struct Tiling {
    int compact;            // 0: oct = 256 consecutive voxels
    int gy, gz, nzo, tiles_y;   // compact: grid Y and Z, octs per Z column, 4x8 blocks along Y
    // first voxel of the 8 consecutive voxels of lane `lane` in oct `q`
    __host__ __device__ __forceinline__ int64_t voxel(int q, int lane) const {
        if (!compact) return (int64_t)q * kOct + lane * kOV;
        const int o = q % nzo, b = q / nzo;
        const int ty = b % tiles_y, tx = b / tiles_y;
        const int ix = tx * kPBx + (lane >> 3), iy = ty * kPBy + (lane & 7);
        return ((int64_t)ix * gy + iy) * gz + o * kOV;
    }
};

struct PlaneArgs {
    Tiling tiling;
    // tables (workspace)
    const uint8_t *cnt8;       // [nw16][n_pad] partial view counts per group of kIdxViews views, position p = oct * 256 + lane * 8 + i
    const uint16_t *offc;      // [n_parts][part_rows][256] offset rows of the octs that see a view, view after view in ring order
    int64_t part_rows;
    const uint16_t *rowcnt;    // [n_parts][nvp] rows in each block
    const uint32_t *ents;      // [n_parts][W][nvp] view | quad mask << 8 | slot A << 16 | slot B << 24 (0: not seen)
    const uint16_t *pairs;     // [n_parts][W][2] the octs of every warp (0xffff: none)
    int nv, nvp, nw16;
    int64_t n_vox, n_pad;
    int n_octs, n_parts, n_units;   // unit = (channel, part); unit u -> c = u / n_parts
    // planes
    const void *feat;
    int64_t sv, sc;            // elements
    uint32_t plane_bytes, plane_pitch;   // smem slot = plane + zero word, padded to plane_pitch
    int stages;                // plane ring: `stages` stages of `group` plane slots (plane_pitch bytes each)
    int group;                 // views per stage (compile-time kG of the kernel instantiation)
    int ring_rows;             // offset-row ring: `ring_rows` rows of 512 B shared by the stages in flight
    int *trace;                // diagnostics: per-warp per-stage clock trace of CTA 0 (tools/lift_trace.py), or null
    int debug;                 // diagnostics (tools/lift_probe.py): 1 = no gather, 4 = no plane copies, 16 = no epilogue,
                               // 32 = index-order pairing (no cost balancing), 8 = no offset-row copies
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
    // try_wait WITHOUT a suspend-time hint: with a hint ptxas emits TRYWAIT + NANOSLEEP.SYNCS <hint>, and the
    // sleeping warp was observed to come back only after the full hint (2 us) instead of at phase completion,
    // which paced the whole ring at one stage per (hint / stages in flight)
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
// Geometry tables.  grid = (octs, 16-view groups); one thread per position.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOct)
k_plane_index(const Tiling tiling, const float *__restrict__ points, const float *__restrict__ proj, int nv, int nvp,
              int n_octs, int64_t n_vox, int64_t n_pad, int height, int width, int elt, uint32_t zero_off,
              uint16_t *__restrict__ off16, uint8_t *__restrict__ cnt8, uint8_t *__restrict__ omask,
              uint8_t *__restrict__ cost16, int view_w) {
    __shared__ float sp[kIdxViews * 12];
    __shared__ unsigned smask;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the dependent kernels start their prologues
    const int v0 = blockIdx.y * kIdxViews;
    const int nvg = min(kIdxViews, nv - v0);
    if (threadIdx.x < nvg * 12) sp[threadIdx.x] = proj[v0 * 12 + threadIdx.x];
    if (threadIdx.x == 0) smask = 0u;
    __syncthreads();
    const int lane_slot = threadIdx.x >> 3, i8 = threadIdx.x & 7;
    const int64_t n = tiling.voxel(blockIdx.x, lane_slot) + i8;
    const int64_t p = (int64_t)blockIdx.x * kOct + threadIdx.x;
    const bool inside = n < n_vox;
    float X = 0.f, Y = 0.f, Z = 0.f;
    if (inside) {
        X = __ldg(points + n);
        Y = __ldg(points + n_vox + n);
        Z = __ldg(points + 2 * n_vox + n);
    }
    int count = 0;
    unsigned m = 0u;
#pragma unroll
    for (int i = 0; i < kIdxViews; ++i) {
        if (i < nvg) {
            float xr, yr, q2;
            const bool ok = project_nearest(sp + i * 12, X, Y, Z, height, width, xr, yr, q2) && inside;
            const uint32_t off = ok ? (uint32_t)((int)yr * width + (int)xr) * (uint32_t)elt : zero_off;
            off16[(int64_t)(v0 + i) * n_pad + p] = (uint16_t)off;
            count += ok ? 1 : 0;
            // warp = 4 lane slots x 8 voxels; quad 0 = voxels 0..3, quad 1 = voxels 4..7 of every lane slot
            const unsigned b = __ballot_sync(0xffffffffu, ok);
            m |= ((b & 0x0f0f0f0fu) ? 1u : 0u) << (2 * i);
            m |= ((b & 0xf0f0f0f0u) ? 2u : 0u) << (2 * i);
        }
    }
    if ((threadIdx.x & 31) == 0 && m != 0u) atomicOr(&smask, m);
    cnt8[(int64_t)blockIdx.y * n_pad + p] = (uint8_t)count;
    __syncthreads();
    const unsigned sm = smask;
    if (threadIdx.x < kIdxViews) omask[(int64_t)blockIdx.x * nvp + v0 + threadIdx.x] = (uint8_t)((sm >> (2 * threadIdx.x)) & 3u);
    // cost of the oct in this view group: active quad-views plus `view_w` per view that sees the oct at all (a tile-view
    // costs its warp a fixed ~400 cycles of list / offset-row handling besides the gathers of its quads)
    if (threadIdx.x == 32) {
        const unsigned any = (sm | (sm >> 1)) & 0x55555555u;
        cost16[(int64_t)blockIdx.y * n_octs + blockIdx.x] = (uint8_t)(__popc(sm) + view_w * __popc(any));
    }
}

// ---------------------------------------------------------------------------------------------
// Pairing + compaction.  grid = (views, parts).  Every block repeats the (tiny) pairing for its part -- ranks the
// octs by cost, pairs the k-th most with the k-th least expensive one (one pair per compute warp), counts per view
// how many of the part's octs see it and where the view's block of offset rows starts -- and then compacts the offset
// rows of ITS view: the rows of the part's octs that see the view are copied into one contiguous block (so that
// k_lift_planes fetches them with a single bulk copy) and every warp gets its entry.  The blocks of view 0 publish the
// per-part tables (pairs, row counts) the lift kernel needs.  (A separate single-block ranking kernel cost 21 us of
// pure latency on the critical path.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_plane_pack(int nv, int nvp, int nw16, int n_octs, int n_parts, int W, int balanced, int ring_rows, int64_t n_pad,
             int64_t part_rows, const uint16_t *__restrict__ off16, const uint8_t *__restrict__ omask,
             const uint8_t *__restrict__ cost16, uint16_t *__restrict__ pairs, uint16_t *__restrict__ rowcnt,
             uint16_t *__restrict__ offc, uint32_t *__restrict__ ents) {
    __shared__ uint16_t s_cost[kPMaxOcts];
    __shared__ uint16_t s_sorted[kPMaxOcts];
    __shared__ uint16_t s_q[2 * kPMaxWarps];       // octs of this part in slot order (warp w: 2 w, 2 w + 1)
    __shared__ uint8_t s_slot[2 * kPMaxWarps];     // their row slot in the view's block (0xff: not seen)
    __shared__ uint32_t s_gstart;                  // first row of this view's block in the part's table
    extern __shared__ uint8_t s_dyn[];             // [2 W][nvp] quad masks of the part's octs, then uint16 [nvp] row counts
    uint8_t *s_m = s_dyn;
    uint16_t *s_c = reinterpret_cast<uint16_t *>(s_dyn + (((size_t)2 * W * nvp + 15) & ~(size_t)15));
    asm volatile("griddepcontrol.wait;" ::: "memory");               // tables of k_plane_index
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // k_lift_planes may start streaming planes
    const int v = blockIdx.x, part = blockIdx.y, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, n_warps = (int)(blockDim.x >> 5);
    if (balanced) {                                                  // n_octs <= kPMaxOcts == blockDim.x
        if (tid < n_octs) {
            int cst = 0;
            for (int g0 = 0; g0 < nw16; g0 += 8) {                   // 8 independent loads in flight
                uint8_t c8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) c8[k] = g0 + k < nw16 ? __ldg(cost16 + (int64_t)(g0 + k) * n_octs + tid) : (uint8_t)0;
#pragma unroll
                for (int k = 0; k < 8; ++k) cst += (int)c8[k];
            }
            s_cost[tid] = (uint16_t)cst;
        }
        __syncthreads();
        if (tid < n_octs) {                                          // rank = number of octs that come before this one
            const int ct = s_cost[tid];
            int r = 0;
            for (int j = 0; j < n_octs; ++j) {
                const int cj = s_cost[j];
                r += (cj > ct || (cj == ct && j < tid)) ? 1 : 0;
            }
            s_sorted[r] = (uint16_t)tid;
        }
        __syncthreads();
    }
    const int n_pairs = (n_octs + 1) / 2;
    if (tid < W) {
        const int pair = tid * n_parts + part;
        uint16_t qa = 0xffffu, qb = 0xffffu;
        if (pair < n_pairs) {
            const int ia = pair, ib = n_octs - 1 - pair;
            qa = balanced ? s_sorted[ia] : (uint16_t)ia;
            if (ib > ia) qb = balanced ? s_sorted[ib] : (uint16_t)ib;
        }
        s_q[2 * tid] = qa;
        s_q[2 * tid + 1] = qb;
        if (v == 0) {
            pairs[((int64_t)part * W + tid) * 2] = qa;
            pairs[((int64_t)part * W + tid) * 2 + 1] = qb;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < 2 * W * nv; idx += blockDim.x) {       // quad masks of the part's octs in every view
        const int i = idx / nv, vv = idx - i * nv;
        const uint16_t q = s_q[i];
        s_m[i * nvp + vv] = q != 0xffffu ? (uint8_t)(__ldg(omask + (int64_t)q * nvp + vv) & 3u) : (uint8_t)0;
    }
    __syncthreads();
    for (int vv = tid; vv < nv; vv += blockDim.x) {                  // rows per view
        int c = 0;
        for (int i = 0; i < 2 * W; ++i) c += s_m[i * nvp + vv] ? 1 : 0;
        s_c[vv] = (uint16_t)c;
        if (v == 0) rowcnt[(int64_t)part * nvp + vv] = (uint16_t)c;
    }
    __syncthreads();
    if (warp == 0) {
        // first row of every view's block: rows + ring paddings (ring_pad) of the views before it
        int tot = 0;
        for (int vv = lane; vv < nv; vv += 32) tot += (int)s_c[vv];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        const int pad_total = (ring_rows - tot % ring_rows) % ring_rows;
        int base = 0;
        for (int vb = 0; vb < nv; vb += 32) {
            const int vv = vb + lane;
            const int eff = vv < nv ? (int)s_c[vv] + pad_total / nv + (vv < pad_total % nv ? 1 : 0) : 0;
            int incl = eff;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (vv == v) s_gstart = (uint32_t)(base + incl - eff);
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        // slot = number of active octs before this one in view v (2 W <= 64)
        const bool a0 = lane < 2 * W && s_m[lane * nvp + v] != 0, a1 = lane + 32 < 2 * W && s_m[(lane + 32) * nvp + v] != 0;
        const unsigned b0 = __ballot_sync(0xffffffffu, a0), b1 = __ballot_sync(0xffffffffu, a1);
        const unsigned below = (1u << lane) - 1u;
        if (lane < 2 * W) s_slot[lane] = a0 ? (uint8_t)__popc(b0 & below) : (uint8_t)0xff;
        if (lane + 32 < 2 * W) s_slot[lane + 32] = a1 ? (uint8_t)(__popc(b0) + __popc(b1 & below)) : (uint8_t)0xff;
    }
    __syncthreads();
    const int64_t g_start = (int64_t)s_gstart;
    if (tid < W) {
        const uint32_t qm = (uint32_t)s_m[(2 * tid) * nvp + v] | ((uint32_t)s_m[(2 * tid + 1) * nvp + v] << 2);
        const uint32_t e = qm ? ((uint32_t)v | (qm << 8) | ((uint32_t)s_slot[2 * tid] << 16) | ((uint32_t)s_slot[2 * tid + 1] << 24)) : 0u;
        ents[((int64_t)part * W + tid) * nvp + v] = e;
    }
    // rows: one warp per row, 16 B per lane
    for (int i = warp; i < 2 * W; i += n_warps) {
        if (!s_m[i * nvp + v]) continue;
        const uint4 *src = reinterpret_cast<const uint4 *>(off16 + (int64_t)v * n_pad + (int64_t)s_q[i] * kOct);
        uint4 *dst = reinterpret_cast<uint4 *>(offc + ((int64_t)part * part_rows + g_start + s_slot[i]) * kOct);
        dst[lane] = src[lane];
    }
}

// ---------------------------------------------------------------------------------------------
// Gather + statistics with the planes of one channel streamed through shared memory.
// ---------------------------------------------------------------------------------------------
// flag words in shared memory (release / acquire at CTA scope): the per-stage hand-shake between 25 consumer warps
// and the producers goes through these instead of mbarriers -- 50 mbarrier operations per stage (25 arrivals plus
// 25 waits on one barrier word) were measured at ~700 cycles per stage and paced the whole ring
// Plain volatile accesses, no fences: a warp's shared-memory instructions execute in program order in the SM's
// load/store pipeline, so the flag store follows the warp's gathers of the stage, and data written by a bulk copy
// is in shared memory before the copy's mbarrier completes (which the forwarder observes before it raises `ready`).
// (st.release / ld.acquire cost ~200 / ~130 cycles each here -- a fence per stage and warp.)
__device__ __forceinline__ void st_release(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

template <typename T>
__device__ __forceinline__ void gather_quad(uint32_t sb, uint32_t w0, uint32_t w1, unsigned long long &s1a,
                                            unsigned long long &s2a, unsigned long long &s1b,
                                            unsigned long long &s2b) {
    const float f0 = lds_elt<T>(sb + (w0 & 0xffffu));
    const float f1 = lds_elt<T>(sb + (w0 >> 16));
    const float f2 = lds_elt<T>(sb + (w1 & 0xffffu));
    const float f3 = lds_elt<T>(sb + (w1 >> 16));
    acc2(s1a, s2a, f0, f1);
    acc2(s1b, s2b, f2, f3);
}

__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// Ring bookkeeping shared by the producer and the consumers.  The offset rows of view v occupy eff(v) consecutive
// (circular) rows of the row ring, eff(v) = rowcnt(v) + pad(v); the pads make the per-unit total a multiple of
// the ring size R, so that the block of view v starts at the same ring row in every unit and the consumers can
// bake absolute ring rows into their entries.
__device__ __forceinline__ int ring_pad(int total_rows, int R, int nv, int v) {
    const int P = (R - total_rows % R) % R;
    return P / nv + (v < P % nv ? 1 : 0);
}

// entry of the active-view list: view | quad mask << 8 | ring row of oct A << 12 | ring row of oct B << 22
constexpr int kPMaxRingRows = 1023;

// kPf (experimental, ND_LIFT_PREFETCH=1, not the default): the consumers walk their list of ENTRIES instead of the stages
// and fetch the offset rows of the next entry before they gather the current one, see the consumer loop.
template <typename T, bool kRaw, bool kDiag, int kG, bool kPf = false>
__global__ void __launch_bounds__((kPMaxWarps + 3) * 32, 1)
k_lift_planes(const PlaneArgs a) {
    const int S = a.stages;                                        // stages of kG consecutive views each
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = (blockDim.x >> 5) - 3;                           // compute warps (+ plane producer, row producer, forwarder)
    // persistent CTA: units blockIdx.x, blockIdx.x + gridDim.x, ...; gridDim.x is a multiple of n_parts,
    // so the part (and with it every oct-dependent table) is the same for all units of a CTA
    const int part = blockIdx.x % a.n_parts;
    const int list_pitch = a.nv + 1;                               // entries per warp (+ sentinel)

    // plane ring (S slots), offset-row ring (R rows); then the barriers, the per-view row counts and the active-view lists
    const int R = a.ring_rows;
    const uint32_t stage_pitch = (uint32_t)kG * a.plane_pitch;
    const uint32_t row_base = smem_u32(smem) + (uint32_t)S * stage_pitch;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)S * stage_pitch + (size_t)R * kRowBytes);
    uint32_t *s_flags = reinterpret_cast<uint32_t *>(bars + 2 * S);                                        // [32] progress of the warps, [32] ready
    uint16_t *s_cnt = reinterpret_cast<uint16_t *>(s_flags + 64);                                          // [nvp] rows to copy
    uint16_t *s_eff = s_cnt + a.nvp;                                                                        // [nvp] rows of ring space
    uint16_t *s_pos = s_eff + a.nvp;                                                                        // [stages per unit] first ring row
    uint16_t *s_sn = s_pos + a.nvp;                                                                         // [stages per unit] ring rows
    uint16_t *s_lag = s_sn + a.nvp;                                                                         // [stages per unit] see below
    uint32_t *s_grow = reinterpret_cast<uint32_t *>(s_lag + a.nvp);                                         // [stages per unit] first row in the part's table
    uint32_t *s_lists = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(s_cnt) + (((size_t)a.nvp * 14 + 15) & ~(size_t)15));
    const uint32_t sm_base = smem_u32(smem);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t f_progress = smem_u32(s_flags), f_ready = smem_u32(s_flags + 32);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_full + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 64) s_flags[threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < S * kG; i += blockDim.x)         // the zero word behind every plane slot
        *reinterpret_cast<uint32_t *>(smem + (size_t)i * a.plane_pitch + a.plane_bytes) = 0u;
    __syncthreads();

    const long long t_cta = kDiag ? clock64() : 0;
    if (warp == W + 2) {
        // ---------------- forwarder: turns the completion of a stage's bulk copies (mbarrier) into a flag word ----------------
        const int n_my = blockIdx.x < a.n_units ? (a.n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int total = n_my * ((a.nv + kG - 1) / kG);
        int s = 0;
        uint32_t par = 0;
        for (int i = 0; i < total; ++i) {
            mbar_wait(bar_full + 8 * s, par);
            if (lane == 0) st_release(f_ready, (uint32_t)(i + 1));
            if (++s == S) { s = 0; par ^= 1u; }
        }
        return;
    }
    if (warp >= W) {
        // ---------------- two producer warps: warp W streams the planes, warp W + 1 the offset rows ----------------
        // Their loops are kept free of integer divisions and of everything that can be tabulated: one elected
        // thread issuing ~100 dependent instructions per view was the bottleneck of the whole CTA.
        const bool plane_warp = warp == W;
        const int64_t view_bytes = a.sv * (int64_t)sizeof(T);
        const int n_my = blockIdx.x < a.n_units ? (a.n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int spu = (a.nv + kG - 1) / kG;                      // stages per unit
        const int total = n_my * spu;                              // stages this CTA runs through
        const int pre = min(S, total);
        const char *feat = reinterpret_cast<const char *>(a.feat);
        const int64_t chan_bytes = a.sc * (int64_t)sizeof(T);
        const bool copy_planes = !kDiag || !(a.debug & 4);
        // the planes are kernel inputs, not products of the index / pack kernels: the planes of the first S stages
        // start streaming before the dependency wait; their barriers get the arrival (and the offset bytes) afterwards
        if (plane_warp && lane == 0 && copy_planes) {
            int u = blockIdx.x, v = 0;
            const char *src = feat + (int64_t)(u / a.n_parts) * chan_bytes;
            for (int i = 0; i < pre; ++i) {
                const int nvs = min(kG, a.nv - v);
                mbar_expect_tx_only(bar_full + 8 * i, (uint32_t)nvs * a.plane_bytes);
                for (int g = 0; g < nvs; ++g)
                    bulk_g2s(sm_base + (uint32_t)(i * kG + g) * a.plane_pitch, src + g * view_bytes, a.plane_bytes, bar_full + 8 * i);
                src += nvs * view_bytes;
                v += nvs;
                if (v == a.nv) { v = 0; u += gridDim.x; src = feat + (int64_t)(u / a.n_parts) * chan_bytes; }
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");         // row counts / offset blocks come from k_plane_pack
        // Per-stage tables (both warps compute the same values): ring rows s_sn (rows + paddings of the stage's views),
        // first ring row s_pos, first row in the part's table s_grow, and s_lag = how many of the preceding stages may
        // still be unreleased when the stage's rows are written (they must fit the ring together; < S).
        int tot = 0;
        for (int v = lane; v < a.nv; v += 32) {
            const int c = (int)a.rowcnt[(int64_t)part * a.nvp + v];
            s_cnt[v] = (uint16_t)c;
            tot += c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        for (int v = lane; v < a.nv; v += 32) s_eff[v] = (uint16_t)((int)s_cnt[v] + ring_pad(tot, R, a.nv, v));
        __syncwarp();
        if (lane == 0) {
            int g_row = 0;
            for (int j = 0; j < spu; ++j) {
                int n = 0;
                for (int g = 0; g < kG && j * kG + g < a.nv; ++g) n += (int)s_eff[j * kG + g];
                s_sn[j] = (uint16_t)n;
                s_grow[j] = (uint32_t)g_row;
                s_pos[j] = (uint16_t)(g_row % R);
                g_row += n;
            }
            for (int j = 0; j < spu; ++j) {
                int sum = (int)s_sn[j], d = 0, k = j;
                while (d < S - 1) {
                    k = k == 0 ? spu - 1 : k - 1;
                    if (sum + (int)s_sn[k] > R) break;
                    sum += (int)s_sn[k];
                    ++d;
                }
                s_lag[j] = (uint16_t)d;
            }
        }
        __syncwarp();
        const char *off_part = reinterpret_cast<const char *>(a.offc) + (int64_t)part * a.part_rows * kRowBytes;
        int n_tr = 0;
        // stage i = (unit u, stage j of the unit): slot s
        int s = 0, j = 0, u = blockIdx.x;
        const char *psrc = feat + (int64_t)(u / a.n_parts) * chan_bytes;
        int released = 0;                                          // stages every consumer warp is done with
        for (int i = 0; i < total; ++i) {
            const int n = (int)s_sn[j];
            const uint32_t fb = bar_full + 8 * s;
            // planes: the slots are free once stage i - S is released; rows: stages older than i - lag released
            const int need = plane_warp ? i - S + 1 : i - (int)s_lag[j];
            const long long tp0 = kDiag && a.trace != nullptr ? clock64() : 0;
            if (released < need) {
                do {                                                // lane w reads the progress word of consumer warp w
                    const uint32_t pr = lane < W ? ld_acquire(f_progress + 4 * lane) : 0x7fffffffu;
                    released = (int)__reduce_min_sync(0xffffffffu, pr);
                } while (released < need);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // their reads before our async-proxy writes
            }
            if (kDiag && a.trace != nullptr && blockIdx.x == 0 && lane == 0 && plane_warp && n_tr < 256) {
                int *t = a.trace + ((size_t)W * 256 + n_tr) * 4;
                t[0] = (int)(tp0 - t_cta);
                t[1] = (int)(clock64() - t_cta);
                t[2] = n;
                t[3] = i - released;
                ++n_tr;
            }
            if (plane_warp) {
                const int nvs = min(kG, a.nv - j * kG);
                const bool with_planes = i >= pre && copy_planes;   // (i < pre: the plane bytes were announced above)
                const bool copy_rows = !kDiag || !(a.debug & 8);
                if (lane == 0)
                    mbar_expect_tx(fb, (with_planes ? (uint32_t)nvs * a.plane_bytes : 0u) + (copy_rows ? (uint32_t)n * kRowBytes : 0u));
                __syncwarp();
                if (with_planes && lane < nvs)
                    bulk_g2s(sm_base + (uint32_t)(s * kG + lane) * a.plane_pitch, psrc + lane * view_bytes, a.plane_bytes, fb);
                psrc += nvs * view_bytes;
            } else if (lane == 0 && n != 0 && (!kDiag || !(a.debug & 8))) {
                const char *src = off_part + (size_t)s_grow[j] * kRowBytes;
                const int pos = (int)s_pos[j];
                const int first = min(n, R - pos);
                bulk_g2s(row_base + (uint32_t)pos * kRowBytes, src, (uint32_t)first * kRowBytes, fb);
                if (n > first)                                      // the block wraps around the end of the ring
                    bulk_g2s(row_base, src + (size_t)first * kRowBytes, (uint32_t)(n - first) * kRowBytes, fb);
            }
            if (++s == S) s = 0;
            if (++j == spu) {                                       // next unit of this CTA: one division per unit
                j = 0;
                u += gridDim.x;
                psrc = feat + (int64_t)(u / a.n_parts) * chan_bytes;
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    asm volatile("griddepcontrol.wait;" ::: "memory");             // tables come from k_plane_index / k_plane_pack (PDL)
    const uint16_t qa16 = __ldg(a.pairs + ((int64_t)part * W + warp) * 2);
    const uint16_t qb16 = __ldg(a.pairs + ((int64_t)part * W + warp) * 2 + 1);
    const int qa = qa16 == 0xffffu ? -1 : (int)qa16, qb = qb16 == 0xffffu ? -1 : (int)qb16;
    const int64_t pa = (int64_t)(qa >= 0 ? qa : 0) * kOct + lane * kOV;
    const int64_t pb = (int64_t)(qb >= 0 ? qb : 0) * kOct + lane * kOV;
    const int64_t n0a = qa >= 0 ? a.tiling.voxel(qa, lane) : a.n_vox;
    const int64_t n0b = qb >= 0 ? a.tiling.voxel(qb, lane) : a.n_vox;

    // view counts of this lane's 16 voxels (sum of the uint8 partials), packed 4 per word
    uint32_t cws[4] = {0u, 0u, 0u, 0u};
    for (int g = 0; g < a.nw16; ++g) {
        if (qa >= 0) {
            const uint2 t = __ldg(reinterpret_cast<const uint2 *>(a.cnt8 + (int64_t)g * a.n_pad + pa));
            cws[0] = __vadd4(cws[0], t.x); cws[1] = __vadd4(cws[1], t.y);
        }
        if (qb >= 0) {
            const uint2 t = __ldg(reinterpret_cast<const uint2 *>(a.cnt8 + (int64_t)g * a.n_pad + pb));
            cws[2] = __vadd4(cws[2], t.x); cws[3] = __vadd4(cws[3], t.y);
        }
    }

    // The views that see this warp's octs, as a list of entries (view | quad mask | ring rows of the two octs):
    // views that see none of the four quads cost the warp nothing but the stage hand-shake.
    uint32_t *act = s_lists + warp * list_pitch;
    {
        const uint32_t *src = a.ents + ((int64_t)part * W + warp) * a.nvp;
        const uint16_t *rc = a.rowcnt + (int64_t)part * a.nvp;
        int tot = 0;
        for (int v = lane; v < a.nv; v += 32) tot += (int)__ldg(rc + v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        int n_act = 0, start = 0;                                  // start: ring row of the block of view vb (before lane prefix)
        for (int vb = 0; vb < a.nv; vb += 32) {
            const int v = vb + lane;
            const uint32_t e = v < a.nv ? __ldg(src + v) : 0u;
            int eff = v < a.nv ? (int)__ldg(rc + v) + ring_pad(tot, R, a.nv, v) : 0;
            int incl = eff;                                        // inclusive prefix of eff over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int my_start = (start + incl - eff) % R;
            start = (start + __shfl_sync(0xffffffffu, incl, 31)) % R;
            const unsigned b = __ballot_sync(0xffffffffu, e != 0u);
            if (e) {
                int ra = my_start + (int)((e >> 16) & 0xffu), rb = my_start + (int)(e >> 24);
                if (ra >= R) ra -= R;
                if (rb >= R) rb -= R;
                if (!(e & 0x300u)) ra = 0;
                if (!(e & 0xc00u)) rb = 0;
                act[n_act + __popc(b & ((1u << lane) - 1u))] = (e & 0xfffu) | ((uint32_t)ra << 12) | ((uint32_t)rb << 22);
            }
            n_act += __popc(b);
        }
        if (lane == 0) act[n_act] = 0xffu;                          // sentinel: view 255 is never reached
        __syncwarp();
    }

    int s = 0, n_tr = 0;
    uint32_t sb = sm_base;                                         // planes of the current stage
    uint32_t gi = 0, ready = 0;                                    // stages this warp is done with / known to have landed
    const uint32_t my_progress = f_progress + 4 * warp;
    const uint32_t lane_row = row_base + lane * (kOV * 2);
    const bool do_gather = !kDiag || !(a.debug & 1);
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        const int c = u / a.n_parts;
        unsigned long long s1[kPV / 2], s2[kPV / 2];
#pragma unroll
        for (int j = 0; j < kPV / 2; ++j) { s1[j] = 0ull; s2[j] = 0ull; }
        const uint32_t *ap = act;
        uint32_t ent = *ap;

        if constexpr (kPf) {
            // Entry-driven walk.  In the stage-driven loop below every view that sees the warp costs two dependent
            // shared-memory round trips before its first gather (list entry -> offset rows -> plane), ~245 cycles per
            // view on the busiest warp (half of its time, tools/sim_lift_smem.py).  Here the offset rows of the NEXT
            // entry are requested before the gathers of the current one are issued (they are in the row ring as soon as
            // their stage has landed, which is certain when it is the current stage and otherwise checked against the
            // forwarder's counter), and stages the warp does not appear in are skipped with one progress update.
            const int spu = (a.nv + kG - 1) / kG;
            const uint32_t g0 = gi;                                 // global index of this unit's first stage
            int cur_j = 0;                                          // stage of the unit the warp stands at: gi == g0 + cur_j
            uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = make_uint4(0u, 0u, 0u, 0u);
            bool have = false;                                      // c0 / c1 already hold the rows of `ent`
            while ((ent & 0xffu) != 0xffu) {
                const int v = (int)(ent & 0xffu), j = v / kG, g = v % kG;
                if (j != cur_j) {                                   // every stage before j is done for this warp
                    s = (s + (j - cur_j)) % S;
                    sb = sm_base + (uint32_t)s * stage_pitch;
                    cur_j = j;
                    gi = g0 + (uint32_t)j;
                    __syncwarp();
                    if (lane == 0) st_release(my_progress, gi);
                }
                if (!have) {
                    while (ready <= gi) ready = ld_acquire(f_ready);
                    if (ent & 0x300u) c0 = lds_u4(lane_row + ((ent >> 12) & 0x3ffu) * kRowBytes);
                    if (ent & 0xc00u) c1 = lds_u4(lane_row + (ent >> 22) * kRowBytes);
                }
                const uint32_t pb = sb + (uint32_t)g * a.plane_pitch;
                const uint32_t nxt = *++ap;
                uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = make_uint4(0u, 0u, 0u, 0u);
                bool nhave = false;
                if ((nxt & 0xffu) != 0xffu) {
                    const uint32_t gn = g0 + (nxt & 0xffu) / kG;
                    if (gn != gi && ready <= gn) ready = ld_acquire(f_ready);      // one look, no spinning
                    if (gn == gi || ready > gn) {
                        if (nxt & 0x300u) n0 = lds_u4(lane_row + ((nxt >> 12) & 0x3ffu) * kRowBytes);
                        if (nxt & 0xc00u) n1 = lds_u4(lane_row + (nxt >> 22) * kRowBytes);
                        nhave = true;
                    }
                }
                if (ent & 0x100u) gather_quad<T>(pb, c0.x, c0.y, s1[0], s2[0], s1[1], s2[1]);
                if (ent & 0x200u) gather_quad<T>(pb, c0.z, c0.w, s1[2], s2[2], s1[3], s2[3]);
                if (ent & 0x400u) gather_quad<T>(pb, c1.x, c1.y, s1[4], s2[4], s1[5], s2[5]);
                if (ent & 0x800u) gather_quad<T>(pb, c1.z, c1.w, s1[6], s2[6], s1[7], s2[7]);
                c0 = n0;
                c1 = n1;
                have = nhave;
                ent = nxt;
            }
            // the rest of the unit's stages hold nothing for this warp
            s = (s + (spu - cur_j)) % S;
            sb = sm_base + (uint32_t)s * stage_pitch;
            gi = g0 + (uint32_t)spu;
            __syncwarp();
            if (lane == 0) st_release(my_progress, gi);
        } else
        for (int v0 = 0; v0 < a.nv; v0 += kG) {
            const long long tc0 = kDiag && a.trace != nullptr ? clock64() : 0;
            while (ready <= gi) ready = ld_acquire(f_ready);       // the forwarder publishes the number of landed stages
            const long long tc1 = kDiag && a.trace != nullptr ? clock64() : 0;
            int n_ent = 0;
#pragma unroll
            for (int g = 0; g < kG; ++g) {
                if ((ent & 0xffu) == (uint32_t)(v0 + g)) {         // warp-uniform: the list is per warp
                    if (do_gather) {
                        const uint32_t pb = sb + (uint32_t)g * a.plane_pitch;
                        if (ent & 0x300u) {
                            const uint4 c0 = lds_u4(lane_row + ((ent >> 12) & 0x3ffu) * kRowBytes);
                            if (ent & 0x100u) gather_quad<T>(pb, c0.x, c0.y, s1[0], s2[0], s1[1], s2[1]);
                            if (ent & 0x200u) gather_quad<T>(pb, c0.z, c0.w, s1[2], s2[2], s1[3], s2[3]);
                        }
                        if (ent & 0xc00u) {
                            const uint4 c1 = lds_u4(lane_row + (ent >> 22) * kRowBytes);
                            if (ent & 0x400u) gather_quad<T>(pb, c1.x, c1.y, s1[4], s2[4], s1[5], s2[5]);
                            if (ent & 0x800u) gather_quad<T>(pb, c1.z, c1.w, s1[6], s2[6], s1[7], s2[7]);
                        }
                    }
                    ent = *++ap;
                    ++n_ent;
                }
            }
            __syncwarp();
            ++gi;
            if (lane == 0) st_release(my_progress, gi);
            if (kDiag && a.trace != nullptr && blockIdx.x == 0 && lane == 0 && n_tr < 256) {
                int *t = a.trace + ((size_t)warp * 256 + n_tr) * 4;
                t[0] = (int)(tc0 - t_cta);
                t[1] = (int)(tc1 - t_cta);
                t[2] = (int)(clock64() - t_cta);
                t[3] = n_ent;
                ++n_tr;
            }
            sb += stage_pitch;
            if (++s == S) { s = 0; sb = sm_base; }
        }
        if (kDiag && (a.debug & 16)) continue;

        // ---------------- epilogue of unit (c, part): the producer is already streaming the next unit ----------------
        const int64_t row = (int64_t)c * a.n_vox;
        const bool base_ok = (row % 4 == 0) &&
                             ((reinterpret_cast<uintptr_t>(a.out_a) | reinterpret_cast<uintptr_t>(a.out_b)) % 16 == 0);
#pragma unroll
        for (int g = 0; g < kPV / 4; ++g) {
            const int64_t nb = (g < 2 ? n0a : n0b) + 4 * (g & 1);  // first of this quad's 4 voxels
            if (nb >= a.n_vox) continue;
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
                if (kRaw) {
                    oa[j] = v1[j];
                    ob[j] = v2[j];
                } else if (cnv[j] > 0) {
                    const float cf = (float)cnv[j];                               // count + 1e-8 == count in fp32
                    const float m = v1[j] / cf;                                   // IEEE divide: the mean is bit-equal to the reference's
                    const float rc = __frcp_rn(cf);
                    float ssd = fmaxf(fmaf(-m, v1[j], v2[j]), 0.0f);              // sum over valid views of (f - m)^2
                    ssd = fmaf((float)(a.n_views_total - cnv[j]) * m, m, ssd);    // invalid views contribute m^2 each
                    float al = 1.0f;
                    if (a.alpha != nullptr && nb + j < a.n_vox) al = __ldg(a.alpha + nb + j);
                    oa[j] = m * al;
                    ob[j] = exp2f(ssd * rc * -1.4426950408889634f);               // exp(-var); ex2.approx, rel. error ~1e-7 (1 + var)
                } else {
                    oa[j] = 0.0f;                                                 // nerfdet.py:176
                    ob[j] = 0.0f;                                                 // exp(-1e6) == 0 (nerfdet.py:180-181)
                }
            }
            const int64_t o = row + nb;
            if (base_ok && nb + 4 <= a.n_vox && nb % 4 == 0) {
                __stcs(reinterpret_cast<float4 *>(a.out_a + o), make_float4(oa[0], oa[1], oa[2], oa[3]));
                if (a.out_b != nullptr)
                    __stcs(reinterpret_cast<float4 *>(a.out_b + o), make_float4(ob[0], ob[1], ob[2], ob[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (nb + j < a.n_vox) {
                        a.out_a[o + j] = oa[j];
                        if (a.out_b != nullptr) a.out_b[o + j] = ob[j];
                    }
                }
            }
            if (c == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t n = nb + j;
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

struct PlaneGeom {
    Tiling tiling;
    int elt, n_pix, nw16, nvp, n_octs, n_pairs, n_parts, warps, stages, group, grid;
    int64_t n_pad, part_rows;
    uint32_t plane_bytes, plane_pitch;
    int ring_rows;
    size_t off_bytes, cnt_bytes, mask_bytes, cost_bytes, offc_bytes, rowcnt_bytes, ents_bytes, pairs_bytes, gstart_bytes,
        total_bytes, smem_bytes;
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
    if (f->n_views > 254) return false;                                     // uint8 view counts, view 255 = list sentinel
    g.plane_bytes = (uint32_t)pb;
    g.nw16 = (f->n_views + kIdxViews - 1) / kIdxViews;      // view groups of the index pass
    g.nvp = g.nw16 * kIdxViews;
    g.tiling = Tiling{0, 0, 0, 0, 0};
    g.n_octs = (int)ceil_div(n_vox, kOct);
    if (g.n_octs > 65000) return false;                                     // uint16 oct ids
    if (opt != nullptr && opt->grid_x > 0 && opt->grid_y > 0 && opt->grid_z > 0 &&
        (int64_t)opt->grid_x * opt->grid_y * opt->grid_z == n_vox && opt->grid_z % kOV == 0 &&
        opt->grid_x % kPBx == 0 && opt->grid_y % kPBy == 0) {
        g.tiling.compact = 1;
        g.tiling.gy = opt->grid_y;
        g.tiling.gz = opt->grid_z;
        g.tiling.nzo = opt->grid_z / kOV;
        g.tiling.tiles_y = opt->grid_y / kPBy;
    }
    g.n_pad = (int64_t)g.n_octs * kOct;
    g.n_pairs = (g.n_octs + 1) / 2;
    int max_warps = kPMaxWarps, stages = 0, group = 2;
    if (const char *e = getenv("ND_LIFT_STAGES")) stages = atoi(e);         // tuning knobs for tools/lift_probe.py
    if (const char *e = getenv("ND_LIFT_GROUP")) group = atoi(e);
    if (group != 1 && group != 2 && group != 4) group = 2;
    if (const char *e = getenv("ND_LIFT_WARPS")) {
        const int v = atoi(e);
        if (v >= 1 && v < max_warps) max_warps = v;
    }
    g.n_parts = (int)ceil_div(g.n_pairs, max_warps);
    g.warps = (int)ceil_div(g.n_pairs, g.n_parts);
    g.plane_pitch = (uint32_t)align_up((size_t)pb + 16, 128);
    const size_t fixed = 2 * kPMaxStages * 8 + 256 + align_up((size_t)g.nvp * 14, 16) + (size_t)g.warps * (f->n_views + 1) * 4 + 16;
    const size_t avail = (size_t)(227 * 1024) - fixed;
    // plane slots S and offset rows R share the rest: R must hold the largest block (2 W rows); by default every
    // stage in flight gets room for about half of the part's octs (the typical share that sees a view)
    const size_t min_rows = (size_t)2 * g.warps * group;                        // one stage's worst case
    const size_t stage_bytes = (size_t)group * g.plane_pitch;
    while (group > 1 && avail < 2 * stage_bytes + min_rows * kRowBytes) group >>= 1;
    if (avail < 2 * (size_t)group * g.plane_pitch + (size_t)2 * g.warps * group * kRowBytes) return false;
    const size_t sb = (size_t)group * g.plane_pitch, mr = (size_t)2 * g.warps * group;
    if (stages <= 0) {
        // default: as many stages as fit (at most 8 plane slots) while the row ring keeps room for about half of
        // the part's octs (the typical share that sees a view) per view in flight
        stages = 2;
        while ((stages + 1) * group <= 8 &&
               (size_t)(stages + 1) * sb + std::max(mr, (size_t)(stages + 1) * group * g.warps) * kRowBytes <= avail)
            ++stages;
    }
    if (stages < 2) stages = 2;
    if (stages > kPMaxStages) stages = kPMaxStages;
    while (stages > 2 && (size_t)stages * sb + mr * kRowBytes > avail) --stages;
    g.group = group;
    for (;;) {
        g.ring_rows = (int)((avail - (size_t)stages * sb) / kRowBytes);
        if (g.ring_rows > kPMaxRingRows) g.ring_rows = kPMaxRingRows;
        // one stage (rows of all the part's octs plus the ring paddings of its views) must fit the ring
        if (g.ring_rows >= group * (2 * g.warps + g.ring_rows / f->n_views + 1)) break;
        if (stages <= 2) return false;
        --stages;
    }
    g.stages = stages;
    // persistent grid: one CTA per SM, a multiple of n_parts (see k_lift_planes)
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (opt != nullptr && opt->sm_limit > 0 && opt->sm_limit < sms) sms = opt->sm_limit;
    if (const char *e = getenv("ND_LIFT_SMS")) { const int v = atoi(e); if (v > 0 && v < sms) sms = v; }
    const int64_t n_units = (int64_t)f->channels * g.n_parts;
    int64_t grid = n_units < sms ? n_units : sms;
    grid -= grid % g.n_parts;
    if (grid < g.n_parts) grid = g.n_parts;
    // the CTAs walk their units in rounds; the smallest grid with the same number of rounds takes the same time and leaves
    // SMs free for a concurrent kernel (512 units: 4 rounds on 148 CTAs, of which 80 idle in the last round -- or on 128)
    if (getenv("ND_LIFT_NO_TRIM") == nullptr) {
        const int64_t rounds = ceil_div(n_units, grid);
        int64_t trimmed = ceil_div(n_units, rounds);
        trimmed = ceil_div(trimmed, g.n_parts) * g.n_parts;
        if (trimmed < grid) grid = trimmed;
    }
    g.grid = (int)grid;
    g.off_bytes = align_up((size_t)f->n_views * g.n_pad * sizeof(uint16_t), 256);
    g.cnt_bytes = align_up((size_t)g.nw16 * g.n_pad, 256);
    g.mask_bytes = align_up((size_t)g.n_octs * g.nvp, 256);
    g.cost_bytes = align_up((size_t)g.nw16 * g.n_octs, 256);
    g.part_rows = (int64_t)f->n_views * 2 * g.warps + g.ring_rows;      // rows + paddings of one part, upper bound
    g.offc_bytes = align_up((size_t)g.n_parts * (size_t)g.part_rows * kRowBytes, 256);
    g.rowcnt_bytes = align_up((size_t)g.n_parts * g.nvp * sizeof(uint16_t), 256);
    g.ents_bytes = align_up((size_t)g.n_parts * g.warps * g.nvp * sizeof(uint32_t), 256);
    g.pairs_bytes = align_up((size_t)g.n_parts * g.warps * 2 * sizeof(uint16_t), 256);
    g.gstart_bytes = align_up((size_t)g.n_parts * g.nvp * sizeof(uint32_t), 256);
    g.total_bytes = g.off_bytes + g.cnt_bytes + g.mask_bytes + g.cost_bytes + g.offc_bytes + g.rowcnt_bytes + g.ents_bytes +
                    g.pairs_bytes + g.gstart_bytes;
    g.smem_bytes = (size_t)g.stages * g.group * g.plane_pitch + (size_t)g.ring_rows * kRowBytes + 2 * g.stages * 8 + 256 +
                   align_up((size_t)g.nvp * 14, 16) + (size_t)g.warps * (f->n_views + 1) * 4;
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

// launch with programmatic stream serialization: the kernel may start while its predecessor drains and
// executes griddepcontrol.wait before touching the predecessor's results
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
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
    uint16_t *off16 = reinterpret_cast<uint16_t *>(wsb);            wsb += g.off_bytes;
    uint8_t *cnt8 = reinterpret_cast<uint8_t *>(wsb);               wsb += g.cnt_bytes;
    uint8_t *omask = reinterpret_cast<uint8_t *>(wsb);              wsb += g.mask_bytes;
    uint8_t *cost16 = reinterpret_cast<uint8_t *>(wsb);             wsb += g.cost_bytes;
    uint16_t *offc = reinterpret_cast<uint16_t *>(wsb);             wsb += g.offc_bytes;
    uint16_t *rowcnt = reinterpret_cast<uint16_t *>(wsb);           wsb += g.rowcnt_bytes;
    uint32_t *ents = reinterpret_cast<uint32_t *>(wsb);             wsb += g.ents_bytes;
    uint16_t *pairs = reinterpret_cast<uint16_t *>(wsb);            wsb += g.pairs_bytes;
    uint32_t *gstart = reinterpret_cast<uint32_t *>(wsb);

    int debug = 0, view_w = 2;
    if (const char *e = getenv("ND_LIFT_DEBUG")) debug = atoi(e);
    if (const char *e = getenv("ND_LIFT_VIEW_W")) view_w = atoi(e);
    k_plane_index<<<dim3((unsigned)g.n_octs, (unsigned)g.nw16), kOct, 0, st>>>(
        g.tiling, points, proj, f->n_views, g.nvp, g.n_octs, n_vox, g.n_pad, f->height, f->width, g.elt, g.plane_bytes,
        off16, cnt8, omask, cost16, view_w);
    ND_CUDA_LAUNCH_CHECK("k_plane_index");
    const int balanced = (g.n_octs <= kPMaxOcts && !(debug & 32)) ? 1 : 0;
    const size_t pack_smem = (((size_t)2 * g.warps * g.nvp + 15) & ~(size_t)15) + (size_t)g.nvp * sizeof(uint16_t);
    ND_REQUIRE(pack_smem <= 40 * 1024, ND_ERR_BAD_SHAPE, "lift: too many views for the pairing pass (%zu bytes)", pack_smem);
    cudaError_t e = launch_pdl(k_plane_pack, dim3((unsigned)f->n_views, (unsigned)g.n_parts), dim3(256), pack_smem, st,
                               (int)f->n_views, g.nvp, g.nw16, g.n_octs, g.n_parts, g.warps, balanced, g.ring_rows, g.n_pad,
                               g.part_rows, (const uint16_t *)off16, (const uint8_t *)omask, (const uint8_t *)cost16, pairs,
                               rowcnt, offc, ents);
    if (e != cudaSuccess) {
        set_error("k_plane_pack: CUDA error %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    (void)gstart;

    PlaneArgs a{};
    a.tiling = g.tiling;
    a.cnt8 = cnt8;
    a.offc = offc;
    a.part_rows = g.part_rows;
    a.rowcnt = rowcnt;
    a.ents = ents;
    a.pairs = pairs;
    a.nv = f->n_views;
    a.nvp = g.nvp;
    a.nw16 = g.nw16;
    a.n_vox = n_vox;
    a.n_pad = g.n_pad;
    a.n_octs = g.n_octs;
    a.n_parts = g.n_parts;
    a.n_units = f->channels * g.n_parts;
    a.feat = f->data;
    a.sv = f->stride_v;
    a.sc = f->stride_c;
    a.plane_bytes = g.plane_bytes;
    a.plane_pitch = g.plane_pitch;
    a.stages = g.stages;
    a.group = g.group;
    a.ring_rows = g.ring_rows;
    a.trace = g_trace;
    a.debug = debug;
    a.n_views_total = f->n_views;
    a.alpha = alpha;
    a.out_a = out_a;
    a.out_b = out_b;
    a.count_i64 = count_i64;
    a.count_f32 = count_f32;
    // the diagnostics build (debug switches, clock trace) is a separate instantiation: the product kernel carries none of it
    const bool diag = debug != 0 || g_trace != nullptr;
    void (*kern)(const PlaneArgs) = nullptr;
    switch (g.group) {
        case 1: kern = diag ? k_lift_planes<T, kRaw, true, 1> : k_lift_planes<T, kRaw, false, 1>; break;
        case 2: kern = diag ? k_lift_planes<T, kRaw, true, 2> : k_lift_planes<T, kRaw, false, 2>; break;
        default: kern = diag ? k_lift_planes<T, kRaw, true, 4> : k_lift_planes<T, kRaw, false, 4>; break;
    }
    // experimental consumer loop (not validated on a GPU yet; opt-in for the next round's A/B, bench shape only: kG = 2)
    if (!diag && g.group == 2 && getenv("ND_LIFT_PREFETCH") != nullptr && atoi(getenv("ND_LIFT_PREFETCH")) == 1)
        kern = k_lift_planes<T, kRaw, false, 2, true>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    if (e != cudaSuccess) {
        set_error("k_lift_planes: cannot reserve %zu bytes of shared memory: %s", g.smem_bytes, cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    e = launch_pdl(kern, dim3((unsigned)g.grid), dim3((unsigned)(g.warps + 3) * 32), g.smem_bytes, st, a);
    if (e != cudaSuccess) {
        set_error("k_lift_planes: CUDA error %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    return ND_OK;
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
