// Plane-resident fused lift for the reference's own layout (NCHW maps whose sliced planes are contiguous, fp32 or
// bf16).  Replaces reference nerfdet.py:164-181 (backproject + mean / all-view variance / count) without the
// [nv][C][N] per-view volume.
//
// Both sides of the lift are CHANNEL-major -- input planes feat[v][c][pixel], output rows mean/cov[c][voxel] -- so no
// transposition is needed, only a random-access memory that holds one (view, channel) plane (18.9 KB at 59x80 fp32):
// shared memory.  Every plane byte leaves HBM once.
//
//   quad            32 lanes x 4 consecutive voxels (128 voxels, 16 B of an output row per lane).  With the lattice
//                   shape known and Z % 4 == 0, X % 4 == 0, Y % 8 == 0 a quad is a compact 4 x 8 x 4 block of voxels, so
//                   whole quads fall outside a camera frustum and cost nothing; otherwise 128 consecutive voxels.
//   geometry PLAN   (nd_lift_plan_build; depends on points / projection / depth only, NOT on the features, so a
//                   caller that lifts several feature stacks with the same cameras builds it once)
//     k_q_index     bit-exact nearest-pixel projection of every voxel-view (nd_common.cuh:project_nearest), stored as a
//                   uint16 BYTE offset into a plane (invalid -> offset of a zero word behind the plane), one 256 B row
//                   per (view, quad); per (quad, view) "any voxel valid"; per-voxel view counts.
//     k_q_rank      one block ranks the quads by the number of views that see them.
//     k_q_pack      deals the ranked quads out to the compute warps in snake order, 4 per warp (equal load per warp);
//                   a PART is the set of warps of one CTA.  Per part the
//                   offset rows of the active (view, quad) pairs are compacted into ONE stream in exactly the order
//                   the lift kernel consumes them: stage (kG views) -> warp -> view -> slot.
//   k_lift_quads    persistent, one CTA per SM.  Work unit = (channel c, part p); units are handed out by a ticket
//                   counter per part.  A producer warp streams, per stage, the kG planes of channel c and the stage's
//                   block of offset rows through an S-stage mbarrier ring with TMA bulk copies.  A compute warp keeps
//                   sum / sum of squares of its 16 voxels per lane in 32 registers (packed f32x2) across all views; per
//                   view that sees one of its quads it reads 8 B of offsets per lane and active quad, issues ALL the
//                   gathers of the view from the plane in shared memory, then accumulates.  The epilogue turns the
//                   accumulators into mean / exp(-var) (or raw S1 / S2 for the view-sharded path).
//   Programmatic dependent launch: a CTA triggers the dependent launch when it has drawn its last ticket, and the
//   kernel itself waits for its predecessor only before its first global WRITE, so the first unit of step i + 1 runs
//   on the SMs that step i's last round leaves idle (256 channels x 2 parts on 148 SMs: 3.46 rounds) -- back-to-back
//   lifts pack without the 13 % quantisation loss.
#include <algorithm>
#include <atomic>

#include "nd_common.cuh"

namespace nd {

constexpr int kQV = 4;                  // voxels per lane and quad
constexpr int kQuad = 32 * kQV;         // voxels per quad
constexpr int kSlots = 4;               // quads per compute warp
constexpr int kQMaxWarps = 25;          // compute warps per CTA (+ 1 producer warp)
constexpr int kQRowBytes = kQuad * 2;   // one offset row: uint16 per voxel of a quad
constexpr int kQBx = 4, kQBy = 8;       // compact quad = kQBx x kQBy columns in (x, y) x kQV in z
constexpr int kVG = 8;                  // views per block of k_q_index
constexpr int kQMaxQuads = 4096;        // bitonic ranking in shared memory
constexpr int kQMaxStages = 8;

struct QTiling {
    int compact;                        // 0: quad = 128 consecutive voxels
    int gy, gz, nzq, tiles_y;           // compact: lattice Y and Z, quads per Z column, 4 x 8 blocks along Y
    // first of the 4 consecutive voxels of lane `lane` in quad `q`
    __host__ __device__ __forceinline__ int64_t voxel(int q, int lane) const {
        if (!compact) return (int64_t)q * kQuad + lane * kQV;
        const int o = q % nzq, b = q / nzq;
        const int ty = b % tiles_y, tx = b / tiles_y;
        const int ix = tx * kQBx + (lane >> 3), iy = ty * kQBy + (lane & 7);
        return ((int64_t)ix * gy + iy) * gz + o * kQV;
    }
};

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) -----------------------------------
__device__ __forceinline__ uint32_t q_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void q_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void q_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool q_mbar_test(uint32_t bar, uint32_t parity) {      // non-blocking
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void q_mbar_wait(uint32_t bar, uint32_t parity) {      // hardware-suspended wait
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
__device__ __forceinline__ void q_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void q_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float q_lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float q_lds_bf16(uint32_t addr) {
    uint16_t h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __uint_as_float((uint32_t)h << 16);
}
__device__ __forceinline__ uint2 q_lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
template <typename T> __device__ __forceinline__ float q_lds_elt(uint32_t addr);
template <> __device__ __forceinline__ float q_lds_elt<float>(uint32_t addr) { return q_lds_f32(addr); }
template <> __device__ __forceinline__ float q_lds_elt<__nv_bfloat16>(uint32_t addr) { return q_lds_bf16(addr); }
// flag words in shared memory, plain volatile accesses: a warp's shared-memory instructions execute in program order
// in the SM's load/store pipe, so a flag store follows the warp's gathers of the stage
__device__ __forceinline__ void q_st_flag(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t q_ld_flag(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// packed fp32 pair accumulate: s1 += f, s2 += f * f  (FADD2 / FFMA2 on sm_100)
__device__ __forceinline__ void q_acc2(unsigned long long &s1, unsigned long long &s2, float fa, float fb) {
    unsigned long long f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(f) : "f"(fa), "f"(fb));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s1) : "l"(f));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s2) : "l"(f));
}
__device__ __forceinline__ float2 q_unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// ---------------------------------------------------------------------------------------------
// Geometry tables.  grid = (quads, groups of kVG views), one warp per block; lane = 4 consecutive voxels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_q_index(const QTiling tiling, const float *__restrict__ points, const float *__restrict__ proj,
          const float *__restrict__ depth, float voxel_z, int nv, int nvp, int n_quads, int64_t n_vox, int height,
          int width, int elt, uint32_t zero_off, uint16_t *__restrict__ off16, uint8_t *__restrict__ act,
          uint32_t *__restrict__ cntp, uint8_t *__restrict__ costp) {
    __shared__ float sp[kVG * 12];
    const int q = blockIdx.x, lane = threadIdx.x;
    const int v0 = blockIdx.y * kVG;
    const int nvg = min(kVG, nv - v0);
    for (int i = lane; i < nvg * 12; i += 32) sp[i] = proj[v0 * 12 + i];
    __syncwarp();
    const int64_t n0 = tiling.voxel(q, lane);
    float X[kQV], Y[kQV], Z[kQV];
    bool inside[kQV];
#pragma unroll
    for (int k = 0; k < kQV; ++k) {
        inside[k] = n0 + k < n_vox;
        const int64_t n = inside[k] ? n0 + k : 0;
        X[k] = __ldg(points + n);
        Y[k] = __ldg(points + n_vox + n);
        Z[k] = __ldg(points + 2 * n_vox + n);
    }
    uint32_t cnt = 0u;                                          // 4 packed uint8 counts
    int n_act = 0;                                              // views of this group that see the quad
    for (int i = 0; i < nvg; ++i) {
        uint32_t off[kQV];
        bool any = false;
#pragma unroll
        for (int k = 0; k < kQV; ++k) {
            float xr, yr, q2;
            bool ok = project_nearest(sp + i * 12, X[k], Y[k], Z[k], height, width, xr, yr, q2) && inside[k];
            const int pix = ok ? (int)yr * width + (int)xr : 0;
            if (ok && depth != nullptr) {                       // B4, nerfdet.py:405-411
                const float d = __ldg(depth + (int64_t)(v0 + i) * height * width + pix);
                ok = (q2 > __fsub_rn(d, voxel_z)) && (q2 < __fadd_rn(d, voxel_z));
            }
            off[k] = ok ? (uint32_t)pix * (uint32_t)elt : zero_off;
            cnt += ok ? (1u << (8 * k)) : 0u;
            any |= ok;
        }
        any = __any_sync(0xffffffffu, any);
        n_act += any ? 1 : 0;
        if (any) {
            uint2 w;
            w.x = off[0] | (off[1] << 16);
            w.y = off[2] | (off[3] << 16);
            *reinterpret_cast<uint2 *>(off16 + ((int64_t)(v0 + i) * n_quads + q) * kQuad + lane * kQV) = w;
        }
        if (lane == 0) act[(int64_t)(v0 + i) * n_quads + q] = any ? 1 : 0;
    }
    cntp[((int64_t)blockIdx.y * n_quads + q) * 32 + lane] = cnt;
    if (lane == 0) costp[(int64_t)blockIdx.y * n_quads + q] = (uint8_t)n_act;
}

// ---------------------------------------------------------------------------------------------
// Ranking: one block sorts the quads by the number of views that see them (descending, ties by index).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_q_rank(const uint8_t *__restrict__ costp, int nvg, int n_quads, int n_quads_pad, uint16_t *__restrict__ ranked) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint32_t *s_key = reinterpret_cast<uint32_t *>(s_dyn);                              // [n_quads_pad]
    const int tid = threadIdx.x;
    for (int q = tid; q < n_quads_pad; q += blockDim.x) {
        uint32_t key = 0xffffffffu;
        if (q < n_quads) {
            int c = 0;
            for (int g = 0; g < nvg; ++g) c += (int)__ldg(costp + (int64_t)g * n_quads + q);        // coalesced over q
            key = ((uint32_t)(255 - min(c, 255)) << 16) | (uint32_t)q;
        }
        s_key[q] = key;
    }
    __syncthreads();
    for (int k = 2; k <= n_quads_pad; k <<= 1) {                                        // bitonic sort, ascending
        for (int s = k >> 1; s > 0; s >>= 1) {
            for (int i = tid; i < n_quads_pad; i += blockDim.x) {
                const int l = i ^ s;
                if (l > i) {
                    const uint32_t x = s_key[i], y = s_key[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { s_key[i] = y; s_key[l] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int q = tid; q < n_quads; q += blockDim.x) ranked[q] = (uint16_t)(s_key[q] & 0xffffu);
}

// ---------------------------------------------------------------------------------------------
// Ownership and compaction.  grid = (stages per unit, parts, copy split).  Every block repeats the (small) per-part
// stage table, then compacts the offset rows of ITS stage and writes its column of the tables.
// ---------------------------------------------------------------------------------------------
struct QPackArgs {
    int nv, nvp, nvg, n_quads, n_quads_pad, n_parts, W, G, spu;
    int64_t part_rows;
    const uint16_t *off16;
    const uint8_t *act;
    const uint32_t *cntp;
    const uint16_t *ranked;   // [n_quads] quad ids by descending cost (k_q_rank)
    uint16_t *quadmap;        // [n_parts][W][4]
    uint32_t *hdr;            // [n_parts][W][spu]  mask (4 bits per view of the stage: slot s of view g = bit 4 g + s) | first row << 16
    uint32_t *nrows;          // [n_parts][spu]
    uint32_t *grow;           // [n_parts][spu]     first row of the stage in the part's stream
    uint32_t *cntc;           // [n_parts][W][4][32] packed view counts of the lane's 4 voxels
    uint16_t *offc;           // [n_parts][part_rows][128]
    unsigned int *tickets;    // [n_parts] unit tickets of the lift kernel, zeroed here
};

__global__ void __launch_bounds__(256)
k_q_pack(const QPackArgs a) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint16_t *s_mask = reinterpret_cast<uint16_t *>(s_dyn);                             // [W][spu]
    const int n2 = (a.W * a.spu + 1) & ~1;
    uint16_t *s_start = s_mask + n2;                                                    // [W][spu]
    uint32_t *s_tot = reinterpret_cast<uint32_t *>(s_start + n2);                       // [spu]
    uint32_t *s_grow = s_tot + a.spu;                                                   // [spu]
    __shared__ uint16_t s_quad[kQMaxWarps * kSlots];
    const int j = blockIdx.x, part = blockIdx.y, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int zi = blockIdx.z, zn = gridDim.z;                 // the copy work of a (stage, part) is split over zn blocks

    // 2. ownership: rank r goes to warp snake(r) of all n_parts * W warps, slot r / (n_parts * W)
    const int nwt = a.n_parts * a.W;
    if (tid < a.W * kSlots) {
        const int w = tid / kSlots, s = tid % kSlots;
        const int gw = w * a.n_parts + part;
        const int r = s * nwt + ((s & 1) ? nwt - 1 - gw : gw);
        s_quad[tid] = r < a.n_quads ? __ldg(a.ranked + r) : (uint16_t)0xffffu;
    }
    __syncthreads();
    // 3. masks of every (warp, stage) of the part
    for (int i = tid; i < a.W * a.spu; i += blockDim.x) {
        const int w = i / a.spu, jj = i - w * a.spu;
        uint32_t m = 0u;
        for (int g = 0; g < a.G; ++g) {
            const int v = jj * a.G + g;
            if (v >= a.nv) break;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const uint16_t q = s_quad[w * kSlots + s];
                if (q != 0xffffu && __ldg(a.act + (int64_t)v * a.n_quads + q)) m |= 1u << (4 * g + s);
            }
        }
        s_mask[i] = (uint16_t)m;
    }
    __syncthreads();
    // 4. first row of every warp inside its stage's block, rows per stage
    for (int jj = tid; jj < a.spu; jj += blockDim.x) {
        uint32_t st = 0u;
        for (int w = 0; w < a.W; ++w) {
            s_start[w * a.spu + jj] = (uint16_t)st;
            st += (uint32_t)__popc((uint32_t)s_mask[w * a.spu + jj]);
        }
        s_tot[jj] = st;
    }
    __syncthreads();
    if (warp == 0) {                                                                    // exclusive prefix over the stages
        uint32_t base = 0u;
        for (int j0 = 0; j0 < a.spu; j0 += 32) {
            const int jj = j0 + lane;
            const uint32_t n = jj < a.spu ? s_tot[jj] : 0u;
            uint32_t incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (jj < a.spu) s_grow[jj] = base + incl - n;
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    // 5. this block's column of the tables
    if (zi == 0 && tid < a.W)
        a.hdr[((int64_t)part * a.W + tid) * a.spu + j] = (uint32_t)s_mask[tid * a.spu + j] | ((uint32_t)s_start[tid * a.spu + j] << 16);
    if (zi == 0 && tid == 0) {
        a.nrows[(int64_t)part * a.spu + j] = s_tot[j];
        a.grow[(int64_t)part * a.spu + j] = s_grow[j];
    }
    if (j == 0) {
        if (zi == 0 && tid < a.W * kSlots) a.quadmap[(int64_t)part * a.W * kSlots + tid] = s_quad[tid];
        if (zi == 0 && tid == 0) a.tickets[part] = 0u;
        // view counts of the part's voxels in slot order (sum of the per-group partials; <= 254 per byte, no carries)
        for (int i = zi * (int)blockDim.x + tid; i < a.W * kSlots * 32; i += zn * (int)blockDim.x) {
            const uint16_t q = s_quad[i >> 5];
            uint32_t c = 0u;
            if (q != 0xffffu)
                for (int g = 0; g < a.nvg; ++g) c += __ldg(a.cntp + ((int64_t)g * a.n_quads + q) * 32 + (i & 31));
            a.cntc[(int64_t)part * a.W * kSlots * 32 + i] = c;
        }
    }
    // 6. offset rows of stage j: warp -> view -> slot, one block warp per row, 8 B per lane
    const int n_bw = (int)(blockDim.x >> 5);
    for (int w = zi * n_bw + warp; w < a.W; w += zn * n_bw) {
        const uint32_t m = (uint32_t)s_mask[w * a.spu + j];
        if (!m) continue;
        int64_t row = (int64_t)part * a.part_rows + s_grow[j] + s_start[w * a.spu + j];
        for (int g = 0; g < a.G; ++g) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                if (m & (1u << (4 * g + s))) {
                    const int v = j * a.G + g;
                    const uint16_t q = s_quad[w * kSlots + s];
                    const uint2 *src = reinterpret_cast<const uint2 *>(a.off16 + ((int64_t)v * a.n_quads + q) * kQuad);
                    uint2 *dst = reinterpret_cast<uint2 *>(a.offc + row * kQuad);
                    dst[lane] = __ldg(src + lane);
                    ++row;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Gather + statistics with the planes of one channel streamed through shared memory.
// ---------------------------------------------------------------------------------------------
struct QArgs {
    QTiling tiling;
    const uint32_t *hdr, *nrows, *grow, *cntc;
    const uint16_t *quadmap, *offc;
    int64_t part_rows;
    unsigned int *tickets;     // [n_parts] running ticket counters (never reset between launches)
    uint32_t ticket_base;      // tickets drawn per part by the earlier launches on this plan
    int nv, spu, W, n_parts, channels;
    int64_t n_vox;
    const void *feat;
    int64_t sv, sc;            // elements
    uint32_t plane_bytes, pitch;   // smem slot = plane + zero word, padded to `pitch`
    int S, R;                  // plane ring: S stages of kG slots; offset-row ring: R rows of 256 B
    int pf;                    // planes are prefetched into L2 this many stages ahead of their copy (0 = off)
    int n_views_total;
    const float *alpha;
    float *out_a, *out_b;
    int64_t *count_i64;
    float *count_f32;
};

// accumulate one quad (4 voxels per lane) into the registers of slot `slot` (warp-uniform)
__device__ __forceinline__ void q_acc_slot(int slot, unsigned long long (&s1)[kSlots * 2], unsigned long long (&s2)[kSlots * 2],
                                           const float *f) {
    switch (slot) {
        case 0: q_acc2(s1[0], s2[0], f[0], f[1]); q_acc2(s1[1], s2[1], f[2], f[3]); break;
        case 1: q_acc2(s1[2], s2[2], f[0], f[1]); q_acc2(s1[3], s2[3], f[2], f[3]); break;
        case 2: q_acc2(s1[4], s2[4], f[0], f[1]); q_acc2(s1[5], s2[5], f[2], f[3]); break;
        default: q_acc2(s1[6], s2[6], f[0], f[1]); q_acc2(s1[7], s2[7], f[2], f[3]); break;
    }
}

template <typename T>
__device__ __forceinline__ void q_gather4(uint32_t pb, const uint2 o, float *f) {
    f[0] = q_lds_elt<T>(pb + (o.x & 0xffffu));
    f[1] = q_lds_elt<T>(pb + (o.x >> 16));
    f[2] = q_lds_elt<T>(pb + (o.y & 0xffffu));
    f[3] = q_lds_elt<T>(pb + (o.y >> 16));
}

// rare paths of the epilogue, kept out of line: unaligned / partial quads and the view counts (one channel only)
__device__ __noinline__ void q_store_partial(float *out_a, float *out_b, int64_t o, int64_t left, float4 va, float4 vb) {
    const float oa[4] = {va.x, va.y, va.z, va.w}, ob[4] = {vb.x, vb.y, vb.z, vb.w};
    for (int t = 0; t < 4 && t < left; ++t) {
        out_a[o + t] = oa[t];
        if (out_b != nullptr) out_b[o + t] = ob[t];
    }
}
__device__ __noinline__ void q_store_counts(int64_t *count_i64, float *count_f32, int64_t nb, int64_t left, uint32_t cw) {
    for (int t = 0; t < 4 && t < left; ++t) {
        const int cn = (int)((cw >> (8 * t)) & 0xffu);
        if (count_i64 != nullptr) count_i64[nb + t] = (int64_t)cn;
        if (count_f32 != nullptr) count_f32[nb + t] = (float)cn;
    }
}

template <typename T, bool kRaw, int kG>
__global__ void __launch_bounds__((kQMaxWarps + 1) * 32, 1)
k_lift_quads(const QArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = a.W, S = a.S, R = a.R, spu = a.spu;
    const int part = blockIdx.x % a.n_parts;

    // planes [S][kG][pitch] | offset rows [R][256] | barriers [S] | progress flags [32] | unit queue [16] |
    // stage tables
    const uint32_t sm_base = q_smem_u32(smem);
    const uint32_t stage_pitch = (uint32_t)kG * a.pitch;
    const uint32_t row_base = sm_base + (uint32_t)S * stage_pitch;
    unsigned char *p_tab = smem + (size_t)S * stage_pitch + (size_t)R * kQRowBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(p_tab);
    uint32_t *s_flags = reinterpret_cast<uint32_t *>(bars + kQMaxStages);      // [32] stages done per warp
    uint32_t *s_units = s_flags + 32;                                           // [16] (ordinal + 1) << 16 | channel + 1 (0: end)
    uint32_t *s_nrows = s_units + 16;                                           // [spu]
    uint32_t *s_grow = s_nrows + spu;                                           // [spu]
    uint32_t *s_pos = s_grow + spu;                                             // [spu] first ring row of the stage (restarts at 0 every unit)
    uint2 *s_tab = reinterpret_cast<uint2 *>(s_pos + spu + (spu & 1));            // [W][spu] item list (4-bit codes), smem address of the warp's first row | items << 24
    uint4 *s_stage = reinterpret_cast<uint4 *>((reinterpret_cast<uintptr_t>(s_tab + W * spu) + 15) & ~(uintptr_t)15);   // [spu] producer: rows, first ring row, lag, first row in the stream
    float *s_rcp = reinterpret_cast<float *>(s_stage + spu);                    // [256] RN(1 / count), [0] = 0
    int32_t *s_qbase = reinterpret_cast<int32_t *>(s_rcp + 256);                // [W][4] first voxel of the quad (lane 0), -1: none
    const uint32_t bar_full = q_smem_u32(bars);
    const uint32_t f_progress = q_smem_u32(s_flags), f_units = q_smem_u32(s_units);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) q_mbar_init(bar_full + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32 + 16) s_flags[threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < S * kG; i += blockDim.x)                      // the zero word behind every plane slot
        *reinterpret_cast<uint32_t *>(smem + (size_t)i * a.pitch + a.plane_bytes) = 0u;
    // the plan's tables are complete when this kernel starts (their producer does not trigger early, see the header)
    for (int i = threadIdx.x; i < spu; i += blockDim.x) {
        s_nrows[i] = __ldg(a.nrows + (int64_t)part * spu + i);
        s_grow[i] = __ldg(a.grow + (int64_t)part * spu + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // ring rows of the stages: a stage's rows never wrap around the ring, and every unit starts at row 0, so that
        // the positions are the same for every unit (the producer checks overlaps with stages still in use)
        uint32_t head = 0;
        for (int j = 0; j < spu; ++j) {
            const uint32_t n = s_nrows[j];
            if (head + n > (uint32_t)R) head = 0;
            s_pos[j] = head;
            head += n;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < spu; j += blockDim.x) {
        // lag: how many of the preceding stages (cyclically: the positions repeat every unit) may still be in use
        // when stage j is written -- at most S - 1 (plane slots), fewer when their ring rows would be overwritten
        const uint32_t n = s_nrows[j], pos = s_pos[j];
        uint32_t lag = (uint32_t)S - 1u;
        for (int d = 1; d < S; ++d) {
            int jj = (j - d) % spu;
            if (jj < 0) jj += spu;
            const uint32_t nn = s_nrows[jj], pp = s_pos[jj];
            if (n != 0u && nn != 0u && pos < pp + nn && pp < pos + n) { lag = (uint32_t)d - 1u; break; }
        }
        s_stage[j] = make_uint4(n, pos, lag, s_grow[j]);
    }
    for (int i = threadIdx.x; i < W * spu; i += blockDim.x) {
        const uint32_t h = __ldg(a.hdr + (int64_t)part * W * spu + i);
        // the warp's items of the stage as a list of 4-bit codes (view g << 2 | slot s), in row order, and their number
        uint32_t desc = 0u, n = 0u;
        for (uint32_t m = h & 0xffffu; m != 0u; m &= m - 1u) desc |= (uint32_t)(__ffs((int)m) - 1) << (4 * n++);
        s_tab[i] = make_uint2(desc, (row_base + (s_pos[i % spu] + (h >> 16)) * kQRowBytes) | (n << 24));
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_rcp[i] = i ? __frcp_rn((float)i) : 0.0f;
    for (int i = threadIdx.x; i < W * kSlots; i += blockDim.x) {
        const uint16_t q = __ldg(a.quadmap + (int64_t)part * W * kSlots + i);
        s_qbase[i] = q != 0xffffu ? (int32_t)a.tiling.voxel((int)q, 0) : -1;
    }
    __syncthreads();

    if (warp == W) {
        // ---------------- producer warp ----------------
        unsigned int *ticket = a.tickets + part;
        auto pull = [&]() -> int {
            uint32_t t = 0u;
            if (lane == 0) t = atomicAdd(ticket, 1u) - a.ticket_base;
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t < (uint32_t)a.channels) return (int)t;
            // this CTA starts no further unit: once every CTA is here the next launch may fill the SMs that fall idle
            // (and finds the ticket counter at rest)
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
            return -1;
        };
        int ch = pull();
        int ch_next = ch >= 0 ? pull() : -1;
        const int64_t view_bytes = a.sv * (int64_t)sizeof(T), chan_bytes = a.sc * (int64_t)sizeof(T);
        const char *off_part = reinterpret_cast<const char *>(a.offc) + (int64_t)part * a.part_rows * kQRowBytes;
        uint32_t i = 0, released = 0;
        int slot = 0, k = 0;
        while (ch >= 0) {
            if (lane == 0) q_st_flag(f_units + 4 * (k & 15), ((uint32_t)(k + 1) << 16) | (uint32_t)(ch + 1));
            const char *psrc = reinterpret_cast<const char *>(a.feat) + (int64_t)ch * chan_bytes + (int64_t)lane * view_bytes;
            // The loop body is one table read, one compare and the copies: this warp runs ahead of 25 others on a
            // serial chain of dependent instructions, and everything it does per stage delays the refill of a slot.
            for (int j = 0; j < spu; ++j) {
                const uint4 st = s_stage[j];                                    // rows, first ring row, lag, first row in the stream
                // stages i-1 .. i-lag may still be in use (lag < S: the plane slot of stage i - S is reused, and the
                // ring rows of stage i overlap those of stage i - lag - 1 at the earliest)
                const uint32_t need = i > st.z ? i - st.z : 0u;
                if (released < need) {
                    do {                                                        // lane w reads the progress word of compute warp w
                        const uint32_t pr = lane < W ? q_ld_flag(f_progress + 4 * lane) : 0xffffffffu;
                        released = __reduce_min_sync(0xffffffffu, pr);
                    } while (released < need);                                  // busy poll: one LDS per look
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // their reads before our async-proxy writes
                }
                const int nvs = min(kG, a.nv - j * kG);
                const uint32_t fb = bar_full + 8 * slot;
                if (lane == 0) q_mbar_expect_tx(fb, (uint32_t)nvs * a.plane_bytes + st.x * kQRowBytes);
                if (lane < nvs)
                    q_bulk_g2s(sm_base + (uint32_t)(slot * kG + lane) * a.pitch, psrc, a.plane_bytes, fb);
                else if (lane == kG && st.x != 0u)
                    q_bulk_g2s(row_base + st.y * kQRowBytes, off_part + (size_t)st.w * kQRowBytes, st.x * kQRowBytes, fb);
                else if (a.pf > 0 && lane >= 16 && lane < 16 + kG) {
                    // optional: HBM -> L2 ahead of the ring, so that the copy into shared memory sees L2 latency
                    int jp = j + a.pf, cp = ch;
                    if (jp >= spu) { jp -= spu; cp = ch_next; }
                    const int vp = jp * kG + (lane - 16);
                    if (cp >= 0 && jp < spu && vp < a.nv)
                        q_prefetch_l2(reinterpret_cast<const char *>(a.feat) + (int64_t)cp * chan_bytes + (int64_t)vp * view_bytes,
                                      a.plane_bytes);
                }
                psrc += (int64_t)kG * view_bytes;
                ++i;
                if (++slot == S) slot = 0;
            }
            ++k;
            ch = ch_next;
            ch_next = ch >= 0 ? pull() : -1;
        }
        if (lane == 0) q_st_flag(f_units + 4 * (k & 15), (uint32_t)(k + 1) << 16);      // end of work
        return;
    }

    // ---------------- compute warps ----------------
    uint32_t gi = 0, par = 0;
    bool landed = false;                                                        // stage gi is known to have landed
    bool waited = false;                                                        // griddepcontrol.wait done
    const int lane_off = (int)(a.tiling.voxel(0, lane) - a.tiling.voxel(0, 0));
    // Loop state kept in registers (made opaque so that the compiler does not re-derive it from the thread index, the
    // shared-memory window or the constant bank inside the item loop): this warp's table row, progress word and lane
    // offset; barrier and plane base of the current slot, advanced incrementally.
    uint32_t my_progress = f_progress + 4 * warp, tab_addr = q_smem_u32(s_tab + warp * spu), lane8 = (uint32_t)lane * 8u;
    uint32_t pitch = a.pitch, bar = bar_full, sb = sm_base;
    const uint32_t bar_end = bar_full + 8u * (uint32_t)S, ring_bytes = (uint32_t)S * stage_pitch;
    asm volatile("" : "+r"(my_progress), "+r"(tab_addr), "+r"(lane8), "+r"(pitch), "+r"(bar), "+r"(sb));
    for (int k = 0;; ++k) {
        uint32_t e;
        do { e = q_ld_flag(f_units + 4 * (k & 15)); } while ((e >> 16) != (uint32_t)(k + 1));
        if ((e & 0xffffu) == 0u) break;
        unsigned long long s1[kSlots * 2], s2[kSlots * 2];
#pragma unroll
        for (int i = 0; i < kSlots * 2; ++i) { s1[i] = 0ull; s2[i] = 0ull; }

        uint32_t ta = tab_addr;
        for (int j = 0; j < spu; ++j, ta += 8) {
            const uint2 t = q_lds_u2(ta);
            if (!landed) q_mbar_wait(bar, par);
            int n = (int)(t.y >> 24);                                           // items of this warp in the stage
            if (n != 0) {                                                       // warp-uniform
                // Two items at a time, software-pipelined: the offset rows of the NEXT pair are requested while the 8
                // gathers of the current pair are in flight.  The cost is proportional to the number of active
                // (view, quad) pairs; nothing is paid for the inactive ones.
                uint32_t desc = t.x;
                uint32_t raddr = (t.y & 0xffffffu) + lane8;
                uint2 o0 = q_lds_u2(raddr), o1 = make_uint2(0u, 0u);
                if (n > 1) o1 = q_lds_u2(raddr + kQRowBytes);
                for (;;) {
                    const uint32_t d0 = desc & 15u, d1 = (desc >> 4) & 15u;
                    desc >>= 8;
                    const bool two = n > 1;
                    float f[8];
                    q_gather4<T>(sb + (d0 >> 2) * pitch, o0, f);
                    if (two) q_gather4<T>(sb + (d1 >> 2) * pitch, o1, f + 4);
                    n -= 2;
                    raddr += 2 * kQRowBytes;
                    if (n > 0) {
                        o0 = q_lds_u2(raddr);
                        if (n > 1) o1 = q_lds_u2(raddr + kQRowBytes);
                    }
                    q_acc_slot((int)(d0 & 3u), s1, s2, f);
                    if (two) q_acc_slot((int)(d1 & 3u), s1, s2, f + 4);
                    if (n <= 0) break;
                }
            }
            // next stage: look now, use the answer in the next iteration (a completed test costs ~90 cycles)
            bar += 8u;
            sb += stage_pitch;
            if (bar == bar_end) { bar = bar_full; sb -= ring_bytes; par ^= 1u; }
            landed = q_mbar_test(bar, par);
            __syncwarp();
            ++gi;
            if (lane == 0) q_st_flag(my_progress, gi);
        }
        const int c = (int)(q_ld_flag(f_units + 4 * (k & 15)) & 0xffffu) - 1;    // this unit's channel

        // ---------------- epilogue of unit (c, part); the producer is already streaming the next unit ----------------
        if (!waited) {                                                          // first global write of this launch
            asm volatile("griddepcontrol.wait;" ::: "memory");
            waited = true;
        }
        float *out_a = a.out_a, *out_b = a.out_b;
        const int64_t row = (int64_t)c * a.n_vox;
        const bool vec_ok = (row % 4 == 0) &&
                            ((reinterpret_cast<uintptr_t>(out_a) | reinterpret_cast<uintptr_t>(out_b)) % 16 == 0);
        const float nvt = (float)a.n_views_total;
        uint32_t cws[kSlots];                                                   // view counts of this lane's 16 voxels (4 loads in flight)
#pragma unroll
        for (int s = 0; s < kSlots; ++s) cws[s] = __ldg(a.cntc + (((int64_t)part * W + warp) * kSlots + s) * 32 + lane);
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            const int32_t qb = s_qbase[warp * kSlots + s];
            const int64_t nb = (int64_t)qb + lane_off;                          // first of this lane's 4 voxels of the quad
            if (qb < 0 || nb >= a.n_vox) continue;
            const uint32_t cw = cws[s];
            const bool full = vec_ok && nb + 4 <= a.n_vox && nb % 4 == 0;
            float4 al = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            if (!kRaw && a.alpha != nullptr) {
                if (full) {
                    al = __ldg(reinterpret_cast<const float4 *>(a.alpha + nb));
                } else {
                    al.x = __ldg(a.alpha + nb);
                    if (nb + 1 < a.n_vox) al.y = __ldg(a.alpha + nb + 1);
                    if (nb + 2 < a.n_vox) al.z = __ldg(a.alpha + nb + 2);
                    if (nb + 3 < a.n_vox) al.w = __ldg(a.alpha + nb + 3);
                }
            }
            const float2 a0 = q_unpack2(s1[2 * s]), a1 = q_unpack2(s1[2 * s + 1]);
            const float2 b0 = q_unpack2(s2[2 * s]), b1 = q_unpack2(s2[2 * s + 1]);
            const float v1[4] = {a0.x, a0.y, a1.x, a1.y};
            const float v2[4] = {b0.x, b0.y, b1.x, b1.y};
            const float alv[4] = {al.x, al.y, al.z, al.w};
            float oa[4], ob[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (kRaw) {
                    oa[t] = v1[t];
                    ob[t] = v2[t];
                } else {
                    const uint32_t cn = (cw >> (8 * t)) & 0xffu;
                    const float cf = (float)cn;                                 // count + 1e-8 == count in fp32
                    const float rc = s_rcp[cn];                                 // RN(1 / count); 0 for count 0
                    // correctly rounded S1 / count from the reciprocal (one Newton step on the quotient): the mean is
                    // bit-equal to the reference's IEEE divide for these integer divisors; S1 == 0 where count == 0
                    const float q0 = v1[t] * rc;
                    const float m = fmaf(fmaf(-q0, cf, v1[t]), rc, q0);         // nerfdet.py:175-176
                    float ssd = fmaxf(fmaf(-m, v1[t], v2[t]), 0.0f);            // sum over valid views of (f - m)^2
                    ssd = fmaf((nvt - cf) * m, m, ssd);                         // invalid views contribute m^2 each (nerfdet.py:179)
                    oa[t] = m * alv[t];
                    // exp(-var); ex2.approx, rel. error ~1e-7 (1 + var); exp(-1e6) == 0 where count == 0 (nerfdet.py:180-181)
                    ob[t] = cn != 0u ? exp2f(ssd * (rc * -1.4426950408889634f)) : 0.0f;
                }
            }
            const int64_t o = row + nb;
            if (full) {
                __stcs(reinterpret_cast<float4 *>(out_a + o), make_float4(oa[0], oa[1], oa[2], oa[3]));
                if (out_b != nullptr)
                    __stcs(reinterpret_cast<float4 *>(out_b + o), make_float4(ob[0], ob[1], ob[2], ob[3]));
            } else {
                q_store_partial(out_a, out_b, o, a.n_vox - nb, make_float4(oa[0], oa[1], oa[2], oa[3]),
                                make_float4(ob[0], ob[1], ob[2], ob[3]));
            }
            if (c == 0) q_store_counts(a.count_i64, a.count_f32, nb, a.n_vox - nb, cw);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
struct QGeom {
    QTiling tiling;
    int elt, n_pix, nv, nvp, nvg, n_quads, n_quads_pad, n_parts, W, G, spu, S, R, grid;
    int64_t part_rows;
    uint32_t plane_bytes, pitch;
    // plan layout (bytes from the start of the plan buffer): the tables the lift kernel reads, then build scratch
    size_t o_hdr, o_nrows, o_grow, o_quadmap, o_cntc, o_tickets, o_offc, o_off16, o_act, o_cntp, o_ranked, o_costp, total_bytes;
    size_t smem_bytes, pack_smem;
};

// Per-device facts, looked up once (idempotent caches of immutable values, not mutable state): the SM count and whether
// a kernel instantiation already carries its dynamic shared-memory attribute on the device.
constexpr int kQMaxDevices = 64;
static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kQMaxDevices ? dev : 0;
}
static int device_sm_count() {
    static std::atomic<int> cache[kQMaxDevices];
    const int dev = current_device();
    int sms = cache[dev].load(std::memory_order_relaxed);
    if (sms == 0) {
        sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cache[dev].store(sms, std::memory_order_relaxed);
    }
    return sms;
}
// dynamic shared memory attribute of kernel variant `slot` (0..15): set when the size grows on this device
static cudaError_t ensure_smem(const void *kern, int slot, size_t bytes) {
    static std::atomic<unsigned> have[kQMaxDevices][16];
    const int dev = current_device();
    if (have[dev][slot].load(std::memory_order_relaxed) >= bytes) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have[dev][slot].store((unsigned)bytes, std::memory_order_relaxed);
    return e;
}

static size_t q_fixed_smem(int W, int spu) {
    return (size_t)kQMaxStages * 8 + (32 + 16) * 4 + (size_t)spu * 16 + 16 + (size_t)4 * spu * 4 + (size_t)W * spu * 8 + 256 * 4 + (size_t)W * kSlots * 4 + 16;
}

static bool quad_geom(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt, QGeom &g) {
    g.elt = f->dtype == ND_F32 ? 4 : 2;
    g.n_pix = f->height * f->width;
    g.nv = f->n_views;
    if (f->stride_x != 1 || f->stride_y != f->width) return false;          // planes must be contiguous
    const int64_t pb = (int64_t)g.n_pix * g.elt;
    if (pb % 16 != 0 || pb + 4 > 65535) return false;                       // TMA granule; uint16 byte offsets
    if ((reinterpret_cast<uintptr_t>(f->data) & 15) != 0 || (f->stride_v * g.elt) % 16 != 0 ||
        (f->stride_c * g.elt) % 16 != 0)
        return false;
    if (f->n_views > 254) return false;                                     // uint8 view counts
    g.plane_bytes = (uint32_t)pb;
    g.pitch = (uint32_t)align_up((size_t)pb + 16, 128);
    g.nvg = (g.nv + kVG - 1) / kVG;
    g.nvp = g.nvg * kVG;
    g.tiling = QTiling{0, 0, 0, 0, 0};
    if (opt != nullptr && opt->grid_x > 0 && opt->grid_y > 0 && opt->grid_z > 0 &&
        (int64_t)opt->grid_x * opt->grid_y * opt->grid_z == n_vox && opt->grid_z % kQV == 0 &&
        opt->grid_x % kQBx == 0 && opt->grid_y % kQBy == 0) {
        g.tiling.compact = 1;
        g.tiling.gy = opt->grid_y;
        g.tiling.gz = opt->grid_z;
        g.tiling.nzq = opt->grid_z / kQV;
        g.tiling.tiles_y = opt->grid_y / kQBy;
    }
    const int64_t nq = ceil_div(n_vox, kQuad);
    if (nq > kQMaxQuads) return false;
    g.n_quads = (int)nq;
    g.n_quads_pad = 32;
    while (g.n_quads_pad < g.n_quads) g.n_quads_pad <<= 1;
    const int nwt = (int)ceil_div(g.n_quads, kSlots);
    g.n_parts = (int)ceil_div(nwt, kQMaxWarps);
    g.W = (int)ceil_div(nwt, g.n_parts);
    // views per stage and ring depth: as many plane slots as fit beside an offset-row ring that holds the worst case of
    // one stage (every quad of the part active in every view of the stage) and 40 % of that (the typical activity) per
    // further stage
    int G = 2, S = 0;
    if (opt != nullptr && (opt->views_per_stage == 1 || opt->views_per_stage == 2)) G = opt->views_per_stage;
    if (opt != nullptr && opt->stages >= 2 && opt->stages <= kQMaxStages) S = opt->stages;
    const size_t cap = (size_t)227 * 1024;
    for (;; G >>= 1) {
        g.G = G;
        g.spu = (g.nv + G - 1) / G;
        const size_t fixed = q_fixed_smem(g.W, g.spu);
        const size_t worst = (size_t)g.W * kSlots * G;                      // rows of one stage, worst case
        int s = S > 0 ? S : kQMaxStages;
        for (; s >= 2; --s) {
            const size_t planes = (size_t)s * G * g.pitch;
            if (planes + fixed + worst * kQRowBytes > cap) continue;
            const size_t rows = (cap - planes - fixed) / kQRowBytes;
            if (S > 0 || rows >= worst + (size_t)(s - 1) * worst * 2 / 5 || s == 2) {
                g.S = s;
                g.R = (int)std::min<size_t>(rows, 4095);
                break;
            }
        }
        if (s >= 2) break;
        if (G == 1) return false;
    }
    g.smem_bytes = (size_t)g.S * g.G * g.pitch + (size_t)g.R * kQRowBytes + q_fixed_smem(g.W, g.spu);
    int sms = device_sm_count();
    if (opt != nullptr && opt->sm_limit > 0 && opt->sm_limit < sms) sms = opt->sm_limit;
    const int64_t n_units = (int64_t)f->channels * g.n_parts;
    int64_t grid = n_units < sms ? n_units : sms;
    grid -= grid % g.n_parts;
    if (grid < g.n_parts) grid = g.n_parts;
    g.grid = (int)grid;
    g.part_rows = (int64_t)g.nv * g.W * kSlots;                             // upper bound of a part's stream
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += align_up(bytes, 256); return at; };
    g.o_hdr = take((size_t)g.n_parts * g.W * g.spu * 4);
    g.o_nrows = take((size_t)g.n_parts * g.spu * 4);
    g.o_grow = take((size_t)g.n_parts * g.spu * 4);
    g.o_quadmap = take((size_t)g.n_parts * g.W * kSlots * 2);
    g.o_cntc = take((size_t)g.n_parts * g.W * kSlots * 32 * 4);
    g.o_tickets = take((size_t)g.n_parts * 4);
    g.o_offc = take((size_t)g.n_parts * (size_t)g.part_rows * kQRowBytes);
    g.o_off16 = take((size_t)g.nv * g.n_quads * kQRowBytes);
    g.o_act = take((size_t)g.n_quads * g.nvp);                        // [nv][n_quads]
    g.o_cntp = take((size_t)g.nvg * g.n_quads * 32 * 4);
    g.o_ranked = take((size_t)g.n_quads * 2);
    g.o_costp = take((size_t)g.nvg * g.n_quads);
    g.total_bytes = o;
    g.pack_smem = (size_t)2 * ((g.W * g.spu + 1) & ~1) * 2 + (size_t)2 * g.spu * 4;
    return g.pack_smem <= 200 * 1024;
}

bool lift_quads_eligible(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt) {
    QGeom g;
    return n_vox > 0 && quad_geom(f, n_vox, opt, g);
}

size_t lift_quads_plan_bytes(const nd_maps *f, int64_t n_vox, const nd_lift_options *opt) {
    QGeom g;
    if (n_vox <= 0 || !quad_geom(f, n_vox, opt, g)) return 0;
    return g.total_bytes;
}

template <typename... KArgs, typename... Args>
static cudaError_t q_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
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

nd_status lift_quads_plan_build(const nd_maps *f, const float *points, const float *proj, int64_t n_vox,
                                const float *depth, float voxel_z, void *plan, size_t plan_bytes,
                                const nd_lift_options *opt, cudaStream_t st) {
    QGeom g;
    ND_REQUIRE(quad_geom(f, n_vox, opt, g), ND_ERR_BAD_ARG, "lift plan: input not eligible for the plane-resident path");
    ND_REQUIRE(plan != nullptr && plan_bytes >= g.total_bytes, ND_ERR_WORKSPACE, "lift plan: buffer too small (%zu < %zu bytes)",
               plan_bytes, g.total_bytes);
    ND_REQUIRE((reinterpret_cast<uintptr_t>(plan) % 256) == 0, ND_ERR_BAD_ALIGNMENT, "lift plan: buffer not 256-byte aligned");
    char *b = reinterpret_cast<char *>(plan);
    uint16_t *off16 = reinterpret_cast<uint16_t *>(b + g.o_off16);
    uint8_t *act = reinterpret_cast<uint8_t *>(b + g.o_act);
    uint32_t *cntp = reinterpret_cast<uint32_t *>(b + g.o_cntp);
    uint8_t *costp = reinterpret_cast<uint8_t *>(b + g.o_costp);
    k_q_index<<<dim3((unsigned)g.n_quads, (unsigned)g.nvg), 32, 0, st>>>(g.tiling, points, proj, depth, voxel_z, g.nv, g.nvp,
                                                                         g.n_quads, n_vox, f->height, f->width, g.elt,
                                                                         g.plane_bytes, off16, act, cntp, costp);
    ND_CUDA_LAUNCH_CHECK("k_q_index");
    uint16_t *ranked = reinterpret_cast<uint16_t *>(b + g.o_ranked);
    k_q_rank<<<1, g.n_quads_pad >= 1024 ? 1024 : std::max(g.n_quads_pad, 64), (size_t)g.n_quads_pad * 4, st>>>(
        costp, g.nvg, g.n_quads, g.n_quads_pad, ranked);
    ND_CUDA_LAUNCH_CHECK("k_q_rank");
    QPackArgs pa{};
    pa.nv = g.nv; pa.nvp = g.nvp; pa.nvg = g.nvg; pa.n_quads = g.n_quads; pa.n_quads_pad = g.n_quads_pad;
    pa.n_parts = g.n_parts; pa.W = g.W; pa.G = g.G; pa.spu = g.spu; pa.part_rows = g.part_rows;
    pa.off16 = off16; pa.act = act; pa.cntp = cntp; pa.ranked = ranked;
    pa.quadmap = reinterpret_cast<uint16_t *>(b + g.o_quadmap);
    pa.hdr = reinterpret_cast<uint32_t *>(b + g.o_hdr);
    pa.nrows = reinterpret_cast<uint32_t *>(b + g.o_nrows);
    pa.grow = reinterpret_cast<uint32_t *>(b + g.o_grow);
    pa.cntc = reinterpret_cast<uint32_t *>(b + g.o_cntc);
    pa.offc = reinterpret_cast<uint16_t *>(b + g.o_offc);
    pa.tickets = reinterpret_cast<unsigned int *>(b + g.o_tickets);
    if (g.pack_smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_q_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.pack_smem);
        if (e != cudaSuccess) {
            set_error("k_q_pack: cannot reserve %zu bytes of shared memory: %s", g.pack_smem, cudaGetErrorString(e));
            return ND_ERR_CUDA;
        }
    }
    k_q_pack<<<dim3((unsigned)g.spu, (unsigned)g.n_parts, 4u), 256, g.pack_smem, st>>>(pa);
    ND_CUDA_LAUNCH_CHECK("k_q_pack");
    return ND_OK;
}

template <typename T, bool kRaw>
nd_status lift_quads_run(const nd_maps *f, const void *plan, size_t plan_bytes, int64_t n_vox, uint32_t launch_index,
                         int n_views_total, const float *alpha, float *out_a, float *out_b, int64_t *count_i64,
                         float *count_f32, const nd_lift_options *opt, cudaStream_t st) {
    QGeom g;
    ND_REQUIRE(quad_geom(f, n_vox, opt, g), ND_ERR_BAD_ARG, "lift: input not eligible for the plane-resident path");
    ND_REQUIRE(plan != nullptr && plan_bytes >= g.total_bytes, ND_ERR_WORKSPACE, "lift: plan buffer too small (%zu < %zu bytes)",
               plan_bytes, g.total_bytes);
    ND_REQUIRE((reinterpret_cast<uintptr_t>(plan) % 256) == 0, ND_ERR_BAD_ALIGNMENT, "lift: plan buffer not 256-byte aligned");
    char *b = const_cast<char *>(reinterpret_cast<const char *>(plan));
    QArgs a{};
    a.tiling = g.tiling;
    a.hdr = reinterpret_cast<const uint32_t *>(b + g.o_hdr);
    a.nrows = reinterpret_cast<const uint32_t *>(b + g.o_nrows);
    a.grow = reinterpret_cast<const uint32_t *>(b + g.o_grow);
    a.cntc = reinterpret_cast<const uint32_t *>(b + g.o_cntc);
    a.quadmap = reinterpret_cast<const uint16_t *>(b + g.o_quadmap);
    a.offc = reinterpret_cast<const uint16_t *>(b + g.o_offc);
    a.part_rows = g.part_rows;
    // tickets: every launch draws channels + (CTAs of the part) tickets per part -- every CTA draws exactly one ticket
    // past the end -- and a launch starts drawing only after its predecessor's CTAs have all drawn their last one
    // (that is when they trigger the dependent launch), so one running counter per part serves all launches
    a.tickets = reinterpret_cast<unsigned int *>(b + g.o_tickets);
    a.ticket_base = launch_index * (uint32_t)(f->channels + g.grid / g.n_parts);
    a.nv = g.nv; a.spu = g.spu; a.W = g.W; a.n_parts = g.n_parts; a.channels = f->channels;
    a.n_vox = n_vox;
    a.feat = f->data; a.sv = f->stride_v; a.sc = f->stride_c;
    a.plane_bytes = g.plane_bytes; a.pitch = g.pitch; a.S = g.S; a.R = g.R;
    a.pf = opt != nullptr && opt->prefetch_stages != 0 ? (opt->prefetch_stages > 0 ? std::min(opt->prefetch_stages, g.spu) : 0) : 0;
    a.n_views_total = n_views_total > 0 ? n_views_total : g.nv;
    a.alpha = alpha; a.out_a = out_a; a.out_b = out_b; a.count_i64 = count_i64; a.count_f32 = count_f32;
    void (*kern)(const QArgs) = nullptr;
    switch (g.G) {
        case 1: kern = k_lift_quads<T, kRaw, 1>; break;
        default: kern = k_lift_quads<T, kRaw, 2>; break;
    }
    const int variant = (sizeof(T) == 4 ? 0 : 1) * 4 + (kRaw ? 2 : 0) + (g.G == 1 ? 0 : 1);
    cudaError_t e = ensure_smem(reinterpret_cast<const void *>(kern), variant, g.smem_bytes);
    if (e != cudaSuccess) {
        set_error("k_lift_quads: cannot reserve %zu bytes of shared memory: %s", g.smem_bytes, cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    e = q_launch_pdl(kern, dim3((unsigned)g.grid), dim3((unsigned)(g.W + 1) * 32), g.smem_bytes, st, a);
    if (e != cudaSuccess) {
        set_error("k_lift_quads: CUDA error %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    return ND_OK;
}

#define ND_INSTANTIATE_QUADS(T, R)                                                                                     \
    template nd_status lift_quads_run<T, R>(const nd_maps *, const void *, size_t, int64_t, uint32_t, int, const float *, \
                                            float *, float *, int64_t *, float *, const nd_lift_options *, cudaStream_t);
ND_INSTANTIATE_QUADS(float, false)
ND_INSTANTIATE_QUADS(float, true)
ND_INSTANTIATE_QUADS(__nv_bfloat16, false)
ND_INSTANTIATE_QUADS(__nv_bfloat16, true)

}  // namespace nd
