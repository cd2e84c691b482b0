// Fast path of the fused lift for the reference's own layout: fp32 NCHW feature maps whose
// sliced planes are contiguous (Wf == full width), C a multiple of 32.
//
// Measured on the B200 at hand (tools/microbench*.cu, profiles/):  L2-hit reads run at
// 17-30 TB/s, HBM reads at ~6.5 TB/s, but stores into L2 only at ~6 TB/s and a copy
// (HBM read + L2-resident store) at ~3.5 TB/s each way.  The NCHW -> pixel-major staging
// pass is therefore the expensive part and the gather (pure L2-hit reads) is cheap, so the
// schedule overlaps them:
//
//   phase 0      : pixel-index table  ||  stage(chunk 0)
//   phase k=1..K-1: gather(chunk k-1) ||  stage(chunk k)          (one launch per phase,
//   phase K      : gather(chunk K-1)                                CTA-level role split)
//
// with 32-channel chunks (one 128-byte row per pixel) in two ping-pong staging buffers of
// nv*P*128 bytes each (2 x 30 MB at nv=50), which stay resident in the 126 MB L2.
// Kernel boundaries are the only grid-wide barriers; the hardware CTA scheduler balances
// the two roles.
#include "nd_common.cuh"

namespace nd {

constexpr int kFastThreads = 512;
constexpr int kFastChunk = 32;          // channels per chunk = lanes of a warp
constexpr int kStagePix = 256;          // pixels per staging tile
constexpr int kGatherVox = kFastThreads / 32;   // voxels per gather CTA (one per warp)
constexpr int kRowsInFlight = 8;

struct FastArgs {
    // ---- index role (phase 0) ----
    const float *points;
    const float *proj;
    int nv, nvp;
    int64_t n_vox;
    int height, width;
    int32_t *pix;               // [N][nvp]
    int n_index_blocks;
    // ---- stage role ----
    const float *feat;
    int64_t sv, sc;
    int n_pix;
    int stage_c0;               // first channel of the chunk being staged
    float *stage_dst;
    int tiles_per_view;
    int n_stage_blocks;
    // ---- gather role ----
    const float *stage_src;
    int gather_c0;
    int n_gather_blocks;
    int n_views_total;
    const float *alpha;
    float *out_a, *out_b;
    int64_t *count_i64;
    float *count_f32;
};

// -------------------------------------------------------------------------------------
__device__ __forceinline__ void role_index(const FastArgs &a, int block, float *smem) {
    // smem: projection matrices [nv][12]
    for (int i = threadIdx.x; i < a.nv * 12; i += blockDim.x) smem[i] = a.proj[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_total = a.n_index_blocks * (kFastThreads / 32);
    for (int64_t n = (int64_t)block * (kFastThreads / 32) + warp; n < a.n_vox; n += warps_total) {
        const float X = __ldg(a.points + n), Y = __ldg(a.points + a.n_vox + n), Z = __ldg(a.points + 2 * a.n_vox + n);
        for (int vb = 0; vb < a.nvp; vb += 32) {
            const int v = vb + lane;
            int32_t out = -1;
            if (v < a.nv) {
                float xr, yr, q2;
                if (project_nearest(smem + v * 12, X, Y, Z, a.height, a.width, xr, yr, q2))
                    out = (int32_t)yr * a.width + (int32_t)xr;
            }
            a.pix[n * a.nvp + v] = out;
        }
    }
}

// -------------------------------------------------------------------------------------
// 32 channels x 256 pixels: coalesced 512 B reads per warp, XOR-swizzled float4 tile in
// shared memory (conflict-free both ways), 4x4 register transpose, full 128 B row stores.
__device__ __forceinline__ void role_stage(const FastArgs &a, int tile_id, float4 *tile) {
    const int v = tile_id / a.tiles_per_view;
    const int p0 = (tile_id - v * a.tiles_per_view) * kStagePix;
    const int tid = threadIdx.x;
    const float *plane0 = a.feat + (int64_t)v * a.sv + (int64_t)a.stage_c0 * a.sc;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = tid + kFastThreads * r;
        const int ch = i >> 6, c4 = i & 63;
        const int p = p0 + 4 * c4;
        const float *src = plane0 + (int64_t)ch * a.sc + p;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p + 3 < a.n_pix) {
            val = ld_stream_f4(reinterpret_cast<const float4 *>(src));
        } else if (p < a.n_pix) {
            val.x = ld_stream(src);
            if (p + 1 < a.n_pix) val.y = ld_stream(src + 1);
            if (p + 2 < a.n_pix) val.z = ld_stream(src + 2);
        }
        tile[ch * 64 + (c4 ^ ((ch >> 2) & 7))] = val;
    }
    __syncthreads();
    const int cq = tid & 7, c4 = tid >> 3;
    const float4 r0 = tile[(4 * cq + 0) * 64 + (c4 ^ cq)];
    const float4 r1 = tile[(4 * cq + 1) * 64 + (c4 ^ cq)];
    const float4 r2 = tile[(4 * cq + 2) * 64 + (c4 ^ cq)];
    const float4 r3 = tile[(4 * cq + 3) * 64 + (c4 ^ cq)];
    const int p = p0 + 4 * c4;
    float4 *dst = reinterpret_cast<float4 *>(a.stage_dst + ((int64_t)v * a.n_pix + p) * kFastChunk) + cq;
    if (p + 0 < a.n_pix) dst[0] = make_float4(r0.x, r1.x, r2.x, r3.x);
    if (p + 1 < a.n_pix) dst[8] = make_float4(r0.y, r1.y, r2.y, r3.y);
    if (p + 2 < a.n_pix) dst[16] = make_float4(r0.z, r1.z, r2.z, r3.z);
    if (p + 3 < a.n_pix) dst[24] = make_float4(r0.w, r1.w, r2.w, r3.w);
}

// -------------------------------------------------------------------------------------
// One warp per voxel, lane = channel of the chunk.  Statistics about the shift K = first
// valid sample (see lift.cu) in three registers per lane.
template <bool kRaw>
__device__ __forceinline__ void role_gather(const FastArgs &a, int tile_id, float *smem) {
    float *st_a = smem;                                   // [32][kGatherVox + 1]
    float *st_b = smem + kFastChunk * (kGatherVox + 1);
    int *s_cnt = reinterpret_cast<int *>(smem + 2 * kFastChunk * (kGatherVox + 1));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n0 = (int64_t)tile_id * kGatherVox;
    const int64_t n = n0 + warp;
    float s1 = 0.f, s2 = 0.f, k0 = 0.f;
    int cnt = 0;
    if (n < a.n_vox) {
        const int32_t *prow = a.pix + n * a.nvp;
        const float *base = a.stage_src + lane;
        bool first = true;
        for (int vb = 0; vb < a.nv; vb += 64) {
            // two batches of 32 views per round trip to the index table
            const int i0 = (vb + lane < a.nv) ? __ldg(prow + vb + lane) : -1;
            const int i1 = (vb + 32 + lane < a.nv) ? __ldg(prow + vb + 32 + lane) : -1;
            // row offset (in floats) of this lane's view: ((v * P) + pix) * 32
            const int64_t o0 = ((int64_t)(vb + lane) * a.n_pix + i0) * kFastChunk;
            const int64_t o1 = ((int64_t)(vb + 32 + lane) * a.n_pix + i1) * kFastChunk;
            unsigned m0 = __ballot_sync(0xffffffffu, i0 >= 0);
            unsigned m1 = __ballot_sync(0xffffffffu, i1 >= 0);
            cnt += __popc(m0) + __popc(m1);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                unsigned m = half ? m1 : m0;
                const int64_t off = half ? o1 : o0;
                while (m) {
                    float f[kRowsInFlight];
                    bool has[kRowsInFlight];
#pragma unroll
                    for (int u = 0; u < kRowsInFlight; ++u) {
                        has[u] = (m != 0);
                        const int b = (__ffs(m) - 1) & 31;
                        m &= (m - 1);
                        const int64_t o = __shfl_sync(0xffffffffu, off, b);
                        f[u] = has[u] ? __ldg(base + o) : 0.f;
                    }
                    if (first) { k0 = f[0]; first = false; }
#pragma unroll
                    for (int u = 0; u < kRowsInFlight; ++u) {
                        if (has[u]) {
                            const float d = f[u] - k0;
                            s1 += d;
                            s2 = fmaf(d, d, s2);
                        }
                    }
                }
            }
        }
    }
    const float cf = (float)cnt;
    float oa, ob;
    if (kRaw) {
        oa = fmaf(cf, k0, s1);
        ob = fmaf(cf * k0, k0, fmaf(2.0f * k0, s1, s2));
    } else if (cnt > 0) {
        const float md = s1 / cf;
        const float mean = k0 + md;
        float ssd = fmaxf(fmaf(-md, s1, s2), 0.0f);
        ssd = fmaf((float)(a.n_views_total - cnt) * mean, mean, ssd);
        float al = 1.0f;
        if (a.alpha != nullptr) al = __ldg(a.alpha + n);
        oa = mean * al;
        ob = expf(-(ssd / cf));
    } else {
        oa = 0.f;
        ob = 0.f;
    }
    st_a[lane * (kGatherVox + 1) + warp] = oa;
    st_b[lane * (kGatherVox + 1) + warp] = ob;
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    // 32 channels x 16 voxels: each half-warp stores one 64 B row segment
    {
        const int vx = threadIdx.x & (kGatherVox - 1);
        const int ch = threadIdx.x / kGatherVox;          // 0..31
        const int64_t nn = n0 + vx;
        if (nn < a.n_vox) {
            const int64_t o = (int64_t)(a.gather_c0 + ch) * a.n_vox + nn;
            st_stream(a.out_a + o, st_a[ch * (kGatherVox + 1) + vx]);
            if (a.out_b != nullptr) st_stream(a.out_b + o, st_b[ch * (kGatherVox + 1) + vx]);
            if (a.gather_c0 == 0 && ch == 0) {
                if (a.count_i64 != nullptr) a.count_i64[nn] = (int64_t)s_cnt[vx];
                if (a.count_f32 != nullptr) a.count_f32[nn] = (float)s_cnt[vx];
            }
        }
    }
}

template <bool kRaw>
__global__ void __launch_bounds__(kFastThreads, 2)
k_lift_phase(const FastArgs a) {
    extern __shared__ float4 dyn_smem[];
    int b = blockIdx.x;
    if (b < a.n_index_blocks) {
        role_index(a, b, reinterpret_cast<float *>(dyn_smem));
        return;
    }
    b -= a.n_index_blocks;
    // interleave the two roles evenly over the launch order
    const int64_t total = (int64_t)a.n_stage_blocks + a.n_gather_blocks;
    const int t_before = (int)(((int64_t)b * a.n_stage_blocks) / total);
    const int t_after = (int)(((int64_t)(b + 1) * a.n_stage_blocks) / total);
    if (t_after > t_before)
        role_stage(a, t_before, dyn_smem);
    else
        role_gather<kRaw>(a, b - t_before, reinterpret_cast<float *>(dyn_smem));
}

// -------------------------------------------------------------------------------------
bool lift_fast_eligible(const nd_maps *f, size_t budget_bytes) {
    if (f->dtype != ND_F32) return false;
    if (f->stride_x != 1 || f->stride_y != f->width) return false;
    if (f->channels % kFastChunk != 0 || f->channels < kFastChunk) return false;
    if ((reinterpret_cast<uintptr_t>(f->data) & 15) != 0 || (f->stride_v & 3) != 0 || (f->stride_c & 3) != 0) return false;
    if (f->n_views * 12 * sizeof(float) > 32 * 1024) return false;
    const size_t stage = (size_t)f->n_views * f->height * f->width * kFastChunk * sizeof(float);
    return 2 * stage <= budget_bytes;
}

size_t lift_fast_workspace_bytes(const nd_maps *f, int64_t n_vox) {
    const size_t nvp = align_up((size_t)f->n_views, 32);
    const size_t pix = align_up((size_t)n_vox * nvp * sizeof(int32_t), 256);
    const size_t stage = align_up((size_t)f->n_views * f->height * f->width * kFastChunk * sizeof(float), 256);
    const size_t n_buf = f->channels > kFastChunk ? 2 : 1;
    return pix + n_buf * stage;
}

template <bool kRaw>
nd_status run_lift_fast(const nd_maps *f, const float *points, const float *proj, int64_t n_vox, const float *alpha,
                        float *out_a, float *out_b, int64_t *count_i64, float *count_f32, void *ws, size_t ws_bytes,
                        cudaStream_t st) {
    const size_t need = lift_fast_workspace_bytes(f, n_vox);
    ND_REQUIRE(ws != nullptr && ws_bytes >= need, ND_ERR_WORKSPACE, "lift: workspace too small (%zu < %zu bytes)",
               ws_bytes, need);
    ND_REQUIRE((reinterpret_cast<uintptr_t>(ws) % 256) == 0, ND_ERR_BAD_ALIGNMENT, "lift: workspace not 256-byte aligned");
    const int nv = f->n_views;
    const int nvp = (int)align_up((size_t)nv, 32);
    const int n_pix = f->height * f->width;
    const size_t pix_bytes = align_up((size_t)n_vox * nvp * sizeof(int32_t), 256);
    const size_t stage_bytes = align_up((size_t)nv * n_pix * kFastChunk * sizeof(float), 256);
    char *wsb = reinterpret_cast<char *>(ws);
    float *stage[2] = {reinterpret_cast<float *>(wsb + pix_bytes), reinterpret_cast<float *>(wsb + pix_bytes + stage_bytes)};
    const int n_chunks = f->channels / kFastChunk;
    if (n_chunks == 1) stage[1] = stage[0];

    constexpr size_t kSmem = 32 * 1024;      // stage tile; the other roles use less
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_lift_phase<kRaw>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
        attr_done = true;
    }

    FastArgs a{};
    a.points = points;
    a.proj = proj;
    a.nv = nv;
    a.nvp = nvp;
    a.n_vox = n_vox;
    a.height = f->height;
    a.width = f->width;
    a.pix = reinterpret_cast<int32_t *>(wsb);
    a.feat = reinterpret_cast<const float *>(f->data);
    a.sv = f->stride_v;
    a.sc = f->stride_c;
    a.n_pix = n_pix;
    a.tiles_per_view = (int)ceil_div(n_pix, kStagePix);
    a.n_views_total = nv;
    a.alpha = alpha;
    a.out_a = out_a;
    a.out_b = out_b;
    a.count_i64 = count_i64;
    a.count_f32 = count_f32;
    const int stage_blocks = a.tiles_per_view * nv;
    const int gather_blocks = (int)ceil_div(n_vox, kGatherVox);
    const int index_blocks = (int)(ceil_div(n_vox, kFastThreads / 32) < 148 * 2 ? ceil_div(n_vox, kFastThreads / 32) : 148 * 2);

    for (int phase = 0; phase <= n_chunks; ++phase) {
        a.n_index_blocks = phase == 0 ? index_blocks : 0;
        a.n_stage_blocks = phase < n_chunks ? stage_blocks : 0;
        a.n_gather_blocks = phase > 0 ? gather_blocks : 0;
        a.stage_c0 = phase * kFastChunk;
        a.stage_dst = stage[phase & 1];
        a.gather_c0 = (phase - 1) * kFastChunk;
        a.stage_src = stage[(phase - 1) & 1];
        const unsigned grid = (unsigned)(a.n_index_blocks + a.n_stage_blocks + a.n_gather_blocks);
        k_lift_phase<kRaw><<<grid, kFastThreads, kSmem, st>>>(a);
        ND_CUDA_LAUNCH_CHECK("k_lift_phase");
    }
    return ND_OK;
}

template nd_status run_lift_fast<false>(const nd_maps *, const float *, const float *, int64_t, const float *, float *,
                                        float *, int64_t *, float *, void *, size_t, cudaStream_t);
template nd_status run_lift_fast<true>(const nd_maps *, const float *, const float *, int64_t, const float *, float *,
                                       float *, int64_t *, float *, void *, size_t, cudaStream_t);

}  // namespace nd
