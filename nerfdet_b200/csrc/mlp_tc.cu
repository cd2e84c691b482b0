// The shared NeRF / geometry MLP on the 5th-generation tensor cores (row M of SURVEY.md section 8a;
// reference nerf_mlp.py:11-234 as instantiated at nerfdet.py:62-69).  Two precisions of the same kernel:
//   bf16    bf16 operands, fp32 accumulation in tensor memory: the path BASELINE.json's 1e-2 tolerance applies to;
//   split   fp32-grade ("3 x bf16"): every operand is carried as hi + lo, two bf16 numbers (16 mantissa bits), and every
//           product as hi*hi + lo*hi + hi*lo on the tensor cores (fp32 accumulation; the dropped lo*lo term and the
//           residuals are ~2^-16 relative), the input encoding uses the reference's own argument rounding.  This is the
//           default, 1e-4 path; the FFMA kernel in mlp.cu remains for architectures this kernel does not take.
//
// One persistent CTA per SM walks tiles of 128 points through the whole network:
//
//   activations   never leave the SM, and the hidden activations never leave TENSOR MEMORY: the epilogue of
//                 layer i reads its fp32 accumulator row (tcgen05.ld), applies bias / relu, packs to bf16 and
//                 stores the result back IN PLACE (tcgen05.st: the 8 packed columns of K step s over the first
//                 8 of the 16 accumulator columns they came from), where layer i + 1 reads it as the A operand
//                 of tcgen05.mma (A-from-TMEM form).  Shared-memory bandwidth, the limiter of the first version
//                 (A re-read from shared memory for every MMA), is left to the weights alone.
//                 IN [128][192] bf16 (posenc 63 | features | zero pad to 144 | view encoding 27 at column
//                 144) lives in shared memory as K-major 128-byte-swizzled blocks of 64 columns (A-from-smem).
//   weights       nd_pack_mlp_weights_tc stores every (layer, 64-column K block) as the exact shared-memory
//                 image of a K-major SWIZZLE_128B B operand, so the producer warp streams one block with ONE
//                 cp.async.bulk (TMA) into a 5-stage mbarrier ring; the weights stay L2-resident (0.8 MB).
//   accumulators  two 128 x 256 fp32 regions in TMEM (all 512 columns).  Jobs (layers) alternate between
//                 them, so the epilogue of layer i runs while layer i + 1 already accumulates into the other
//                 region: the epilogue publishes its A operand in chunks of 64 K columns and the MMA warp
//                 starts the next layer's K block j as soon as chunk j is there.
//   roles         warps 0-7 epilogue (TMEM lane quarter = warp % 4, 32-column half of a chunk = warp / 4),
//                 warps 8-11 input encoding of the NEXT tile (held in registers until the IN buffer is
//                 released), warp 12 weight producer, warp 13 MMA issuer (one elected thread), TMEM owner.
//   heads         sigma (K -> 1) and the RGB output layer (128 -> 3) are dot products of the fp32 rows the
//                 epilogue threads already hold (thread = point); the [h, in] skip input of the heads is
//                 extra K blocks of the same job, the `in` share of sigma is computed in fp32 by the
//                 encoding warps.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "nd_common.cuh"

namespace nd {

constexpr int kTcTile = 128;            // points per tile = UMMA M
constexpr int kTcWidth = 256;           // net_width
constexpr int kTcCondWidth = 128;       // net_width_condition
constexpr int kTcPosOct = 10, kTcViewOct = 4;
constexpr int kTcPosDim = 3 + 6 * kTcPosOct;     // 63
constexpr int kTcViewDim = 3 + 6 * kTcViewOct;   // 27
constexpr int kTcInPad = 144;           // [posenc | features | 0] columns of the IN buffer (9 K steps)
constexpr int kTcCondCol = 144;         // view encoding: IN columns 144 .. 175 (2 K steps)
constexpr int kTcInChunks = 22;         // 16-byte chunks (8 bf16) a point's IN row is written in
constexpr int kTcBlockBytes = 16384;    // 128 rows x 128 B: one A block
constexpr int kTcUnitN = 256;           // output columns per B unit = UMMA N (128: a 256-wide layer is two units per K block)
constexpr int kTcStageBytes = kTcUnitN * 128;    // rows x 128 B: one B unit
constexpr int kTcStages = 160 * 1024 / kTcStageBytes;          // bf16: 5 weight stages beside 3 IN blocks
constexpr int kTcStagesSplit = 3;                              // split: 3 weight stages beside 6 IN blocks (hi and lo)
constexpr int kTcMaxUnits = 64, kTcMaxJobs = 12, kTcMaxDepth = 8;
constexpr int kTcEpiThreads = 256;      // warps 0-7: TMEM lane quarter = warp % 4, column half of a 64-column chunk = warp / 4
constexpr int kTcThreads = 448;         // + warps 8-11 input encoding, warp 12 weight producer, warp 13 MMA issuer
constexpr uint32_t kOffIn = 0, kOffB = kOffIn + 3 * kTcBlockBytes, kOffTail = kOffB + kTcStages * kTcStageBytes;   // 212992
constexpr uint32_t kOffBSplit = kOffIn + 6 * kTcBlockBytes;    // split: IN hi blocks 0-2, IN lo blocks 3-5, then the weight ring
static_assert(kOffBSplit + kTcStagesSplit * kTcStageBytes <= kOffTail, "split layout must fit below the tail");
constexpr uint32_t kTailBars = 0, kTailTmem = 256, kTailHead = 512, kTailSigIn = kTailHead + 2 * kTcTile * 16,
                   kTailUnits = kTailSigIn + 4 * kTcTile * 4, kTailPar = kTailUnits + kTcMaxUnits * 16;

enum : uint8_t { kJobRelu = 1, kJobWritesH = 2, kJobSigma = 4, kJobRgb = 8, kJobFreesIn = 16 };
// unit flags: 1 = first unit of its K block (wait for the A operand), 2 = first K block (overwrite D),
// split precision: 4 = B holds the hi halves (issue A hi and A lo against it), 8 = B holds the lo halves (A hi only)
enum : uint8_t { kUnitFirst = 1, kUnitOverwrite = 2, kUnitBHi = 4, kUnitBLo = 8 };

struct TcUnit {                 // kTcUnitN (or 128) output columns x one K block of one job = one B stage
    uint32_t src_off;           // byte offset of the unit in the weight image
    uint8_t a_blk;              // A operand: 0..3 = chunk of the previous job's output (TMEM), 4..6 = IN block (smem)
    uint8_t k0, nk;             // first K step (16 columns each) inside the block and number of steps
    uint8_t flags;              // 1: first unit of its K block (wait for the A operand), 2: first K block (overwrite D)
    uint16_t d_col;             // accumulator column of the unit's first output
    uint16_t n;                 // output columns (UMMA N)
};
struct TcJob {
    uint8_t first_unit, n_units, flags, pad;
    uint16_t n, bias_off;       // output width (UMMA N), float offset of the bias in the small-parameter block
};
struct TcPlan {
    int n_jobs, n_units;        // jobs per tile in the mode at hand
    TcJob jobs[kTcMaxJobs];
    TcUnit units[kTcMaxUnits];
};
struct TcPackUnit {             // source of one block for the packing kernel
    const float *w;             // reference weight [n][k_ref]
    int n, row0, k_ref, col_base, lo, hi;   // rows row0 .. row0 + n; block column kk in [lo, hi) <- reference column col_base + kk
    uint32_t dst_off;
    int low_half;               // split precision: 0 = bf16(w), 1 = bf16(w - bf16(w))
};
struct TcLayout {
    TcPlan full, density;       // with / without the colour branch
    TcPackUnit pack[kTcMaxUnits];
    int n_pack;
    size_t image_bytes;         // weight image
    int n_par;                  // small parameters (floats): biases, sigma row, rgb output layer
    int off_ws_h, off_ws_in, off_bsig, off_wo, off_bo;
    size_t total_bytes;
    int in_dim, feat;
};

struct TcArgs {
    TcPlan plan;
    const uint8_t *wimg;
    const float *par;
    int n_par, off_ws_h, off_ws_in, off_bsig, off_wo, off_bo;
    const float *x, *feat, *cond;
    int64_t n_points;
    int samples_per_ray, feat_dim, n_tiles;
    float *sigma, *alpha, *rgb;
};

// ---- plan / layout (host) -------------------------------------------------------------------------------
static bool tc_layout(const nd_mlp_weights *w, TcLayout &L, bool report, bool split = false) {
    const int depth = w->net_depth, feat = w->feature_dim, skip = w->skip_layer;
    const int in_dim = kTcPosDim + feat;
    if (depth < 1 || depth > kTcMaxDepth || w->net_width != kTcWidth || w->cond_width != kTcCondWidth || feat < 0 ||
        in_dim > kTcInPad || w->pos_octaves != kTcPosOct || w->view_octaves != kTcViewOct || skip < 0) {
        if (report)
            set_error("nerf_mlp (tensor-core path): unsupported architecture (depth %d, width %d, cond width %d, feature_dim %d, "
                      "octaves %d/%d): needs net_width 256, net_width_condition 128, 63 + feature_dim <= 144, octaves 10/4",
                      depth, w->net_width, w->cond_width, feat, w->pos_octaves, w->view_octaves);
        return false;
    }
    memset(&L, 0, sizeof(L));
    L.in_dim = in_dim;
    L.feat = feat;
    TcPlan &P = L.full;
    uint32_t img = 0;
    int par = 0;
    auto add_job = [&](int n, uint8_t flags) -> TcJob & {
        TcJob &j = P.jobs[P.n_jobs++];
        j.first_unit = (uint8_t)P.n_units;
        j.n_units = 0;
        j.flags = flags;
        j.n = (uint16_t)n;
        j.bias_off = (uint16_t)par;
        par += n;
        return j;
    };
    auto add_unit = [&](TcJob &j, const float *wsrc, int k_ref, int a_blk, int k0, int nk, int col_base, int lo, int hi) {
        const bool first_kb = j.n_units == 0;
        const int un = j.n < kTcUnitN ? j.n : kTcUnitN;
        for (int nh = 0; nh < j.n / un; ++nh) {
            for (int half = 0; half < (split ? 2 : 1); ++half) {                 // split: the hi weights, then the lo weights
                if (P.n_units >= kTcMaxUnits) { ++P.n_units; continue; }         // counted, rejected below
                TcUnit &u = P.units[P.n_units];
                TcPackUnit &pu = L.pack[P.n_units];
                ++P.n_units;
                ++j.n_units;
                u.src_off = img;
                u.a_blk = (uint8_t)a_blk;
                u.k0 = (uint8_t)k0;
                u.nk = (uint8_t)nk;
                u.flags = (uint8_t)((nh == 0 && half == 0 ? kUnitFirst : 0) | (first_kb && half == 0 ? kUnitOverwrite : 0) |
                                    (split ? (half == 0 ? kUnitBHi : kUnitBLo) : 0));
                u.d_col = (uint16_t)(nh * un);
                u.n = (uint16_t)un;
                pu.w = wsrc;
                pu.n = un;
                pu.row0 = nh * un;
                pu.k_ref = k_ref;
                pu.col_base = col_base;
                pu.lo = lo;
                pu.hi = hi;
                pu.dst_off = img;
                pu.low_half = half;
                img += (uint32_t)un * 128u;
            }
        }
    };
    auto add_h_units = [&](TcJob &j, const float *wsrc, int k_ref) {
        for (int c = 0; c < 4; ++c) add_unit(j, wsrc, k_ref, c, 0, 4, c * 64, 0, 64);
    };
    auto add_in_units = [&](TcJob &j, const float *wsrc, int k_ref, int ref0) {   // ref0: reference column of in[0]
        for (int b = 0; b < 3; ++b) {
            const int cols = b < 2 ? 64 : kTcInPad - 128;
            int hi = in_dim - b * 64;
            hi = hi < 0 ? 0 : (hi > cols ? cols : hi);
            add_unit(j, wsrc, k_ref, 4 + b, 0, cols / 16, ref0 + b * 64, 0, hi);
        }
    };
    bool cat = false;
    for (int i = 0; i < depth; ++i) {
        TcJob &j = add_job(kTcWidth, kJobRelu | kJobWritesH);
        const int k_ref = i == 0 ? in_dim : kTcWidth + (cat ? in_dim : 0);
        if (i == 0) {
            add_in_units(j, w->base_w[0], k_ref, 0);
        } else {
            add_h_units(j, w->base_w[i], k_ref);
            if (cat) add_in_units(j, w->base_w[i], k_ref, kTcWidth);
        }
        cat = skip > 0 && i % skip == 0 && i > 0;
    }
    P.jobs[P.n_jobs - 1].flags |= kJobSigma;
    const int head_ref = kTcWidth + (cat ? in_dim : 0);
    // the density-only plan stops here: its last job does not publish H, and IN is free after the last job reading it
    L.density = P;
    {
        TcPlan &D = L.density;
        D.jobs[D.n_jobs - 1].flags &= (uint8_t)~kJobWritesH;
        int last_in = 0;
        for (int j = 0; j < D.n_jobs; ++j)
            for (int u = 0; u < D.jobs[j].n_units; ++u)
                if (D.units[D.jobs[j].first_unit + u].a_blk >= 4) last_in = j;
        D.jobs[last_in].flags |= kJobFreesIn;
    }
    {   // bottleneck (no activation) and the colour hidden layer
        TcJob &jb = add_job(kTcWidth, kJobWritesH);
        add_h_units(jb, w->bottleneck_w, head_ref);
        if (cat) add_in_units(jb, w->bottleneck_w, head_ref, kTcWidth);
        TcJob &jr = add_job(kTcCondWidth, kJobRelu | kJobRgb | kJobFreesIn);
        add_h_units(jr, w->rgb_hidden_w, kTcWidth + kTcViewDim);
        // view encoding: IN block 2, block columns 16 .. 47 (IN columns 144 .. 175)
        add_unit(jr, w->rgb_hidden_w, kTcWidth + kTcViewDim, 6, 1, 2, kTcWidth - 16, 16, 16 + kTcViewDim);
    }
    L.n_pack = P.n_units;
    L.image_bytes = img;
    L.off_ws_h = par;  par += kTcWidth;
    L.off_ws_in = par; par += kTcInPad;
    L.off_bsig = par;  par += 4;
    L.off_wo = par;    par += 3 * kTcCondWidth;
    L.off_bo = par;    par += 4;
    L.n_par = par;
    L.total_bytes = align_up(L.image_bytes, 256) + (size_t)par * sizeof(float);
    if (P.n_units > kTcMaxUnits || 1024 + kOffTail + kTailPar + (size_t)par * sizeof(float) > 227 * 1024) {
        if (report)
            set_error("nerf_mlp (tensor-core path): net_depth %d with this skip pattern needs %d weight blocks / %zu bytes of shared "
                      "memory per CTA (limits: %d, %d); use the fp32 path", depth, P.n_units,
                      1024 + kOffTail + kTailPar + (size_t)par * sizeof(float), kTcMaxUnits, 227 * 1024);
        return false;
    }
    (void)head_ref;
    return true;
}

// ---- packing --------------------------------------------------------------------------------------------
struct TcPackArgs {                // kernel parameters are limited to 4 KB: 64 units per launch
    TcPackUnit u[64];
};

// K-major SWIZZLE_128B image of an [n][64] bf16 block: row r at (r / 8) * 1024 + (r % 8) * 128, its 16-byte chunk c
// at position c ^ (r % 8)
__device__ __host__ __forceinline__ uint32_t sw128_offset(int r, int col) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((col >> 3) ^ (r & 7)) & 7) << 4) + (col & 7) * 2);
}

__global__ void k_pack_tc_units(const __grid_constant__ TcPackArgs a, uint8_t *__restrict__ img) {
    const TcPackUnit &u = a.u[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u.n * 64) return;
    const int r = i >> 6, kk = i & 63;
    float v = 0.0f;
    if (kk >= u.lo && kk < u.hi) v = u.w[(size_t)(u.row0 + r) * u.k_ref + u.col_base + kk];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    if (u.low_half) h = __float2bfloat16_rn(v - __bfloat162float(h));      // what the hi half leaves of the weight
    *reinterpret_cast<__nv_bfloat16 *>(img + u.dst_off + sw128_offset(r, kk)) = h;
}

struct TcParSrc {
    const float *bias[kTcMaxJobs];
    int bias_off[kTcMaxJobs], bias_n[kTcMaxJobs], n_jobs;
    const float *sigma_w, *sigma_b, *wo, *bo;
    int head_ref, in_dim, cat;
    int off_ws_h, off_ws_in, off_bsig, off_wo, off_bo, n_par;
};

__global__ void k_pack_tc_params(const __grid_constant__ TcParSrc s, float *__restrict__ par) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n_par) return;
    float v = 0.0f;
    for (int j = 0; j < s.n_jobs; ++j)
        if (i >= s.bias_off[j] && i < s.bias_off[j] + s.bias_n[j]) v = s.bias[j][i - s.bias_off[j]];
    if (i >= s.off_ws_h && i < s.off_ws_h + kTcWidth) v = s.sigma_w[i - s.off_ws_h];
    if (i >= s.off_ws_in && i < s.off_ws_in + kTcInPad) {
        const int k = i - s.off_ws_in;
        v = (s.cat && k < s.in_dim) ? s.sigma_w[kTcWidth + k] : 0.0f;
    }
    if (i == s.off_bsig) v = s.sigma_b[0];
    if (i >= s.off_wo && i < s.off_wo + 3 * kTcCondWidth) v = s.wo[i - s.off_wo];
    if (i >= s.off_bo && i < s.off_bo + 3) v = s.bo[i - s.off_bo];
    par[i] = v;
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------
namespace tc {
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// one lane of a converged warp (lets ptxas keep the MMA operands in uniform registers and drop its per-lane issue loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
    return r;
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = (uint64_t)((addr & 0x3ffffu) >> 4);   // start address, bits [0, 14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16, 30)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset, bits [32, 46)
    d |= (uint64_t)1 << 46;                            // descriptor version 1 (sm_100), bits [46, 48)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B, bits [61, 64)
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A rows = TMEM lanes, K pairs packed into 32-bit columns
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// relu fused into the bf16 conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
}  // namespace tc

// ---- input encoding -----------------------------------------------------------------------------------
// [x, sin(2^k x) (k outer, xyz inner), sin(2^k x + pi/2)] (nerf_mlp.py:181-197) by angle doubling from one accurate
// sincos per coordinate: the doubling error (2^k * 6e-8) and the reference's fp32 rounding of 2^k x + pi/2
// (|delta arg| <= 1.3e-4) are both far below the bf16 rounding the values get on their way into the tensor core.
template <int kOct>
__device__ __forceinline__ void encode_doubling(const float (&x)[3], float *dst) {
    float s[3], c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dst[i] = x[i];
        sincosf(x[i], &s[i], &c[i]);
    }
#pragma unroll
    for (int k = 0; k < kOct; ++k) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            dst[3 + k * 3 + i] = s[i];
            dst[3 + kOct * 3 + k * 3 + i] = c[i];
            const float s2 = 2.0f * s[i] * c[i];
            const float c2 = fmaf(-2.0f * s[i], s[i], 1.0f);
            s[i] = s2;
            c[i] = c2;
        }
    }
}

// The reference's own arithmetic (nerf_mlp.py:189-196): sin(fl(x * 2^k)) and sin(fl(fl(x * 2^k) + fl(pi / 2))) with an
// accurate sinf -- the fp32 rounding of arguments up to ~4000 rad moves the value by up to 1e-4, so the fp32-grade
// precision must round where the reference rounds.
template <int kOct>
__device__ __forceinline__ void encode_exact(const float (&x)[3], float *dst) {
    const float half_pi = 1.5707963267948966f;
#pragma unroll
    for (int i = 0; i < 3; ++i) dst[i] = x[i];
#pragma unroll
    for (int k = 0; k < kOct; ++k) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float sc = __fmul_rn(x[i], (float)(1 << k));
            dst[3 + k * 3 + i] = sinf(sc);
            dst[3 + kOct * 3 + k * 3 + i] = sinf(__fadd_rn(sc, half_pi));
        }
    }
}

// hi / lo halves of an fp32 pair as packed bf16x2 words: hi = bf16(v), lo = bf16(v - hi)
__device__ __forceinline__ void split_bf16(float a0, float a1, uint32_t &hi, uint32_t &lo) {
    hi = tc::pack_bf16(a0, a1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = tc::pack_bf16(a0 - h0, a1 - h1);
}

// kSplit: the fp32-grade "3 x bf16" precision (see the header): A hi / lo halves side by side in tensor memory and in the
// IN buffer, hi and lo weight units alternating in the weight stream.
template <bool kSplit>
__global__ void __launch_bounds__(kTcThreads, 1)
k_nerf_mlp_tc(const __grid_constant__ TcArgs a) {
    constexpr int kStages = kSplit ? kTcStagesSplit : kTcStages;
    constexpr uint32_t kOffBk = kSplit ? kOffBSplit : kOffB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B operands need 1024-byte alignment
    uint8_t *sm = smem_raw + (base - raw);
    const uint32_t sIN = base + kOffIn, sB = base + kOffBk;
    const uint32_t bars = base + kOffTail + kTailBars;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(sm + kOffTail + kTailTmem);
    float *s_head = reinterpret_cast<float *>(sm + kOffTail + kTailHead);          // [2][128][4] (slot t % 2): sigma | rgb partials of column half 1
    // [4][128]: slot t % 4.  The encoding warps may run ahead of the epilogue by the tile whose first job is in
    // flight plus the one they hold in registers; four slots keep the writer of tile t + 4 behind the reader of tile t
    // in every job plan (two would race in the density-only plan, where IN is released after the first job).
    float *s_sig_in = reinterpret_cast<float *>(sm + kOffTail + kTailSigIn);
    float *s_par = reinterpret_cast<float *>(sm + kOffTail + kTailPar);
    const uint32_t b_full = bars, b_empty = bars + 8 * kStages, d_full = bars + 8 * 2 * kStages,
                   d_empty = d_full + 8 * 2, h_ready = d_full + 8 * 4, in_ready = d_full + 8 * 8, in_free = d_full + 8 * 9;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { tc::mbar_init(b_full + 8 * i, 1); tc::mbar_init(b_empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(d_full + 8 * i, 1); tc::mbar_init(d_empty + 8 * i, kTcEpiThreads / 32); }
        for (int i = 0; i < 4; ++i) tc::mbar_init(h_ready + 8 * i, kTcEpiThreads / 32);   // one arrival per warp
        tc::mbar_init(in_ready, 4);
        tc::mbar_init(in_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 13) tc::tmem_alloc(tc::smem_u32(s_tmem), 512);
    for (int i = threadIdx.x; i < a.n_par; i += kTcThreads) s_par[i] = a.par[i];
    // The MMA warp's per-unit record, precomputed once: one LDS.128 per unit instead of dependent constant-bank loads
    // and descriptor arithmetic between the MMAs (the issuing thread, not the tensor pipe, paced the first version:
    // ~224 cycles per M128 x N256 x K16 instruction against 138 for the same instruction issued back to back,
    // tools/microbench7.cu).
    //   x: A operand of the first K step -- TMEM column offset inside the previous job's region, or the low word of the
    //      shared-memory descriptor of the IN block;  y: instruction descriptor;  z: 2 * k0 | nk << 8 | flags << 16 |
    //      a_blk << 24;  w: accumulator column
    uint4 *s_units = reinterpret_cast<uint4 *>(sm + kOffTail + kTailUnits);
    for (int i = threadIdx.x; i < a.plan.n_units; i += kTcThreads) {
        const TcUnit u = a.plan.units[i];
        uint4 r;
        r.x = u.a_blk < 4 ? (uint32_t)u.a_blk * 64u + 16u * u.k0
                          : (uint32_t)tc::smem_desc(base + kOffIn + (uint32_t)(u.a_blk - 4) * kTcBlockBytes) + 2u * u.k0;
        r.y = tc::instr_desc(u.n);
        r.z = 2u * u.k0 | ((uint32_t)u.nk << 8) | ((uint32_t)u.flags << 16) | ((uint32_t)u.a_blk << 24);
        r.w = u.d_col;
        s_units[i] = r;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const TcPlan &P = a.plan;

    if (warp == 12) {
        // ---------------- weight producer ----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                for (int u = 0; u < P.n_units; ++u) {
                    tc::mbar_wait(b_empty + 8 * stage, ph ^ 1u);
                    const uint32_t bytes = (uint32_t)P.units[u].n * 128u;
                    tc::mbar_expect_tx(b_full + 8 * stage, bytes);
                    tc::bulk_g2s(sB + stage * kTcStageBytes, a.wimg + P.units[u].src_off, bytes, b_full + 8 * stage);
                    if (++stage == kStages) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 13) {
        // ---------------- MMA issuer ----------------
        // The whole warp walks the job / unit lists in lock-step (uniform control flow, waits included); the tcgen05
        // instructions are issued by one elected lane.  Everything per unit comes from one prefetched LDS.128: a single
        // thread retires a dependent instruction every ~4-5 cycles, so ~40 instructions between two MMAs already cost
        // more than the 138 cycles the tensor pipe needs for an M128 x N256 x K16 step (tools/microbench7.cu).
        {
            int stage = 0;
            uint32_t ph = 0, hpar = 0, in_par = 0, dpar = 3u;       // dpar bit d: parity to wait for on d_empty[d]
            uint32_t jc = 0;                                           // running job counter -> accumulator buffer
            const uint64_t desc_hi = tc::smem_desc(0) & 0xffffffff00000000ull;        // layout / stride bits of every operand
            const uint32_t bd0 = (uint32_t)tc::smem_desc(sB);                          // low word of stage 0's B descriptor
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                bool in_ok = false;
                for (int j = 0; j < P.n_jobs; ++j, ++jc) {
                    const uint32_t first_unit = P.jobs[j].first_unit, n_units = P.jobs[j].n_units, jflags = P.jobs[j].flags;
                    const uint32_t d = jc & 1u;
                    uint4 ur = s_units[first_unit];
                    tc::mbar_wait(d_empty + 8 * d, (dpar >> d) & 1u);
                    dpar ^= 1u << d;
                    const uint32_t a_region = tmem + (d ^ 1u) * 256u;             // the previous job's region holds this job's A
                    const uint32_t d_region = tmem + d * 256u;
                    for (uint32_t uu = 0; uu < n_units; ++uu) {
                        const uint4 u = ur;
                        if (uu + 1 < n_units) ur = s_units[first_unit + uu + 1];  // next record in flight during the waits
                        const uint32_t a_blk = u.z >> 24, uflags = (u.z >> 16) & 0xffu, nk = (u.z >> 8) & 0xffu;
                        const bool from_tmem = a_blk < 4;
                        if (from_tmem) {
                            if (uflags & kUnitFirst) {
                                tc::mbar_wait(h_ready + 8 * a_blk, (hpar >> a_blk) & 1u);
                                hpar ^= 1u << a_blk;
                            }
                        } else if (!in_ok) {
                            tc::mbar_wait(in_ready, in_par);
                            in_par ^= 1u;
                            in_ok = true;
                        }
                        tc::mbar_wait(b_full + 8 * stage, ph);
                        tc::tc_fence_after();
                        const uint64_t bd = desc_hi | (uint64_t)(bd0 + (uint32_t)stage * (kTcStageBytes >> 4) + (u.z & 0xffu));
                        const uint32_t d_tmem = d_region + u.w;
                        const uint32_t acc0 = (uflags & kUnitOverwrite) ? 0u : 1u;     // first K block of the job overwrites D
                        if (tc::elect_one()) {
                            if (from_tmem) {
                                // K step s of chunk c: 8 packed columns at column 64 c + 16 s of the previous region
                                // (split precision: the hi halves; the lo halves sit in the 8 columns behind them)
                                const uint32_t a_tmem = a_region + u.x;
                                tc::umma_bf16_ts(d_tmem, a_tmem, bd, u.y, acc0);
                                if (nk > 1) tc::umma_bf16_ts(d_tmem, a_tmem + 16u, bd + 2u, u.y, 1u);
                                if (nk > 2) tc::umma_bf16_ts(d_tmem, a_tmem + 32u, bd + 4u, u.y, 1u);
                                if (nk > 3) tc::umma_bf16_ts(d_tmem, a_tmem + 48u, bd + 6u, u.y, 1u);
                                if (kSplit && (uflags & kUnitBHi)) {                // A lo x B hi
                                    tc::umma_bf16_ts(d_tmem, a_tmem + 8u, bd, u.y, 1u);
                                    if (nk > 1) tc::umma_bf16_ts(d_tmem, a_tmem + 24u, bd + 2u, u.y, 1u);
                                    if (nk > 2) tc::umma_bf16_ts(d_tmem, a_tmem + 40u, bd + 4u, u.y, 1u);
                                    if (nk > 3) tc::umma_bf16_ts(d_tmem, a_tmem + 56u, bd + 6u, u.y, 1u);
                                }
                            } else {
                                const uint64_t ad = desc_hi | (uint64_t)u.x;           // 32 B per K step = 2 descriptor units
                                tc::umma_bf16(d_tmem, ad, bd, u.y, acc0);
                                if (nk > 1) tc::umma_bf16(d_tmem, ad + 2u, bd + 2u, u.y, 1u);
                                if (nk > 2) tc::umma_bf16(d_tmem, ad + 4u, bd + 4u, u.y, 1u);
                                if (nk > 3) tc::umma_bf16(d_tmem, ad + 6u, bd + 6u, u.y, 1u);
                                if (kSplit && (uflags & kUnitBHi)) {                // IN lo (three blocks further on) x B hi
                                    const uint64_t al = ad + (uint64_t)(3u * (kTcBlockBytes >> 4));
                                    tc::umma_bf16(d_tmem, al, bd, u.y, 1u);
                                    if (nk > 1) tc::umma_bf16(d_tmem, al + 2u, bd + 2u, u.y, 1u);
                                    if (nk > 2) tc::umma_bf16(d_tmem, al + 4u, bd + 4u, u.y, 1u);
                                    if (nk > 3) tc::umma_bf16(d_tmem, al + 6u, bd + 6u, u.y, 1u);
                                }
                            }
                            tc::umma_commit(b_empty + 8 * stage);                   // frees the B stage once these MMAs are done
                            if (uu + 1 == n_units) {
                                tc::umma_commit(d_full + 8 * d);
                                if (jflags & kJobFreesIn) tc::umma_commit(in_free);
                            }
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ---------------- input encoding of the next tile (thread = point) ----------------
        const int r = threadIdx.x - kTcEpiThreads;
        const uint32_t row_addr = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const float *ws_in = s_par + a.off_ws_in;
        int t = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++t) {
            const int64_t gp = (int64_t)tile * kTcTile + r;
            const bool ok = gp < a.n_points;
            if constexpr (kSplit) {
                // fp32-grade inputs: the expensive part (60 + 24 accurate sines) is done ahead of the IN buffer's release and
                // held as fp32; the features are read and everything is split into hi / lo bf16 halves afterwards
                float enc[kTcPosDim], venc[32];
                float xyz[3] = {0.f, 0.f, 0.f};
                if (ok) { xyz[0] = a.x[gp * 3]; xyz[1] = a.x[gp * 3 + 1]; xyz[2] = a.x[gp * 3 + 2]; }
                encode_exact<kTcPosOct>(xyz, enc);
#pragma unroll
                for (int i = 0; i < 32; ++i) venc[i] = 0.0f;
                if (a.cond != nullptr && ok) {
                    const int64_t ray = gp / a.samples_per_ray;
                    const float dir[3] = {a.cond[ray * 3], a.cond[ray * 3 + 1], a.cond[ray * 3 + 2]};
                    encode_exact<kTcViewOct>(dir, venc);
                }
                if (t > 0) tc::mbar_wait(in_free, (uint32_t)(t - 1) & 1u);     // the previous tile's last reader of IN is done
                const float *fr = a.feat + gp * a.feat_dim;
                float sig_in = 0.0f;
#pragma unroll
                for (int c = 0; c < kTcInChunks; ++c) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int idx = c * 8 + e;
                        if (c >= kTcInPad / 8) v[e] = venc[idx - kTcInPad];
                        else if (idx < kTcPosDim) v[e] = ok ? enc[idx] : 0.0f;
                        else v[e] = (ok && idx - kTcPosDim < a.feat_dim) ? __ldg(fr + (idx - kTcPosDim)) : 0.0f;
                        if (c < kTcInPad / 8) sig_in = fmaf(v[e], ws_in[idx], sig_in);
                    }
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_bf16(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
                    const int blk = c >> 3, cc = c & 7;
                    const uint32_t dst = sIN + blk * kTcBlockBytes + row_addr + (uint32_t)((cc ^ (r & 7)) << 4);
                    tc::sts_u4(dst, hi[0], hi[1], hi[2], hi[3]);
                    tc::sts_u4(dst + 3 * kTcBlockBytes, lo[0], lo[1], lo[2], lo[3]);
                }
                s_sig_in[(t & 3) * kTcTile + r] = sig_in;
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(in_ready);
                continue;
            }
            uint32_t pk[kTcInChunks * 4];
            float sig_in = 0.0f;
            {
                float enc[kTcPosDim];
                float xyz[3] = {0.f, 0.f, 0.f};
                if (ok) { xyz[0] = a.x[gp * 3]; xyz[1] = a.x[gp * 3 + 1]; xyz[2] = a.x[gp * 3 + 2]; }
                encode_doubling<kTcPosOct>(xyz, enc);
                const float *fr = a.feat + gp * a.feat_dim;
#pragma unroll
                for (int c = 0; c < kTcInPad / 8; ++c) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int idx = c * 8 + e;
                        if (idx < kTcPosDim) v[e] = ok ? enc[idx] : 0.0f;
                        else v[e] = (ok && idx - kTcPosDim < a.feat_dim) ? __ldg(fr + (idx - kTcPosDim)) : 0.0f;
                        sig_in = fmaf(v[e], ws_in[idx], sig_in);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[c * 4 + e] = tc::pack_bf16(v[2 * e], v[2 * e + 1]);
                }
            }
            {
                float enc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) enc[i] = 0.0f;
                if (a.cond != nullptr && ok) {
                    const int64_t ray = gp / a.samples_per_ray;
                    const float dir[3] = {a.cond[ray * 3], a.cond[ray * 3 + 1], a.cond[ray * 3 + 2]};
                    encode_doubling<kTcViewOct>(dir, enc);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[(kTcInPad / 8 + c) * 4 + e] = tc::pack_bf16(enc[c * 8 + 2 * e], enc[c * 8 + 2 * e + 1]);
            }
            if (t > 0) tc::mbar_wait(in_free, (uint32_t)(t - 1) & 1u);     // the previous tile's last reader of IN is done
#pragma unroll
            for (int c = 0; c < kTcInChunks; ++c) {
                const int blk = c >> 3, cc = c & 7;
                tc::sts_u4(sIN + blk * kTcBlockBytes + row_addr + (uint32_t)((cc ^ (r & 7)) << 4), pk[c * 4], pk[c * 4 + 1],
                           pk[c * 4 + 2], pk[c * 4 + 3]);
            }
            s_sig_in[(t & 3) * kTcTile + r] = sig_in;
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(in_ready);
        }
    } else {
        // ---------------- epilogue: thread = (point = TMEM lane, 32-column half of every 64-column chunk) ----------------
        const int q = warp & 3, half = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float *ws_h = s_par + a.off_ws_h, *wo = s_par + a.off_wo;
        uint32_t dfull_par = 0, jc = 0;
        int t = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++t) {
            const int64_t gp = (int64_t)tile * kTcTile + r;
            for (int j = 0; j < P.n_jobs; ++j, ++jc) {
                const TcJob job = P.jobs[j];
                const uint32_t d = jc & 1u;
                tc::mbar_wait(d_full + 8 * d, (dfull_par >> d) & 1u);
                dfull_par ^= 1u << d;
                tc::tc_fence_after();
                const float *bias = s_par + job.bias_off;
                const bool relu = job.flags & kJobRelu, writes_h = job.flags & kJobWritesH;
                const bool sig_head = job.flags & kJobSigma, rgb_head = job.flags & kJobRgb;
                float sig = 0.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
                const int n_chunks = job.n >> 6;
                for (int ch = 0; ch < n_chunks; ++ch) {
                    const int col = ch * 64 + half * 32;
                    uint32_t v[32];
                    tc::tmem_ld32(tmem + lane_base + d * 256u + (uint32_t)col, v);
                    tc::tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(bias + col + i);
                        f[i] = __uint_as_float(v[i]) + b4.x;
                        f[i + 1] = __uint_as_float(v[i + 1]) + b4.y;
                        f[i + 2] = __uint_as_float(v[i + 2]) + b4.z;
                        f[i + 3] = __uint_as_float(v[i + 3]) + b4.w;
                    }
                    if (sig_head) {                            // fp32 row x fp32 sigma weights, relu'd values
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 w4 = *reinterpret_cast<const float4 *>(ws_h + col + i);
                            sig = fmaf(fmaxf(f[i], 0.f), w4.x, sig); sig = fmaf(fmaxf(f[i + 1], 0.f), w4.y, sig);
                            sig = fmaf(fmaxf(f[i + 2], 0.f), w4.z, sig); sig = fmaf(fmaxf(f[i + 3], 0.f), w4.w, sig);
                        }
                    }
                    if (rgb_head) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 w0 = *reinterpret_cast<const float4 *>(wo + col + i);
                            const float4 w1 = *reinterpret_cast<const float4 *>(wo + kTcCondWidth + col + i);
                            const float4 w2 = *reinterpret_cast<const float4 *>(wo + 2 * kTcCondWidth + col + i);
                            const float t0 = fmaxf(f[i], 0.f), t1 = fmaxf(f[i + 1], 0.f), t2 = fmaxf(f[i + 2], 0.f), t3 = fmaxf(f[i + 3], 0.f);
                            c0 = fmaf(t0, w0.x, c0); c0 = fmaf(t1, w0.y, c0); c0 = fmaf(t2, w0.z, c0); c0 = fmaf(t3, w0.w, c0);
                            c1 = fmaf(t0, w1.x, c1); c1 = fmaf(t1, w1.y, c1); c1 = fmaf(t2, w1.z, c1); c1 = fmaf(t3, w1.w, c1);
                            c2 = fmaf(t0, w2.x, c2); c2 = fmaf(t1, w2.y, c2); c2 = fmaf(t2, w2.z, c2); c2 = fmaf(t3, w2.w, c2);
                        }
                    }
                    if (writes_h) {
                        // bf16 pairs back into the columns they came from: K step s at column 16 s of this region
                        uint32_t pkd[16];
                        const uint32_t t0 = tmem + lane_base + d * 256u + (uint32_t)col;
                        if constexpr (kSplit) {
                            // hi halves where the bf16 precision keeps its operand, lo halves in the 8 columns behind them
                            uint32_t pkl[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float x0 = relu ? fmaxf(f[2 * i], 0.f) : f[2 * i], x1 = relu ? fmaxf(f[2 * i + 1], 0.f) : f[2 * i + 1];
                                split_bf16(x0, x1, pkd[i], pkl[i]);
                            }
                            tc::tmem_st8(t0 + 8u, pkl);
                            tc::tmem_st8(t0 + 24u, pkl + 8);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                pkd[i] = relu ? tc::pack_bf16_relu(f[2 * i], f[2 * i + 1]) : tc::pack_bf16(f[2 * i], f[2 * i + 1]);
                        }
                        tc::tmem_st8(t0, pkd);
                        tc::tmem_st8(t0 + 16u, pkd + 8);
                        tc::tmem_st_wait();
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(h_ready + 8 * ch);
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(d_empty + 8 * d);   // this accumulator region may be overwritten
                if (sig_head || rgb_head) {
                    // the two column halves of a point sit in warps q and q + 4: half 1 hands its partial sums over
                    float *slot = s_head + ((t & 1) * kTcTile + r) * 4;
                    if (half == 1) {
                        if (sig_head) slot[0] = sig;
                        else { slot[1] = c0; slot[2] = c1; slot[3] = c2; }
                    }
                    tc::bar_sync(1 + q, 64);
                    if (half == 0 && gp < a.n_points) {
                        if (sig_head) {
                            const float sg = fmaxf(sig + slot[0] + s_sig_in[(t & 3) * kTcTile + r] + s_par[a.off_bsig], 0.0f);
                            if (a.sigma != nullptr) a.sigma[gp] = sg;
                            if (a.alpha != nullptr) a.alpha[gp] = 1.0f - expf(-sg);        // nerfdet.py:258
                        } else {
                            const float *bo = s_par + a.off_bo;
                            a.rgb[gp * 3] = 1.0f / (1.0f + expf(-(c0 + slot[1] + bo[0])));
                            a.rgb[gp * 3 + 1] = 1.0f / (1.0f + expf(-(c1 + slot[2] + bo[1])));
                            a.rgb[gp * 3 + 2] = 1.0f / (1.0f + expf(-(c2 + slot[3] + bo[2])));
                        }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, 512);
    }
}

}  // namespace nd

using namespace nd;

extern "C" {

static size_t tc_packed_bytes(const nd_mlp_weights *w, bool split) {
    TcLayout L;
    if (w == nullptr || !tc_layout(w, L, false, split)) return 0;
    return L.total_bytes;
}

static int tc_pack(const nd_mlp_weights *w, void *packed, size_t packed_bytes, void *stream, bool split) {
    ND_REQUIRE(w != nullptr && packed != nullptr, ND_ERR_BAD_ARG, "nd_pack_mlp_weights_tc: null pointer");
    TcLayout L;
    if (!tc_layout(w, L, true, split)) return ND_ERR_BAD_SHAPE;
    ND_REQUIRE(packed_bytes >= L.total_bytes, ND_ERR_WORKSPACE, "nd_pack_mlp_weights_tc: buffer too small (%zu < %zu)",
               packed_bytes, L.total_bytes);
    ND_REQUIRE((reinterpret_cast<uintptr_t>(packed) % 256) == 0, ND_ERR_BAD_ALIGNMENT,
               "nd_pack_mlp_weights_tc: buffer not 256-byte aligned");
    for (int i = 0; i < w->net_depth; ++i)
        ND_REQUIRE(w->base_w[i] != nullptr && w->base_b[i] != nullptr, ND_ERR_BAD_ARG, "nd_pack_mlp_weights_tc: layer %d missing", i);
    ND_REQUIRE(w->sigma_w && w->sigma_b && w->bottleneck_w && w->bottleneck_b && w->rgb_hidden_w && w->rgb_hidden_b &&
                   w->rgb_out_w && w->rgb_out_b,
               ND_ERR_BAD_ARG, "nd_pack_mlp_weights_tc: head weights missing");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *img = reinterpret_cast<uint8_t *>(packed);
    float *par = reinterpret_cast<float *>(img + align_up(L.image_bytes, 256));
    for (int u0 = 0; u0 < L.n_pack; u0 += 64) {
        TcPackArgs pa;
        memset(&pa, 0, sizeof(pa));
        const int nb = L.n_pack - u0 < 64 ? L.n_pack - u0 : 64;
        for (int i = 0; i < nb; ++i) pa.u[i] = L.pack[u0 + i];
        k_pack_tc_units<<<dim3((unsigned)ceil_div(kTcWidth * 64, 256), (unsigned)nb), 256, 0, st>>>(pa, img);
        ND_CUDA_LAUNCH_CHECK("k_pack_tc_units");
    }
    TcParSrc ps;
    memset(&ps, 0, sizeof(ps));
    const TcPlan &P = L.full;
    ps.n_jobs = P.n_jobs;
    for (int j = 0; j < P.n_jobs; ++j) {
        ps.bias_off[j] = P.jobs[j].bias_off;
        ps.bias_n[j] = P.jobs[j].n;
        ps.bias[j] = j < w->net_depth ? w->base_b[j] : (j == w->net_depth ? w->bottleneck_b : w->rgb_hidden_b);
    }
    ps.sigma_w = w->sigma_w; ps.sigma_b = w->sigma_b; ps.wo = w->rgb_out_w; ps.bo = w->rgb_out_b;
    ps.in_dim = L.in_dim;
    {
        const int skip = w->skip_layer, i = w->net_depth - 1;
        ps.cat = (skip > 0 && i % skip == 0 && i > 0) ? 1 : 0;
    }
    ps.off_ws_h = L.off_ws_h; ps.off_ws_in = L.off_ws_in; ps.off_bsig = L.off_bsig; ps.off_wo = L.off_wo; ps.off_bo = L.off_bo;
    ps.n_par = L.n_par;
    k_pack_tc_params<<<(unsigned)ceil_div(L.n_par, 256), 256, 0, st>>>(ps, par);
    ND_CUDA_LAUNCH_CHECK("k_pack_tc_params");
    return ND_OK;
}

static int tc_forward(const nd_mlp_weights *arch, const void *packed, const float *x, const float *features, const float *cond,
                      int64_t n_points, int samples_per_ray, float *sigma, float *alpha, float *rgb, void *stream, bool split) {
    ND_REQUIRE(n_points >= 0, ND_ERR_BAD_SHAPE, "nd_nerf_mlp_fwd_tc: negative point count");
    if (n_points == 0) return ND_OK;
    ND_REQUIRE(arch != nullptr && packed != nullptr && x != nullptr, ND_ERR_BAD_ARG, "nd_nerf_mlp_fwd_tc: null pointer");
    TcLayout L;
    if (!tc_layout(arch, L, true, split)) return ND_ERR_BAD_SHAPE;
    ND_REQUIRE(L.feat == 0 || features != nullptr, ND_ERR_BAD_ARG, "nd_nerf_mlp_fwd_tc: features missing");
    ND_REQUIRE(rgb == nullptr || (cond != nullptr && samples_per_ray > 0 && n_points % samples_per_ray == 0),
               ND_ERR_BAD_SHAPE, "nd_nerf_mlp_fwd_tc: rgb needs cond [P / samples_per_ray][3]");
    ND_REQUIRE((reinterpret_cast<uintptr_t>(packed) % 256) == 0, ND_ERR_BAD_ALIGNMENT,
               "nd_nerf_mlp_fwd_tc: packed buffer not 256-byte aligned");
    const bool want_rgb = rgb != nullptr && cond != nullptr;
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.plan = want_rgb ? L.full : L.density;
    a.wimg = reinterpret_cast<const uint8_t *>(packed);
    a.par = reinterpret_cast<const float *>(a.wimg + align_up(L.image_bytes, 256));
    a.n_par = L.n_par;
    a.off_ws_h = L.off_ws_h; a.off_ws_in = L.off_ws_in; a.off_bsig = L.off_bsig; a.off_wo = L.off_wo; a.off_bo = L.off_bo;
    a.x = x; a.feat = features; a.cond = want_rgb ? cond : nullptr;
    a.n_points = n_points;
    a.samples_per_ray = samples_per_ray > 0 ? samples_per_ray : 1;
    a.feat_dim = L.feat;
    a.n_tiles = (int)ceil_div(n_points, kTcTile);
    a.sigma = sigma; a.alpha = alpha; a.rgb = want_rgb ? rgb : nullptr;
    const size_t smem = 1024 + kOffTail + kTailPar + (size_t)L.n_par * sizeof(float);
    ND_REQUIRE(smem <= 227 * 1024, ND_ERR_BAD_SHAPE, "nd_nerf_mlp_fwd_tc: %zu bytes of shared memory needed", smem);
    void (*kern)(const TcArgs) = split ? k_nerf_mlp_tc<true> : k_nerf_mlp_tc<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("k_nerf_mlp_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.n_tiles < sms ? a.n_tiles : sms;
    kern<<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(a);
    ND_CUDA_LAUNCH_CHECK("k_nerf_mlp_tc");
    return ND_OK;
}

size_t nd_mlp_tc_packed_bytes(const nd_mlp_weights *w) { return tc_packed_bytes(w, false); }
int nd_pack_mlp_weights_tc(const nd_mlp_weights *w, void *packed, size_t packed_bytes, void *stream) {
    return tc_pack(w, packed, packed_bytes, stream, false);
}
int nd_nerf_mlp_fwd_tc(const nd_mlp_weights *arch, const void *packed, const float *x, const float *features,
                       const float *cond, int64_t n_points, int samples_per_ray, float *sigma, float *alpha, float *rgb,
                       void *stream) {
    return tc_forward(arch, packed, x, features, cond, n_points, samples_per_ray, sigma, alpha, rgb, stream, false);
}

size_t nd_mlp_tc3_packed_bytes(const nd_mlp_weights *w) { return tc_packed_bytes(w, true); }
int nd_pack_mlp_weights_tc3(const nd_mlp_weights *w, void *packed, size_t packed_bytes, void *stream) {
    return tc_pack(w, packed, packed_bytes, stream, true);
}
int nd_nerf_mlp_fwd_tc3(const nd_mlp_weights *arch, const void *packed, const float *x, const float *features,
                        const float *cond, int64_t n_points, int samples_per_ray, float *sigma, float *alpha, float *rgb,
                        void *stream) {
    return tc_forward(arch, packed, x, features, cond, n_points, samples_per_ray, sigma, alpha, rgb, stream, true);
}

}  // extern "C"
