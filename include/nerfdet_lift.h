/*
 * nerfdet_lift.h -- C ABI of libnerfdet_lift.so, the B200 (sm_100a) implementation of
 * NeRF-Det's multi-view 2D->3D lifting path.
 *
 * The reference has no FFI layer on this path: its "operator API" is a set of Python
 * callables.  Every entry point below cites the reference code
 * it replaces as  <file>:<lines>  relative to the reference tree
 * (mmdet3d/models/detectors/nerfdet.py            -> "nerfdet.py",
 *  mmdet3d/models/model_utils/projection.py       -> "projection.py",
 *  mmdet3d/models/model_utils/render_ray.py       -> "render_ray.py",
 *  mmdet3d/models/model_utils/nerf_mlp.py         -> "nerf_mlp.py").
 * The Python package nerfdet_b200 binds these with ctypes and registers them as torch
 * custom ops; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; all buffers are
 *    owned by the caller and outlive the call; the library never allocates, frees or
 *    retains device memory (nd_*_workspace_bytes tells the caller what to provide);
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant,
 *    and never synchronise the host;
 *  - strides are in ELEMENTS; shapes are row-major unless strides are given;
 *  - return value: nd_status (0 = ok); nd_last_error_string() describes the last
 *    failure on the calling thread.  No C++ exception crosses this boundary.
 */
#ifndef NERFDET_LIFT_H_
#define NERFDET_LIFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ND_VERSION 100 /* 0.1.0 */

typedef enum nd_status {
    ND_OK = 0,
    ND_ERR_BAD_ARG = 1,        /* null pointer / negative size / unsupported combination */
    ND_ERR_BAD_SHAPE = 2,
    ND_ERR_BAD_ALIGNMENT = 3,
    ND_ERR_WORKSPACE = 4,      /* workspace missing or too small */
    ND_ERR_UNSUPPORTED_ARCH = 5,
    ND_ERR_CUDA = 6            /* launch / runtime error, see nd_last_error_string */
} nd_status;

typedef enum nd_dtype { ND_F32 = 0, ND_BF16 = 1 } nd_dtype;

/* A (possibly non-contiguous) stack of per-view 2-D maps, e.g. the reference's
 * `feature[:, :, :height, :width]` view of the FPN output (nerfdet.py:165). */
typedef struct nd_maps {
    const void *data;
    int32_t dtype;      /* nd_dtype */
    int32_t n_views;
    int32_t channels;
    int32_t height;     /* of the slice that is addressed */
    int32_t width;
    int64_t stride_v, stride_c, stride_y, stride_x;   /* elements */
} nd_maps;

/* Tuning knobs; pass NULL for defaults. */
typedef enum nd_lift_path {
    ND_LIFT_PATH_AUTO = 0,     /* plane-resident kernel when the maps are contiguous NCHW planes, else staged */
    ND_LIFT_PATH_STAGED = 1    /* force the generic pixel-major staging path (any strides / channels-last) */
} nd_lift_path;

typedef struct nd_lift_options {
    size_t scratch_budget_bytes;   /* staged path: pixel-major staging kept L2-resident; 0 = default (64 MiB) */
    int32_t voxels_per_cta;        /* 0 = default */
    int32_t path;                  /* nd_lift_path */
    int32_t grid_x, grid_y, grid_z; /* optional: the voxel lattice behind `points` (Z fastest, X*Y*Z == n_voxels),
                                      as in get_points (nerfdet.py:381-390); lets the plane-resident kernel
                                      use spatially compact warp tiles.  0 = unknown (results are identical) */
    int32_t sm_limit;              /* plane-resident kernel: occupy at most this many SMs (0 = all), e.g. to leave room
                                      for a concurrent exchange kernel */
    int32_t views_per_stage;       /* plane-resident kernel: views per pipeline stage, 1 or 2 (0 = default 2) */
    int32_t stages;                /* plane-resident kernel: pipeline stages, 2..8 (0 = as many as fit) */
    int32_t prefetch_stages;       /* plane-resident kernel: L2 prefetch distance in stages (0 = off) */
} nd_lift_options;

int nd_version(void);
const char *nd_last_error_string(void);

/* ---------------------------------------------------------------------------------------
 * B3  nerfdet.py:396-403  (projection part of `backproject`)
 * points [3][N] f32, projection [nv][3][4] f32 ->
 *   x, y  int64 [nv][N]  (round-half-even of q0/q2, q1/q2; unspecified where !valid)
 *   valid uint8 [nv][N]  = x>=0 & y>=0 & x<width & y<height & q2>0
 * Bit-exact contract: K=4 FMA chain in k order, IEEE divide, rint.
 * ------------------------------------------------------------------------------------- */
int nd_project_voxels(const float *points, const float *projection, int n_views, int64_t n_voxels,
                      int height, int width, int64_t *x, int64_t *y, uint8_t *valid, void *stream);

/* ---------------------------------------------------------------------------------------
 * B3+B5  nerfdet.py:393-420  `backproject(features, points, projection, depth, voxel_size)`
 * Compatibility (materialising) path: volume f32 [nv][C][N] (0 where invalid),
 * valid uint8 [nv][N].  depth_resized: optional f32 [nv][height][width], the depth map
 * already bilinearly resized to the feature resolution (nerfdet.py:406); voxel_z =
 * voxel_size[-1] of the gate at nerfdet.py:409-410.
 * ------------------------------------------------------------------------------------- */
int nd_backproject(const nd_maps *features, const float *points, const float *projection,
                   int64_t n_voxels, const float *depth_resized, float voxel_z,
                   float *volume, uint8_t *valid, void *stream);

/* ---------------------------------------------------------------------------------------
 * B3+B5+B6  nerfdet.py:164-181  fused: project + nearest gather + masked mean +
 * all-view variance + count, without materialising the per-view volume.
 *   mean  f32 [C][N]            (0 where count == 0)
 *   cov   f32 [C][N] or NULL    exp(-var), var summed over ALL views / count (0 where count == 0)
 *   count int64 [N]
 *   alpha f32 [N] or NULL: when given, `mean` receives alpha*mean (nerfdet.py:259-261).
 * workspace: nd_lift_workspace_bytes(), 256-byte aligned.
 * ------------------------------------------------------------------------------------- */
size_t nd_lift_workspace_bytes(const nd_maps *features, int64_t n_voxels, const nd_lift_options *opt);

/* Number of kernel launches nd_lift_mean_var / nd_lift_accumulate will issue for this input
 * (bench.py reports it as gpu_launches). */
int nd_lift_launch_count(const nd_maps *features, int64_t n_voxels, const nd_lift_options *opt);

int nd_lift_mean_var(const nd_maps *features, const float *points, const float *projection,
                     int64_t n_voxels, const float *alpha, float *mean, float *cov, int64_t *count,
                     void *workspace, size_t workspace_bytes, const nd_lift_options *opt, void *stream);

/* ---------------------------------------------------------------------------------------
 * The same in two steps: the GEOMETRY PLAN (everything that depends on points / projection / depth only: pixel
 * offsets, view counts, frustum culling, work distribution) is built once and reused for every feature stack lifted
 * with the same cameras and lattice -- nerfdet.py:155-160 recomputes projection and points per scene, so one plan per
 * scene; several lifts of one scene (features + any further stack, both GPUs' halves of a sharded lift, a fixed camera
 * rig) share it.
 *   nd_lift_plan_bytes   size of the caller-owned plan buffer (256-byte aligned); 0 = these maps are not eligible for
 *                        the plane-resident kernel (non-contiguous planes, planes > 64 KB, > 254 views, > 524288 voxels):
 *                        use nd_lift_mean_var, which then takes the staged path.  Only dtype / shape / strides of
 *                        `features` are read.
 *   nd_lift_plan_build   depth_resized f32 [nv][height][width] or NULL and voxel_z: the depth gate of nerfdet.py:405-411.
 *   nd_lift_plan_mean_var / nd_lift_plan_accumulate   outputs as nd_lift_mean_var / nd_lift_accumulate.
 *     launch_index: number of nd_lift_plan_* launches already issued on this plan buffer since nd_lift_plan_build,
 *       counted by the caller (0, 1, 2, ...; the library keeps no state).  All launches on one plan must use the same
 *       channel count and options, and be issued on one stream (or otherwise ordered).
 *     n_views_total: divisor-side view count of the all-view variance (0 = the plan's own view count).
 * Consecutive launches overlap through programmatic dependent launch: a launch reads its inputs as soon as SMs are
 * free and waits for its predecessor only before its first write.
 * ------------------------------------------------------------------------------------- */
size_t nd_lift_plan_bytes(const nd_maps *features, int64_t n_voxels, const nd_lift_options *opt);
int nd_lift_plan_build(const nd_maps *features, const float *points, const float *projection, int64_t n_voxels,
                       const float *depth_resized, float voxel_z, void *plan, size_t plan_bytes,
                       const nd_lift_options *opt, void *stream);
int nd_lift_plan_mean_var(const nd_maps *features, const void *plan, size_t plan_bytes, int64_t n_voxels,
                          uint32_t launch_index, int n_views_total, const float *alpha, float *mean, float *cov,
                          int64_t *count, const nd_lift_options *opt, void *stream);
int nd_lift_plan_accumulate(const nd_maps *features, const void *plan, size_t plan_bytes, int64_t n_voxels,
                            uint32_t launch_index, float *s1, float *s2, float *cnt, const nd_lift_options *opt,
                            void *stream);

/* ---------------------------------------------------------------------------------------
 * View-sharded form of the same: each rank runs nd_lift_accumulate on
 * its own views, the caller all-reduces (sum) the accumulators, then nd_lift_finalize with
 * the GLOBAL view count.
 *   s1, s2 f32 [C][N]  (sum and sum of squares over this rank's valid views; s2 may be NULL when no variance is wanted)
 *   cnt    f32 [N]     (valid views on this rank; exact in fp32 up to 2^24)
 * ------------------------------------------------------------------------------------- */
int nd_lift_accumulate(const nd_maps *features, const float *points, const float *projection,
                       int64_t n_voxels, float *s1, float *s2, float *cnt,
                       void *workspace, size_t workspace_bytes, const nd_lift_options *opt, void *stream);

int nd_lift_finalize(const float *s1, const float *s2, const float *cnt, int n_views_total,
                     int channels, int64_t n_voxels, const float *alpha,
                     float *mean, float *cov, int64_t *count, void *stream);

/* ---------------------------------------------------------------------------------------
 * Backward of nd_lift_mean_var: what torch autograd computes for
 * nerfdet.py:164-181 with respect to `features` (the reference trains the backbone through this path).
 *   mean, cov, count   the forward's outputs (cov may be NULL when grad_cov is NULL); the forward must have run
 *                      WITHOUT alpha (the alpha product of nerfdet.py:259-261 stays an autograd op on the caller's side)
 *   grad_mean, grad_cov  f32 [C][N] incoming gradients, either may be NULL
 *   grad_features      [nv][C][height][width] CONTIGUOUS, dtype of `features`; every element is written
 *   depth_resized, voxel_z, n_views_total   as in the forward (n_views_total 0 = the view count of `features`)
 * The validity masks come from the same device function as the forward's (bit-identical).  The sums are gathers over a
 * CSR of the voxels of every pixel, added in ascending voxel order: the gradient is deterministic.  Planes up to 20480
 * pixels (ND_ERR_BAD_SHAPE beyond), any lattice size (parts of 25 600 voxels, one gather launch per part).
 * workspace: nd_lift_backward_workspace_bytes(), 256-byte aligned.  2 + parts launches.
 * ------------------------------------------------------------------------------------- */
size_t nd_lift_backward_workspace_bytes(const nd_maps *features, int64_t n_voxels);
int nd_lift_backward(const nd_maps *features, const float *points, const float *projection, int64_t n_voxels,
                     const float *depth_resized, float voxel_z, int n_views_total, const float *mean, const float *cov,
                     const int64_t *count, const float *grad_mean, const float *grad_cov, void *grad_features,
                     void *workspace, size_t workspace_bytes, void *stream);


/* ---------------------------------------------------------------------------------------
 * Exchange step of the view-sharded lift over NVLink peer memory (no counterpart in the
 * reference, which lifts one scene per GPU).  Replaces "all-reduce the accumulators, then nd_lift_finalize on every
 * rank" with one kernel per rank that loads its channel slice of S1 / S2 / count from EVERY rank's accumulators
 * (P2P reads), finalises it with the global view count (nerfdet.py:171-181) and stores the rows into EVERY rank's
 * mean / cov (P2P writes); a one-block wait kernel then holds `stream` until all peers are done with this rank's
 * buffers.  Ranks meet through epoch flags inside the segments, so no NCCL call is needed on the data path.
 *
 *   nd_peer_alloc   cudaMalloc + zero a segment on the current device and export its CUDA IPC handle (64 bytes, host).
 *                   The one place where the library owns device memory: an IPC handle covers a whole allocation.
 *   nd_peer_open    map another process's segment into the current device's address space (lazy peer access);
 *   nd_peer_close / nd_peer_free   undo the two above.
 *   nd_lift_finalize_peers
 *     acc_host / mean_host / cov_host / flags_host: HOST arrays of `world` DEVICE pointers, entry g = rank g's
 *       accumulators [S1 (C*N) | S2 (C*N) | count (N)] f32 as written by nd_lift_accumulate, its mean [C][N], its cov
 *       [C][N] and its flag block.  cov_host NULL: the variance is not wanted (the live path never reads the 256-channel
 *       volume_cov, nerfdet.py:179-181 vs :232-261) -- the accumulators are then [S1 (C*N) | count (N)] (nd_lift_accumulate
 *       with s2 == NULL) and half as many bytes cross the links.  Flag block: (ND_PEER_FLAG_WORDS uint32, zero before the first epoch);
 *       entry `rank` is the local segment.  world <= ND_MAX_PEERS.
 *     epoch: 1, 2, 3, ... the same on every rank for the same step.
 *     count: int64 [N] local, or NULL.  alpha: f32 [N] local or NULL (alpha * mean, nerfdet.py:259-261).
 *     acc_mc / mean_mc / cov_mc: NVLS multicast addresses of the same three buffers (one address that reaches every
 *       rank's copy through the NVSwitch), or all NULL.  When given, the sums are taken in the switch
 *       (multimem.ld_reduce) and the rows leave with one multimem.st, which roughly halves the bytes a GPU receives;
 *       the per-rank tables are then only used for the flag blocks.  Needs N % 4 == 0.
 *     max_ctas: 0 = a grid that fills the GPU; > 0 = at most that many one-per-SM CTAs (512 / 1024 threads) looping over
 *       the work, so that the exchange of one scene can run beside the accumulate of the next on the SMs that
 *       nd_lift_options.sm_limit keeps free (the exchange is bound by the links, not by the SMs).
 *     timeout_ms: bound of every wait for a peer (0 = 4000).
 *     owner_rank: -1 = every rank receives the finished rows (all-gather); r >= 0 = only rank r does (the rank whose
 *       scene this is: a data-parallel detector runs neck and heads of a scene on one GPU) -- the other ranks' outputs
 *       are not written, and a GPU's links carry (G-1)/G x 52.5 MB of partial sums plus, averaged over rotating owners,
 *       1/G of the rows instead of all of them.  All ranks must pass the same value.
 *   Outputs are complete on `stream` when the call's kernels have run.  A peer that does not arrive within the
 *   time-out raises word 2 * ND_MAX_PEERS + 1 (the error word) of EVERY rank's flag block instead of hanging; a rank
 *   that finds its error word set performs no reduce and no peer store, fills the rows it owns with NaN in its own
 *   mean / cov and still completes the hand-shake, so the step fails visibly on every rank (the caller polls the
 *   error word, e.g. with an asynchronous copy to pinned host memory after each step).
 * ------------------------------------------------------------------------------------- */
#define ND_MAX_PEERS 8
#define ND_PEER_FLAG_WORDS 32

int nd_peer_alloc(size_t bytes, void **ptr, unsigned char *handle64_host);
int nd_peer_open(const unsigned char *handle64_host, void **ptr);
int nd_peer_close(void *ptr);
int nd_peer_free(void *ptr);
int nd_lift_finalize_peers(const void *const *acc_host, void *const *mean_host, void *const *cov_host,
                           void *const *flags_host, int world, int rank, uint32_t epoch, int n_views_total,
                           int channels, int64_t n_voxels, const float *alpha, int64_t *count, const void *acc_mc,
                           void *mean_mc, void *cov_mc, int max_ctas, int timeout_ms, int owner_rank, void *stream);


/* ---------------------------------------------------------------------------------------
 * B7  nerfdet.py:190-197  the per-pixel Linear(C -> 32) ("mapping") on the sliced 2-D features:
 *   mapped[v][y][x][j] = bias[j] + sum_c weight[j][c] * features[v][c][y][x]      (fp32 FMA, channels ascending)
 * features: NCHW planes read in place (contiguous planes, 16-byte aligned; f32 or bf16); weight f32 [32][C]
 * row-major (the reference nn.Linear.weight), bias f32 [32] or NULL; mapped f32 [nv][h][w][32] -- the
 * CHANNELS-LAST layout nd_live_stats and nd_render_gather_stats gather from (stride_c == 1).
 * ------------------------------------------------------------------------------------- */
int nd_map_features(const nd_maps *features, const float *weight, const float *bias, int out_channels,
                    float *mapped, void *stream);

/* ---------------------------------------------------------------------------------------
 * B8+B9  nerfdet.py:200-210, 232-253  live 35-channel voxel statistics that feed the density MLP.
 *   mapped  [nv][Cm][Hf][Wf] f32/bf16: the 2-D mapped features (B7, nerfdet.py:190-197), gathered with the
 *           FEATURE-level projection; an invalid voxel-view contributes map_bias (Linear(0), nerfdet.py:233-237)
 *   rgb     [nv][3][h][w] f32: denorm_images[:, :, :h, :w], gathered with its own stride-1 projection
 *           (0 where that projection is invalid; its validity does not enter the count, nerfdet.py:204-210)
 *   global_volume f32 [N][2*(3+Cm)]: channel-INTERLEAVED rows [m0, c0, m1, c1, ...] exactly as the reference's
 *           cat(dim=1) + view produces them (nerfdet.py:251-253); mean is NOT zeroed where count == 0
 *   mean35 / cov35 f32 [3+Cm][N] (optional, may be NULL), count int64 [N] (optional)
 * ------------------------------------------------------------------------------------- */
int nd_live_stats(const nd_maps *mapped, const nd_maps *rgb, const float *points, const float *projection,
                  const float *rgb_projection, int64_t n_voxels, const float *map_bias, float *global_volume,
                  float *mean35, float *cov35, int64_t *count, void *stream);

/* The same statistics with the depth gate of backproject (row B4, nerfdet.py:405-411) on BOTH gathers, as
 * extract_feat(depth=...) applies it (nerfdet.py:164-169 for the features that `mapping` then sees, :204-210 for the RGB
 * volume): a voxel-view that projects into the map is kept only if |z - depth[v][y][x]| < voxel_z.
 *   depth_mapped  f32 [nv][h][w]  the depth maps resized to the MAPPED feature resolution (F.interpolate bilinear, done
 *                 by the caller like the reference does), depth_rgb f32 [nv][H][W] resized to the rgb maps' resolution;
 *                 both NULL = no gate (nd_live_stats).  A gated-out view counts as invalid: bias for the mapped
 *                 channels, 0 for RGB, not counted. */
int nd_live_stats_gated(const nd_maps *mapped, const nd_maps *rgb, const float *points, const float *projection,
                        const float *rgb_projection, int64_t n_voxels, const float *map_bias, const float *depth_mapped,
                        const float *depth_rgb, float voxel_z, float *global_volume, float *mean35, float *cov35,
                        int64_t *count, void *stream);

/* Backward of the mapped channels of nd_live_stats (autograd of nerfdet.py:232-253 with
 * respect to the mapped features and, through the invalid views, the mapping's bias -- SURVEY.md section 0.6).
 *   mapped          f32 channels-last [nv][h][w][Cm] contiguous, the forward's input
 *   global_volume   the forward's output [N][2 * (3 + Cm)], grad_global_volume the incoming gradient
 *   grad_mapped     f32 [nv][h][w][Cm] and grad_bias f32 [Cm], both ZEROED by the caller (accumulated with reductions)
 * The RGB channels carry no gradient (input images). */
int nd_live_stats_bwd(const nd_maps *mapped, const float *points, const float *projection, int64_t n_voxels,
                      const float *map_bias, const float *global_volume, const float *grad_global_volume,
                      float *grad_mapped, float *grad_bias, void *stream);

/* Backward of nd_live_stats_gated: the same validity (depth gate on the mapped gather included) as its forward. */
int nd_live_stats_bwd_gated(const nd_maps *mapped, const float *points, const float *projection, int64_t n_voxels,
                            const float *map_bias, const float *depth_mapped, float voxel_z, const float *global_volume,
                            const float *grad_global_volume, float *grad_mapped, float *grad_bias, void *stream);

/* ---------------------------------------------------------------------------------------
 * M  nerf_mlp.py:11-234  VanillaNeRFRadianceField (NerfMLP + SinusoidalEncoder), as instantiated at
 * nerfdet.py:62-69.  The struct carries the reference state_dict tensors (row-major [out][in] f32 DEVICE
 * pointers; only the dimensions are read by nd_mlp_packed_bytes / nd_nerf_mlp_fwd).
 * ------------------------------------------------------------------------------------- */
typedef struct nd_mlp_weights {
    const float *base_w[8], *base_b[8];              /* mlp.base.hidden_layers.<i>.{weight,bias} */
    const float *sigma_w, *sigma_b;                  /* mlp.sigma_layer.output_layer */
    const float *bottleneck_w, *bottleneck_b;        /* mlp.bottleneck_layer.output_layer */
    const float *rgb_hidden_w, *rgb_hidden_b;        /* mlp.rgb_layer.hidden_layers.0 */
    const float *rgb_out_w, *rgb_out_b;              /* mlp.rgb_layer.output_layer */
    int32_t net_depth, net_width, skip_layer, feature_dim, cond_width, pos_octaves, view_octaves, reserved;
} nd_mlp_weights;

size_t nd_mlp_packed_bytes(const nd_mlp_weights *w);                       /* 0 = unsupported architecture */
/* Transposes the weights once into the caller-owned `packed` buffer (k-major, paddings zeroed). */
int nd_pack_mlp_weights(const nd_mlp_weights *w, void *packed, size_t packed_bytes, void *stream);
/* forward (nerf_mlp.py:217-234) / query_density (nerf_mlp.py:224-227) for P points:
 *   x [P][3], features [P][feature_dim], cond [P / samples_per_ray][3] (ray directions, broadcast over the
 *   samples of a ray, nerf_mlp.py:153-157) ->  sigma [P] (relu'd), alpha [P] = 1 - exp(-sigma) (nerfdet.py:258),
 *   rgb [P][3] (sigmoid'd).  Any output may be NULL; rgb == NULL or cond == NULL stops after the density head. */
int nd_nerf_mlp_fwd(const nd_mlp_weights *arch, const void *packed, const float *x, const float *features,
                    const float *cond, int64_t n_points, int samples_per_ray, float *sigma, float *alpha, float *rgb,
                    void *stream);

/* The same network on the tcgen05 tensor cores: bf16 operands, fp32 accumulation in tensor memory (csrc/mlp_tc.cu).
 * This is the path BASELINE.json's 1e-2 (bf16) tolerance applies to; nd_nerf_mlp_fwd stays the 1e-4 (fp32) one.
 * Own packed layout (every 64-column K block of every layer stored as the shared-memory image of a K-major
 * 128-byte-swizzled UMMA operand, then biases / head rows in fp32); `packed` must be 256-byte aligned.
 * Supported: net_width 256, net_width_condition 128, 63 + feature_dim <= 144, octaves 10 / 4, net_depth <= 8
 * (nd_mlp_tc_packed_bytes returns 0 otherwise).  Same arguments and outputs as nd_nerf_mlp_fwd. */
size_t nd_mlp_tc_packed_bytes(const nd_mlp_weights *w);
int nd_pack_mlp_weights_tc(const nd_mlp_weights *w, void *packed, size_t packed_bytes, void *stream);
int nd_nerf_mlp_fwd_tc(const nd_mlp_weights *arch, const void *packed, const float *x, const float *features,
                       const float *cond, int64_t n_points, int samples_per_ray, float *sigma, float *alpha, float *rgb,
                       void *stream);

/* The same tensor-core kernel at fp32-grade precision ("3 x bf16"): every operand -- inputs, activations, weights -- is
 * carried as a hi + lo pair of bf16 numbers (16 mantissa bits) and every product as hi*hi + lo*hi + hi*lo with fp32
 * accumulation in tensor memory; the positional encoding rounds its arguments where the reference does
 * (nerf_mlp.py:189-196).  Three times the tensor-core work of nd_nerf_mlp_fwd_tc, within BASELINE.json's 1e-4 (fp32)
 * tolerance: the default precision of the drop-in module.  Own packed layout (hi and lo weight images); same supported
 * architectures, arguments and outputs as nd_nerf_mlp_fwd_tc. */
size_t nd_mlp_tc3_packed_bytes(const nd_mlp_weights *w);
int nd_pack_mlp_weights_tc3(const nd_mlp_weights *w, void *packed, size_t packed_bytes, void *stream);
int nd_nerf_mlp_fwd_tc3(const nd_mlp_weights *arch, const void *packed, const float *x, const float *features,
                        const float *cond, int64_t n_points, int samples_per_ray, float *sigma, float *alpha, float *rgb,
                        void *stream);

/* ---------------------------------------------------------------------------------------
 * R2  render_ray.py:145-189  sample_along_camera_ray: z_vals [R][S] = near + i * (far - near) / (S - 1), optional
 * stratified jitter with caller-supplied uniforms t_rand [R][S] (torch.rand_like drawn by the caller; NULL = det),
 * pts [R][S][3] = z * d + o (mul and add rounded separately).
 * ------------------------------------------------------------------------------------- */
int nd_sample_rays(const float *ray_o, const float *ray_d, int64_t n_rays, int n_samples, float near_depth,
                   float far_depth, const float *t_rand, float *pts, float *z_vals, void *stream);

/* ---------------------------------------------------------------------------------------
 * R4+R5+R6  projection.py:24-151 (Projector.compute, grid_sample branch) + render_ray.py:71-93, 301-303.
 *   pts [P][3]; cameras [nv][34] = [h, w, K 4x4, E 4x4] (render_ray.py:48-69); images [nv][3][Hp][Wp] f32 and
 *   featmaps [nv][D][Hf][Wf] f32/bf16, both contiguous planes, sampled bilinearly (zeros padding,
 *   align_corners = True) at the SAME normalised coordinates.
 *   globalfeat f32 [P][2*(3+D)] = [mean(3+D), exp(-var)(3+D)] (concatenated, render_ray.py:303)
 *   view_mask u8 [P][nv] (inbound & in-front, optional), pixel_mask u8 [P] = (sum of view_mask > 1, optional),
 *   pixel_locations f32 [nv][P][2] (optional; clamped like projection.py:61), in_front u8 [nv][P] (optional),
 *   view_features f32 [P][nv][3+D] (optional): the materialised per-view samples Projector.compute returns
 *   (projection.py:147, compatibility path only).
 * The reference's [rays, samples, views, 3+D] tensor is never materialised.
 * ------------------------------------------------------------------------------------- */
int nd_render_gather_stats(const float *pts, int64_t n_points, const float *cameras, int n_views,
                           const nd_maps *images, const nd_maps *featmaps, float *globalfeat, uint8_t *view_mask,
                           uint8_t *pixel_mask, float *pixel_locations, uint8_t *in_front, float *view_features, void *stream);

/* ---------------------------------------------------------------------------------------
 * Backward of nd_render_gather_stats with respect to the mapped feature maps (what autograd
 * computes for projection.py:91-151 + render_ray.py:71-93).
 *   featmaps        f32 channels-last [nv][h][w][D] contiguous (D <= 32, D % 4 == 0): the forward's input
 *   globalfeat      f32 [P][2 * (3 + D)]: the forward's output;  grad_globalfeat: the incoming gradient, same shape
 *   grad_featmaps   f32 [nv][h][w][D], ZEROED by the caller: the gradient is accumulated with vector reductions
 *   image_height / image_width: the size of the image stack the forward sampled (only its masks enter here)
 * The images and the sample positions get no gradient.
 * ------------------------------------------------------------------------------------- */
int nd_render_gather_stats_bwd(const float *pts, int64_t n_points, const float *cameras, int n_views, int image_height,
                               int image_width, const nd_maps *featmaps, const float *globalfeat, const float *grad_globalfeat,
                               float *grad_featmaps, void *stream);

/* ---------------------------------------------------------------------------------------
 * R7  render_ray.py:196-247  raw2outputs: alpha = 1 - exp(-sigma) (no interval term), T = cumprod(1 - alpha + 1e-10)
 * shifted, weights = alpha * T, rgb = sum w rgb, depth = sum w z / (sum w + 1e-8) clamped to z_bounds = {min, max} of
 * the whole batch's z_vals (render_ray.py:236; 2 floats in DEVICE memory so that no host sync is needed),
 * ray_mask = sum(pixel_mask) > 8.
 *   rgb [R][S][3], sigma [R][S], z_vals [R][S], pixel_mask u8 [R][S] (optional)
 *   out_rgb [R][3], out_depth [R]; weights / alpha / transparency [R][S] and ray_mask u8 [R] optional.
 * ------------------------------------------------------------------------------------- */
int nd_composite(const float *rgb, const float *sigma, const float *z_vals, const uint8_t *pixel_mask, int64_t n_rays,
                 int n_samples, const float *z_bounds, int white_bkgd, float *out_rgb, float *out_depth,
                 float *weights, float *alpha, float *transparency, uint8_t *ray_mask, void *stream);

/* ---------------------------------------------------------------------------------------
 * R8  render_ray.py:26-46  volume_sampling: trilinear grid_sample (border padding, align_corners = True) of
 * volume [C][D0][D1][D2] at (pts - aabb_min) * 2 / (aabb_max - aabb_min) - 1, normalised x indexing the LAST
 * axis (applied literally like the reference).  out [P][C], inside u8 [P] (all three strictly within (-1, 1)).
 * aabb_* are HOST pointers to 3 floats.
 * ------------------------------------------------------------------------------------- */
int nd_volume_sample_trilinear(const float *volume, int channels, int d0, int d1, int d2, const float *pts,
                               int64_t n_points, const float *aabb_min_host, const float *aabb_max_host, float *out,
                               uint8_t *inside, void *stream);

/* ---------------------------------------------------------------------------------------
 * Row N3: the producers of the render branch's inputs on the GPU.
 *
 * nd_generate_rays   datasets/pipelines/multi_view.py:124-132 + data_augment_utils.py:410-424 (get_dtu_raydir) +
 *   formating.py:70-75.  intrinsic3x3_host: the NeRF intrinsics (rows 0-1 already divided by ori_h / img_h), row-major,
 *   HOST memory; rot f64 [nt][3][3] (camrotc2w) and lightpos f32 [nt][3] on the device.  Pixel grid
 *   [margin, width - margin) x [margin, height - margin), rows = y.  ray_d, ray_o f32 [nt][(H-2m)*(W-2m)][3].
 * nd_denorm_images   multi_view.py:107-110 (mmcv.imdenormalize(img, mean, std, to_bgr).astype(uint8) / 255.0) +
 *   formating.py:87-91.  img f32 [n][3][H][W] = the normalised network input; mean3_host / std3_host f64 HOST;
 *   out f32 [n][3][H][W] in [0, 1], channels swapped when to_bgr.
 * ------------------------------------------------------------------------------------- */
int nd_generate_rays(const float *intrinsic3x3_host, const double *rot, const float *lightpos, int n_target_views, int height,
                     int width, int margin, float *ray_d, float *ray_o, void *stream);
int nd_denorm_images(const float *img, const double *mean3_host, const double *std3_host, int to_bgr, int64_t n_images,
                     int height, int width, float *out, void *stream);

/* ---------------------------------------------------------------------------------------
 * Row N4: the metrics of the render_testing evaluator, batched over the rendered views
 * (mmdet3d/models/model_utils/save_rendered_img.py:10-78; SSIM = skimage.metrics.structural_similarity of the pinned
 * scikit-image 0.18.1 as the reference ends up calling it: per channel, 7 x 7 uniform window, sample covariance,
 * K1 = 0.01, K2 = 0.03, float64, map cropped by 3 pixels; data_range = 2 for float images there).
 *   pred f32 [nv][H][W][3]; target f32 or f64 [nv][H][W][3]; psnr_ssim f64 [nv][2] = {PSNR, SSIM} per view.
 *   nd_depth_sqerr: out f64 [n_pixels] = mean over the views of (depth - gt_depth)^2 (the reference's "rsme" map).
 * ------------------------------------------------------------------------------------- */
size_t nd_image_metrics_workspace_bytes(int n_views, int height, int width);
int nd_image_metrics(const float *pred, const void *target, int target_is_f64, int n_views, int height, int width,
                     double data_range, double *psnr_ssim, void *workspace, size_t workspace_bytes, void *stream);
int nd_depth_sqerr(const float *depth, const void *gt_depth, int gt_is_f64, int n_views, int64_t n_pixels, double *out,
                   void *stream);

/* ---------------------------------------------------------------------------------------
 * Row N2: hand-over to FastIndoorImVoxelNeck (necks/imvoxelnet.py:8-67, nerfdet.py:262-267).
 *   volume f32 [C][N] (alpha * mean as the lift writes it) -> out [N][C] in out_dtype (ND_F32 / ND_BF16): the
 *   channels-last-3D layout of a [1, C, X, Y, Z] tensor, which cuDNN's Conv3d takes without a conversion pass;
 *   valid f32 [N] or NULL = (float)count, the `valids.float()` the head up-samples (nerfdet.py:287, imvoxel_head_v2.py:93).
 * ------------------------------------------------------------------------------------- */
int nd_volume_to_neck(const float *volume, const int64_t *count, int channels, int64_t n_voxels, int out_dtype, void *out,
                      float *valid, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFDET_LIFT_H_ */
