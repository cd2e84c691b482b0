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
    int32_t reserved;
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
 * View-sharded form of the same: each rank runs nd_lift_accumulate on
 * its own views, the caller all-reduces (sum) the accumulators, then nd_lift_finalize with
 * the GLOBAL view count.
 *   s1, s2 f32 [C][N]  (sum and sum of squares over this rank's valid views)
 *   cnt    f32 [N]     (valid views on this rank; exact in fp32 up to 2^24)
 * ------------------------------------------------------------------------------------- */
int nd_lift_accumulate(const nd_maps *features, const float *points, const float *projection,
                       int64_t n_voxels, float *s1, float *s2, float *cnt,
                       void *workspace, size_t workspace_bytes, const nd_lift_options *opt, void *stream);

int nd_lift_finalize(const float *s1, const float *s2, const float *cnt, int n_views_total,
                     int channels, int64_t n_voxels, const float *alpha,
                     float *mean, float *cov, int64_t *count, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFDET_LIFT_H_ */
