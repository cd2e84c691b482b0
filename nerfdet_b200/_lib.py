"""ctypes binding of ``libnerfdet_lift.so`` (C ABI declared in include/nerfdet_lift.h).

There is NO fallback: if the library is missing the import of any op raises, and on a
machine with a GPU every op runs the CUDA kernels or fails loudly."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

from . import build as _build

ND_F32, ND_BF16 = 0, 1
ND_LIFT_PATH_AUTO, ND_LIFT_PATH_STAGED = 0, 1
ND_MAX_PEERS, ND_PEER_FLAG_WORDS = 8, 32


class NdMaps(ctypes.Structure):
    _fields_ = [('data', c_void_p), ('dtype', c_int32), ('n_views', c_int32), ('channels', c_int32),
                ('height', c_int32), ('width', c_int32), ('stride_v', c_int64), ('stride_c', c_int64),
                ('stride_y', c_int64), ('stride_x', c_int64)]


class NdLiftOptions(ctypes.Structure):
    _fields_ = [('scratch_budget_bytes', c_size_t), ('voxels_per_cta', c_int32), ('path', c_int32),
                ('grid_x', c_int32), ('grid_y', c_int32), ('grid_z', c_int32), ('sm_limit', c_int32),
                ('views_per_stage', c_int32), ('stages', c_int32), ('prefetch_stages', c_int32)]


class NdMlpWeights(ctypes.Structure):
    _fields_ = [('base_w', c_void_p * 8), ('base_b', c_void_p * 8),
                ('sigma_w', c_void_p), ('sigma_b', c_void_p),
                ('bottleneck_w', c_void_p), ('bottleneck_b', c_void_p),
                ('rgb_hidden_w', c_void_p), ('rgb_hidden_b', c_void_p),
                ('rgb_out_w', c_void_p), ('rgb_out_b', c_void_p),
                ('net_depth', c_int32), ('net_width', c_int32), ('skip_layer', c_int32), ('feature_dim', c_int32),
                ('cond_width', c_int32), ('pos_octaves', c_int32), ('view_octaves', c_int32), ('reserved', c_int32)]


# name -> (restype, argtypes); must list every symbol of include/nerfdet_lift.h
SIGNATURES = {
    'nd_version': (c_int, []),
    'nd_last_error_string': (ctypes.c_char_p, []),
    'nd_project_voxels': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    'nd_backproject': (c_int, [POINTER(NdMaps), c_void_p, c_void_p, c_int64, c_void_p, c_float, c_void_p,
                               c_void_p, c_void_p]),
    'nd_lift_workspace_bytes': (c_size_t, [POINTER(NdMaps), c_int64, POINTER(NdLiftOptions)]),
    'nd_lift_launch_count': (c_int, [POINTER(NdMaps), c_int64, POINTER(NdLiftOptions)]),
    'nd_lift_mean_var': (c_int, [POINTER(NdMaps), c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_size_t, POINTER(NdLiftOptions), c_void_p]),
    'nd_lift_accumulate': (c_int, [POINTER(NdMaps), c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_size_t, POINTER(NdLiftOptions), c_void_p]),
    'nd_lift_plan_bytes': (c_size_t, [POINTER(NdMaps), c_int64, POINTER(NdLiftOptions)]),
    'nd_lift_plan_build': (c_int, [POINTER(NdMaps), c_void_p, c_void_p, c_int64, c_void_p, c_float, c_void_p, c_size_t,
                                   POINTER(NdLiftOptions), c_void_p]),
    'nd_lift_plan_mean_var': (c_int, [POINTER(NdMaps), c_void_p, c_size_t, c_int64, ctypes.c_uint32, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, POINTER(NdLiftOptions), c_void_p]),
    'nd_lift_plan_accumulate': (c_int, [POINTER(NdMaps), c_void_p, c_size_t, c_int64, ctypes.c_uint32, c_void_p, c_void_p,
                                        c_void_p, POINTER(NdLiftOptions), c_void_p]),
    'nd_lift_finalize': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    'nd_lift_backward_workspace_bytes': (c_size_t, [POINTER(NdMaps), c_int64]),
    'nd_lift_backward': (c_int, [POINTER(NdMaps), c_void_p, c_void_p, c_int64, c_void_p, ctypes.c_float, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'nd_generate_rays': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'nd_denorm_images': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p]),
    'nd_image_metrics_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'nd_image_metrics': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.c_double, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    'nd_depth_sqerr': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p]),
    'nd_volume_to_neck': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    'nd_render_gather_stats_bwd': (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, POINTER(NdMaps), c_void_p, c_void_p,
                                           c_void_p, c_void_p]),
    'nd_live_stats_bwd': (c_int, [POINTER(NdMaps), c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    'nd_peer_alloc': (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    'nd_peer_open': (c_int, [c_void_p, POINTER(c_void_p)]),
    'nd_peer_close': (c_int, [c_void_p]),
    'nd_peer_free': (c_int, [c_void_p]),
    'nd_lift_finalize_peers': (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                       c_int, c_int, ctypes.c_uint32, c_int, c_int, c_int64, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'nd_live_stats': (c_int, [POINTER(NdMaps), POINTER(NdMaps), c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nd_live_stats_gated': (c_int, [POINTER(NdMaps), POINTER(NdMaps), c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                    c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nd_live_stats_bwd_gated': (c_int, [POINTER(NdMaps), c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_float, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    'nd_map_features': (c_int, [POINTER(NdMaps), c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    'nd_mlp_packed_bytes': (c_size_t, [POINTER(NdMlpWeights)]),
    'nd_pack_mlp_weights': (c_int, [POINTER(NdMlpWeights), c_void_p, c_size_t, c_void_p]),
    'nd_nerf_mlp_fwd': (c_int, [POINTER(NdMlpWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    'nd_mlp_tc_packed_bytes': (c_size_t, [POINTER(NdMlpWeights)]),
    'nd_pack_mlp_weights_tc': (c_int, [POINTER(NdMlpWeights), c_void_p, c_size_t, c_void_p]),
    'nd_nerf_mlp_fwd_tc': (c_int, [POINTER(NdMlpWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    'nd_mlp_tc3_packed_bytes': (c_size_t, [POINTER(NdMlpWeights)]),
    'nd_pack_mlp_weights_tc3': (c_int, [POINTER(NdMlpWeights), c_void_p, c_size_t, c_void_p]),
    'nd_nerf_mlp_fwd_tc3': (c_int, [POINTER(NdMlpWeights), c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    'nd_sample_rays': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    'nd_render_gather_stats': (c_int, [c_void_p, c_int64, c_void_p, c_int, POINTER(NdMaps), POINTER(NdMaps),
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nd_composite': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'nd_volume_sample_trilinear': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64,
                                           POINTER(c_float), POINTER(c_float), c_void_p, c_void_p, c_void_p]),
}

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load():
    """Loads (building first if the sources changed and nvcc is present) the C-ABI library."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not _build.is_current():
        try:
            _build.build_library()
        except Exception as e:  # no nvcc on this machine: use the shipped .so if there is one
            if not os.path.isfile(path):
                raise RuntimeError(
                    f'libnerfdet_lift.so is missing and could not be built ({e}); '
                    'the nerfdet_b200 ops have no CPU or eager fallback') from e
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().nd_last_error_string()
        raise RuntimeError(f'{what} failed with nd_status {status}: {msg.decode() if msg else ""}')
