"""torch custom ops (``torch.ops.nerfdet_b200.*``) over the C ABI of libnerfdet_lift.so.

Each op validates shapes / dtypes / devices in Python (the reference itself only asserts
``stride == 4`` and ``B == 1``), allocates outputs and workspaces with torch, and launches
on ``torch.cuda.current_stream()``.  CUDA tensors only -- there is no CPU path."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import ND_BF16, ND_F32, NdLiftOptions, NdMaps

_NS = 'nerfdet_b200'


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors: Optional[Tensor]):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('nerfdet_b200 ops run on CUDA tensors only (no CPU fallback)')


def _ptr(t: Optional[Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _maps(features: Tensor) -> NdMaps:
    if features.dim() != 4:
        raise ValueError(f'features must be [n_views, C, H, W], got {tuple(features.shape)}')
    if features.dtype == torch.float32:
        dt = ND_F32
    elif features.dtype == torch.bfloat16:
        dt = ND_BF16
    else:
        raise TypeError(f'features must be float32 or bfloat16, got {features.dtype}')
    nv, c, h, w = features.shape
    sv, sc, sy, sx = features.stride()
    return NdMaps(features.data_ptr(), dt, nv, c, h, w, sv, sc, sy, sx)


def _check_geometry(points: Tensor, projection: Tensor, n_views: int):
    if points.dtype != torch.float32 or projection.dtype != torch.float32:
        raise TypeError('points and projection must be float32')
    if points.dim() not in (2, 4) or points.shape[0] != 3:
        raise ValueError(f'points must be [3, N] or [3, X, Y, Z], got {tuple(points.shape)}')
    if tuple(projection.shape) != (n_views, 3, 4):
        raise ValueError(f'projection must be [{n_views}, 3, 4], got {tuple(projection.shape)}')


def _flat_points(points: Tensor) -> Tuple[Tensor, Tuple[int, int, int]]:
    """[3, N] points plus the lattice shape when ``points`` came as [3, X, Y, Z] (get_points)."""
    grid = tuple(points.shape[1:]) if points.dim() == 4 else (0, 0, 0)
    return points.reshape(3, -1).contiguous(), grid


def _options(scratch_budget_bytes: int, grid=(0, 0, 0)) -> NdLiftOptions:
    """``scratch_budget_bytes`` > 0 selects the generic staged path (any strides) with that much
    L2-resident staging; 0 = automatic (plane-resident kernel for contiguous NCHW planes)."""
    path = _lib.ND_LIFT_PATH_STAGED if scratch_budget_bytes > 0 else _lib.ND_LIFT_PATH_AUTO
    return NdLiftOptions(max(scratch_budget_bytes, 0), 0, path, int(grid[0]), int(grid[1]), int(grid[2]), 0)


# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{_NS}::project_voxels', mutates_args=())
def project_voxels(points: Tensor, projection: Tensor, height: int, width: int) -> Tuple[Tensor, Tensor, Tensor]:
    """x, y int64 [nv, N]; valid bool [nv, N]  (reference nerfdet.py:396-403)."""
    _need_cuda(points, projection)
    nv = projection.shape[0]
    _check_geometry(points, projection, nv)
    points = points.contiguous()
    projection = projection.contiguous()
    n = points.shape[1]
    x = torch.empty((nv, n), dtype=torch.int64, device=points.device)
    y = torch.empty_like(x)
    valid = torch.empty((nv, n), dtype=torch.bool, device=points.device)
    lib = _lib.load()
    _lib.check(lib.nd_project_voxels(_ptr(points), _ptr(projection), nv, n, height, width, _ptr(x), _ptr(y),
                                     _ptr(valid), _stream()), 'nd_project_voxels')
    return x, y, valid


@project_voxels.register_fake
def _(points, projection, height, width):
    nv, n = projection.shape[0], points.shape[1]
    x = points.new_empty((nv, n), dtype=torch.int64)
    return x, torch.empty_like(x), points.new_empty((nv, n), dtype=torch.bool)


# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{_NS}::backproject', mutates_args=())
def backproject(features: Tensor, points: Tensor, projection: Tensor, depth_resized: Optional[Tensor],
                voxel_z: float) -> Tuple[Tensor, Tensor]:
    """Materialised volume f32 [nv, C, N] and valid bool [nv, N]  (reference nerfdet.py:393-420)."""
    _need_cuda(features, points, projection, depth_resized)
    m = _maps(features)
    _check_geometry(points, projection, m.n_views)
    points = points.contiguous()
    projection = projection.contiguous()
    if depth_resized is not None:
        if depth_resized.dtype != torch.float32 or tuple(depth_resized.shape) != (m.n_views, m.height, m.width):
            raise ValueError('depth_resized must be float32 [n_views, H, W] at the feature resolution')
        depth_resized = depth_resized.contiguous()
    n = points.shape[1]
    volume = torch.empty((m.n_views, m.channels, n), dtype=torch.float32, device=features.device)
    valid = torch.empty((m.n_views, n), dtype=torch.bool, device=features.device)
    lib = _lib.load()
    _lib.check(lib.nd_backproject(ctypes.byref(m), _ptr(points), _ptr(projection), n, _ptr(depth_resized),
                                  float(voxel_z), _ptr(volume), _ptr(valid), _stream()), 'nd_backproject')
    return volume, valid


@backproject.register_fake
def _(features, points, projection, depth_resized, voxel_z):
    nv, c = features.shape[:2]
    n = points.shape[1]
    return (features.new_empty((nv, c, n), dtype=torch.float32), features.new_empty((nv, n), dtype=torch.bool))


# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{_NS}::lift_mean_var', mutates_args=())
def lift_mean_var(features: Tensor, points: Tensor, projection: Tensor, alpha: Optional[Tensor],
                  want_cov: bool, scratch_budget_bytes: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Fused project + gather + mean / all-view variance / count (reference nerfdet.py:164-181).
    Returns mean f32 [C, N] (times alpha when given), exp(-var) f32 [C, N] (empty when
    ``want_cov`` is False) and count int64 [N]."""
    _need_cuda(features, points, projection, alpha)
    m = _maps(features)
    _check_geometry(points, projection, m.n_views)
    points, grid = _flat_points(points)
    projection = projection.contiguous()
    n = points.shape[1]
    if alpha is not None:
        if alpha.dtype != torch.float32 or alpha.numel() != n:
            raise ValueError('alpha must be float32 with one value per voxel')
        alpha = alpha.contiguous()
    dev = features.device
    mean = torch.empty((m.channels, n), dtype=torch.float32, device=dev)
    cov = torch.empty((m.channels, n) if want_cov else (0,), dtype=torch.float32, device=dev)
    count = torch.empty((n,), dtype=torch.int64, device=dev)
    lib = _lib.load()
    opt = _options(scratch_budget_bytes, grid)
    optp = ctypes.byref(opt)
    ws_bytes = lib.nd_lift_workspace_bytes(ctypes.byref(m), n, optp)
    ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
    _lib.check(lib.nd_lift_mean_var(ctypes.byref(m), _ptr(points), _ptr(projection), n, _ptr(alpha), _ptr(mean),
                                    _ptr(cov) if want_cov else None, _ptr(count), _ptr(ws), ws_bytes, optp,
                                    _stream()), 'nd_lift_mean_var')
    return mean, cov, count


@lift_mean_var.register_fake
def _(features, points, projection, alpha, want_cov, scratch_budget_bytes):
    c, n = features.shape[1], points[0].numel()
    mean = features.new_empty((c, n), dtype=torch.float32)
    cov = features.new_empty((c, n) if want_cov else (0,), dtype=torch.float32)
    return mean, cov, features.new_empty((n,), dtype=torch.int64)


# ------------------------------------------------------------------------------------------
@torch.library.custom_op(f'{_NS}::lift_accumulate', mutates_args=())
def lift_accumulate(features: Tensor, points: Tensor, projection: Tensor, scratch_budget_bytes: int) -> Tensor:
    """Per-rank accumulators of the view-sharded lift, one flat f32 buffer
    ``[S1 (C*N) | S2 (C*N) | count (N)]`` ready for a single all-reduce."""
    _need_cuda(features, points, projection)
    m = _maps(features)
    _check_geometry(points, projection, m.n_views)
    points, grid = _flat_points(points)
    projection = projection.contiguous()
    n = points.shape[1]
    c = m.channels
    dev = features.device
    acc = torch.empty(((2 * c + 1) * n,), dtype=torch.float32, device=dev)
    lib = _lib.load()
    opt = _options(scratch_budget_bytes, grid)
    optp = ctypes.byref(opt)
    ws_bytes = lib.nd_lift_workspace_bytes(ctypes.byref(m), n, optp)
    ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
    base = acc.data_ptr()
    _lib.check(lib.nd_lift_accumulate(ctypes.byref(m), _ptr(points), _ptr(projection), n,
                                      ctypes.c_void_p(base), ctypes.c_void_p(base + 4 * c * n),
                                      ctypes.c_void_p(base + 8 * c * n), _ptr(ws), ws_bytes, optp, _stream()),
               'nd_lift_accumulate')
    return acc


@lift_accumulate.register_fake
def _(features, points, projection, scratch_budget_bytes):
    c, n = features.shape[1], points[0].numel()
    return features.new_empty(((2 * c + 1) * n,), dtype=torch.float32)


@torch.library.custom_op(f'{_NS}::lift_finalize', mutates_args=())
def lift_finalize(acc: Tensor, n_views_total: int, channels: int, n_voxels: int, alpha: Optional[Tensor],
                  want_cov: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """mean / exp(-var) / count from (all-reduced) accumulators and the GLOBAL view count."""
    _need_cuda(acc, alpha)
    c, n = channels, n_voxels
    if acc.dtype != torch.float32 or acc.numel() != (2 * c + 1) * n or not acc.is_contiguous():
        raise ValueError('acc must be the contiguous float32 buffer returned by lift_accumulate')
    if alpha is not None:
        alpha = alpha.contiguous()
    dev = acc.device
    mean = torch.empty((c, n), dtype=torch.float32, device=dev)
    cov = torch.empty((c, n) if want_cov else (0,), dtype=torch.float32, device=dev)
    count = torch.empty((n,), dtype=torch.int64, device=dev)
    base = acc.data_ptr()
    lib = _lib.load()
    _lib.check(lib.nd_lift_finalize(ctypes.c_void_p(base), ctypes.c_void_p(base + 4 * c * n),
                                    ctypes.c_void_p(base + 8 * c * n), n_views_total, c, n, _ptr(alpha),
                                    _ptr(mean), _ptr(cov) if want_cov else None, _ptr(count), _stream()),
               'nd_lift_finalize')
    return mean, cov, count


@lift_finalize.register_fake
def _(acc, n_views_total, channels, n_voxels, alpha, want_cov):
    mean = acc.new_empty((channels, n_voxels))
    cov = acc.new_empty((channels, n_voxels) if want_cov else (0,))
    return mean, cov, acc.new_empty((n_voxels,), dtype=torch.int64)


def lift_launch_count(features: Tensor, n_voxels: int, scratch_budget_bytes: int = 0) -> int:
    """Kernel launches one fused lift of ``features`` issues (for bench.py's ``gpu_launches``)."""
    m = _maps(features)
    opt = _options(scratch_budget_bytes)
    return int(_lib.load().nd_lift_launch_count(ctypes.byref(m), n_voxels, ctypes.byref(opt)))
