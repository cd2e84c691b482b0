"""TEST INFRASTRUCTURE -- CPU oracle for the voxel lifting path.  NOT product code.

A restatement, in plain torch-CPU fp32 ops, of what the reference computes in
``mmdet3d/models/detectors/nerfdet.py`` (citations are file:line into
/root/reference).  It is pinned against the reference itself: ``oracle/make_golden.py``
runs the unmodified reference in the build container and writes the fixtures
under ``tests/golden/`` that ``tests/test_oracle_golden.py`` checks this file
against.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
/ ``--impl reference`` legs may import it; the product package never does.

Parity status: the reference ships no test / golden vector for this path
 -> pinned by reference-generated fixtures, not by upstream tests.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# host-side geometry (reference: nerfdet.py:364-390)
# --------------------------------------------------------------------------- #
def compute_projection(img_meta, stride: int) -> torch.Tensor:
    """[nv, 3, 4] = K' @ E[:3] with K' = K[:3,:3], rows 0-1 divided by
    ``ori_h / (img_h / stride)``  (nerfdet.py:364-378, ``angles=None`` branch)."""
    k = torch.tensor(img_meta['lidar2img']['intrinsic'][:3, :3])
    ratio = img_meta['ori_shape'][0] / (img_meta['img_shape'][0] / stride)
    k[:2] /= ratio
    mats = [k @ torch.tensor(e)[:3] for e in img_meta['lidar2img']['extrinsic']]
    return torch.stack(mats)


def get_points(n_voxels, voxel_size, origin) -> torch.Tensor:
    """Voxel lattice [3, X, Y, Z]: idx * voxel_size + (origin - n/2 * voxel_size),
    min-corner convention (nerfdet.py:380-390)."""
    n_voxels = torch.as_tensor(n_voxels)
    voxel_size = torch.as_tensor(voxel_size, dtype=torch.float32)
    origin = torch.as_tensor(origin, dtype=torch.float32)
    axes = [torch.arange(int(n)) for n in n_voxels]
    grid = torch.stack(torch.meshgrid(axes, indexing='ij'))
    corner = origin - n_voxels / 2.0 * voxel_size
    return grid * voxel_size.view(3, 1, 1, 1) + corner.view(3, 1, 1, 1)


# --------------------------------------------------------------------------- #
# B3: projection + nearest pixel index (nerfdet.py:394-403)
# --------------------------------------------------------------------------- #
def project_voxels(points: torch.Tensor, projection: torch.Tensor, height: int, width: int):
    """Returns x, y (int64 [nv, N]), valid (bool [nv, N]) and the homogeneous
    coordinates q [nv, 3, N].  ``x = round(q0/q2)`` is round-half-to-even."""
    nv = projection.shape[0]
    flat = points.reshape(1, 3, -1).expand(nv, 3, -1)
    homo = torch.cat((flat, torch.ones_like(flat[:, :1])), dim=1)
    q = torch.bmm(projection, homo)
    x = (q[:, 0] / q[:, 2]).round().long()
    y = (q[:, 1] / q[:, 2]).round().long()
    valid = (x >= 0) & (y >= 0) & (x < width) & (y < height) & (q[:, 2] > 0)
    return x, y, valid, q


def depth_gate(valid, x, y, z, depth, height, width, voxel_z: float):
    """B4 (nerfdet.py:405-411): keep a voxel-view only if |z - depth[y,x]| < voxel_z,
    depth bilinearly resized to the feature resolution."""
    d = F.interpolate(depth.unsqueeze(1), size=(height, width), mode='bilinear').squeeze(1)
    out = valid.clone()
    for i in range(valid.shape[0]):
        keep = z[i] > 0
        sel = valid[i]
        dz = d[i, y[i, sel], x[i, sel]]
        keep[sel] = (z[i, sel] > dz - voxel_z) & (z[i, sel] < dz + voxel_z)
        out[i] &= keep
    return out


# --------------------------------------------------------------------------- #
# B5: materialising gather (nerfdet.py:414-420)
# --------------------------------------------------------------------------- #
def backproject(features: torch.Tensor, points: torch.Tensor, projection: torch.Tensor,
                depth: Optional[torch.Tensor] = None, voxel_size: Optional[Sequence[float]] = None):
    """volume [nv, C, X, Y, Z] f32 (0 where invalid), valid [nv, 1, X, Y, Z] bool."""
    nv, c, height, width = features.shape
    gx, gy, gz = points.shape[-3:]
    x, y, valid, q = project_voxels(points, projection, height, width)
    if depth is not None:
        valid = depth_gate(valid, x, y, q[:, 2], depth, height, width, voxel_size[-1])
    n = gx * gy * gz
    volume = torch.zeros((nv, c, n), dtype=torch.float32)
    for i in range(nv):
        sel = valid[i]
        volume[i][:, sel] = features[i][:, y[i, sel], x[i, sel]].float()
    return volume.view(nv, c, gx, gy, gz), valid.view(nv, 1, gx, gy, gz)


# --------------------------------------------------------------------------- #
# B6: mean / all-view variance / count (nerfdet.py:171-181)
# --------------------------------------------------------------------------- #
def mean_var(volume: torch.Tensor, valid: torch.Tensor):
    """mean over valid views; variance summed over ALL views (invalid ones are 0,
    SURVEY.md §0.3) divided by the valid count; unobserved voxels: mean 0, cov
    exp(-1e6) = 0.  Returns (mean [C,X,Y,Z], cov [C,X,Y,Z], count [1,X,Y,Z] int64)."""
    count = valid.sum(dim=0)
    denom = count + 1e-8
    mean = volume.sum(dim=0) / denom
    empty = count[0] == 0
    mean[:, empty] = 0.0
    cov = ((volume - mean.unsqueeze(0)) ** 2).sum(dim=0) / denom
    cov[:, empty] = 1e6
    return mean, torch.exp(-cov), count


def lift_mean_var(features, points, projection):
    """B3+B5+B6 chained exactly as ``extract_feat`` does (nerfdet.py:164-181)."""
    volume, valid = backproject(features, points, projection)
    return mean_var(volume, valid)


# --------------------------------------------------------------------------- #
# B7-B9: live 35-channel statistics (nerfdet.py:190-253)
# --------------------------------------------------------------------------- #
def map_features_2d(features: torch.Tensor, w_map: torch.Tensor, b_map: torch.Tensor):
    """Per-pixel Linear(C -> 32) of the sliced feature maps (nerfdet.py:194-197)."""
    nv, c, h, w = features.shape
    flat = features.reshape(nv, c, h * w).permute(0, 2, 1).contiguous()
    out = F.linear(flat, w_map, b_map)
    return out.permute(0, 2, 1).contiguous().view(nv, -1, h, w)


def live_stats(volume, valid, rgb_volume, w_map, b_map):
    """35-channel mean / exp(-var) fed to the density MLP (nerfdet.py:234-253).

    volume [nv,C,X,Y,Z], valid [nv,1,X,Y,Z], rgb_volume [nv,3,X,Y,Z].  Invalid
    views enter as Linear(0) = bias.  Returns
    (global_volume [N, 70] in the reference's *interleaved* order, SURVEY.md §0.10,
    mean35 [35,X,Y,Z], cov35 [35,X,Y,Z])."""
    nv, c, gx, gy, gz = volume.shape
    count = valid.sum(dim=0)
    denom = count + 1e-8
    flat = volume.reshape(nv, c, -1).permute(0, 2, 1).contiguous()
    mapped = F.linear(flat, w_map, b_map).permute(0, 2, 1).contiguous().view(nv, -1, gx, gy, gz)
    allch = torch.cat([rgb_volume, mapped], dim=1)
    mean = allch.sum(dim=0) / denom
    cov = ((allch - mean.unsqueeze(0)) ** 2).sum(dim=0) / denom
    cov[:, count[0] == 0] = 1e6
    cov = torch.exp(-cov)
    # The reference concatenates two 4-D [35,X,Y,Z] tensors along dim=1 (= X) and
    # then views the result as [-1, X*Y*Z]; the rows come out channel-interleaved.
    glob = torch.cat([mean, cov], dim=1).view(-1, gx * gy * gz).permute(1, 0).contiguous()
    return glob, mean, cov


def density_volume(points, global_volume, volume_mean, count, field):
    """B10 (nerfdet.py:254-261): alpha = 1 - exp(-relu(sigma)); out = alpha * mean,
    0 where unobserved.  ``field`` is any object with ``query_density(x, feats)``."""
    gx, gy, gz = volume_mean.shape[-3:]
    pts = points.reshape(3, -1).permute(1, 0).contiguous()
    sigma = field.query_density(pts, global_volume)
    alpha = 1 - torch.exp(-sigma)
    out = alpha.view(1, gx, gy, gz) * volume_mean
    out[:, count[0] == 0] = 0.0
    return out, alpha


def extract_lift(features_sliced, img_meta, n_voxels, voxel_size, denorm_images=None,
                 w_map=None, b_map=None, field=None, stride: int = 4, depth=None):
    """The whole voxel side of ``extract_feat`` for one scene (nerfdet.py:152-261,
    minus ``render_rays``).  Returns a dict of every intermediate the parity tests
    compare.  ``depth [nv, Hp, Wp]``: the depth prior ``extract_feat`` hands to both
    ``backproject`` calls (nerfdet.py:164-169, 204-210)."""
    proj = compute_projection(img_meta, stride)
    pts = get_points(n_voxels, voxel_size, img_meta['lidar2img']['origin'])
    volume, valid = backproject(features_sliced, pts, proj, depth, voxel_size)
    mean, cov, count = mean_var(volume, valid)
    out = dict(projection=proj, points=pts, volume_mean=mean, volume_cov=cov, count=count)
    if denorm_images is None:
        return out
    h, w = img_meta['img_shape'][:2]
    imgs = denorm_images.reshape([-1] + list(denorm_images.shape)[2:])
    rgb_proj = compute_projection(img_meta, 1)
    rgb_volume, _ = backproject(imgs[:, :, :h, :w], pts, rgb_proj, depth, voxel_size)
    glob, mean35, cov35 = live_stats(volume, valid, rgb_volume, w_map, b_map)
    out.update(rgb_projection=rgb_proj, global_volume=glob, mean35=mean35, cov35=cov35,
               feature_2d=map_features_2d(features_sliced, w_map, b_map))
    if field is not None:
        x_scene, alpha = density_volume(pts, glob, mean, count, field)
        out.update(x_scene=x_scene, alpha=alpha)
    return out
