"""The live voxel side of ``nerfdet.extract_feat`` (reference ``nerfdet.py:152-261``, the
``nerf_mode='image'``, ``nerf_density=True`` branch -- the only one that runs):

    B7   map_features_2d      Linear(C -> 32) per pixel of the sliced feature maps          nerfdet.py:190-197
    B8+9 live_statistics      RGB + mapped-feature gathers, 35-ch mean / exp(-var)         nerfdet.py:200-210, 232-253
    B10  density_volume       alpha = 1 - exp(-relu(sigma));  x = alpha * mean             nerfdet.py:254-261

``lift_scene`` chains them with the fused 256-channel lift (lifting.lift_mean_var) and returns what
``extract_feat`` hands to ``neck_3d`` (the ``[C, X, Y, Z]`` volume) and to the head (``valids``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import lifting, ops


def map_features_2d(feature_2d: torch.Tensor, mapping) -> torch.Tensor:
    """``self.mapping`` applied per pixel (nerfdet.py:190-197): ``[nv, C, h, w] -> [nv, 32, h, w]``.
    ``mapping`` is the reference's ``nn.Sequential(nn.Linear(256, 32))`` (nerfdet.py:103-105) or any
    callable on ``[nv, h*w, C]``.

    For the reference's ``Linear(C, 32)`` this is ``nd_map_features`` (csrc/mapping.cu): one fp32 FFMA kernel that
    reads the NCHW planes in place -- the reference's 242 MB ``permute(0, 2, 1).contiguous()`` copy is not made -- and
    writes the channels-last layout the gather kernels want: the returned tensor has the reference's logical shape
    ``[nv, 32, h, w]`` with channels-last strides.  Other mappings go through torch."""
    nv, c, h, w = feature_2d.shape
    lin = mapping[0] if isinstance(mapping, torch.nn.Sequential) and len(mapping) == 1 else mapping
    elt = feature_2d.element_size()
    if (isinstance(lin, torch.nn.Linear) and lin.out_features == 32 and c % 16 == 0 and feature_2d.is_cuda
            and (h * w * elt) % 16 == 0):
        sv, sc, sy, sx = feature_2d.stride()
        if not (sx == 1 and sy == w and (sc * elt) % 16 == 0 and (sv * elt) % 16 == 0 and feature_2d.data_ptr() % 16 == 0):
            feature_2d = feature_2d.contiguous()
        if torch.is_grad_enabled() and (feature_2d.requires_grad or lin.weight.requires_grad):
            # training: through the registered op, whose autograd formula is two cuBLAS GEMMs
            y = ops.map_features(feature_2d, lin.weight, lin.bias)
        else:
            y = ops.direct.map_features(feature_2d, lin.weight.detach(), lin.bias.detach() if lin.bias is not None else None)
        return y.permute(0, 3, 1, 2)
    flat = feature_2d.reshape(nv, c, h * w).permute(0, 2, 1).contiguous().float()
    return mapping(flat).view(nv, h, w, -1).permute(0, 3, 1, 2)


def _mapping_bias(mapping) -> torch.Tensor:
    """The bias the invalid views enter the statistics with; attached to the graph when it trains."""
    lin = mapping[0] if isinstance(mapping, torch.nn.Sequential) else mapping
    return lin.bias if (torch.is_grad_enabled() and lin.bias.requires_grad) else lin.bias.detach()


class _LiveStatsFn(torch.autograd.Function):
    """``nd_live_stats`` with gradients for the mapped features and the mapping bias (row N1, ``nd_live_stats_bwd``)."""

    @staticmethod
    def forward(ctx, mapped_2d, rgb_images, points, projection, rgb_projection, map_bias, depth_mapped, depth_rgb, voxel_z):
        glob, _, _, count = ops.direct.live_stats(mapped_2d.detach(), rgb_images, points, projection, rgb_projection,
                                                  map_bias.detach(), False, depth_mapped, depth_rgb, voxel_z)
        ctx.save_for_backward(mapped_2d, points, projection, map_bias, glob)
        ctx.depth_mapped, ctx.voxel_z = depth_mapped, voxel_z
        ctx.mark_non_differentiable(count)
        return glob, count

    @staticmethod
    def backward(ctx, g_glob, _g_count):
        mapped_2d, points, projection, map_bias, glob = ctx.saved_tensors
        g_mapped, g_bias = ops.direct.live_stats_bwd(mapped_2d.detach(), points, projection, map_bias.detach(), glob, g_glob,
                                                     ctx.depth_mapped, ctx.voxel_z)
        return g_mapped, None, None, None, None, g_bias, None, None, None


def live_statistics(mapped_2d, rgb_images, points, projection, rgb_projection, map_bias, want_planes=False,
                    depth: Optional[torch.Tensor] = None, voxel_size=None):
    """B8 + B9.  ``mapped_2d [nv, 32, h, w]`` (B7 output), ``rgb_images [nv, 3, H, W]`` = the
    ``denorm_images[:, :, :img_h, :img_w]`` slice, projections at feature level (stride 4) and image level
    (stride 1).  Returns ``global_volume [N, 70]`` (interleaved rows), the feature-level
    count ``[1, X, Y, Z]`` and, on request, ``mean35`` / ``cov35 [35, X, Y, Z]``.

    ``depth [nv, Hp, Wp]`` with ``voxel_size`` gates both gathers like ``backproject`` (nerfdet.py:405-411) does when
    ``extract_feat`` is called with a depth prior: resized bilinearly to each map's resolution (``F.interpolate``, like
    the reference), a voxel-view is kept only if ``|z - depth[y, x]| < voxel_size[-1]``."""
    gx, gy, gz = points.shape[-3:]
    depth_mapped = depth_rgb = None
    voxel_z = 0.0
    if depth is not None:
        if voxel_size is None:
            raise ValueError('depth needs voxel_size (the gate is |z - depth| < voxel_size[-1])')
        with torch.no_grad():
            d = depth.float().unsqueeze(1)
            depth_mapped = F.interpolate(d, size=tuple(mapped_2d.shape[-2:]), mode='bilinear').squeeze(1)
            depth_rgb = F.interpolate(d, size=tuple(rgb_images.shape[-2:]), mode='bilinear').squeeze(1)
        voxel_z = float(voxel_size[-1])
    if torch.is_grad_enabled() and (mapped_2d.requires_grad or map_bias.requires_grad):
        if want_planes:
            raise NotImplementedError('mean35 / cov35 planes are forward-only outputs')
        glob, count = _LiveStatsFn.apply(mapped_2d, rgb_images, points, projection, rgb_projection, map_bias,
                                         depth_mapped, depth_rgb, voxel_z)
        mean35 = cov35 = None
    else:
        glob, mean35, cov35, count = ops.direct.live_stats(mapped_2d, rgb_images, points, projection, rgb_projection, map_bias,
                                                           want_planes, depth_mapped, depth_rgb, voxel_z)
    out = dict(global_volume=glob, count=count.view(1, gx, gy, gz))
    if want_planes:
        out['mean35'] = mean35.view(-1, gx, gy, gz)
        out['cov35'] = cov35.view(-1, gx, gy, gz)
    return out


def lift_scene(feature: torch.Tensor, img_meta: Dict, n_voxels, voxel_size, mapping=None, nerf_mlp=None,
               denorm_images: Optional[torch.Tensor] = None, stride: int = 4, want_cov: bool = False,
               depth: Optional[torch.Tensor] = None) -> Dict:
    """One scene of ``extract_feat``'s loop body (nerfdet.py:152-261) without ``render_rays``.

    ``feature [nv, C, Hf, Wf]`` is the un-sliced FPN output of the scene's views.  With ``mapping``,
    ``nerf_mlp`` and ``denorm_images [nv, 3, Hp, Wp]`` given, the returned ``volume`` is
    ``alpha * volume_mean`` (the density-weighted volume ``neck_3d`` receives); without them it is the plain
    ``volume_mean``.  Keys: volume [C,X,Y,Z], valid [1,X,Y,Z] int64, and, when available, volume_cov,
    feature_2d [nv,32,h,w], global_volume [N,70], alpha [N].

    ``depth [nv, Hp, Wp]`` (``extract_feat(depth=...)``) gates every gather of the scene like ``backproject`` does
    (nerfdet.py:164-169, 204-210, 405-411): the 256-channel lift in its geometry plan, the live 35-channel statistics
    (mapped features and RGB) in ``nd_live_stats_gated``."""
    dev = feature.device
    projection = lifting.to_device(lifting.compute_projection(img_meta, stride), dev)
    points = lifting.get_points_device(n_voxels, voxel_size, img_meta['lidar2img']['origin'], dev)
    height = img_meta['img_shape'][0] // stride
    width = img_meta['img_shape'][1] // stride
    sliced = feature[:, :, :height, :width]
    out = {}
    alpha = None
    if mapping is not None and nerf_mlp is not None and denorm_images is not None:
        feature_2d = map_features_2d(sliced, mapping)
        rgb_projection = lifting.to_device(lifting.compute_projection(img_meta, 1), dev)
        rgb = denorm_images[:, :, :img_meta['img_shape'][0], :img_meta['img_shape'][1]]
        live = live_statistics(feature_2d, rgb, points, projection, rgb_projection, _mapping_bias(mapping), depth=depth,
                               voxel_size=voxel_size if depth is not None else None)
        pts = points.view(3, -1).permute(1, 0).contiguous()
        _, alpha = nerf_mlp.query_density(pts, live['global_volume'], return_alpha=True)
        out.update(feature_2d=feature_2d, global_volume=live['global_volume'], alpha=alpha.view(-1),
                   rgb_projection=rgb_projection)
    mean, cov, valid = lifting.lift_mean_var(sliced, points, projection, alpha=alpha, want_cov=want_cov, depth=depth,
                                             voxel_size=voxel_size if depth is not None else None)
    out.update(volume=mean, valid=valid, points=points, projection=projection)
    if want_cov:
        out['volume_cov'] = cov
    return out
