"""The NeRF rendering branch with the reference call signatures
(``mmdet3d/models/model_utils/render_ray.py``): ``render_rays`` / ``render_rays_func`` and their helpers.

Host code keeps what the reference keeps on the host -- ray selection with a numpy ``RandomState(234)``
(render_ray.py:20, 422), ``torch.rand_like`` for the stratified jitter (so both random streams match the
reference's), camera packing (render_ray.py:48-69); everything per ray sample runs in the CUDA kernels of
``csrc/render.cu`` and ``csrc/mlp.cu``.  The reference's ``[rays, samples, views, 35]`` tensor is not built."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import ops
from .projection import Projector  # noqa: F401  (re-exported like the reference module layout)

rng = np.random.RandomState(234)          # render_ray.py:20


def _compute_projection(img_meta) -> torch.Tensor:
    """[1, n_views, 34] = [h, w, K 4x4 (rows 0-1 / (ori_h / img_h)), E 4x4] on the CPU (render_ray.py:48-69).
    One numpy block instead of the reference's per-view tensor constructors (240 us of host time at 50 views);
    the values are the same float32 numbers."""
    l2i = img_meta['lidar2img']
    extrinsic = l2i['extrinsic']
    views = len(extrinsic)
    intrinsic = l2i['intrinsic']
    intrinsic = (intrinsic.detach().cpu().numpy() if isinstance(intrinsic, torch.Tensor) else np.asarray(intrinsic))[:4, :4]
    intrinsic = intrinsic.astype(np.float32)                                 # a copy: the caller's matrix is not touched
    intrinsic[:2] /= np.float32(img_meta['ori_shape'][0] / img_meta['img_shape'][0])
    out = np.empty((1, views, 34), dtype=np.float32)
    out[0, :, 0] = img_meta['img_shape'][0]
    out[0, :, 1] = img_meta['img_shape'][1]
    out[0, :, 2:18] = intrinsic.reshape(16)
    if views and isinstance(extrinsic[0], torch.Tensor):
        out[0, :, 18:] = torch.stack(list(extrinsic)).detach().cpu().numpy().reshape(views, 16)
    else:
        out[0, :, 18:] = np.asarray(extrinsic, dtype=np.float32).reshape(views, 16)
    return torch.from_numpy(out)


_CAMERA_CACHE: "OrderedDict[tuple, torch.Tensor]" = OrderedDict()


def _device_cameras(img_meta, device) -> torch.Tensor:
    """The packed cameras ``[n_views, 34]`` on ``device``.  A pageable host-to-device copy blocks the host until the
    stream has drained, which serialises every call behind the previous one; the upload goes through pinned memory
    and is skipped when these exact bytes are already on the device (same scene as one of the last calls)."""
    cams = _compute_projection(img_meta)[0]
    key = (cams.numpy().tobytes(), str(device), torch.cuda.current_stream(device).cuda_stream)
    hit = _CAMERA_CACHE.get(key)
    if hit is not None:
        _CAMERA_CACHE.move_to_end(key)
        return hit
    dev_cams = cams.pin_memory().to(device, non_blocking=True)
    _CAMERA_CACHE[key] = dev_cams
    while len(_CAMERA_CACHE) > 16:
        _CAMERA_CACHE.popitem(last=False)
    return dev_cams


def sample_along_camera_ray(ray_o, ray_d, depth_range, N_samples, inv_uniform=False, det=False):
    """pts [rays, samples, 3], z_vals [rays, samples] (render_ray.py:145-189)."""
    if inv_uniform:
        raise NotImplementedError('inv_uniform=True is never used by NeRF-Det')
    t_rand = None
    if not det:
        t_rand = torch.rand((ray_o.shape[0], N_samples), dtype=torch.float32, device=ray_o.device)   # == rand_like(z_vals)
    return ops.direct.sample_rays(ray_o, ray_d, float(depth_range[0]), float(depth_range[1]), int(N_samples), t_rand)


def volume_sampling(sample_pts, features, aabb):
    """Trilinear lookup of ``features [1, C, D0, D1, D2]`` at ``sample_pts [rays, samples, 3]``
    (render_ray.py:26-46): ``[rays, samples, C]`` and the strict in-box mask."""
    assert features.shape[0] == 1
    r, s = sample_pts.shape[:2]
    out, inside = ops.direct.volume_sample(features[0], sample_pts.reshape(-1, 3), [float(v) for v in aabb[0]],
                                    [float(v) for v in aabb[1]])
    return out.view(r, s, -1), inside.view(r, s)


def raw2outputs(raw, z_vals, mask, white_bkgd=False):
    """Alpha compositing (render_ray.py:196-247) of ``raw [rays, samples, 4]`` = (rgb, sigma)."""
    return _composite(raw[..., :3].contiguous(), raw[..., 3].contiguous(), z_vals, mask, white_bkgd)


def _composite(rgb_pts, sigma_pts, z_vals, mask, white_bkgd=False, det=False):
    """``raw2outputs`` on separate colour / density tensors (no cat + slice round trip).  The depth clamp bounds are
    batch-global (render_ray.py:236) and stay on the device; with deterministic sampling every ray has the same
    samples, so the first ray's end points are the bounds."""
    if det and z_vals.shape[0] > 0:
        bounds = torch.stack([z_vals[0, 0], z_vals[0, -1]])
    else:
        bounds = torch.stack(torch.aminmax(z_vals))
    if torch.is_grad_enabled() and (rgb_pts.requires_grad or sigma_pts.requires_grad):
        rgb, depth, weights, alpha, trans, ray_mask = _CompositeFn.apply(rgb_pts.contiguous(), sigma_pts.contiguous(), z_vals, mask,
                                                                         bounds, bool(white_bkgd))
    else:
        rgb, depth, weights, alpha, trans, ray_mask = ops.direct.composite(rgb_pts, sigma_pts, z_vals, mask, bounds, bool(white_bkgd))
    return OrderedDict([('rgb', rgb), ('depth', depth), ('weights', weights),
                        ('mask', ray_mask if mask is not None else None), ('alpha', alpha), ('z_vals', z_vals),
                        ('transparency', trans)])


class _GatherStatsFn(torch.autograd.Function):
    """``render_gather_stats`` with a gradient for the mapped feature maps (row N1, ``nd_render_gather_stats_bwd``)."""

    @staticmethod
    def forward(ctx, flat, cameras, img, features_2D):
        glob, _, pixel_mask, _, _, _ = ops.direct.render_gather_stats(flat, cameras, img, features_2D.detach(), False, False, False)
        ctx.save_for_backward(flat, cameras, features_2D, glob)
        ctx.img_hw = tuple(img.shape[-2:])
        ctx.mark_non_differentiable(pixel_mask)
        return glob, pixel_mask

    @staticmethod
    def backward(ctx, g_glob, _g_mask):
        flat, cameras, features_2D, glob = ctx.saved_tensors
        grad = ops.direct.render_gather_stats_bwd(flat, cameras, ctx.img_hw[0], ctx.img_hw[1], features_2D.detach(), glob, g_glob)
        return None, None, None, grad


class _CompositeFn(torch.autograd.Function):
    """``nd_composite`` forward; the backward differentiates the same formula (render_ray.py:196-247) written in torch ops
    on the saved inputs -- a few element-wise kernels on ``[rays, samples]`` tensors."""

    @staticmethod
    def forward(ctx, rgb_pts, sigma_pts, z_vals, mask, bounds, white_bkgd):
        out = ops.direct.composite(rgb_pts, sigma_pts, z_vals, mask, bounds, white_bkgd)
        ctx.save_for_backward(rgb_pts, sigma_pts, z_vals, bounds)
        ctx.white_bkgd = white_bkgd
        ctx.mark_non_differentiable(out[5])
        return out

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_weights, g_alpha, g_trans, _g_mask):
        rgb_pts, sigma_pts, z_vals, bounds = ctx.saved_tensors
        with torch.enable_grad():
            rgb = rgb_pts.detach().requires_grad_(True)
            sigma = sigma_pts.detach().requires_grad_(True)
            alpha = 1. - torch.exp(-sigma)
            trans = torch.cumprod(1. - alpha + 1e-10, dim=-1)[:, :-1]
            trans = torch.cat((torch.ones_like(trans[:, 0:1]), trans), dim=-1)
            weights = alpha * trans
            rgb_map = torch.sum(weights.unsqueeze(2) * rgb, dim=1)
            if ctx.white_bkgd:
                rgb_map = rgb_map + (1. - torch.sum(weights, dim=-1, keepdim=True))
            depth = torch.sum(weights * z_vals, dim=-1) / (torch.sum(weights, dim=-1) + 1e-8)
            depth = torch.maximum(torch.minimum(depth, bounds[1]), bounds[0])
            outs = [rgb_map, depth, weights, alpha, trans]
            gouts = [g_rgb, g_depth, g_weights, g_alpha, g_trans]
            pairs = [(o, g) for o, g in zip(outs, gouts) if g is not None]
            g_in = torch.autograd.grad([o for o, _ in pairs], [rgb, sigma], [g for _, g in pairs], allow_unused=True)
        return g_in[0], g_in[1], None, None, None, None


def render_rays_func(ray_o, ray_d, mean_volume, cov_volume, features_2D, img, aabb, near_far_range, N_samples,
                     N_rand=4096, nerf_mlp=None, img_meta=None, projector=None, mode='volume', nerf_sample_view=3,
                     inv_uniform=False, N_importance=0, det=False, is_train=True, white_bkgd=False, gt_rgb=None,
                     gt_depth=None):
    """render_ray.py:250-367 (coarse pass; the ``N_importance > 0`` fine pass references undefined names in the
    reference and is not built)."""
    if N_importance > 0:
        raise NotImplementedError('the fine pass (N_importance > 0) is dead code in the reference')
    ret = {'outputs_coarse': None, 'outputs_fine': None, 'gt_rgb': gt_rgb, 'gt_depth': gt_depth}
    pts, z_vals = sample_along_camera_ray(ray_o, ray_d, near_far_range, N_samples, inv_uniform, det)
    n_rays, n_samples = pts.shape[:2]
    cameras = _device_cameras(img_meta, pts.device)
    flat = pts.view(-1, 3)
    if mode == 'image':
        if torch.is_grad_enabled() and features_2D.requires_grad:
            glob, pixel_mask = _GatherStatsFn.apply(flat, cameras, img, features_2D)
        else:
            glob, _, pixel_mask, _, _, _ = ops.direct.render_gather_stats(flat, cameras, img, features_2D, False, False, False)
        rgb_pts, density_pts = nerf_mlp(pts, ray_d, glob.view(n_rays, n_samples, -1))
        ret['sigma'] = density_pts
    elif mode == 'volume':
        mean_pts, inbound = volume_sampling(pts, mean_volume, aabb)
        cov_pts, inbound = volume_sampling(pts, cov_volume, aabb)
        empty = img.new_zeros((img.shape[0], 0) + tuple(img.shape[2:]))
        _, _, pixel_mask, _, _, _ = ops.direct.render_gather_stats(flat, cameras, img, empty, False, False, False)
        rgb_pts, density_pts = nerf_mlp(pts, ray_d, torch.cat([mean_pts, cov_pts], dim=-1))
        density_pts = density_pts * inbound.unsqueeze(dim=-1)
    else:
        raise ValueError(f'unknown mode {mode!r}')
    ret['outputs_coarse'] = _composite(rgb_pts, density_pts.reshape(n_rays, n_samples), z_vals,
                                       pixel_mask.view(n_rays, n_samples), white_bkgd, det=det)
    return ret


def render_rays(ray_batch, mean_volume, cov_volume, features_2D, img, aabb, near_far_range, N_samples, N_rand=4096,
                nerf_mlp=None, img_meta=None, projector=None, mode='volume', nerf_sample_view=3, inv_uniform=False,
                N_importance=0, det=False, is_train=True, white_bkgd=False, render_testing=False):
    """render_ray.py:371-520.  ``img [nv, 3, Hp, Wp]`` are the de-normalised source images, ``features_2D
    [nv, D, h, w]`` the mapped feature maps; tensors of ``ray_batch`` may live on the host or the device."""
    dev = img.device
    ray_o = ray_batch['ray_o']
    ray_d = ray_batch['ray_d']
    gt_rgb = ray_batch['gt_rgb']
    gt_depth = ray_batch['gt_depth']
    nerf_sizes = ray_batch['nerf_sizes']

    def common(extra_det):
        return dict(nerf_mlp=nerf_mlp, img_meta=img_meta, projector=projector, mode=mode,
                    nerf_sample_view=nerf_sample_view, inv_uniform=inv_uniform, N_importance=N_importance,
                    det=extra_det, is_train=is_train, white_bkgd=white_bkgd)

    if is_train:
        ray_o = ray_o.view(-1, 3)
        ray_d = ray_d.view(-1, 3)
        gt_rgb = gt_rgb.view(-1, 3)
        if len(gt_depth) != 0:
            gt_depth = gt_depth.view(-1, 1)
            keep = (gt_depth > 0).squeeze(-1)
            ray_o, ray_d, gt_rgb, gt_depth = ray_o[keep], ray_d[keep], gt_rgb[keep], gt_depth[keep]
        else:
            gt_depth = None
        total_rays = ray_d.shape[0]
        select_inds = rng.choice(total_rays, size=(N_rand,), replace=False)
        ray_o, ray_d, gt_rgb = ray_o[select_inds], ray_d[select_inds], gt_rgb[select_inds]
        if gt_depth is not None:
            gt_depth = gt_depth[select_inds]
        return render_rays_func(ray_o.to(dev).float(), ray_d.to(dev).float(), mean_volume, cov_volume, features_2D, img,
                                aabb, near_far_range, N_samples, N_rand, gt_rgb=gt_rgb.to(dev),
                                gt_depth=gt_depth.to(dev) if gt_depth is not None else None, **common(det))
    if render_testing:
        nerf_size = nerf_sizes[0]
        view_num = ray_o.shape[1]
        H, W = int(nerf_size[0][0]), int(nerf_size[0][1])
        ray_o = ray_o.view(-1, 3).to(dev).float()
        ray_d = ray_d.view(-1, 3).to(dev).float()
        gt_rgb = gt_rgb.view(-1, 3)
        gt_depth = gt_depth.view(-1, 1) if len(gt_depth) != 0 else None
        assert view_num * H * W == ray_o.shape[0]
        rgbs, depths = [], []
        # the reference walks the image in chunks of N_rand rays (render_ray.py:480-500); with deterministic sampling every ray
        # is independent of its batch (same z_vals, hence the same depth clamp bounds), so larger chunks give the same pixels
        # with 32x fewer launch sequences at N_rand = 2048
        chunk = max(int(N_rand), 1 << 16)
        for i in range(0, ray_o.shape[0], chunk):
            ret = render_rays_func(ray_o[i:i + chunk], ray_d[i:i + chunk], mean_volume, cov_volume, features_2D, img,
                                   aabb, near_far_range, N_samples, N_rand, gt_rgb=gt_rgb, gt_depth=gt_depth,
                                   **common(True))
            rgbs.append(ret['outputs_coarse']['rgb'])
            depths.append(ret['outputs_coarse']['depth'])
        return {'outputs_coarse': {'rgb': torch.cat(rgbs, dim=0).view(view_num, H, W, 3),
                                   'depth': torch.cat(depths, dim=0).view(view_num, H, W, 1)},
                'gt_rgb': gt_rgb.view(view_num, H, W, 3),
                'gt_depth': gt_depth.view(view_num, H, W, 1) if gt_depth is not None else None}
    return None
