"""TEST INFRASTRUCTURE -- CPU oracle for the NeRF rendering branch.  NOT product code.

Restates, in torch-CPU fp32, ``mmdet3d/models/model_utils/render_ray.py`` and
``mmdet3d/models/model_utils/projection.py`` of the reference (file:line cited
per function).  Pinned by fixtures generated from the unmodified reference
(``oracle/make_golden.py`` -> ``tests/golden/``); no upstream test covers this
path.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# R2  sample_along_camera_ray  (render_ray.py:145-189)
# --------------------------------------------------------------------------- #
def sample_along_rays(ray_o, ray_d, near: float, far: float, n_samples: int,
                      det: bool, t_rand: Optional[torch.Tensor] = None):
    """z_i = near + i*step (two roundings), optional stratified jitter with the
    supplied uniform numbers, pts = z*d + o (separate mul and add)."""
    assert near > 0 and far > near
    ones = torch.ones_like(ray_d[..., 0])
    lo = near * ones
    step = (far * ones - lo) / (n_samples - 1)
    z = torch.stack([lo + i * step for i in range(n_samples)], dim=1)
    if not det:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mid, z[:, -1:]], dim=-1)
        lower = torch.cat([z[:, :1], mid], dim=-1)
        if t_rand is None:
            t_rand = torch.rand_like(z)
        z = lower + (upper - lower) * t_rand
    pts = z.unsqueeze(2) * ray_d.unsqueeze(1) + ray_o.unsqueeze(1)
    return pts, z


# --------------------------------------------------------------------------- #
# R3  camera packing  (render_ray.py:48-69)
# --------------------------------------------------------------------------- #
def pack_cameras(img_meta) -> torch.Tensor:
    """[1, nv, 34] = [h, w, K4x4 (rows 0-1 / (ori_h/img_h)), E4x4]."""
    ext = img_meta['lidar2img']['extrinsic']
    nv = len(ext)
    k = torch.tensor(img_meta['lidar2img']['intrinsic'][:4, :4])
    k[:2] /= img_meta['ori_shape'][0] / img_meta['img_shape'][0]
    hw = torch.tensor([float(img_meta['img_shape'][0]), float(img_meta['img_shape'][1])])
    rows = [torch.cat([hw, k.reshape(16), torch.tensor(e, dtype=torch.float32).reshape(16)])
            for e in ext]
    return torch.stack(rows).view(1, nv, 34)


# --------------------------------------------------------------------------- #
# R4  per-sample projection  (projection.py:24-64)
# --------------------------------------------------------------------------- #
def project_samples(xyz: torch.Tensor, cameras: torch.Tensor):
    """xyz [R, S, 3], cameras [nv, 34] -> pixel [nv, R, S, 2] f32, in_front [nv, R, S] bool."""
    r, s = xyz.shape[:2]
    nv = cameras.shape[0]
    k = cameras[:, 2:18].reshape(nv, 4, 4)
    e = cameras[:, 18:34].reshape(nv, 4, 4)
    flat = xyz.reshape(-1, 3)
    homo = torch.cat([flat, torch.ones_like(flat[:, :1])], dim=-1)
    q = k.bmm(e).bmm(homo.t()[None].repeat(nv, 1, 1)).permute(0, 2, 1)
    pix = q[..., :2] / torch.clamp(q[..., 2:3], min=1e-8)
    pix = torch.clamp(pix, min=-1e6, max=1e6)
    return pix.reshape(nv, r, s, 2), (q[..., 2] > 0).reshape(nv, r, s)


def gather_views(xyz, images, cameras, featmaps):
    """R5 (projection.py:91-151, ``grid_sample=True`` branch).

    images [nv, 3, Hp, Wp] (NCHW, *padded* size), featmaps [nv, D, Hf, Wf].
    Both are bilinearly sampled (zeros padding, align_corners=True) at the same
    normalised coordinates 2*pix/[w-1, h-1]-1 with (h, w) = img_shape.  Returns
    rgb_feat [R, S, nv, 3+D] and mask [R, S, nv, 1] f32 (in-bounds & in-front)."""
    h, w = cameras[0][:2]
    pix, front = project_samples(xyz, cameras)
    scale = torch.tensor([w - 1.0, h - 1.0])[None, None, :]
    grid = 2 * pix / scale - 1.0
    rgb = F.grid_sample(images, grid, align_corners=True).permute(2, 3, 0, 1)
    feat = None
    if featmaps is not None:
        fs = F.grid_sample(featmaps, grid, align_corners=True).permute(2, 3, 0, 1)
        feat = torch.cat([rgb, fs], dim=-1)
    inb = (pix[..., 0] <= w - 1.0) & (pix[..., 0] >= 0) & (pix[..., 1] <= h - 1.0) & (pix[..., 1] >= 0)
    mask = (inb * front).float().permute(1, 2, 0)[..., None]
    return feat, mask


# --------------------------------------------------------------------------- #
# R6  masked mean / all-view variance  (render_ray.py:71-93)
# --------------------------------------------------------------------------- #
def view_statistics(feat: torch.Tensor, mask: torch.Tensor):
    """feat [R,S,nv,C], mask [R,S,nv,1] -> mean, exp(-var), each [R,S,1,C].
    The variance sums over every view (masked or not) and divides by the mask
    count."""
    cnt = mask.sum(dim=2, keepdim=True) + 1e-8
    mean = (feat * (mask / cnt)).sum(dim=2, keepdim=True)
    var = ((feat - mean) ** 2).sum(dim=2, keepdim=True) / cnt
    return mean, torch.exp(-var)


# --------------------------------------------------------------------------- #
# R7  alpha compositing  (render_ray.py:196-247)
# --------------------------------------------------------------------------- #
def composite(rgb, sigma, z_vals, pixel_mask, white_bkgd: bool = False):
    """rgb [R,S,3], sigma [R,S], z_vals [R,S], pixel_mask [R,S] bool."""
    alpha = 1.0 - torch.exp(-sigma)
    trans = torch.cumprod(1.0 - alpha + 1e-10, dim=-1)[:, :-1]
    trans = torch.cat((torch.ones_like(trans[:, :1]), trans), dim=-1)
    weights = alpha * trans
    rgb_map = (weights.unsqueeze(2) * rgb).sum(dim=1)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - weights.sum(dim=-1, keepdim=True))
    ray_mask = pixel_mask.float().sum(dim=1) > 8 if pixel_mask is not None else None
    depth = (weights * z_vals).sum(dim=-1) / (weights.sum(dim=-1) + 1e-8)
    depth = torch.clamp(depth, z_vals.min(), z_vals.max())
    return OrderedDict(rgb=rgb_map, depth=depth, weights=weights, mask=ray_mask,
                       alpha=alpha, z_vals=z_vals, transparency=trans)


# --------------------------------------------------------------------------- #
# R8  trilinear volume lookup  (render_ray.py:26-46)
# --------------------------------------------------------------------------- #
def volume_lookup(pts, volume, aabb):
    """pts [R,S,3]; volume [1,C,D0,D1,D2].  Normalised x indexes the LAST volume
    axis (grid_sample convention, applied literally by the reference)."""
    assert volume.shape[0] == 1
    c = volume.shape[1]
    r, s, _ = pts.shape
    box = torch.tensor(aabb, dtype=torch.float32)
    inv = 1.0 / (box[1] - box[0]) * 2
    norm = (pts.view(1, r * s, 1, 1, 3) - box[0]) * inv - 1
    out = F.grid_sample(volume, norm, align_corners=True, padding_mode='border')
    inside = ((norm < 1) & (norm > -1)).float().sum(dim=-1).view(r, s) == 3
    return out.view(c, r, s).permute(1, 2, 0).contiguous(), inside


# --------------------------------------------------------------------------- #
# render_rays_func, image mode  (render_ray.py:250-327)
# --------------------------------------------------------------------------- #
def render_image_mode(ray_o, ray_d, featmaps, images, near_far, n_samples, field, img_meta,
                      det: bool, t_rand=None, white_bkgd: bool = False) -> Dict:
    """images [nv, 3, Hp, Wp]; featmaps [nv, D, Hf, Wf]; field(x, cond, feats)."""
    pts, z = sample_along_rays(ray_o, ray_d, near_far[0], near_far[1], n_samples, det, t_rand)
    cams = pack_cameras(img_meta)[0]
    feat, mask = gather_views(pts, images, cams, featmaps)
    pixel_mask = mask[..., 0].sum(dim=2) > 1
    mean, var = view_statistics(feat, mask)
    glob = torch.cat([mean, var], dim=-1).squeeze(2)
    rgb, sigma = field(pts, ray_d, glob)
    out = composite(rgb, sigma[..., 0], z, pixel_mask, white_bkgd)
    return dict(outputs_coarse=out, sigma=sigma, globalfeat=glob, pts=pts, pixel_mask=pixel_mask,
                rgb_pts=rgb)


def select_training_rays(ray_batch: Dict, n_rand: int, rng: np.random.RandomState):
    """R1 train branch (render_ray.py:408-427): flatten, drop gt_depth <= 0, draw
    ``n_rand`` rays without replacement from the supplied host RNG."""
    ray_o = ray_batch['ray_o'].view(-1, 3)
    ray_d = ray_batch['ray_d'].view(-1, 3)
    gt_rgb = ray_batch['gt_rgb'].view(-1, 3)
    gt_depth = ray_batch['gt_depth']
    if len(gt_depth) != 0:
        gt_depth = gt_depth.view(-1, 1)
        keep = (gt_depth > 0).squeeze(-1)
        ray_o, ray_d, gt_rgb, gt_depth = ray_o[keep], ray_d[keep], gt_rgb[keep], gt_depth[keep]
    else:
        gt_depth = None
    sel = rng.choice(ray_d.shape[0], size=(n_rand,), replace=False)
    return (ray_o[sel], ray_d[sel], gt_rgb[sel],
            gt_depth[sel] if gt_depth is not None else None, sel)
