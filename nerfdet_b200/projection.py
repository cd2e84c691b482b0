"""``Projector`` with the reference interface (``mmdet3d/models/model_utils/projection.py:20-151``).

``compute`` returns the materialised ``[rays, samples, views, 3 + D]`` samples like the reference (the
compatibility path, used by tests); the renderer itself calls ``ops.render_gather_stats`` which reduces over
views on the fly and never writes that tensor."""
from __future__ import annotations

import torch

from . import ops


class Projector:
    def __init__(self, device='cuda'):
        self.device = device

    @staticmethod
    def inbound(pixel_locations, h, w):
        """projection.py:24-33"""
        return (pixel_locations[..., 0] <= w - 1.) & (pixel_locations[..., 0] >= 0) & \
               (pixel_locations[..., 1] <= h - 1.) & (pixel_locations[..., 1] >= 0)

    @staticmethod
    def normalize(pixel_locations, h, w):
        """projection.py:35-40"""
        resize_factor = torch.tensor([w - 1., h - 1.]).to(pixel_locations.device)[None, None, :]
        return 2 * pixel_locations / resize_factor - 1.

    def _run(self, xyz, train_imgs, train_cameras, featmaps, want_pixels, want_features):
        if train_cameras.dim() == 3:
            assert train_cameras.shape[0] == 1, 'only support batch_size=1 for now'     # projection.py:100-101
            train_cameras = train_cameras[0]
        if train_imgs.dim() == 5:
            assert train_imgs.shape[0] == 1, 'only support batch_size=1 for now'
            train_imgs = train_imgs[0].permute(0, 3, 1, 2)                               # projection.py:103
        lead = xyz.shape[:-1]
        if featmaps is None:
            featmaps = train_imgs.new_zeros((train_imgs.shape[0], 0) + tuple(train_imgs.shape[2:]))
        glob, vmask, pmask, pix, front, vf = ops.direct.render_gather_stats(xyz.reshape(-1, 3), train_cameras, train_imgs,
                                                                     featmaps, want_pixels, want_features)
        return lead, glob, vmask, pmask, pix, front, vf

    def compute_projections(self, xyz, train_cameras):
        """pixel_locations [nv, rays, samples, 2] (clamped to +-1e6) and the in-front mask (projection.py:42-64)."""
        cams = train_cameras if train_cameras.dim() == 2 else train_cameras[0]
        nv = cams.shape[0]
        dummy = xyz.new_zeros((nv, 3, 2, 2))
        lead, _, _, _, pix, front, _ = self._run(xyz, dummy, cams, None, True, False)
        return pix.view(nv, *lead, 2), front.view(nv, *lead)

    def compute(self, xyz, train_imgs, train_cameras, featmaps=None, grid_sample=True):
        """projection.py:91-151: ``rgb_feat [rays, samples, views, 3 + D]`` (None when ``featmaps`` is None) and
        ``mask [rays, samples, views, 1]`` float (in-bounds and in-front)."""
        if not grid_sample:
            raise NotImplementedError('grid_sample=False (projection.py:131-144) is never taken by NeRF-Det')
        lead, _, vmask, _, _, _, vf = self._run(xyz, train_imgs, train_cameras, featmaps, False, featmaps is not None)
        nv = vmask.shape[1]
        mask = vmask.view(*lead, nv, 1).float()
        rgb_feat = vf.view(*lead, nv, -1) if featmaps is not None else None
        return rgb_feat, mask
