#!/usr/bin/env python
"""One launch of the render gather kernel at the bench shape (ncu target)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerfdet_b200 import ops, render  # noqa: E402

dev = torch.device('cuda', 0)
cfg, sc, state = bench.render_scene()
imgs = sc.denorm_images[0].to(dev)
f2d = sc.features[:, :, :bench.FEAT_HW[0], :bench.FEAT_HW[1]].contiguous().to(dev).contiguous(memory_format=torch.channels_last)
rb = sc.ray_batch
sel = np.random.RandomState(5).choice(rb['ray_o'].view(-1, 3).shape[0], 2048, replace=False)
ro, rd = rb['ray_o'].view(-1, 3)[sel].float().to(dev), rb['ray_d'].view(-1, 3)[sel].float().to(dev)
pts, z = render.sample_along_camera_ray(ro, rd, cfg.near_far_range, 64, det=True)
cams = render._compute_projection(sc.img_meta)[0].to(dev)
for _ in range(3):
    out = ops.direct.render_gather_stats(pts.view(-1, 3), cams, imgs, f2d, False, False)
torch.cuda.synchronize()
vm = ops.direct.render_gather_stats(pts.view(-1, 3), cams, imgs, f2d, True, False)[1]
print('views seeing a sample (mean):', float(vm.float().sum(-1).mean()))
