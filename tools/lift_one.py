#!/usr/bin/env python
"""A few launches of the plan-based fused lift at the bench shape (the target of `ncu -k regex:k_lift_quads`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfdet_b200 import lifting, ops  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_features, make_scene  # noqa: E402

kw = {}
for a in sys.argv[1:]:
    k, v = a.split('=')
    kw[k] = int(v)
nv, c, grid = kw.pop('nv', 50), kw.pop('c', 256), (40, 40, 16)
cfg = SceneConfig(n_views=nv, n_voxels=grid, voxel_size=(0.16, 0.16, 0.2), channels=4)
sc = make_scene(cfg, seed=1000, with_images=False, with_features=False)
proj = lifting.compute_projection(sc.img_meta, 4).cuda()
pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin']).cuda()
sets = [torch.from_numpy(make_features(np.random.RandomState(2000 + i), (nv, c, 60, 80))).cuda() for i in range(3)]
plan = ops.LiftPlan(sets[0][:, :, :59, :80], pts, proj, **kw)
for i in range(6):
    out = plan.mean_var(sets[i % 3][:, :, :59, :80])
    torch.cuda.synchronize()
print('ok', float(out[0].sum()))
