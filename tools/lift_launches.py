#!/usr/bin/env python
"""One-shot lifts (geometry plan + lift per call) at two voxel grids: the target of
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list (k_q_index, k_q_pack, k_lift_quads)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfdet_b200 import lifting, ops  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_features, make_scene  # noqa: E402

nv, c = 50, 256
feats = torch.from_numpy(make_features(np.random.RandomState(1), (nv, c, 60, 80))).cuda()
for grid, vs in (((40, 40, 16), (.16, .16, .2)), ((80, 80, 32), (.08, .08, .08))):
    cfg = SceneConfig(n_views=nv, n_voxels=grid, voxel_size=vs, channels=4)
    sc = make_scene(cfg, seed=1000, with_images=False, with_features=False)
    proj = lifting.compute_projection(sc.img_meta, 4).cuda()
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin']).cuda()
    for i in range(4):
        out = ops.lift_mean_var(feats[:, :, :59, :80], pts, proj, None, True, 0)
        torch.cuda.synchronize()
print('ok', float(out[0].sum()))
