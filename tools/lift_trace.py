#!/usr/bin/env python
"""Per-warp, per-stage clock trace of CTA 0 of k_lift_planes (development tool).
usage: lift_trace.py <stages> <group>"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['ND_LIFT_STAGES'], os.environ['ND_LIFT_GROUP'] = sys.argv[1], sys.argv[2]
from nerfdet_b200 import _lib, lifting  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_scene  # noqa: E402

dev = torch.device('cuda', 0)
nv, ch = 50, 256
cfg = SceneConfig(n_views=nv, n_voxels=(40, 40, 16), voxel_size=(0.16, 0.16, 0.2), channels=ch)
sc = make_scene(cfg, seed=1000, with_images=False, with_features=False)
proj = lifting.compute_projection(sc.img_meta, 4).to(dev)
pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin']).to(dev)
feats = [torch.randn(nv, ch, 60, 80, device=dev) for _ in range(2)]
for i in range(3):
    lifting.lift_mean_var(feats[i % 2][:, :, :59, :80], pts, proj)
torch.cuda.synchronize()
W = 26
trace = torch.zeros((W, 256, 4), dtype=torch.int32, device=dev)
lib = _lib.load()
lib.nd_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.nd_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
lifting.lift_mean_var(feats[1][:, :, :59, :80], pts, proj)
torch.cuda.synchronize()
lib.nd_debug_set_trace(None)
t = trace.cpu().numpy()
n_st = int((t[0, :, 2] > 0).sum())
print('stages traced per warp:', n_st)
cons = t[:25, :n_st]
wait = (cons[:, :, 1] - cons[:, :, 0])
work = (cons[:, :, 2] - cons[:, :, 1])
ents = cons[:, :, 3]
print('CTA 0 total cycles (last stage end):', cons[:, n_st - 1, 2].max())
print('per warp: total wait cycles, total work cycles, entries, work cycles per entry')
for w in range(25):
    e = max(int(ents[w].sum()), 1)
    print(f'  warp {w:2d}: wait {wait[w].sum():8d}  work {work[w].sum():8d}  entries {e:4d}  cyc/entry {work[w].sum() / e:7.1f}')
print('stage end time of the slowest warp, first 30 stages:', cons[:, :30, 2].max(axis=0).tolist())
print('producer: (wait begin, wait end) first 12 stages:', t[25, :12, :2].tolist())
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gpurun_out', 'lift_trace.npy'), t)
# producer (plane warp) per-stage detail: wait begin, wait end, gap to the next wait begin
pr = t[25, :n_st if n_st < 256 else 255]
print('producer stages 20..40: (wait, issue-to-next-wait):', [(int(pr[i, 1] - pr[i, 0]), int(pr[i + 1, 0] - pr[i, 1])) for i in range(20, 40)])
c0 = t[0]
print('consumer warp 0 stages 20..40: (wait, work):', [(int(c0[i, 1] - c0[i, 0]), int(c0[i, 2] - c0[i, 1])) for i in range(20, 40)])
print('consumer warp 0 stage period:', [int(c0[i + 1, 0] - c0[i, 0]) for i in range(20, 40)])
