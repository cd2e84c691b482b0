#!/usr/bin/env python
"""Times the fused lift at a given view count with the tuning/diagnostic env knobs
(ND_LIFT_STAGES / ND_LIFT_WARPS / ND_LIFT_DEBUG) -- a probe for kernel development, not a benchmark.
usage: lift_probe.py "<nv> <channels> <stages> <group> <warps> <debug>" ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfdet_b200 import lifting  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_scene  # noqa: E402


def run(nv, ch, steps=40):
    dev = torch.device('cuda', 0)
    cfg = SceneConfig(n_views=nv, n_voxels=(40, 40, 16), voxel_size=(0.16, 0.16, 0.2), channels=ch)
    sc = make_scene(cfg, seed=1000, with_images=False, with_features=False)
    proj = lifting.compute_projection(sc.img_meta, 4).to(dev)
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin']).to(dev)
    sets = [torch.randn(nv, ch, 60, 80, device=dev) for _ in range(3)]
    for i in range(5):
        lifting.lift_mean_var(sets[i % 3][:, :, :59, :80], pts, proj)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        lifting.lift_mean_var(sets[i % 3][:, :, :59, :80], pts, proj)
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / steps * 1e3
    if os.environ.get('PROBE_GRAPH', '1') == '0':
        return eager, float('nan')
    # the same steps replayed from a CUDA graph: no host launch overhead
    g = torch.cuda.CUDAGraph()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for i in range(6):
                out = lifting.lift_mean_var(sets[i % 3][:, :, :59, :80], pts, proj)
        g.replay()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps // 6 + 1):
            g.replay()
        e1.record(stream)
    torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / ((steps // 6 + 1) * 6) * 1e3
    return eager, graph


for cfg in sys.argv[1:]:
    nv, ch, st, grp, w, dbg, l2a = (cfg.split() + ['0'])[:7]
    os.environ['ND_LIFT_L2AHEAD'] = l2a
    os.environ['ND_LIFT_STAGES'], os.environ['ND_LIFT_WARPS'], os.environ['ND_LIFT_DEBUG'] = st, w, dbg
    os.environ['ND_LIFT_GROUP'] = grp
    print(f'nv={nv} C={ch} stages={st} group={grp} warps={w} debug={dbg} l2ahead={l2a}: eager/graph us per step = %.1f / %.1f' % run(int(nv), int(ch)), flush=True)
