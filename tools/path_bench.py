#!/usr/bin/env python
"""Timings of the hot path beyond bench.py's headline step (BASELINE.json configs[1..4] on one GPU):

  lift sweep     fused lift (nerfdet.py:164-181) over view counts and voxel grids (configs[1], [3], [4]); every shape
                 with its algorithmic bytes and the fraction of the measured HBM peak
  live path      one scene of extract_feat's loop body without render_rays (nerfdet.py:152-261): 2-D mapping, live
                 35-channel statistics, density MLP, fused lift with alpha -- fp32 and bf16 (tcgen05) MLP
  render         render_rays_func at N_rand = 2048, N_samples = 64 (configs[2]) -- fp32 and bf16 MLP

Each number: CUDA events around a CUDA-graph replay of the calls (no host launch gaps) and, beside it, the eager
time through the Python API.  Inputs rotate over 3 feature sets so that no call finds its planes in L2.
One JSON line per measurement."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfdet_b200 import lifting, live, render  # noqa: E402
from nerfdet_b200.nerf_mlp import VanillaNeRFRadianceField  # noqa: E402
from nerfdet_b200.projection import Projector  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_mlp_state, make_scene  # noqa: E402

DEV = torch.device('cuda', 0)
PEAK = 6551.0
try:
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'MEASURED_PEAKS.json')) as fh:
        PEAK = float(json.load(fh)['hbm_gbs'])
except Exception:
    pass


def timed(fn, n_variants=3, reps=20):
    for i in range(4):
        fn(i % n_variants)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % n_variants)
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / reps * 1e3
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    try:
        with torch.cuda.stream(side):
            fn(0)
            with torch.cuda.graph(g, stream=side):
                for i in range(2 * n_variants):
                    fn(i % n_variants)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / (10 * n_variants) * 1e3
    except Exception as exc:                      # host-synchronising call: cannot be captured
        torch.cuda.synchronize()
        graph = float('nan')
        print(json.dumps({'note': f'graph capture failed: {type(exc).__name__}: {exc}'[:200]}))
    return eager, graph


def lift_sweep():
    shapes = [(20, (40, 40, 16), (.16, .16, .2)), (50, (40, 40, 16), (.16, .16, .2)), (100, (40, 40, 16), (.16, .16, .2)),
              (50, (56, 56, 16), (.16, .16, .2)), (50, (64, 64, 24), (.1, .1, .13)), (50, (80, 80, 32), (.08, .08, .08))]
    for nv, grid, vs in shapes:
        cfg = SceneConfig(n_views=nv, n_voxels=grid, voxel_size=vs, channels=256)
        sc = make_scene(cfg, seed=1000, with_images=False, with_features=False)
        proj = lifting.compute_projection(sc.img_meta, 4).to(DEV)
        pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin']).to(DEV)
        sets = [torch.randn(nv, 256, 60, 80, device=DEV) for _ in range(3)]
        eager, graph = timed(lambda i: lifting.lift_mean_var(sets[i][:, :, :59, :80], pts, proj))
        n = int(np.prod(grid))

        def report(dtype, elt, eager, graph):
            byts = nv * 256 * 59 * 80 * elt + 2 * 256 * n * 4 + n * 8 + nv * 48
            t = graph if graph == graph else eager
            print(json.dumps({'what': 'lift', 'features': dtype, 'views': nv, 'grid': list(grid), 'us': round(t, 1),
                              'us_eager': round(eager, 1), 'gsamples_per_s': round(nv * n / t / 1e3, 2),
                              'algorithmic_mb': round(byts / 1e6, 1), 'gbs': round(byts / t / 1e3, 1),
                              'frac_of_hbm_peak': round(byts / t / 1e3 / PEAK, 3)}), flush=True)
        report('f32', 4, eager, graph)
        if nv == 50 and grid == (40, 40, 16):                    # bf16 feature variant of configs[1]
            sets16 = [s.to(torch.bfloat16) for s in sets]
            e16, g16 = timed(lambda i: lifting.lift_mean_var(sets16[i][:, :, :59, :80], pts, proj))
            report('bf16', 2, e16, g16)
            del sets16
        del sets


def live_and_render():
    cfg = SceneConfig(n_views=50, n_voxels=(40, 40, 16), voxel_size=(.16, .16, .2), channels=256, n_target_views=2)
    sc = make_scene(cfg, seed=1000, with_images=True, with_features=False)
    state = make_mlp_state(191)
    mapping = torch.nn.Sequential(torch.nn.Linear(256, 32))
    mapping.load_state_dict({'0.weight': state['mapping.0.weight'], '0.bias': state['mapping.0.bias']})
    mapping = mapping.to(DEV)
    imgs = sc.denorm_images[0].to(DEV)
    sets = [torch.randn(50, 256, 60, 80, device=DEV) for _ in range(3)]
    rb = sc.ray_batch
    rs = np.random.RandomState(5)
    sel = rs.choice(rb['ray_o'].view(-1, 3).shape[0], 2048, replace=False)
    ray_o = rb['ray_o'].view(-1, 3)[sel].float().to(DEV)
    ray_d = rb['ray_d'].view(-1, 3)[sel].float().to(DEV)
    for prec in ('fp32', 'bf16'):
        field = VanillaNeRFRadianceField(4, 256, 3, 70, 1, 128, precision=prec)
        field.load_state_dict({k: v for k, v in state.items() if not k.startswith('mapping.')})
        field = field.to(DEV)
        with torch.no_grad():
            eager, graph = timed(lambda i: live.lift_scene(sets[i], sc.img_meta, cfg.n_voxels, cfg.voxel_size, mapping, field, imgs))
            print(json.dumps({'what': 'live path (mapping + 35-ch statistics + density MLP + fused lift)', 'mlp': prec,
                              'views': 50, 'us': round(graph, 1), 'us_eager': round(eager, 1)}), flush=True)
            out = live.lift_scene(sets[0], sc.img_meta, cfg.n_voxels, cfg.voxel_size, mapping, field, imgs)
            f2d = out['feature_2d']
            proj = Projector()
            eager, graph = timed(lambda i: render.render_rays_func(
                ray_o, ray_d, None, None, f2d, imgs, cfg.aabb, cfg.near_far_range, 64, 2048, field, sc.img_meta, proj,
                'image', 3, False, 0, True))
            print(json.dumps({'what': 'render_rays_func 2048 rays x 64 samples, 50 source views', 'mlp': prec,
                              'us': round(graph, 1), 'us_eager': round(eager, 1),
                              'mrays_per_s': round(2048 / (graph if graph == graph else eager), 2)}), flush=True)


if __name__ == '__main__':
    torch.backends.cuda.matmul.allow_tf32 = False
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'lift'):
        lift_sweep()
    if which in ('all', 'path'):
        live_and_render()
