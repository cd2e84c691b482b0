#!/usr/bin/env python
"""Where the time of one ``render_rays_func`` call (2048 rays x 64 samples, 50 views) goes: the host time of the
call and of its camera packing, and each kernel on its own with CUDA events.

  python tools/render_dev.py [--mlp bf16|fp32]      (GPU box)
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerfdet_b200 import ops, render  # noqa: E402
from nerfdet_b200.nerf_mlp import VanillaNeRFRadianceField  # noqa: E402
from nerfdet_b200.projection import Projector  # noqa: E402


def timed(fn, steps=50, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


def host(fn, steps=200):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps * 1e6
    torch.cuda.synchronize()
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mlp', default='bf16')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    cfg, sc, state = bench.render_scene()
    n_rand, n_samples = 2048, 64
    field = VanillaNeRFRadianceField(4, 256, 3, 70, 1, 128, precision=args.mlp)
    field.load_state_dict({k: v for k, v in state.items() if not k.startswith('mapping.')})
    field = field.to(dev)
    imgs = sc.denorm_images[0].to(dev)
    f2d = sc.features[:, :, :bench.FEAT_HW[0], :bench.FEAT_HW[1]].contiguous().to(dev).contiguous(memory_format=torch.channels_last)
    rb = sc.ray_batch
    sel = np.random.RandomState(5).choice(rb['ray_o'].view(-1, 3).shape[0], n_rand, replace=False)
    ro, rd = rb['ray_o'].view(-1, 3)[sel].float().to(dev), rb['ray_d'].view(-1, 3)[sel].float().to(dev)
    proj = Projector()

    with torch.no_grad():
        def step():
            return render.render_rays_func(ro, rd, None, None, f2d, imgs, cfg.aabb, cfg.near_far_range, n_samples, n_rand, field,
                                           sc.img_meta, proj, 'image', 3, False, 0, True)
        print(f'render_rays_func: {timed(step):.1f} us/step on the device, {host(step):.1f} us host time per call', flush=True)
        print(f'  camera packing on the host: {host(lambda: render._compute_projection(sc.img_meta)):.1f} us; with the upload: '
              f'{host(lambda: render._compute_projection(sc.img_meta)[0].to(dev)):.1f} us', flush=True)
        # the stages of one call in order, events between them (averaged over 30 calls)
        names = ['sample_rays', 'camera upload', 'gather', 'mlp', 'bounds', 'composite']
        acc = np.zeros(len(names))
        for rep in range(35):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            torch.cuda.synchronize()
            ev[0].record()
            pts, z = render.sample_along_camera_ray(ro, rd, cfg.near_far_range, n_samples, det=True)
            ev[1].record()
            cams = render._compute_projection(sc.img_meta)[0].to(dev)
            ev[2].record()
            glob, _, pm, _, _, _ = ops.direct.render_gather_stats(pts.view(-1, 3), cams, imgs, f2d, False, False, False)
            ev[3].record()
            rgb, sig = field(pts, rd, glob.view(n_rand, n_samples, -1))
            ev[4].record()
            bounds = torch.stack([z[0, 0], z[0, -1]])
            ev[5].record()
            out = ops.direct.composite(rgb, sig.reshape(n_rand, n_samples), z, pm.view(n_rand, n_samples), bounds, False)
            ev[6].record()
            torch.cuda.synchronize()
            if rep >= 5:
                acc += np.array([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(len(names))])
        print('  stages of one call, host not running ahead (us): ' + ', '.join(f'{n} {t / 30:.1f}' for n, t in zip(names, acc)), flush=True)
        pts, z = render.sample_along_camera_ray(ro, rd, cfg.near_far_range, n_samples, det=True)
        cams = render._compute_projection(sc.img_meta)[0].to(dev)
        flat = pts.view(-1, 3)
        print(f'  sample_rays: {timed(lambda: render.sample_along_camera_ray(ro, rd, cfg.near_far_range, n_samples, det=True)):.1f} us', flush=True)
        print(f'  render_gather_stats: {timed(lambda: ops.direct.render_gather_stats(flat, cams, imgs, f2d, False, False)):.1f} us '
              f'(host {host(lambda: ops.direct.render_gather_stats(flat, cams, imgs, f2d, False, False)):.1f})', flush=True)
        imcl = imgs.contiguous(memory_format=torch.channels_last)
        print(f'  render_gather_stats, no view mask: {timed(lambda: ops.direct.render_gather_stats(flat, cams, imgs, f2d, False, False, False)):.1f} us; '
              f'channels-last images: {timed(lambda: ops.direct.render_gather_stats(flat, cams, imcl, f2d, False, False, False)):.1f} us; '
              f'the layout change itself: {timed(lambda: imgs.contiguous(memory_format=torch.channels_last)):.1f} us', flush=True)
        a = ops.direct.render_gather_stats(flat, cams, imgs, f2d, False, False, False)[0]
        b = ops.direct.render_gather_stats(flat, cams, imcl, f2d, False, False, False)[0]
        print('  same result from both image layouts:', bool(torch.equal(a, b)), flush=True)
        glob, _, pm, _, _, _ = ops.direct.render_gather_stats(flat, cams, imgs, f2d, False, False)
        g3 = glob.view(n_rand, n_samples, -1)
        print(f'  nerf_mlp: {timed(lambda: field(pts, rd, g3)):.1f} us (host {host(lambda: field(pts, rd, g3)):.1f})', flush=True)
        rgb, sig = field(pts, rd, g3)
        sg = sig.reshape(n_rand, n_samples)
        pmv = pm.view(n_rand, n_samples)
        print(f'  composite: {timed(lambda: render._composite(rgb, sg, z, pmv, False, det=True)):.1f} us '
              f'(host {host(lambda: render._composite(rgb, sg, z, pmv, False, det=True)):.1f})', flush=True)


if __name__ == '__main__':
    main()
