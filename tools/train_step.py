#!/usr/bin/env python
"""Forward + backward times of the differentiable path at the bench shapes (row N1): the fused lift, the whole voxel side
(live.lift_scene: mapping, live statistics, density, lift) and the render branch (2048 rays x 64 samples).

  python tools/train_step.py      (GPU box)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerfdet_b200 import lifting, live, render  # noqa: E402
from nerfdet_b200.nerf_mlp import VanillaNeRFRadianceField  # noqa: E402
from nerfdet_b200.projection import Projector  # noqa: E402

DEV = torch.device('cuda', 0)


def timed(fn, steps=20, warmup=3):
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


def main():
    torch.cuda.set_device(0)
    from nerfdet_b200.synthetic import SceneConfig, make_mlp_state, make_scene
    cfg = SceneConfig(n_views=50, n_target_views=1)
    sc = make_scene(cfg, seed=3000)
    state = make_mlp_state(3100)
    field = VanillaNeRFRadianceField(4, 256, 3, 70, 1, 128, precision='fp32')
    field.load_state_dict({k: v for k, v in state.items() if not k.startswith('mapping.')})
    field = field.to(DEV)
    lin = torch.nn.Linear(256, 32).to(DEV)
    with torch.no_grad():
        lin.weight.copy_(state['mapping.0.weight'])
        lin.bias.copy_(state['mapping.0.bias'])
    mapping = torch.nn.Sequential(lin)
    feats = sc.features.to(DEV).requires_grad_(True)
    imgs = sc.denorm_images[0].to(DEV)
    proj = lifting.compute_projection(sc.img_meta, 4).to(DEV)
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin']).to(DEV)
    sliced = feats[:, :, :bench.FEAT_HW[0], :bench.FEAT_HW[1]]

    def lift_fb():
        mean, cov, _ = lifting.lift_mean_var(sliced, pts, proj)
        (mean.sum() + cov.sum()).backward()
        feats.grad = None
    with torch.no_grad():
        f_us = timed(lambda: lifting.lift_mean_var(sliced.detach(), pts, proj))
    print(f'fused lift (50 x 256 x 59x80 -> 40x40x16): forward {f_us:.0f} us, forward + backward {timed(lift_fb):.0f} us', flush=True)

    def scene_fb():
        res = live.lift_scene(feats, sc.img_meta, cfg.n_voxels, cfg.voxel_size, mapping=mapping, nerf_mlp=field, denorm_images=imgs)
        res['volume'].sum().backward()
        feats.grad = None
        for p in list(field.parameters()) + list(lin.parameters()):
            p.grad = None
    with torch.no_grad():
        s_us = timed(lambda: live.lift_scene(feats.detach(), sc.img_meta, cfg.n_voxels, cfg.voxel_size, mapping=mapping,
                                             nerf_mlp=field, denorm_images=imgs))
    print(f'live.lift_scene (mapping + live statistics + density + lift): forward {s_us:.0f} us, forward + backward '
          f'{timed(scene_fb):.0f} us', flush=True)

    with torch.no_grad():
        f2d0 = live.map_features_2d(sliced.detach(), mapping)
    f2d = f2d0.detach().clone().requires_grad_(True)
    rb = sc.ray_batch
    sel = np.random.RandomState(5).choice(rb['ray_o'].view(-1, 3).shape[0], 2048, replace=False)
    ro, rd = rb['ray_o'].view(-1, 3)[sel].float().to(DEV), rb['ray_d'].view(-1, 3)[sel].float().to(DEV)
    pj = Projector()

    def call():
        return render.render_rays_func(ro, rd, None, None, f2d, imgs, cfg.aabb, cfg.near_far_range, 64, 2048, field, sc.img_meta, pj,
                                       'image', 3, False, 0, True)['outputs_coarse']

    def render_fb():
        out = call()
        (out['rgb'].sum() + out['depth'].sum()).backward()
        f2d.grad = None
        for p in field.parameters():
            p.grad = None
    with torch.no_grad():
        r_us = timed(call)
    print(f'render_rays_func (2048 rays x 64 samples x 50 views, fp32-grade MLP): forward {r_us:.0f} us, forward + backward '
          f'{timed(render_fb):.0f} us', flush=True)
    field.precision = 'bf16'
    with torch.no_grad():
        r_us = timed(call)
    print(f'render_rays_func, bf16 MLP forward (its backward re-evaluates the fp32 formula): forward {r_us:.0f} us, forward + backward '
          f'{timed(render_fb):.0f} us', flush=True)


    # row N4: one full target image (220 x 300 rays x 64 samples) through render_rays(render_testing=True) + the metrics
    from nerfdet_b200 import evaluate
    rbt = {k: (v[:, :1] if torch.is_tensor(v) and v.dim() >= 3 else v) for k, v in rb.items() if k != 'camrotc2w'}
    rbt['nerf_sizes'] = rb['nerf_sizes'][:, :1]
    with torch.no_grad():
        def full():
            return render.render_rays(rbt, None, None, f2d.detach(), imgs, cfg.aabb, cfg.near_far_range, 64, 2048, field, sc.img_meta,
                                      pj, 'image', is_train=False, render_testing=True)
        us = timed(full, 5, 2)
        out = full()
        m_us = timed(lambda: evaluate.image_metrics(out['outputs_coarse']['rgb'], out['gt_rgb'].to(DEV)), 20, 3)
    n_rays = out['outputs_coarse']['rgb'].shape[1] * out['outputs_coarse']['rgb'].shape[2]
    print(f'render_testing, one {out["outputs_coarse"]["rgb"].shape[1]} x {out["outputs_coarse"]["rgb"].shape[2]} image ({n_rays} rays x 64 '
          f'samples, bf16 MLP): {us / 1e3:.2f} ms = {n_rays / us:.2f} M rays/s; PSNR + SSIM of the image: {m_us:.0f} us', flush=True)


if __name__ == '__main__':
    main()
