"""TEST INFRASTRUCTURE -- seeded inputs of the golden cases.  NOT product code.

``oracle/make_golden.py`` feeds these to the unmodified reference and stores its
outputs under ``tests/golden/<case>.npz``; the tests rebuild the very same inputs
here (numpy ``RandomState`` => identical on every machine) and compare the oracle
and the CUDA path with the stored reference outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nerfdet_b200.synthetic import SceneConfig, make_features, make_mlp_state, make_scene  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

_TINY = dict(ori_shape=(968, 1296), img_shape=(59, 80), pad_shape=(60, 80))

CASES = {
    # B3/B5/B6: tiny image, coarse grid
    'lift_tiny': dict(kind='lift', seed=1, n_views=6, n_voxels=(10, 10, 4),
                      voxel_size=(0.64, 0.64, 0.8), channels=8, **_TINY),
    # B3/B5/B6: the real low-res geometry (239x320 -> 59x80 slice of 60x80), full 40x40x16 grid
    'lift_lowres': dict(kind='lift', seed=2, n_views=6, n_voxels=(40, 40, 16),
                        voxel_size=(0.16, 0.16, 0.2), channels=2),
    # img_shape a multiple of the stride (240 rows -> 60x80, no slicing), shifted origin
    'lift_h240_shift': dict(kind='lift', seed=3, n_views=3, n_voxels=(12, 12, 6),
                            voxel_size=(0.5, 0.5, 0.5), channels=4, ori_shape=(968, 1296),
                            img_shape=(240, 320), pad_shape=(240, 320), shift_origin=True),
    # B4: depth-gated backproject
    'lift_depth': dict(kind='lift', seed=4, n_views=4, n_voxels=(10, 10, 4),
                       voxel_size=(0.64, 0.64, 0.8), channels=4, with_depth=True, **_TINY),
    # one view only (count is 0 or 1)
    'lift_oneview': dict(kind='lift', seed=5, n_views=1, n_voxels=(8, 8, 4),
                         voxel_size=(0.8, 0.8, 0.8), channels=4, **_TINY),
    # whole extract_feat (train mode, rays, density volume) at C = 256
    'extract_small': dict(kind='extract', seed=6, torch_seed=1234, n_views=5, n_voxels=(8, 8, 4),
                          voxel_size=(0.8, 0.8, 0.8), channels=256, n_target_views=2,
                          N_samples=8, N_rand=64, **_TINY),
    # the voxel side of extract_feat with a depth prior: backproject's gate on the 256-channel volume AND on the RGB volume
    # (nerfdet.py:164-169, 204-210), i.e. depth-gated live statistics and density
    'extract_depth': dict(kind='extract_depth', seed=16, torch_seed=4321, n_views=6, n_voxels=(10, 10, 4),
                          voxel_size=(0.64, 0.64, 0.8), channels=256, n_target_views=1,
                          N_samples=8, N_rand=64, **_TINY),
    # render_rays_func, deterministic sampling, with intermediates
    'render_det': dict(kind='render_det', seed=7, n_views=6, n_rays=48, N_samples=16, **_TINY),
    'mlp_small': dict(kind='mlp', seed=8, n_rays=32, N_samples=8),
    # N1: gradients of render_rays_func (rgb, depth) w.r.t. the mapped features and the field's weights, reference autograd
    'render_grad': dict(kind='render_grad', seed=7, n_views=6, n_rays=48, N_samples=16, **_TINY),
    # N1: gradients of nerfdet.py:164-181 with respect to the features, from the reference's own autograd
    'lift_grad_tiny': dict(kind='lift_grad', seed=11, n_views=6, n_voxels=(10, 10, 4),
                           voxel_size=(0.64, 0.64, 0.8), channels=8, **_TINY),
    'lift_grad_depth': dict(kind='lift_grad', seed=12, n_views=4, n_voxels=(10, 10, 4),
                            voxel_size=(0.64, 0.64, 0.8), channels=4, with_depth=True, **_TINY),
    'volume_lookup': dict(kind='volume_lookup', seed=9),
    # N3: ray directions of two target cameras (reference get_dtu_raydir) and de-normalised images (cv2 arithmetic)
    'rays_small': dict(kind='rays', seed=13, n_target_views=2, height=48, width=64, margin=10, n_images=3,
                       ori_shape=(968, 1296), img_shape=(47, 64)),
}


def _scene_cfg(case, **over) -> SceneConfig:
    keys = ('n_views', 'n_voxels', 'voxel_size', 'channels', 'ori_shape', 'img_shape', 'pad_shape',
            'shift_origin', 'n_target_views')
    kw = {k: case[k] for k in keys if k in case}
    kw.update(over)
    return SceneConfig(**kw)


def lift_inputs(case):
    cfg = _scene_cfg(case)
    sc = make_scene(cfg, seed=case['seed'], with_images=False)
    h = cfg.img_shape[0] // cfg.stride
    w = cfg.img_shape[1] // cfg.stride
    out = dict(img_meta=sc.img_meta, stride=cfg.stride, n_voxels=cfg.n_voxels,
               voxel_size=cfg.voxel_size, features=sc.features,
               features_sliced=sc.features[:, :, :h, :w])
    if case.get('with_depth'):
        rs = np.random.RandomState(case['seed'] + 1000)
        out['depth'] = torch.from_numpy(
            rs.uniform(0.5, 4.0, (cfg.n_views,) + tuple(cfg.pad_shape)).astype(np.float32))
    return out


def lift_grad_upstream(case):
    """Seeded incoming gradients (g_mean, g_cov), each [C, X, Y, Z] float32."""
    rs = np.random.RandomState(case['seed'] + 500)
    shape = (case['channels'],) + tuple(case['n_voxels'])
    return (torch.from_numpy(rs.standard_normal(shape).astype(np.float32)),
            torch.from_numpy(rs.standard_normal(shape).astype(np.float32)))


IMG_NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375])


def rays_inputs(case):
    """Seeded cameras (float64 rotations like the dataset's camrotc2w) and normalised images (HWC float32)."""
    rs = np.random.RandomState(case['seed'])
    nt = case['n_target_views']
    q, _ = np.linalg.qr(rs.standard_normal((nt, 3, 3)))
    k = np.array([[1170.19, 0., 647.75, 0.], [0., 1170.19, 483.75, 0.], [0., 0., 1., 0.], [0., 0., 0., 1.]], dtype=np.float32)
    img_meta = dict(lidar2img=dict(intrinsic=k), ori_shape=case['ori_shape'] + (3,), img_shape=case['img_shape'] + (3,))
    pixels = rs.randint(0, 256, (case['n_images'], case['height'], case['width'], 3)).astype(np.float32)
    mean, std = np.array(IMG_NORM['mean'], dtype=np.float32), np.array(IMG_NORM['std'], dtype=np.float32)
    img = ((pixels - mean) / std).astype(np.float32)            # what Normalize hands on (RGB, HWC)
    img[:, -1] = 0.0                                            # a padded row, as Pad leaves it
    return dict(img_meta=img_meta, camrotc2w=q.astype(np.float64), lightpos=rs.standard_normal((nt, 3)).astype(np.float32),
                img_hwc=img, height=case['height'], width=case['width'], margin=case['margin'])


RENDER_GRAD_KEYS = ('mlp.base.hidden_layers.0.weight', 'mlp.base.hidden_layers.3.bias', 'mlp.sigma_layer.output_layer.weight',
                    'mlp.bottleneck_layer.output_layer.bias', 'mlp.rgb_layer.hidden_layers.0.weight',
                    'mlp.rgb_layer.output_layer.weight', 'mlp.rgb_layer.output_layer.bias')


def render_grad_upstream(case):
    """Seeded incoming gradients for the rendered colour [rays, 3] and depth [rays]."""
    rs = np.random.RandomState(case['seed'] + 700)
    return (torch.from_numpy(rs.standard_normal((case['n_rays'], 3)).astype(np.float32)),
            torch.from_numpy(rs.standard_normal((case['n_rays'],)).astype(np.float32)))


def extract_inputs(case):
    cfg = _scene_cfg(case)
    sc = make_scene(cfg, seed=case['seed'])
    return dict(img_meta=sc.img_meta, n_voxels=cfg.n_voxels, voxel_size=cfg.voxel_size,
                features=sc.features, pad_shape=cfg.pad_shape, aabb=cfg.aabb,
                near_far_range=list(cfg.near_far_range), N_samples=case['N_samples'],
                N_rand=case['N_rand'], ray_batch=sc.ray_batch,
                state=make_mlp_state(case['seed'] + 100))


def extract_depth_inputs(case):
    """extract_inputs plus a seeded depth prior ``[nv, Hp, Wp]`` that keeps roughly a third of the voxel-views."""
    inp = extract_inputs(case)
    rs = np.random.RandomState(case['seed'] + 1000)
    inp['depth'] = torch.from_numpy(rs.uniform(0.5, 4.0, (case['n_views'],) + tuple(inp['pad_shape'])).astype(np.float32))
    return inp


def render_inputs(case):
    cfg = _scene_cfg(case, channels=32, n_target_views=1)
    sc = make_scene(cfg, seed=case['seed'])
    rs = np.random.RandomState(case['seed'] + 2000)
    rb = sc.ray_batch
    total = rb['ray_d'].view(-1, 3).shape[0]
    sel = rs.choice(total, size=(case['n_rays'],), replace=False)
    h = cfg.img_shape[0] // cfg.stride
    w = cfg.img_shape[1] // cfg.stride
    return dict(img_meta=sc.img_meta, ray_o=rb['ray_o'].view(-1, 3)[sel].contiguous(),
                ray_d=rb['ray_d'].view(-1, 3)[sel].contiguous(),
                featmaps=sc.features[:, :, :h, :w].contiguous(), images=sc.denorm_images[0],
                aabb=cfg.aabb, near_far_range=list(cfg.near_far_range),
                N_samples=case['N_samples'], state=make_mlp_state(case['seed'] + 100))


def mlp_inputs(case):
    rs = np.random.RandomState(case['seed'])
    r, s = case['n_rays'], case['N_samples']
    pts = torch.from_numpy(rs.uniform(-3.5, 3.5, (r, s, 3)).astype(np.float32))
    ray_d = torch.from_numpy(rs.normal(0, 0.7, (r, 3)).astype(np.float32))
    feats = torch.from_numpy(np.concatenate(
        [rs.normal(0, 1.0, (r, s, 35)), rs.uniform(0, 1, (r, s, 35))], axis=-1).astype(np.float32))
    return dict(pts=pts, ray_d=ray_d, feats=feats, state=make_mlp_state(case['seed'] + 100))


def volume_lookup_inputs(case):
    rs = np.random.RandomState(case['seed'])
    volume = torch.from_numpy(rs.standard_normal((1, 5, 6, 7, 4)).astype(np.float32))
    aabb = ([-2.7, -2.7, -0.78], [3.7, 3.7, 1.78])
    pts = torch.from_numpy(rs.uniform(-3.5, 4.5, (16, 8, 3)).astype(np.float32))
    pts[..., 2] = pts[..., 2] * 0.4
    return dict(volume=volume, aabb=aabb, pts=pts)


def load_golden(name):
    path = os.path.join(GOLDEN_DIR, f'{name}.npz')
    with np.load(path) as z:
        return {k: z[k] for k in z.files}
