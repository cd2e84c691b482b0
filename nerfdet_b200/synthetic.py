"""Synthetic ScanNet-shaped scenes for the lifting path.

Everything is drawn from ``numpy.random.RandomState(seed)`` so the same seed
gives the same scene on every machine and torch version; tensors are created on
the CPU and moved by the caller.  The contract mimicked is what the reference's
data layer hands to ``nerfdet.extract_feat``:

* ``img_meta['lidar2img']`` = ``{intrinsic 4x4 f32, extrinsic [nv x 4x4 f32] (world->camera),
  origin f32[3]}`` (reference ``mmdet3d/datasets/scannet_monocular_dataset.py:44-53``),
* ``img_shape`` / ``ori_shape`` (reference ``configs/nerfdet/nerfdet_res50_2x_low_res.py:95-97``),
* stride-4 FPN-like features ``[nv, C, Hpad/4, Wpad/4]``,
* ``denorm_images [1, nv, 3, Hpad, Wpad]`` in [0, 1],
* rays: ``lightpos`` / ``raydirs`` (un-normalised, camera z = 1; reference
  ``mmdet3d/datasets/pipelines/multi_view.py:124-132`` and
  ``data_augment_utils.py:410-424``), ``gt_images`` / ``gt_depths`` float64.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# ScanNet colour intrinsics for 1296x968 frames
_FX = 1170.19
_FY = 1170.19
_CX = 647.75
_CY = 483.75


@dataclass
class SceneConfig:
    """Shape of one synthetic scene.  Defaults = nerfdet_res50_2x_low_res."""

    n_views: int = 50
    n_voxels: Tuple[int, int, int] = (40, 40, 16)
    voxel_size: Tuple[float, float, float] = (0.16, 0.16, 0.2)
    channels: int = 256
    ori_shape: Tuple[int, int] = (968, 1296)
    img_shape: Tuple[int, int] = (239, 320)      # after Resize(keep_ratio)
    pad_shape: Tuple[int, int] = (240, 320)      # after Pad
    stride: int = 4
    origin: Tuple[float, float, float] = (0.0, 0.0, 0.5)
    n_target_views: int = 0                       # >0 => rays are generated
    margin: int = 10
    aabb: Tuple[Tuple[float, float, float], Tuple[float, float, float]] = (
        (-2.7, -2.7, -0.78), (3.7, 3.7, 1.78))
    near_far_range: Tuple[float, float] = (0.2, 8.0)
    shift_origin: bool = False

    @property
    def feat_hw(self) -> Tuple[int, int]:
        return self.pad_shape[0] // self.stride, self.pad_shape[1] // self.stride

    @property
    def n_voxels_total(self) -> int:
        return int(np.prod(self.n_voxels))


@dataclass
class Scene:
    cfg: SceneConfig
    img_meta: Dict
    features: torch.Tensor                 # [nv, C, Hf_pad, Wf_pad] f32 (un-sliced)
    denorm_images: torch.Tensor            # [1, nv, 3, Hpad, Wpad] f32
    ray_batch: Optional[Dict] = None
    c2w: List[np.ndarray] = field(default_factory=list)


def _look_at_c2w(cam: np.ndarray, target: np.ndarray) -> np.ndarray:
    """OpenCV camera (x right, y down, z forward), world z up."""
    fwd = target - cam
    fwd = fwd / np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right = right / np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4, dtype=np.float64)
    c2w[:3, 0] = right
    c2w[:3, 1] = down
    c2w[:3, 2] = fwd
    c2w[:3, 3] = cam
    return c2w


def _draw_cameras(rs: np.random.RandomState, n: int) -> List[np.ndarray]:
    cams = []
    for _ in range(n):
        radius = rs.uniform(1.0, 2.5)
        height = rs.uniform(1.0, 1.6)
        az = rs.uniform(0.0, 2.0 * np.pi)
        target = np.array([rs.normal(0.0, 0.5), rs.normal(0.0, 0.5), 0.5])
        cam = np.array([radius * np.cos(az), radius * np.sin(az), height])
        cams.append(_look_at_c2w(cam, target))
    return cams


def make_intrinsic(cfg: SceneConfig) -> np.ndarray:
    k = np.eye(4, dtype=np.float32)
    sy = cfg.ori_shape[0] / 968.0
    sx = cfg.ori_shape[1] / 1296.0
    k[0, 0] = _FX * sx
    k[1, 1] = _FY * sy
    k[0, 2] = _CX * sx
    k[1, 2] = _CY * sy
    return k


def make_features(rs: np.random.RandomState, shape, dtype=np.float32) -> np.ndarray:
    """FPN-like: 1.5*relu(N(0,1)) + 0.1*N(0,1)  (SURVEY.md §8d)."""
    a = rs.standard_normal(shape).astype(np.float32)
    b = rs.standard_normal(shape).astype(np.float32)
    return (1.5 * np.maximum(a, 0.0) + 0.1 * b).astype(dtype)


def make_scene(cfg: SceneConfig, seed: int = 0, with_images: bool = True,
               with_features: bool = True) -> Scene:
    rs = np.random.RandomState(seed)
    c2ws = _draw_cameras(rs, cfg.n_views)
    extrinsic = [np.linalg.inv(m).astype(np.float32) for m in c2ws]
    origin = np.array(cfg.origin, dtype=np.float32)
    if cfg.shift_origin:
        origin = origin + rs.normal(0.0, [0.7, 0.7, 0.0]).astype(np.float32)
    img_meta = dict(
        lidar2img=dict(intrinsic=make_intrinsic(cfg), extrinsic=extrinsic, origin=origin),
        img_shape=(cfg.img_shape[0], cfg.img_shape[1], 3),
        ori_shape=(cfg.ori_shape[0], cfg.ori_shape[1], 3),
        pad_shape=(cfg.pad_shape[0], cfg.pad_shape[1], 3),
    )
    hf, wf = cfg.feat_hw
    if with_features:
        feats = torch.from_numpy(make_features(rs, (cfg.n_views, cfg.channels, hf, wf)))
    else:
        feats = torch.empty(0)
    if with_images:
        imgs = torch.from_numpy(
            rs.uniform(0.0, 1.0, (1, cfg.n_views, 3, cfg.pad_shape[0], cfg.pad_shape[1])
                       ).astype(np.float32))
    else:
        imgs = torch.empty(0)

    ray_batch = None
    if cfg.n_target_views > 0:
        ray_batch = make_rays(cfg, rs, img_meta)
        ray_batch['denorm_images'] = imgs
    return Scene(cfg=cfg, img_meta=img_meta, features=feats, denorm_images=imgs,
                 ray_batch=ray_batch, c2w=c2ws)


def make_rays(cfg: SceneConfig, rs: np.random.RandomState, img_meta: Dict) -> Dict:
    """Rays of ``n_target_views`` extra cameras, shaped like the collated batch
    (``[B=1, nt, n_pix, 3]``)."""
    nt = cfg.n_target_views
    tcams = _draw_cameras(rs, nt)
    height, width = cfg.pad_shape
    ratio = cfg.ori_shape[0] / cfg.img_shape[0]
    k = img_meta['lidar2img']['intrinsic'].copy()
    k[:2] = k[:2] / ratio
    px, py = np.meshgrid(
        np.arange(cfg.margin, width - cfg.margin).astype(np.float32),
        np.arange(cfg.margin, height - cfg.margin).astype(np.float32))
    raydirs, lightpos = [], []
    for c2w in tcams:
        x = (px + 0.5 - k[0, 2]) / k[0, 0]
        y = (py + 0.5 - k[1, 2]) / k[1, 1]
        d = np.stack([x, y, np.ones_like(x)], axis=-1) @ c2w[:3, :3].T
        d = d.reshape(-1, 3).astype(np.float32)
        raydirs.append(d)
        lightpos.append(np.broadcast_to(c2w[:3, 3].astype(np.float32), d.shape).copy())
    npix = raydirs[0].shape[0]
    gt_rgb = rs.uniform(0.0, 1.0, (1, nt, npix, 3))                  # float64
    gt_depth = rs.uniform(0.5, 1.5, (1, nt, npix))                   # float64
    # a few zero depths so the reference's gt_depth > 0 filter does something
    gt_depth[rs.uniform(size=gt_depth.shape) < 0.02] = 0.0
    nerf_sizes = np.tile(np.array([[py.shape[0], py.shape[1], 3]]), (nt, 1))[None]
    return dict(
        ray_o=torch.from_numpy(np.stack(lightpos)[None]),
        ray_d=torch.from_numpy(np.stack(raydirs)[None]),
        gt_rgb=torch.from_numpy(gt_rgb),
        gt_depth=torch.from_numpy(gt_depth),
        nerf_sizes=torch.from_numpy(nerf_sizes),
        camrotc2w=torch.from_numpy(np.stack([c[:3, :3] for c in tcams]).astype(np.float64)),   # the target cameras themselves
    )


def make_mlp_state(seed: int, feature_dim: int = 70, net_width: int = 256,
                   net_depth: int = 4, cond_width: int = 128, bias_std: float = 0.1,
                   map_in: int = 256, map_out: int = 32) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference ``state_dict`` keys:
    xavier-uniform weights; biases get N(0, bias_std) so that the
    bias-for-invalid-view quirk is exercised."""
    rs = np.random.RandomState(seed)

    def xavier(out_f, in_f):
        a = np.sqrt(6.0 / (in_f + out_f))
        return torch.from_numpy(rs.uniform(-a, a, (out_f, in_f)).astype(np.float32))

    def bias(n):
        return torch.from_numpy((bias_std * rs.standard_normal(n)).astype(np.float32))

    in_dim = 63 + feature_dim
    sd = {}
    d_in = in_dim
    for i in range(net_depth):
        sd[f'mlp.base.hidden_layers.{i}.weight'] = xavier(net_width, d_in)
        sd[f'mlp.base.hidden_layers.{i}.bias'] = bias(net_width)
        d_in = net_width
    hid = net_width + in_dim
    sd['mlp.sigma_layer.output_layer.weight'] = xavier(1, hid)
    sd['mlp.sigma_layer.output_layer.bias'] = bias(1)
    sd['mlp.bottleneck_layer.output_layer.weight'] = xavier(net_width, hid)
    sd['mlp.bottleneck_layer.output_layer.bias'] = bias(net_width)
    sd['mlp.rgb_layer.hidden_layers.0.weight'] = xavier(cond_width, net_width + 27)
    sd['mlp.rgb_layer.hidden_layers.0.bias'] = bias(cond_width)
    sd['mlp.rgb_layer.output_layer.weight'] = xavier(3, cond_width)
    sd['mlp.rgb_layer.output_layer.bias'] = bias(3)
    sd['posi_encoder.scales'] = torch.tensor([2 ** i for i in range(10)])
    sd['view_encoder.scales'] = torch.tensor([2 ** i for i in range(4)])
    a = 1.0 / np.sqrt(map_in)
    sd['mapping.0.weight'] = torch.from_numpy(
        rs.uniform(-a, a, (map_out, map_in)).astype(np.float32))
    sd['mapping.0.bias'] = bias(map_out)
    return sd
