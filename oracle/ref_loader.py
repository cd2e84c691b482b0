"""TEST INFRASTRUCTURE -- not product code.

Loads the *unmodified* reference implementation of the lifting path from
``/root/reference`` (read-only; exists only in the build container, never on the
GPU box) so that (i) the oracle restatement in this directory can be validated
against it and (ii) golden vectors can be generated (``oracle/make_golden.py``).

``projection.py``, ``render_ray.py`` and ``nerf_mlp.py`` import only torch/numpy and
are loaded by file path.  ``nerfdet.py`` imports mmdet / mmdet3d.core, which are
not installed; they are replaced by minimal stand-ins in ``sys.modules``
(SURVEY.md §8c).  Nothing here is imported by the product package.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get('NERFDET_REFERENCE_ROOT', '/root/reference')
_MU = 'mmdet3d/models/model_utils'
_cache = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _MU, 'render_ray.py'))


def _load(name: str, rel: str, package: str | None = None):
    path = os.path.join(REFERENCE_ROOT, rel)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _install_stubs():
    if 'mmdet' in sys.modules and getattr(sys.modules['mmdet'], '_nd_stub', False):
        return

    class _Registry:
        def register_module(self, *a, **k):
            def deco(cls):
                return cls
            return deco

    class _StubNet(nn.Module):
        def __init__(self, cfg=None):
            super().__init__()
            self.cfg = cfg

        def init_weights(self, *a, **k):
            pass

        def forward(self, x):
            return x

    class BaseDetector(nn.Module):
        def init_weights(self, pretrained=None):
            pass

    def _build(cfg):
        if isinstance(cfg, nn.Module):
            return cfg
        if isinstance(cfg, dict) and isinstance(cfg.get('module'), nn.Module):
            return cfg['module']
        return _StubNet(cfg)

    mmdet = types.ModuleType('mmdet')
    mmdet._nd_stub = True
    models = types.ModuleType('mmdet.models')
    models.DETECTORS = _Registry()
    models.build_backbone = _build
    models.build_neck = _build
    models.build_head = _build
    detectors = types.ModuleType('mmdet.models.detectors')
    detectors.BaseDetector = BaseDetector
    mmdet.models = models
    models.detectors = detectors
    sys.modules['mmdet'] = mmdet
    sys.modules['mmdet.models'] = models
    sys.modules['mmdet.models.detectors'] = detectors

    mmdet3d = types.ModuleType('mmdet3d')
    mmdet3d.__path__ = []
    core = types.ModuleType('mmdet3d.core')
    core.bbox3d2result = lambda *a, **k: None
    mmdet3d.core = core
    sys.modules['mmdet3d'] = mmdet3d
    sys.modules['mmdet3d.core'] = core


def load():
    """Returns a namespace with the reference modules:
    ``projection``, ``render_ray``, ``nerf_mlp``, ``nerfdet`` (module objects)."""
    if 'ns' in _cache:
        return _cache['ns']
    if not available():
        raise FileNotFoundError(f'reference not found under {REFERENCE_ROOT}')
    _install_stubs()
    pkg_models = types.ModuleType('ndref.models')
    pkg_models.__path__ = []
    pkg_mu = types.ModuleType('ndref.models.model_utils')
    pkg_mu.__path__ = []
    pkg_det = types.ModuleType('ndref.models.detectors')
    pkg_det.__path__ = []
    root = types.ModuleType('ndref')
    root.__path__ = []
    sys.modules['ndref'] = root
    sys.modules['ndref.models'] = pkg_models
    sys.modules['ndref.models.model_utils'] = pkg_mu
    sys.modules['ndref.models.detectors'] = pkg_det

    projection = _load('ndref.models.model_utils.projection', f'{_MU}/projection.py')
    render_ray = _load('ndref.models.model_utils.render_ray', f'{_MU}/render_ray.py')
    nerf_mlp = _load('ndref.models.model_utils.nerf_mlp', f'{_MU}/nerf_mlp.py')
    sri = types.ModuleType('ndref.models.model_utils.save_rendered_img')
    sri.save_rendered_img = lambda *a, **k: (0.0, 0.0, 0.0)
    sys.modules['ndref.models.model_utils.save_rendered_img'] = sri
    nerfdet = _load('ndref.models.detectors.nerfdet', 'mmdet3d/models/detectors/nerfdet.py',
                    package='ndref.models.detectors')
    ns = types.SimpleNamespace(projection=projection, render_ray=render_ray,
                               nerf_mlp=nerf_mlp, nerfdet=nerfdet)
    _cache['ns'] = ns
    return ns


class _FeatureBackbone(nn.Module):
    """Stand-in backbone: ignores the image and returns pre-made features so the
    reference ``extract_feat`` runs its own lifting code on chosen inputs."""

    def __init__(self, feats):
        super().__init__()
        self.feats = feats

    def init_weights(self, *a, **k):
        pass

    def forward(self, img):
        return [self.feats]


class _Identity(nn.Module):
    def init_weights(self, *a, **k):
        pass

    def forward(self, x):
        return x


class _DummyHead(nn.Module):
    voxel_size = None

    def init_weights(self, *a, **k):
        pass


def build_reference_detector(features, n_voxels, voxel_size, aabb, near_far_range,
                             N_samples, N_rand, out_channels=256, **kw):
    """Instantiates the reference ``nerfdet`` class around fixed features
    ``[nv, C, Hf, Wf]`` (batch of one scene)."""
    ns = load()
    det = ns.nerfdet.nerfdet(
        backbone=dict(module=_FeatureBackbone(features)),
        neck=dict(module=_Identity(), out_channels=out_channels),
        neck_3d=dict(module=_Identity()),
        bbox_head=dict(module=_DummyHead()),
        n_voxels=n_voxels, voxel_size=voxel_size, aabb=aabb,
        near_far_range=near_far_range, N_samples=N_samples, N_rand=N_rand,
        nerf_mode='image', nerf_density=True, **kw)
    return det
