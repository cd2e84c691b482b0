"""TEST INFRASTRUCTURE -- CPU oracle for the shared NeRF / geometry MLP.  NOT product code.

Restates ``VanillaNeRFRadianceField`` (reference
``mmdet3d/models/model_utils/nerf_mlp.py:11-234``, instantiated at
``mmdet3d/models/detectors/nerfdet.py:62-69``) as explicit matmuls over the
reference ``state_dict`` keys.  Checked against the reference module itself in
``tests/test_oracle_golden.py`` (build container) and against fixtures made by
``oracle/make_golden.py``.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def sinusoidal_encode(x: torch.Tensor, n_octaves: int) -> torch.Tensor:
    """[x, sin(x*2^k) (k outer, xyz inner), sin(x*2^k + pi/2)]  (nerf_mlp.py:181-197).
    The cosine half is sin of the fp32-rounded sum."""
    if n_octaves == 0:
        return x
    scales = torch.tensor([2 ** k for k in range(n_octaves)], device=x.device)
    scaled = (x[..., None, :] * scales[:, None]).reshape(*x.shape[:-1], n_octaves * x.shape[-1])
    both = torch.cat([scaled, scaled + 0.5 * math.pi], dim=-1)
    return torch.cat([x, torch.sin(both)], dim=-1)


class FieldOracle:
    """Callable like the reference module: ``field(x, condition, features)`` ->
    (rgb, sigma) and ``field.query_density(x, features)`` -> sigma."""

    def __init__(self, state: Dict[str, torch.Tensor], net_depth: int = 4):
        self.s = {k: v for k, v in state.items()}
        self.net_depth = net_depth

    def _lin(self, key, x):
        return F.linear(x, self.s[key + '.weight'], self.s[key + '.bias'])

    def _trunk(self, x, features):
        inp = torch.cat([sinusoidal_encode(x, 10), features], dim=-1)
        h = inp
        for i in range(self.net_depth):
            h = torch.relu(self._lin(f'mlp.base.hidden_layers.{i}', h))
        # skip connection is appended after the last hidden layer (skip_layer=3,
        # net_depth=4: i % 3 == 0 and i > 0 only for i = 3; nerf_mlp.py:85-86)
        return torch.cat([h, inp], dim=-1)

    def query_density(self, x, features):
        h = self._trunk(x, features)
        return torch.relu(self._lin('mlp.sigma_layer.output_layer', h))

    def __call__(self, x, condition, features):
        h = self._trunk(x, features)
        sigma = torch.relu(self._lin('mlp.sigma_layer.output_layer', h))
        bott = self._lin('mlp.bottleneck_layer.output_layer', h)
        cond = sinusoidal_encode(condition, 4)
        if cond.shape[:-1] != bott.shape[:-1]:
            cond = cond.view([cond.shape[0]] + [1] * (bott.dim() - cond.dim()) + [cond.shape[-1]]
                             ).expand(*bott.shape[:-1], cond.shape[-1])
        t = torch.relu(self._lin('mlp.rgb_layer.hidden_layers.0', torch.cat([bott, cond], dim=-1)))
        rgb = torch.sigmoid(self._lin('mlp.rgb_layer.output_layer', t))
        return rgb, sigma


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class FieldOracleBf16(FieldOracle):
    """The same network with the operand roundings of the tensor-core kernel (nerfdet_b200/csrc/mlp_tc.cu): inputs,
    weights and hidden activations rounded to bf16 where they enter a matrix product, products accumulated in fp32;
    the sigma row and the 128 -> 3 output layer stay fp32 on the un-rounded rows.  Used by the GPU tests to tell a
    layout / pipeline bug (differs from this by > 1e-3) from bf16 rounding (differs from ``FieldOracle`` by ~1e-2)."""

    def _lin16(self, key, x):
        return F.linear(_bf16(x), _bf16(self.s[key + '.weight']), self.s[key + '.bias'])

    def _trunk16(self, x, features):
        inp = torch.cat([sinusoidal_encode(x, 10), features], dim=-1)
        h = inp
        for i in range(self.net_depth):
            h = torch.relu(self._lin16(f'mlp.base.hidden_layers.{i}', h))
        return h, inp

    def query_density(self, x, features):
        h, inp = self._trunk16(x, features)
        return torch.relu(self._lin('mlp.sigma_layer.output_layer', torch.cat([h, inp], dim=-1)))

    def __call__(self, x, condition, features):
        h, inp = self._trunk16(x, features)
        cat = torch.cat([h, inp], dim=-1)
        sigma = torch.relu(self._lin('mlp.sigma_layer.output_layer', cat))
        bott = self._lin16('mlp.bottleneck_layer.output_layer', cat)
        cond = sinusoidal_encode(condition, 4)
        if cond.shape[:-1] != bott.shape[:-1]:
            cond = cond.view([cond.shape[0]] + [1] * (bott.dim() - cond.dim()) + [cond.shape[-1]]
                             ).expand(*bott.shape[:-1], cond.shape[-1])
        t = torch.relu(self._lin16('mlp.rgb_layer.hidden_layers.0', torch.cat([bott, cond], dim=-1)))
        rgb = torch.sigmoid(self._lin('mlp.rgb_layer.output_layer', t))
        return rgb, sigma
