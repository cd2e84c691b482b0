"""The shared NeRF / geometry MLP with the reference's module interface
(``mmdet3d/models/model_utils/nerf_mlp.py:200-234``, instantiated at ``nerfdet.py:62-69``).

``VanillaNeRFRadianceField`` keeps the reference constructor, ``forward(x, condition, features)``,
``query_density(x, features)`` and -- through an identical sub-module tree -- the reference
``state_dict`` keys (``mlp.base.hidden_layers.<i>.weight`` ...), so reference checkpoints load
unchanged.  The arithmetic runs on the tcgen05 tensor cores (csrc/mlp_tc.cu): ``precision='fp32'`` (default) carries
every operand as a hi + lo pair of bf16 numbers with fp32 accumulation in tensor memory (``nd_nerf_mlp_fwd_tc3``, 1e-4),
``precision='bf16'`` uses plain bf16 operands (``nd_nerf_mlp_fwd_tc``, 1e-2); ``precision='fp32_ffma'`` is the FFMA kernel
(``nd_nerf_mlp_fwd``, csrc/mlp.cu, 1e-4) that also takes the architectures the tensor-core kernel does not.  There is no
eager fallback.

Autograd (row N1 of SURVEY.md section 8f): the forward is always the CUDA kernel; when an input or a parameter requires
grad, the backward re-evaluates the network's fp32 formula with torch ops under autograd (``_torch_math``: its GEMMs go
to cuBLAS -- library GEMMs, not a kernel of this repository; dgrad / wgrad on tcgen05 are not built) and returns the
gradients of that formula, so the module trains as a drop-in.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def _encode(x: torch.Tensor, n_octaves: int) -> torch.Tensor:
    """``SinusoidalEncoder`` (nerf_mlp.py:181-197): [x, sin(2^k x), sin(2^k x + pi / 2)], k outer, xyz inner."""
    scales = torch.tensor([2 ** k for k in range(n_octaves)], device=x.device, dtype=x.dtype)
    scaled = (x[..., None, :] * scales[:, None]).reshape(*x.shape[:-1], n_octaves * x.shape[-1])
    return torch.cat([x, torch.sin(torch.cat([scaled, scaled + 0.5 * math.pi], dim=-1))], dim=-1)


class _FieldFn(torch.autograd.Function):
    """Forward: the CUDA kernel.  Backward: torch autograd through ``field._torch_math`` on the saved inputs."""

    @staticmethod
    def forward(ctx, field, want_rgb, spr, x, cond, feats, *params):
        sigma, rgb, _ = ops.direct.nerf_mlp_fwd(field.packed_weights(), field.dims, x, feats, cond, spr, want_rgb, False,
                                                field.precision)
        ctx.field, ctx.want_rgb, ctx.spr, ctx.has_cond = field, want_rgb, spr, cond is not None
        ctx.save_for_backward(x, cond if cond is not None else x.new_empty(0), feats)
        if not want_rgb:
            rgb = x.new_empty((0, 3))
            ctx.mark_non_differentiable(rgb)
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        field = ctx.field
        x, cond, feats = ctx.saved_tensors
        params = list(field._weights().values())
        need = ctx.needs_input_grad[3:]
        with torch.enable_grad():
            leaves = [x.detach().requires_grad_(need[0]),
                      cond.detach().requires_grad_(bool(need[1])) if ctx.has_cond else None,
                      feats.detach().requires_grad_(need[2])]
            sigma, rgb = field._torch_math(leaves[0], leaves[1], leaves[2], ctx.spr, ctx.want_rgb)
            outs, gouts = [sigma], [g_sigma.reshape(sigma.shape)]
            if ctx.want_rgb:
                outs.append(rgb)
                gouts.append(g_rgb.reshape(rgb.shape))
            cands = leaves + params
            wanted = [t for t, n in zip(cands, need) if n and t is not None]
            grads = iter(torch.autograd.grad(outs, wanted, gouts, allow_unused=True)) if wanted else iter(())
        result = [next(grads) if (n and t is not None) else None for t, n in zip(cands, need)]
        return (None, None, None) + tuple(result)


class _Linears(nn.Module):
    """Container with the attribute names of the reference ``MLP`` (nerf_mlp.py:11-90):
    ``hidden_layers`` (ModuleList of Linear) and, when enabled, ``output_layer``."""

    def __init__(self, input_dim: int, output_dim: Optional[int], net_depth: int, net_width: int,
                 skip_layer: Optional[int], output_enabled: bool = True):
        super().__init__()
        self.hidden_layers = nn.ModuleList()
        in_features = input_dim
        for i in range(net_depth):
            self.hidden_layers.append(nn.Linear(in_features, net_width))
            if skip_layer is not None and i % skip_layer == 0 and i > 0:
                in_features = net_width + input_dim
            else:
                in_features = net_width
        if output_enabled:
            self.output_layer = nn.Linear(in_features, output_dim)
            self.output_dim = output_dim
        else:
            self.output_dim = in_features
        for m in self.modules():                     # nerf_mlp.py:60-78: xavier-uniform weights, zero biases
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)


class _NerfMLP(nn.Module):
    """Sub-module names of the reference ``NerfMLP`` (nerf_mlp.py:103-161)."""

    def __init__(self, input_dim, condition_dim, feature_dim, net_depth, net_width, skip_layer,
                 net_depth_condition, net_width_condition):
        super().__init__()
        self.base = _Linears(input_dim + feature_dim, None, net_depth, net_width, skip_layer, output_enabled=False)
        hidden = self.base.output_dim
        self.sigma_layer = _Linears(hidden, 1, 0, net_width, None)
        self.bottleneck_layer = _Linears(hidden, net_width, 0, net_width, None)
        self.rgb_layer = _Linears(net_width + condition_dim, 3, net_depth_condition, net_width_condition, None)


class _Encoder(nn.Module):
    """Holds the ``scales`` buffer of the reference ``SinusoidalEncoder`` (nerf_mlp.py:164-197); the encoding
    itself is evaluated inside the kernel."""

    def __init__(self, x_dim: int, min_deg: int, max_deg: int):
        super().__init__()
        self.x_dim, self.min_deg, self.max_deg = x_dim, min_deg, max_deg
        self.register_buffer('scales', torch.tensor([2 ** i for i in range(min_deg, max_deg)]))

    @property
    def latent_dim(self) -> int:
        return (1 + (self.max_deg - self.min_deg) * 2) * self.x_dim


class VanillaNeRFRadianceField(nn.Module):
    def __init__(self, net_depth: int = 8, net_width: int = 256, skip_layer: int = 4, feature_dim: int = 0,
                 net_depth_condition: int = 1, net_width_condition: int = 128, *, precision: str = 'fp32') -> None:
        super().__init__()
        if precision not in ('fp32', 'bf16', 'fp32_ffma'):
            raise ValueError("precision must be 'fp32', 'bf16' or 'fp32_ffma'")
        self.precision = precision                 # may be switched at any time; the packed copy follows
        if net_depth_condition != 1:
            raise NotImplementedError('net_depth_condition != 1 is not used by NeRF-Det and not built')
        self.posi_encoder = _Encoder(3, 0, 10)
        self.view_encoder = _Encoder(3, 0, 4)
        self.mlp = _NerfMLP(self.posi_encoder.latent_dim, self.view_encoder.latent_dim, feature_dim, net_depth,
                            net_width, skip_layer, net_depth_condition, net_width_condition)
        self.dims = [net_depth, net_width, skip_layer or 0, feature_dim, net_width_condition, 10, 4]
        # 'fp32' runs on the tensor cores with split operands; architectures that kernel does not take fall to the FFMA one
        if precision == 'fp32' and not ops.mlp_precision_supported(self.dims, 'fp32'):
            self.precision = 'fp32_ffma'
        self._packed = None
        self._packed_key = None

    # ---- weights ---------------------------------------------------------------------------
    def _weights(self):
        m = self.mlp
        w = {}
        for i, lin in enumerate(m.base.hidden_layers):
            w[f'base_w{i}'], w[f'base_b{i}'] = lin.weight, lin.bias
        w['sigma_w'], w['sigma_b'] = m.sigma_layer.output_layer.weight, m.sigma_layer.output_layer.bias
        w['bottleneck_w'], w['bottleneck_b'] = (m.bottleneck_layer.output_layer.weight,
                                                m.bottleneck_layer.output_layer.bias)
        w['rgb_hidden_w'], w['rgb_hidden_b'] = m.rgb_layer.hidden_layers[0].weight, m.rgb_layer.hidden_layers[0].bias
        w['rgb_out_w'], w['rgb_out_b'] = m.rgb_layer.output_layer.weight, m.rgb_layer.output_layer.bias
        return w

    def packed_weights(self) -> torch.Tensor:
        """Kernel-layout copy of the weights, rebuilt whenever a parameter was modified or moved."""
        w = self._weights()
        key = (self.precision,) + tuple((t.data_ptr(), t._version) for t in w.values())
        if self._packed is None or key != self._packed_key:
            self._packed = ops.pack_mlp_weights({k: t.detach().contiguous() for k, t in w.items()}, self.dims,
                                                self.precision)
            self._packed_key = key
        return self._packed

    # ---- the fp32 formula in torch ops (backward only) -----------------------------------------
    def _torch_math(self, x, cond, feats, spr: int, want_rgb: bool):
        """(relu(sigma) [P, 1], sigmoid(rgb) [P, 3] or None) exactly as nerf_mlp.py:80-161, 224-234 chains them."""
        m, skip = self.mlp, self.dims[2]
        inp = torch.cat([_encode(x, 10), feats], dim=-1) if feats.shape[-1] else _encode(x, 10)
        h = inp
        for i, lin in enumerate(m.base.hidden_layers):
            h = torch.relu(F.linear(h, lin.weight, lin.bias))
            if skip and i % skip == 0 and i > 0:
                h = torch.cat([h, inp], dim=-1)
        sigma = torch.relu(F.linear(h, m.sigma_layer.output_layer.weight, m.sigma_layer.output_layer.bias))
        if not want_rgb:
            return sigma, None
        bott = F.linear(h, m.bottleneck_layer.output_layer.weight, m.bottleneck_layer.output_layer.bias)
        enc = _encode(cond, 4)
        if spr > 1:
            enc = enc.repeat_interleave(spr, dim=0)
        lin = m.rgb_layer.hidden_layers[0]
        t = torch.relu(F.linear(torch.cat([bott, enc], dim=-1), lin.weight, lin.bias))
        return sigma, torch.sigmoid(F.linear(t, m.rgb_layer.output_layer.weight, m.rgb_layer.output_layer.bias))

    def _needs_grad(self, *tensors) -> bool:
        return torch.is_grad_enabled() and (any(t is not None and t.requires_grad for t in tensors)
                                            or any(p.requires_grad for p in self._weights().values()))

    # ---- reference interface ---------------------------------------------------------------
    def query_density(self, x, features=None, return_alpha: bool = False):
        """relu(sigma) for points ``x [..., 3]`` with ``features [..., feature_dim]`` (nerf_mlp.py:224-227).
        ``return_alpha`` additionally returns 1 - exp(-sigma) (nerfdet.py:258) from the same launch."""
        lead = x.shape[:-1]
        feats = self._features(x, features)
        if self._needs_grad(x, feats):
            sigma, _ = _FieldFn.apply(self, False, 1, x.reshape(-1, 3), None, feats, *self._weights().values())
            sigma = sigma.view(*lead, 1)
            return (sigma, 1 - torch.exp(-sigma)) if return_alpha else sigma
        sigma, _, alpha = ops.direct.nerf_mlp_fwd(self.packed_weights(), self.dims, x.reshape(-1, 3), feats, None, 1, False,
                                           return_alpha, self.precision)
        sigma = sigma.view(*lead, 1)
        return (sigma, alpha.view(*lead, 1)) if return_alpha else sigma

    def forward(self, x, condition=None, features=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(sigmoid(rgb) [..., 3], relu(sigma) [..., 1]) (nerf_mlp.py:229-234).  ``condition`` is either
        per point (``x.shape[:-1] + (3,)``) or per ray ``[num_rays, 3]`` broadcast over the remaining
        leading dimensions of ``x`` (nerf_mlp.py:153-157)."""
        if condition is None:
            raise ValueError('the colour branch needs `condition` (ray directions); use query_density for sigma only')
        lead = x.shape[:-1]
        feats = self._features(x, features)
        p = feats.shape[0]
        if condition.shape[:-1] == lead:
            cond, spr = condition.reshape(-1, 3), 1
        else:
            if condition.dim() != 2 or condition.shape[0] != lead[0]:
                raise ValueError(f'condition {tuple(condition.shape)} does not broadcast over x {tuple(x.shape)}')
            cond, spr = condition, p // condition.shape[0]
        if self._needs_grad(x, cond, feats):
            sigma, rgb = _FieldFn.apply(self, True, spr, x.reshape(-1, 3), cond, feats, *self._weights().values())
            return rgb.view(*lead, 3), sigma.view(*lead, 1)
        sigma, rgb, _ = ops.direct.nerf_mlp_fwd(self.packed_weights(), self.dims, x.reshape(-1, 3), feats, cond, spr, True, False,
                                         self.precision)
        return rgb.view(*lead, 3), sigma.view(*lead, 1)

    def _features(self, x, features):
        fd = self.dims[3]
        if fd == 0:
            return x.new_empty((x.numel() // 3, 0))
        if features is None:
            raise ValueError(f'this field was built with feature_dim={fd}: `features` is required')
        return features.reshape(-1, fd)
