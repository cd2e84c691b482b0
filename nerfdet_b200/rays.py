"""Row N3 of SURVEY.md section 8f: the producers of the render branch's inputs, on the GPU.

The reference builds the rays of the target views and the de-normalised source images in its CPU data pipeline
(``mmdet3d/datasets/pipelines/multi_view.py:107-132``, ``data_augment_utils.py:410-424``, ``formating.py:70-91``) and ships
them to the device with every sample: 660 000 rays x 2 tensors and a second copy of the images.  Both are functions of
data the device already has (camera matrices; the normalised network input), so they are computed where they are used."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

from . import ops
from .lifting import to_device


def nerf_intrinsics(img_meta) -> torch.Tensor:
    """``intrinsics_nerf`` of multi_view.py:117-118: the intrinsic matrix with rows 0-1 divided by ``ori_h / img_h``
    (host float32, like the reference's numpy array)."""
    k = np.array(img_meta['lidar2img']['intrinsic'], dtype=np.float32, copy=True)
    ratio = img_meta['ori_shape'][0] / img_meta['img_shape'][0]
    k[:2] = k[:2] / ratio
    return torch.from_numpy(k)


def generate_ray_batch(img_meta, camrotc2w, lightpos, height: int, width: int, margin: int, device) -> Dict[str, torch.Tensor]:
    """``ray_o`` / ``ray_d`` ``[1, nt, n_pix, 3]`` float32 and ``nerf_sizes [1, nt, 3]`` as the collated batch holds them
    (multi_view.py:124-132, 149; formating.py:55-75), for ``nt`` target cameras given by their camera-to-world rotations
    ``camrotc2w [nt, 3, 3]`` and centres ``lightpos [nt, 3]`` (numpy or torch, any float dtype)."""
    rot = to_device(torch.as_tensor(np.asarray(camrotc2w), dtype=torch.float64), device)
    pos = to_device(torch.as_tensor(np.asarray(lightpos), dtype=torch.float32), device)
    ray_d, ray_o = ops.direct.generate_rays(nerf_intrinsics(img_meta), rot, pos, int(height), int(width), int(margin))
    nt = rot.shape[0]
    sizes = torch.tensor([[height - 2 * margin, width - 2 * margin, 3]] * nt).unsqueeze(0)
    return {'ray_o': ray_o.unsqueeze(0), 'ray_d': ray_d.unsqueeze(0), 'nerf_sizes': sizes}


def denorm_images(img: torch.Tensor, mean: Sequence[float], std: Sequence[float], to_bgr: bool = True) -> torch.Tensor:
    """``denorm_images [n, 3, H, W]`` in [0, 1] from the normalised network input ``img [n, 3, H, W]``
    (multi_view.py:107-110 + formating.py:87-91: ``imdenormalize(...).astype(uint8) / 255``, quantisation included)."""
    return ops.direct.denorm_images(img, [float(v) for v in mean], [float(v) for v in std], bool(to_bgr))
