"""TEST INFRASTRUCTURE -- CPU oracle of row N3 (ray generation, de-normalised images).  NOT product code.

``raydirs`` follows ``mmdet3d/datasets/pipelines/multi_view.py:124-132`` and ``data_augment_utils.py:410-424``
(``get_dtu_raydir``); it is pinned against the reference's own ``get_dtu_raydir`` run in the build container
(``oracle/make_golden.py:gen_rays``: the function is taken out of the unmodified source file, whose module imports mmcv).
``denorm`` follows ``multi_view.py:107-110``: ``mmcv.imdenormalize`` (mmcv 1.x, ``mmcv/image/photometric.py``; the package
is absent here) is ``cv2.multiply(img, std); cv2.add(img, mean, img); cv2.cvtColor(img, RGB2BGR, img)`` with mean / std as
float64 ``(1, 3)`` arrays -- restated in plain numpy below and pinned against cv2 itself (``gen_rays`` stores cv2's output).
Parity status: pinned by reference- / OpenCV-generated fixtures (tests/golden/rays_small.npz)."""
from __future__ import annotations

import numpy as np


def raydirs(intrinsics_nerf: np.ndarray, camrotc2w: np.ndarray, height: int, width: int, margin: int) -> np.ndarray:
    """[(H - 2m) * (W - 2m), 3] float32 (multi_view.py:124-132, data_augment_utils.py:410-424)."""
    px, py = np.meshgrid(np.arange(margin, width - margin).astype(np.float32),
                         np.arange(margin, height - margin).astype(np.float32))
    pixelcoords = np.stack((px, py), axis=-1).astype(np.float32)
    x = (pixelcoords[..., 0] + 0.5 - intrinsics_nerf[0, 2]) / intrinsics_nerf[0, 0]
    y = (pixelcoords[..., 1] + 0.5 - intrinsics_nerf[1, 2]) / intrinsics_nerf[1, 1]
    z = np.ones_like(x)
    dirs = np.stack([x, y, z], axis=-1) @ camrotc2w[:, :].T
    return np.reshape(dirs.astype(np.float32), (-1, 3))


def denorm(img_hwc: np.ndarray, mean, std, to_bgr: bool = True) -> np.ndarray:
    """``imdenormalize(img, mean, std, to_bgr).astype(uint8) / 255.0`` for one float32 HWC image: the product in float64
    rounded to float32, the sum in float32 (OpenCV's arithmetic for a 32F matrix and a float64 scalar)."""
    std = np.asarray(std, dtype=np.float64).reshape(1, 1, 3)
    mean32 = np.asarray(mean, dtype=np.float64).reshape(1, 1, 3).astype(np.float32)
    v = (img_hwc.astype(np.float64) * std).astype(np.float32)
    v = (v + mean32).astype(np.float32)
    if to_bgr:
        v = v[..., ::-1]
    return v.astype(np.int32).astype(np.uint8) / 255.0
