"""Row N4 of SURVEY.md section 8f: the evaluator of the ``render_testing`` branch
(``mmdet3d/models/model_utils/save_rendered_img.py``) as a batched GPU pass.

``render.render_rays(..., render_testing=True)`` renders the full target images chunk by chunk on the device; the
reference then pulls every view to the host, computes PSNR in torch, SSIM in scikit-image and writes annotated PNGs
(``:40-78``).  Here the metrics of all views come from one launch; the PNG writing (cv2 / file system) is not part of the
path."""
from __future__ import annotations

import torch

from . import ops

FLOAT_DATA_RANGE = 2.0      # scikit-image 0.18.1 (the reference's pin): data_range of a float image = its dtype range (-1, 1)


def image_metrics(pred: torch.Tensor, target: torch.Tensor, data_range: float = FLOAT_DATA_RANGE) -> torch.Tensor:
    """float64 ``[nv, 2]`` = {PSNR, SSIM} per view for ``pred``, ``target`` ``[nv, H, W, 3]`` on the device."""
    return ops.direct.image_metrics(pred.float(), target if target.dtype == torch.float64 else target.float(), float(data_range))


def compute_psnr(pred, target, mask=None):
    """``compute_psnr`` (save_rendered_img.py:13-20) for one view ``[H, W, 3]``; a numpy scalar like the reference's."""
    if mask is not None:
        raise NotImplementedError('masked PSNR is never used by NeRF-Det (save_rendered_img.py:66 passes mask=None)')
    return image_metrics(pred.unsqueeze(0), target.unsqueeze(0))[0, 0].cpu().numpy()


def compute_ssim(pred, target, mask=None):
    """``compute_ssim`` (save_rendered_img.py:22-38) for one view ``[H, W, 3]``."""
    if mask is not None:
        raise NotImplementedError('masked SSIM is never used by NeRF-Det (save_rendered_img.py:68 passes mask=None)')
    return float(image_metrics(pred.unsqueeze(0), target.unsqueeze(0))[0, 1])


def evaluate_rendered(rendered_results):
    """What ``save_rendered_img`` returns -- (mean PSNR, mean SSIM, mean over the views of the squared depth error map) --
    for ``rendered_results`` as ``render_rays(render_testing=True)`` produces them (the last entry counts, like the
    reference's loop at :45-49), without the image files."""
    ret = rendered_results[-1]
    rgb, depth = ret['outputs_coarse']['rgb'], ret['outputs_coarse']['depth']
    gt, gt_depth = ret['gt_rgb'].to(rgb.device), ret['gt_depth'].to(rgb.device)
    m = image_metrics(rgb, gt)
    rsme = ops.direct.depth_sqerr(depth.float(), gt_depth if gt_depth.dtype == torch.float64 else gt_depth.float())
    return float(m[:, 0].mean()), float(m[:, 1].mean()), rsme.cpu().numpy()
