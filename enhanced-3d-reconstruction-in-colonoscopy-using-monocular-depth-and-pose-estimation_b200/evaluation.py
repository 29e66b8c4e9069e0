"""Drop-in for the hot-path functions of the reference's ``eval/evaluation.py``:
``compute_errors`` (:16-60) and ``compose_poses`` (:279-382), computed by libdav2_b200.so.

Extra (not in the reference): ``test_step_metrics`` fuses the mask of lightning_model.py:304-306
with compute_errors in ONE pass over pred/gt (no boolean-mask compaction, no per-metric sync), and
``metric_partials`` / ``finalize_compute_errors`` expose the partial sums so that multi-GPU runs can
all-reduce sums and finalise afterwards."""
from __future__ import annotations

import logging
from typing import Optional

import torch

from . import ops

logger = logging.getLogger(__name__)


def finalize_compute_errors(partials: torch.Tensor) -> dict:
    """partials fp64 [8] -> {"d1","abs_rel","rmse","l1"} as 0-d fp32 tensors (same device, no sync)."""
    p = partials.to(torch.float64)
    n = p[0]
    return {
        "d1": (p[5] / n).to(torch.float32),
        "abs_rel": (p[2] / n).to(torch.float32),
        "rmse": torch.sqrt(p[3] / n).to(torch.float32),
        "l1": (p[1] / n).to(torch.float32),
    }


def metric_partials(pred: torch.Tensor, gt: torch.Tensor, min_depth: float = 1e-6, max_depth: float = 20.0,
                    per_frame: bool = False) -> torch.Tensor:
    """Masked partial sums of test_step (valid = min_depth <= gt <= max_depth); see include/dav2_b200.h."""
    assert pred.shape == gt.shape
    return ops.depth_metric_partials(pred.float().contiguous(), gt.float().contiguous(), min_depth, max_depth, 0, per_frame)


def test_step_metrics(pred: torch.Tensor, gt: torch.Tensor, min_depth: float = 1e-6, max_depth: float = 20.0) -> dict:
    """lightning_model.py:304-313 in one kernel: batch-wide mask + compute_errors."""
    return finalize_compute_errors(metric_partials(pred, gt, min_depth, max_depth))


test_step_metrics.__test__ = False


def compute_errors(pred: torch.Tensor, gt: torch.Tensor, check_finite: bool = True) -> dict:
    """eval/evaluation.py:16-60.  ``pred`` / ``gt`` are the already-masked 1-D tensors the reference passes."""
    assert pred.shape == gt.shape
    part = ops.depth_metric_partials(pred.detach().float().contiguous().view(1, -1),
                                     gt.detach().float().contiguous().view(1, -1), 0.0, 0.0, 2, False)
    if check_finite:  # the reference's two .any() syncs, folded into one read
        n_nan, n_inf = part[6:8].tolist()
        if n_nan:
            logger.warning("NaN values detected in predictions")
        if n_inf:
            logger.warning("Inf values detected in predictions")
    return finalize_compute_errors(part)


def compose_poses(relative_poses: torch.Tensor, initial_pose: Optional[torch.Tensor] = None) -> torch.Tensor:
    """eval/evaluation.py:279-382: [N,7] (or [b,N,7] -> batch 0 only, or [7]) -> absolute poses [N+1,7]."""
    rel = relative_poses
    if rel.dim() == 3:
        rel = rel[0]
    if rel.dim() == 1:
        rel = rel.unsqueeze(0)
    if initial_pose is not None and initial_pose.dim() > 1:
        initial_pose = initial_pose.squeeze()
    return ops.compose_poses(rel.detach().float().contiguous(), initial_pose)


def poses_to_transforms(abs_poses: torch.Tensor) -> torch.Tensor:
    """[n,7] absolute poses -> fp64 [n,12] rows of [R|t] (depth_to_pointcloud.py:168-173 semantics)."""
    n = abs_poses.shape[0]
    zero_rel = torch.zeros(0, 7, dtype=torch.float32, device=abs_poses.device)
    out = []
    for i in range(n):  # tiny: used for parity tests; bulk path is ops.compose_poses(want_T12=True)
        _, T = ops.compose_poses(zero_rel, abs_poses[i], want_T12=True)
        out.append(T[0])
    return torch.stack(out)
