"""Drop-in for ``calculate_metrics.calculate_metrics`` (reference calculate_metrics.py:17-55)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops

_KEYS = ("rmse", "mae", "abs_rel", "sq_rel", "delta1", "delta2", "delta3")


def finalize_calculate_metrics(p) -> dict:
    """fp64 partial sums [8] (host) -> the reference's dict; NaN dict when nothing is valid."""
    n = float(p[0])
    if n == 0:
        return {k: np.nan for k in _KEYS}
    mae, msq, mgt = p[1] / n, p[3] / n, p[4] / n
    return {"rmse": float(np.sqrt(msq)), "mae": float(mae), "abs_rel": float(mae / mgt), "sq_rel": float(msq / mgt),
            "delta1": float(p[5] / n), "delta2": float(p[6] / n), "delta3": float(p[7] / n)}


def _to_dev(a, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def calculate_metrics_batch(gt, pred, mask_invalid: bool = True, device="cuda"):
    """Per-frame metrics for [B,H,W] stacks in one launch -> list of dicts."""
    g, p = _to_dev(gt, device), _to_dev(pred, device)
    assert g.shape == p.shape and g.dim() == 3
    part = ops.depth_metric_partials(p, g, 0.0, 0.0, 1 if mask_invalid else 3, True).cpu().numpy()
    return [finalize_calculate_metrics(row) for row in part]


def calculate_metrics(gt, pred, mask_invalid: bool = True, device="cuda") -> dict:
    """gt, pred: [H,W] numpy arrays or tensors (metres)."""
    g, p = _to_dev(gt, device), _to_dev(pred, device)
    return calculate_metrics_batch(g.unsqueeze(0), p.unsqueeze(0), mask_invalid, device)[0]
