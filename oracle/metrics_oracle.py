"""CPU oracle (numpy) for the two depth-metric definitions on the hot path.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.

* ``compute_errors``     restates eval/evaluation.py:16-60 (torch): l1 = mean|d|;
                         abs_rel = mean(|d|/(gt+1e-6)); rmse = sqrt(mean d^2);
                         d1 = mean(max(gt/pred, pred/gt) < 1.1).  No positivity guard.
* ``test_step_metrics``  restates lightning_model.py:301-313: valid = (gt>=min_depth)&(gt<=max_depth)
                         over the WHOLE batch, then compute_errors on the masked, flattened pixels.
* ``calculate_metrics``  restates calculate_metrics.py:17-55 (numpy): valid=(gt>0)&(pred>0)&~inf;
                         delta<1.25^k; rmse; mae; abs_rel = mean|d|/mean(gt); sq_rel = mean d^2/mean(gt);
                         NaN dict when nothing is valid.

Sums are accumulated in float64 (the reference uses fp32 pairwise sums; the difference is ~1e-7
relative, far inside the 1e-4 gate).  PINNED against the reference's own functions executed in the
build container: tests/golden/metrics_*.npz (generator scripts/make_golden.py) and the spot values
recorded in SURVEY.md section 8c.
"""
from __future__ import annotations

import numpy as np

PARTIAL_FIELDS = ("n", "sum_abs", "sum_absrel_eps", "sum_sq", "sum_gt", "n_d_a", "n_d_b", "n_d_c")


def compute_errors(pred, gt) -> dict:
    pred = np.asarray(pred, dtype=np.float32).reshape(-1)
    gt = np.asarray(gt, dtype=np.float32).reshape(-1)
    assert pred.shape == gt.shape
    diff = (pred - gt).astype(np.float32)
    ad = np.abs(diff)
    with np.errstate(divide="ignore", invalid="ignore"):
        thresh = np.maximum(gt / pred, pred / gt)
        rel = ad / (gt + np.float32(1e-6))
    n = max(pred.size, 1)
    return {
        "d1": float((thresh < np.float32(1.1)).sum(dtype=np.float64) / n),
        "abs_rel": float(rel.sum(dtype=np.float64) / n),
        "rmse": float(np.sqrt((diff.astype(np.float64) ** 2).sum() / n)),
        "l1": float(ad.sum(dtype=np.float64) / n),
    }


def test_step_metrics(pred, gt, min_depth: float = 1e-6, max_depth: float = 20.0) -> dict:
    pred = np.asarray(pred, dtype=np.float32)
    gt = np.asarray(gt, dtype=np.float32)
    m = (gt >= np.float32(min_depth)) & (gt <= np.float32(max_depth))
    return compute_errors(pred[m], gt[m])


test_step_metrics.__test__ = False  # not a pytest test


def calculate_metrics(gt, pred, mask_invalid: bool = True) -> dict:
    gt = np.asarray(gt, dtype=np.float32)
    pred = np.asarray(pred, dtype=np.float32)
    if mask_invalid:
        m = (gt > 0) & (pred > 0) & (~np.isinf(gt)) & (~np.isinf(pred))
        gt, pred = gt[m], pred[m]
    if gt.size == 0:
        return {k: float("nan") for k in ("rmse", "mae", "abs_rel", "sq_rel", "delta1", "delta2", "delta3")}
    n = gt.size
    with np.errstate(divide="ignore", invalid="ignore"):
        thresh = np.maximum(gt / pred, pred / gt)
    d = (gt - pred).astype(np.float32)
    mean_gt = gt.sum(dtype=np.float64) / n
    msq = (d.astype(np.float64) ** 2).sum() / n
    mae = np.abs(d).sum(dtype=np.float64) / n
    return {
        "rmse": float(np.sqrt(msq)),
        "mae": float(mae),
        "abs_rel": float(mae / mean_gt),
        "sq_rel": float(msq / mean_gt),
        "delta1": float((thresh < 1.25).sum(dtype=np.float64) / n),
        "delta2": float((thresh < 1.25 ** 2).sum(dtype=np.float64) / n),
        "delta3": float((thresh < 1.25 ** 3).sum(dtype=np.float64) / n),
    }
