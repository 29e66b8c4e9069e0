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


# ------------------------------------------------------------------------------------------------
# Pose metrics (eval/evaluation.py:63-254) -- trajectory-level scalars over N x 7 arrays: host numpy, like the
# reference; the pose CHAIN inside evaluate_trajectory runs on the GPU kernel when the inputs are CUDA tensors.
# ------------------------------------------------------------------------------------------------
def _np(a):
    import numpy as np
    return a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)


def _quat_to_matrix(q):
    import numpy as np
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quaternion_distance(q1, q2):
    """eval/evaluation.py:63-82: geodesic angle in degrees."""
    import numpy as np
    q1, q2 = _np(q1).astype(np.float64), _np(q2).astype(np.float64)
    q1, q2 = q1 / np.linalg.norm(q1), q2 / np.linalg.norm(q2)
    return float(np.degrees(2 * np.arccos(np.clip(np.abs(np.dot(q1, q2)), -1.0, 1.0))))


def compute_ate(gt_trans, pred_trans):
    """eval/evaluation.py:85-99."""
    import numpy as np
    e = np.linalg.norm(_np(gt_trans) - _np(pred_trans), axis=1)
    return np.sqrt(np.mean(e ** 2))


def compute_rte(gt_trans, pred_trans):
    """eval/evaluation.py:102-120."""
    import numpy as np
    return np.mean(np.linalg.norm(np.diff(_np(gt_trans), axis=0) - np.diff(_np(pred_trans), axis=0), axis=1))


def compute_rot_error(gt_rots, pred_rots):
    """eval/evaluation.py:123-161: mean angle of R_gt^T R_pred in degrees; zero predicted quaternion -> identity."""
    import numpy as np
    errs = []
    for q_gt, q_pred in zip(_np(gt_rots).astype(np.float64), _np(pred_rots).astype(np.float64)):
        if np.linalg.norm(q_pred) < 1e-8:
            q_pred = np.array([0.0, 0.0, 0.0, 1.0])
            logger.warning("Zero quaternion detected, using identity quaternion instead")
        q_gt, q_pred = q_gt / np.linalg.norm(q_gt), q_pred / np.linalg.norm(q_pred)
        r = _quat_to_matrix(q_gt).T @ _quat_to_matrix(q_pred)
        errs.append(np.degrees(np.arccos(np.clip((np.trace(r) - 1) / 2, -1.0, 1.0))))
    return np.mean(errs)


def compute_pose_errors(pred_positions: torch.Tensor, gt_positions: torch.Tensor) -> dict:
    """eval/evaluation.py:164-208."""
    import numpy as np
    p, g = _np(pred_positions).copy(), _np(gt_positions).copy()
    pq, gq = p[:, 3:], g[:, 3:]
    pq = pq / np.maximum(np.linalg.norm(pq, axis=1, keepdims=True), 1e-8)
    gq = gq / np.maximum(np.linalg.norm(gq, axis=1, keepdims=True), 1e-8)
    pq[np.sum(gq * pq, axis=1) < 0] *= -1
    return {"ate": torch.as_tensor(compute_ate(g[:, :3], p[:, :3])), "rte": torch.as_tensor(compute_rte(g[:, :3], p[:, :3])),
            "rote": torch.as_tensor(compute_rot_error(gq, pq))}


def calculate_scale_factor(pred_rel_poses, gt_rel_poses):
    """eval/evaluation.py:257-276."""
    pt, gt = pred_rel_poses[:, :3], gt_rel_poses[:, :3]
    return torch.sum(pt * gt) / torch.sum(pt * pt)


def evaluate_trajectory(pred_rel_poses: torch.Tensor, gt_rel_poses: torch.Tensor, initial_pose: Optional[torch.Tensor] = None):
    """eval/evaluation.py:211-254: scale-align, compose both chains (GPU kernel), ATE / RTE / ROT."""
    scale = calculate_scale_factor(pred_rel_poses, gt_rel_poses)
    scaled = pred_rel_poses.clone()
    scaled[:, :3] *= scale
    dev = scaled.device if scaled.is_cuda else torch.device("cuda")
    pred_abs = compose_poses(scaled.to(dev), None if initial_pose is None else initial_pose.to(dev))
    gt_abs = compose_poses(gt_rel_poses.to(dev), None if initial_pose is None else initial_pose.to(dev))
    return {"rte": compute_rte(scaled[:, :3], gt_rel_poses[:, :3]), "ate": compute_ate(gt_abs[:, :3], pred_abs[:, :3]),
            "rote": compute_rot_error(gt_abs[:, 3:], pred_abs[:, 3:])}


class RunningMeans:
    """Stand-in for the reference's ``MetricCollection`` of ``MeanMetric`` (lightning_model.py:143-150,
    pose_estimation_model.py:146-152): per key a running (sum, count) kept where the values live -- on the device for
    device tensors, so nothing synchronises until ``compute``."""

    def __init__(self, keys):
        self.keys = tuple(keys)
        self._acc: Optional[torch.Tensor] = None  # fp64 [len(keys)] sums
        self._n = 0

    def reset(self) -> None:
        self._acc, self._n = None, 0

    def update(self, metrics: dict) -> None:
        vals = [torch.as_tensor(metrics[k]).detach().to(torch.float64).reshape(()) for k in self.keys]
        dev = next((v.device for v in vals if v.is_cuda), vals[0].device)
        v = torch.stack([x.to(dev) for x in vals])
        self._acc = v.clone() if self._acc is None else self._acc + v.to(self._acc.device)
        self._n += 1

    def compute(self) -> dict:
        if self._acc is None:
            return {k: float("nan") for k in self.keys}
        return dict(zip(self.keys, (self._acc / self._n).tolist()))


# ------------------------------------------------------------------------------------------------
# Per-procedure aggregation of test_lightning.py:27-111 / :240-274 and pose_estimation_lightning.py:43-185 (host
# bookkeeping; no Lightning needed)
# ------------------------------------------------------------------------------------------------
DEPTH_KEYS = ("l1", "abs_rel", "d1", "rmse")
POSE_KEYS = ("ate", "rte", "rote")


class ProcedureMetricCollector:
    """The reference appends the BATCH metric once per frame to the frame's procedure bucket
    (test_lightning.py:76-109) and reports mean over procedures of per-procedure means (:244-274).
    ``keys=POSE_KEYS`` gives the pose test's collector (pose_estimation_lightning.py:43-185: same bucketing over
    ``batch["dataset"]`` / ``batch["id"]``, metrics ate / rte / rote)."""

    def __init__(self, keys=DEPTH_KEYS):
        from collections import defaultdict
        self.KEYS = tuple(keys)
        self.metrics_by_procedure = defaultdict(list)

    @staticmethod
    def procedure_of(dataset_path: str, frame_id: str):
        colon = None
        for part in str(dataset_path).split("/"):
            if part.startswith("SyntheticColon_"):
                colon = part
        proc = None
        for tag in ("S", "B", "O"):
            if tag in frame_id:
                proc = f"Frames_{tag}{frame_id.split(tag)[1].split('_')[0]}"
                break
        return None if colon is None or proc is None else f"{colon}/{proc}"

    def on_test_batch_end(self, outputs: dict, batch: dict):
        if not all(k in outputs for k in self.KEYS):
            raise ValueError("Missing expected keys in outputs")
        vals = torch.stack([torch.as_tensor(outputs[k]).detach().to(torch.float64).reshape(()) for k in self.KEYS])
        m = dict(zip(self.KEYS, vals.tolist()))  # one device read per batch
        for ds, fid in zip(batch["dataset"], batch["id"]):
            p = self.procedure_of(str(ds), str(fid))
            if p is not None:
                self.metrics_by_procedure[p].append(m)

    def summary(self) -> dict:
        import numpy as np
        per, allm = {}, {k: [] for k in self.KEYS}
        for proc, lst in self.metrics_by_procedure.items():
            arr = np.array([[m[k] for k in self.KEYS] for m in lst])
            mean = arr.mean(axis=0)
            per[proc] = dict(zip(self.KEYS, mean.tolist()))
            for k, v in zip(self.KEYS, mean):
                allm[k].append(v)
        overall = {k: {"mean": float(np.mean(v)) if v else float("nan"), "std": float(np.std(v)) if v else float("nan")}
                   for k, v in allm.items()}
        return {"per_procedure": per, "overall_metrics": overall}
