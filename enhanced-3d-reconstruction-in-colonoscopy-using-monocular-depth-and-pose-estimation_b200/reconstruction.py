"""Full reconstruction pass (BASELINE config 4): depth for every frame, relative poses for every consecutive
pair, the pose chain, and world-frame back-projection -- the composition of reference stages
(run.py / depth_to_pointcloud.py / pose_estimation_model.py / eval/evaluation.py) that the reference never wires
together itself.

Sharding: rank r owns the contiguous frames [a, b).  Pair i = (frame i, frame i+1) belongs to the
rank that owns frame i, so a rank needs ONE halo frame (b) whose depth it recomputes locally instead of
exchanging it.  The chain itself is 28 bytes per pair: relative poses are all-gathered and every rank composes the
whole trajectory redundantly."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops, sharding
from .pose_estimation_model import stack_pairs  # noqa: F401  (the windowed form below is the same stacking)


def calculate_scale_factor(pred_rel_poses: torch.Tensor, gt_rel_poses: torch.Tensor) -> torch.Tensor:
    """eval/evaluation.py:257-276: sum(pred_t . gt_t) / sum(pred_t . pred_t)."""
    pt, gt = pred_rel_poses[:, :3], gt_rel_poses[:, :3]
    return torch.sum(pt * gt) / torch.sum(pt * pt)


def pair_counts(total: int, world: int) -> list:
    """Pairs (i, i+1) owned by each rank = the rank's frames except the video's last one; 0 for an empty shard."""
    out = []
    for r in range(world):
        a, b = sharding.frame_range(total, r, world)
        out.append(max(min(b, total - 1) - a, 0))
    return out


@torch.no_grad()
def reconstruct(frames: torch.Tensor, depth_model, pose_model, k4, scale: float | torch.Tensor = 1.0,
                batch: int = 16, depth_for_pose_scale: float = 1.0, initial_pose: Optional[torch.Tensor] = None,
                rank: int = 0, world: int = 1, group=None):
    """frames: [N,3,H,W] normalised RGB of ONE video (every rank passes the same tensor or at least its shard + halo).

    Returns dict(depth [n,H,W], rel [N-1,7] (scaled), abs [N,7], T12 [N,12], xyz [n,H*W,3], valid [n,H*W], counts [n],
    frame_range (a, b)) for this rank's frames; gather clouds with ``sharding.gather_clouds`` when shards are equal."""
    N, _, H, W = frames.shape
    a, b = sharding.frame_range(N, rank, world)
    hi = min(b + 1, N)  # halo frame for the last local pair
    dev = next(depth_model.parameters()).device
    if b <= a:
        # empty shard (N < world): no kernel runs here, but the rank still takes part in the pose all-gather below
        hi = a
    depth = torch.empty(max(hi - a, 0), H, W, dtype=torch.float32, device=dev)
    # relative poses of the local pairs (i, i+1), i in [a, min(b, N-1)): streamed batch by batch -- every batch of frames is
    # uploaded once, its depth computed, and its pairs (plus the pair that straddles the previous batch) go through the
    # pose network straight away, so neither the frames nor the [pairs, 8, H, W] stack is ever resident as a whole
    n_pairs = max(min(b, N - 1) - a, 0)
    rel_local = torch.zeros(n_pairs, 7, dtype=torch.float32, device=dev)
    prev = None  # last (rgb, scaled depth) frame of the previous batch
    for s in range(a, hi, batch):
        e = min(s + batch, hi)
        x = frames[s:e].to(dev, non_blocking=True)
        d = depth_model(x)
        depth[s - a:e - a] = d
        f = torch.cat([x, d[:, None] * depth_for_pose_scale], dim=1)          # [n, 4, H, W]
        first = s                                                              # global index of the first pair's frame i
        if prev is not None:
            f = torch.cat([prev, f], dim=0)
            first = s - 1
        if f.shape[0] > 1:
            pairs = torch.cat([f[:-1], f[1:]], dim=1).contiguous()            # == stack_pairs on this window
            rel_local[first - a:first - a + pairs.shape[0]] = pose_model(pairs)
        prev = f[-1:]
    # trajectory: gather the (tiny) relative poses, compose redundantly on every rank
    if world > 1:
        import torch.distributed as dist
        counts = pair_counts(N, world)
        parts = [torch.zeros(c, 7, dtype=torch.float32, device=dev) for c in counts]
        dist.all_gather(parts, rel_local, group=group) if len(set(counts)) == 1 else _all_gather_ragged(parts, rel_local, rank, group)
        rel = torch.cat(parts)
    else:
        rel = rel_local
    rel = rel.clone()
    rel[:, :3] *= scale  # the network predicts unit translation directions (data_processing/pose_estimation.py:256-258)
    abs7, T12 = ops.compose_poses(rel.contiguous(), initial_pose, want_T12=True)
    if b > a:
        xyz, valid, counts_v = ops.backproject(depth[:b - a].contiguous(), k4, T12[a:b].contiguous())
    else:
        xyz = torch.empty(0, H * W, 3, dtype=torch.float32, device=dev)
        valid = torch.empty(0, H * W, dtype=torch.uint8, device=dev)
        counts_v = torch.empty(0, dtype=torch.int32, device=dev)
    return {"depth": depth[:b - a], "rel": rel, "abs": abs7, "T12": T12, "xyz": xyz, "valid": valid, "counts": counts_v,
            "frame_range": (a, b)}


def _all_gather_ragged(parts, mine, rank, group):
    import torch.distributed as dist
    for r, buf in enumerate(parts):
        if r == rank:
            buf.copy_(mine)
        if buf.numel():
            dist.broadcast(buf, src=r, group=group)
