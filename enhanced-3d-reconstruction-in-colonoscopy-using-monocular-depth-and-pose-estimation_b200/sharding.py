"""Multi-GPU plumbing for the frame-parallel hot path.

Frames are independent, so every rank (one process per GPU) owns a CONTIGUOUS frame range and runs
the whole path on it with no data-path collective.  Only two exchanges exist, both after the kernels:
  * metric partial SUMS -> all_reduce(SUM), finalised afterwards so sharded == single-GPU results
    (compute_errors is a whole-batch reduction, lightning_model.py:310-313);
  * per-frame point clouds (dense xyz + validity mask, fixed shape) -> all_gather in frame order.
Backend agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def frame_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop): first ``total % world`` ranks get one extra frame; covers every frame once."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_partials(partials: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the fp64 metric partials over ranks (in place); no-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def gather_clouds(xyz: torch.Tensor, valid: Optional[torch.Tensor] = None, group=None):
    """xyz [b,HW,3] (+ valid [b,HW]) of equal-size shards -> [world*b,HW,3] (+ [world*b,HW]) in frame order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return xyz, valid
    world = dist.get_world_size(group)
    out = torch.empty((world * xyz.shape[0],) + tuple(xyz.shape[1:]), dtype=xyz.dtype, device=xyz.device)
    dist.all_gather_into_tensor(out.view(-1), xyz.contiguous().view(-1), group=group)
    vout = None
    if valid is not None:
        vout = torch.empty((world * valid.shape[0],) + tuple(valid.shape[1:]), dtype=valid.dtype, device=valid.device)
        dist.all_gather_into_tensor(vout.view(-1), valid.contiguous().view(-1), group=group)
    return out, vout
