"""Multi-GPU plumbing for the frame-parallel hot path.

Frames are independent, so every rank (one process per GPU) owns a CONTIGUOUS frame range and runs
the whole path on it with no data-path collective.  Only two exchanges exist, both after the kernels:
  * metric partial SUMS -> all_reduce(SUM), finalised afterwards so sharded == single-GPU results
    (compute_errors is a whole-batch reduction, lightning_model.py:310-313);
  * per-frame point clouds (dense xyz + validity mask, fixed shape) -> all_gather in frame order.
Backend agnostic (NCCL on the GPUs, gloo in the CPU tests).

``CloudGather`` is the B200 form of the second exchange: every rank's gather buffer is peer-mapped into every other
rank (CUDA IPC over NVLink / NVSwitch) and the back-projection kernel stores each point straight into all of them, so
the kernel's output write IS the all-gather -- no NCCL kernel, no staging copy, and the transfer overlaps the math."""
from __future__ import annotations

import ctypes as C
import math
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


class _DevMem:
    """A raw device allocation exposed through __cuda_array_interface__ so torch can view it without owning it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class CloudGather:
    """Peer-mapped all-gather target for the fused back-projection (``dav2_backproject_gather``).

    Every rank holds ``n_buffers`` gathered clouds [world*frames_per_rank, HW, 3] (+ validity mask + per-frame counts);
    ``backproject`` makes the local kernel write this rank's frames into the current buffer of ALL ranks.  Ordering:
    readers must run after every rank's kernel has finished -- any later collective on the same streams gives that (the
    step's metric all-reduce does; ``complete()`` issues a 4-byte one otherwise) -- and because consecutive steps alternate
    buffers, a rank may start the next step while a slower peer still reads the previous cloud."""

    def __init__(self, frames_per_rank: int, HW: int, device, group=None, with_valid: bool = True, n_buffers: int = 2):
        from . import _lib
        self._lib = _lib.load()
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        if self.world > 8:
            raise ValueError("CloudGather: at most 8 ranks (one NVSwitch domain)")
        self.device = torch.device(device)
        self.fpr, self.HW, self.with_valid, self.n_buffers = int(frames_per_rank), int(HW), with_valid, int(n_buffers)
        F = self.world * self.fpr
        self.F = F
        self.off_valid = _align(F * self.HW * 12)
        self.off_counts = self.off_valid + _align(F * self.HW if with_valid else 0)
        self.per = self.off_counts + _align(F * 4)
        total = self.per * self.n_buffers
        with torch.cuda.device(self.device):
            base = C.c_void_p()
            _lib.check(self._lib.dav2_peer_alloc(C.byref(base), total), "dav2_peer_alloc")
            self.base = int(base.value)
            handle = (C.c_uint8 * 64)()
            _lib.check(self._lib.dav2_peer_export(self.base, handle), "dav2_peer_export")
            self.peers = [self.base]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle), group=group)
                self.peers = []
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self.peers.append(self.base)
                    else:
                        p = C.c_void_p()
                        buf = (C.c_uint8 * 64).from_buffer_copy(h)
                        _lib.check(self._lib.dav2_peer_open(buf, C.byref(p)), "dav2_peer_open")
                        self.peers.append(int(p.value))
        self._mem = torch.as_tensor(_DevMem(self.base, total), device=self.device)
        self._step = 0
        self._closed = False

    def views(self, buf: int):
        """(xyz [F,HW,3] fp32, valid [F,HW] u8 | None, counts [F] i32) of local buffer ``buf``."""
        m = self._mem[buf * self.per:(buf + 1) * self.per]
        F, HW = self.F, self.HW
        xyz = m[:F * HW * 12].view(torch.float32).view(F, HW, 3)
        valid = m[self.off_valid:self.off_valid + F * HW].view(F, HW) if self.with_valid else None
        counts = m[self.off_counts:self.off_counts + F * 4].view(torch.int32)
        return xyz, valid, counts

    def backproject(self, depth: torch.Tensor, K4, T12=None, depth_scale: float = 1.0, depth_trunc: float = math.inf,
                    gt: Optional[torch.Tensor] = None, min_depth: float = 1e-6, max_depth: float = 20.0):
        """depth [frames_per_rank,H,W] of THIS rank -> views of the gathered cloud (complete after ``complete()`` or any
        later collective on this stream).  With ``gt`` the same pass also produces this rank's test_step metric partial
        sums (fp64 [8]) and the call returns (xyz, valid, counts, partials)."""
        from . import ops
        assert depth.shape[0] == self.fpr and depth.shape[1] * depth.shape[2] == self.HW
        buf = self._step % self.n_buffers
        self._step += 1
        o = buf * self.per
        part = ops.backproject_gather(depth, K4, T12, [p + o for p in self.peers],
                                      [p + o + self.off_valid for p in self.peers] if self.with_valid else None,
                                      [p + o + self.off_counts for p in self.peers], frame_offset=self.rank * self.fpr,
                                      depth_scale=depth_scale, depth_trunc=depth_trunc, gt=gt, min_depth=min_depth,
                                      max_depth=max_depth)
        return self.views(buf) if gt is None else self.views(buf) + (part,)

    def complete(self):
        """Order this stream after every rank's back-projection kernel (a 4-byte all-reduce)."""
        if self.world > 1:
            dist.all_reduce(torch.zeros(1, dtype=torch.int32, device=self.device), group=self.group)

    def close(self, collective: bool = True):
        """Unmap the peers and free the local buffer.  ``collective=False`` skips the barriers (error paths in which
        not every rank owns a CloudGather); the caller must then know that no peer kernel can still write here."""
        if self._closed:
            return
        self._closed = True
        from . import _lib
        torch.cuda.synchronize(self.device)
        if self.world > 1 and collective:
            dist.barrier(group=self.group)  # nobody unmaps / frees while a peer kernel may still write
        self._mem = None
        with torch.cuda.device(self.device):
            for r, p in enumerate(self.peers):
                if r != self.rank:
                    _lib.check(self._lib.dav2_peer_close(p), "dav2_peer_close")
            if self.world > 1 and collective:
                dist.barrier(group=self.group)
            _lib.check(self._lib.dav2_peer_free(self.base), "dav2_peer_free")
