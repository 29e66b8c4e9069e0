"""World-size-2 gloo test of the multi-GPU host logic: contiguous frame shards, all-reduce of metric
partial SUMS then finalise == whole-batch metrics, cloud gather in frame order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partials(pred, gt, lo=1e-6, hi=20.0):
    m = (gt >= lo) & (gt <= hi)
    p, g = pred[m].astype(np.float32), gt[m].astype(np.float32)
    d = p - g
    t = np.maximum(g / p, p / g)
    return np.array([p.size, np.abs(d).sum(dtype=np.float64), (np.abs(d) / (g + np.float32(1e-6))).sum(dtype=np.float64),
                     (d.astype(np.float64) ** 2).sum(), g.sum(dtype=np.float64), (t < np.float32(1.1)).sum(), 0, 0], np.float64)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dav2_b200 import evaluation as ev
    from dav2_b200 import sharding
    rng = np.random.default_rng(7)
    gt = np.clip(rng.gamma(2.0, 0.15, size=(total, 1, 12, 10)), 0, 1).astype(np.float32)
    pred = (np.where(gt > 0, gt, 0.3) * rng.normal(1, 0.07, size=gt.shape)).astype(np.float32)
    a, b = sharding.frame_range(total, rank, world)
    part = torch.from_numpy(_partials(pred[a:b], gt[a:b]))
    sharding.allreduce_partials(part)
    out = ev.finalize_compute_errors(part)
    xyz = torch.arange(a, b, dtype=torch.float32).view(-1, 1, 1).expand(b - a, 5, 3).contiguous()
    valid = torch.full((b - a, 5), rank, dtype=torch.uint8)
    allx, allv = sharding.gather_clouds(xyz, valid)
    if rank == 0:
        q.put(({k: float(v) for k, v in out.items()}, allx[:, 0, 0].tolist(), allv[:, 0].tolist()))
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_process():
    from oracle import metrics_oracle as met
    world, total = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    metrics, order, vmask = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    gt = np.clip(rng.gamma(2.0, 0.15, size=(total, 1, 12, 10)), 0, 1).astype(np.float32)
    pred = (np.where(gt > 0, gt, 0.3) * rng.normal(1, 0.07, size=gt.shape)).astype(np.float32)
    ref = met.test_step_metrics(pred, gt)
    for k in ref:
        assert abs(metrics[k] - ref[k]) < 1e-6, k
    assert order == [float(i) for i in range(total)]   # gather keeps frame order
    assert vmask == [0, 0, 0, 1, 1, 1]


def test_frame_range_partition():
    from dav2_b200 import sharding
    for total in (0, 1, 7, 64, 4096, 1000):
        for world in (1, 2, 3, 4, 8):
            rs = [sharding.frame_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frame_range(4, 2, 2)


def test_cloud_gather_needs_the_library_and_a_gpu():
    """CloudGather is GPU-only plumbing: on a CPU box it must fail loudly (no silent fallback), and its layout maths
    (256-byte aligned xyz | valid | counts sections, double buffered) is checkable without a device."""
    from dav2_b200 import sharding
    assert sharding._align(1) == 256 and sharding._align(256) == 256 and sharding._align(257) == 512
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            sharding.CloudGather(2, 16, "cuda")


def test_pair_counts_with_empty_shards():
    """reconstruction.reconstruct: a rank whose shard is empty (fewer frames than ranks) owns no pair and runs no kernel,
    but the per-rank pair counts every rank derives must still add up to N - 1."""
    from dav2_b200 import reconstruction as rc
    assert rc.pair_counts(10, 1) == [9]
    assert rc.pair_counts(10, 4) == [3, 3, 2, 1]
    assert rc.pair_counts(3, 8) == [1, 1, 0, 0, 0, 0, 0, 0]
    assert rc.pair_counts(1, 2) == [0, 0] and rc.pair_counts(0, 2) == [0, 0]
    for total in (0, 1, 2, 5, 1000):
        for world in (1, 2, 3, 8):
            assert sum(rc.pair_counts(total, world)) == max(total - 1, 0)
