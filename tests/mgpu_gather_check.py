"""Run under torchrun with one rank per GPU: the fused back-projection + peer-memory gather (sharding.CloudGather)
must equal an NCCL all_gather of the per-rank clouds, for several steps (double buffering) with per-rank poses."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from dav2_b200 import ops, sharding


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, H, W = 4, 70, 98
    k4 = (90.0, 91.0, 48.5, 35.2)
    cg = sharding.CloudGather(B, H * W, dev)
    for step in range(4):
        g = torch.Generator(device=dev).manual_seed(100 * step + rank)
        depth = torch.rand(B, H, W, generator=g, device=dev) * 3.0
        depth[:, :2, :5] = 0.0
        T12 = (torch.eye(4, dtype=torch.float64)[:3].reshape(1, 12).repeat(B, 1) + 0.01 * (rank + 1)).to(dev)
        xyz_all, valid_all, counts_all = cg.backproject(depth, k4, T12)
        part = torch.ones(8, dtype=torch.float64, device=dev)
        sharding.allreduce_partials(part)  # the step's metric all-reduce orders readers after all writers
        rx, rv, rc = ops.backproject(depth, k4, T12)
        ex, ev = sharding.gather_clouds(rx, rv)
        ec = torch.empty(world * B, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(ec, rc)
        assert torch.equal(xyz_all, ex), f"rank {rank} step {step}: xyz mismatch"
        assert torch.equal(valid_all, ev), f"rank {rank} step {step}: valid mismatch"
        assert torch.equal(counts_all, ec), f"rank {rank} step {step}: counts mismatch"
        assert float(part[0]) == world
    cg.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_GATHER_OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
