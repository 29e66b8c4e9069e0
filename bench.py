#!/usr/bin/env python
"""Headline benchmark: frames/s of DepthAnythingV2-L 518x518 depth + point cloud on N B200s.

  python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path = oracle port, rank 0)

One "step" = one batch of `--batch` synthetic SimCol-shaped frames per GPU through the hot path:
depth (DINOv2-L + DPT head on tcgen05 kernels) -> pose chain -> fused back-projection + SE(3) +
validity -> depth-metric partial sums (-> NCCL all-reduce of the sums + all-gather of the clouds when N>1).
Prints ONE JSON line (see the keys at the bottom)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "frames/s DAv2-L 518^2 depth+pointcloud"
GFLOP_PER_FRAME = {"vits": 115.3, "vitb": 380.7, "vitl": 1304.2}  # SURVEY.md 8d, dense 2*MAC @518^2


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1393.4), d.get("bf16_tflops", 1655.9), d.get("hbm_gbs", 6438.8), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nme, val in zip(names, r[4:8]):
                if val.strip().lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_gt(B, H, W, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    # Gamma(2, 0.15) = sum of two exponentials; clipped to [0,1]; 2 % invalid
    e = -0.15 * (torch.log(torch.rand(B, 1, H, W, generator=g, device=device).clamp_min(1e-9)) +
                 torch.log(torch.rand(B, 1, H, W, generator=g, device=device).clamp_min(1e-9)))
    gt = e.clamp(0, 1)
    gt[torch.rand(B, 1, H, W, generator=g, device=device) < 0.02] = 0.0
    return gt


def synth_frames(B, H, W, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    u = torch.rand(B, 3, H, W, generator=g, device=device)
    mean = torch.tensor([0.485, 0.456, 0.406], device=device).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=device).view(1, 3, 1, 1)
    return ((u - mean) / std).contiguous()


def synth_rel_poses(n, device, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(n, 3, generator=g) * 0.01
    rv = torch.randn(n, 3, generator=g) * (2.0 * np.pi / 180.0)
    ang = rv.norm(dim=1, keepdim=True).clamp_min(1e-12)
    q = torch.cat([rv / ang * torch.sin(ang / 2), torch.cos(ang / 2)], dim=1)
    return torch.cat([t, q], dim=1).float().to(device)


K518 = tuple(v * 518.0 / 475.0 for v in (156.0418, 155.7529, 178.5604, 181.8043))  # datasets/UnityCam/cam.txt:1


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_frames(encoder, size, state_dict, n_warm, n_timed, budget_s=None):
    """Mirrors run.py:195-262 (batch 1) + depth_to_pointcloud back-projection + test_step metrics."""
    from oracle import dav2_oracle as O
    from oracle import geometry_oracle as geo
    from oracle import metrics_oracle as met

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.MODEL_CONFIGS[encoder]
    m = O.DepthAnythingV2(encoder, cfg["features"], cfg["out_channels"], max_depth=20.0).eval()
    if state_dict is not None:
        m.load_state_dict(state_dict)
    else:
        m.load_state_dict(O.make_state_dict(encoder, 0))
    T = geo.make_transform([0.1, 0.2, 0.3], [0.0, 0.0174524, 0.0, 0.9998477])
    times = []
    t_start = time.perf_counter()
    for i in range(n_warm + n_timed):
        x = O.synthetic_frames(1, size, size, seed=100 + i)
        gt = synth_gt(1, size, size, "cpu", 7 + i).numpy()
        t0 = time.perf_counter()
        with torch.no_grad():
            d = m(x)
        pts, valid = geo.backproject(d[0].numpy(), K518, T)
        met.test_step_metrics(d[:, None].numpy(), gt, 1e-6, 20.0)
        dt = time.perf_counter() - t0
        if i >= n_warm:
            times.append(dt)
        if budget_s is not None and i >= n_warm and time.perf_counter() - t_start > budget_s:
            break
    return times, cores, torch.get_num_threads()


def workload_config(args, world=1, gather=None, gather_note=None):
    """The `config` object shared by both arms (BASELINE configs[2] on one GPU, the same per GPU for N > 1)."""
    B, S = args.batch, args.size
    multi = ""
    if world > 1:
        multi = ("; NCCL all-reduce of sums; clouds gathered by the back-projection kernel's stores into peer-mapped buffers"
                 if gather == "fused" else "; NCCL all-reduce of sums + all-gather of clouds" if gather == "nccl" else
                 "; NCCL all-reduce of sums")
    cfg = {"workload": f"BASELINE configs[2]: DepthAnythingV2 {args.encoder} batch {B}/GPU, {S}x{S} synthetic SimCol-shaped "
                       "frames, random-init weights; depth + pose chain + fused back-projection/SE(3)/validity + metric "
                       "partial sums" + multi,
           "encoder": args.encoder, "batch_per_gpu": B, "size": S,
           "l2": f"inputs re-read every step are {B * 3 * S * S * 4 / 1e6:.0f} MB and activations >10 GB, larger than the 126 MB L2"}
    if gather_note:
        cfg["gather_note"] = gather_note
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores, threads = cpu_reference_frames(args.encoder, args.size, None, args.warmup, args.steps)
    ms = 1e3 * float(np.mean(times))
    fps = 1e3 / ms
    sample = (f"1 frame/step (batch 1 like run.py:195-262), {len(times)} timed steps, {args.encoder} {args.size}x{args.size} fp32 "
              "oracle port (reference model code is an un-vendored external checkout) + numpy back-projection + metrics")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample, "host_cores": cores},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--encoder", default="vitl")
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--size", type=int, default=518)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"],
                    help="tensor-core operand format (fp16 = the reference's AMP 16-mixed; fp32 accumulate either way)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the cloud gather (N>1)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N>1: 'fused' = the back-projection kernel stores into every rank's peer-mapped gather buffer "
                         "(sharding.CloudGather); 'nccl' = local back-projection + all_gather_into_tensor")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from dav2_b200 import _lib, evaluation, ops, sharding, weights
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S = args.batch, args.size
    HW = S * S

    model = DepthAnythingV2(**MODEL_CONFIGS[args.encoder], max_depth=20.0, precision=args.precision)
    weights.randomize_(model, seed=0)
    cpu_sd = {k: v.clone() for k, v in model.state_dict().items()} if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    model = model.to(dev).eval()
    weights.calibrate_(model, dev)  # spread the synthetic depth over (0, max_depth) instead of saturating the sigmoid
    if cpu_sd is not None:
        cpu_sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}

    # device-resident inputs for `value`; distinct frames per rank (weak scaling)
    x_dev = synth_frames(B, S, S, dev, 1234 + rank)
    gt_dev = synth_gt(B, S, S, dev, 99 + rank)
    rel = synth_rel_poses(B, dev, 5 + rank)
    k4 = torch.tensor(K518, dtype=torch.float64, device=dev)
    xyz = torch.empty(B, HW, 3, dtype=torch.float32, device=dev)
    cloud_all = mask_all = fused = gather_note = None
    if world > 1 and not args.no_gather:
        if args.gather == "fused":
            # peer-mapped gather buffers (CUDA IPC over NVLink), double buffered.  If peer mapping is unavailable on this
            # box (IPC disabled, no P2P between some pair) EVERY rank falls back to the NCCL all-gather together.
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            try:
                fused = sharding.CloudGather(B, HW, dev)
            except Exception as e:  # noqa: BLE001
                gather_note = f"fused gather unavailable ({type(e).__name__}: {e}); NCCL all-gather used"
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0:
                if fused is not None:
                    fused.close(collective=False)  # a peer failed to map: do not wait for it
                fused = None
                gather_note = gather_note or "fused gather unavailable on a peer rank; NCCL all-gather used"
        if fused is None:
            cloud_all = torch.empty(world, B, HW, 3, dtype=torch.float32, device=dev)
            mask_all = torch.empty(world, B, HW, dtype=torch.uint8, device=dev)

    def step(x, gt):
        depth = model(x)
        _, T12 = ops.compose_poses(rel, None, want_T12=True)
        if fused is not None:
            # the kernel's stores ARE the all-gather; the metric all-reduce below orders readers after all writers
            _, _, counts_all = fused.backproject(depth, k4, T12[1:])
            counts = counts_all[rank * B:(rank + 1) * B]
        else:
            _, valid, counts = ops.backproject(depth, k4, T12[1:], out_xyz=xyz)
        part = evaluation.metric_partials(depth[:, None], gt, 1e-6, 20.0)
        if world > 1:
            dist.all_reduce(part)  # partial SUMS, finalised after the reduce
            if cloud_all is not None:
                dist.all_gather_into_tensor(cloud_all.view(-1), xyz.view(-1))
                dist.all_gather_into_tensor(mask_all.view(-1), valid.view(-1))
        return depth, counts, part

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(x_dev, gt_dev)
    barrier()

    # ---- timed region: device-resident inputs -------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.profile_enable(True)
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if os.environ.get("BENCH_CUDA_PROFILER"):  # ncu --profile-from-start off: capture exactly the timed steps
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(args.steps):
        depth, counts, part = step(x_dev, gt_dev)
    ev1.record()
    barrier()
    if os.environ.get("BENCH_CUDA_PROFILER"):
        torch.cuda.profiler.stop()
    launches = _lib.launch_count() - n0
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    clocks = sampler.stop() if sampler else None
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = world * B / (ms_step / 1e3)

    # ---- e2e: host (pinned) inputs through the public API, H2D + D2H inside the timed region ----------
    hx = [torch.empty(B, 3, S, S, dtype=torch.float32).pin_memory() for _ in range(2)]
    hg = [torch.empty(B, 1, S, S, dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(2):
        hx[i].copy_(x_dev.cpu()); hg[i].copy_(gt_dev.cpu())
    dx = [torch.empty_like(x_dev) for _ in range(2)]
    dg = [torch.empty_like(gt_dev) for _ in range(2)]
    h_part = torch.empty(8, dtype=torch.float64).pin_memory()
    h_counts = torch.empty(B, dtype=torch.int32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            dx[s].copy_(hx[s], non_blocking=True)
            dg[s].copy_(hg[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in range(2):
            consumed[s].record(main_stream)
        upload(0)
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                upload(i + 1)  # overlaps with this step's compute
            main_stream.wait_event(ready[s])
            depth, counts, part = step(dx[s], dg[s])
            consumed[s].record(main_stream)
            h_part.copy_(part, non_blocking=True)
            h_counts.copy_(counts, non_blocking=True)
            main_stream.synchronize()  # the host consumes the step's result
        return evaluation.finalize_compute_errors(h_part)

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    e2e_loop(args.steps)
    ev1.record()
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), 0.0)
    wall_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([max(e2e_ms, wall_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (float(t) / args.steps / 1e3)
    h2d = B * 3 * HW * 4 + B * HW * 4
    d2h = 8 * 8 + 4 * B

    fused_used = fused is not None
    if fused is not None:
        fused.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    sustained, burst, hbm, peak_src = load_peaks()
    mm = {k: prof.get(k, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0}) for k in ("gemm_tcgen05", "conv_tcgen05")}
    dom_ms = mm["gemm_tcgen05"]["ms"] + mm["conv_tcgen05"]["ms"]
    dom_fl = mm["gemm_tcgen05"]["flops"] + mm["conv_tcgen05"]["flops"]
    dom_n = mm["gemm_tcgen05"]["launches"] + mm["conv_tcgen05"]["launches"]
    achieved = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get("gemm_tcgen05_kernel_bytes_per_launch")
    bp = prof.get("backproject")
    breakdown = {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                     "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None,
                     "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 and v["flops"] == 0 else None}
                 for k, v in prof.items()}
    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if args.precision == "fp16" else "bf16",
        "data": "synthetic",
        "config": workload_config(args, world, None if (world == 1 or args.no_gather) else ("fused" if fused_used else "nccl"),
                                  gather_note),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "pinned host frames+gt -> H2D (double buffered on a copy stream) -> dav2 API -> D2H metric sums + counts; clouds stay in HBM"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (linear + implicit-GEMM conv instantiations)",
                     "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained if sustained else None,
                     "traffic": traffic, "launches_per_step": dom_n / args.steps, "avg_launch_ms": dom_ms / max(dom_n, 1),
                     "peak_source": f"{peak_src} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
                     "pipeline_frac": value / world * GFLOP_PER_FRAME.get(args.encoder, 0) / 1e3 / sustained if S == 518 else None},
        "kernels": breakdown,
    }
    if bp and bp["ms"] > 0:
        out["roofline_backproject"] = {"bound": "hbm", "achieved": bp["bytes"] / (bp["ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                       "frac": bp["bytes"] / (bp["ms"] * 1e-3) / 1e9 / hbm, "traffic": None}
    if world == 1 and not args.no_cpu_baseline:
        times, cores, threads = cpu_reference_frames(args.encoder, S, cpu_sd, 1, 3, budget_s=25.0)
        fps = 1.0 / float(np.mean(times))
        out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "host_cores": cores,
                               "sample": f"{len(times)} frames, batch 1 (run.py loop), same weights, fp32 oracle port + numpy "
                                         "back-projection + metrics, after 1 warm-up frame"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
