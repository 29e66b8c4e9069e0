#!/usr/bin/env python
"""Headline benchmark: frames/s of DepthAnythingV2 depth + point cloud on N B200s.

  python bench.py --gpus N --steps K --warmup W [--config 3]    (ours; N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...        (the reference's CPU path = oracle port, rank 0)

--config selects the BASELINE.json configuration (1-based like SURVEY.md 8d; default 3 = the headline):
  2  vitb, batch 32, 518^2: depth + AbsRel / d1 / RMSE metrics (both metric definitions), one GPU
  3  vitl, batch 64 per GPU, 518^2: depth -> pose chain -> fused back-projection + SE(3) + validity -> metric partial
     sums (-> clouds gathered by the back-projection kernel's peer stores when N > 1); weak scaling
  4  full reconstruction pass over a 1000-frame video (dav2_b200.reconstruction.reconstruct): vitl depth, ResNet-18 pose
     on stacked pairs, pose chain, world-frame clouds; frames sharded over the ranks with a one-frame halo; strong scaling
  5  vitl, batch 16 per GPU, 1036^2 (5477 tokens): the same step as config 3
One "step" = one batch per GPU through that path (config 4: the whole video).  Prints ONE JSON line (keys at the bottom)."""
import argparse
import hashlib
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

GFLOP_PER_FRAME = {("vits", 518): 115.3, ("vitb", 518): 380.7, ("vitl", 518): 1304.2, ("vitl", 1036): 7424.0}
POSE_GFLOP_PER_PAIR = 22.2
CONFIGS = {
    2: dict(encoder="vitb", batch=32, size=518, scaling="weak",
            metric="frames/s DAv2-B 518^2 depth+metrics",
            what="depth + compute_errors partial sums (test_step mask) + per-frame calculate_metrics partial sums"),
    3: dict(encoder="vitl", batch=64, size=518, scaling="weak",
            metric="frames/s DAv2-L 518^2 depth+pointcloud",
            what="depth + pose chain + ONE fused pass: back-projection/SE(3)/validity + metric partial sums"),
    4: dict(encoder="vitl", batch=64, size=518, scaling="strong", frames=1000,
            metric="frames/s DAv2-L 518^2 full reconstruction pass (depth + ResNet-18 pose + chain + world cloud)",
            what="reconstruction.reconstruct over a 1000-frame video: vitl depth, ResNet-18 pose on stacked pairs "
                 "[rgb_i, d_i, rgb_i+1, d_i+1], compose_poses, world-frame back-projection; frames sharded with a one-frame halo"),
    5: dict(encoder="vitl", batch=16, size=1036, scaling="weak",
            metric="frames/s DAv2-L 1036^2 depth+pointcloud",
            what="depth (5477 tokens) + pose chain + fused back-projection/SE(3)/validity + metric partial sums"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1393.4), d.get("bf16_tflops", 1655.9), d.get("hbm_gbs", 6438.8), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def csrc_sha16():
    """Identity of the kernel sources: profiles/traffic_*.json carries the value it was measured at."""
    d = os.path.join(ROOT, "enhanced-3d-reconstruction-in-colonoscopy-using-monocular-depth-and-pose-estimation_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(key="gemm_tcgen05_kernel_bytes_per_launch"):
    """DRAM bytes per launch of the dominant kernel from the committed ncu launch list -- only if that list was taken
    with THESE kernel sources (otherwise null: a constant from an older build would be stale)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if f.startswith("traffic") and f.endswith(".json"):
            try:
                d = json.load(open(os.path.join(pdir, f)))
            except ValueError:
                continue
            if d.get("csrc_sha16") == csrc_sha16():
                best = (d.get(key), f)
    return best if best else (None, None)


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for nme, val in zip(names, r[4:8]):
                if val.strip().lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


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


K475 = (156.0418, 155.7529, 178.5604, 181.8043)  # datasets/UnityCam/cam.txt:1


def k_for(size):
    return tuple(v * size / 475.0 for v in K475)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_frames(cfgno, encoder, size, state_dict, n_warm, n_timed, budget_s=None):
    """Mirrors run.py:195-262 (batch 1) + depth_to_pointcloud back-projection + test_step metrics; config 4 adds the
    ResNet-18 pose network on the stacked pair and one compose_poses step per frame."""
    from oracle import dav2_oracle as O
    from oracle import geometry_oracle as geo
    from oracle import metrics_oracle as met

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.MODEL_CONFIGS[encoder]
    m = O.DepthAnythingV2(encoder, cfg["features"], cfg["out_channels"], max_depth=20.0).eval()
    m.load_state_dict(state_dict if state_dict is not None else O.make_state_dict(encoder, 0))
    pose = None
    if cfgno == 4:
        from oracle import pose_oracle
        pose = pose_oracle.build_pose_oracle(8, 0).eval()
    T = geo.make_transform([0.1, 0.2, 0.3], [0.0, 0.0174524, 0.0, 0.9998477])
    k4 = k_for(size)
    times = []
    t_start = time.perf_counter()
    prev = None
    for i in range(n_warm + n_timed):
        x = O.synthetic_frames(1, size, size, seed=100 + i)
        gt = synth_gt(1, size, size, "cpu", 7 + i).numpy()
        t0 = time.perf_counter()
        with torch.no_grad():
            d = m(x)
            if pose is not None:
                cur = torch.cat([x, d[:, None]], dim=1)
                if prev is not None:
                    rel = pose(torch.cat([prev, cur], dim=1)).numpy()
                    geo.compose_poses(rel)
                prev = cur
        geo.backproject(d[0].numpy(), k4, T)
        if cfgno != 4:
            met.test_step_metrics(d[:, None].numpy(), gt, 1e-6, 20.0)
        dt = time.perf_counter() - t0
        if i >= n_warm:
            times.append(dt)
        if budget_s is not None and i >= n_warm and time.perf_counter() - t_start > budget_s:
            break
    return times, cores, torch.get_num_threads()


def workload_config(args, world=1, gather=None, gather_note=None):
    """The `config` object shared by both arms."""
    c = CONFIGS[args.config]
    B, S = args.batch, args.size
    multi = ""
    if world > 1 and args.config in (3, 5):
        multi = ("; clouds gathered by the back-projection kernel's stores into peer-mapped buffers (NVLink), metric sums "
                 "accumulated on device and all-reduced once per timed region" if gather == "fused" else
                 "; NCCL all-gather of clouds every step, metric sums all-reduced once per timed region" if gather == "nccl" else
                 "; metric sums all-reduced once per timed region")
    per = (f"{c['frames']} frames per step over all GPUs (batches of {B})" if args.config == 4 else f"batch {B}/GPU")
    cfg = {"workload": f"BASELINE configs[{args.config - 1}] (SURVEY.md 8d config {args.config}): DepthAnythingV2 {args.encoder} "
                       f"{per}, {S}x{S} synthetic SimCol-shaped frames, random-init weights; {c['what']}{multi}",
           "baseline_config": args.config, "encoder": args.encoder, "batch_per_gpu": B, "size": S,
           "l2": f"inputs re-read every step are {B * 3 * S * S * 4 / 1e6:.0f} MB and activations >10 GB, larger than the 126 MB L2"}
    if gather_note:
        cfg["gather_note"] = gather_note
    if world > 1 and gather == "fused":
        cfg["gather_overlap"] = ("side stream: the gather stores of step i overlap the encoder of step i+1" if not args.no_overlap
                                 else "none (main stream)")
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores, threads = cpu_reference_frames(args.config, args.encoder, args.size, None, args.warmup, args.steps)
    ms = 1e3 * float(np.mean(times))
    fps = 1e3 / ms
    sample = (f"1 frame/step (batch 1 like run.py:195-262), {len(times)} timed steps, {args.encoder} {args.size}x{args.size} fp32 "
              "oracle port (reference model code is an un-vendored external checkout) + numpy back-projection + metrics"
              + (" + ResNet-18 pose on the stacked pair + compose_poses" if args.config == 4 else ""))
    print(json.dumps({
        "impl": "reference", "metric": CONFIGS[args.config]["metric"], "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": CONFIGS[args.config]["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample, "host_cores": cores},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def gpu_eager_baseline(encoder, size, state_dict, dev, batch):
    """SURVEY.md 8d "GPU baseline beside it": the oracle architecture in eager PyTorch on the SAME GPU with the
    reference's own GPU settings (fp16 autocast + TF32: configs/trainer/default.yaml:4, test_lightning.py:24), same
    weights, CUDA-event timed, forward only."""
    from oracle import dav2_oracle as O
    cfg = O.MODEL_CONFIGS[encoder]
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        m = O.DepthAnythingV2(encoder, cfg["features"], cfg["out_channels"], max_depth=20.0).eval()
        m.load_state_dict(state_dict)
        m = m.to(dev)
        x = synth_frames(batch, size, size, dev, 4321)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            for _ in range(2):
                m(x)
            torch.cuda.synchronize(dev)
            e0.record()
            n = 3
            for _ in range(n):
                m(x)
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / n
        del m, x
        torch.cuda.empty_cache()
        return {"value": batch / ms * 1e3, "unit": "frames/s", "batch": batch, "ms_per_batch": ms, "steps": n,
                "what": "oracle architecture in eager PyTorch, fp16 autocast + TF32 (the reference's GPU settings), same GPU and "
                        "weights, depth forward only (no back-projection / metrics), CUDA events"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE configuration (SURVEY.md 8d numbering)")
    ap.add_argument("--encoder", default=None)
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU per step")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--frames", type=int, default=None, help="config 4: frames of the video (default 1000)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"],
                    help="tensor-core operand format (fp16 = the reference's AMP 16-mixed; fp32 accumulate either way)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the cloud gather (N>1)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N>1, fused gather: keep the back-projection/gather kernel on the main stream instead of a side stream "
                         "(where its NVLink stores overlap the next batch's encoder)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N>1: 'fused' = the back-projection kernel stores into every rank's peer-mapped gather buffer "
                         "(sharding.CloudGather); 'nccl' = local back-projection + all_gather_into_tensor")
    args = ap.parse_args()
    c = CONFIGS[args.config]
    args.encoder = args.encoder or c["encoder"]
    args.batch = args.batch or c["batch"]
    args.size = args.size or c["size"]
    args.frames = args.frames or c.get("frames")
    if args.steps is None:
        args.steps = 3 if args.config == 4 else 8
    args.warmup = max(args.warmup, 3) if (args.impl == "ours" and args.config != 4) else max(args.warmup, 1)

    if args.impl == "reference":
        return run_reference(args)
    if args.config == 4:
        return run_reconstruction(args)

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
    cloud = args.config in (3, 5)

    model = DepthAnythingV2(**MODEL_CONFIGS[args.encoder], max_depth=20.0, precision=args.precision)
    weights.randomize_(model, seed=0)
    model = model.to(dev).eval()
    weights.calibrate_(model, dev)  # spread the synthetic depth over (0, max_depth) instead of saturating the sigmoid
    want_sd = rank == 0 and world == 1 and not (args.no_cpu_baseline and args.no_gpu_baseline)
    cpu_sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()} if want_sd else None

    # device-resident inputs for `value`; distinct frames per rank (weak scaling)
    x_dev = synth_frames(B, S, S, dev, 1234 + rank)
    gt_dev = synth_gt(B, S, S, dev, 99 + rank)
    rel = synth_rel_poses(B, dev, 5 + rank)
    k4 = torch.tensor(k_for(S), dtype=torch.float64, device=dev)
    xyz_buf = [torch.empty(B, HW, 3, dtype=torch.float32, device=dev) for _ in range(2)] if cloud else None
    cloud_all = mask_all = fused = gather_note = None
    if cloud and world > 1 and not args.no_gather:
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

    part_acc = torch.zeros(8, dtype=torch.float64, device=dev)   # metric partial SUMS accumulate on the device ...
    step_no = [0]

    # N > 1, fused gather: the back-projection kernel's NVLink stores (1.86 GB per step at N = 8: 2.4 ms) run on a SIDE stream,
    # ordered after the depth of their own step only, so they overlap the next batch's encoder instead of sitting on the
    # critical path (round 1: un-attributed step gap 0.7 -> 6.1 ms from N = 1 to 8).
    main_stream = torch.cuda.current_stream(dev)
    cloud_stream = torch.cuda.Stream(device=dev) if (fused is not None and not args.no_overlap) else main_stream

    def step(x, gt, before_cloud=None):
        """One batch through the path.  Returns (depth, xyz, valid, counts, partials, done) of THIS rank's frames; `done` is
        an event after the step's last kernel.  `before_cloud`: event the cloud stage must wait for (its output buffers)."""
        i = step_no[0]
        step_no[0] += 1
        depth = model(x)
        xyz = valid = counts = None
        done = torch.cuda.Event()
        if cloud:
            if cloud_stream is not main_stream:
                have_depth = torch.cuda.Event()
                have_depth.record(main_stream)
                cloud_stream.wait_event(have_depth)
                depth.record_stream(cloud_stream)
            if before_cloud is not None:
                cloud_stream.wait_event(before_cloud)
            with torch.cuda.stream(cloud_stream):
                _, T12 = ops.compose_poses(rel, None, want_T12=True)
                # ONE pass over the depth map: back-projection + SE(3) + validity + the metric partial sums
                if fused is not None:
                    # the kernel's stores ARE the all-gather (no collective in the step; CloudGather.complete() orders a
                    # reader of the whole gathered cloud after all writers)
                    xa, va, ca, part = fused.backproject(depth, k4, T12[1:], gt=gt, min_depth=1e-6, max_depth=20.0)
                    xyz, valid, counts = xa[rank * B:(rank + 1) * B], va[rank * B:(rank + 1) * B], ca[rank * B:(rank + 1) * B]
                else:
                    xyz, valid, counts, part = ops.backproject_metrics(depth, gt, k4, T12[1:], 1e-6, 20.0, out_xyz=xyz_buf[i % 2])
                    if cloud_all is not None:
                        dist.all_gather_into_tensor(cloud_all.view(-1), xyz.view(-1))
                        dist.all_gather_into_tensor(mask_all.view(-1), valid.view(-1))
                part_acc.add_(part)  # sums accumulate on the device, all-reduced ONCE per timed region
                done.record(cloud_stream)
        else:
            if before_cloud is not None:
                main_stream.wait_event(before_cloud)
            part = evaluation.metric_partials(depth[:, None], gt, 1e-6, 20.0)           # compute_errors / test_step
            ops.depth_metric_partials(depth[:, None].contiguous(), gt, 0.0, 0.0, 1, True)  # calculate_metrics per frame
            part_acc.add_(part)
            done.record(main_stream)
        return depth, xyz, valid, counts, part, done

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (+ self-verification of the fused gather where the driver can see it) -----------------
    gather_verified = None
    for w in range(args.warmup):
        depth, xyz, valid, counts, _, done = step(x_dev, gt_dev)
        main_stream.wait_event(done)
        if w == 0 and fused is not None:
            fused.complete()  # every rank's kernel has finished writing into every buffer
            buf = (fused._step - 1) % fused.n_buffers
            xa, va, ca = fused.views(buf)
            _, T12 = ops.compose_poses(rel, None, want_T12=True)
            rx, rv, rc = ops.backproject(depth, k4, T12[1:])
            ex = torch.empty(world * B, HW, 3, dtype=torch.float32, device=dev)
            ev = torch.empty(world * B, HW, dtype=torch.uint8, device=dev)
            ec = torch.empty(world * B, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(ex.view(-1), rx.view(-1))
            dist.all_gather_into_tensor(ev.view(-1), rv.view(-1))
            dist.all_gather_into_tensor(ec, rc)
            okv = torch.tensor([int(torch.equal(xa, ex) and torch.equal(va, ev) and torch.equal(ca, ec))], dtype=torch.int32, device=dev)
            dist.all_reduce(okv, op=dist.ReduceOp.MIN)
            gather_verified = bool(int(okv))
            del ex, ev, ec, rx, rv, rc
    barrier()

    # ---- timed region: device-resident inputs -------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.profile_enable(True)
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    part_acc.zero_()
    barrier()
    if os.environ.get("BENCH_CUDA_PROFILER"):  # ncu --profile-from-start off: capture exactly the timed steps
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(args.steps):
        done = step(x_dev, gt_dev)[-1]
    main_stream.wait_event(done)  # the last step's cloud stage (the side stream is in order)
    if world > 1:
        dist.all_reduce(part_acc)  # the one exchange of metric sums; also orders any reader after every rank's last kernel
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
    metrics_all = {k: float(v) for k, v in evaluation.finalize_compute_errors(part_acc).items()}

    # ---- e2e: pinned host inputs -> H2D -> public API -> D2H of the PRODUCT (depth, cloud, mask, metric sums) -----------
    # Three streams: uploads of step i+1 and downloads of step i-1 overlap the compute of step i; the host blocks on the
    # download of step i-1 before it queues step i+1 ("the caller consumes the product"), so at most two steps are in flight.
    hx = [torch.empty(B, 3, S, S, dtype=torch.float32).pin_memory() for _ in range(2)]
    hg = [torch.empty(B, 1, S, S, dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(2):
        hx[i].copy_(x_dev.cpu()); hg[i].copy_(gt_dev.cpu())
    dx = [torch.empty_like(x_dev) for _ in range(2)]
    dg = [torch.empty_like(gt_dev) for _ in range(2)]
    h_part = [torch.empty(8, dtype=torch.float64).pin_memory() for _ in range(2)]
    h_depth = [torch.empty(B, S, S, dtype=torch.float32).pin_memory() for _ in range(2)]
    h_xyz = [torch.empty(B, HW, 3, dtype=torch.float32).pin_memory() for _ in range(2)] if cloud else None
    h_valid = [torch.empty(B, HW, dtype=torch.uint8).pin_memory() for _ in range(2)] if cloud else None
    h_counts = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)] if cloud else None
    up_stream, down_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    computed = [torch.cuda.Event() for _ in range(2)]
    downloaded = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        s = i % 2
        with torch.cuda.stream(up_stream):
            up_stream.wait_event(consumed[s])
            dx[s].copy_(hx[s], non_blocking=True)
            dg[s].copy_(hg[s], non_blocking=True)
            ready[s].record(up_stream)

    def e2e_loop(n, product):
        for s in range(2):
            consumed[s].record(main_stream)
            downloaded[s].record(down_stream)
        upload(0)
        keep = [None, None]
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                upload(i + 1)  # overlaps with this step's compute
            main_stream.wait_event(ready[s])
            # step i-2's product must have left the device buffers this step overwrites (cloud stage waits for it)
            depth, xyz, valid, counts, part, done = step(dx[s], dg[s], before_cloud=downloaded[s])
            consumed[s] = done   # the upload stream may overwrite dx[s] / dg[s] once the whole step has read them
            computed[s] = done
            keep[s] = (depth, xyz, valid, counts, part)  # keep the tensors alive until their download is done
            with torch.cuda.stream(down_stream):
                down_stream.wait_event(computed[s])
                h_part[s].copy_(part, non_blocking=True)
                if product:
                    h_depth[s].copy_(depth, non_blocking=True)
                    if cloud:
                        h_xyz[s].copy_(xyz, non_blocking=True)
                        h_valid[s].copy_(valid, non_blocking=True)
                if cloud:
                    h_counts[s].copy_(counts, non_blocking=True)
                downloaded[s].record(down_stream)
            if i >= 1:
                downloaded[(i - 1) % 2].synchronize()  # the host consumes step i-1's product
        downloaded[(n - 1) % 2].synchronize()
        return evaluation.finalize_compute_errors(h_part[(n - 1) % 2])

    def timed_e2e(product):
        e2e_loop(2, product)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        e2e_loop(args.steps, product)
        ev1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        tt = torch.tensor([max(ev0.elapsed_time(ev1), wall_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * B / (float(tt) / args.steps / 1e3)

    e2e_value = timed_e2e(True)
    e2e_metrics_only = timed_e2e(False)
    h2d = B * 3 * HW * 4 + B * HW * 4
    d2h_small = 8 * 8 + (4 * B if cloud else 0)
    d2h = d2h_small + B * HW * 4 + (B * HW * 13 if cloud else 0)

    # the back-projection (+ metric sums) pass alone, back to back: inside the step it runs at the power-capped clock of the
    # dense kernels around it and is instruction-issue bound there, so the in-step figure (roofline_backproject) is reported
    # next to this one
    bp_alone = None
    if cloud:
        _, T12 = ops.compose_poses(rel, None, want_T12=True)
        xyz_t = xyz_buf[0]
        g2 = synth_gt(B, S, S, dev, 99 + rank)
        d2 = model(synth_frames(B, S, S, dev, 1234 + rank))
        for _ in range(3):
            ops.backproject_metrics(d2, g2, k4, T12[1:], 1e-6, 20.0, out_xyz=xyz_t)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(20):
            ops.backproject_metrics(d2, g2, k4, T12[1:], 1e-6, 20.0, out_xyz=xyz_t)
        ev1.record()
        torch.cuda.synchronize()
        bp_alone = ev0.elapsed_time(ev1) / 20.0  # ms per launch incl. the two 0.3 KB memsets of each call
        del d2, g2

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
    traffic, traffic_file = measured_traffic()
    breakdown = {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                     "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None,
                     "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 and v["flops"] == 0 else None}
                 for k, v in prof.items()}
    gf = GFLOP_PER_FRAME.get((args.encoder, S))
    out = {
        "metric": c["metric"], "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None,
        "dtype": "f16" if args.precision == "fp16" else "bf16",
        "data": "synthetic",
        "config": workload_config(args, world, None if (world == 1 or args.no_gather or not cloud) else ("fused" if fused_used else "nccl"),
                                  gather_note),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "pinned host frames+gt -> H2D (copy stream, double buffered) -> dav2 API -> D2H of the product into pinned host "
                        "memory on a third stream: depth [B,H,W] fp32" + (", cloud xyz [B,HW,3] fp32, validity mask, per-frame counts" if cloud else "")
                        + ", metric sums; the host blocks on step i-1's product before queueing step i+1"},
        "e2e_metrics_only": {"value": e2e_metrics_only, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_small,
                             "note": "same loop returning only metric sums + counts (round 1's e2e definition)"},
        "gpu_launches": int(launches),
        "metrics": metrics_all,
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (linear + implicit-GEMM conv instantiations)",
                     "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained if sustained else None,
                     "traffic": traffic, "traffic_source": traffic_file or "null: no ncu launch list under profiles/ was taken with these kernel sources",
                     "csrc_sha16": csrc_sha16(),
                     "launches_per_step": dom_n / args.steps, "avg_launch_ms": dom_ms / max(dom_n, 1),
                     "peak_source": f"{peak_src} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
                     "pipeline_frac": value / world * gf / 1e3 / sustained if gf else None},
        "kernels": breakdown,
    }
    if gather_verified is not None:
        out["gather_verified"] = gather_verified
    # (config 3 / 5: the metric sums ride in the back-projection pass -- its 21 B/px are in "backproject")
    for key, name in (("backproject", "roofline_backproject"), ("depth_metrics", "roofline_depth_metrics")):
        bp = prof.get(key)
        if bp and bp["ms"] > 0:
            out[name] = {"bound": "hbm", "achieved": bp["bytes"] / (bp["ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": bp["bytes"] / (bp["ms"] * 1e-3) / 1e9 / hbm,
                         # DRAM bytes per launch from the committed launch list (same kernel sources only); the depth map
                         # the pass reads was just written by the head conv, so part of it is served by L2
                         "traffic": measured_traffic("backproject_kernel_bytes_per_launch")[0] if key == "backproject" and world == 1 and args.config == 3 else None}
    if bp_alone:
        gbs = B * HW * 21.0 / (bp_alone * 1e-3) / 1e9
        out["roofline_backproject_alone"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                             "us_per_launch": bp_alone * 1e3, "traffic": None,
                                             "what": "the same fused pass (21 B/px, single destination) launched 20x back to back right after the timed steps "
                                                     "(the SM clock is still at the power-capped level of the step; at free clocks the "
                                                     "kernel takes 72 us = 0.78, profiles/README.md)"}
    if world == 1 and not args.no_gpu_baseline:
        del x_dev, gt_dev, dx, dg
        torch.cuda.empty_cache()
        try:
            out["gpu_eager_baseline"] = gpu_eager_baseline(args.encoder, S, cpu_sd, dev, min(B, 16 if S <= 518 else 2))
        except Exception as e:  # noqa: BLE001  (e.g. out of memory on a shared box: the headline must still print)
            out["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if world == 1 and not args.no_cpu_baseline:
        times, cores, threads = cpu_reference_frames(args.config, args.encoder, S, cpu_sd, 1, 3, budget_s=25.0 if S <= 518 else 60.0)
        fps = 1.0 / float(np.mean(times))
        out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "host_cores": cores,
                               "sample": f"{len(times)} frames, batch 1 (run.py loop), same weights, fp32 oracle port + numpy "
                                         "back-projection + metrics, after 1 warm-up frame"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reconstruction(args):
    """BASELINE configs[3]: one step = the whole video through reconstruction.reconstruct."""
    import torch.distributed as dist
    from dav2_b200 import _lib, reconstruction, sharding, weights
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    from dav2_b200.pose_estimation_model import PoseEstimationNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c = CONFIGS[4]
    N, B, S = args.frames, args.batch, args.size
    HW = S * S
    depth_model = DepthAnythingV2(**MODEL_CONFIGS[args.encoder], max_depth=20.0, precision=args.precision)
    weights.randomize_(depth_model, seed=0)
    depth_model = depth_model.to(dev).eval()
    weights.calibrate_(depth_model, dev)
    torch.manual_seed(0)
    pose_model = PoseEstimationNet(8, precision=args.precision).to(dev).eval()
    a, b = sharding.frame_range(N, rank, world)
    hi = min(b + 1, N)
    # this rank's shard + halo; frame f is seeded by its index, so shards are reproducible for any GPU count
    g = torch.Generator(device=dev)
    frames = torch.empty(N if world == 1 else hi - a, 3, S, S, dtype=torch.float32, device=dev)
    base = 0 if world == 1 else a
    for f in range(base, base + frames.shape[0]):
        g.manual_seed(1234 + f)
        frames[f - base] = torch.randn(3, S, S, generator=g, device=dev)
    k4 = k_for(S)

    class _Shard:
        """reconstruct() indexes the video by GLOBAL frame number; serve those indices from the local shard + halo."""

        def __init__(self, t, first, total):
            self.t, self.first, self.shape = t, first, (total,) + tuple(t.shape[1:])

        def __getitem__(self, sl):
            return self.t[sl.start - self.first:sl.stop - self.first]

    video = frames if world == 1 else _Shard(frames, a, N)

    def one_pass(src):
        return reconstruction.reconstruct(src, depth_model, pose_model, k4, scale=0.01, batch=B, rank=rank, world=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = one_pass(video)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.profile_enable(True)
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = one_pass(video)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - n0
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = N / (ms_step / 1e3)
    valid_points = out["counts"].sum().to(torch.int64)
    if world > 1:
        dist.all_reduce(valid_points)

    # e2e: the shard comes from pinned host memory (reconstruct uploads batch by batch) and the product (depth, world-frame
    # cloud, mask, counts, trajectory) is copied back to pinned host memory inside the timed region
    h_frames = torch.empty(frames.shape, dtype=torch.float32).pin_memory()
    h_frames.copy_(frames.cpu())
    nloc = b - a
    h_depth = torch.empty(nloc, S, S, dtype=torch.float32).pin_memory()
    h_xyz = torch.empty(nloc, HW, 3, dtype=torch.float32).pin_memory()
    h_valid = torch.empty(nloc, HW, dtype=torch.uint8).pin_memory()
    h_abs = torch.empty(N, 7, dtype=torch.float32).pin_memory()
    h_counts = torch.empty(nloc, dtype=torch.int32).pin_memory()
    del frames, video
    torch.cuda.empty_cache()
    hvideo = h_frames if world == 1 else _Shard(h_frames, a, N)

    def e2e_pass():
        o = one_pass(hvideo)
        h_depth.copy_(o["depth"], non_blocking=True)
        h_xyz.copy_(o["xyz"], non_blocking=True)
        h_valid.copy_(o["valid"], non_blocking=True)
        h_counts.copy_(o["counts"], non_blocking=True)
        h_abs.copy_(o["abs"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        e2e_pass()
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    tt = torch.tensor([max(ev0.elapsed_time(ev1), wall_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = N / (float(tt) / args.steps / 1e3)
    if rank != 0:
        dist.destroy_process_group()
        return
    sustained, burst, hbm, peak_src = load_peaks()
    mm = {k: prof.get(k, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0}) for k in ("gemm_tcgen05", "conv_tcgen05")}
    dom_ms = mm["gemm_tcgen05"]["ms"] + mm["conv_tcgen05"]["ms"]
    dom_fl = mm["gemm_tcgen05"]["flops"] + mm["conv_tcgen05"]["flops"]
    dom_n = mm["gemm_tcgen05"]["launches"] + mm["conv_tcgen05"]["launches"]
    achieved = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic, traffic_file = measured_traffic()
    gf = GFLOP_PER_FRAME[(args.encoder, S)] + POSE_GFLOP_PER_PAIR if (args.encoder, S) in GFLOP_PER_FRAME else None
    nl0 = sharding.frame_range(N, 0, world)
    outj = {
        "metric": c["metric"], "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16" if args.precision == "fp16" else "bf16", "data": "synthetic",
        "config": workload_config(args, world), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int((min(nl0[1] + 1, N) - nl0[0]) * 3 * HW * 4),
                "d2h_bytes_per_step": int((nl0[1] - nl0[0]) * HW * 17 + (nl0[1] - nl0[0]) * 4 + N * 28),
                "note": "per rank: pinned host frames of the shard + halo -> H2D batch by batch inside reconstruct -> D2H of depth, "
                        "world-frame xyz, validity mask, counts and the trajectory into pinned host memory"},
        "gpu_launches": int(launches), "valid_points": int(valid_points), "trajectory_len": int(out["abs"].shape[0]),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (linear + implicit-GEMM conv instantiations; depth + pose networks)",
                     "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained if sustained else None,
                     "traffic": traffic, "traffic_source": traffic_file, "csrc_sha16": csrc_sha16(),
                     "launches_per_step": dom_n / args.steps, "avg_launch_ms": dom_ms / max(dom_n, 1),
                     "peak_source": f"{peak_src} MEASURED_PEAKS.json bf16_tflops_sustained",
                     "pipeline_frac": value / world * gf / 1e3 / sustained if gf else None},
        "kernels": {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                        "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None}
                    for k, v in prof.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        times, cores, threads = cpu_reference_frames(4, args.encoder, S, None, 1, 3, budget_s=30.0)
        fps = 1.0 / float(np.mean(times))
        outj["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "host_cores": cores,
                                "sample": f"{len(times)} frames, batch 1: fp32 oracle depth + ResNet-18 pose on the stacked pair + "
                                          "compose_poses + numpy back-projection, after 1 warm-up frame"}
    print(json.dumps(outj))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
