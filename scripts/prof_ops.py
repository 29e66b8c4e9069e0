"""Operator-shaped launches of the hot kernels at the vitl B=64 sizes, for ncu (GPU box only).
Usage: python scripts/prof_ops.py [gemm|conv|attn|all] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dav2_b200 import _lib
if os.environ.get("PROF_LIB"):  # A/B another build of the library (profiling only): set before the first load()
    _lib.LIB_PATH = os.environ["PROF_LIB"]
from dav2_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dt = torch.float16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def rnd(*shape, scale=1.0, dtype=dt):
    return (torch.randn(*shape, generator=g, device=dev) * scale).to(dtype)

M, D = 64 * 1370, 1024
if what in ("gemm", "all"):
    a = rnd(M, D); w_qkv = rnd(3 * D, D, scale=D ** -0.5); b_qkv = rnd(3 * D, dtype=torch.float32)
    w_fc1 = rnd(4 * D, D, scale=D ** -0.5); b_fc1 = rnd(4 * D, dtype=torch.float32)
    hid = rnd(M, 4 * D); w_fc2 = rnd(D, 4 * D, scale=(4 * D) ** -0.5); b_d = rnd(D, dtype=torch.float32)
    w_proj = rnd(D, D, scale=D ** -0.5); gamma = rnd(D, dtype=torch.float32)
    x = rnd(M, D, dtype=torch.float32)
    for _ in range(reps):
        ops.linear_h16(a, w_qkv, b_qkv, 0)          # qkv   87680 x 3072 x 1024
        ops.linear_h16(a, w_fc1, b_fc1, 1)          # fc1   87680 x 4096 x 1024 + GELU
        ops.linear_resid_(x, hid, w_fc2, b_d, gamma)  # fc2   87680 x 1024 x 4096 + residual
        ops.linear_resid_(x, a, w_proj, b_d, gamma)   # proj  87680 x 1024 x 1024 + residual
    if os.environ.get("DAV2_TIME"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for name, fn, fl in (("qkv", lambda: ops.linear_h16(a, w_qkv, b_qkv, 0), 2.0 * M * 3 * D * D),
                             ("fc1+gelu", lambda: ops.linear_h16(a, w_fc1, b_fc1, 1), 2.0 * M * 4 * D * D),
                             ("fc2+resid", lambda: ops.linear_resid_(x, hid, w_fc2, b_d, gamma), 2.0 * M * 4 * D * D),
                             ("proj+resid", lambda: ops.linear_resid_(x, a, w_proj, b_d, gamma), 2.0 * M * D * D)):
            fn(); torch.cuda.synchronize(); e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("%-10s %.3f ms  %.0f TFLOP/s" % (name, ms, fl / ms / 1e9))
if what in ("conv", "all"):
    B = 64
    x1 = rnd(B, 148, 148, 256); w1 = ops.pack_conv3x3_weight(rnd(256, 256, 3, 3, scale=(9 * 256) ** -0.5))
    bias = rnd(256, dtype=torch.float32)
    x2 = rnd(B, 296, 296, 256); w2 = ops.pack_conv3x3_weight(rnd(128, 256, 3, 3, scale=(9 * 256) ** -0.5))
    bias2 = rnd(128, dtype=torch.float32)
    for _ in range(reps):
        ops.conv3x3_h16(x1, w1, bias, None, None, 2)    # RCU conv1 at 148^2, F=256
        if what == "conv":
            ops.conv3x3_h16(x1, w1, bias, x1, x1, 0, want_relu=True)  # RCU conv2: two skip adds + dual output
        ops.conv3x3_h16(x2, w2, bias2, None, None, 0)   # output_conv1 at 296^2, 256 -> 128
    if os.environ.get("DAV2_TIME"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for name, fn, fl in (("RCU conv 256->256 @148^2", lambda: ops.conv3x3_h16(x1, w1, bias, None, None, 2), 2.0 * B * 148 * 148 * 256 * 256 * 9),
                             ("output_conv1 256->128 @296^2", lambda: ops.conv3x3_h16(x2, w2, bias2, None, None, 0), 2.0 * B * 296 * 296 * 256 * 128 * 9)):
            fn(); torch.cuda.synchronize(); e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("%-30s %.3f ms  %.0f TFLOP/s" % (name, ms, fl / ms / 1e9))
if what == "attn5477":  # BASELINE config 5: 16 frames of 1036^2 -> 5477 tokens per image
    qkv5 = rnd(16 * 5477, 3 * D, scale=float(os.environ.get("DAV2_QKV_SCALE", "1.0")))
    for _ in range(reps):
        ops.attention_h16(qkv5, 16, 5477, D)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ops.attention_h16(qkv5, 16, 5477, D)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("attention 16 x 5477 tokens: %.3f ms  %.0f TFLOP/s" % (ms, 4 * 16 * 16 * 5477 * 5477 * 64 / ms / 1e9))
if what in ("attn", "all", "attntrace"):
    qkv = rnd(M, 3 * D, scale=float(os.environ.get("DAV2_QKV_SCALE", "1.0")))
    for _ in range(reps):
        ops.attention_h16(qkv, 64, 1370, D)
    if os.environ.get("DAV2_TIME"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): ops.attention_h16(qkv, 64, 1370, D)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("attention EMU=%s: %.3f ms  %.0f TFLOP/s" % (os.environ.get("DAV2_ATTN_EMU", "default"), ms, 4 * 64 * 16 * 1370 * 1370 * 64 / ms / 1e9))
if what == "attntrace":
    # in-kernel timeline (library built with -DATTN_TRACE, selected through PROF_LIB)
    import ctypes, numpy as np
    from dav2_b200 import _lib
    lib = _lib.load()
    trace = torch.zeros(11 * 11 * 16, dtype=torch.int64, device=dev)
    ops.attention_h16(qkv, 64, 1370, D); torch.cuda.synchronize()
    lib.dav2_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
    assert lib.dav2_debug_set_attn_trace(trace.data_ptr()) == 0
    ops.attention_h16(qkv, 64, 1370, D); torch.cuda.synchronize()
    np.save("gpurun_out/attn_trace.npy", trace.cpu().numpy().reshape(11, 11, 16))
if what in ("geom",):
    B, H, W = 64, 518, 518
    depth = torch.rand(B, H, W, generator=g, device=dev) * 20.0
    gt = torch.rand(B, H, W, generator=g, device=dev)
    T12 = torch.eye(4, dtype=torch.float64)[:3].reshape(1, 12).repeat(B, 1) + 0.01
    for _ in range(reps):
        xyz, valid, counts = ops.backproject(depth, (170.1677, 169.8526, 194.7248, 198.2624), T12)
        ops.depth_metric_partials(depth, gt, 1e-6, 20.0, 0, False)
        ops.depth_metric_partials(depth, gt, 1e-6, 20.0, 1, True)
        ops.backproject_metrics(depth, gt, (170.1677, 169.8526, 194.7248, 198.2624), T12, 1e-6, 20.0, out_xyz=xyz)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    k4dev = torch.tensor([170.1677, 169.8526, 194.7248, 198.2624], dtype=torch.float64, device=dev)
    T12 = T12.to(dev)
    for name, fn in (("backproject", lambda: ops.backproject(depth, k4dev, T12, out_xyz=xyz)),
                     ("metrics0", lambda: ops.depth_metric_partials(depth, gt, 1e-6, 20.0, 0, False)),
                     ("backproject+metrics fused (21 B/px)", lambda: ops.backproject_metrics(depth, gt, k4dev, T12, 1e-6, 20.0, out_xyz=xyz))):
        fn(); torch.cuda.synchronize()
        ev[0].record()
        for _ in range(20): fn()
        ev[1].record(); torch.cuda.synchronize()
        us = ev[0].elapsed_time(ev[1]) / 20 * 1e3
        bpp = {"backproject": 17.0, "metrics0": 8.0}.get(name, 21.0)
        print(name, "us/launch (warm, back to back) %.1f  -> %.0f GB/s algorithmic" % (us, B * H * W * bpp / us / 1e3))
    pts = xyz[:8].reshape(-1, 3)
    ev[0].record(); q, _ = ops.voxel_downsample(pts, 0.01, valid=valid[:8].reshape(-1)); ev[1].record(); torch.cuda.synchronize()
    print("voxel 8 frames", pts.shape[0], "->", q.shape[0], "ms", ev[0].elapsed_time(ev[1]))
torch.cuda.synchronize()
print("ok")
