"""BASELINE config 4 on one GPU: the full reconstruction pass (vitl depth for every frame, ResNet-18 relative poses for
every consecutive pair, the pose chain, world-frame back-projection) through dav2_b200.reconstruction.reconstruct.
Synthetic frames, random-init weights.  Prints one JSON line for DESIGN.md (not the driver's bench)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dav2_b200 import reconstruction, weights
from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
from dav2_b200.pose_estimation_model import PoseEstimationNet

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = "cuda"
depth_model = DepthAnythingV2(**MODEL_CONFIGS["vitl"], max_depth=20.0)
weights.randomize_(depth_model, seed=0)
depth_model = depth_model.to(dev).eval()
weights.calibrate_(depth_model, dev)
pose_model = PoseEstimationNet(8).to(dev).eval()
g = torch.Generator(device=dev).manual_seed(0)
frames = torch.randn(N, 3, 518, 518, generator=g, device=dev)
k4 = (170.1677, 169.8526, 194.7248, 198.2624)
out = reconstruction.reconstruct(frames[:batch + 1], depth_model, pose_model, k4, batch=batch)  # warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = reconstruction.reconstruct(frames, depth_model, pose_model, k4, batch=batch)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"config": "BASELINE configs[3]: full reconstruction pass, vitl depth + ResNet-18 pose + pose chain + world-frame cloud",
                  "frames": N, "batch": batch, "ms_total": ms, "frames_per_s": N / ms * 1e3,
                  "valid_points": int(out["counts"].sum()), "trajectory_len": int(out["abs"].shape[0])}))
