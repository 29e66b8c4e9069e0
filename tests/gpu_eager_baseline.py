"""GPU baseline 'the kernel to beat' (SURVEY.md 8d): the oracle model (the reference's architecture in plain PyTorch)
run eagerly on the same B200 with fp16 autocast + TF32, the reference's own GPU settings (configs/trainer/default.yaml:4,
test_lightning.py:24).  Not part of the product or of bench.py (it lives under tests/ because it executes the oracle); prints one JSON line
for DESIGN.md.  Usage on the GPU box: python tests/gpu_eager_baseline.py [encoder] [batch] [size]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import dav2_oracle as O

enc = sys.argv[1] if len(sys.argv) > 1 else "vitl"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 518
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
m = O.build_oracle(enc, seed=0).cuda().eval()
x = O.synthetic_frames(B, S, S, seed=1).cuda()
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    for _ in range(2): m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 4
    for _ in range(n): m(x)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"baseline": "torch eager fp16 autocast + TF32 (oracle architecture)", "encoder": enc, "batch": B, "size": S,
                  "ms_per_batch": ms, "frames_per_s": B / ms * 1e3}))
