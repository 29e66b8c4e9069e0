import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dav2_b200 import ops
x = (torch.randn(64, 296, 296, 128, device="cuda") ).half()
for _ in range(2): y = ops.bilinear_nhwc_h16(x, 518, 518)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = ops.bilinear_nhwc_h16(x, 518, 518)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("bilinear 296->518 x128ch B=64: %.3f ms  %.0f GB/s" % (ms, (x.numel() + y.numel()) * 2 / ms / 1e6))
