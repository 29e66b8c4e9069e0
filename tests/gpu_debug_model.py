"""Stage-by-stage comparison of the engine's internal buffers against the fp32 oracle (GPU box only).
Usage: python scripts/gpu_debug_model.py [enc] [H] [W] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import dav2_oracle as O
from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2

enc = sys.argv[1] if len(sys.argv) > 1 else "vits"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 70
W = int(sys.argv[3]) if len(sys.argv) > 3 else 98
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
prec = sys.argv[5] if len(sys.argv) > 5 else "fp16"
HD = torch.bfloat16 if prec == "bf16" else torch.float16
oracle = O.build_oracle(enc, seed=0)
m = DepthAnythingV2(**MODEL_CONFIGS[enc], precision=prec).cuda().eval()
m.load_state_dict(oracle.state_dict())
x = O.synthetic_frames(B, H, W, seed=11)
ph, pw = H // 14, W // 14
P = ph * pw
D = oracle.pretrained.embed_dim
Fe = MODEL_CONFIGS[enc]["features"]
oc = MODEL_CONFIGS[enc]["out_channels"]

cap = {}
def hook(name):
    def f(mod, inp, out):
        cap[name] = out.detach()
    return f
s = oracle.depth_head.scratch
for i in range(4):
    oracle.depth_head.projects[i].register_forward_hook(hook(f"proj{i}"))
    oracle.depth_head.resize_layers[i].register_forward_hook(hook(f"lvl{i}"))
    getattr(s, f"layer{i+1}_rn").register_forward_hook(hook(f"rn{i}"))
    getattr(s, f"refinenet{i+1}").register_forward_hook(hook(f"path{i+1}"))
s.output_conv1.register_forward_hook(hook("out1"))
oracle.pretrained.blocks[0].register_forward_hook(hook("blk0"))
with torch.no_grad():
    ref = oracle(x)
    tok0 = oracle.pretrained.prepare_tokens(x)
    taps = oracle.forward_taps(x)
got = m(x.cuda()).cpu()

def rep(name, g, r):
    g = g.float().cpu(); r = r.float()
    e = (g - r).abs()
    print(f"{name:12s} shape {tuple(r.shape)} max|ref| {float(r.abs().max()):9.4f} max err {float(e.max()):9.5f} mean err {float(e.mean()):9.6f} rel {float(e.max()/r.abs().max().clamp_min(1e-9)):8.5f}")

hh = [4 * ph, 2 * ph, ph, (ph + 1) // 2]; ww = [4 * pw, 2 * pw, pw, (pw + 1) // 2]
for i, (t, _c) in enumerate(taps):
    rep(f"tap{i}", m.debug_buffer(f"tap{i}", HD, (B, P, D)), t)
for i in range(4):
    rep(f"proj{i}", m.debug_buffer(f"proj{i}", HD, (B, ph, pw, oc[i])), cap[f"proj{i}"].permute(0, 2, 3, 1))
    if i != 2:
        nm = f"lvl{i}"
        rep(nm, m.debug_buffer(nm, HD, (B, hh[i], ww[i], oc[i])), cap[f"lvl{i}"].permute(0, 2, 3, 1))
    rep(f"rn{i}", m.debug_buffer(f"rn{i}", HD, (B, hh[i], ww[i], Fe)), cap[f"rn{i}"].permute(0, 2, 3, 1))
for i in (4, 3, 2, 1):
    r = cap[f"path{i}"].permute(0, 2, 3, 1)
    rep(f"path{i}", m.debug_buffer(f"path{i}", HD, tuple(r.shape)), r)
r = cap["out1"].permute(0, 2, 3, 1)
rep("out1", m.debug_buffer("out1", HD, tuple(r.shape)), r)
rep("depth", got, ref)
print("ref depth mean/std", float(ref.mean()), float(ref.std()))
