"""Precision study on the GPU box (not a pytest file; executes the oracle, hence under tests/).

For one (encoder, size) it compares against the fp32 oracle (torch, TF32 off, on the GPU):
  * the engine in fp16-operand and bf16-operand mode,
  * the SAME oracle under torch.autocast(fp16 / bf16) -- i.e. what the reference's own AMP path loses at that precision,
with three measures: of-range (max|d|/max depth, mean|d|/mean depth), per-pixel relative (median / p99 of |d|/depth over
pixels with depth > 1 % of the range) and pre-sigmoid logits (max / mean abs).  Prints one JSON line per row.
Usage: python tests/gpu_precision_study.py [encoder] [size] [batch] [seed]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
from oracle import dav2_oracle as O

enc = sys.argv[1] if len(sys.argv) > 1 else "vitl"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 518
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 0
MD = 20.0
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def stats(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    err = (got - ref).abs()
    sel = ref > 0.01 * MD
    rel = (err[sel] / ref[sel])
    lg = torch.log(got.clamp(1e-9, MD - 1e-9) / (MD - got.clamp(1e-9, MD - 1e-9)))
    lr = torch.log(ref.clamp(1e-9, MD - 1e-9) / (MD - ref.clamp(1e-9, MD - 1e-9)))
    mid = (ref > 0.02 * MD) & (ref < 0.98 * MD)  # logits recovered from depth are ill-conditioned at the rails
    le = (lg - lr).abs()[mid]
    return {"range_max": float(err.max() / ref.abs().max()), "range_mean": float(err.mean() / ref.abs().mean()),
            "pix_rel_median": float(rel.median()), "pix_rel_p99": float(rel.quantile(0.99)) if rel.numel() < 16e6 else float(rel[::7].quantile(0.99)),
            "pix_rel_max": float(rel.max()), "logit_max": float(le.max()), "logit_mean": float(le.mean()),
            "logit_std_ref": float(lr[mid].std()), "frac_sel": float(sel.double().mean())}


oracle = O.build_oracle(enc, seed=seed).cuda().eval()
x = O.synthetic_frames(B, S, S, seed=11).cuda()
with torch.no_grad():
    ref = oracle(x)
rows = []
for prec in ("fp16", "bf16"):
    m = DepthAnythingV2(**MODEL_CONFIGS[enc], max_depth=MD, precision=prec)
    m.load_state_dict(oracle.state_dict())
    m = m.cuda().eval()
    rows.append({"impl": f"engine {prec}", **stats(m(x), ref)})
    del m
for dt, nm in ((torch.float16, "fp16"), (torch.bfloat16, "bf16")):
    with torch.no_grad(), torch.autocast("cuda", dtype=dt):
        got = oracle(x)
    rows.append({"impl": f"torch autocast {nm} (the reference's AMP path at this precision)", **stats(got.float(), ref)})
for r in rows:
    print(json.dumps({"encoder": enc, "size": S, "batch": B, **r}))
