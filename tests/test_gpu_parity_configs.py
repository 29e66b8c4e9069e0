"""Round-2 parity cases: the BASELINE configurations the first round covered only by proxy, and the
finer error measures.

Measures (all against the fp32 CPU oracle on identical seeded, NON-degenerate weights; depth spans (0, max_depth)):
  * of-range      max|d| / max depth and mean|d| / mean depth            (the round-1 measure; gate 1e-2)
  * per pixel     |d| / depth over pixels with depth > 1 % of the range: median (gate 1e-2) and p99 (bounded at 2e-2 AND
                  at 1.25x the p99 of the reference's own GPU precision -- the oracle under torch.autocast(fp16) on the
                  same device -- because the tail is set by the logit scale of the synthetic weights: d(depth)/depth =
                  (1 - sigmoid) * d(logit), and 16-bit operands put ~3e-3 * std on a logit whatever the kernel does)
  * logits        |logit - logit_oracle| before the saturating sigmoid, max and mean, reported and bounded
  * taps          the four encoder taps (after the final LayerNorm), all encoders incl. vitl
DESIGN.md section 2 states which measure the north-star's "1e-2 relative" is read on: of-range max, of-range mean and
per-pixel median, all three.
"""
import os

import numpy as np
import pytest
import torch

from oracle import dav2_oracle as O
from oracle import geometry_oracle as geo

pytestmark = pytest.mark.gpu

MD = 20.0
GATE = 1e-2
P99_BOUND = 2e-2


def _build(enc, seed=0, precision="fp16"):
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    oracle = O.build_oracle(enc, seed=seed)
    m = DepthAnythingV2(**MODEL_CONFIGS[enc], max_depth=MD, precision=precision)
    m.load_state_dict(oracle.state_dict())
    return oracle, m.cuda().eval()


def error_stats(got: torch.Tensor, ref: torch.Tensor, md: float = MD) -> dict:
    got, ref = got.double().cpu(), ref.double().cpu()
    err = (got - ref).abs()
    sel = ref > 0.01 * md
    rel = (err[sel] / ref[sel]).flatten()
    if rel.numel() > 4_000_000:  # torch.quantile input limit
        rel_q = rel[:: rel.numel() // 4_000_000 + 1]
    else:
        rel_q = rel
    return {"range_max": float(err.max() / ref.abs().max()), "range_mean": float(err.mean() / ref.abs().mean()),
            "pix_median": float(rel.median()), "pix_p99": float(rel_q.quantile(0.99)), "pix_max": float(rel.max()),
            "frac_pixels": float(sel.double().mean())}


def _check_depth(oracle, m, x, label, gate=GATE, want_logits=True):
    with torch.no_grad():
        ref_logits = oracle.forward_logits(x)
        ref = MD * torch.sigmoid(ref_logits)
    if want_logits:
        m.capture_logits(True)
    got = m(x.cuda())
    st = error_stats(got, ref)
    assert float(ref.std()) > 0.5 and st["frac_pixels"] > 0.5, "degenerate oracle output would make parity vacuous"
    line = f"{label}: " + " ".join(f"{k}={v:.2e}" for k, v in st.items())
    if want_logits:
        B, H, W = got.shape
        lg = m.debug_buffer("logits", torch.float32, (B, H, W)).double().cpu()
        le = (lg - ref_logits.double().reshape(B, H, W)).abs()
        st["logit_max"], st["logit_mean"], st["logit_ref_std"] = float(le.max()), float(le.mean()), float(ref_logits.std())
        line += f" logit_max={st['logit_max']:.2e} logit_mean={st['logit_mean']:.2e} (oracle logit std {st['logit_ref_std']:.2f})"
        # depth = MD * sigmoid(logit): the engine's depth must be consistent with its own logits
        assert float((MD * torch.sigmoid(lg) - got.double().cpu()).abs().max()) < 1e-4
    # the reference's own GPU precision beside it: the oracle under fp16 autocast on this device (Lightning "16-mixed")
    og = oracle.cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        amp = og(x.cuda()).float()
    oracle.cpu()
    st_amp = error_stats(amp, ref)
    line += f" | torch fp16 autocast: range_max={st_amp['range_max']:.2e} pix_median={st_amp['pix_median']:.2e} pix_p99={st_amp['pix_p99']:.2e}"
    print(line)
    assert st["range_max"] < gate and st["range_mean"] < gate, line
    assert st["pix_median"] < gate, line
    assert st["pix_p99"] < P99_BOUND and st["pix_p99"] < 1.25 * st_amp["pix_p99"] + 1e-3, line
    return st


# BASELINE configs[1] (vitb @518^2, batch > 1), configs[2] architecture with every measure, configs[4] (vitl @1036^2: 5477
# tokens, interpolated position table; the CPU oracle needs ~1 minute)
@pytest.mark.parametrize("enc,B,S", [("vitb", 2, 518), ("vitl", 1, 518), ("vitl", 1, 1036)])
def test_depth_logits_per_pixel(enc, B, S):
    oracle, m = _build(enc)
    st = _check_depth(oracle, m, O.synthetic_frames(B, S, S, seed=11), f"{enc} B={B} {S}^2 fp16")
    assert st["logit_mean"] < 2e-2 and st["logit_max"] < 0.25  # logits of std ~2: 16-bit operand noise, no structure error


@pytest.mark.parametrize("enc,S", [("vitl", 518), ("vitb", 518), ("vits", 1036)])
def test_all_taps(enc, S):
    """All four encoder taps (final LayerNorm applied, cls dropped) -- vitl taps at blocks 4 / 11 / 17 / 23."""
    oracle, m = _build(enc, seed=2)
    x = O.synthetic_frames(1, S, S, seed=5)
    with torch.no_grad():
        taps = oracle.forward_taps(x)
    m(x.cuda())
    D = oracle.pretrained.embed_dim
    P = (S // 14) ** 2
    for i, (t, _cls) in enumerate(taps):
        got = m.debug_buffer(f"tap{i}", torch.float16, (P, D)).float().cpu()
        err = (got - t[0]).abs()
        rms = float(t[0].pow(2).mean().sqrt())
        # normalised features (rms ~1): the error grows with depth (up to 24 fp16-operand blocks for vitl)
        print(f"{enc} {S}^2 tap{i}: max err {float(err.max()):.3e} mean err {float(err.mean()):.3e} (feature rms {rms:.2f})")
        assert float(err.max()) < 0.04 * max(rms, 1.0) and float(err.mean()) < 4e-3 * max(rms, 1.0), (i, float(err.max()))


def test_bf16_mode_vs_reference_at_bf16():
    """north_star "1e-2 relative (bf16)".  bf16 operands carry 8 mantissa bits (fp16: 11): rounding every activation to
    bf16 costs 8x the fp16 noise whatever the kernel does.  Measured waiver: the bf16 engine is
    held to (a) the bound round 1 measured and (b) being NO WORSE than the reference's own AMP path at bf16 (the oracle
    under torch.autocast(bfloat16)) -- the engine keeps the residual stream, LayerNorm, softmax and every accumulator
    in fp32, which autocast does too."""
    oracle, m = _build("vits", precision="bf16")
    x = O.synthetic_frames(1, 518, 518, seed=11)
    with torch.no_grad():
        ref = oracle(x)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            amp = oracle(x).float()
    got = m(x.cuda())
    ours, theirs = error_stats(got, ref), error_stats(amp, ref)
    print("bf16 engine      :", " ".join(f"{k}={v:.2e}" for k, v in ours.items()))
    print("torch bf16 autocast:", " ".join(f"{k}={v:.2e}" for k, v in theirs.items()))
    assert ours["range_max"] < 6e-2 and ours["range_mean"] < 2.5e-2
    assert ours["range_mean"] < 1.5 * theirs["range_mean"] + 1e-3 and ours["pix_median"] < 1.5 * theirs["pix_median"] + 1e-3


def test_config1_fixture_frame(golden_dir):
    """BASELINE configs[0]: the reference's own frame (left 475^2 crop of FrameBuffer_0051.png, committed under
    tests/golden by scripts/make_golden.py) -> INTER_CUBIC 518^2 -> infer_image (run.py:233-234) -> explicit
    back-projection (depth_to_pointcloud_dav2.py:300-313) with K scaled to 518."""
    import cv2
    from dav2_b200 import ops
    g = np.load(os.path.join(golden_dir, "config1_vits.npz"))
    crop = cv2.imread(os.path.join(golden_dir, "FrameBuffer_0051_left475.png"))
    assert crop.shape == (475, 475, 3)
    img518 = cv2.resize(crop, (518, 518), interpolation=cv2.INTER_CUBIC)
    oracle, m = _build("vits")
    stride = int(g["stride"])
    # the oracle run here reproduces the committed sample (thread-count dependent summation order only)
    ref = oracle.infer_image(img518, 518)
    assert np.abs(ref[::stride, ::stride] - g["depth_sub"]).max() < 1e-3 * MD
    got = m.infer_image(img518, 518)
    assert got.shape == (518, 518) and got.dtype == np.float32
    st = error_stats(torch.from_numpy(got), torch.from_numpy(ref))
    print("config 1 (fixture frame, vits):", " ".join(f"{k}={v:.2e}" for k, v in st.items()))
    assert st["range_max"] < GATE and st["pix_p99"] < GATE
    assert np.abs(got[::stride, ::stride] - g["depth_sub"]).max() / g["depth_sub"].max() < GATE
    # run.py path on the 475^2 frame itself (resize to 518 inside infer_image, bilinear back to 475)
    got475, ref475 = m.infer_image(crop, 518), oracle.infer_image(crop, 518)
    assert got475.shape == (475, 475)
    assert np.abs(got475 - ref475).max() / np.abs(ref475).max() < GATE
    # points: identical depth in, the reference formula's points out (no pose, no validity filter in that script)
    xyz, valid, counts = ops.backproject(torch.from_numpy(ref)[None].cuda(), tuple(g["k4"]))
    sub = np.zeros((518, 518), bool)
    sub[::stride, ::stride] = True
    pts = xyz[0].cpu().numpy().astype(np.float64)[sub.reshape(-1)]
    refp = g["points_sub"]
    # the committed points were computed from the fixture-time depth: compare through the depth each was built from
    z_now, z_fix = ref[::stride, ::stride].reshape(-1).astype(np.float64), g["depth_sub"].reshape(-1).astype(np.float64)
    scale = (z_now / z_fix)[:, None]
    rel = np.linalg.norm(pts - refp * scale, axis=1) / np.linalg.norm(refp * scale, axis=1)
    assert rel.max() < 1e-5 and int(counts[0]) == 518 * 518 and bool(valid.all())
    full, _ = geo.backproject(ref, tuple(g["k4"]), np.eye(4))
    rel_full = np.linalg.norm(xyz[0].cpu().numpy() - full, axis=1) / np.linalg.norm(full, axis=1)
    assert rel_full.max() < 1e-5


def _read_ply(path):
    with open(path, "rb") as f:
        raw = f.read()
    head, body = raw.split(b"end_header\n", 1)
    n = int([l for l in head.decode().splitlines() if l.startswith("element vertex")][0].split()[-1])
    rec = np.frombuffer(body, dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")], count=n)
    return np.stack([rec["x"], rec["y"], rec["z"]], 1), np.stack([rec["r"], rec["g"], rec["b"]], 1)


def test_pointcloud_dav2_frame_loop(golden_dir, tmp_path):
    """depth_to_pointcloud_dav2.py:247-326 end to end on the reference's frame and a second frame of another size:
    infer_image(image, height) -> every pixel back-projected with the cam file's K -> PLY.  Depth against the oracle's
    infer_image, points against the script's fp64 numpy formula on that depth, colours = the file's RGB."""
    import cv2
    from dav2_b200 import depth_to_pointcloud_dav2 as pc
    crop = cv2.imread(os.path.join(golden_dir, "FrameBuffer_0051_left475.png"))
    cv2.imwrite(str(tmp_path / "frame_a.png"), crop)
    cv2.imwrite(str(tmp_path / "frame_b.png"), crop[100:324, 50:386])   # 224 x 336
    cv2.imwrite(str(tmp_path / "frame_c.png"), crop[::-1].copy())       # same size as a: shares its batch
    cam = tmp_path / "cam.txt"
    cam.write_text("156.0432 0 178.5605 0 155.7543 181.8043 0 0 1\n")
    oracle, m = _build("vits")
    names = [str(tmp_path / f"frame_{c}.png") for c in "abc"]
    out = pc.process_frames(m, names, str(tmp_path / "out"), cam_file=str(cam), batch=3)
    assert [os.path.basename(p) for p in out] == ["frame_a.ply", "frame_b.ply", "frame_c.ply"]
    fx, fy, cx, cy = 156.0432, 155.7543, 178.5605, 181.8043
    for name, ply in zip(names, out):
        img = cv2.imread(name)
        h, w = img.shape[:2]
        pts, cols = _read_ply(ply)
        assert pts.shape == (h * w, 3)
        assert np.array_equal(cols.reshape(h, w, 3), img[:, :, ::-1])
        ref = oracle.infer_image(img, h)
        z = pts[:, 2].reshape(h, w)
        assert np.abs(z - ref).max() / np.abs(ref).max() < GATE
        assert np.array_equal(z.astype(np.float32), m.infer_image(img, h))  # batched == per-frame infer_image
        x, y = np.meshgrid(np.arange(w), np.arange(h))
        want = np.stack(((x - cx) / fx * z, (y - cy) / fy * z, z), -1).reshape(-1, 3)
        rel = np.linalg.norm(pts - want, axis=1) / np.linalg.norm(want, axis=1)
        assert rel.max() < 1e-6, rel.max()
