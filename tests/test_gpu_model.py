"""End-to-end GPU parity: dav2_b200.dpt.DepthAnythingV2 (C ABI, 16-bit tensor-core operands, fp32
accumulate / residual / softmax) against the fp32 CPU oracle on identical seeded NON-DEGENERATE
weights (pre-sigmoid logits std ~2: the depth spans (0, max_depth)) and synthetic frames.

Tolerance (north_star): depth within 1e-2 relative in half-precision mode.  "relative" is measured
against the frame's depth range: max-abs error / max-abs oracle depth, and mean-abs error / mean depth.
The default operand format is fp16 -- the reference's own GPU precision (Lightning "16-mixed",
configs/trainer/default.yaml:4) -- and meets 1e-2.  The optional bf16 format has 8x coarser
rounding (2^-8 per stored activation through ~70 layers into a saturating sigmoid); its measured
error on these deliberately wide-range weights is 2-5e-2 max / <2e-2 mean, asserted at 6e-2 / 2.5e-2."""
import numpy as np
import pytest
import torch

from oracle import dav2_oracle as O

pytestmark = pytest.mark.gpu

DEPTH_TOL = 1e-2
BF16_TOL_MAX, BF16_TOL_MEAN = 6e-2, 2.5e-2


def _build(enc, seed=0, precision="fp16"):
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    oracle = O.build_oracle(enc, seed=seed)
    m = DepthAnythingV2(**MODEL_CONFIGS[enc], max_depth=20.0, precision=precision)
    missing, unexpected = m.load_state_dict(oracle.state_dict(), strict=True)
    assert not missing and not unexpected
    return oracle, m.cuda().eval()


def _check(oracle, m, x, tol_max=DEPTH_TOL, tol_mean=DEPTH_TOL):
    with torch.no_grad():
        ref = oracle(x)
    got = m(x.cuda()).cpu()
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert float(ref.std()) > 0.5, "degenerate oracle output would make parity vacuous"
    err = (got - ref).abs()
    rel_max = float(err.max() / ref.abs().max())
    rel_mean = float(err.mean() / ref.abs().mean())
    assert rel_max < tol_max and rel_mean < tol_mean, (rel_max, rel_mean)
    return rel_max, rel_mean


@pytest.mark.parametrize("enc,B,H,W", [("vits", 1, 518, 518), ("vits", 2, 70, 98), ("vitb", 1, 140, 140), ("vitl", 1, 70, 70),
                                       ("vitl", 1, 518, 518),      # the benchmarked architecture at its real frame size
                                       ("vits", 1, 1036, 1036)])   # BASELINE config 5: 5477 tokens, interpolated pos-embed
def test_forward_matches_oracle(enc, B, H, W):
    oracle, m = _build(enc)
    _check(oracle, m, O.synthetic_frames(B, H, W, seed=11))


@pytest.mark.parametrize("enc,B,H,W", [("vits", 1, 518, 518), ("vitb", 1, 140, 140)])
def test_forward_bf16_mode(enc, B, H, W):
    oracle, m = _build(enc, precision="bf16")
    _check(oracle, m, O.synthetic_frames(B, H, W, seed=11), BF16_TOL_MAX, BF16_TOL_MEAN)


FP32_TOL = 1e-4  # north_star: depth within 1e-4 relative in fp32 mode


@pytest.mark.parametrize("enc,B,H,W", [("vits", 2, 70, 98), ("vits", 1, 518, 518), ("vitb", 1, 140, 140), ("vitl", 1, 70, 70)])
def test_forward_fp32_mode(enc, B, H, W):
    """precision="fp32": the all-fp32 validation engine meets the 1e-4 gate (max AND mean, same measure as above)."""
    oracle, m = _build(enc, precision="fp32")
    rel_max, rel_mean = _check(oracle, m, O.synthetic_frames(B, H, W, seed=11), FP32_TOL, FP32_TOL)
    print(f"fp32 mode {enc} {B}x{H}x{W}: rel max {rel_max:.2e} mean {rel_mean:.2e}")


def test_taps_match_oracle():
    oracle, m = _build("vits", seed=2)
    x = O.synthetic_frames(1, 518, 518, seed=5)
    with torch.no_grad():
        taps = oracle.forward_taps(x)
    m(x.cuda())
    for i, (t, _cls) in enumerate(taps):
        got = m.debug_buffer(f"tap{i}", torch.float16, (1369, 384)).float().cpu()
        err = float((got - t[0]).abs().max())
        assert err < 0.02, (i, err)  # normalised features (std 1, max ~5) after up to 12 fp16-operand blocks


def test_batch_consistency_and_strict_false():
    """Frames are independent: a batch equals its frames run one by one (what sharding relies on)."""
    oracle, m = _build("vits", seed=4)
    x = O.synthetic_frames(3, 98, 126, seed=9).cuda()
    whole = m(x)
    for b in range(3):
        single = m(x[b:b + 1].contiguous())
        # different batch sizes take different kernels (1-CTA RMW vs 2-CTA bulk-reduce epilogue, fma vs mul+add):
        # agreement is at the 16-bit rounding-noise level, not bitwise
        assert float((single[0] - whole[b]).abs().max() / whole[b].abs().max()) < DEPTH_TOL
    # lightning_model.py:130-140: encoder-only partial load with strict=False must work
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    m2 = DepthAnythingV2(**MODEL_CONFIGS["vits"])
    sd = {k: v for k, v in oracle.state_dict().items() if "pretrained" in k}
    res = m2.load_state_dict(sd, strict=False)
    assert all(k.startswith("depth_head.") for k in res.missing_keys) and not res.unexpected_keys
    assert all(("pretrained" in n) == n.startswith("pretrained.") for n, _ in m2.named_parameters())


def test_full_size_batch_properties_vitl():
    """BASELINE configs[2] at its real size (vitl, 64 frames of 518^2 -- too large for the CPU oracle): size-independent
    properties instead.  (1) frames are independent: frames 0 / 31 / 63 of the batch equal the same frames run alone to
    16-bit rounding noise; (2) permuting the batch permutes the output; (3) the fp32 engine agrees on one frame at its
    own 1e-4 gate x the fp16 gate; (4) the output is finite, inside (0, max_depth) and non-degenerate; (5) frames 0 and 63
    OF THE 64-FRAME BATCH match the CPU oracle run on them alone (the oracle needs ~4 s per vitl frame)."""
    from dav2_b200 import weights
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    m = DepthAnythingV2(**MODEL_CONFIGS["vitl"], max_depth=20.0)
    weights.randomize_(m, seed=0)
    m = m.cuda().eval()
    weights.calibrate_(m, "cuda")
    x = O.synthetic_frames(64, 518, 518, seed=21).cuda()
    whole = m(x)
    assert whole.shape == (64, 518, 518) and torch.isfinite(whole).all()
    assert float(whole.min()) > 0.0 and float(whole.max()) < 20.0 and float(whole.std()) > 0.5
    for b in (0, 31, 63):
        single = m(x[b:b + 1].contiguous())
        assert float((single[0] - whole[b]).abs().max() / whole[b].abs().max()) < DEPTH_TOL, b
    perm = torch.arange(63, -1, -1, device="cuda")
    flipped = m(x[perm].contiguous())
    assert float((flipped[perm] - whole).abs().max() / whole.abs().max()) < DEPTH_TOL
    # (5) and against the CPU oracle itself on the first and last frame of the 64-frame batch (same weights)
    oracle = O.DepthAnythingV2("vitl", 256, [256, 512, 1024, 1024], max_depth=20.0).eval()
    oracle.load_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()})
    with torch.no_grad():
        for b in (0, 63):
            ref_b = oracle(x[b:b + 1].cpu())[0]
            err = (whole[b].cpu() - ref_b).abs()
            assert float(err.max() / ref_b.abs().max()) < DEPTH_TOL and float(err.mean() / ref_b.abs().mean()) < DEPTH_TOL, b
    m32 = DepthAnythingV2(**MODEL_CONFIGS["vitl"], max_depth=20.0, precision="fp32")
    m32.load_state_dict(m.state_dict())
    m32 = m32.cuda().eval()
    ref = m32(x[5:6].contiguous())
    assert float((ref[0] - whole[5]).abs().max() / ref.abs().max()) < DEPTH_TOL


def test_infer_image_matches_oracle():
    oracle, m = _build("vits", seed=6)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, size=(95, 120, 3), dtype=np.uint8)  # -> 140 x 182 network input
    ref = oracle.infer_image(img, 140)
    got = m.infer_image(img, 140)
    assert got.shape == ref.shape == (95, 120) and got.dtype == np.float32
    assert np.abs(got - ref).max() / np.abs(ref).max() < DEPTH_TOL


@pytest.mark.parametrize("h,w,size", [(95, 120, 140), (475, 475, 518), (140, 140, 140), (108, 135, 70)])
def test_gpu_preprocess_matches_opencv(h, w, size):
    """image2tensor on the GPU vs the upstream OpenCV path (cv2.INTER_CUBIC on the fp64 RGB/255 image)."""
    from dav2_b200 import ops
    from dav2_b200.dpt import DepthAnythingV2
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 255, size=(h, w, 3), dtype=np.uint8)
    ref, _ = O.image2tensor(img, size)
    nh, nw = DepthAnythingV2.target_size(h, w, size)
    assert (nh, nw) == tuple(ref.shape[-2:])
    got = ops.preprocess_bgr_u8(torch.from_numpy(img).cuda(), nh, nw).cpu()
    assert float((got - ref).abs().max()) < 2e-5


def test_cpu_input_is_rejected():
    from dav2_b200._lib import Dav2Error
    _, m = _build("vits")
    with pytest.raises(Dav2Error):
        m(O.synthetic_frames(1, 70, 70))


def test_lightning_module_test_and_predict():
    """lightning_model.DepthAnythingV2Module.test_step / predict_step (lightning_model.py:289-330, :343-360) driven by
    the trainer.test mirror with the per-procedure collector (test_lightning.py:47-111): per-batch metrics equal the
    oracle's mask + compute_errors on the module's own prediction, the logged epoch means are the means over batches,
    and the prediction itself is the oracle's depth."""
    from oracle import metrics_oracle as MO
    from dav2_b200 import lightning_model as lm
    from dav2_b200.evaluation import ProcedureMetricCollector
    oracle = O.build_oracle("vits", seed=0)
    mod = lm.DepthAnythingV2Module(encoder="vits", min_depth=1e-6, max_depth=20.0)
    mod.load_state_dict({f"model.{k}": v for k, v in oracle.state_dict().items()})
    mod = mod.cuda().eval()
    assert mod.device.type == "cuda"
    g = torch.Generator().manual_seed(5)
    batches = []
    for i, B in enumerate((3, 2)):
        x = O.synthetic_frames(B, 70, 98, seed=20 + i)
        gt = (torch.rand(B, 1, 70, 98, generator=g) * 25.0)          # some above max_depth = 20 -> masked out
        gt[torch.rand(B, 1, 70, 98, generator=g) < 0.05] = 0.0       # and some invalid zeros
        batches.append({"image": x, "depth": gt, "dataset": ["data/SyntheticColon_I"] * B,
                        "id": [f"S{i + 1}_{j:04d}" for j in range(B)]})
    coll = ProcedureMetricCollector()
    logged = lm.test(mod, batches, callbacks=[coll])
    preds = lm.predict(mod, batches)
    per_batch = []
    for b, p in zip(batches, preds):
        assert p.shape == (b["image"].shape[0], 70, 98) and p.is_cuda
        with torch.no_grad():
            ref = oracle(b["image"])
        assert float((p.cpu() - ref).abs().max() / ref.abs().max()) < DEPTH_TOL
        per_batch.append(MO.test_step_metrics(p.cpu().numpy()[:, None], b["depth"].numpy(), 1e-6, 20.0))
    for k in ("d1", "abs_rel", "rmse", "l1"):
        want = float(np.mean([m[k] for m in per_batch]))
        assert abs(logged[f"Test/test_{k}"] - want) <= 1e-4 * max(abs(want), 1e-6), (k, logged, want)
    s = coll.summary()["per_procedure"]
    assert set(s) == {"SyntheticColon_I/Frames_S1", "SyntheticColon_I/Frames_S2"}
    assert len(coll.metrics_by_procedure["SyntheticColon_I/Frames_S1"]) == 3
    for i, name in enumerate(("SyntheticColon_I/Frames_S1", "SyntheticColon_I/Frames_S2")):
        for k in ("d1", "abs_rel", "rmse", "l1"):
            assert abs(s[name][k] - per_batch[i][k]) <= 1e-4 * max(abs(per_batch[i][k]), 1e-6)
