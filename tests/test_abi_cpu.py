"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/dav2_b200.h declares; host-side logic that needs no GPU."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    from dav2_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "dav2_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dav2_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = built.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/dav2_b200.h but not exported"
    assert declared == set(built.SIGNATURES), declared ^ set(built.SIGNATURES)
    assert b"sm_100a" in lib.dav2_version()


def test_no_fallback_without_gpu(built):
    from dav2_b200 import ops
    from dav2_b200._lib import Dav2Error
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    with pytest.raises(Dav2Error):
        ops.backproject(torch.zeros(1, 4, 4), (1.0, 1.0, 0.0, 0.0))
    m = DepthAnythingV2(**MODEL_CONFIGS["vits"])
    with pytest.raises(Dav2Error):
        m(torch.zeros(1, 3, 28, 28))
    if not torch.cuda.is_available():
        import ctypes as C
        cfg = built.Dav2Config(384, 12, 6, 64, (C.c_int32 * 4)(48, 96, 192, 384), (C.c_int32 * 4)(2, 5, 8, 11), 20.0)
        h = C.c_void_p()
        assert built.load().dav2_create(C.byref(h), C.byref(cfg)) != 0  # fails loudly, no CPU path
        assert len(built.load().dav2_last_error()) > 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "enhanced-3d-reconstruction-in-colonoscopy-using-monocular-depth-and-pose-estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_module_surface_matches_reference_contract():
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    m = DepthAnythingV2(encoder="vitb", features=128, out_channels=[96, 192, 384, 768], max_depth=100.0)
    assert m.encoder == "vitb" and m.max_depth == 100.0 and hasattr(m, "pretrained") and hasattr(m, "depth_head")
    assert sum(p.numel() for p in m.parameters()) == 97470785
    sd = m.state_dict()
    # run.py:134-144 strips a "model." prefix from Lightning checkpoints before load_state_dict
    lightning = {"model." + k: v for k, v in sd.items()}
    stripped = {k[len("model."):]: v for k, v in lightning.items()}
    assert m.load_state_dict(stripped).missing_keys == []
    assert m.eval() is m
    with pytest.raises(ValueError):
        DepthAnythingV2(encoder="vitx")


def test_image2tensor_matches_upstream_rule():
    from dav2_b200.dpt import DepthAnythingV2
    from oracle import dav2_oracle as O
    rng = np.random.default_rng(0)
    for h, w, size in [(475, 475, 518), (95, 120, 140), (108, 135, 70)]:
        img = rng.integers(0, 255, size=(h, w, 3), dtype=np.uint8)
        a, hw = DepthAnythingV2.image2tensor(img, size)
        b, hw2 = O.image2tensor(img, size)
        assert hw == hw2 == (h, w) and a.shape == b.shape
        assert torch.equal(a, b)
        assert tuple(a.shape[-2:]) == O.resize_target(h, w, size)


def test_pos_table_matches_upstream_interpolation():
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    from oracle import dav2_oracle as O
    o = O.build_oracle("vits", calibrate=False)
    m = DepthAnythingV2(**MODEL_CONFIGS["vits"])
    m.load_state_dict(o.state_dict())
    for ph, pw in [(5, 7), (37, 78), (74, 74)]:
        ref = o.pretrained.interpolate_pos_encoding(ph * pw, ph * 14, pw * 14)[0]
        assert torch.allclose(m._pos_table(ph, pw), ref, atol=1e-6)


def test_finalizers_on_host():
    from dav2_b200 import calculate_metrics as cm
    from dav2_b200 import evaluation as ev
    from oracle import metrics_oracle as met
    rng = np.random.default_rng(1)
    gt = np.clip(rng.gamma(2.0, 0.15, size=(2000,)), 1e-3, 1).astype(np.float32)
    pred = (gt * rng.normal(1, 0.07, size=gt.shape)).astype(np.float32)
    d = pred - gt
    t = np.maximum(gt / pred, pred / gt)
    part = np.array([gt.size, np.abs(d).sum(), (np.abs(d) / (gt + np.float32(1e-6))).sum(), (d.astype(np.float64) ** 2).sum(),
                     gt.sum(dtype=np.float64), (t < np.float32(1.1)).sum(), 0, 0], dtype=np.float64)
    got = ev.finalize_compute_errors(torch.from_numpy(part))
    ref = met.compute_errors(pred, gt)
    for k in ref:
        assert abs(float(got[k]) - ref[k]) < 1e-6
    part[5:] = [(t < 1.25).sum(), (t < 1.25 ** 2).sum(), (t < 1.25 ** 3).sum()]
    got2 = cm.finalize_calculate_metrics(part)
    ref2 = met.calculate_metrics(gt, pred)
    for k in ref2:
        assert abs(got2[k] - ref2[k]) < 1e-6
    assert all(np.isnan(v) for v in cm.finalize_calculate_metrics(np.zeros(8)).values())


def test_bench_traffic_is_keyed_to_kernel_sources(tmp_path, monkeypatch):
    """bench.py reports roofline.traffic only from a launch list taken with the CURRENT kernel sources: the committed
    json carries the sha256 of csrc/ and a mismatch yields null instead of a stale constant."""
    import json
    import bench
    sha = bench.csrc_sha16()
    assert len(sha) == 16 and sha == bench.csrc_sha16()
    committed = json.load(open(os.path.join(bench.ROOT, "profiles", "traffic_r02.json")))
    assert committed["csrc_sha16"] == sha, "profiles/traffic_r02.json was taken with other kernel sources: re-run scripts/gpu_profile.sh + summarise_profiles.py"
    val, src = bench.measured_traffic()
    assert src == "traffic_r02.json" and val == committed["gemm_tcgen05_kernel_bytes_per_launch"] > 1e8
    assert bench.measured_traffic("backproject_kernel_bytes_per_launch")[0] == committed["backproject_kernel_bytes_per_launch"]
    # a list from other sources is ignored
    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "traffic_r02.json").write_text(json.dumps({**committed, "csrc_sha16": "0" * 16}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "csrc_sha16", lambda: sha)
    assert bench.measured_traffic() == (None, None)
