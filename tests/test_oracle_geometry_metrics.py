"""Oracle (oracle/*.py) pinned against fixtures produced by EXECUTING the reference
(scripts/make_golden.py) and against scipy."""
import os

import numpy as np
import pytest

from oracle import geometry_oracle as geo
from oracle import metrics_oracle as met

CM_KEYS = ("rmse", "mae", "abs_rel", "sq_rel", "delta1", "delta2", "delta3")
CE_KEYS = ("d1", "abs_rel", "rmse", "l1")


def test_calculate_metrics_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics_small.npz"))
    for b in range(g["gt"].shape[0]):
        m = met.calculate_metrics(g["gt"][b], g["pred"][b])
        np.testing.assert_allclose([m[k] for k in CM_KEYS], g["calculate_metrics"][b], rtol=2e-6, atol=1e-9)


def test_compute_errors_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics_small.npz"))
    m = met.test_step_metrics(g["pred"][:, None], g["gt"][:, None], 1e-6, 20.0)
    np.testing.assert_allclose([m[k] for k in CE_KEYS], g["compute_errors"], rtol=2e-6, atol=1e-9)


def test_spot_values_from_survey(golden_dir):
    """SURVEY.md 8c spot case: seed 0, gt=clip(Gamma(2,.03),0,.2) 518^2, pred=gt*N(1,.05)."""
    g = np.load(os.path.join(golden_dir, "metrics_small.npz"))
    r0 = np.random.default_rng(0)
    gt0 = np.clip(r0.gamma(2.0, 0.03, size=(518, 518)), 0, 0.2).astype(np.float32)
    pred0 = (gt0 * r0.normal(1.0, 0.05, size=(518, 518))).astype(np.float32)
    m = met.calculate_metrics(gt0, pred0)
    np.testing.assert_allclose([m[k] for k in CM_KEYS], g["spot518_calculate_metrics"], rtol=5e-6)
    c = met.test_step_metrics(pred0, gt0)
    np.testing.assert_allclose([c[k] for k in CE_KEYS], g["spot518_compute_errors"], rtol=5e-6)
    assert abs(m["rmse"] - 0.0036354) < 1e-6 and abs(c["d1"] - 0.9429458) < 1e-6


def test_empty_gives_nan(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics_small.npz"))
    assert g["empty_is_nan"].all()
    m = met.calculate_metrics(np.zeros((4, 4), np.float32), np.ones((4, 4), np.float32))
    assert all(np.isnan(v) for v in m.values())


def test_compose_poses_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "poses_small.npz"))
    np.testing.assert_allclose(geo.compose_poses(g["rel"]), g["abs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(geo.compose_poses(g["rel"], g["init"]), g["abs_init"], rtol=1e-5, atol=1e-6)
    # 3-D input silently uses batch 0 (eval/evaluation.py:306-308)
    np.testing.assert_allclose(geo.compose_poses(np.stack([g["rel"], g["rel"] * 0])), g["abs"], rtol=1e-5, atol=1e-6)


def test_quat_to_matrix_matches_scipy(golden_dir):
    g = np.load(os.path.join(golden_dir, "poses_small.npz"))
    for p, Rm in zip(g["abs_init"], g["rot"]):
        np.testing.assert_allclose(geo.quat_to_matrix(p[3:]), Rm, atol=1e-12)
    T12 = geo.poses_to_T12(g["abs_init"])
    np.testing.assert_allclose(T12[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]].reshape(-1, 3, 3), g["rot"], atol=1e-12)


def test_backproject_matches_reference_formula(golden_dir):
    g = np.load(os.path.join(golden_dir, "backproject_small.npz"))
    pts, valid = geo.backproject(g["depth"], g["k4"])
    np.testing.assert_allclose(pts[valid], g["points"][valid], rtol=1e-13, atol=0)
    assert (valid == (g["depth"].reshape(-1) > 0)).all()
    w, _ = geo.backproject(g["depth"], g["k4"], T=g["T"])
    np.testing.assert_allclose(w[valid], g["world"][valid], rtol=1e-12, atol=1e-15)


def test_backproject_open3d_rule():
    d = np.array([[0, 500, 2999, 3000, 65535]], dtype=np.uint16)
    pts, valid = geo.backproject(d, (100.0, 100.0, 2.0, 0.0), depth_scale=1000.0, depth_trunc=3.0)
    assert valid.tolist() == [False, True, True, False, False]
    np.testing.assert_allclose(pts[1], [(1 - 2) * 0.5 / 100, 0.0, 0.5])


def test_parse_intrinsics_comma_and_space():
    a = geo.parse_intrinsics("156.0418,0,178.5604,0,155.7529,181.8043,0,0,1")
    b = geo.parse_intrinsics("156.0418 0 178.5604\n0 155.7529 181.8043\n0 0 1\n")
    assert a == b == geo.SIMCOL_K_475
    k518 = geo.scale_intrinsics(a)
    np.testing.assert_allclose(k518, (170.1677, 169.8526, 194.7248, 198.2624), atol=1e-4)


def test_voxel_down_sample_hand_case():
    """Open3D VoxelDownSample rule on a hand-computed cloud (depth_to_pointcloud.py:357-359)."""
    # min bound (0,0,0) -> voxel origin -0.5; voxel 1.0: x in [-0.5,0.5) -> 0, [0.5,1.5) -> 1, ...
    pts = np.array([[0.0, 0.0, 0.0], [0.4, 0.2, 0.1], [0.6, 0.0, 0.0], [1.4, 0.4, 0.4], [3.0, 3.0, 3.0]])
    cols = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [1.0, 1.0, 1.0], [0.5, 0.5, 0.5]])
    out, oc, idx = geo.voxel_down_sample(pts, 1.0, cols)
    np.testing.assert_array_equal(idx, [[0, 0, 0], [1, 0, 0], [3, 3, 3]])
    np.testing.assert_allclose(out, [[0.2, 0.1, 0.05], [1.0, 0.2, 0.2], [3.0, 3.0, 3.0]], atol=1e-15)
    np.testing.assert_allclose(oc, [[0.5, 0.5, 0], [0.5, 0.5, 1.0], [0.5, 0.5, 0.5]], atol=1e-15)
    # idempotent on an already down-sampled cloud whose means stay in their voxels; count never grows
    rng = np.random.default_rng(0)
    p = rng.normal(size=(5000, 3))
    a, _, ia = geo.voxel_down_sample(p, 0.25)
    assert a.shape[0] == np.unique(ia, axis=0).shape[0] <= 5000
    with pytest.raises(ValueError):
        geo.voxel_down_sample(p, 0.0)
    with pytest.raises(RuntimeError):
        geo.voxel_down_sample(p * 1e6, 1e-6)
    e, ec, _ = geo.voxel_down_sample(np.zeros((0, 3)), 0.1, np.zeros((0, 3)))
    assert e.shape == (0, 3) and ec.shape == (0, 3)
