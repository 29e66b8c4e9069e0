"""GPU parity (through the C ABI) for the HBM-bound kernels: back-projection + SE(3), metric
reductions, pose chain -- against the oracle and the reference-generated golden fixtures."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import geometry_oracle as geo
from oracle import metrics_oracle as met

pytestmark = pytest.mark.gpu

CM_KEYS = ("rmse", "mae", "abs_rel", "sq_rel", "delta1", "delta2", "delta3")
CE_KEYS = ("d1", "abs_rel", "rmse", "l1")
POINT_RTOL = 1e-5   # north_star: point coordinates within 1e-5 relative given identical depth
METRIC_TOL = 1e-4   # north_star: metric values within 1e-4


def _rel_err_points(a, b):
    """relative error measured on the point norm, per point."""
    na = np.linalg.norm(b, axis=-1)
    return np.linalg.norm(a - b, axis=-1) / np.maximum(na, 1e-12)


def test_backproject_golden_fixture(golden_dir):
    from dav2_b200 import ops
    g = np.load(os.path.join(golden_dir, "backproject_small.npz"))
    d = torch.from_numpy(g["depth"])[None].cuda()
    xyz, valid, counts = ops.backproject(d, tuple(g["k4"]))
    v = valid[0].bool().cpu().numpy()
    assert (v == (g["depth"].reshape(-1) > 0)).all() and int(counts[0]) == int(v.sum())
    assert _rel_err_points(xyz[0].cpu().numpy()[v].astype(np.float64), g["points"][v]).max() < POINT_RTOL
    T12 = torch.from_numpy(g["T"][:3, :4].reshape(1, 12))
    w, _, _ = ops.backproject(d, tuple(g["k4"]), T12)
    assert _rel_err_points(w[0].cpu().numpy()[v].astype(np.float64), g["world"][v]).max() < POINT_RTOL


@pytest.mark.parametrize("B,H,W", [(1, 1, 1), (2, 7, 5), (3, 37, 53), (2, 518, 518),
                                   (2, 4, 1), (1, 8, 2), (1, 4, 3), (1, 2, 6)])  # strips narrower than one 4-pixel quad
def test_backproject_vs_oracle(B, H, W):
    from dav2_b200 import ops
    rng = np.random.default_rng(B * 1000 + H)
    depth = np.clip(rng.gamma(2.0, 0.03, size=(B, H, W)), 0, 0.2).astype(np.float32)
    depth[rng.random((B, H, W)) < 0.05] = 0.0
    if H * W > 4:
        depth[0].reshape(-1)[1] = np.nan
        depth[0].reshape(-1)[2] = np.inf
        depth[0].reshape(-1)[3] = -1.0
    k4 = geo.scale_intrinsics(geo.SIMCOL_K_475, 475, max(W, 2))
    Ts = []
    for b in range(B):
        q = rng.normal(size=4)
        Ts.append(geo.make_transform(rng.normal(size=3), q))
    T12 = torch.from_numpy(np.stack([T[:3, :4].reshape(-1) for T in Ts]))
    xyz, valid, counts = ops.backproject(torch.from_numpy(depth).cuda(), k4, T12)
    for b in range(B):
        ref, rv = geo.backproject(depth[b], k4, Ts[b])
        got = xyz[b].cpu().numpy().astype(np.float64)
        gv = valid[b].bool().cpu().numpy()
        assert (gv == rv).all()
        assert int(counts[b]) == int(rv.sum())
        if rv.any():
            assert _rel_err_points(got[rv], ref[rv]).max() < POINT_RTOL
        assert (got[~rv] == 0).all()


def test_backproject_open3d_scale_trunc():
    from dav2_b200 import ops
    d = torch.tensor([[[0.0, 500.0, 2999.0, 3000.0, 65535.0, 1.0, 2.0]]]).cuda()
    xyz, valid, counts = ops.backproject(d, (100.0, 100.0, 2.0, 0.0), None, 1000.0, 3.0)
    assert valid[0].tolist() == [0, 1, 1, 0, 0, 1, 1] and int(counts[0]) == 4
    np.testing.assert_allclose(xyz[0, 1].cpu().numpy(), [(1 - 2) * 0.5 / 100, 0.0, 0.5], rtol=1e-6)


def test_backproject_full_size_properties():
    """BASELINE full size (64 x 518^2): size-independent properties instead of an oracle pass."""
    from dav2_b200 import ops
    B, H, W = 64, 518, 518
    g = torch.Generator(device="cuda").manual_seed(5)
    depth = torch.rand(B, H, W, generator=g, device="cuda") * 0.2
    depth[:, ::7, ::5] = 0.0
    k4 = geo.scale_intrinsics(geo.SIMCOL_K_475)
    xyz, valid, counts = ops.backproject(depth, k4)
    assert torch.equal(valid.view(B, H, W).bool(), depth > 0)
    assert torch.equal(counts.long(), (depth > 0).view(B, -1).sum(1))
    # z channel is the depth itself (identity pose); x/z is the pixel ray, independent of depth (linearity)
    assert torch.equal(xyz[..., 2].view(B, H, W), depth)
    xyz2, _, _ = ops.backproject(depth * 2, k4)
    m = valid.bool()
    assert torch.allclose(xyz2[m], 2 * xyz[m], rtol=1e-6, atol=0)
    # rigid transform preserves pairwise distances: rotate by a pose and compare norms about the centroid
    T = geo.make_transform([0.3, -0.2, 0.1], [0.1, 0.2, 0.3, 0.9])
    T12 = torch.from_numpy(np.tile(T[:3, :4].reshape(1, 12), (B, 1)))
    w, _, _ = ops.backproject(depth, k4, T12)
    a = (xyz[0][m[0]].double() - torch.tensor(0.0)).norm(dim=1)
    b = (w[0][m[0]].double() - torch.tensor(T[:3, 3]).cuda()).norm(dim=1)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_metrics_golden_fixture(golden_dir):
    from dav2_b200 import calculate_metrics as cm
    from dav2_b200 import evaluation as ev
    g = np.load(os.path.join(golden_dir, "metrics_small.npz"))
    rows = cm.calculate_metrics_batch(g["gt"], g["pred"])
    for b, m in enumerate(rows):
        np.testing.assert_allclose([m[k] for k in CM_KEYS], g["calculate_metrics"][b], rtol=METRIC_TOL, atol=1e-7)
    one = cm.calculate_metrics(g["gt"][1], g["pred"][1])
    np.testing.assert_allclose([one[k] for k in CM_KEYS], g["calculate_metrics"][1], rtol=METRIC_TOL, atol=1e-7)
    gt, pred = torch.from_numpy(g["gt"]).cuda()[:, None], torch.from_numpy(g["pred"]).cuda()[:, None]
    fused = ev.test_step_metrics(pred, gt, 1e-6, 20.0)
    np.testing.assert_allclose([float(fused[k]) for k in CE_KEYS], g["compute_errors"], rtol=METRIC_TOL, atol=1e-7)
    mask = (gt >= 1e-6) & (gt <= 20.0)
    plain = ev.compute_errors(pred[mask].flatten(), gt[mask].flatten())
    np.testing.assert_allclose([float(plain[k]) for k in CE_KEYS], g["compute_errors"], rtol=METRIC_TOL, atol=1e-7)
    assert plain["d1"].dim() == 0 and plain["d1"].is_cuda and plain["d1"].dtype == torch.float32


def test_metrics_empty_and_nonfinite():
    from dav2_b200 import calculate_metrics as cm
    from dav2_b200 import evaluation as ev
    m = cm.calculate_metrics(np.zeros((4, 4), np.float32), np.ones((4, 4), np.float32))
    assert all(math.isnan(v) for v in m.values())
    pred = torch.tensor([1.0, float("nan"), 2.0, float("inf")]).cuda()
    gt = torch.tensor([1.0, 1.0, 2.0, 1.0]).cuda()
    out = ev.compute_errors(pred, gt)
    ref = met.compute_errors(pred.cpu().numpy(), gt.cpu().numpy())
    assert math.isnan(float(out["l1"])) == math.isnan(ref["l1"])
    np.testing.assert_allclose(float(out["d1"]), ref["d1"], rtol=1e-6)
    e = ev.compute_errors(torch.zeros(0).cuda(), torch.zeros(0).cuda())
    assert math.isnan(float(e["rmse"]))


@pytest.mark.parametrize("B,H,W", [(1, 3, 5), (4, 70, 98), (32, 518, 518)])
def test_metrics_vs_oracle(B, H, W):
    from dav2_b200 import calculate_metrics as cm
    from dav2_b200 import evaluation as ev
    rng = np.random.default_rng(B + H)
    gt = np.clip(rng.gamma(2.0, 0.15, size=(B, 1, H, W)), 0, 1).astype(np.float32)
    gt[rng.random(gt.shape) < 0.02] = 0.0
    pred = (np.where(gt > 0, gt, 0.3) * rng.normal(1.0, 0.07, size=gt.shape)).astype(np.float32)
    ref = met.test_step_metrics(pred, gt, 1e-6, 20.0)
    got = ev.test_step_metrics(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), 1e-6, 20.0)
    for k in CE_KEYS:
        assert abs(float(got[k]) - ref[k]) <= METRIC_TOL * max(1.0, abs(ref[k])), k
    nb = min(B, 3)
    rows = cm.calculate_metrics_batch(gt[:nb, 0], pred[:nb, 0])
    for b in range(nb):
        r = met.calculate_metrics(gt[b, 0], pred[b, 0])
        for k in CM_KEYS:
            assert abs(rows[b][k] - r[k]) <= METRIC_TOL * max(1.0, abs(r[k])), k
    # sharding property (multi-GPU contract): partial sums over frame slices add up to the whole batch
    if B >= 2:
        p, g = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
        whole = ev.metric_partials(p, g)
        parts = ev.metric_partials(p[: B // 2].contiguous(), g[: B // 2].contiguous()) + \
            ev.metric_partials(p[B // 2:].contiguous(), g[B // 2:].contiguous())
        assert torch.allclose(whole, parts, rtol=1e-12)


def test_compose_poses_golden(golden_dir):
    from dav2_b200 import evaluation as ev
    from dav2_b200 import ops
    g = np.load(os.path.join(golden_dir, "poses_small.npz"))
    rel = torch.from_numpy(g["rel"]).cuda()
    np.testing.assert_allclose(ev.compose_poses(rel).cpu().numpy(), g["abs"], rtol=1e-5, atol=1e-6)
    a, T12 = ops.compose_poses(rel, torch.from_numpy(g["init"]).cuda(), want_T12=True)
    np.testing.assert_allclose(a.cpu().numpy(), g["abs_init"], rtol=1e-5, atol=1e-6)
    # 3-D input uses batch 0 only; 1-D input is a single step
    np.testing.assert_allclose(ev.compose_poses(torch.stack([rel, rel * 0])).cpu().numpy(), g["abs"], rtol=1e-5, atol=1e-6)
    assert ev.compose_poses(rel[0]).shape == (2, 7)
    T = T12.cpu().numpy()
    np.testing.assert_allclose(T[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]].reshape(-1, 3, 3), g["rot"], atol=1e-6)
    np.testing.assert_allclose(T[:, [3, 7, 11]], g["abs_init"][:, :3], atol=1e-7)
    # long chain (config 4: ~1000 frames) against the oracle
    rng = np.random.default_rng(1)
    N = 1000
    q = rng.normal(size=(N, 4)); q[:, 3] += 8; q /= np.linalg.norm(q, axis=1, keepdims=True)
    relN = np.concatenate([rng.normal(0, 0.01, size=(N, 3)), q], axis=1).astype(np.float32)
    got = ev.compose_poses(torch.from_numpy(relN).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, geo.compose_poses(relN), rtol=1e-4, atol=1e-5)


def test_point_cloud_api(tmp_path, golden_dir):
    """generate_point_cloud / load_* signatures on files (Open3D semantics: z=d/1000, z>=3 dropped, BGR colours)."""
    import cv2
    from dav2_b200 import depth_to_pointcloud as d2p
    rng = np.random.default_rng(3)
    root = tmp_path / "SyntheticColon_I"
    (root / "Frames_S1").mkdir(parents=True)
    H = W = 32
    depth = rng.integers(0, 4000, size=(H, W)).astype(np.uint16)
    color = rng.integers(0, 255, size=(H, W, 3)).astype(np.uint8)
    cv2.imwrite(str(root / "Frames_S1" / "Depth_0000.png"), depth)
    cv2.imwrite(str(root / "Frames_S1" / "FrameBuffer_0000.png"), color)
    (root / "cam.txt").write_text("20.0,0,15.5,0,21.0,16.5,0,0,1")
    (root / "SavedPosition_S1.txt").write_text("0.1 0.2 0.3\n1 2 3\n")
    (root / "SavedRotationQuaternion_S1.txt").write_text("0.1 0.2 0.3 0.9\n0 0 0 1\n")
    rgb = str(root / "Frames_S1" / "FrameBuffer_0000.png")
    cam, pos, rot = d2p.get_procedure_files(rgb)
    pc = d2p.generate_point_cloud(str(root / "Frames_S1" / "Depth_0000.png"), rgb, cam, pos, rot, 0)
    T = geo.make_transform([0.1, 0.2, 0.3], [0.1, 0.2, 0.3, 0.9])
    np.testing.assert_allclose(d2p.load_transformation(pos, rot, 0), T, atol=1e-12)
    ref, rv = geo.backproject(depth, (20.0, 21.0, 15.5, 16.5), T, depth_scale=1000.0, depth_trunc=3.0)
    assert len(pc) == int(rv.sum())
    assert _rel_err_points(pc.points, ref[rv]).max() < POINT_RTOL
    np.testing.assert_allclose(pc.colors, color.reshape(-1, 3)[rv] / 255.0, atol=1e-6)
    # PointCloud.transform (o3d semantics, in place, returns self): fp64 on the device
    T2 = geo.make_transform([0.5, -0.25, 2.0], [0.3, -0.1, 0.2, 0.9])
    moved = d2p.PointCloud(pc.points_tensor.clone()).transform(T2)
    want_pts = pc.points @ T2[:3, :3].T + T2[:3, 3]
    assert _rel_err_points(moved.points, want_pts).max() < POINT_RTOL
    both = d2p.PointCloud()
    both += pc
    both += pc
    assert len(both) == 2 * len(pc) and both.points.shape == (2 * len(pc), 3)
    d2p.write_ply(str(tmp_path / "c.ply"), both)
    assert os.path.getsize(tmp_path / "c.ply") > 27 * len(both)
    # main (depth_to_pointcloud.py:316-371): frames batched into one back-projection launch == the per-frame clouds fused
    depth1 = rng.integers(0, 4000, size=(H, W)).astype(np.uint16)
    cv2.imwrite(str(root / "Frames_S1" / "Depth_0001.png"), depth1)
    cv2.imwrite(str(root / "Frames_S1" / "FrameBuffer_0001.png"), color[::-1].copy())
    dps = [str(root / "Frames_S1" / f"Depth_000{i}.png") for i in (0, 1)]
    cps = [str(root / "Frames_S1" / f"FrameBuffer_000{i}.png") for i in (0, 1)]
    fused = d2p.main(dps, cps, str(tmp_path / "out"))
    per_frame = d2p.PointCloud()
    for i in (0, 1):
        per_frame += d2p.generate_point_cloud(dps[i], cps[i], cam, pos, rot, i)
    want = per_frame.voxel_down_sample(0.01)
    assert len(fused) == len(want) > 0
    np.testing.assert_allclose(fused.points, want.points, rtol=1e-6, atol=1e-7)
    assert os.path.exists(tmp_path / "out" / "combined_point_cloud.ply")


@pytest.mark.parametrize("n,voxel,with_rgb,with_valid", [(1, 0.01, False, False), (1000, 0.05, True, False),
                                                         (200_000, 0.01, True, True), (3_000_000, 0.01, False, True)])
def test_voxel_downsample_vs_oracle(n, voxel, with_rgb, with_valid):
    """dav2_voxel_downsample == Open3D rule restated in the oracle: same voxel set, same counts, means to fp32 rounding."""
    from dav2_b200 import ops
    rng = np.random.default_rng(n)
    # a colon-like tube, ~0.3 units across, so voxels hold from one to many points
    t = rng.uniform(0, 1.0, n)
    pts = np.stack([0.15 * np.cos(40 * t) + 0.02 * rng.normal(size=n), 0.15 * np.sin(40 * t) + 0.02 * rng.normal(size=n),
                    t * (0.5 if n > 1000 else 0.1)], -1).astype(np.float32)
    rgb = rng.uniform(0, 1, (n, 3)).astype(np.float32) if with_rgb else None
    valid = (rng.uniform(size=n) > 0.1).astype(np.uint8) if with_valid else None
    if with_valid:
        pts[::97] = np.nan  # non-finite rows are dropped too
    keep = np.isfinite(pts).all(-1) & ((valid > 0) if with_valid else True)
    ref, refc, _ = geo.voxel_down_sample(pts[keep], voxel, rgb[keep] if with_rgb else None)
    got, gotc = ops.voxel_downsample(torch.from_numpy(pts).cuda(), voxel, None if rgb is None else torch.from_numpy(rgb).cuda(),
                                     None if valid is None else torch.from_numpy(valid).cuda())
    assert got.shape == (ref.shape[0], 3)
    g = got.cpu().numpy().astype(np.float64)
    # both are sorted by (ix,iy,iz): compare row by row; means agree to one fp32 rounding
    np.testing.assert_allclose(g, ref, rtol=0, atol=1e-7 * max(1.0, np.abs(ref).max()))
    if with_rgb:
        np.testing.assert_allclose(gotc.cpu().numpy(), refc, rtol=0, atol=2e-7)
    # property: point count is conserved through the voxel populations (mean of means weighted == global mean)
    assert got.shape[0] <= int(keep.sum())


def test_voxel_downsample_edges_and_api():
    from dav2_b200 import depth_to_pointcloud as d2p
    from dav2_b200 import ops
    z = torch.zeros(0, 3, device="cuda")
    a, b = ops.voxel_downsample(z, 0.01)
    assert a.shape == (0, 3) and b is None
    with pytest.raises(ValueError):
        ops.voxel_downsample(torch.zeros(4, 3, device="cuda"), 0.0)
    far = torch.tensor([[0.0, 0, 0], [1e6, 0, 0]], device="cuda")
    with pytest.raises(RuntimeError, match="too small"):
        ops.voxel_downsample(far, 1e-3)
    # all points masked out -> empty cloud
    a, _ = ops.voxel_downsample(torch.rand(10, 3, device="cuda"), 0.1, valid=torch.zeros(10, dtype=torch.uint8, device="cuda"))
    assert a.shape == (0, 3)
    # PointCloud.voxel_down_sample (depth_to_pointcloud.py:357-359) incl. colours; idempotence of the voxel SET
    rng = np.random.default_rng(5)
    p = torch.from_numpy(rng.uniform(0, 0.2, (50_000, 3)).astype(np.float32)).cuda()
    c = torch.from_numpy(rng.uniform(0, 1, (50_000, 3)).astype(np.float32)).cuda()
    pc = d2p.PointCloud(p, c)
    ds = pc.voxel_down_sample(voxel_size=0.01)
    ref, refc, _ = geo.voxel_down_sample(p.cpu().numpy(), 0.01, c.cpu().numpy())
    assert len(ds) == ref.shape[0] and ds.colors.shape == ds.points.shape
    np.testing.assert_allclose(ds.points, ref, atol=1e-7)
    np.testing.assert_allclose(ds.colors, refc, atol=2e-7)
    assert len(d2p.PointCloud().voxel_down_sample(0.01)) == 0


def test_calculate_metrics_unmasked_variant():
    """calculate_metrics(gt, pred, mask_invalid=False) (calculate_metrics.py:17-21 skipped): every pixel counts."""
    from dav2_b200 import calculate_metrics as cm
    rng = np.random.default_rng(11)
    gt = rng.gamma(2.0, 0.05, (70, 98)).astype(np.float32) + 1e-3
    pred = (gt * rng.normal(1.0, 0.2, gt.shape)).astype(np.float32).clip(1e-3)
    got = cm.calculate_metrics(gt, pred, mask_invalid=False)
    ref = met.calculate_metrics(gt, pred, mask_invalid=False)
    for k in CM_KEYS:
        assert abs(got[k] - ref[k]) <= METRIC_TOL * max(1.0, abs(ref[k])), k


def test_backproject_gather_multi_destination():
    """dav2_backproject_gather: identical results in every destination at frame_offset (single process, two local
    destinations stand in for peer-mapped buffers); rows outside the written slice stay untouched."""
    from dav2_b200 import ops
    torch.manual_seed(0)
    B, H, W, F, off = 3, 37, 52, 8, 4
    depth = torch.rand(B, H, W, device="cuda") * 5
    depth[0, 0, :7] = 0.0
    k4 = (30.0, 31.0, 25.5, 18.2)
    T12 = (torch.eye(4, dtype=torch.float64)[:3].reshape(1, 12).repeat(B, 1) + 0.01 * torch.arange(B)[:, None]).cuda()
    ref_xyz, ref_valid, ref_counts = ops.backproject(depth, k4, T12)
    dst_xyz = [torch.full((F, H * W, 3), -7.0, device="cuda") for _ in range(2)]
    dst_valid = [torch.full((F, H * W), 9, dtype=torch.uint8, device="cuda") for _ in range(2)]
    dst_counts = [torch.full((F,), -1, dtype=torch.int32, device="cuda") for _ in range(2)]
    ops.backproject_gather(depth, k4, T12, dst_xyz, dst_valid, dst_counts, frame_offset=off)
    for x, v, c in zip(dst_xyz, dst_valid, dst_counts):
        assert torch.equal(x[off:off + B], ref_xyz) and torch.equal(v[off:off + B], ref_valid)
        assert torch.equal(c[off:off + B], ref_counts)
        assert (x[:off] == -7.0).all() and (x[off + B:] == -7.0).all() and (v[:off] == 9).all() and (c[off + B:] == -1).all()
    # pointer-list form without mask / counts
    y = torch.zeros(F, H * W, 3, device="cuda")
    ops.backproject_gather(depth, k4, T12, [y.data_ptr()], None, None, frame_offset=0)
    assert torch.equal(y[:B], ref_xyz)


def test_cloud_gather_single_rank():
    """sharding.CloudGather without a process group: peer alloc / export / views / double buffering / close."""
    from dav2_b200 import ops, sharding
    B, H, W = 2, 28, 42
    cg = sharding.CloudGather(B, H * W, "cuda")
    k4 = (30.0, 31.0, 20.5, 14.2)
    for step in range(3):
        depth = torch.rand(B, H, W, device="cuda") + step
        xyz, valid, counts = cg.backproject(depth, k4)
        cg.complete()
        rx, rv, rc = ops.backproject(depth, k4)
        assert torch.equal(xyz, rx) and torch.equal(valid, rv) and torch.equal(counts, rc)
    assert cg.views(0)[0].data_ptr() != cg.views(1)[0].data_ptr()
    cg.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_cloud_gather_two_gpus():
    """Fused back-projection + gather over peer memory == NCCL all_gather of per-rank clouds (tests/mgpu_gather_check.py)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29631", os.path.join(root, "tests", "mgpu_gather_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "MGPU_GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("B,H,W,pose", [(2, 37, 52, True), (3, 518, 518, True), (1, 70, 98, False), (2, 7, 5, True)])
def test_backproject_metrics_fused_equals_separate(B, H, W, pose):
    """dav2_backproject_metrics: ONE pass == dav2_backproject (bit for bit) + dav2_depth_metrics variant 0 (sums to fp32
    accumulation order), incl. invalid / non-finite depths, the per-frame form and the odd-sized fallback (7x5)."""
    from dav2_b200 import evaluation as ev
    from dav2_b200 import ops
    rng = np.random.default_rng(B * 100 + H)
    gt = np.clip(rng.gamma(2.0, 0.15, size=(B, H, W)), 0, 1).astype(np.float32)
    gt[rng.random((B, H, W)) < 0.02] = 0.0
    depth = (np.where(gt > 0, gt, 0.3) * rng.normal(1.0, 0.07, size=gt.shape)).astype(np.float32)
    depth[rng.random((B, H, W)) < 0.01] = 0.0
    if H * W > 8:
        depth[0].reshape(-1)[1] = np.nan
        depth[0].reshape(-1)[2] = np.inf
        depth[0].reshape(-1)[3] = -1.0
    k4 = geo.scale_intrinsics(geo.SIMCOL_K_475, 475, W)
    T12 = None
    if pose:
        T12 = torch.from_numpy(np.stack([geo.make_transform(rng.normal(size=3), rng.normal(size=4))[:3, :4].reshape(-1) for _ in range(B)]))
    d, g = torch.from_numpy(depth).cuda(), torch.from_numpy(gt).cuda()
    xyz, valid, counts = ops.backproject(d, k4, T12)
    for per_frame in (False, True):
        want = ops.depth_metric_partials(d, g, 1e-6, 20.0, 0, per_frame)
        fx, fv, fc, part = ops.backproject_metrics(d, g, k4, T12, 1e-6, 20.0, per_frame=per_frame)
        assert torch.equal(fx.view(torch.int32), xyz.view(torch.int32)) and torch.equal(fv, valid) and torch.equal(fc, counts)
        assert part.shape == want.shape
        w, p = want.cpu().numpy(), part.cpu().numpy()
        assert np.array_equal(w[..., [0, 5, 6, 7]], p[..., [0, 5, 6, 7]])          # counts are exact
        np.testing.assert_allclose(p[..., 1:5], w[..., 1:5], rtol=2e-6, atol=1e-9)  # fp32 partial sums, fp64 across warps
    # and against the oracle, finalised (the 1e-4 metric gate)
    _, _, _, part = ops.backproject_metrics(d, g, k4, T12, 1e-6, 20.0)
    got = ev.finalize_compute_errors(part)
    ref = met.test_step_metrics(depth[:, None], gt[:, None], 1e-6, 20.0)  # NaN / inf predictions poison some entries: skipped below
    for k in CE_KEYS:
        if np.isfinite(ref[k]):
            assert abs(float(got[k]) - ref[k]) < METRIC_TOL * max(1.0, abs(ref[k])), (k, float(got[k]), ref[k])
