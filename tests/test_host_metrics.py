"""Host-side pieces of the path: pose metrics vs fixtures produced by executing the reference
(eval/evaluation.py:63-208), the per-procedure collector (test_lightning.py:27-111,240-274), run.py writers."""
import os

import numpy as np
import pytest
import torch


def test_pose_errors_match_reference(golden_dir):
    from dav2_b200 import evaluation as ev
    from oracle import geometry_oracle as geo
    g = np.load(os.path.join(golden_dir, "pose_metrics_small.npz"))
    absg, absp = geo.compose_poses(g["gt"]), geo.compose_poses(g["pred"])
    pe = ev.compute_pose_errors(torch.from_numpy(absp), torch.from_numpy(absg))
    np.testing.assert_allclose([float(pe["ate"]), float(pe["rte"]), float(pe["rote"])], g["pose_errors"], rtol=1e-4, atol=1e-6)
    # the reference evaluates arccos near 1 in float32 (inputs are float32); ours in float64: agree to ~3e-4 deg
    assert abs(ev.quaternion_distance(g["gt"][0, 3:], g["pred"][0, 3:]) - g["qdist"][0]) < 2e-3
    s = ev.calculate_scale_factor(torch.from_numpy(g["pred"]), torch.from_numpy(g["gt"]))
    assert torch.isfinite(s)


@pytest.mark.gpu
def test_evaluate_trajectory_matches_reference(golden_dir):
    from dav2_b200 import evaluation as ev
    g = np.load(os.path.join(golden_dir, "pose_metrics_small.npz"))
    out = ev.evaluate_trajectory(torch.from_numpy(g["pred"]).cuda(), torch.from_numpy(g["gt"]).cuda())
    np.testing.assert_allclose([float(out["rte"]), float(out["ate"]), float(out["rote"])], g["traj"], rtol=1e-3, atol=1e-5)


def test_procedure_collector_semantics():
    from dav2_b200.evaluation import ProcedureMetricCollector
    c = ProcedureMetricCollector()
    assert c.procedure_of("datasets/SyntheticColon/SyntheticColon_I", "S3_0012") == "SyntheticColon_I/Frames_S3"
    assert c.procedure_of("datasets/SyntheticColon/SyntheticColon_II", "B10_0001") == "SyntheticColon_II/Frames_B10"
    assert c.procedure_of("datasets/other", "S3_0012") is None
    b1 = {"dataset": ["x/SyntheticColon_I"] * 3, "id": ["S1_0000", "S1_0001", "S2_0000"]}
    b2 = {"dataset": ["x/SyntheticColon_I"] * 2, "id": ["S2_0001", "S2_0002"]}
    c.on_test_batch_end({"l1": 1.0, "abs_rel": 2.0, "d1": 0.5, "rmse": 3.0}, b1)
    c.on_test_batch_end({"l1": 3.0, "abs_rel": 4.0, "d1": 0.7, "rmse": 5.0}, b2)
    s = c.summary()
    # the BATCH value is replicated per frame; overall = mean over procedures of per-procedure means
    assert s["per_procedure"]["SyntheticColon_I/Frames_S1"]["l1"] == 1.0
    assert abs(s["per_procedure"]["SyntheticColon_I/Frames_S2"]["l1"] - (1.0 + 3.0 + 3.0) / 3) < 1e-12
    assert abs(s["overall_metrics"]["l1"]["mean"] - (1.0 + 7.0 / 3) / 2) < 1e-12
    with pytest.raises(ValueError):
        c.on_test_batch_end({"l1": 1.0}, b1)


def test_run_writers():
    from dav2_b200 import run
    d = np.linspace(0.5, 4.5, 12, dtype=np.float32).reshape(3, 4)
    u8 = run.depth_to_uint8(d)
    assert u8.dtype == np.uint8 and u8.min() == 0 and u8.max() == 255
    assert run.colorize(u8, grayscale=True).shape == (3, 4, 3)
    # run.py:160,245-248: matplotlib "Spectral" (not reversed), integer image indexes the 256-entry table; BGR out.
    # Known answers of matplotlib.colormaps["Spectral"]: entry 0 / 255 are the first / last ColorBrewer anchors,
    # entry 128 = (0.99807766, 0.99923106, 0.74602076).
    lut = run.spectral_lut()
    assert lut.shape == (256, 3)
    np.testing.assert_allclose(lut[0], np.array([158, 1, 66]) / 255.0, atol=1e-12)
    np.testing.assert_allclose(lut[255], np.array([94, 79, 162]) / 255.0, atol=1e-12)
    np.testing.assert_allclose(lut[128], [0.99807766, 0.99923106, 0.74602076], atol=1e-8)
    col = run.colorize(np.array([[0, 255, 128]], dtype=np.uint8), grayscale=False)
    assert col.dtype == np.uint8 and col[0, 0].tolist() == [66, 1, 158] and col[0, 1].tolist() == [162, 79, 94]
    assert col[0, 2].tolist() == [190, 254, 254]
    assert run.output_path("/a/b/FrameBuffer_0051.png", "/out") == "/out/FrameBuffer_0051.png"


@pytest.mark.gpu
def test_run_frames_loop(tmp_path):
    """run.py:195-262: writes <stem>.png (+ .npy), skips files whose PNG exists, batches equal-shape frames."""
    import cv2
    from dav2_b200 import run, weights
    from dav2_b200.dpt import MODEL_CONFIGS, DepthAnythingV2
    m = DepthAnythingV2(**MODEL_CONFIGS["vits"])
    weights.randomize_(m, 1)
    m = weights.calibrate_(m.cuda().eval())
    rng = np.random.default_rng(0)
    files = []
    for i, (h, w) in enumerate([(60, 80), (60, 80), (50, 90)]):
        p = str(tmp_path / f"frame_{i}.jpg")
        cv2.imwrite(p, rng.integers(0, 255, size=(h, w, 3), dtype=np.uint8))
        files.append(p)
    out = str(tmp_path / "out")
    written = run.run_frames(m, files, out, input_size=70, save_numpy=True, pred_only=False, batch=4)
    assert len(written) == 3
    d0 = np.load(os.path.join(out, "frame_0.npy"))
    assert d0.shape == (60, 80) and d0.dtype == np.float32
    single = m.infer_image(cv2.imread(files[0]), 70)
    assert np.abs(single - d0).max() / np.abs(single).max() < 1e-2
    img = cv2.imread(os.path.join(out, "frame_2.png"))
    assert img.shape == (50, 90 + 50 + 90, 3)
    assert run.run_frames(m, files, out, input_size=70) == []  # everything exists -> skipped


def test_input_output_files(tmp_path):
    """depth_to_pointcloud.py:53-122 on a SimCol-shaped tree, a 'testing' folder, list files and a single image."""
    from types import SimpleNamespace as NS
    from dav2_b200 import depth_to_pointcloud as d2p
    base = tmp_path / "SyntheticColon"
    for sub, frames in (("I", "Frames_S1"), ("II", "Frames_B2")):
        (base / f"SyntheticColon_{sub}" / frames).mkdir(parents=True)
        (base / f"SyntheticColon_{sub}" / f"{frames}_OP" / "depth").mkdir(parents=True)
        for i in (1, 0):
            (base / f"SyntheticColon_{sub}" / frames / f"FrameBuffer_{i:04d}.png").write_bytes(b"x")
            (base / f"SyntheticColon_{sub}" / f"{frames}_OP" / "depth" / f"Depth_{i:04d}.png").write_bytes(b"x")
        (base / f"SyntheticColon_{sub}" / f"{frames}_OP" / f"FrameBuffer_9999.png").write_bytes(b"x")  # "_OP" is filtered
    a = NS(img_path=str(base), depth_path="", ds_type="simcol", outdir=None)
    rgb, dep, out = d2p.input_output_files(a)
    assert [os.path.basename(p) for p in rgb] == ["FrameBuffer_0000.png", "FrameBuffer_0001.png"] * 2
    assert all("_OP" not in p for p in rgb) and "SyntheticColon_I/" in rgb[0] and "SyntheticColon_II/" in rgb[2]
    assert [os.path.basename(p) for p in dep] == ["Depth_0000.png", "Depth_0001.png"] * 2
    assert out == str(base) and a.outdir == str(base)
    t = tmp_path / "testing"
    t.mkdir()
    for n in ("frame_2.jpg", "frame_1.jpg", "other.jpg"):
        (t / n).write_bytes(b"x")
    rgb, dep, out = d2p.input_output_files(NS(img_path=str(t), depth_path="", ds_type="testing", outdir="o"))
    assert [os.path.basename(p) for p in rgb] == ["frame_1.jpg", "frame_2.jpg"] and dep == [] and out == "o"
    (tmp_path / "rgb.txt").write_text("a.png\nb.png\n")
    (tmp_path / "dep.txt").write_text("da.png\ndb.png\n")
    rgb, dep, out = d2p.input_output_files(NS(img_path=str(tmp_path / "rgb.txt"), depth_path=str(tmp_path / "dep.txt"),
                                              ds_type="simcol", outdir=None))
    assert rgb == ["a.png", "b.png"] and dep == ["da.png", "db.png"] and out is None
    single = str(t / "frame_1.jpg")
    rgb, dep, out = d2p.input_output_files(NS(img_path=single, depth_path="d.png", ds_type="simcol", outdir=None))
    assert rgb == [single] and dep == [] and out == str(t)


def test_lightning_module_surface():
    """lightning_model.DepthAnythingV2Module mirror: constructor / hparams / checkpoint key handling run without a
    GPU (the engine is created on first forward)."""
    import torch
    from dav2_b200 import lightning_model as lm
    mod = lm.DepthAnythingV2Module(encoder="vits", min_depth=1e-6, max_depth=20.0, encoder_lr=5e-6)
    assert mod.hparams.encoder == "vits" and mod.hparams.max_depth == 20.0 and mod.hparams.encoder_lr == 5e-6
    assert mod.model.max_depth == 20.0 and mod.device.type == "cpu"
    sd = mod.state_dict()
    assert all(k.startswith("model.") for k in sd) and any("pretrained" in k for k in sd)
    # a Lightning checkpoint's state_dict loads back (test_lightning.py:114-130); stray keys are rejected when strict
    key = next(k for k in sd if k.endswith("cls_token"))
    sd2 = {k: (torch.full_like(v, 0.25) if k == key else v) for k, v in sd.items()}
    mod.load_state_dict(sd2)
    assert float(mod.model.state_dict()[key[len("model."):]].flatten()[0]) == 0.25
    with pytest.raises(RuntimeError):
        mod.load_state_dict({**sd2, "loss.weight": torch.zeros(1)})
    with pytest.raises(NotImplementedError):
        mod.training_step({}, 0)
    with pytest.raises(ValueError):
        lm.DepthAnythingV2Module(encoder="vitg")
    with pytest.raises(FileNotFoundError):
        lm.DepthAnythingV2Module(encoder="vits", pretrained_from="/nonexistent/ckpt.pth")
    from dav2_b200.evaluation import RunningMeans
    means = RunningMeans(lm.METRIC_KEYS)
    assert all(np.isnan(v) for v in means.compute().values())
    means.update({"d1": 0.5, "abs_rel": 1.0, "rmse": 2.0, "l1": 3.0})
    means.update({"d1": torch.tensor(1.0), "abs_rel": torch.tensor(3.0), "rmse": torch.tensor(4.0), "l1": torch.tensor(5.0)})
    assert means.compute() == {"d1": 0.75, "abs_rel": 2.0, "rmse": 3.0, "l1": 4.0}


def test_pointcloud_dav2_host_side(tmp_path):
    """depth_to_pointcloud_dav2.py host helpers: camera file, input listing, camera selection, pose file, checkpoint keys."""
    import torch
    from scipy.spatial.transform import Rotation
    from dav2_b200 import depth_to_pointcloud_dav2 as pc
    cam = tmp_path / "cam.txt"
    cam.write_text("227.60416 0 227.5 0 237.5 237.5 0 0 1\nignored second line\n")
    assert pc.read_cam_file(str(cam)) == {"fx": 227.60416, "fy": 237.5, "cx": 227.5, "cy": 237.5}
    # listing (:189-240): txt list, single file, simcol tree without the _OP folders, testing jpgs
    lst = tmp_path / "list.txt"
    lst.write_text("a.png\nb.png")
    assert pc.collect_filenames(str(lst)) == (["a.png", "b.png"], None)
    one = tmp_path / "one.png"
    one.write_bytes(b"")
    assert pc.collect_filenames(str(one)) == ([str(one)], str(tmp_path))
    assert pc.collect_filenames(str(one), outdir="o") == ([str(one)], "o")
    root = tmp_path / "SyntheticColon"
    for d in ("SyntheticColon_I/Frames_S1", "SyntheticColon_I/Frames_S1_OP", "SyntheticColon_III/Frames_O2"):
        (root / d).mkdir(parents=True)
        (root / d / "FrameBuffer_0000.png").write_bytes(b"")
        (root / d / "Depth_0000.png").write_bytes(b"")
    files, outdir = pc.collect_filenames(str(root), "simcol")
    assert outdir == str(root) and sorted(os.path.relpath(f, root) for f in files) == [
        "SyntheticColon_I/Frames_S1/FrameBuffer_0000.png", "SyntheticColon_III/Frames_O2/FrameBuffer_0000.png"]
    (tmp_path / "frame_01.jpg").write_bytes(b"")
    assert pc.collect_filenames(str(tmp_path), "testing")[0] == [str(tmp_path / "frame_01.jpg")]
    # camera selection (:253-268), including the reference's substring behaviour
    assert pc.cam_file_for("x/SyntheticColon_I/Frames_S1/f.png", "simcol", None) == "datasets/SyntheticColon/SyntheticColon_I/cam.txt"
    assert pc.cam_file_for("x/SyntheticColon_III/Frames_O2/f.png", "simcol", None) == "datasets/SyntheticColon/SyntheticColon_I/cam.txt"
    assert pc.cam_file_for("f.png", None, "my_cam.txt") == "my_cam.txt"
    with pytest.raises(ValueError):
        pc.cam_file_for("x/other/f.png", "simcol", None)
    with pytest.raises(ValueError):
        pc.cam_file_for("f.png", None, None)
    # pose (:53-69) against scipy, un-normalised quaternion in the file
    (tmp_path / "p.txt").write_text("0.1 -0.2 0.3")
    (tmp_path / "q.txt").write_text("0.2 0.4 -0.2 1.6")
    T = pc.load_transformation(str(tmp_path / "p.txt"), str(tmp_path / "q.txt"))
    np.testing.assert_allclose(T[:3, :3], Rotation.from_quat([0.2, 0.4, -0.2, 1.6]).as_matrix(), atol=1e-15)
    np.testing.assert_allclose(T[:, 3], [0.1, -0.2, 0.3, 1.0])
    # checkpoint (:163-185)
    lin = torch.nn.Linear(2, 2)
    w = torch.full((2, 2), 0.5)
    torch.save({"state_dict": {"model.weight": w, "model.bias": torch.zeros(2)}}, tmp_path / "a.ckpt")
    pc.load_checkpoint(lin, str(tmp_path / "a.ckpt"))
    assert torch.equal(lin.weight.data, w)
    torch.save({"weight": 2 * w, "bias": torch.zeros(2)}, tmp_path / "b.pth")
    pc.load_checkpoint(lin, str(tmp_path / "b.pth"))
    assert torch.equal(lin.weight.data, 2 * w)


def test_relative_pose_targets_match_item_code():
    """data_processing.relative_pose_targets (all pairs at once) == the reference loader's per-item code
    (pose_estimation.py:245-303, restated in oracle/pose_oracle.py), including a repeated pose (zero motion)."""
    from oracle import pose_oracle as PO
    from dav2_b200 import data_processing as dp
    rng = np.random.default_rng(3)
    N = 40
    pos = np.cumsum(rng.normal(0, 0.01, (N, 3)), axis=0)
    q = rng.normal(0, 1, (N, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    poses = np.concatenate([pos, q], 1)
    poses[7] = poses[6]  # identical consecutive poses: 0 / 1e-8 translation, identity rotation
    got = dp.relative_pose_targets(poses)
    assert got.shape == (N - 1, 7) and got.dtype == torch.float32
    want = torch.stack([PO.relative_pose_item(poses[i], poses[i + 1]) for i in range(N - 1)])
    assert torch.allclose(got, want, rtol=0, atol=2e-7), float((got - want).abs().max())
    assert torch.equal(got[6, :3], torch.zeros(3)) and abs(float(got[6, 6])) > 0.999999
    assert torch.allclose(got[:, 3:].norm(dim=1), torch.ones(N - 1), atol=1e-6)
    with pytest.raises(ValueError):
        dp.relative_pose_targets(np.zeros((4, 6)))


def _pose_module_with_fixture_preds(g):
    """PoseEstimationModule whose network is replaced by the fixture's predictions: the hooks' bookkeeping is under test."""
    from dav2_b200.pose_estimation_model import PoseEstimationModule
    mod = PoseEstimationModule(in_channels=8, lr=1e-4)
    preds = iter(torch.from_numpy(g["pred"]))
    mod.model = lambda x: next(preds).to(x.device)
    return mod


def test_pose_module_test_step_matches_reference(golden_dir):
    """PoseEstimationModule.test_step (pose_estimation_model.py:302-317): per-batch ate / rte / rote equal the executed
    reference's; state-dict prefix handling; pose collector keys."""
    from dav2_b200.evaluation import POSE_KEYS, ProcedureMetricCollector
    from dav2_b200.pose_estimation_model import PoseEstimationModule
    g = np.load(os.path.join(golden_dir, "pose_module_small.npz"))
    real = PoseEstimationModule(in_channels=8, lr=1e-4)
    assert real.hparams.lr == 1e-4 and all(k.startswith("model.") for k in real.state_dict())
    real.load_state_dict(real.state_dict())
    with pytest.raises(RuntimeError):
        real.load_state_dict({**real.state_dict(), "criterion.w": torch.zeros(1)})
    with pytest.raises(NotImplementedError):
        real.training_step({})
    mod = _pose_module_with_fixture_preds(g)
    coll = ProcedureMetricCollector(POSE_KEYS)
    mod.on_test_epoch_start()
    for b in range(g["gt"].shape[0]):
        batch = {"input": torch.zeros(8, 8, 4, 4), "target": torch.from_numpy(g["gt"][b]),
                 "dataset": ["d/SyntheticColon_II"] * 8, "id": [f"B{b + 1}_{j:04d}" for j in range(8)]}
        out = mod.test_step(batch, b)
        np.testing.assert_allclose([float(out[k]) for k in POSE_KEYS], g["per_batch"][b], rtol=1e-4, atol=1e-6)
        coll.on_test_batch_end(out, batch)
    assert len(mod.current_trajectory_preds) == 5 and mod.current_trajectory_preds[0].shape == (8, 7)
    s = coll.summary()
    assert abs(s["per_procedure"]["SyntheticColon_II/Frames_B2"]["rote"] - g["per_batch"][1][2]) < 1e-4
    np.testing.assert_allclose(mod.metric.compute()["ate"], g["per_batch"][:, 0].mean(), rtol=1e-4)


@pytest.mark.gpu
def test_pose_module_epoch_end_matches_reference(golden_dir):
    """on_test_epoch_end (:319-343): evaluate_trajectory on the stacked per-batch tensors, as the reference executes it."""
    g = np.load(os.path.join(golden_dir, "pose_module_small.npz"))
    mod = _pose_module_with_fixture_preds(g)
    mod.on_test_epoch_start()
    for b in range(g["gt"].shape[0]):
        mod.test_step({"input": torch.zeros(8, 8, 4, 4), "target": torch.from_numpy(g["gt"][b])}, b)
    out = mod.on_test_epoch_end()
    np.testing.assert_allclose([float(out["trajectory"][k]) for k in ("ate", "rte", "rote")], g["traj"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose([out["mean"][k] for k in ("ate", "rte", "rote")], g["per_batch"].mean(0), rtol=1e-4)
    assert set(mod.logged) == {f"Test/{p}_{k}" for p in ("test", "trajectory") for k in ("ate", "rte", "rote")}
    assert mod.current_trajectory_preds == []
