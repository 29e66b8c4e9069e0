"""Generate tests/golden/*.npz by EXECUTING the reference's own importable functions.

Run in the build container only (needs /root/reference):  python scripts/make_golden.py
The fixtures travel to the GPU box; /root/reference does not.

Reference functions executed (unmodified, imported from /root/reference):
  calculate_metrics.calculate_metrics            calculate_metrics.py:17-55
  eval.evaluation.compute_errors                 eval/evaluation.py:16-60
  eval.evaluation.compose_poses (+quaternion_*)  eval/evaluation.py:279-485
  scipy Rotation.from_quat(..).as_matrix()       as used by depth_to_pointcloud.py:168
and the explicit back-projection block of depth_to_pointcloud_dav2.py:300-313 (inline code in
``main``; not importable as a function, so its six numpy lines are evaluated here on the fixture).
``config1_fixture`` adds BASELINE configs[0]: the left 475x475 crop of the reference's FrameBuffer_0051.png (copied as a
PNG fixture) through the oracle's infer_image and the same explicit block.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth_depth_pair(rng, H, W, invalid_frac=0.02):
    gt = np.clip(rng.gamma(2.0, 0.15, size=(H, W)), 0, 1).astype(np.float32)
    gt[rng.random((H, W)) < invalid_frac] = 0.0
    pred = (np.where(gt > 0, gt, 0.3) * rng.normal(1.0, 0.07, size=(H, W))).astype(np.float32)
    return gt, pred


def main():
    sys.path.insert(0, REF)
    import calculate_metrics as ref_cm  # noqa: E402
    from eval import evaluation as ref_ev  # noqa: E402
    from scipy.spatial.transform import Rotation as R  # noqa: E402

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- metrics: per-frame calculate_metrics and per-batch compute_errors -----------------
    B, H, W = 3, 70, 98
    gts, preds, cm_rows, = [], [], []
    for _ in range(B):
        gt, pred = synth_depth_pair(rng, H, W)
        pred[rng.random((H, W)) < 0.01] = 0.0  # exercises the pred>0 guard / division by zero
        gts.append(gt)
        preds.append(pred)
        m = ref_cm.calculate_metrics(gt.copy(), pred.copy())
        cm_rows.append([m[k] for k in ("rmse", "mae", "abs_rel", "sq_rel", "delta1", "delta2", "delta3")])
    gt_b, pred_b = np.stack(gts), np.stack(preds)
    # lightning_model.py:304-313 : batch-wide mask then compute_errors
    tg, tp = torch.from_numpy(gt_b)[:, None], torch.from_numpy(pred_b)[:, None]
    mask = (tg >= 1e-6) & (tg <= 20.0)
    ce = ref_ev.compute_errors(tp[mask].flatten(), tg[mask].flatten())
    # SURVEY 8c spot case (seed 0, 518^2)
    r0 = np.random.default_rng(0)
    gt0 = np.clip(r0.gamma(2.0, 0.03, size=(518, 518)), 0, 0.2).astype(np.float32)
    pred0 = (gt0 * r0.normal(1.0, 0.05, size=(518, 518))).astype(np.float32)
    m0 = ref_cm.calculate_metrics(gt0.copy(), pred0.copy())
    t0g, t0p = torch.from_numpy(gt0), torch.from_numpy(pred0)
    mk = (t0g >= 1e-6) & (t0g <= 20.0)
    c0 = ref_ev.compute_errors(t0p[mk].flatten(), t0g[mk].flatten())
    empty = ref_cm.calculate_metrics(np.zeros((4, 4), np.float32), np.ones((4, 4), np.float32))
    np.savez_compressed(
        os.path.join(OUT, "metrics_small.npz"),
        gt=gt_b, pred=pred_b,
        calculate_metrics=np.asarray(cm_rows, dtype=np.float64),
        compute_errors=np.asarray([float(ce[k]) for k in ("d1", "abs_rel", "rmse", "l1")], dtype=np.float64),
        spot518_calculate_metrics=np.asarray([m0[k] for k in ("rmse", "mae", "abs_rel", "sq_rel", "delta1", "delta2", "delta3")], dtype=np.float64),
        spot518_compute_errors=np.asarray([float(c0[k]) for k in ("d1", "abs_rel", "rmse", "l1")], dtype=np.float64),
        empty_is_nan=np.asarray([np.isnan(v) for v in empty.values()]),
    )

    # ---- pose chain ---------------------------------------------------------------------
    N = 64
    t = rng.normal(0, 0.01, size=(N, 3)).astype(np.float32)
    rv = np.deg2rad(rng.normal(0, 2.0, size=(N, 3)))
    q = R.from_rotvec(rv).as_quat().astype(np.float32)  # xyzw
    rel = np.concatenate([t, q], axis=1).astype(np.float32)
    rel[17, 3:] = 0.0  # zero quaternion -> identity branch (eval/evaluation.py:331-338)
    rel[30, 3:] *= 1.3  # non-unit quaternion: the reference does not normalise
    absp = ref_ev.compose_poses(torch.from_numpy(rel)).numpy()
    init = np.array([0.1, -0.2, 0.3, *R.from_rotvec([0.1, 0.2, -0.3]).as_quat()], dtype=np.float32)
    absp_init = ref_ev.compose_poses(torch.from_numpy(rel), torch.from_numpy(init)).numpy()
    mats = np.stack([R.from_quat(p[3:].astype(np.float64)).as_matrix() for p in absp_init])
    np.savez_compressed(os.path.join(OUT, "poses_small.npz"), rel=rel, abs=absp, init=init, abs_init=absp_init, rot=mats)

    # ---- explicit back-projection block (depth_to_pointcloud_dav2.py:300-313) -------------
    H, W = 37, 53
    z = np.clip(rng.gamma(2.0, 0.03, size=(H, W)), 0, 0.2).astype(np.float32)
    fx, fy, cx, cy = 156.0418 * W / 475, 155.7529 * H / 475, 178.5604 * W / 475, 181.8043 * H / 475
    x, y = np.meshgrid(np.arange(W), np.arange(H))
    x = (x - cx) / fx
    y = (y - cy) / fy
    points = np.stack((np.multiply(x, z), np.multiply(y, z), z), axis=-1).reshape(-1, 3)
    T = np.eye(4)
    T[:3, :3] = R.from_quat(absp_init[5, 3:].astype(np.float64)).as_matrix()  # depth_to_pointcloud.py:168-173
    T[:3, 3] = absp_init[5, :3]
    world = points @ T[:3, :3].T + T[:3, 3]
    np.savez_compressed(os.path.join(OUT, "backproject_small.npz"), depth=z, k4=np.array([fx, fy, cx, cy]), points=points, T=T, world=world)
    print("wrote", sorted(os.listdir(OUT)))


def pose_metric_fixtures():
    """tests/golden/pose_metrics_small.npz: eval/evaluation.py:63-254 executed on a seeded synthetic trajectory."""
    import sys
    import numpy as np
    import torch
    from scipy.spatial.transform import Rotation as R
    sys.path.insert(0, REF)
    from eval import evaluation as ref
    rng = np.random.default_rng(77)
    N = 40

    def rel(n, s):
        t = rng.normal(0, s, size=(n, 3))
        q = R.from_rotvec(np.deg2rad(rng.normal(0, 2.0, size=(n, 3)))).as_quat()
        return np.concatenate([t, q], 1).astype(np.float32)

    gt = rel(N, 0.01)
    pred = gt.copy()
    pred[:, :3] = pred[:, :3] / np.linalg.norm(pred[:, :3], axis=1, keepdims=True) + rng.normal(0, 0.05, size=(N, 3)).astype(np.float32)
    pred[:, 3:] += rng.normal(0, 0.01, size=(N, 4)).astype(np.float32)
    pred[5, 3:] = 0
    traj = ref.evaluate_trajectory(torch.from_numpy(pred), torch.from_numpy(gt))
    absg, absp = ref.compose_poses(torch.from_numpy(gt)), ref.compose_poses(torch.from_numpy(pred))
    pe = ref.compute_pose_errors(absp, absg)
    np.savez_compressed(os.path.join(OUT, "pose_metrics_small.npz"), gt=gt, pred=pred,
                        traj=np.array([float(traj["rte"]), float(traj["ate"]), float(traj["rote"])]),
                        pose_errors=np.array([float(pe["ate"]), float(pe["rte"]), float(pe["rote"])]),
                        qdist=np.array([ref.quaternion_distance(gt[0, 3:], pred[0, 3:])]))


def config1_fixture():
    """BASELINE configs[0]: left 475x475 crop of the reference's own FrameBuffer_0051.png ->
    cv2.INTER_CUBIC to 518x518 -> infer_image(img, 518) (vits, seed-0 oracle weights) -> the explicit back-projection
    block of depth_to_pointcloud_dav2.py:300-313 evaluated with K scaled to 518 (datasets/UnityCam/cam.txt:1).
    Writes the crop (PNG) and a stride-7 sample of the oracle depth + the reference formula's points."""
    import cv2
    sys.path.insert(0, os.path.dirname(OUT.rstrip("/")).rsplit("/tests", 1)[0])
    from oracle import dav2_oracle as O
    img = cv2.imread(os.path.join(REF, "FrameBuffer_0051.png"))
    crop = np.ascontiguousarray(img[:, :475])
    assert crop.shape == (475, 475, 3)
    cv2.imwrite(os.path.join(OUT, "FrameBuffer_0051_left475.png"), crop)
    img518 = cv2.resize(crop, (518, 518), interpolation=cv2.INTER_CUBIC)
    torch.set_num_threads(8)
    oracle = O.build_oracle("vits", seed=0)
    depth = oracle.infer_image(img518, 518)
    assert depth.shape == (518, 518) and depth.dtype == np.float32
    cam = [float(v) for v in open(os.path.join(REF, "datasets/UnityCam/cam.txt")).readline().split(",")]
    s = 518.0 / 475.0
    fx, fy, cx, cy = cam[0] * s, cam[4] * s, cam[2] * s, cam[5] * s
    width = height = 518
    # --- depth_to_pointcloud_dav2.py:300-313, verbatim arithmetic -------------------------------------------------
    x, y = np.meshgrid(np.arange(width), np.arange(height))
    x = (x - cx) / fx
    y = (y - cy) / fy
    z = np.array(depth)
    points = np.stack((np.multiply(x, z), np.multiply(y, z), z), axis=-1).reshape(-1, 3)
    # ---------------------------------------------------------------------------------------------------------------
    sub = np.zeros((518, 518), bool)
    sub[::7, ::7] = True
    np.savez_compressed(os.path.join(OUT, "config1_vits.npz"), k4=np.array([fx, fy, cx, cy]), stride=np.array(7),
                        depth_sub=depth[::7, ::7], points_sub=points[sub.reshape(-1)],
                        depth_stats=np.array([depth.min(), depth.max(), depth.mean(), depth.std()], dtype=np.float64))
    print("config1:", depth.min(), depth.max(), depth.mean(), depth.std())


def pose_module_fixture():
    """tests/golden/pose_module_small.npz: what PoseEstimationModule's test hooks compute (pose_estimation_model.py:302-343)
    for 5 batches of 8 pairs -- per-batch compute_pose_errors and evaluate_trajectory on the STACKED [5,8,7] tensors, both
    by executing the reference's eval/evaluation.py."""
    import numpy as np
    import torch
    from scipy.spatial.transform import Rotation as R
    sys.path.insert(0, REF)
    from eval import evaluation as ref
    rng = np.random.default_rng(91)
    nb, B = 5, 8
    t = rng.normal(0, 1.0, size=(nb, B, 3))
    t /= np.linalg.norm(t, axis=2, keepdims=True)
    q = R.from_rotvec(np.deg2rad(rng.normal(0, 2.0, size=(nb * B, 3)))).as_quat().reshape(nb, B, 4)
    gt = np.concatenate([t, q], 2).astype(np.float32)
    pred = gt + rng.normal(0, 0.03, size=gt.shape).astype(np.float32)
    per_batch = []
    for b in range(nb):
        m = ref.compute_pose_errors(torch.from_numpy(pred[b]), torch.from_numpy(gt[b]))
        per_batch.append([float(m["ate"]), float(m["rte"]), float(m["rote"])])
    traj = ref.evaluate_trajectory(pred_rel_poses=torch.from_numpy(pred.copy()), gt_rel_poses=torch.from_numpy(gt.copy()),
                                   initial_pose=None)
    np.savez_compressed(os.path.join(OUT, "pose_module_small.npz"), gt=gt, pred=pred, per_batch=np.array(per_batch),
                        traj=np.array([float(traj["ate"]), float(traj["rte"]), float(traj["rote"])]))
    print("pose module:", per_batch[0], traj)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "config1":
        config1_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "pose_module":
        pose_module_fixture()
    else:
        main()
        pose_metric_fixtures()
        pose_module_fixture()
        config1_fixture()
