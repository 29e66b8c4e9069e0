"""CPU oracle (numpy, float64) for the back-projection + world transform + pose chain.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.

Restates, citing the reference (paths relative to /root/reference):

* ``backproject``            depth_to_pointcloud_dav2.py:300-313 (explicit pinhole formula:
                             x=(u-cx)/fx, y=(v-cy)/fy, points=(x*z, y*z, z), row-major, float64)
                             + the Open3D rule the other script relies on
                             (depth_to_pointcloud.py:218-231, defaults depth_scale=1000,
                             depth_trunc=3.0: z=d/scale; z>=trunc -> 0; emit iff z>0)
                             + depth_to_pointcloud.py:239 ``point_cloud.transform(T)``: X_w = R X + t.
* ``quat_to_matrix``         depth_to_pointcloud.py:168 ``R.from_quat(q).as_matrix()`` (xyzw,
                             normalised by scipy).
* ``make_transform``         depth_to_pointcloud.py:170-173.
* ``parse_intrinsics``       depth_to_pointcloud.py:140-142 (+ comma separated datasets/UnityCam/cam.txt:1).
* ``compose_poses``          eval/evaluation.py:279-382 with quaternion_multiply :385-424 and
                             quaternion_rotate_vector :427-485 (float32, sequential, xyzw, zero-norm
                             relative quaternion -> identity, q NOT normalised).

* ``voxel_down_sample``      depth_to_pointcloud.py:357-359 -> Open3D ``PointCloud::VoxelDownSample`` (Open3D >=0.18,
                             requirements.txt:10; cpp/open3d/geometry/PointCloud.cpp): voxel_min_bound =
                             min_bound - voxel/2; index = floor((p - voxel_min_bound)/voxel) per axis; one
                             output point per occupied voxel = mean (float64 accumulation) of its points and
                             colours.  Open3D emits hash-map order; the oracle emits ascending (ix,iy,iz).
                             Open3D is not installable here: "parity unpinned" by execution, pinned by the
                             hand-computed case in tests/test_oracle_geometry_metrics.py.

PINNING: quat_to_matrix is checked against scipy (installed); compose_poses against the reference's
own ``eval.evaluation.compose_poses`` executed in the build container (fixtures in
tests/golden/, generator scripts/make_golden.py).  Open3D itself is not installed anywhere we
can run, so the Open3D validity rule is restated from its documented defaults: that part is
"parity unpinned" by execution and pinned only by the explicit formula in
depth_to_pointcloud_dav2.py.
"""
from __future__ import annotations

import numpy as np

SIMCOL_K_475 = (156.0418, 155.7529, 178.5604, 181.8043)  # fx, fy, cx, cy  (datasets/UnityCam/cam.txt:1)


def scale_intrinsics(k4, src: int = 475, dst: int = 518):
    s = dst / src
    return tuple(float(v) * s for v in k4)


def parse_intrinsics(text: str):
    vals = np.array([float(t) for t in text.replace(",", " ").split()], dtype=np.float64).reshape(3, 3)
    return float(vals[0, 0]), float(vals[1, 1]), float(vals[0, 2]), float(vals[1, 2])


def quat_to_matrix(q_xyzw) -> np.ndarray:
    q = np.asarray(q_xyzw, dtype=np.float64)
    q = q / np.linalg.norm(q)
    x, y, z, w = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ],
        dtype=np.float64,
    )


def make_transform(position, q_xyzw) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = quat_to_matrix(q_xyzw)
    T[:3, 3] = np.asarray(position, dtype=np.float64)
    return T


def backproject(depth, k4, T=None, depth_scale: float = 1.0, depth_trunc: float = float("inf")):
    """depth [H,W] -> (xyz float64 [H*W,3] dense row-major, valid bool [H*W]).

    Invalid pixels (z<=0 after scale/trunc, or non-finite) carry xyz = 0 in the dense output; the
    Open3D-style compacted cloud is ``xyz[valid]``.
    """
    d = np.asarray(depth)
    H, W = d.shape
    fx, fy, cx, cy = (float(v) for v in k4)
    z = d.astype(np.float64) / float(depth_scale)
    z = np.where(np.isfinite(z), z, 0.0)
    z = np.where(z >= depth_trunc, 0.0, z)
    valid = z > 0
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    x = (u - cx) / fx
    y = (v - cy) / fy
    pts = np.stack((x * z, y * z, z), axis=-1).reshape(-1, 3)
    if T is not None:
        T = np.asarray(T, dtype=np.float64)
        pts = pts @ T[:3, :3].T + T[:3, 3]
    valid = valid.reshape(-1)
    pts = np.where(valid[:, None], pts, 0.0)
    return pts, valid


def _qmul(q1, q2):
    x1, y1, z1, w1 = q1
    x2, y2, z2, w2 = q2
    f = np.float32
    w = f(f(f(w1 * w2) - f(x1 * x2)) - f(y1 * y2)) - f(z1 * z2)
    x = f(f(f(w1 * x2) + f(x1 * w2)) + f(y1 * z2)) - f(z1 * y2)
    y = f(f(f(w1 * y2) - f(x1 * z2)) + f(y1 * w2)) + f(z1 * x2)
    z = f(f(f(w1 * z2) + f(x1 * y2)) - f(y1 * x2)) + f(z1 * w2)
    return np.array([x, y, z, w], dtype=np.float32)


def _cross(a, b):
    f = np.float32
    return np.array(
        [f(a[1] * b[2]) - f(a[2] * b[1]), f(a[2] * b[0]) - f(a[0] * b[2]), f(a[0] * b[1]) - f(a[1] * b[0])],
        dtype=np.float32,
    )


def _qrot(q, v):
    qv, qw = q[:3], q[3]
    uv = _cross(qv, v)
    uuv = _cross(qv, uv)
    return (v + np.float32(2) * (uv * qw + uuv)).astype(np.float32)


def compose_poses(rel, initial=None) -> np.ndarray:
    """rel [N,7] (t | q xyzw) -> abs [N+1,7]; float32, strictly sequential like the reference."""
    rel = np.asarray(rel, dtype=np.float32)
    if rel.ndim == 3:
        rel = rel[0]  # eval/evaluation.py:306-308: only batch 0 is used
    if rel.ndim == 1:
        rel = rel[None]
    cur = np.array([0, 0, 0, 0, 0, 0, 1], dtype=np.float32) if initial is None else np.asarray(initial, np.float32).reshape(-1)
    out = [cur]
    for r in rel:
        rq = r[3:]
        if np.linalg.norm(rq) < 1e-8:
            rq = np.array([0, 0, 0, 1], dtype=np.float32)
        nq = _qmul(cur[3:], rq)
        nt = (cur[:3] + _qrot(cur[3:], r[:3])).astype(np.float32)
        cur = np.concatenate([nt, nq]).astype(np.float32)
        out.append(cur)
    return np.stack(out)


def poses_to_T12(abs7) -> np.ndarray:
    """abs [N,7] -> [N,12] row-major [R|t] rows, float64 (depth_to_pointcloud.py:168-173 semantics)."""
    abs7 = np.asarray(abs7, dtype=np.float64)
    out = np.zeros((abs7.shape[0], 12))
    for i, p in enumerate(abs7):
        T = make_transform(p[:3], p[3:])
        out[i] = T[:3, :4].reshape(-1)
    return out


def voxel_down_sample(points, voxel_size: float, colors=None):
    """points [n,3] (any float dtype; promoted to float64 like Open3D's Vector3d) -> (means [m,3], colour means | None,
    voxel indices [m,3]) sorted by (ix, iy, iz)."""
    if not voxel_size > 0.0:
        raise ValueError("voxel_size <= 0.")
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    if p.shape[0] == 0:
        return p, (None if colors is None else np.zeros((0, 3))), np.zeros((0, 3), dtype=np.int64)
    vmin = p.min(axis=0) - 0.5 * voxel_size
    vmax = p.max(axis=0) + 0.5 * voxel_size
    if voxel_size * np.iinfo(np.int32).max < (vmax - vmin).max():
        raise RuntimeError("voxel_size is too small.")
    idx = np.floor((p - vmin) / voxel_size).astype(np.int64)
    uniq, inv = np.unique(idx, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    cnt = np.bincount(inv, minlength=uniq.shape[0]).astype(np.float64)
    acc = np.zeros((uniq.shape[0], 3))
    np.add.at(acc, inv, p)  # sequential in index order, like the reference's loop
    out_c = None
    if colors is not None:
        c = np.asarray(colors, dtype=np.float64).reshape(-1, 3)
        accc = np.zeros((uniq.shape[0], 3))
        np.add.at(accc, inv, c)
        out_c = accc / cnt[:, None]
    return acc / cnt[:, None], out_c, uniq
