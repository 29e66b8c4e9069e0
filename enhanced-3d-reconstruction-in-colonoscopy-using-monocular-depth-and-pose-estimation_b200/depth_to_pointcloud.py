"""Drop-in for the hot-path functions of the reference's ``depth_to_pointcloud.py``.

The reference builds each frame's cloud with Open3D on the CPU (depth_to_pointcloud.py:214-239):
RGBD image (z = d/1000, z >= 3 -> 0), pinhole back-projection of z > 0 pixels, 4x4 world
transform from the ground-truth pose.  Here ONE fused kernel (dav2_backproject) does
back-projection + SE(3) + validity on the GPU; compaction / colour gather are torch indexing on
the device.  ``PointCloud`` mimics the small part of ``o3d.geometry.PointCloud`` the script uses
(``.points``, ``.colors``, ``.transform``, ``+=``, ``.voxel_down_sample``).  Poisson meshing (:245-281, :361-362) is out of scope
(a global sparse solve inside Open3D).
"""
from __future__ import annotations

import glob
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch

from . import ops


@dataclass
class PinholeCameraIntrinsic:
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float

    @property
    def intrinsic_matrix(self):
        return np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1.0]])

    def k4(self):
        return (self.fx, self.fy, self.cx, self.cy)


class PointCloud:
    """Device-resident cloud: points fp32 [n,3], colours fp32 [n,3] in [0,1]."""

    def __init__(self, points: torch.Tensor | None = None, colors: torch.Tensor | None = None):
        self._p = [points] if points is not None else []
        self._c = [colors] if colors is not None else []

    def _cat(self):
        if len(self._p) > 1:
            self._p = [torch.cat(self._p)]
            self._c = [torch.cat(self._c)] if self._c else []
        return (self._p[0] if self._p else None), (self._c[0] if self._c else None)

    @property
    def points_tensor(self):
        return self._cat()[0]

    @property
    def points(self) -> np.ndarray:
        p = self._cat()[0]
        return np.zeros((0, 3)) if p is None else p.double().cpu().numpy()

    @property
    def colors(self) -> np.ndarray:
        c = self._cat()[1]
        return np.zeros((0, 3)) if c is None else c.double().cpu().numpy()

    def __len__(self):
        return sum(int(t.shape[0]) for t in self._p)

    def __iadd__(self, other: "PointCloud"):
        # depth_to_pointcloud.py:354 ``combined += point_cloud``: O(1) append, one concat at read time
        self._p += other._p
        self._c += other._c
        return self

    def voxel_down_sample(self, voxel_size: float) -> "PointCloud":
        """o3d PointCloud.voxel_down_sample (depth_to_pointcloud.py:357-359) on the GPU (dav2_voxel_downsample)."""
        p, c = self._cat()
        if p is None or p.shape[0] == 0:
            return PointCloud()
        if c is not None and c.shape[0] != p.shape[0]:
            c = None
        q, qc = ops.voxel_downsample(p, voxel_size, c)
        return PointCloud(q, qc)

    def transform(self, T):
        p, _ = self._cat()
        if p is not None and p.shape[0]:
            self._p = [ops.transform_points_(p.contiguous(), np.asarray(T, dtype=np.float64)[:3, :4])]  # fp64 math, one rounding
        return self


def load_camera_intrinsics(file_path: str, width: int, height: int) -> PinholeCameraIntrinsic:
    """depth_to_pointcloud.py:126-151 (accepts whitespace- and comma-separated 3x3, cf. datasets/UnityCam/cam.txt)."""
    with open(file_path, "r", encoding="utf-8") as f:
        vals = np.array([float(t) for t in f.read().replace(",", " ").split()], dtype=np.float64).reshape(3, 3)
    return PinholeCameraIntrinsic(width, height, vals[0, 0], vals[1, 1], vals[0, 2], vals[1, 2])


def quat_to_matrix(q_xyzw) -> np.ndarray:
    x, y, z, w = np.asarray(q_xyzw, dtype=np.float64) / np.linalg.norm(q_xyzw)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


_pose_cache: dict = {}


def _load_pose_table(path: str) -> np.ndarray:
    # the reference re-parses both files for EVERY frame (depth_to_pointcloud.py:160-161); parse once
    key = (path, os.path.getmtime(path))
    if key not in _pose_cache:
        _pose_cache[key] = np.loadtxt(path, delimiter="," if "," in open(path).readline() else None)
    return _pose_cache[key]


def load_transformation(position_file, rotation_file, frame_idx) -> np.ndarray:
    """depth_to_pointcloud.py:154-174 -> 4x4 float64."""
    position = _load_pose_table(str(position_file))[frame_idx]
    quaternion = _load_pose_table(str(rotation_file))[frame_idx]
    T = np.eye(4)
    T[:3, :3] = quat_to_matrix(quaternion)
    T[:3, 3] = position
    return T


def point_cloud_from_depth(depth, color, intrinsics: PinholeCameraIntrinsic, transformation=None,
                           depth_scale: float = 1000.0, depth_trunc: float = 3.0, device="cuda") -> PointCloud:
    """Array-level core of generate_point_cloud: depth [H,W] (any numeric dtype), colour [H,W,3] u8 or None."""
    d = torch.as_tensor(np.ascontiguousarray(depth).astype(np.float32) if not torch.is_tensor(depth) else depth)
    d = d.to(device=device, dtype=torch.float32).contiguous()[None]
    T12 = None
    if transformation is not None:
        T12 = torch.as_tensor(np.asarray(transformation, dtype=np.float64)[:3, :4].reshape(1, 12))
    xyz, valid, _ = ops.backproject(d, intrinsics.k4(), T12, depth_scale, depth_trunc, want_counts=False)
    keep = valid[0].bool()
    pts = xyz[0][keep]
    cols = None
    if color is not None:
        c = torch.as_tensor(np.ascontiguousarray(color)).to(device).reshape(-1, 3)
        cols = c[keep].float() / 255.0  # channel order as read from the file (BGR), like the reference
    return PointCloud(pts, cols)


def point_clouds_from_depth_batch(depths: np.ndarray, colors: np.ndarray | None, k4s: np.ndarray, t12s: np.ndarray | None,
                                  depth_scale: float = 1000.0, depth_trunc: float = 3.0, device="cuda") -> PointCloud:
    """Batched core: depths [B,H,W], colours [B,H,W,3] u8 or None, k4s [B,4], t12s [B,12] or None -> ONE cloud holding the
    frames' valid points in frame order (one dav2_backproject launch for the whole batch)."""
    d = torch.as_tensor(np.ascontiguousarray(depths).astype(np.float32)).to(device)
    K = torch.as_tensor(np.asarray(k4s, dtype=np.float64))
    T = None if t12s is None else torch.as_tensor(np.asarray(t12s, dtype=np.float64))
    xyz, valid, _ = ops.backproject(d, K, T, depth_scale, depth_trunc, want_counts=False)
    keep = valid.bool().reshape(-1)
    pts = xyz.reshape(-1, 3)[keep]
    cols = None
    if colors is not None:
        c = torch.as_tensor(np.ascontiguousarray(colors)).to(device).reshape(-1, 3)
        cols = c[keep].float() / 255.0
    return PointCloud(pts, cols)


def generate_point_cloud(depth_image_path: str, color_image_path: str, intrinsics_path: str, position_file: str,
                         rotation_file: str, frame_idx: int) -> PointCloud:
    """depth_to_pointcloud.py:178-241."""
    import cv2

    depth_image = cv2.imread(depth_image_path, cv2.IMREAD_UNCHANGED)
    color_image = cv2.imread(color_image_path)
    width, height = color_image.shape[:2]  # sic: the reference swaps the names (square frames)
    depth_image = cv2.resize(depth_image, (width, height), interpolation=cv2.INTER_NEAREST)
    intr = load_camera_intrinsics(intrinsics_path, width, height)
    T = load_transformation(position_file, rotation_file, frame_idx)
    return point_cloud_from_depth(depth_image, color_image, intr, T)


def get_procedure_files(rgb_filename: str) -> tuple:
    """depth_to_pointcloud.py:284-312."""
    path = Path(rgb_filename)
    procedure_dir = path.parent.parent
    sub = path.parent.name.split("_")[1]
    return (str(procedure_dir / "cam.txt"), str(procedure_dir / f"SavedPosition_{sub}.txt"),
            str(procedure_dir / f"SavedRotationQuaternion_{sub}.txt"))


def write_ply(path: str, cloud: PointCloud) -> None:
    """Binary little-endian PLY (x y z double, r g b uchar) -- the on-disk format of :368-371."""
    pts, cols = cloud.points, cloud.colors
    n = pts.shape[0]
    has_c = cols.shape[0] == n and n > 0
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {n}", "property double x", "property double y",
           "property double z"]
    dt = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if has_c:
        hdr += ["property uchar red", "property uchar green", "property uchar blue"]
        dt += [("r", "u1"), ("g", "u1"), ("b", "u1")]
    hdr.append("end_header")
    rec = np.empty(n, dtype=dt)
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if has_c:
        c8 = np.clip(np.round(cols * 255.0), 0, 255).astype(np.uint8)
        rec["r"], rec["g"], rec["b"] = c8[:, 0], c8[:, 1], c8[:, 2]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        f.write(rec.tobytes())


def input_output_files(args) -> tuple:
    """depth_to_pointcloud.py:53-122: (rgb files, depth files, output dir) from the CLI namespace (``img_path``,
    ``depth_path``, ``ds_type``, ``outdir``; ``outdir`` is filled in on ``args`` like the reference does).

    Quirks kept: with a FILE ``img_path`` the single-image branch hangs off the ``depth_path`` test (so a ``.txt`` image list
    with a non-``.txt`` depth path is replaced by ``[img_path]``); SimCol RGB frames are the ``Frames_*`` folders without
    ``_OP``, depths come from ``Frames_*_OP/depth``; lists are sorted per SyntheticColon_{I,II,III} sub-set."""
    rgb_filenames, depth_filenames = [], []
    if os.path.isfile(args.img_path):
        if args.img_path.endswith("txt"):
            with open(args.img_path, "r", encoding="utf-8") as f:
                rgb_filenames = f.read().splitlines()
        if args.depth_path.endswith("txt"):
            with open(args.depth_path, "r", encoding="utf-8") as f:
                depth_filenames = f.read().splitlines()
        else:
            rgb_filenames = [args.img_path]
            if args.outdir is None:
                args.outdir = str(Path(args.img_path).parent)
    elif args.ds_type == "simcol":
        base_dir = Path(args.img_path)
        for suffix in ("I", "II", "III"):
            rgb = glob.glob(str(base_dir / f"SyntheticColon_{suffix}/Frames_*/FrameBuffer_*.png"), recursive=True)
            rgb_filenames.extend(sorted(p for p in rgb if "_OP" not in str(p)))
            depth_filenames.extend(sorted(glob.glob(str(base_dir / f"SyntheticColon_{suffix}/Frames_*_OP/depth/Depth_*.png"),
                                                    recursive=True)))
        if args.outdir is None:
            args.outdir = str(base_dir)
    elif args.ds_type == "testing":
        base_dir = Path(args.img_path)
        rgb_filenames.extend(sorted(glob.glob(str(base_dir / "frame_*.jpg"), recursive=True)))
        if args.outdir is None:
            args.outdir = str(base_dir)
    return rgb_filenames, depth_filenames, args.outdir


def main(depth_image_paths: list, color_image_paths: list, output_dir: str, batch: int = 32) -> PointCloud:
    """depth_to_pointcloud.py:316-371 without the Poisson mesh (out of scope).  Frames are read ``batch`` at a time and
    every run of same-shaped frames is back-projected by ONE launch (per-frame intrinsics and poses are kernel inputs);
    the reference builds one Open3D cloud per frame."""
    import cv2

    combined = PointCloud()
    pairs = list(zip(depth_image_paths, color_image_paths))
    for s in range(0, len(pairs), batch):
        depths, colors, k4s, t12s = [], [], [], []
        for frame_idx, (dp, cp) in enumerate(pairs[s:s + batch], start=s):
            cam, pos, rot = get_procedure_files(cp)
            depth_image = cv2.imread(dp, cv2.IMREAD_UNCHANGED)
            color_image = cv2.imread(cp)
            width, height = color_image.shape[:2]  # sic (generate_point_cloud)
            depths.append(cv2.resize(depth_image, (width, height), interpolation=cv2.INTER_NEAREST))
            colors.append(color_image)
            k4s.append(load_camera_intrinsics(cam, width, height).k4())
            t12s.append(load_transformation(pos, rot, frame_idx)[:3, :4].reshape(12))
        i = 0
        while i < len(depths):  # runs of equal shape -> one launch each
            j = i + 1
            while j < len(depths) and depths[j].shape == depths[i].shape and colors[j].shape == colors[i].shape:
                j += 1
            combined += point_clouds_from_depth_batch(np.stack(depths[i:j]), np.stack(colors[i:j]), np.asarray(k4s[i:j]),
                                                      np.asarray(t12s[i:j]))
            i = j
    combined = combined.voxel_down_sample(voxel_size=0.01)  # :357-359
    os.makedirs(output_dir, exist_ok=True)
    write_ply(f"{output_dir}/combined_point_cloud.ply", combined)
    return combined
