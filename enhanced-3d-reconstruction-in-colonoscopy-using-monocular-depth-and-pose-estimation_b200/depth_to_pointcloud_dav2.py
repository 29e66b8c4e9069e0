"""Drop-in for the reference's ``depth_to_pointcloud_dav2.py``: image -> ``infer_image(image, height)`` ->
back-projection of EVERY pixel with the camera file's intrinsics (no validity filter, no pose) -> one PLY per frame
(depth_to_pointcloud_dav2.py:35-50 ``read_cam_file``, :53-69 ``load_transformation``, :159-187 checkpoint loading,
:189-240 input listing, :247-326 the frame loop).

Differences by design: frames are processed in BATCHES (equal-sized frames share one upload, one pre-processing
launch, one forward, one ``dav2_backproject`` launch); points are the kernel's fp32 roundings of the fp64 result the
reference's numpy code computes (<= 6e-8 relative, tests/test_gpu_parity_configs.py).  Two reference quirks are kept
so that outputs land where the reference puts them: for ``ds_type == "simcol"`` the per-sequence ``<Frames_x>_PC``
directory is created but the PLY is written to ``outdir`` (:279-285 vs :317-326), and the camera file is chosen by the
reference's substring tests (:253-264; see ``cam_file_for``)."""
from __future__ import annotations

import glob
import os
from pathlib import Path
from typing import Iterable, List, Optional

import numpy as np
import torch

from . import ops
from .depth_to_pointcloud import PointCloud, quat_to_matrix, write_ply
from .run import infer_images


def read_cam_file(cam_file: str) -> dict:
    """First line of the file = the row-major 3x3 K (:35-50)."""
    with open(cam_file, "r", encoding="utf-8") as f:
        v = [float(t) for t in f.readline().split()]
    return {"fx": v[0], "fy": v[4], "cx": v[2], "cy": v[5]}


def load_transformation(position_file: str, rotation_file: str) -> np.ndarray:
    """Single-pose files: position xyz + quaternion xyzw (normalised like scipy does) -> 4x4 (:53-69)."""
    T = np.eye(4)
    T[:3, :3] = quat_to_matrix(np.loadtxt(rotation_file))
    T[:3, 3] = np.loadtxt(position_file)
    return T


def load_checkpoint(model, path: str):
    """:163-185: a Lightning checkpoint's ``state_dict`` with the ``model.`` prefix stripped, or a bare state dict."""
    ckpt = torch.load(path, map_location="cpu")
    if "state_dict" in ckpt:
        sd = {(k[6:] if k.startswith("model.") else k): v for k, v in ckpt["state_dict"].items()}
        return model.load_state_dict(sd)
    return model.load_state_dict(ckpt)


def collect_filenames(img_path: str, ds_type: Optional[str] = None, outdir: Optional[str] = None) -> tuple:
    """(filenames, outdir) of :189-240."""
    filenames: List[str] = []
    if os.path.isfile(img_path):
        if img_path.endswith("txt"):
            with open(img_path, "r", encoding="utf-8") as f:
                filenames = f.read().splitlines()
        else:
            filenames = [img_path]
            if outdir is None:
                outdir = str(Path(img_path).parent)
    elif ds_type == "simcol":
        for suffix in ("I", "II", "III"):
            pattern = str(Path(img_path) / f"SyntheticColon_{suffix}/Frames_*/FrameBuffer_*.png")
            filenames.extend(p for p in glob.glob(pattern, recursive=True) if "_OP" not in str(p))
        if outdir is None:
            outdir = str(Path(img_path))
    elif ds_type == "testing":
        filenames.extend(glob.glob(str(Path(img_path) / "frame_*.jpg"), recursive=True))
        if outdir is None:
            outdir = str(Path(img_path))
    return filenames, outdir


def cam_file_for(filename: str, ds_type: Optional[str], cam_file: Optional[str]) -> str:
    """:253-268."""
    if ds_type == "simcol":
        base = Path("datasets/SyntheticColon")
        for suffix in ("I", "II", "III"):
            # the reference's plain substring tests in its order: "SyntheticColon_I" is a prefix of the other two, so
            # every SimCol frame resolves to SyntheticColon_I/cam.txt there -- kept, results must match the reference's
            if f"SyntheticColon_{suffix}" in str(Path(filename)):
                return str(base / f"SyntheticColon_{suffix}/cam.txt")
        raise ValueError(f"Unknown SyntheticColon suffix in {filename}")
    if cam_file:
        return cam_file
    raise ValueError("No camera file specified. Use --cam-file.")


@torch.no_grad()
def image_point_clouds(model, raws: List[np.ndarray], k4s) -> List[PointCloud]:
    """:287-314 for a list of BGR uint8 frames: depth at the frame's own resolution, then x = (u - cx) / fx * z,
    y = (v - cy) / fy * z for every pixel; colours RGB / 255.  ``k4s``: one (fx, fy, cx, cy) per frame."""
    dev = next(model.parameters()).device
    depths = infer_images(model, raws, None, dev)
    out: List[Optional[PointCloud]] = [None] * len(raws)
    groups = {}
    for i, r in enumerate(raws):
        groups.setdefault(tuple(r.shape[:2]), []).append(i)
    for idxs in groups.values():
        d = torch.stack([depths[i] for i in idxs])
        K = torch.as_tensor(np.asarray([k4s[i] for i in idxs], dtype=np.float64))
        xyz, _, _ = ops.backproject(d, K, None, 1.0, float("inf"), want_counts=False)
        # the reference keeps every pixel; the kernel writes zeros where z <= 0, which is what x*0, y*0, 0 is
        for j, i in enumerate(idxs):
            rgb = torch.from_numpy(np.ascontiguousarray(raws[i][:, :, ::-1])).to(dev).reshape(-1, 3)
            out[i] = PointCloud(xyz[j], rgb.float() / 255.0)
    return out


def process_frames(model, filenames: Iterable[str], outdir: str, ds_type: Optional[str] = None,
                   cam_file: Optional[str] = None, batch: int = 16) -> List[str]:
    """The frame loop of :247-326; returns the written PLY paths."""
    import cv2

    os.makedirs(outdir, exist_ok=True)
    filenames = list(filenames)
    cams = {}
    written = []
    for s in range(0, len(filenames), batch):
        names = filenames[s:s + batch]
        k4s = []
        for f in names:
            cf = cam_file_for(f, ds_type, cam_file)
            if cf not in cams:
                c = read_cam_file(cf)
                cams[cf] = (c["fx"], c["fy"], c["cx"], c["cy"])
            k4s.append(cams[cf])
            if ds_type == "simcol":
                fp = Path(f).parent
                os.makedirs(str(fp.parent / (fp.name + "_PC")), exist_ok=True)
        raws = [cv2.imread(f) for f in names]
        for f, cloud in zip(names, image_point_clouds(model, raws, k4s)):
            path = os.path.join(outdir, os.path.splitext(os.path.basename(f))[0] + ".ply")
            write_ply(path, cloud)
            written.append(path)
    return written
