"""Frame-loop semantics of the reference's ``run.py:195-262`` on top of the B200 engine, with the output writers
of ``run.py:239-262`` (raw ``.npy`` depth, min-max normalised 8-bit PNG, optional side-by-side with the input).

Differences by design: frames are processed in BATCHES (the reference runs batch 1 with one sync per frame).
The colour output uses matplotlib's ``Spectral`` map exactly as ``run.py:160,245-248`` does; matplotlib is not in
this image, so its 256-entry lookup table is rebuilt here from the 11 ColorBrewer anchors with matplotlib's own
``LinearSegmentedColormap`` arithmetic (``spectral_lut``) -- colour PNGs never silently degrade to grayscale."""
from __future__ import annotations

import os
from typing import Iterable, List

import numpy as np
import torch

from . import ops


def depth_to_uint8(depth: np.ndarray) -> np.ndarray:
    """run.py:242-243: (d - min) / (max - min) * 255 -> uint8."""
    d = (depth - depth.min()) / (depth.max() - depth.min()) * 255.0
    return d.astype(np.uint8)


# ColorBrewer 11-class "Spectral" anchors = matplotlib._cm._Spectral_data (8-bit values / 255)
_SPECTRAL_ANCHORS = np.array([(158, 1, 66), (213, 62, 79), (244, 109, 67), (253, 174, 97), (254, 224, 139), (255, 255, 191),
                              (230, 245, 152), (171, 221, 164), (102, 194, 165), (50, 136, 189), (94, 79, 162)],
                             dtype=np.float64) / 255.0
_SPECTRAL_LUT = None


def spectral_lut() -> np.ndarray:
    """matplotlib.colormaps.get_cmap("Spectral") as its [256,3] float64 lookup table (run.py:160): piecewise-linear through the
    anchors at x = linspace(0,1,11), sampled at linspace(0,1,256) the way matplotlib.colors._create_lookup_table does."""
    global _SPECTRAL_LUT
    if _SPECTRAL_LUT is None:
        x = np.linspace(0.0, 1.0, len(_SPECTRAL_ANCHORS))
        xind = np.linspace(0.0, 1.0, 256)
        ind = np.searchsorted(x, xind)[1:-1]
        dist = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
        y = _SPECTRAL_ANCHORS
        mid = dist[:, None] * (y[ind] - y[ind - 1]) + y[ind - 1]
        _SPECTRAL_LUT = np.clip(np.concatenate([y[:1], mid, y[-1:]]), 0.0, 1.0)
    return _SPECTRAL_LUT


def colorize(depth_u8: np.ndarray, grayscale: bool) -> np.ndarray:
    """run.py:245-248: grayscale x3, or ``(cmap(depth)[:, :, :3] * 255)[:, :, ::-1].astype(uint8)`` with cmap = matplotlib
    'Spectral' (an integer image indexes the colour table directly) -> BGR."""
    if grayscale:
        return np.repeat(depth_u8[..., np.newaxis], 3, axis=-1)
    return (spectral_lut()[depth_u8] * 255)[:, :, ::-1].astype(np.uint8)


def output_path(filename: str, outdir: str) -> str:
    return os.path.join(outdir, os.path.splitext(os.path.basename(filename))[0] + ".png")


@torch.no_grad()
def infer_images(model, raws: List[np.ndarray], input_size=518, device=None) -> List[torch.Tensor]:
    """``infer_image`` for a list of BGR uint8 frames, batched: frames of equal size share ONE upload of the uint8
    batch, one pre-processing launch (OpenCV-compatible bicubic + normalisation on the GPU, upstream image2tensor), one
    forward and one resize back.  ``input_size=None`` means "the frame's own height", the way
    depth_to_pointcloud_dav2.py:291 calls it.  Returns per-frame device tensors [h,w] in input order."""
    dev = device if device is not None else next(model.parameters()).device
    groups = {}
    for i, r in enumerate(raws):
        groups.setdefault(tuple(r.shape[:2]), []).append(i)
    depths = [None] * len(raws)
    for (h, w), idxs in groups.items():
        nh, nw = model.target_size(h, w, h if input_size is None else input_size)
        u8 = torch.from_numpy(np.stack([raws[i] for i in idxs])).to(dev, non_blocking=True)
        d = ops.resize_depth(model(ops.preprocess_bgr_u8(u8, nh, nw)), h, w)
        for j, i in enumerate(idxs):
            depths[i] = d[j]
    return depths


@torch.no_grad()
def run_frames(model, filenames: Iterable[str], outdir: str, input_size: int = 518, save_numpy: bool = False,
               pred_only: bool = True, grayscale: bool = True, batch: int = 16, skip_existing: bool = True) -> List[str]:
    """Process image files like the run.py loop; returns the list of written PNG paths."""
    import cv2

    os.makedirs(outdir, exist_ok=True)
    todo = [f for f in filenames if not (skip_existing and os.path.exists(output_path(f, outdir)))]  # run.py:228-230
    written = []
    dev = next(model.parameters()).device
    for s in range(0, len(todo), batch):
        names = todo[s:s + batch]
        raws = [cv2.imread(f) for f in names]
        depths = [d.cpu().numpy() for d in infer_images(model, raws, input_size, dev)]
        for name, raw, depth in zip(names, raws, depths):
            stem = os.path.join(outdir, os.path.splitext(os.path.basename(name))[0])
            if save_numpy:
                np.save(stem + ".npy", depth)  # run.py:239-240 (same stem as the PNG)
            vis = colorize(depth_to_uint8(depth), grayscale)
            if not pred_only:
                split = np.ones((raw.shape[0], 50, 3), dtype=np.uint8) * 255  # run.py:253-262
                vis = cv2.hconcat([raw, split, vis])
            cv2.imwrite(stem + ".png", vis)
            written.append(stem + ".png")
    return written
