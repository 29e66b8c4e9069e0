"""Frame-loop semantics of the reference's ``run.py:195-262`` on top of the B200 engine, with the output writers
of ``run.py:239-262`` (raw ``.npy`` depth, min-max normalised 8-bit PNG, optional side-by-side with the input).

Differences by design: frames are processed in BATCHES (the reference runs batch 1 with one sync per frame);
the Spectral colour map needs matplotlib, which the reference imports but this image lacks, so colour output
falls back to grayscale unless matplotlib is importable."""
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


def colorize(depth_u8: np.ndarray, grayscale: bool) -> np.ndarray:
    """run.py:245-248: grayscale x3, or matplotlib 'Spectral_r' -> BGR."""
    if not grayscale:
        try:
            import matplotlib
            cmap = matplotlib.colormaps.get_cmap("Spectral_r")
            return (cmap(depth_u8)[:, :, :3] * 255)[:, :, ::-1].astype(np.uint8)
        except ImportError:
            pass
    return np.repeat(depth_u8[..., np.newaxis], 3, axis=-1)


def output_path(filename: str, outdir: str) -> str:
    return os.path.join(outdir, os.path.splitext(os.path.basename(filename))[0] + ".png")


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
        tens = [model.image2tensor(r, input_size) for r in raws]
        # batch frames that share a network input shape; others run alone (infer_image semantics per frame)
        groups = {}
        for i, (t, hw) in enumerate(tens):
            groups.setdefault(tuple(t.shape[-2:]), []).append(i)
        depths = [None] * len(names)
        for shape, idxs in groups.items():
            x = torch.cat([tens[i][0] for i in idxs]).to(dev)
            d = model(x)
            for j, i in enumerate(idxs):
                h, w = tens[i][1]
                depths[i] = ops.resize_depth(d[j:j + 1].contiguous(), h, w)[0].cpu().numpy()
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
