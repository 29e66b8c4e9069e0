"""GPU form of the SimCol loader's per-item transforms (reference ``data_processing/simcol.py:104-135`` transform_input /
transform_output and ``:161-168`` __getitem__): ToTensor -> Resize((size, size), BICUBIC, antialias=True) [-> ImageNet
Normalize] for the frame, ``/ 65535`` + the same resize for the 16-bit ground-truth depth.

The reference runs them per item on the CPU (torchvision on float tensors inside DataLoader workers); here a whole batch
of decoded uint8 / uint16 arrays is uploaded as is (3 B / 2 B per pixel instead of 12 / 4) and ONE kernel per tensor does
division, torch's anti-aliased bicubic resampling and the normalisation (``dav2_resize_aa``).  Out of scope stays out:
file listing, train/val splitting and the Lightning DataModule."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _to_device(a, device):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    return t.to(device, non_blocking=True).contiguous()


class SimColTransforms:
    """``transform_input`` / ``transform_output`` of ``SimColDataset`` for decoded arrays or batches of them."""

    def __init__(self, size: int = 518, device="cuda"):
        self.size = int(size)
        self.device = torch.device(device)

    def transform_input(self, image) -> torch.Tensor:
        """uint8 RGB [H,W,3(+alpha)] or [B,H,W,3] (numpy or tensor) -> normalised fp32 [3,S,S] or [B,3,S,S] on the GPU."""
        single = image.ndim == 3
        x = _to_device(image[..., :3], self.device)  # simcol.py:161 ``[:, :, :3]``
        if x.dtype != torch.uint8:
            raise TypeError("transform_input expects the decoded uint8 frame (the /255 happens in the kernel)")
        out = ops.resize_aa(x[None] if single else x, self.size, self.size)
        return out[0] if single else out

    def transform_output(self, depth) -> torch.Tensor:
        """uint16 depth [H,W] or [B,H,W] -> fp32 [1,S,S] or [B,1,S,S] in [0,1] (simcol.py:163-165: / 65535, then resize)."""
        single = depth.ndim == 2
        x = _to_device(depth, self.device)
        if x.dtype not in (torch.uint16, torch.float32):
            raise TypeError("transform_output expects the decoded uint16 depth (or float32, divided by 1)")
        out = ops.resize_aa(x[None] if single else x, self.size, self.size)
        return out[0] if single else out

    def __call__(self, image, depth):
        return {"image": self.transform_input(image), "depth": self.transform_output(depth)}


def load_item(input_path: str, target_path: str, transforms: SimColTransforms) -> dict:
    """simcol.py:150-176 for one (frame, depth) pair of files."""
    from PIL import Image

    image = np.array(Image.open(input_path))[:, :, :3]
    depth = np.array(Image.open(target_path))
    if depth.dtype != np.uint16:
        depth = depth.astype(np.float32)
    return transforms(image, depth)
