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


# ------------------------------------------------------------------------------------------------
# Pose-pair items (reference ``data_processing/pose_estimation.py:205-311``): the same frame / depth transforms, two
# consecutive frames stacked to 8 channels, and the relative-pose target computed from the absolute poses.
# ------------------------------------------------------------------------------------------------
def relative_pose_targets(poses) -> torch.Tensor:
    """Absolute poses [N,7] (xyz + quaternion xyzw) -> the N-1 targets of ``PoseDataset.__getitem__`` (:245-303), all pairs
    at once with the item code's own fp32 operations in its order: unit translation direction
    ``(p2 - p1) / (|p2 - p1| + 1e-8)`` and ``normalize(q2 (x) conj(q1), eps=1e-8)`` with the product written as the
    reference writes it."""
    P = torch.as_tensor(np.asarray(poses) if not torch.is_tensor(poses) else poses).to(torch.float32)
    if P.dim() != 2 or P.shape[1] != 7:
        raise ValueError(f"poses must be [N,7], got {tuple(P.shape)}")
    p1, p2, qa, q2 = P[:-1, :3], P[1:, :3], P[:-1, 3:], P[1:, 3:]
    rel = p2 - p1
    rel = rel / (torch.linalg.vector_norm(rel, dim=1, keepdim=True) + 1e-8)
    q1 = qa * torch.tensor([-1.0, -1.0, -1.0, 1.0], dtype=torch.float32, device=P.device)  # conjugate (:262-264)
    x = q2[:, 0] * q1[:, 3] + q2[:, 1] * q1[:, 2] - q2[:, 2] * q1[:, 1] + q2[:, 3] * q1[:, 0]
    y = -q2[:, 0] * q1[:, 2] + q2[:, 1] * q1[:, 3] + q2[:, 2] * q1[:, 0] + q2[:, 3] * q1[:, 1]
    z = q2[:, 0] * q1[:, 1] - q2[:, 1] * q1[:, 0] + q2[:, 2] * q1[:, 3] + q2[:, 3] * q1[:, 2]
    w = -q2[:, 0] * q1[:, 0] - q2[:, 1] * q1[:, 1] - q2[:, 2] * q1[:, 2] + q2[:, 3] * q1[:, 3]
    q = torch.nn.functional.normalize(torch.stack([x, y, z, w], dim=1), dim=1, eps=1e-8)
    return torch.cat([rel, q], dim=1)


def pose_pair_items(images, depths, poses, transforms: SimColTransforms) -> dict:
    """A sequence of N decoded frames (uint8 [N,H,W,3]) and depths (uint16 [N,H,W]) with absolute poses [N,7] ->
    ``{"input": [N-1,8,S,S], "target": [N-1,7]}``: every frame goes through the transforms ONCE (the reference's loader
    transforms each frame twice, as the second half of pair i-1 and the first half of pair i), then
    ``cat(rgb_i, depth_i, rgb_{i+1}, depth_{i+1})`` (:229-243)."""
    from .pose_estimation_model import stack_pairs
    rgb, d = transforms.transform_input(images), transforms.transform_output(depths)
    return {"input": stack_pairs(rgb, d), "target": relative_pose_targets(poses).to(rgb.device)}
