"""CPU oracle for the pose network (TEST INFRASTRUCTURE ONLY).

Restates ``PoseEstimationNet`` of the reference, pose_estimation_model.py:35-105 (that module imports
``lightning`` at import time, which is not installed, so it cannot be imported here): torchvision
ResNet-18, conv1 replaced by an ``in_channels``-wide 7x7/2 convolution without bias (:56-63), fc -> 256
(:66-67), head ReLU/Dropout(0.3)/Linear(256,128)/ReLU/Dropout(0.2)/Linear(128,64)/ReLU/Dropout(0.1)/
Linear(64,7) (:75-90), forward = head(backbone(x)) (:92-105).  The arithmetic itself is torchvision's
(installed), so the backbone is executed, not restated; parity is pinned by that execution."""
import torch
import torch.nn as nn


def build_pose_oracle(in_channels: int = 8, seed: int = 0) -> nn.Module:
    from torchvision.models import resnet18

    g = torch.Generator().manual_seed(seed)
    net = nn.Module()
    net.backbone = resnet18(weights=None)
    net.backbone.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
    net.backbone.fc = nn.Linear(512, 256)
    net.pose_head = nn.Sequential(nn.ReLU(), nn.Dropout(0.3), nn.Linear(256, 128), nn.ReLU(), nn.Dropout(0.2),
                                  nn.Linear(128, 64), nn.ReLU(), nn.Dropout(0.1), nn.Linear(64, 7))
    with torch.no_grad():  # seeded, non-trivial BatchNorm statistics so that the folding is really exercised
        for n, p in net.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.6 / fan_in ** 0.5))
            elif "bn" in n or "downsample.1" in n:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g) if n.endswith("weight") else 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
        for n, b in net.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(0.2 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))
    net.eval()

    def forward(x):
        return net.pose_head(net.backbone(x))

    net.forward = forward
    return net


def relative_pose_item(pose1, pose2) -> torch.Tensor:
    """ONE item's target exactly as ``PoseDataset.__getitem__`` computes it (data_processing/pose_estimation.py:245-303):
    per-item scalar torch fp32 code, restated line by line (the dataset class itself needs the SimCol files)."""
    import torch.nn.functional as F

    pos1 = torch.tensor(pose1[:3], dtype=torch.float32)
    pos2 = torch.tensor(pose2[:3], dtype=torch.float32)
    quat1 = torch.tensor(pose1[3:], dtype=torch.float32)
    quat2 = torch.tensor(pose2[3:], dtype=torch.float32)
    relative_pos = pos2 - pos1
    relative_pos = relative_pos / (torch.norm(relative_pos) + 1e-8)
    q1_inv = quat1 * torch.tensor([-1, -1, -1, 1], dtype=torch.float32)
    rq = torch.zeros(4, dtype=torch.float32)
    rq[0] = quat2[0] * q1_inv[3] + quat2[1] * q1_inv[2] - quat2[2] * q1_inv[1] + quat2[3] * q1_inv[0]
    rq[1] = -quat2[0] * q1_inv[2] + quat2[1] * q1_inv[3] + quat2[2] * q1_inv[0] + quat2[3] * q1_inv[1]
    rq[2] = quat2[0] * q1_inv[1] - quat2[1] * q1_inv[0] + quat2[2] * q1_inv[3] + quat2[3] * q1_inv[2]
    rq[3] = -quat2[0] * q1_inv[0] - quat2[1] * q1_inv[1] - quat2[2] * q1_inv[2] + quat2[3] * q1_inv[3]
    rq = F.normalize(rq, dim=0, eps=1e-8)
    return torch.cat([relative_pos, rq])
