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
