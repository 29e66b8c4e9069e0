"""Seeded random-init weights for benchmarks and demos (there are no checkpoints offline).

Plain default init makes the sigmoid see ~0 everywhere (depth == max_depth/2), so scales are chosen
to keep activations alive through all blocks; the last 1x1 conv is rescaled on a probe input so
that pre-sigmoid logits have unit-order spread."""
from __future__ import annotations

import math

import torch


def _seed_for(key: str, seed: int) -> int:
    h = 1469598103
    for ch in key.encode():
        h = ((h ^ ch) * 16777619) & 0x7FFFFFFF
    return (h + 7919 * seed) & 0x7FFFFFFF


@torch.no_grad()
def randomize_(model, seed: int = 0):
    """In-place seeded init of a dav2_b200.dpt.DepthAnythingV2 (CPU tensors; call before .cuda())."""
    for k, p in model.state_dict().items():
        g = torch.Generator().manual_seed(_seed_for(k, seed))
        shape = tuple(p.shape)
        if k.endswith("gamma"):
            t = 0.5 + 0.5 * torch.rand(shape, generator=g)
        elif "norm" in k:
            t = 0.1 * torch.randn(shape, generator=g) + (1.0 if k.endswith("weight") else 0.0)
        elif k.endswith(("pos_embed", "cls_token", "mask_token")) or k.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[0] if ("resize_layers.0" in k or "resize_layers.1" in k) else int(math.prod(shape[1:]))
            t = torch.randn(shape, generator=g) * (1.4 / math.sqrt(fan_in))
        p.copy_(t)
    # keep logits O(1): output_conv2.2 sees 32 ReLU channels of O(1) variance
    w = model.depth_head.scratch.output_conv2[2].weight
    w.mul_(2.0 / max(float(w.norm()), 1e-6))
    model.mark_weights_changed()
    return model


@torch.no_grad()
def calibrate_(model, device="cuda", target_std=(3.0, 7.0), max_iter=12):
    """Rescale the last 1x1 conv ON THE GPU until the depth map is spread over (0, max_depth) without
    saturating: >15 % of pixels pinned within 2.5 % of either end -> halve the logit scale; depth std
    below the target -> grow it.  Uses the engine itself (no CPU model exists in the product)."""
    g = torch.Generator().manual_seed(4242)
    u = torch.rand(2, 3, 98, 98, generator=g)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = ((u - mean) / std).to(device)
    model.to(device)
    w = model.depth_head.scratch.output_conv2[2].weight
    model.depth_head.scratch.output_conv2[2].bias.zero_()
    md = float(model.max_depth)
    for _ in range(max_iter):
        model.mark_weights_changed()
        d = model(x)
        sat = float(((d < 0.025 * md) | (d > 0.975 * md)).float().mean())
        sd = float(d.std())
        if sat > 0.15:
            w.mul_(0.5)
        elif sd < target_std[0] * md / 20.0:
            w.mul_(1.6)
        elif sd > target_std[1] * md / 20.0:
            w.mul_(0.75)
        else:
            break
    model.mark_weights_changed()
    return model
