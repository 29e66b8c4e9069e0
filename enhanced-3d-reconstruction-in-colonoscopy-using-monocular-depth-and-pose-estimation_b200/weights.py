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
