"""CPU oracle: torch restatement of DepthAnythingV2 (DINOv2 ViT + DPT head).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  The product (``dav2_b200``) never falls back to this code.

What it restates.  The reference imports the model from an UN-VENDORED third-party checkout
(``Depth_Anything_V2.metric_depth.depth_anything_v2.dpt`` -- reference ``run.py:44``,
``lightning_model.py:16``, ``depth_to_pointcloud_dav2.py:32``); the package is absent from
/root/reference and from this image and is unpinned (no submodule, no version in
``requirements.txt:1-24``).  This file therefore restates the *published* algorithm of
github.com/DepthAnything/Depth-Anything-V2 ``metric_depth/depth_anything_v2/{dpt.py,dinov2.py,
dinov2_layers/*,util/blocks.py,util/transform.py}`` (structure summarised in SURVEY.md App. A)
and anchors parity on the reference's own call sites:

* constructor kwargs            run.py:97-125, lightning_model.py:116-121
* ``forward(x[B,3,H,W]) -> [B,H,W]``   lightning_model.py:301-302 (caller unsqueezes)
* ``infer_image(bgr_u8, input_size)``  run.py:234, depth_to_pointcloud_dav2.py:291
* state-dict key contract       run.py:128-147, lightning_model.py:130-140 ("pretrained" filter)

PARITY PINNING.  The reference has no tests and no golden vectors for the model
(SURVEY.md section 4 / 8c), so the model arithmetic is "parity unpinned" by the reference itself.
What pins this oracle instead (tests/test_oracle_model.py):
  1. integer parameter counts 24 785 089 / 97 470 785 / 335 315 649 (vits/vitb/vitl);
  2. exact state-dict key set of the upstream checkpoints (App. A.4);
  3. numerical agreement (fp32, <=1e-5) with the independent implementation shipped in
     ``transformers`` (``DepthAnythingForDepthEstimation``, metric head) through a key remap, at
     518x518 where no position-embedding interpolation occurs.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# run.py:97-118 (vitg omitted: no DINOv2-g metric checkpoint is used by the reference configs)
MODEL_CONFIGS: Dict[str, dict] = {
    "vits": {"encoder": "vits", "features": 64, "out_channels": [48, 96, 192, 384]},
    "vitb": {"encoder": "vitb", "features": 128, "out_channels": [96, 192, 384, 768]},
    "vitl": {"encoder": "vitl", "features": 256, "out_channels": [256, 512, 1024, 1024]},
}
# upstream dinov2.py vit_small / vit_base / vit_large
ENCODER_DIMS = {
    "vits": dict(embed_dim=384, depth=12, num_heads=6),
    "vitb": dict(embed_dim=768, depth=12, num_heads=12),
    "vitl": dict(embed_dim=1024, depth=24, num_heads=16),
}
# upstream dpt.py ``intermediate_layer_idx``
TAP_LAYERS = {"vits": [2, 5, 8, 11], "vitb": [2, 5, 8, 11], "vitl": [4, 11, 17, 23]}

PATCH = 14
IMG_SIZE = 518
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------------------
# DINOv2 encoder (dinov2.py + dinov2_layers/{patch_embed,attention,mlp,layer_scale,block}.py)
# --------------------------------------------------------------------------------------
class PatchEmbed(nn.Module):
    def __init__(self, embed_dim: int):
        super().__init__()
        self.proj = nn.Conv2d(3, embed_dim, kernel_size=PATCH, stride=PATCH)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)  # B, ph*pw, D (row-major over py,px)


class Attention(nn.Module):
    """Plain (non-xformers) attention: q is scaled BEFORE q@k^T."""

    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, bias=True)
        self.act = nn.GELU()  # exact erf
        self.fc2 = nn.Linear(hidden, dim, bias=True)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class LayerScale(nn.Module):
    def __init__(self, dim: int, init_values: float = 1.0):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class Block(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = LayerScale(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, dim * 4)
        self.ls2 = LayerScale(dim)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        x = x + self.ls2(self.mlp(self.norm2(x)))
        return x


class DinoVisionTransformer(nn.Module):
    def __init__(self, embed_dim: int, depth: int, num_heads: int):
        super().__init__()
        self.embed_dim = embed_dim
        self.patch_size = PATCH
        self.interpolate_offset = 0.1
        self.patch_embed = PatchEmbed(embed_dim)
        n = (IMG_SIZE // PATCH) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))  # present in checkpoints, unused in eval
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)

    def interpolate_pos_encoding(self, npatch: int, w: int, h: int) -> torch.Tensor:
        """Upstream dinov2.py ``interpolate_pos_encoding`` (bicubic, scale_factor with +0.1 offset)."""
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        pos = self.pos_embed.float()
        cls_pos, patch_pos = pos[:, 0], pos[:, 1:]
        dim = pos.shape[-1]
        w0, h0 = w // PATCH, h // PATCH
        sqrt_n = math.sqrt(N)
        sx = float(w0 + self.interpolate_offset) / sqrt_n
        sy = float(h0 + self.interpolate_offset) / sqrt_n
        patch_pos = F.interpolate(
            patch_pos.reshape(1, int(sqrt_n), int(sqrt_n), dim).permute(0, 3, 1, 2),
            scale_factor=(sx, sy),
            mode="bicubic",
            antialias=False,
        )
        assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
        patch_pos = patch_pos.permute(0, 2, 3, 1).view(1, -1, dim)
        return torch.cat((cls_pos.unsqueeze(0), patch_pos), dim=1)

    def prepare_tokens(self, x):
        B, _, w, h = x.shape  # upstream naming: dim 2 is called "w"
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        return x + self.interpolate_pos_encoding(x.shape[1] - 1, w, h)

    def get_intermediate_layers(self, x, idx: Sequence[int], return_class_token=True, norm=True):
        x = self.prepare_tokens(x)
        outs = []
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            if i in idx:
                outs.append(x)
        if norm:
            outs = [self.norm(o) for o in outs]
        cls = [o[:, 0] for o in outs]
        outs = [o[:, 1:] for o in outs]
        return tuple(zip(outs, cls)) if return_class_token else tuple(outs)


# --------------------------------------------------------------------------------------
# DPT head (dpt.py + util/blocks.py)
# --------------------------------------------------------------------------------------
class ResidualConvUnit(nn.Module):
    def __init__(self, features: int):
        super().__init__()
        self.conv1 = nn.Conv2d(features, features, 3, padding=1, bias=True)
        self.conv2 = nn.Conv2d(features, features, 3, padding=1, bias=True)

    def forward(self, x):
        out = F.relu(x)  # activation is nn.ReLU(False): the skip keeps the pre-ReLU x
        out = self.conv1(out)
        out = F.relu(out)
        out = self.conv2(out)
        return out + x


class FeatureFusionBlock(nn.Module):
    def __init__(self, features: int, size=None):
        super().__init__()
        self.out_conv = nn.Conv2d(features, features, 1, bias=True)
        self.resConfUnit1 = ResidualConvUnit(features)
        self.resConfUnit2 = ResidualConvUnit(features)
        self.size = size

    def forward(self, *xs, size=None):
        out = xs[0]
        if len(xs) == 2:
            out = out + self.resConfUnit1(xs[1])
        out = self.resConfUnit2(out)
        if size is None and self.size is None:
            mod = {"scale_factor": 2}
        elif size is None:
            mod = {"size": self.size}
        else:
            mod = {"size": size}
        out = F.interpolate(out, **mod, mode="bilinear", align_corners=True)
        return self.out_conv(out)


class _Scratch(nn.Module):
    pass


class DPTHead(nn.Module):
    def __init__(self, in_channels: int, features: int, out_channels: List[int]):
        super().__init__()
        oc = out_channels
        self.projects = nn.ModuleList([nn.Conv2d(in_channels, c, 1) for c in oc])
        self.resize_layers = nn.ModuleList(
            [
                nn.ConvTranspose2d(oc[0], oc[0], kernel_size=4, stride=4, padding=0),
                nn.ConvTranspose2d(oc[1], oc[1], kernel_size=2, stride=2, padding=0),
                nn.Identity(),
                nn.Conv2d(oc[3], oc[3], kernel_size=3, stride=2, padding=1),
            ]
        )
        s = _Scratch()
        s.layer1_rn = nn.Conv2d(oc[0], features, 3, padding=1, bias=False)
        s.layer2_rn = nn.Conv2d(oc[1], features, 3, padding=1, bias=False)
        s.layer3_rn = nn.Conv2d(oc[2], features, 3, padding=1, bias=False)
        s.layer4_rn = nn.Conv2d(oc[3], features, 3, padding=1, bias=False)
        s.refinenet1 = FeatureFusionBlock(features)
        s.refinenet2 = FeatureFusionBlock(features)
        s.refinenet3 = FeatureFusionBlock(features)
        s.refinenet4 = FeatureFusionBlock(features)
        s.output_conv1 = nn.Conv2d(features, features // 2, 3, padding=1)
        s.output_conv2 = nn.Sequential(
            nn.Conv2d(features // 2, 32, 3, padding=1),
            nn.ReLU(True),
            nn.Conv2d(32, 1, 1),
            nn.Sigmoid(),
        )
        self.scratch = s

    def forward(self, taps, ph: int, pw: int, return_logits: bool = False):
        out = []
        for i, (x, _cls) in enumerate(taps):  # use_clstoken=False: cls ignored
            x = x.permute(0, 2, 1).reshape(x.shape[0], x.shape[-1], ph, pw)
            x = self.projects[i](x)
            x = self.resize_layers[i](x)
            out.append(x)
        l1, l2, l3, l4 = out
        s = self.scratch
        l1r, l2r, l3r, l4r = s.layer1_rn(l1), s.layer2_rn(l2), s.layer3_rn(l3), s.layer4_rn(l4)
        p4 = s.refinenet4(l4r, size=l3r.shape[2:])
        p3 = s.refinenet3(p4, l3r, size=l2r.shape[2:])
        p2 = s.refinenet2(p3, l2r, size=l1r.shape[2:])
        p1 = s.refinenet1(p2, l1r)
        o = s.output_conv1(p1)
        o = F.interpolate(o, (int(ph * PATCH), int(pw * PATCH)), mode="bilinear", align_corners=True)
        if return_logits:  # pre-sigmoid logits, for non-vacuous parity
            o = s.output_conv2[0](o)
            o = s.output_conv2[1](o)
            return s.output_conv2[2](o)
        return s.output_conv2(o)


class DepthAnythingV2(nn.Module):
    """Same constructor / forward / infer_image surface as upstream ``dpt.DepthAnythingV2``."""

    def __init__(self, encoder="vitl", features=256, out_channels=(256, 512, 1024, 1024),
                 use_bn=False, use_clstoken=False, max_depth=20.0):
        super().__init__()
        assert not use_bn and not use_clstoken, "reference configs never enable these (run.py:97-125)"
        self.intermediate_layer_idx = TAP_LAYERS
        self.max_depth = max_depth
        self.encoder = encoder
        self.pretrained = DinoVisionTransformer(**ENCODER_DIMS[encoder])
        self.depth_head = DPTHead(self.pretrained.embed_dim, features, list(out_channels))

    def forward_taps(self, x):
        return self.pretrained.get_intermediate_layers(
            x, self.intermediate_layer_idx[self.encoder], return_class_token=True)

    def forward(self, x):
        ph, pw = x.shape[-2] // PATCH, x.shape[-1] // PATCH
        feats = self.forward_taps(x)
        depth = self.depth_head(feats, ph, pw) * self.max_depth
        return depth.squeeze(1)

    def forward_logits(self, x):
        ph, pw = x.shape[-2] // PATCH, x.shape[-1] // PATCH
        return self.depth_head(self.forward_taps(x), ph, pw, return_logits=True).squeeze(1)

    @torch.no_grad()
    def infer_image(self, raw_image: np.ndarray, input_size: int = 518) -> np.ndarray:
        image, (h, w) = image2tensor(raw_image, input_size)
        image = image.to(next(self.parameters()).device)
        depth = self.forward(image)
        depth = F.interpolate(depth[:, None], (h, w), mode="bilinear", align_corners=True)[0, 0]
        return depth.cpu().numpy()


# --------------------------------------------------------------------------------------
# infer_image pre-processing (upstream dpt.py image2tensor + util/transform.py)
# --------------------------------------------------------------------------------------
def resize_target(h: int, w: int, input_size: int, multiple: int = PATCH) -> Tuple[int, int]:
    """Resize(keep_aspect_ratio=True, ensure_multiple_of=14, resize_method='lower_bound')."""
    scale_h, scale_w = input_size / h, input_size / w
    # lower_bound: scale so that both sides are >= input_size, i.e. use the larger scale
    if scale_w > scale_h:
        scale_h = scale_w
    else:
        scale_w = scale_h

    def _mult(x):
        y = int(np.round(x / multiple) * multiple)
        if y < input_size:  # min_val = input_size for lower_bound
            y = int(np.ceil(x / multiple) * multiple)
        return y

    return _mult(scale_h * h), _mult(scale_w * w)


def image2tensor(raw_image: np.ndarray, input_size: int = 518):
    import cv2

    h, w = raw_image.shape[:2]
    image = cv2.cvtColor(raw_image, cv2.COLOR_BGR2RGB) / 255.0  # float64
    nh, nw = resize_target(h, w, input_size)
    image = cv2.resize(image, (nw, nh), interpolation=cv2.INTER_CUBIC)
    image = (image - np.asarray(IMAGENET_MEAN)) / np.asarray(IMAGENET_STD)
    image = np.ascontiguousarray(np.transpose(image, (2, 0, 1))).astype(np.float32)
    return torch.from_numpy(image).unsqueeze(0), (h, w)


# --------------------------------------------------------------------------------------
# Seeded, NON-DEGENERATE weights: default init gives depth == max_depth/2
# everywhere, which would make any parity test vacuous.
# --------------------------------------------------------------------------------------
def _stable_seed(key: str, seed: int) -> int:
    h = 2166136261
    for ch in key.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return (h ^ (seed * 2654435761)) & 0x7FFFFFFF


def make_state_dict(encoder: str, seed: int = 0, features=None, out_channels=None) -> Dict[str, torch.Tensor]:
    """Deterministic fp32 state dict keyed by upstream names; independent of torch's global RNG."""
    cfg = MODEL_CONFIGS[encoder]
    m = DepthAnythingV2(encoder, features or cfg["features"], out_channels or cfg["out_channels"])
    sd = {}
    for k, v in m.state_dict().items():
        g = torch.Generator().manual_seed(_stable_seed(k, seed))
        shape = tuple(v.shape)
        if k.endswith("gamma"):
            t = 0.5 + 0.5 * torch.rand(shape, generator=g)
        elif ".norm" in k or k.endswith("pretrained.norm.weight") or k.endswith("pretrained.norm.bias"):
            t = 0.1 * torch.randn(shape, generator=g)
            if k.endswith("weight"):
                t = t + 1.0
        elif k.endswith("pos_embed") or k.endswith("cls_token") or k.endswith("mask_token"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif k.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:  # conv / linear / conv-transpose weights
            if "resize_layers.0" in k or "resize_layers.1" in k:
                fan_in = shape[0]  # ConvTranspose2d weight is [Cin, Cout, k, k]; each output sees Cin taps
            else:
                fan_in = int(np.prod(shape[1:]))
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
            if ".mlp.fc1." in k or "conv" in k or "layer" in k and "_rn" in k:
                t = t * 1.4  # keep activations alive through GELU / ReLU stages
        sd[k] = t.to(torch.float32)
    return sd


def calibrate_logit_scale(model: DepthAnythingV2, x: torch.Tensor, target_std: float = 2.0) -> float:
    """Rescale the last 1x1 conv so pre-sigmoid logits have std ~= target_std on ``x``."""
    with torch.no_grad():
        logits = model.forward_logits(x)
        w = model.depth_head.scratch.output_conv2[2]
        b = float(w.bias)
        std = float((logits - b).std())
        s = target_std / max(std, 1e-12)
        w.weight.mul_(s)
        w.bias.zero_()
    return s


def build_oracle(encoder: str, seed: int = 0, max_depth: float = 20.0, calibrate: bool = True) -> DepthAnythingV2:
    """Oracle model with seeded non-degenerate weights (fp32, eval mode, CPU)."""
    cfg = MODEL_CONFIGS[encoder]
    m = DepthAnythingV2(encoder, cfg["features"], cfg["out_channels"], max_depth=max_depth)
    m.load_state_dict(make_state_dict(encoder, seed))
    m.eval()
    if calibrate:
        g = torch.Generator().manual_seed(4242)
        x = synthetic_frames(1, 98, 98, generator=g)  # 7x7 patches: cheap, input independent enough
        calibrate_logit_scale(m, x)
    return m


def synthetic_frames(B: int, H: int = 518, W: int = 518, generator=None, seed: int = 1234) -> torch.Tensor:
    """SimCol-shaped synthetic input: u~U[0,1) then ImageNet-normalised."""
    g = generator or torch.Generator().manual_seed(seed)
    u = torch.rand(B, 3, H, W, generator=g)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return (u - mean) / std


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
