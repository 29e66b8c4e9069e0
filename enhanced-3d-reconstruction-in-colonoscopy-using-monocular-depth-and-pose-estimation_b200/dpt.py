"""Drop-in for the external ``Depth_Anything_V2.metric_depth.depth_anything_v2.dpt`` module the
reference imports (run.py:44, lightning_model.py:16, depth_to_pointcloud_dav2.py:32).

``DepthAnythingV2`` keeps the constructor, ``forward`` / ``infer_image`` signatures, attribute
names and state-dict keys of upstream, so ``run.py:120-149`` and ``lightning_model.py:116-140``
work unchanged, but the arithmetic runs in libdav2_b200.so (tcgen05 GEMMs / implicit-GEMM convs,
fused attention, ...).  One extra keyword, ``precision`` ("fp16" default = the reference's Lightning
``16-mixed`` AMP, configs/trainer/default.yaml:4; or "bf16"), picks the tensor-core operand format;
accumulation, the residual stream, LayerNorm statistics and softmax are always fp32.  ``precision="fp32"`` selects the
slow all-fp32 validation engine (the 1e-4 parity gate).  The ``nn`` layers below are PARAMETER CONTAINERS ONLY (they give
the upstream parameter names and shapes); they are never called.  No CPU path exists: calling
``forward`` with the module or input off the GPU raises.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from ._lib import Dav2Config, Dav2Error, check

# run.py:97-118
MODEL_CONFIGS = {
    "vits": {"encoder": "vits", "features": 64, "out_channels": [48, 96, 192, 384]},
    "vitb": {"encoder": "vitb", "features": 128, "out_channels": [96, 192, 384, 768]},
    "vitl": {"encoder": "vitl", "features": 256, "out_channels": [256, 512, 1024, 1024]},
}
_ENCODERS = {"vits": (384, 12, 6), "vitb": (768, 12, 12), "vitl": (1024, 24, 16)}
_TAPS = {"vits": [2, 5, 8, 11], "vitb": [2, 5, 8, 11], "vitl": [4, 11, 17, 23]}
_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


class _Holder(nn.Module):
    """Attribute bag; never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise Dav2Error("parameter container: the computation lives in libdav2_b200.so")


def _encoder_params(D: int, depth: int) -> nn.Module:
    enc = _Holder()
    enc.embed_dim = D
    enc.cls_token = nn.Parameter(torch.zeros(1, 1, D))
    enc.pos_embed = nn.Parameter(torch.zeros(1, 37 * 37 + 1, D))
    enc.mask_token = nn.Parameter(torch.zeros(1, D))
    enc.patch_embed = _Holder()
    enc.patch_embed.proj = nn.Conv2d(3, D, 14, 14)
    blocks = []
    for _ in range(depth):
        b = _Holder()
        b.norm1 = nn.LayerNorm(D, eps=1e-6)
        b.attn = _Holder()
        b.attn.qkv = nn.Linear(D, 3 * D)
        b.attn.proj = nn.Linear(D, D)
        b.ls1 = _Holder()
        b.ls1.gamma = nn.Parameter(torch.ones(D))
        b.norm2 = nn.LayerNorm(D, eps=1e-6)
        b.mlp = _Holder()
        b.mlp.fc1 = nn.Linear(D, 4 * D)
        b.mlp.fc2 = nn.Linear(4 * D, D)
        b.ls2 = _Holder()
        b.ls2.gamma = nn.Parameter(torch.ones(D))
        blocks.append(b)
    enc.blocks = nn.ModuleList(blocks)
    enc.norm = nn.LayerNorm(D, eps=1e-6)
    return enc


def _head_params(D: int, Fe: int, oc) -> nn.Module:
    h = _Holder()
    h.projects = nn.ModuleList([nn.Conv2d(D, c, 1) for c in oc])
    h.resize_layers = nn.ModuleList([
        nn.ConvTranspose2d(oc[0], oc[0], 4, 4), nn.ConvTranspose2d(oc[1], oc[1], 2, 2), nn.Identity(),
        nn.Conv2d(oc[3], oc[3], 3, 2, 1)])
    s = _Holder()
    for i in range(4):
        setattr(s, f"layer{i + 1}_rn", nn.Conv2d(oc[i], Fe, 3, padding=1, bias=False))
        r = _Holder()
        r.out_conv = nn.Conv2d(Fe, Fe, 1)
        for u in (1, 2):
            rcu = _Holder()
            rcu.conv1 = nn.Conv2d(Fe, Fe, 3, padding=1)
            rcu.conv2 = nn.Conv2d(Fe, Fe, 3, padding=1)
            setattr(r, f"resConfUnit{u}", rcu)
        setattr(s, f"refinenet{i + 1}", r)
    s.output_conv1 = nn.Conv2d(Fe, Fe // 2, 3, padding=1)
    s.output_conv2 = nn.Sequential(nn.Conv2d(Fe // 2, 32, 3, padding=1), nn.ReLU(True), nn.Conv2d(32, 1, 1), nn.Sigmoid())
    h.scratch = s
    return h


class DepthAnythingV2(nn.Module):
    def __init__(self, encoder="vitl", features=256, out_channels=(256, 512, 1024, 1024), use_bn=False,
                 use_clstoken=False, max_depth=20.0, precision="fp16"):
        super().__init__()
        if precision not in ("fp16", "bf16", "fp32"):
            raise ValueError("precision must be 'fp16' (the reference's AMP 16-mixed; default), 'bf16', or 'fp32' "
                             "(the slow all-fp32 validation engine of the 1e-4 parity gate)")
        self.precision = precision
        if encoder not in _ENCODERS:
            raise ValueError(f"unsupported encoder {encoder!r} (vits | vitb | vitl)")
        if use_bn or use_clstoken:
            raise NotImplementedError("use_bn / use_clstoken are never enabled by the reference (run.py:97-125)")
        D, depth, heads = _ENCODERS[encoder]
        self.intermediate_layer_idx = _TAPS
        self.encoder = encoder
        self.max_depth = max_depth
        self.pretrained = _encoder_params(D, depth)
        self.depth_head = _head_params(D, features, list(out_channels))
        self._cfg = (D, depth, heads, int(features), [int(c) for c in out_channels])
        self._handle = None
        self._handle_device = None
        self._dirty = True
        self._pos_grids = set()

    # ---- engine lifecycle ---------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._dirty = True  # .to() / .half() / .cuda(): re-pack on next use
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._dirty = True
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def mark_weights_changed(self):
        """Call after mutating parameters in place (e.g. ``p.data.copy_``) so they are re-packed."""
        self._dirty = True

    def _release(self):
        if self._handle is not None:
            _lib.load().dav2_destroy(self._handle)
            self._handle = None
        self._handle_device = None
        self._pos_grids = set()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_engine(self, device):
        """One engine (packed weights + workspace) on ONE device, used from one stream at a time (include/dav2_b200.h:
        "one handle per (device, host thread)").  An input on another GPU re-packs the engine there instead of launching
        against the first device's weights."""
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._handle is not None and not self._dirty and self._handle_device == device:
            return
        lib = _lib.load()
        self._release()
        D, depth, heads, Fe, oc = self._cfg
        cfg = Dav2Config(D, depth, heads, Fe, (C.c_int32 * 4)(*oc), (C.c_int32 * 4)(*_TAPS[self.encoder]), float(self.max_depth),
                         {"fp16": _lib.FMT_F16, "bf16": _lib.FMT_BF16, "fp32": _lib.FMT_F32}[self.precision])
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.dav2_create(C.byref(h), C.byref(cfg)), "dav2_create")
            self._handle = h
            for k, v in self.state_dict().items():
                t = v.detach().to(device="cpu", dtype=torch.float32).contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                check(lib.dav2_set_weight(h, k.encode(), t.data_ptr(), shape, t.dim()), f"dav2_set_weight({k})")
            if not lib.dav2_weights_complete(h):
                raise Dav2Error("weights incomplete: " + lib.dav2_last_error().decode())
            if getattr(self, "_capture_logits", False):
                check(lib.dav2_set_capture_logits(h, 1), "dav2_set_capture_logits")
        self._handle_device = device
        self._dirty = False

    def _pos_table(self, ph: int, pw: int) -> torch.Tensor:
        """Upstream DinoVisionTransformer.interpolate_pos_encoding (bicubic, +0.1 offset), host side, once per grid."""
        pos = self.pretrained.pos_embed.detach().float().cpu()
        N = pos.shape[1] - 1
        D = pos.shape[-1]
        side = int(math.sqrt(N))
        # upstream names dim 2 of x "w": scale factors follow (dim2, dim3) = (rows, cols) of the patch grid
        sx, sy = float(ph + 0.1) / side, float(pw + 0.1) / side
        patch = F.interpolate(pos[:, 1:].reshape(1, side, side, D).permute(0, 3, 1, 2), scale_factor=(sx, sy),
                              mode="bicubic", antialias=False)
        assert patch.shape[-2:] == (ph, pw)
        patch = patch.permute(0, 2, 3, 1).reshape(ph * pw, D)
        return torch.cat([pos[0, :1], patch], dim=0).contiguous()

    # ---- reference surface ----------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected x of shape [B,3,H,W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise Dav2Error("DepthAnythingV2.forward needs a CUDA tensor on a B200: dav2_b200 has no CPU path")
        B, _, H, W = x.shape
        if H % 14 or W % 14:
            raise ValueError(f"H={H}, W={W} must be multiples of 14")
        x = x.detach().to(torch.float32).contiguous()
        self._ensure_engine(x.device)
        lib = _lib.load()
        ph, pw = H // 14, W // 14
        with torch.cuda.device(x.device):
            if (ph, pw) != (37, 37) and (ph, pw) not in self._pos_grids:
                tbl = self._pos_table(ph, pw)
                check(lib.dav2_set_pos_embed(self._handle, ph, pw, tbl.data_ptr()), "dav2_set_pos_embed")
                self._pos_grids.add((ph, pw))
            depth = torch.empty(B, H, W, dtype=torch.float32, device=x.device)
            check(lib.dav2_forward(self._handle, x.data_ptr(), B, H, W, depth.data_ptr(),
                                   _lib.current_stream_ptr(x.device)), "dav2_forward")
        return depth

    def capture_logits(self, on: bool = True) -> None:
        """Parity instrumentation: keep the pre-sigmoid logits of the following forwards (``debug_buffer("logits", ...)``)."""
        self._capture_logits = bool(on)
        if self._handle is not None:
            check(_lib.load().dav2_set_capture_logits(self._handle, 1 if on else 0), "dav2_set_capture_logits")

    def debug_buffer(self, name: str, dtype, shape) -> torch.Tensor:
        """Copy of an internal activation of the last forward (parity tests)."""
        dev = next(self.parameters()).device
        out = torch.empty(shape, dtype=dtype, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().dav2_debug_read(self._handle, name.encode(), out.data_ptr(), out.numel() * out.element_size(),
                                              _lib.current_stream_ptr(dev)), "dav2_debug_read")
        return out

    @torch.no_grad()
    def infer_image(self, raw_image: np.ndarray, input_size: int = 518) -> np.ndarray:
        """run.py:234.  The uint8 frame (3 B/px) is uploaded and pre-processed ON the GPU (bicubic resize +
        normalisation kernel), instead of building a 12 B/px float tensor with OpenCV on the host."""
        dev = next(self.parameters()).device
        h, w = raw_image.shape[:2]
        nh, nw = self.target_size(h, w, input_size)
        raw = torch.from_numpy(np.ascontiguousarray(raw_image)).to(dev)
        image = ops.preprocess_bgr_u8(raw, nh, nw)
        depth = self.forward(image)
        depth = ops.resize_depth(depth, h, w)[0]
        return depth.cpu().numpy()

    @staticmethod
    def target_size(h: int, w: int, input_size: int = 518):
        """Resize(keep_aspect_ratio, ensure_multiple_of=14, 'lower_bound') of upstream util/transform.py."""
        scale = max(input_size / h, input_size / w)

        def _mult(v):
            y = int(np.round(v / 14) * 14)
            return y if y >= input_size else int(np.ceil(v / 14) * 14)

        return _mult(scale * h), _mult(scale * w)

    @staticmethod
    def image2tensor(raw_image: np.ndarray, input_size: int = 518):
        """Upstream image2tensor: BGR->RGB, /255, lower-bound resize to a multiple of 14 (cubic), normalise."""
        import cv2

        h, w = raw_image.shape[:2]
        scale = max(input_size / h, input_size / w)

        def _mult(v):
            y = int(np.round(v / 14) * 14)
            return y if y >= input_size else int(np.ceil(v / 14) * 14)

        nh, nw = _mult(scale * h), _mult(scale * w)
        image = cv2.cvtColor(raw_image, cv2.COLOR_BGR2RGB) / 255.0
        image = cv2.resize(image, (nw, nh), interpolation=cv2.INTER_CUBIC)
        image = (image - np.asarray(_MEAN)) / np.asarray(_STD)
        image = np.ascontiguousarray(image.transpose(2, 0, 1)).astype(np.float32)
        return torch.from_numpy(image).unsqueeze(0), (h, w)
