"""Operator-level host wrappers over the C ABI (torch tensors in, torch tensors out).

torch is used for device memory and streams only; every computation below is a hand-written
sm_100a kernel inside libdav2_b200.so.  All tensors must live on the GPU."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import check, current_stream_ptr, require_cuda


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _fmt(*tensors) -> int:
    """fp16 / bf16 operand format code from the tensors' (common) dtype."""
    dt = tensors[0].dtype
    if any(t is not None and t.dtype != dt for t in tensors):
        raise _lib.Dav2Error("all 16-bit operands of one call must share a dtype")
    if dt == torch.float16:
        return _lib.FMT_F16
    if dt == torch.bfloat16:
        return _lib.FMT_BF16
    raise _lib.Dav2Error(f"16-bit operand must be torch.float16 or torch.bfloat16, got {dt}")


def linear_h16(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, act: int = 0) -> torch.Tensor:
    """act(a[M,K] @ w[N,K]^T + bias) -> [M,N] in the operands' 16-bit dtype   (act: 0 none, 1 GELU(erf), 2 ReLU)."""
    require_cuda(a, "a"); require_cuda(w, "w")
    fmt = _fmt(a, w)
    assert a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=a.dtype, device=a.device)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    check(_lib.load().dav2_linear_h16(a.data_ptr(), w.data_ptr(), _ptr(bias), out.data_ptr(), M, N, K, act, fmt,
                                      current_stream_ptr(a.device)), "dav2_linear_h16")
    return out


def linear_resid_(x: torch.Tensor, a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """x[M,N] (fp32, in place) += gamma * (a @ w^T + bias)."""
    require_cuda(x, "x"); require_cuda(a, "a"); require_cuda(w, "w")
    fmt = _fmt(a, w)
    assert x.dtype == torch.float32
    M, K = a.shape
    N = w.shape[0]
    assert x.shape == (M, N)
    check(_lib.load().dav2_linear_resid(a.data_ptr(), w.data_ptr(), _ptr(bias), gamma.data_ptr(), x.data_ptr(), M, N, K, fmt,
                                        current_stream_ptr(a.device)), "dav2_linear_resid")
    return x


def pack_conv3x3_weight(w: torch.Tensor, dtype=None) -> torch.Tensor:
    """Conv2d weight [Cout,Cin,3,3] -> 16-bit [Cout, 9*Cpad], tap-major, Cpad=ceil64(Cin)."""
    dtype = dtype or w.dtype
    Cout, Cin = w.shape[:2]
    Cpad = (Cin + 63) // 64 * 64
    p = torch.zeros(Cout, 9, Cpad, dtype=dtype, device=w.device)
    p[:, :, :Cin] = w.permute(0, 2, 3, 1).reshape(Cout, 9, Cin).to(dtype)
    return p.reshape(Cout, 9 * Cpad).contiguous()


def conv3x3_h16(x, wp, bias=None, add1=None, add2=None, act: int = 0, want_relu: bool = False):
    """3x3/pad1 conv on NHWC fp16/bf16: x [B,H,W,Cin], wp from pack_conv3x3_weight -> out [B,H,W,Cout] (, relu(out))."""
    require_cuda(x, "x"); require_cuda(wp, "wp")
    fmt = _fmt(x, wp, add1, add2)
    B, H, W, Cin = x.shape
    Cout = wp.shape[0]
    out = torch.empty(B, H, W, Cout, dtype=x.dtype, device=x.device)
    out_relu = torch.empty_like(out) if want_relu else None
    check(_lib.load().dav2_conv3x3_h16(x.data_ptr(), wp.data_ptr(), _ptr(bias), _ptr(add1), _ptr(add2), out.data_ptr(),
                                       _ptr(out_relu), B, H, W, Cin, Cout, act, fmt, current_stream_ptr(x.device)),
          "dav2_conv3x3_h16")
    return (out, out_relu) if want_relu else out


def attention_h16(qkv: torch.Tensor, B: int, N: int, D: int) -> torch.Tensor:
    """qkv fp16/bf16 [B*N, 3D] (q pre-scaled) -> softmax(q k^T) v, same dtype [B*N, D]; heads of 64."""
    require_cuda(qkv, "qkv")
    fmt = _fmt(qkv)
    assert qkv.shape == (B * N, 3 * D)
    out = torch.empty(B * N, D, dtype=qkv.dtype, device=qkv.device)
    check(_lib.load().dav2_attention_h16(qkv.data_ptr(), out.data_ptr(), B, N, D, fmt, current_stream_ptr(qkv.device)),
          "dav2_attention_h16")
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6, out_dtype=torch.float16) -> torch.Tensor:
    require_cuda(x, "x")
    assert x.dtype == torch.float32
    rows, D = x.shape
    out = torch.empty(rows, D, dtype=out_dtype, device=x.device)
    check(_lib.load().dav2_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), rows, D, eps, _fmt(out),
                                     current_stream_ptr(x.device)), "dav2_layernorm")
    return out


def bilinear_nhwc_h16(x: torch.Tensor, Ho: int, Wo: int) -> torch.Tensor:
    require_cuda(x, "x")
    B, Hi, Wi, Cc = x.shape
    out = torch.empty(B, Ho, Wo, Cc, dtype=x.dtype, device=x.device)
    check(_lib.load().dav2_bilinear_nhwc_h16(x.data_ptr(), out.data_ptr(), B, Hi, Wi, Ho, Wo, Cc, _fmt(x),
                                             current_stream_ptr(x.device)), "dav2_bilinear_nhwc_h16")
    return out


def resize_depth(depth: torch.Tensor, Ho: int, Wo: int) -> torch.Tensor:
    """F.interpolate(depth[:,None], (Ho,Wo), mode='bilinear', align_corners=True)[:,0] for fp32 [B,H,W]."""
    require_cuda(depth, "depth")
    assert depth.dtype == torch.float32 and depth.dim() == 3
    B, Hi, Wi = depth.shape
    out = torch.empty(B, Ho, Wo, dtype=torch.float32, device=depth.device)
    check(_lib.load().dav2_resize_depth(depth.data_ptr(), B, Hi, Wi, out.data_ptr(), Ho, Wo,
                                        current_stream_ptr(depth.device)), "dav2_resize_depth")
    return out


def backproject(depth: torch.Tensor, K4, T12=None, depth_scale: float = 1.0, depth_trunc: float = math.inf,
                want_valid: bool = True, want_counts: bool = True, out_xyz: torch.Tensor | None = None):
    """depth fp32 [B,H,W] -> (xyz fp32 [B,H*W,3], valid u8 [B,H*W] | None, counts i32 [B] | None).

    K4: (fx,fy,cx,cy) tuple, or fp64 tensor [4] / [B,4]; T12: fp64 tensor [B,12] of row-major [R|t] or None."""
    require_cuda(depth, "depth")
    assert depth.dtype == torch.float32 and depth.dim() == 3
    B, H, W = depth.shape
    dev = depth.device
    if not torch.is_tensor(K4):
        K4 = torch.tensor([float(v) for v in K4], dtype=torch.float64)
    K4 = K4.to(device=dev, dtype=torch.float64).contiguous()
    k_per_frame = 1 if K4.dim() == 2 else 0
    if k_per_frame:
        assert K4.shape == (B, 4)
    if T12 is not None:
        T12 = T12.to(device=dev, dtype=torch.float64).contiguous()
        assert T12.shape == (B, 12)
    xyz = out_xyz if out_xyz is not None else torch.empty(B, H * W, 3, dtype=torch.float32, device=dev)
    valid = torch.empty(B, H * W, dtype=torch.uint8, device=dev) if want_valid else None
    counts = torch.empty(B, dtype=torch.int32, device=dev) if want_counts else None
    check(_lib.load().dav2_backproject(depth.data_ptr(), B, H, W, K4.data_ptr(), k_per_frame, _ptr(T12),
                                       float(depth_scale), float(depth_trunc), xyz.data_ptr(), _ptr(valid), _ptr(counts),
                                       current_stream_ptr(dev)), "dav2_backproject")
    return xyz, valid, counts


def backproject_gather(depth: torch.Tensor, K4, T12, dst_xyz, dst_valid=None, dst_counts=None, frame_offset: int = 0,
                       depth_scale: float = 1.0, depth_trunc: float = math.inf, gt: torch.Tensor | None = None,
                       min_depth: float = 1e-6, max_depth: float = 20.0, per_frame: bool = False):
    """Fused back-projection + all-gather store (dav2_backproject_gather): results for depth [B,H,W] are written at frame
    index ``frame_offset + b`` of EVERY destination.  dst_* are lists of device pointers (ints; peer-mapped buffers from
    ``sharding.CloudGather``) or of tensors ([F,H*W,3] fp32 / [F,H*W] u8 / [F] i32).

    With ``gt`` (fp32, B*H*W elements) the SAME pass also accumulates the test_step metric partial sums of (depth, gt)
    (dav2_backproject_metrics) and returns them (fp64 [8] or [B,8]); without, returns None."""
    require_cuda(depth, "depth")
    assert depth.dtype == torch.float32 and depth.dim() == 3
    B, H, W = depth.shape
    dev = depth.device
    if not torch.is_tensor(K4):
        K4 = torch.tensor([float(v) for v in K4], dtype=torch.float64)
    K4 = K4.to(device=dev, dtype=torch.float64).contiguous()
    k_per_frame = 1 if K4.dim() == 2 else 0
    if T12 is not None:
        T12 = T12.to(device=dev, dtype=torch.float64).contiguous()
        assert T12.shape == (B, 12)
    n = len(dst_xyz)
    assert 1 <= n <= 8, "1..8 destinations"

    def ptr_array(items):
        if items is None:
            return None
        assert len(items) == n
        return (C.c_void_p * n)(*[(t.data_ptr() if torch.is_tensor(t) else int(t)) for t in items])

    lib = _lib.load()
    if gt is None:
        check(lib.dav2_backproject_gather(depth.data_ptr(), B, H, W, K4.data_ptr(), k_per_frame, _ptr(T12),
                                          float(depth_scale), float(depth_trunc), ptr_array(dst_xyz), ptr_array(dst_valid),
                                          ptr_array(dst_counts), n, int(frame_offset), current_stream_ptr(dev)),
              "dav2_backproject_gather")
        return None
    require_cuda(gt, "gt")
    assert gt.dtype == torch.float32 and gt.numel() == depth.numel()
    part = torch.empty((B, 8) if per_frame else (8,), dtype=torch.float64, device=dev)
    check(lib.dav2_backproject_metrics(depth.data_ptr(), gt.data_ptr(), B, H, W, K4.data_ptr(), k_per_frame, _ptr(T12),
                                       float(depth_scale), float(depth_trunc), ptr_array(dst_xyz), ptr_array(dst_valid),
                                       ptr_array(dst_counts), n, int(frame_offset), float(min_depth), float(max_depth),
                                       1 if per_frame else 0, part.data_ptr(), current_stream_ptr(dev)),
          "dav2_backproject_metrics")
    return part


def backproject_metrics(depth: torch.Tensor, gt: torch.Tensor, K4, T12=None, min_depth: float = 1e-6, max_depth: float = 20.0,
                        depth_scale: float = 1.0, depth_trunc: float = math.inf, per_frame: bool = False,
                        out_xyz: torch.Tensor | None = None):
    """One pass over depth [B,H,W]: (xyz [B,H*W,3], valid [B,H*W], counts [B], metric partial sums fp64 [8] | [B,8]) --
    ``backproject`` + ``evaluation.metric_partials`` (test_step mask min_depth <= gt <= max_depth) without re-reading depth."""
    B, H, W = depth.shape
    dev = depth.device
    xyz = out_xyz if out_xyz is not None else torch.empty(B, H * W, 3, dtype=torch.float32, device=dev)
    valid = torch.empty(B, H * W, dtype=torch.uint8, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    part = backproject_gather(depth, K4, T12, [xyz], [valid], [counts], 0, depth_scale, depth_trunc, gt=gt.contiguous(),
                              min_depth=min_depth, max_depth=max_depth, per_frame=per_frame)
    return xyz, valid, counts, part


def voxel_downsample(xyz: torch.Tensor, voxel_size: float, rgb: torch.Tensor | None = None,
                     valid: torch.Tensor | None = None):
    """Open3D-style voxel_down_sample: xyz fp32 [n,3] (+ rgb fp32 [n,3], valid u8 [n]) -> (xyz [m,3], rgb [m,3] | None),
    one mean point per occupied voxel, ordered by voxel index.  One host sync (the voxel count)."""
    require_cuda(xyz, "xyz")
    assert xyz.dtype == torch.float32 and xyz.dim() == 2 and xyz.shape[1] == 3
    if not voxel_size > 0.0:
        raise ValueError("voxel_size <= 0.")  # Open3D's message
    xyz = xyz.contiguous()
    n, dev = xyz.shape[0], xyz.device
    if rgb is not None:
        rgb = rgb.to(device=dev, dtype=torch.float32).contiguous()
        assert rgb.shape == xyz.shape
    if valid is not None:
        valid = valid.to(device=dev, dtype=torch.uint8).contiguous()
        assert valid.numel() == n
    out_xyz = torch.empty_like(xyz)
    out_rgb = torch.empty_like(xyz) if rgb is not None else None
    count = torch.empty(1, dtype=torch.int64, device=dev)
    check(_lib.load().dav2_voxel_downsample(xyz.data_ptr(), _ptr(rgb), _ptr(valid), n, float(voxel_size), out_xyz.data_ptr(),
                                            _ptr(out_rgb), count.data_ptr(), current_stream_ptr(dev)), "dav2_voxel_downsample")
    m = int(count.item())
    if m < 0:
        raise RuntimeError("voxel_size is too small.")  # Open3D's message
    return out_xyz[:m], (out_rgb[:m] if out_rgb is not None else None)


def depth_metric_partials(pred: torch.Tensor, gt: torch.Tensor, lo: float, hi: float, variant: int,
                          per_frame: bool) -> torch.Tensor:
    """fp64 partial sums [B,8] or [8] (see include/dav2_b200.h) for pred/gt fp32 [B, ...]."""
    require_cuda(pred, "pred"); require_cuda(gt, "gt")
    assert pred.dtype == torch.float32 and gt.dtype == torch.float32 and pred.shape == gt.shape
    B = pred.shape[0] if pred.dim() > 1 else 1
    HW = pred.numel() // B
    out = torch.empty((B, 8) if per_frame else (8,), dtype=torch.float64, device=pred.device)
    check(_lib.load().dav2_depth_metrics(pred.data_ptr(), gt.data_ptr(), B, HW, float(lo), float(hi), int(variant),
                                         1 if per_frame else 0, out.data_ptr(), current_stream_ptr(pred.device)),
          "dav2_depth_metrics")
    return out


def transform_points_(xyz: torch.Tensor, T) -> torch.Tensor:
    """In place p <- R p + t for xyz fp32 [n,3] on the GPU; T = 4x4 (or 3x4 / 12) row-major [R|t], evaluated in fp64."""
    require_cuda(xyz, "xyz")
    assert xyz.dtype == torch.float32 and xyz.dim() == 2 and xyz.shape[1] == 3
    T12 = torch.as_tensor(T, dtype=torch.float64).reshape(-1)[:12].to(xyz.device).contiguous()
    check(_lib.load().dav2_transform_points(xyz.data_ptr(), xyz.shape[0], T12.data_ptr(), current_stream_ptr(xyz.device)),
          "dav2_transform_points")
    return xyz


def compose_poses(rel: torch.Tensor, init7: torch.Tensor | None = None, want_T12: bool = False):
    require_cuda(rel, "rel")
    assert rel.dtype == torch.float32 and rel.dim() == 2 and rel.shape[1] == 7
    N = rel.shape[0]
    abs7 = torch.empty(N + 1, 7, dtype=torch.float32, device=rel.device)
    T12 = torch.empty(N + 1, 12, dtype=torch.float64, device=rel.device) if want_T12 else None
    if init7 is not None:
        init7 = init7.to(device=rel.device, dtype=torch.float32).reshape(-1).contiguous()
        assert init7.numel() == 7
    check(_lib.load().dav2_compose_poses(rel.data_ptr(), _ptr(init7), N, abs7.data_ptr(), _ptr(T12),
                                         current_stream_ptr(rel.device)), "dav2_compose_poses")
    return (abs7, T12) if want_T12 else abs7


def preprocess_bgr_u8(img_u8: torch.Tensor, nh: int, nw: int) -> torch.Tensor:
    """BGR uint8 [H,W,3] or [B,H,W,3] (device) -> normalised RGB fp32 [B,3,nh,nw] with OpenCV-compatible bicubic resize
    (upstream image2tensor; one launch for the whole batch)."""
    require_cuda(img_u8, "img")
    if img_u8.dim() == 3:
        img_u8 = img_u8[None]
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[3] == 3
    B, H, W = img_u8.shape[:3]
    out = torch.empty(B, 3, nh, nw, dtype=torch.float32, device=img_u8.device)
    check(_lib.load().dav2_preprocess_bgr_u8_batch(img_u8.data_ptr(), B, H, W, out.data_ptr(), nh, nw,
                                                   current_stream_ptr(img_u8.device)), "dav2_preprocess_bgr_u8_batch")
    return out


def resize_aa(x: torch.Tensor, Ho: int, Wo: int, div_in: float | None = None) -> torch.Tensor:
    """torch's anti-aliased bicubic ``Resize((Ho, Wo), BICUBIC, antialias=True)`` after ToTensor, on the GPU (dav2_resize_aa).

    uint8 [B,H,W,3] RGB -> /255 -> resize -> ImageNet normalisation -> fp32 [B,3,Ho,Wo]   (SimCol transform_input)
    uint16 / float32 [B,H,W] -> / div_in (default 65535 for uint16, 1 for float) -> resize -> fp32 [B,1,Ho,Wo]
    (SimCol transform_output)."""
    require_cuda(x, "x")
    if x.dtype == torch.uint8:
        assert x.dim() == 4 and x.shape[3] == 3, "uint8 input must be [B,H,W,3] RGB"
        mode, C, s = 0, 3, (255.0 if div_in is None else div_in)
    elif x.dtype == torch.uint16:
        assert x.dim() == 3, "uint16 input must be [B,H,W]"
        mode, C, s = 1, 1, (65535.0 if div_in is None else div_in)
    elif x.dtype == torch.float32:
        assert x.dim() == 3, "float32 input must be [B,H,W]"
        mode, C, s = 2, 1, (1.0 if div_in is None else div_in)
    else:
        raise _lib.Dav2Error(f"resize_aa: unsupported dtype {x.dtype}")
    B, H, W = x.shape[:3]
    out = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=x.device)
    check(_lib.load().dav2_resize_aa(mode, x.data_ptr(), B, H, W, out.data_ptr(), Ho, Wo, float(s), current_stream_ptr(x.device)),
          "dav2_resize_aa")
    return out
