"""ctypes binding of libdav2_b200.so (include/dav2_b200.h).  Fails loudly when the library or a
CUDA device is missing -- there is deliberately no fallback path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# The product always loads the in-tree build.  (Profiling scripts that A/B another build of the same library assign
# ``_lib.LIB_PATH`` themselves before the first ``load()``; no environment variable can swap the library.)
LIB_PATH = os.path.join(_HERE, "libdav2_b200.so")

c_void_p, c_int, c_i64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class Dav2Config(C.Structure):
    _fields_ = [
        ("embed_dim", c_int), ("depth", c_int), ("num_heads", c_int), ("features", c_int),
        ("out_channels", c_int * 4), ("tap_layers", c_int * 4), ("max_depth", c_float), ("precision", c_int),
    ]


FMT_F16, FMT_BF16 = 0, 1  # tensor-core operand formats (== include/dav2_b200.h `precision` / `fmt`)
FMT_F32 = 2               # dav2_config.precision only: the fp32 validation engine


class Dav2Error(RuntimeError):
    pass


# name -> (restype, argtypes); exactly the declarations of include/dav2_b200.h
SIGNATURES = {
    "dav2_create": (c_int, [C.POINTER(c_void_p), C.POINTER(Dav2Config)]),
    "dav2_destroy": (None, [c_void_p]),
    "dav2_set_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, C.POINTER(c_i64), c_int]),
    "dav2_weights_complete": (c_int, [c_void_p]),
    "dav2_set_pos_embed": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "dav2_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dav2_set_capture_logits": (c_int, [c_void_p, c_int]),
    "dav2_debug_buffer": (c_int, [c_void_p, C.c_char_p, C.POINTER(c_void_p), C.POINTER(c_i64)]),
    "dav2_debug_read": (c_int, [c_void_p, C.c_char_p, c_void_p, c_i64, c_void_p]),
    "dav2_resize_depth": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "dav2_preprocess_bgr_u8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "dav2_preprocess_bgr_u8_batch": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "dav2_resize_aa": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p]),
    "dav2_backproject": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_float, c_float,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "dav2_backproject_gather": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_float, c_float,
                                        C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_i64, c_void_p]),
    "dav2_backproject_metrics": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_float, c_float,
                                         C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p), c_int, c_i64,
                                         c_float, c_float, c_int, c_void_p, c_void_p]),
    "dav2_peer_alloc": (c_int, [C.POINTER(c_void_p), c_i64]),
    "dav2_peer_free": (c_int, [c_void_p]),
    "dav2_peer_export": (c_int, [c_void_p, c_void_p]),
    "dav2_peer_open": (c_int, [c_void_p, C.POINTER(c_void_p)]),
    "dav2_peer_close": (c_int, [c_void_p]),
    "dav2_voxel_downsample": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, C.c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dav2_depth_metrics": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_float, c_float, c_int, c_int, c_void_p, c_void_p]),
    "dav2_transform_points": (c_int, [c_void_p, c_i64, c_void_p, c_void_p]),
    "dav2_compose_poses": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dav2_pose_create": (c_int, [C.POINTER(c_void_p), c_int]),
    "dav2_pose_destroy": (None, [c_void_p]),
    "dav2_pose_set_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, C.POINTER(c_i64), c_int]),
    "dav2_pose_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dav2_linear_h16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dav2_linear_resid": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dav2_conv3x3_h16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dav2_attention_h16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dav2_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_float, c_int, c_void_p]),
    "dav2_bilinear_nhwc_h16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dav2_profile_enable": (None, [c_int]),
    "dav2_profile_report": (c_int, [C.c_char_p, c_int]),
    "dav2_last_error": (C.c_char_p, []),
    "dav2_launch_count": (c_i64, []),
    "dav2_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Dav2Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). dav2_b200 has no CPU / PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dav2_last_error().decode(errors="replace")
        raise Dav2Error(f"{what or 'dav2 call'} failed (rc={rc}): {msg}")


def profile_enable(on: bool) -> None:
    load().dav2_profile_enable(1 if on else 0)


def profile_report() -> dict:
    import json

    buf = C.create_string_buffer(8192)
    check(load().dav2_profile_report(buf, len(buf)), "dav2_profile_report")
    return json.loads(buf.value.decode())


def launch_count() -> int:
    return int(load().dav2_launch_count())


def current_stream_ptr(device=None) -> int:
    import torch

    return int(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise Dav2Error(f"{name} must be a CUDA tensor: dav2_b200 has no CPU path (got device {t.device})")
    if not t.is_contiguous():
        raise Dav2Error(f"{name} must be contiguous")
    return t
