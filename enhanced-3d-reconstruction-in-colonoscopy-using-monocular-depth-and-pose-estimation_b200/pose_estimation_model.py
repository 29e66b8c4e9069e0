"""Drop-in for ``PoseEstimationNet`` of the reference's ``pose_estimation_model.py:35-105``.

Same constructor (``in_channels``), attribute names (``backbone`` = torchvision ResNet-18 with an
``in_channels``-wide 7x7/2 stem and ``fc -> 256``; ``pose_head`` = the ReLU/Dropout/Linear stack) and
state-dict keys, so checkpoints of the reference load unchanged.  The torchvision / nn layers are
PARAMETER CONTAINERS; ``forward`` runs the network in libdav2_b200.so (tcgen05 GEMM / implicit-GEMM
kernels, BatchNorm folded into the convolutions at engine-build time, eval semantics only)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import Dav2Error, check


def _fold_bn(conv_w: torch.Tensor, bn) -> tuple:
    """conv (no bias) followed by eval-mode BatchNorm2d -> (weight, bias) of the equivalent conv."""
    scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    w = conv_w.detach().double() * scale.view(-1, 1, 1, 1)
    b = bn.bias.detach().double() - bn.running_mean.detach().double() * scale
    return w.float().cpu().contiguous(), b.float().cpu().contiguous()


class PoseEstimationNet(nn.Module):
    def __init__(self, in_channels: int = 8, precision: str = "fp16") -> None:
        super().__init__()
        if in_channels != 8:
            raise NotImplementedError("the engine packs 8-channel frame pairs (rgb1, d1, rgb2, d2)")
        from torchvision.models import resnet18

        self.backbone = resnet18(weights=None)
        self.backbone.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.backbone.fc = nn.Linear(self.backbone.fc.in_features, 256)
        self.pose_head = nn.Sequential(nn.ReLU(), nn.Dropout(0.3), nn.Linear(256, 128), nn.ReLU(), nn.Dropout(0.2),
                                       nn.Linear(128, 64), nn.ReLU(), nn.Dropout(0.1), nn.Linear(64, 7))
        self.precision = precision
        self._handle = None
        self._handle_device = None
        self._dirty = True

    def _apply(self, fn, *a, **k):
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._dirty = True
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def mark_weights_changed(self):
        self._dirty = True

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().dav2_pose_destroy(self._handle)
        except Exception:
            pass

    def _ensure_engine(self, device):
        # one engine on ONE device, one stream at a time (include/dav2_b200.h); an input on another GPU re-packs it there
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._handle is not None and not self._dirty and self._handle_device == device:
            return
        lib = _lib.load()
        if self._handle is not None:
            lib.dav2_pose_destroy(self._handle)
            self._handle = None
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.dav2_pose_create(C.byref(h), _lib.FMT_BF16 if self.precision == "bf16" else _lib.FMT_F16), "dav2_pose_create")
            self._handle = h

            def put(key, t):
                t = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
                check(lib.dav2_pose_set_weight(h, key.encode(), t.data_ptr(), (C.c_int64 * t.dim())(*t.shape), t.dim()),
                      f"dav2_pose_set_weight({key})")

            bb = self.backbone
            w, b = _fold_bn(bb.conv1.weight, bb.bn1)
            put("conv1.weight", w); put("conv1.bias", b)
            for L in (1, 2, 3, 4):
                layer = getattr(bb, f"layer{L}")
                for i, blk in enumerate(layer):
                    for j, (cv, bn) in enumerate(((blk.conv1, blk.bn1), (blk.conv2, blk.bn2)), start=1):
                        w, b = _fold_bn(cv.weight, bn)
                        put(f"layer{L}.{i}.conv{j}.weight", w); put(f"layer{L}.{i}.conv{j}.bias", b)
                    if blk.downsample is not None:
                        w, b = _fold_bn(blk.downsample[0].weight, blk.downsample[1])
                        put(f"layer{L}.{i}.downsample.weight", w); put(f"layer{L}.{i}.downsample.bias", b)
            put("fc.weight", bb.fc.weight); put("fc.bias", bb.fc.bias)
            for n, idx in enumerate((2, 5, 8)):
                put(f"head.{n}.weight", self.pose_head[idx].weight); put(f"head.{n}.bias", self.pose_head[idx].bias)
        self._handle_device = device
        self._dirty = False

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise Dav2Error("PoseEstimationNet engine implements eval semantics only: call .eval()")
        if x.dim() != 4 or x.shape[1] != 8:
            raise ValueError(f"expected [B,8,H,W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise Dav2Error("PoseEstimationNet.forward needs a CUDA tensor on a B200: dav2_b200 has no CPU path")
        x = x.detach().to(torch.float32).contiguous()
        self._ensure_engine(x.device)
        B, _, H, W = x.shape
        out = torch.empty(B, 7, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(_lib.load().dav2_pose_forward(self._handle, x.data_ptr(), B, H, W, out.data_ptr(),
                                                _lib.current_stream_ptr(x.device)), "dav2_pose_forward")
        return out


def stack_pairs(rgb: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] normalised RGB + [N,1,H,W] depth -> [N-1,8,H,W] = cat(rgb_i, d_i, rgb_{i+1}, d_{i+1})
    (data_processing/pose_estimation.py:229-243)."""
    f = torch.cat([rgb, depth], dim=1)
    return torch.cat([f[:-1], f[1:]], dim=1).contiguous()


class PoseEstimationModule:
    """Inference-side drop-in for the reference's ``PoseEstimationModule`` (pose_estimation_model.py:108-170 constructor /
    forward, :295-343 test hooks): ``test_step`` predicts the relative poses of a batch of stacked pairs, keeps them for the
    trajectory evaluation and returns the per-batch ``compute_pose_errors``; ``on_test_epoch_end`` scale-aligns and composes
    the whole trajectory (``evaluation.evaluate_trajectory``, pose chain on the GPU kernel).  No Lightning import; the
    training hooks are outside the path and raise."""

    KEYS = ("ate", "rte", "rote")

    def __init__(self, in_channels: int = 8, precision: str = "fp16", **hparams):
        from types import SimpleNamespace
        from . import evaluation
        self.hparams = SimpleNamespace(in_channels=in_channels, **hparams)
        self.model = PoseEstimationNet(in_channels=in_channels, precision=precision)
        self.metric = evaluation.RunningMeans(self.KEYS)
        self.current_trajectory_preds: list = []
        self.current_trajectory_gts: list = []
        self.trajectory_metrics: list = []
        self.logged: dict = {}

    @property
    def device(self) -> torch.device:
        return next(self.model.parameters()).device

    def to(self, device):
        self.model = self.model.to(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def eval(self):
        self.model.eval()
        return self

    def state_dict(self) -> dict:
        return {f"model.{k}": v for k, v in self.model.state_dict().items()}

    def load_state_dict(self, state_dict: dict, strict: bool = True):
        sd = {k[len("model."):]: v for k, v in state_dict.items() if k.startswith("model.")}
        extra = [k for k in state_dict if not k.startswith("model.")]
        if strict and extra:
            raise RuntimeError(f"unexpected keys in state_dict: {extra[:4]}")
        return self.model.load_state_dict(sd, strict=strict)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)

    __call__ = forward

    def log(self, name: str, value) -> None:
        self.logged[name] = float(value)

    def on_test_epoch_start(self) -> None:
        self.metric.reset()
        self.current_trajectory_preds, self.current_trajectory_gts, self.trajectory_metrics = [], [], []

    @torch.no_grad()
    def test_step(self, batch: dict, batch_idx: int = 0) -> dict:
        from . import evaluation
        pred = self(batch["input"])
        target = batch["target"]
        self.current_trajectory_preds.append(pred.detach().cpu())
        self.current_trajectory_gts.append(target.detach().cpu())
        metrics = evaluation.compute_pose_errors(pred.detach(), target.detach())
        self.metric.update(metrics)
        for k, v in metrics.items():
            self.log(f"Test/test_{k}", v)
        return metrics

    def on_test_epoch_end(self) -> dict:
        from . import evaluation
        # the reference stacks the per-batch tensors ([batches, B, 7], equal batch sizes required) and hands the 3-D stack
        # to evaluate_trajectory (:321-330); kept as is
        pred = torch.stack(self.current_trajectory_preds)
        gt = torch.stack(self.current_trajectory_gts)
        traj = evaluation.evaluate_trajectory(pred_rel_poses=pred, gt_rel_poses=gt, initial_pose=None)
        for k, v in traj.items():
            self.log(f"Test/trajectory_{k}", v)
        self.trajectory_metrics.append(traj)
        self.current_trajectory_preds, self.current_trajectory_gts = [], []
        final = self.metric.compute()
        for k, v in final.items():
            self.log(f"Test/test_{k}", v)
        self.metric.reset()
        return {"trajectory": traj, "mean": final}

    def training_step(self, *a, **k):
        raise NotImplementedError("dav2_b200 is the inference path; training stays with the reference")

    validation_step = configure_optimizers = pose_loss = training_step
