"""Inference-side drop-in for the reference's ``lightning_model.DepthAnythingV2Module``
(lightning_model.py:72-152 constructor, :154-168 ``_preprocess_batch``, :285-330 ``test_step`` and the
test-epoch hooks, :343-360 ``predict_step``) and for the ``trainer.test`` loop that drives it
(test_lightning.py:222-236).  No Lightning / torchmetrics import: the hot path needs neither.

What differs from the reference, on purpose:
  * ``test_step`` computes mask + metrics in ONE kernel pass over (pred, gt)
    (``evaluation.test_step_metrics``) instead of two boolean-mask compactions + four reductions, and
    it does not ``.item()`` anything: the running means of ``self.metric`` live on the device and are
    read once in ``on_test_epoch_end``.
  * training (``training_step`` / ``validation_step`` / ``configure_optimizers``) is outside the path
    (inference only) and raises.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Iterable, Optional

import torch

from . import evaluation
from .dpt import MODEL_CONFIGS, DepthAnythingV2  # MODEL_CONFIGS: lightning_model.py:80-107 (vitg is not built)

METRIC_KEYS = ("d1", "abs_rel", "rmse", "l1")


class DepthAnythingV2Module:
    """``DepthAnythingV2Module(encoder, min_depth, max_depth, ...)`` with the reference's test / predict
    surface.  ``pretrained_from``: a ``depth_anything_v2_metric_hypersim_{encoder}.pth``-style file whose
    ``"pretrained"`` keys are loaded with ``strict=False`` like lightning_model.py:130-140; ``None`` (default)
    looks for the reference's relative path and skips the load when the file is not there (there is no
    checkpoint in this image)."""

    model_configs = MODEL_CONFIGS

    def __init__(self, encoder: str = "vitl", min_depth: float = 1e-6, max_depth: float = 20.0,
                 pretrained_from: Optional[str] = None, precision: str = "fp16", **hparams):
        if encoder not in MODEL_CONFIGS:
            raise ValueError(f"encoder must be one of {sorted(MODEL_CONFIGS)} (vitg is not built), got {encoder!r}")
        self.hparams = SimpleNamespace(encoder=encoder, min_depth=min_depth, max_depth=max_depth, **hparams)
        self.model = DepthAnythingV2(**{**MODEL_CONFIGS[encoder], "max_depth": max_depth}, precision=precision)
        path = pretrained_from or f"./base_checkpoints/depth_anything_v2_metric_hypersim_{encoder}.pth"
        if os.path.exists(path):
            sd = torch.load(path, map_location="cpu")
            self.model.load_state_dict({k: v for k, v in sd.items() if "pretrained" in k}, strict=False)
        elif pretrained_from is not None:
            raise FileNotFoundError(pretrained_from)
        self.metric = evaluation.RunningMeans(METRIC_KEYS)
        self.logged: dict = {}

    # ---- nn.Module-like plumbing the test driver uses -------------------------------------------------
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
        """Lightning checkpoints prefix the network's keys with ``model.`` (test_lightning.py:114-130,
        run.py:134-144)."""
        sd = {k[len("model."):]: v for k, v in state_dict.items() if k.startswith("model.")}
        extra = [k for k in state_dict if not k.startswith("model.")]
        if strict and extra:
            raise RuntimeError(f"unexpected keys in state_dict: {extra[:4]}")
        return self.model.load_state_dict(sd, strict=strict)

    # ---- the hot path ------------------------------------------------------------------------------
    def _preprocess_batch(self, batch: dict) -> tuple:
        return batch["image"], batch["depth"]

    @torch.no_grad()
    def test_step(self, batch: dict, batch_idx: int = 0) -> dict:
        img, depth = self._preprocess_batch(batch)
        pred = self.model(img).unsqueeze(1)
        assert pred.shape == depth.shape, (pred.shape, depth.shape)
        metrics = evaluation.test_step_metrics(pred, depth, self.hparams.min_depth, self.hparams.max_depth)
        self.metric.update(metrics)
        return {k: metrics[k] for k in METRIC_KEYS}

    @torch.no_grad()
    def predict_step(self, batch: dict, batch_idx: int = 0) -> torch.Tensor:
        img, _ = self._preprocess_batch(batch)
        return self.model(img)

    def on_test_epoch_start(self) -> None:
        self.metric.reset()

    def on_test_epoch_end(self) -> dict:
        final = self.metric.compute()
        for k, v in final.items():
            self.log(f"Test/test_{k}", v)
        self.metric.reset()
        return final

    def log(self, name: str, value) -> None:
        self.logged[name] = float(value)

    def training_step(self, *a, **k):
        raise NotImplementedError("dav2_b200 is the inference path; training stays with the reference")

    validation_step = configure_optimizers = training_step


def _to_device(batch: dict, device) -> dict:
    return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}


def test(module: DepthAnythingV2Module, dataloader: Iterable[dict], callbacks: Iterable = ()) -> dict:
    """What ``trainer.test(model, datamodule)`` does for this module (test_lightning.py:222-236): epoch-start
    hook, ``test_step`` per batch with the batch moved to the module's device, every callback's
    ``on_test_batch_end(outputs, batch)``, epoch-end hook.  Returns the logged ``Test/test_*`` means."""
    callbacks = list(callbacks)
    module.eval()
    module.on_test_epoch_start()
    dev = module.device
    for i, batch in enumerate(dataloader):
        batch = _to_device(batch, dev)
        out = module.test_step(batch, i)
        for cb in callbacks:
            cb.on_test_batch_end(out, batch)
    module.on_test_epoch_end()
    return dict(module.logged)


test.__test__ = False


def predict(module: DepthAnythingV2Module, dataloader: Iterable[dict]) -> list:
    """``trainer.predict``: the list of per-batch depth tensors [B,H,W]."""
    module.eval()
    dev = module.device
    return [module.predict_step(_to_device(b, dev), i) for i, b in enumerate(dataloader)]
