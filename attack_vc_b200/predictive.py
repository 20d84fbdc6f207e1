"""Host-side mirror of the reference's VSMask ``PredictiveModel`` (models/predictive_model.py:53-110) on
top of the C-ABI (``avc_pm_*`` in include/avc_b200.h).  PyTorch supplies device memory and the stream
only; there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import WeightView
from .engine import AvcError


def _views(tensors: Dict[str, Tensor]):
    keep, views = [], (WeightView * len(tensors))()
    for i, (k, t) in enumerate(tensors.items()):
        keep.append(k.encode())
        views[i].name = keep[-1]
        views[i].data = t.data_ptr()
        views[i].ndim = t.dim()
        for j, s in enumerate(t.shape):
            views[i].shape[j] = int(s)
    return views, keep


class PredictiveEngine:
    """One ``avc_pm_handle`` bound to a PredictiveModel's weights on one CUDA device.

    ``model`` is the reference's ``PredictiveModel`` (or anything with the same ``state_dict()``), or the
    state dict itself.  ``forward(x)`` follows ``model.training`` semantics through the ``training`` flag:
    False = running statistics (vsmask.py:30), True = batch statistics (train_predictive.py:64)."""

    def __init__(self, model, device: Optional[torch.device] = None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise AvcError("attack_vc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        sd = model if isinstance(model, dict) else model.state_dict()
        sd = {k: v.detach() for k, v in sd.items() if v.dtype.is_floating_point}
        if device is None:
            device = next(iter(sd.values())).device
        device = torch.device(device)
        if device.type != "cuda":
            raise AvcError("attack_vc_b200 needs the model on a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self.shapes = {k: tuple(v.shape) for k, v in sd.items()}
        h = C.c_void_p()
        rc = self._lib.avc_pm_create(C.byref(h), self.device.index)
        if rc != 0:
            raise AvcError(f"avc_pm_create failed ({rc}): {self._lib.avc_pm_last_error(None).decode()}")
        self._h = h
        dev_sd = {k: v.to(device=self.device, dtype=torch.float32).contiguous() for k, v in sd.items()}
        views, keep = _views(dev_sd)
        torch.cuda.synchronize(self.device)
        self._check(self._lib.avc_pm_load_weights(self._h, views, len(dev_sd)))
        del keep

    def close(self):
        if getattr(self, "_h", None):
            self._lib.avc_pm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.avc_pm_last_error(self._h).decode()
            if rc == -1:
                raise ValueError(f"libavc_b200: {msg}")
            raise AvcError(f"libavc_b200 error {rc}: {msg}")

    def _x(self, x: Tensor) -> Tensor:
        if not isinstance(x, Tensor) or x.device != self.device:
            raise AvcError(f"x must be a tensor on {self.device} (no CPU fallback)")
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"x must be float32 [B, 1, F, T] (got {x.dtype} {tuple(x.shape)})")
        return x.contiguous()

    @staticmethod
    def out_shape(F: int, T: int) -> Tuple[int, int]:
        lib = _lib.load()
        a, b = C.c_int32(), C.c_int32()
        if lib.avc_pm_out_shape(F, T, C.byref(a), C.byref(b)) != 0:
            raise ValueError(f"input {F}x{T} is too small for the model")
        return int(a.value), int(b.value)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.avc_pm_kernel_launches(self._h))

    def forward(self, x: Tensor, training: bool = False) -> Tensor:
        x = self._x(x)
        B, _, F, T = x.shape
        Fo, To = self.out_shape(F, T)
        with torch.cuda.device(self.device):
            out = torch.empty(B, 1, Fo, To, device=self.device, dtype=torch.float32)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.avc_pm_forward(self._h, x.data_ptr(), out.data_ptr(), B, F, T, 1 if training else 0, st))
        return out

    __call__ = forward

    def train_step(self, x: Tensor, want_grad_x: bool = False):
        """model.train(); out = model(x); loss = out.square().mean(); loss.backward() -> dict with
        ``loss``, ``out``, ``grads`` (by state_dict key), ``new_stats`` and optionally ``grad_x``."""
        x = self._x(x)
        B, _, F, T = x.shape
        Fo, To = self.out_shape(F, T)
        with torch.cuda.device(self.device):
            out = torch.empty(B, 1, Fo, To, device=self.device, dtype=torch.float32)
            loss = torch.zeros(1, device=self.device, dtype=torch.float32)
            gx = torch.empty_like(x) if want_grad_x else None
            bufs = {k: torch.zeros(s, device=self.device, dtype=torch.float32) for k, s in self.shapes.items()}
            views, keep = _views(bufs)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.avc_pm_train_step(self._h, x.data_ptr(), B, F, T, out.data_ptr(), loss.data_ptr(),
                                                    gx.data_ptr() if gx is not None else None, views, len(bufs), st))
            del keep
        grads = {k: v for k, v in bufs.items() if "running" not in k}
        stats = {k: v for k, v in bufs.items() if "running" in k}
        return {"loss": loss[0], "out": out, "grads": grads, "new_stats": stats, "grad_x": gx}
