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


def allreduce_callback(comm: Tensor, group=None):
    """The ``avc_allreduce_fn`` the library calls for sums that couple the ranks: all_reduce(SUM) of ``comm[:n]`` in
    place, ordered on the library's stream.  ``comm`` may be a CPU tensor under gloo (tests of this host logic)."""
    import torch.distributed as dist

    def fn(_ctx, _ptr, n, stream):
        try:
            view = comm[: int(n)]
            if comm.is_cuda:
                s = torch.cuda.ExternalStream(int(stream), device=comm.device) if stream else torch.cuda.default_stream(comm.device)
                with torch.cuda.stream(s):
                    dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
            else:
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:      # an exception must not unwind through the C frames
            import sys
            print(f"attack_vc_b200: all-reduce callback failed: {e!r}", file=sys.stderr)
            return 1
    return fn


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
            for t in list(getattr(self, "_trainers", ())):      # a trainer's buffers live in this handle's pool: end it first
                t.close()
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

    # ---- data parallel (BASELINE config 5 "across 2/4/8 GPUs") ---------------------------------------
    def set_process_group(self, group=None, world_size: Optional[int] = None):
        """Couple this handle with the other ranks' handles: BatchNorm batch statistics, their backward sums and (in the
        trainer) the parameter gradients are summed over ``group`` with ``torch.distributed.all_reduce`` (NCCL over
        NVLink), so world_size ranks x B windows compute what one device computes on world_size*B windows.  The library
        calls back into :func:`allreduce_callback`'s closure; the buffer is a torch tensor owned by this object."""
        import torch.distributed as dist
        world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        if world <= 1:
            self._check(self._lib.avc_pm_set_allreduce(self._h, _lib.ALLREDUCE_FN(0), None, None, 0, 1))
            self._comm = self._cb = None
            return
        n = int(self._lib.avc_pm_param_count(self._h))
        self._comm = torch.zeros(n, device=self.device, dtype=torch.float32)
        self._cb = _lib.ALLREDUCE_FN(allreduce_callback(self._comm, group))      # kept alive: C holds the pointer
        self._check(self._lib.avc_pm_set_allreduce(self._h, self._cb, None, self._comm.data_ptr(), n, world))

    def export_state_dict(self, like: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
        """model.state_dict() of the handle's CURRENT weights (after trainer steps: train_predictive.py:137-146)."""
        with torch.cuda.device(self.device):
            bufs = {k: torch.empty(s, device=self.device, dtype=torch.float32) for k, s in self.shapes.items()}
            views, keep = _views(bufs)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.avc_pm_export_weights(self._h, views, len(bufs), st))
            del keep
        return bufs

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

    def train_step(self, x: Tensor, want_grad_x: bool = False, reuse_buffers: bool = False):
        """model.train(); out = model(x); loss = out.square().mean(); loss.backward() -> dict with
        ``loss``, ``out``, ``grads`` (by state_dict key), ``new_stats`` and optionally ``grad_x``.
        ``reuse_buffers``: the gradients are views of ONE persistent flat tensor (``flat``: a data-parallel caller
        all-reduces it in a single collective) and no per-call allocation is made for them."""
        x = self._x(x)
        B, _, F, T = x.shape
        Fo, To = self.out_shape(F, T)
        with torch.cuda.device(self.device):
            out = torch.empty(B, 1, Fo, To, device=self.device, dtype=torch.float32)
            loss = torch.zeros(1, device=self.device, dtype=torch.float32)
            gx = torch.empty_like(x) if want_grad_x else None
            flat = None
            if reuse_buffers:
                if getattr(self, "_gbufs", None) is None:
                    names = [k for k in self.shapes if "running" not in k] + [k for k in self.shapes if "running" in k]
                    sizes = [max(1, int(torch.Size(self.shapes[k]).numel())) for k in names]
                    self._gflat = torch.zeros(sum(sizes), device=self.device, dtype=torch.float32)
                    self._gparams = sum(n for k, n in zip(names, sizes) if "running" not in k)
                    self._gbufs, o = {}, 0
                    for k, n in zip(names, sizes):
                        self._gbufs[k] = self._gflat[o:o + n].view(self.shapes[k]); o += n
                    self._gviews = _views(self._gbufs)
                bufs, (views, keep) = self._gbufs, self._gviews
                flat = self._gflat[: self._gparams]
            else:
                bufs = {k: torch.zeros(s, device=self.device, dtype=torch.float32) for k, s in self.shapes.items()}
                views, keep = _views(bufs)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            self._check(self._lib.avc_pm_train_step(self._h, x.data_ptr(), B, F, T, out.data_ptr(), loss.data_ptr(),
                                                    gx.data_ptr() if gx is not None else None, views, len(bufs), st))
            del keep
        grads = {k: v for k, v in bufs.items() if "running" not in k}
        stats = {k: v for k, v in bufs.items() if "running" in k}
        return {"loss": loss[0], "out": out, "grads": grads, "new_stats": stats, "grad_x": gx, "flat": flat}


class PredictiveTrainer:
    """The loop body of the reference's ``train_predictive_model`` (train_predictive.py:92-127) as one call per batch.

    ``pm`` is a :class:`PredictiveEngine` (its weights, BatchNorm running statistics and the Adam moments live on the
    device and are updated in place), ``speaker`` an :class:`attack_vc_b200.Engine` whose AdaIN-VC SpeakerEncoder plays
    ``speaker_encoder``.  Hyper-parameters carry the reference's argparse names and defaults (:173-184).  The two
    repairs the reference needs before the loop can run at all (crop of the [95,63] prediction to the 80 mel rows, the
    constraint applied on the mel axis) are documented in include/avc_b200.h and oracle/vsmask_train_oracle.py."""

    def __init__(self, pm: PredictiveEngine, speaker, batch_size: int, n_mels: int = 80, window_size: int = 100,
                 future_steps: int = 10, epsilon1: float = 0.1, epsilon2: float = 0.05, epsilon3: float = 0.08,
                 lambda_param: float = 0.5, betas=(0.9, 0.999), adam_eps: float = 1e-8, inv_norm: float = 0.0):
        if pm.device != speaker.device:
            raise AvcError(f"predictive model on {pm.device}, speaker encoder on {speaker.device}")
        self.pm, self.speaker, self._lib = pm, speaker, pm._lib
        self.shape = (int(batch_size), 1, int(n_mels), int(window_size))
        a = _lib.PmTrainerArgs(B=batch_size, F=n_mels, T=window_size, future_steps=future_steps, eps1=epsilon1, eps2=epsilon2,
                               eps3=epsilon3, lam=lambda_param, beta1=betas[0], beta2=betas[1], adam_eps=adam_eps,
                               inv_norm=float(inv_norm))
        t = C.c_void_p()
        with torch.cuda.device(pm.device):
            st = C.c_void_p(torch.cuda.current_stream(pm.device).cuda_stream)
            pm._check(self._lib.avc_pm_trainer_begin(pm._h, speaker._h, C.byref(a), st, C.byref(t)))
        self._t = t
        speaker._sessions += 1
        if not hasattr(pm, "_trainers"):
            pm._trainers = []
        pm._trainers.append(self)
        self._loss = torch.zeros(1, device=pm.device, dtype=torch.float32)

    def step(self, source_mels: Tensor, target_mels: Tensor, lr: float = 1e-3) -> Tensor:
        """One optimiser step; returns the loss as a 0-d device tensor (no host synchronisation on the value)."""
        for nm, x in (("source_mels", source_mels), ("target_mels", target_mels)):
            if not isinstance(x, Tensor) or x.device != self.pm.device:
                raise AvcError(f"{nm} must be a tensor on {self.pm.device} (no CPU fallback)")
            if x.dtype != torch.float32 or tuple(x.shape) != self.shape:
                raise ValueError(f"{nm} must be float32 {self.shape} (got {x.dtype} {tuple(x.shape)})")
        s, t = source_mels.contiguous(), target_mels.contiguous()
        with torch.cuda.device(self.pm.device):
            st = C.c_void_p(torch.cuda.current_stream(self.pm.device).cuda_stream)
            loss = torch.empty(1, device=self.pm.device, dtype=torch.float32)
            self.pm._check(self._lib.avc_pm_trainer_step(self._t, s.data_ptr(), t.data_ptr(), float(lr), loss.data_ptr(), st))
        return loss[0]

    def grads(self) -> Dict[str, Tensor]:
        """d loss / d parameter of the last step, by state_dict key (after the gradient all-reduce when sharded)."""
        with torch.cuda.device(self.pm.device):
            bufs = {k: torch.empty(s, device=self.pm.device, dtype=torch.float32) for k, s in self.pm.shapes.items() if "running" not in k}
            views, keep = _views(bufs)
            st = C.c_void_p(torch.cuda.current_stream(self.pm.device).cuda_stream)
            self.pm._check(self._lib.avc_pm_trainer_grads(self._t, views, len(bufs), st))
            del keep
        return bufs

    def state_dict(self) -> Dict[str, Tensor]:
        return self.pm.export_state_dict()

    def close(self):
        if getattr(self, "_t", None):
            self._lib.avc_pm_trainer_end(self._t)
            self._t = None
            self.speaker._sessions -= 1
            if self in getattr(self.pm, "_trainers", ()):
                self.pm._trainers.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
