"""Drop-in for the reference's ``models/header_model.py`` (class name, constructor, ``header`` attribute,
``optimize`` / ``apply_header`` / ``save`` / ``load`` -- header_model.py:7-103), with ``optimize`` running on
libavc_b200.so (avc_header_optimize) instead of a Python loop of autograd steps.

What ``optimize`` needs to know about its callable arguments, which the reference leaves duck-typed:
  * ``speaker_encoder``: the AdaIN-VC model (an ``nn.Module`` with a ``speaker_encoder`` child, as
    ``attack.py`` loads it), or anything with an ``avc_model`` attribute naming one.  The kernels are the
    AdaIN-VC speaker encoder's; an arbitrary Python callable cannot be compiled and raises.
  * ``optimizer``: ``torch.optim.Adam([header.header], lr=...)`` as train_header.py:46 builds it.  Only lr is
    read; non-default betas / eps / weight_decay / amsgrad raise.  The optimizer's own moment state is not
    written back (the reference never reads it after ``optimize``).
There is no CPU path: tensors must live on the CUDA device of the model.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .engine import engine_for


class UniversalPerturbationHeader:
    def __init__(self, mel_bins: int = 80, time_length: int = 100, device: str = "cuda"):
        self.mel_bins, self.time_length, self.device = mel_bins, time_length, device
        self.header = torch.zeros((1, 1, mel_bins, time_length), device=device)      # header_model.py:22-23
        self.header.requires_grad = True

    @staticmethod
    def _resolve_model(speaker_encoder):
        m = getattr(speaker_encoder, "avc_model", speaker_encoder)
        if not (isinstance(m, torch.nn.Module) and hasattr(m, "speaker_encoder")):
            raise TypeError("speaker_encoder must be the AdaIN-VC model (or carry it as .avc_model); "
                            "arbitrary callables cannot run on the CUDA path")
        return m

    @staticmethod
    def _adam_lr(optimizer) -> float:
        if not isinstance(optimizer, torch.optim.Adam) or len(optimizer.param_groups) != 1:
            raise TypeError("optimizer must be torch.optim.Adam over [header.header] (train_header.py:46)")
        g = optimizer.param_groups[0]
        if tuple(g["betas"]) != (0.9, 0.999) or g["eps"] != 1e-8 or g["weight_decay"] != 0 or g.get("amsgrad", False):
            raise ValueError("only Adam's default betas / eps / weight_decay / amsgrad are implemented")
        return float(g["lr"])

    def optimize(self, source_mel: Tensor, target_mel: Tensor, speaker_encoder, optimizer,
                 num_iterations: int = 1000, epsilon: float = 0.1, lambda_param: float = 0.5) -> None:
        """header_model.py:25-68: num_iterations Adam steps on the shared header, clamped to +-epsilon."""
        eng = engine_for(self._resolve_model(speaker_encoder))
        lr = self._adam_lr(optimizer)
        # The fused loop starts Adam from m = v = 0, step 0 and does not write the moments back: a caller that continues with
        # an optimizer that already stepped (chunked optimisation) would silently get another trajectory than the reference,
        # whose moments and bias-correction step live in the optimizer.  train_header.py:77-81 calls optimize() once.
        if any(int(st.get("step", 0)) > 0 for st in optimizer.state.values()):
            raise RuntimeError("UniversalPerturbationHeader.optimize: the Adam optimizer has already stepped; the B200 path "
                               "restarts the moments on every call -- pass a fresh torch.optim.Adam([header.header], lr)")
        new = eng.header_optimize(source_mel, target_mel, num_iterations, epsilon, lambda_param, lr,
                                  header0=self.header.detach()[0, 0])
        with torch.no_grad():
            self.header.data = new                                                     # header_model.py:64-65

    def apply_header(self, source_mel: Tensor) -> Tensor:
        """header_model.py:70-95: add the header to the first min(T, time_length) frames, clamp to [-1, 1]."""
        n = min(source_mel.shape[3], self.time_length)
        out = source_mel.clone()
        out[:, :, :, :n] += self.header[:, :, :, :n]
        return torch.clamp(out, -1.0, 1.0)

    def save(self, path: str) -> None:
        torch.save(self.header, path)

    def load(self, path: str) -> None:
        self.header = torch.load(path, map_location=self.device)
        self.header.requires_grad = True
