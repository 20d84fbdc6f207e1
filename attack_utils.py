"""Drop-in replacement for the reference's ``attack_utils.py`` (same names, signatures, argument
meaning and return value -- attack_utils.py:7-14, 51-53, 89-96), running on libavc_b200.so.

``attack.py`` imports ``e2e_attack, emb_attack, fb_attack`` from this module (attack.py:6) and calls
them at attack.py:60-64; nothing else changes for the CLI.  Differences from the reference, all
documented in DESIGN.md: the returned tensor is detached (callers only use ``.data``,
attack.py:69-70), ``model.*.grad`` is not populated (an unread side effect of ``loss.backward()``),
and CPU tensors raise instead of running slowly -- there is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor
from tqdm import tqdm

from attack_vc_b200 import engine_for

_BAR_CHUNKS = 20


def _run(kind: str, model: nn.Module, vc_src, vc_tgt: Tensor, adv_tgt: Tensor, eps: float, n_iters: int) -> Tensor:
    eng = engine_for(model)
    # same RNG consumption as the reference: one N(0,1) draw shaped like vc_tgt on its device
    ptb = torch.zeros_like(vc_tgt).normal_(0, 1)
    n = int(n_iters)
    pbar = tqdm(total=n)                     # the reference shows tqdm.trange(n_iters), attack_utils.py:33,71,115
    if n < 2 * _BAR_CHUNKS:
        out = eng.attack(kind, vc_tgt, adv_tgt, eps, n, vc_src=vc_src, w0=ptb)
        pbar.update(n)
    else:
        # a bar that moves: the iterations are enqueued in ~20 chunks and the stream is drained after each one
        # (20 synchronisations per attack: ~0.1 ms on a 600 ms run of 1500 iterations)
        ses = eng.begin(kind, vc_tgt, adv_tgt, eps, n, vc_src=vc_src, w0=ptb)
        try:
            done, chunk = 0, -(-n // _BAR_CHUNKS)
            while done < n:
                k = min(chunk, n - done)
                ses.step(k)
                torch.cuda.current_stream(vc_tgt.device).synchronize()
                done += k
                pbar.update(k)
        finally:
            out, _ = ses.end()
    pbar.close()
    return out


def e2e_attack(model: nn.Module, vc_src: Tensor, vc_tgt: Tensor, adv_tgt: Tensor, eps: float, n_iters) -> Tensor:
    """End-to-end attack: perturb vc_tgt so that model.inference(vc_src, .) moves towards the
    conversion of adv_tgt and away from the original conversion (attack_utils.py:7-48)."""
    return _run("e2e", model, vc_src, vc_tgt, adv_tgt, eps, n_iters)


def emb_attack(model: nn.Module, vc_tgt: Tensor, adv_tgt: Tensor, eps: float, n_iters: int) -> Tensor:
    """Embedding attack on model.speaker_encoder (attack_utils.py:51-86)."""
    return _run("emb", model, None, vc_tgt, adv_tgt, eps, n_iters)


def fb_attack(model: nn.Module, vc_src: Tensor, vc_tgt: Tensor, adv_tgt: Tensor, eps: float, n_iters: int) -> Tensor:
    """Feedback attack on the speaker embedding of the converted utterance (attack_utils.py:89-130)."""
    return _run("fb", model, vc_src, vc_tgt, adv_tgt, eps, n_iters)
