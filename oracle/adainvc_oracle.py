"""CPU oracle for the attack-vc adversarial perturbation loop.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (functional PyTorch, fp32 or fp64) of the one hot
path this repository accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product path
(``attack_vc_b200`` + ``libavc_b200.so``) never does and has no CPU fallback.

Parity pinning: the reference ships no tests, golden vectors or checkpoints (SURVEY.md
§4, §8c), and its arithmetic lives in PyTorch (third party, version unpinned by the
reference; torch 2.11.0 here).  The oracle is therefore pinned by EXECUTING the unmodified
reference in the build container: ``tests/test_oracle_vs_reference.py`` imports
``/root/reference/models.py`` + ``attack_utils.py`` and asserts equality, and
``tests/tools/make_golden.py`` stores reference outputs under ``tests/golden/`` so the same
check travels to the GPU box where ``/root/reference`` does not exist.

Reference lines followed (``/root/reference``):
  pad_layer            models.py:10-30      -> _pad_conv
  pixel_shuffle_1d     models.py:33-49      -> _pixel_shuffle
  upsample             models.py:52-63      -> nearest ``repeat_interleave``
  append_cond          models.py:66-79      -> _adain
  conv_bank            models.py:82-104     -> _bank
  get_act              models.py:107-118    -> _act
  ContentEncoder.fwd   models.py:181-210    -> content_encoder
  SpeakerEncoder.fwd   models.py:285-343    -> speaker_encoder
  Decoder.fwd          models.py:403-435    -> decoder
  AdaInVC.inference    models.py:472-485    -> inference
  emb/e2e/fb_attack    attack_utils.py:7-130 -> run_attack
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from attack_vc_b200.synthetic import (SYNTH_CONFIG, ParamTree, make_inputs, make_state_dict,  # noqa: F401,E402
                                      param_shapes)

# --------------------------------------------------------------------------------------
# Functional model (channels-first [B, C, T], exactly the reference's tensor convention)
# --------------------------------------------------------------------------------------
class ActProbe:
    """Test instrument for the kink arbiter (tests/test_attacks_gpu.py).  While active it numbers every activation call
    whose input depends on the attacked tensor (autograd recording, input requires grad), records the pre-activations
    and can force chosen units onto their other branch.  The network is piecewise linear: where a ReLU unit sits within
    rounding distance of zero, two correct implementations may legitimately take different branches, and each branch
    has its own exact gradient.  Never active outside tests."""
    active: Optional["ActProbe"] = None

    def __init__(self, flip: Optional[Dict] = None, record: bool = True):
        self.flip = flip or {}          # {call index: BoolTensor (shape of the pre-activation)}: units forced to the other branch
        self.pre: Dict[int, Tensor] = {}
        self.record = record
        self.n = 0

    def __enter__(self):
        ActProbe.active = self
        return self

    def __exit__(self, *exc):
        ActProbe.active = None


def _act(x: Tensor, act: str) -> Tensor:
    # get_act, models.py:107-118: "lrelu" -> LeakyReLU() (slope 0.01), anything else ReLU.
    probe = ActProbe.active
    if probe is not None and torch.is_grad_enabled() and x.requires_grad:
        i = probe.n
        probe.n += 1
        if probe.record:
            probe.pre[i] = x.detach().clone()
        if i in probe.flip:
            on = (x.detach() > 0) ^ probe.flip[i]
            return torch.where(on, x, x * (0.01 if act == "lrelu" else 0.0))
    return F.leaky_relu(x, 0.01) if act == "lrelu" else F.relu(x)


def _pad_conv(sd, key: str, x: Tensor, stride: int = 1) -> Tensor:
    # pad_layer, models.py:23-30: reflect pad (k//2, k//2) for odd k, (k//2, k//2-1) for even k.
    w, b = sd[key + ".weight"], sd[key + ".bias"]
    k = w.shape[-1]
    left, right = k // 2, (k // 2 if k % 2 else k // 2 - 1)
    if left or right:
        x = F.pad(x, (left, right), mode="reflect")
    return F.conv1d(x, w, b, stride=stride)


def _bank(sd, prefix: str, x: Tensor, n_bank: int, act: str) -> Tensor:
    # conv_bank, models.py:98-104: [act(conv_k(x)) for k] + [x] concatenated on channels, x LAST.
    outs = [_act(_pad_conv(sd, f"{prefix}conv_bank.{i}", x), act) for i in range(n_bank)]
    return torch.cat(outs + [x], dim=1)


def _inorm(x: Tensor) -> Tensor:
    # nn.InstanceNorm1d(affine=False): per (b, c) over time, biased variance, eps=1e-5,
    # identical in train and eval (no running stats) -- models.py:176,396.
    # F.instance_norm is the ATen op the reference's module dispatches to.
    return F.instance_norm(x, eps=1e-5)


def _adain(x: Tensor, cond: Tensor) -> Tensor:
    # append_cond, models.py:76-79: first half of cond = mean, second half = std.
    p = cond.shape[1] // 2
    return x * cond[:, p:, None] + cond[:, :p, None]


def _pixel_shuffle(x: Tensor, r: int) -> Tensor:
    # pixel_shuffle_1d, models.py:44-49: out[b, c, r*t + s] = in[b, r*c + s, t].
    b, c, t = x.shape
    return x.reshape(b, c // r, r, t).permute(0, 1, 3, 2).reshape(b, c // r, t * r)


def _n_bank(c: Dict) -> int:
    return len(range(c["bank_scale"], c["bank_size"] + 1, c["bank_scale"]))


def speaker_encoder(sd, x: Tensor, cfg: Dict = SYNTH_CONFIG, prefix: str = "speaker_encoder.") -> Tensor:
    c = cfg["SpeakerEncoder"]
    act = c["act"]
    h = _bank(sd, prefix, x, _n_bank(c), act)                       # models.py:336
    h = _act(_pad_conv(sd, prefix + "in_conv_layer", h), act)        # :337-338
    for l in range(c["n_conv_blocks"]):                              # :285-305 (no norm in SE)
        y = _act(_pad_conv(sd, f"{prefix}first_conv_layers.{l}", h), act)
        y = _act(_pad_conv(sd, f"{prefix}second_conv_layers.{l}", y, stride=c["subsample"][l]), act)
        if c["subsample"][l] > 1:
            h = F.avg_pool1d(h, kernel_size=c["subsample"][l], ceil_mode=True)
        h = y + h
    v = h.mean(dim=2)                                                # AdaptiveAvgPool1d(1), :340
    for l in range(c["n_dense_blocks"]):                             # :307-325
        y = _act(F.linear(v, sd[f"{prefix}first_dense_layers.{l}.weight"], sd[f"{prefix}first_dense_layers.{l}.bias"]), act)
        y = _act(F.linear(y, sd[f"{prefix}second_dense_layers.{l}.weight"], sd[f"{prefix}second_dense_layers.{l}.bias"]), act)
        v = y + v
    return F.linear(v, sd[prefix + "output_layer.weight"], sd[prefix + "output_layer.bias"])  # :342


def content_encoder(sd, x: Tensor, cfg: Dict = SYNTH_CONFIG, prefix: str = "content_encoder."):
    c = cfg["ContentEncoder"]
    act = c["act"]
    h = _bank(sd, prefix, x, _n_bank(c), act)                       # models.py:191
    h = _act(_inorm(_pad_conv(sd, prefix + "in_conv_layer", h)), act)  # :192-194
    for l in range(c["n_conv_blocks"]):                              # :196-207
        y = _act(_inorm(_pad_conv(sd, f"{prefix}first_conv_layers.{l}", h)), act)
        y = _act(_inorm(_pad_conv(sd, f"{prefix}second_conv_layers.{l}", y, stride=c["subsample"][l])), act)
        if c["subsample"][l] > 1:
            h = F.avg_pool1d(h, kernel_size=c["subsample"][l], ceil_mode=True)
        h = y + h
    return _pad_conv(sd, prefix + "mean_layer", h), _pad_conv(sd, prefix + "std_layer", h)  # :208-209


def decoder(sd, z: Tensor, emb: Tensor, cfg: Dict = SYNTH_CONFIG, prefix: str = "decoder.") -> Tensor:
    c = cfg["Decoder"]
    act = c["act"]
    h = _act(_inorm(_pad_conv(sd, prefix + "in_conv_layer", z)), act)  # models.py:413-415
    for l in range(c["n_conv_blocks"]):                                # :417-432
        a0 = F.linear(emb, sd[f"{prefix}conv_affine_layers.{2 * l}.weight"], sd[f"{prefix}conv_affine_layers.{2 * l}.bias"])
        a1 = F.linear(emb, sd[f"{prefix}conv_affine_layers.{2 * l + 1}.weight"], sd[f"{prefix}conv_affine_layers.{2 * l + 1}.bias"])
        y = _act(_adain(_inorm(_pad_conv(sd, f"{prefix}first_conv_layers.{l}", h)), a0), act)
        y = _pad_conv(sd, f"{prefix}second_conv_layers.{l}", y)
        up = c["upsample"][l]
        if up > 1:
            y = _pixel_shuffle(y, up)                                   # shuffle BEFORE the norm, :423-426
        y = _act(_adain(_inorm(y), a1), act)
        h = y + (h.repeat_interleave(up, dim=2) if up > 1 else h)      # nearest upsample, :430-431
    return _pad_conv(sd, prefix + "out_conv_layer", h)                  # :434


def inference(sd, src: Tensor, tgt: Tensor, cfg: Dict = SYNTH_CONFIG) -> Tensor:
    mu, _ = content_encoder(sd, src, cfg)        # models.py:482 (mu only, no sampling)
    emb = speaker_encoder(sd, tgt, cfg)          # :483
    return decoder(sd, mu, emb, cfg)             # :484


# --------------------------------------------------------------------------------------
# nn.Module facade with the reference's state_dict keys and attribute names, so the
# drop-in ``attack_utils`` API can be driven without /root/reference (GPU box).
# --------------------------------------------------------------------------------------
class _SubNet(nn.Module):
    def __init__(self, owner: "OracleAdaInVC", which: str):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        self._which = which

    def forward(self, *args):
        own = self._owner
        sd = own.live_state()
        if self._which == "speaker_encoder":
            return speaker_encoder(sd, args[0], own.cfg)
        if self._which == "content_encoder":
            return content_encoder(sd, args[0], own.cfg)
        return decoder(sd, args[0], args[1], own.cfg)


class OracleAdaInVC(ParamTree):
    """Same public surface as the reference's ``AdaInVC`` (models.py:438-485): attributes
    ``content_encoder`` / ``speaker_encoder`` / ``decoder`` (callables), ``inference``, and a
    ``state_dict()`` with identical keys and shapes.  Parameters keep requires_grad=True and
    the module stays in train mode, as ``load_model`` leaves the reference (data_utils.py:220-221)."""

    def __init__(self, cfg: Dict = SYNTH_CONFIG, seed: int = 0, dtype=torch.float32,
                 state: Optional[Dict[str, Tensor]] = None):
        super().__init__(cfg, seed, dtype, state, subnet=lambda which: _SubNet(self, which))

    def live_state(self) -> Dict[str, Tensor]:
        params = dict(self.named_parameters())
        return {k: params[k] for k in self._keys}

    def inference(self, src: Tensor, tgt: Tensor) -> Tensor:
        return inference(self.live_state(), src, tgt, self.cfg)


# --------------------------------------------------------------------------------------
# Attack loops (attack_utils.py:7-130), instrumented: take w0, record loss and w.grad.
# ``model`` is anything with .speaker_encoder(x) and .inference(src, tgt): the reference's
# AdaInVC or OracleAdaInVC.  Structure mirrors the reference: Adam([w]) defaults, MSELoss
# (mean over ALL elements), targets under no_grad, loss = mse(out,tgt) - 0.1*mse(out,org).
# With a model whose params require grad this also does the reference's unread wgrad work
# and recomputes the loop-invariant content encoder every iteration, so timing it is a
# faithful CPU baseline.
# --------------------------------------------------------------------------------------
def run_attack(kind: str, model, vc_tgt: Tensor, adv_tgt: Tensor, eps: float, n_iters: int,
               w0: Tensor, vc_src: Optional[Tensor] = None, record_grads: Iterable[int] = (),
               progress=None, record_w: bool = False, inv_norm: Optional[float] = None) -> Dict[str, object]:
    if kind not in ("emb", "e2e", "fb"):
        raise NotImplementedError(kind)
    record = set(record_grads)
    w = w0.detach().clone().requires_grad_(True)              # attack_utils.py:30,68,112 (w0 injected)
    opt = torch.optim.Adam([w])                               # :31,69,113  lr 1e-3, betas (.9,.999), eps 1e-8
    mse = nn.MSELoss()                                        # :32,70,114
    if inv_norm is not None:
        # sharded call (attack_vc_b200/distributed.py): same loss, but normalised by the GLOBAL element
        # count 1/(B_total*D) instead of this slice's -- the only coupling between utterances
        mse = lambda a, b: (a - b).square().sum() * inv_norm  # noqa: E731

    def fwd(x: Tensor) -> Tensor:
        if kind == "emb":
            return model.speaker_encoder(x)                   # :79
        if kind == "e2e":
            return model.inference(vc_src, x)                 # :41
        return model.speaker_encoder(model.inference(vc_src, x))  # :123

    with torch.no_grad():                                     # :35-37, 73-75, 117-119
        org = fwd(vc_tgt)
        tgt = model.speaker_encoder(adv_tgt) if kind in ("emb", "fb") else model.inference(vc_src, adv_tgt)

    losses: List[float] = []
    grads: Dict[int, Tensor] = {}
    ws: Dict[int, Tensor] = {}
    it = range(n_iters) if progress is None else progress(range(n_iters))
    for i in it:
        adv = vc_tgt + eps * w.tanh()                         # :40,78,122
        out = fwd(adv)
        loss = mse(out, tgt) - 0.1 * mse(out, org)            # :43,81,125
        opt.zero_grad()
        loss.backward()
        if i in record:
            grads[i] = w.grad.detach().clone()
            if record_w:
                ws[i] = w.detach().clone()      # w BEFORE this iteration's Adam step (teacher forcing)
        losses.append(loss.detach())                          # no .item(): the reference never reads the loss in the loop either
        opt.step()
    with torch.no_grad():
        final = vc_tgt + eps * w.tanh()                       # :48,86,130
    return {"adv": final.detach(), "w": w.detach().clone(),
            "losses": torch.stack(losses).double().cpu() if losses else torch.zeros(0, dtype=torch.float64),
            "grads": grads, "ws": ws, "org": org.detach(), "tgt": tgt.detach()}


def run_header(model, source_mel: Tensor, target_mel: Tensor, num_iterations: int, epsilon: float = 0.1,
               lambda_param: float = 0.5, lr: float = 1e-3, header0: Optional[Tensor] = None,
               inv_norm: Optional[float] = None) -> Dict[str, object]:
    """UniversalPerturbationHeader.optimize (models/header_model.py:25-68) driven as train_header.py:46,77-81
    does (Adam([header], lr)), with speaker_encoder = model.speaker_encoder on mel.squeeze(1).
    source_mel / target_mel: [B,1,80,T].  Records the loss of every iteration and the last header gradient."""
    header = (torch.zeros((1, 1) + tuple(source_mel.shape[2:]), dtype=source_mel.dtype) if header0 is None
              else header0.detach().clone().reshape((1, 1) + tuple(source_mel.shape[2:]))).requires_grad_(True)   # :22-23
    opt = torch.optim.Adam([header], lr=lr)
    enc = lambda m: model.speaker_encoder(m.squeeze(1))
    mse = F.mse_loss if inv_norm is None else (lambda a, b: (a - b).square().sum() * inv_norm)
    losses: List[float] = []
    grad = None
    for _ in range(num_iterations):
        perturbed = torch.clamp(source_mel + header, -1.0, 1.0)            # :42-45
        with torch.no_grad():
            e_src, e_tgt = enc(source_mel), enc(target_mel)                # :48-49 (recomputed every iteration there)
        e = enc(perturbed)                                                 # :50
        loss = mse(e, e_tgt) - lambda_param * mse(e, e_src)                # :53-56
        opt.zero_grad()
        loss.backward()
        grad = header.grad.detach().clone()
        opt.step()                                                         # :59-61
        with torch.no_grad():
            header.data = torch.clamp(header.data, -epsilon, epsilon)      # :64-65
        losses.append(float(loss.detach()))
    return {"header": header.detach().clone(), "losses": torch.tensor(losses, dtype=torch.float64), "grad": grad}
