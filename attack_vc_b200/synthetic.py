"""Synthetic AdaIN-VC configuration, seeded random-init weights and inputs.

The reference keeps its hyper-parameters and weights outside the repository (config.yaml / model.ckpt
on Google Drive, README.md:11-12), and there is no network here, so benchmarks and tests use the
AdaIN-VC hyper-parameters with 80-bin mels and random-init weights.  ``ParamTree`` is a weights-only
``nn.Module`` with the reference's ``state_dict()`` keys and the attributes the engine reads
(``speaker_encoder.subsample`` ...); it has NO forward math, so nothing can fall back to it.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, Optional

import torch
from torch import Tensor, nn

# AdaIN-VC hyper-parameters; constructor args of models.py:126-139, 218-232, 351-363
SYNTH_CONFIG: Dict[str, Dict] = {
    "SpeakerEncoder": dict(c_in=80, c_h=128, c_out=128, kernel_size=5, bank_size=8, bank_scale=1,
                           c_bank=128, n_conv_blocks=6, n_dense_blocks=6,
                           subsample=[1, 2, 1, 2, 1, 2], act="relu", dropout_rate=0.0),
    "ContentEncoder": dict(c_in=80, c_h=128, c_out=128, kernel_size=5, bank_size=8, bank_scale=1,
                           c_bank=128, n_conv_blocks=6, subsample=[1, 2, 1, 2, 1, 2],
                           act="relu", dropout_rate=0.0),
    "Decoder": dict(c_in=128, c_cond=128, c_h=128, c_out=80, kernel_size=5, n_conv_blocks=6,
                    upsample=[2, 1, 2, 1, 2, 1], act="relu", sn=False, dropout_rate=0.0),
}


def _encoder_shapes(prefix: str, c: Dict, dense: bool) -> "OrderedDict[str, tuple]":
    out: "OrderedDict[str, tuple]" = OrderedDict()
    ks = list(range(c["bank_scale"], c["bank_size"] + 1, c["bank_scale"]))
    for i, k in enumerate(ks):
        out[f"{prefix}conv_bank.{i}.weight"] = (c["c_bank"], c["c_in"], k)
        out[f"{prefix}conv_bank.{i}.bias"] = (c["c_bank"],)
    c_cat = c["c_bank"] * len(ks) + c["c_in"]
    out[f"{prefix}in_conv_layer.weight"] = (c["c_h"], c_cat, 1)
    out[f"{prefix}in_conv_layer.bias"] = (c["c_h"],)
    for name in ("first_conv_layers", "second_conv_layers"):
        for l in range(c["n_conv_blocks"]):
            out[f"{prefix}{name}.{l}.weight"] = (c["c_h"], c["c_h"], c["kernel_size"])
            out[f"{prefix}{name}.{l}.bias"] = (c["c_h"],)
    if dense:
        for name in ("first_dense_layers", "second_dense_layers"):
            for l in range(c["n_dense_blocks"]):
                out[f"{prefix}{name}.{l}.weight"] = (c["c_h"], c["c_h"])
                out[f"{prefix}{name}.{l}.bias"] = (c["c_h"],)
        out[f"{prefix}output_layer.weight"] = (c["c_out"], c["c_h"])
        out[f"{prefix}output_layer.bias"] = (c["c_out"],)
    else:
        for name in ("mean_layer", "std_layer"):
            out[f"{prefix}{name}.weight"] = (c["c_out"], c["c_h"], 1)
            out[f"{prefix}{name}.bias"] = (c["c_out"],)
    return out


def param_shapes(cfg: Dict = SYNTH_CONFIG) -> "OrderedDict[str, tuple]":
    """state_dict key -> shape, in the reference's registration order
    (models.py:159-179 CE, :258-283 SE, :383-401 DEC, :448-452 AdaInVC)."""
    out: "OrderedDict[str, tuple]" = OrderedDict()
    out.update(_encoder_shapes("content_encoder.", cfg["ContentEncoder"], dense=False))
    out.update(_encoder_shapes("speaker_encoder.", cfg["SpeakerEncoder"], dense=True))
    d = cfg["Decoder"]
    p = "decoder."
    out[p + "in_conv_layer.weight"] = (d["c_h"], d["c_in"], 1)
    out[p + "in_conv_layer.bias"] = (d["c_h"],)
    for l in range(d["n_conv_blocks"]):
        out[f"{p}first_conv_layers.{l}.weight"] = (d["c_h"], d["c_h"], d["kernel_size"])
        out[f"{p}first_conv_layers.{l}.bias"] = (d["c_h"],)
    for l in range(d["n_conv_blocks"]):
        out[f"{p}second_conv_layers.{l}.weight"] = (d["c_h"] * d["upsample"][l], d["c_h"], d["kernel_size"])
        out[f"{p}second_conv_layers.{l}.bias"] = (d["c_h"] * d["upsample"][l],)
    for l in range(2 * d["n_conv_blocks"]):
        out[f"{p}conv_affine_layers.{l}.weight"] = (2 * d["c_h"], d["c_cond"])
        out[f"{p}conv_affine_layers.{l}.bias"] = (2 * d["c_h"],)
    out[p + "out_conv_layer.weight"] = (d["c_out"], d["c_h"], 1)
    out[p + "out_conv_layer.bias"] = (d["c_out"],)
    return out


def make_state_dict(cfg: Dict = SYNTH_CONFIG, seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """Seeded random-init weights: every tensor ~ U(-1/sqrt(fan_in), +1/sqrt(fan_in)), the
    bound PyTorch's default Conv1d/Linear init uses.  One generator, tensors in
    ``param_shapes`` order, always drawn in float64 then cast, so fp32 and fp64 copies agree."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * (seed + 1))
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    shapes = param_shapes(cfg)
    for key, shape in shapes.items():
        wshape = shapes[key[: -len("bias")] + "weight"] if key.endswith("bias") else shape
        fan_in = math.prod(wshape[1:])
        bound = 1.0 / math.sqrt(fan_in)
        t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2.0 - 1.0) * bound
        sd[key] = t.to(dtype)
    return sd


def make_inputs(kind: str, B: int, T: int, seed: int = 1, T_src: Optional[int] = None,
                T_adv: Optional[int] = None, dtype=torch.float32) -> Dict[str, Tensor]:
    """Synthetic 80-bin log-mel utterances ~N(0,1) and the initial w0 ~N(0,1)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * (seed + 1))
    T_src = T if T_src is None else T_src
    T_adv = T if T_adv is None else T_adv
    d = {
        "vc_tgt": torch.randn(B, 80, T, generator=g, dtype=torch.float64).to(dtype),
        "adv_tgt": torch.randn(B, 80, T_adv, generator=g, dtype=torch.float64).to(dtype),
        "w0": torch.randn(B, 80, T, generator=g, dtype=torch.float64).to(dtype),
    }
    if kind != "emb":
        d["vc_src"] = torch.randn(B, 80, T_src, generator=g, dtype=torch.float64).to(dtype)
    return d


class _Node(nn.Module):
    pass


class ParamTree(nn.Module):
    """Weights-only module tree with the reference's parameter names (models.py:438-452)."""

    def __init__(self, cfg: Dict = SYNTH_CONFIG, seed: int = 0, dtype=torch.float32,
                 state: Optional[Dict[str, Tensor]] = None, subnet: Optional[Callable[[str], nn.Module]] = None):
        super().__init__()
        self.cfg = cfg
        state = make_state_dict(cfg, seed, dtype) if state is None else state
        make = subnet if subnet is not None else (lambda which: _Node())
        self.content_encoder = make("content_encoder")
        self.speaker_encoder = make("speaker_encoder")
        self.decoder = make("decoder")
        for which, key in (("content_encoder", "ContentEncoder"), ("speaker_encoder", "SpeakerEncoder"),
                           ("decoder", "Decoder")):
            sub = getattr(self, which)
            for k, v in cfg[key].items():       # c_in, subsample, upsample, act ... like the reference attrs
                setattr(sub, k, v)
            sub.dropout_layer = nn.Dropout(p=cfg[key]["dropout_rate"])
        for key, value in state.items():
            parts = key.split(".")
            node: nn.Module = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            node.register_parameter(parts[-1], nn.Parameter(value.clone()))
        self._keys = list(state.keys())


# ---- VSMask PredictiveModel: architecture table and seeded synthetic weights ----
# (c_in, c_out, (stride_h, stride_w)) -- models/predictive_model.py:65-73 and :76-82
PM_DOWN = [(1, 32, (1, 2)), (32, 64, (2, 2)), (64, 128, (2, 2)), (128, 256, (2, 2)), (256, 256, (2, 2)),
           (256, 512, (2, 2)), (512, 512, (2, 2))]
PM_UP = [(512, 256), (256, 128), (128, 64), (64, 32), (32, 1)]


def pm_param_shapes() -> "OrderedDict[str, tuple]":
    """state_dict keys/shapes in the reference's registration order (num_batches_tracked included)."""
    out: "OrderedDict[str, tuple]" = OrderedDict()
    for i, (ci, co, _) in enumerate(PM_DOWN):
        p = f"down_blocks.{i}.conv."
        out[p + "1.weight"] = (co, ci, 3, 3)
        out[p + "1.bias"] = (co,)
        out[p + "2.weight"] = (co,)
        out[p + "2.bias"] = (co,)
        out[p + "2.running_mean"] = (co,)
        out[p + "2.running_var"] = (co,)
        out[p + "2.num_batches_tracked"] = ()
        out[p + "3.weight"] = (1,)
    for i, (ci, co) in enumerate(PM_UP):
        p = f"up_blocks.{i}.conv_transpose.0."
        out[p + "weight"] = (ci, co, 3, 3)
        out[p + "bias"] = (co,)
    return out


def pm_make_state_dict(seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """Seeded synthetic weights.  Conv / ConvTranspose: U(+-1/sqrt(fan_in)) like PyTorch's default;
    BatchNorm affine, running statistics and the PReLU slope are randomised too (the defaults 1/0/0/1/0.25
    would leave those code paths untested)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(424243 * (seed + 1))
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    u = lambda shape, lo, hi: (torch.rand(shape, generator=g, dtype=torch.float64) * (hi - lo) + lo)
    for key, shape in pm_param_shapes().items():
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        if ".conv.2." in key:
            if key.endswith("running_var"):
                t = u(shape, 0.5, 1.5)
            elif key.endswith("running_mean"):
                t = u(shape, -0.2, 0.2)
            elif key.endswith("weight"):
                t = u(shape, 0.8, 1.2)
            else:
                t = u(shape, -0.1, 0.1)
        elif ".conv.3." in key:
            t = u(shape, 0.1, 0.4)
        else:
            wkey = key[: -len("bias")] + "weight" if key.endswith("bias") else key
            wshape = pm_param_shapes()[wkey]
            # Conv2d fan_in = c_in*9; ConvTranspose2d weight is [c_in, c_out, 3, 3] and PyTorch uses size(1)*9
            fan_in = wshape[1] * 9
            b = 1.0 / math.sqrt(fan_in)
            t = u(shape, -b, b)
        sd[key] = t.to(dtype)
    return sd
