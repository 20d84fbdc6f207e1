"""HBM-bound kernels at sizes where the HBM roofline applies (tensors >> the 126 MB L2 in aggregate):
InstanceNorm+AdaIN+act forward / backward and the fused tanh+Adam update, through the C-ABI unit entry points.
Prints algorithmic GB/s (fwd 8 B/element (+4 with residual), bwd 12 B/element, update 36 B/element)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree

dev = torch.device("cuda:0")
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to(dev))
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6550.7


def timed(fn, reps=10):
    eng.unit_timing(reps)
    try:
        ts = []
        for _ in range(3):
            fn(); ts.append(eng.unit_last_ms())
    finally:
        eng.unit_timing(0)
    return sorted(ts)[1]


out = {}
for B, T in ((512, 256), (2048, 256), (512, 1024)):
    C = 128
    g = torch.Generator(device="cpu").manual_seed(0)
    y = torch.randn(B, T, C, generator=g).to(dev)
    cond = torch.randn(B, 2 * C, generator=g).to(dev)
    res = torch.randn(B, T, C, generator=g).to(dev)
    n = B * T * C
    ms = timed(lambda: eng.instnorm_adain_act_fwd(y, cond, None, 1, 0.0))
    out[f"norm_fwd_{B}x{T}x{C}"] = (8 * n / ms / 1e6, ms)
    ms = timed(lambda: eng.instnorm_adain_act_fwd(y, cond, res, 1, 0.0))
    out[f"norm_fwd_res_{B}x{T}x{C}"] = (12 * n / ms / 1e6, ms)
    o, stats = eng.instnorm_adain_act_fwd(y, cond, None, 1, 0.0)
    ms = timed(lambda: eng.instnorm_adain_act_bwd(res, y, stats, cond, 0.0))
    out[f"norm_bwd_{B}x{T}x{C}"] = (12 * n / ms / 1e6, ms)
    del y, cond, res, o, stats
for B, T in ((512, 512), (2048, 512)):
    n = B * T * 80
    t = [torch.randn(n, device=dev) for _ in range(5)]
    t[4] = t[4].abs() * 1e-6
    ms = timed(lambda: eng.adam_tanh_step(t[0], t[1], t[2], t[3], t[4], 0.1, 1))
    out[f"update_{B}x{T}x80"] = (36 * n / ms / 1e6, ms)
    del t
for k, (gbs, ms) in out.items():
    print(f"{k:28s} {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs/peak:6.1%} of {peak:.0f} GB/s")
