"""tcgen05/TMA wgrad vs the fp32 CUDA-core wgrad and fp64 autograd: error and time per case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree

eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to("cuda"))


def ref_conv(x, w, s):
    k = w.shape[2]
    pad = (k // 2, k // 2) if k % 2 else (k // 2, k // 2 - 1)
    return F.conv1d(F.pad(x.transpose(1, 2), pad, mode="reflect"), w, None, stride=s).transpose(1, 2)


cases = [(2, 64, 128, 128, 5, 1), (3, 50, 80, 128, 3, 1), (2, 37, 128, 128, 4, 2), (4, 33, 128, 80, 1, 1), (2, 40, 128, 256, 5, 1),
         (5, 24, 1104, 128, 1, 1), (2, 30, 128, 128, 8, 3), (64, 512, 128, 128, 5, 1), (128, 512, 128, 128, 5, 2)]
for B, T, ci, co, k, s in cases:
    g = torch.Generator(device="cuda").manual_seed(B + T + k)
    x = torch.randn(B, T, ci, device="cuda", generator=g)
    w = (torch.randn(co, ci, k, device="cuda", generator=g) / (ci * k) ** 0.5).double().requires_grad_(True)
    y = ref_conv(x.double(), w, s)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    (dw_ref,) = torch.autograd.grad(y, (w,), dy.double())
    out = {}
    for impl in (1, 2):
        dw, _ = eng.conv1d_wgrad(x, dy, k, stride=s, bias=False, impl=impl)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5):
            eng.conv1d_wgrad(x, dy, k, stride=s, bias=False, impl=impl)
        b.record(); torch.cuda.synchronize()
        err = float((dw.double() - dw_ref).norm() / dw_ref.norm())
        out[impl] = (err, a.elapsed_time(b) / 5)
    fl = 2.0 * B * y.shape[1] * ci * co * k
    print(f"B{B} T{T} {ci}->{co} k{k} s{s}: simt err {out[1][0]:.2e} {out[1][1]*1e3:8.1f} us | tcgen05 err {out[2][0]:.2e} {out[2][1]*1e3:8.1f} us ({fl/out[2][1]/1e9:.1f} TF/s)")
