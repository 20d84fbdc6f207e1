import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree
dev = torch.device("cuda:0")
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to(dev))
def ref_conv(x_tl, w, b, stride):
    k = w.shape[-1]; pl, pr = k // 2, (k // 2 if k % 2 else k // 2 - 1)
    x = x_tl.transpose(1, 2)
    if pl or pr: x = F.pad(x, (pl, pr), mode="reflect")
    return F.conv1d(x, w, b, stride=stride).transpose(1, 2)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for pos in (False, True):
  for (B, T, ci, co, k) in [(8, 256, 128, 128, 1), (8, 256, 128, 128, 5), (8, 256, 1104, 128, 1), (8, 256, 128, 128, 8)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, T, ci, device=dev, generator=g); w = torch.randn(co, ci, k, device=dev, generator=g) / (ci * k) ** 0.5; b = torch.randn(co, device=dev, generator=g)
    if pos: x, w = x.abs(), w.abs()
    ref = ref_conv(x.double(), w.double(), b.double(), 1)
    errs = []
    for impl in (1, 4, 2, 5):
        y = eng.conv1d_fwd(x, w, b, 1, impl)
        errs.append(rel(y, ref))
        if impl == 2: bias = float(((y.double() - ref) / ref.abs().clamp_min(1e-3)).mean())
    print(f"{'positive' if pos else 'random  '} K={ci*k:5d} (k{k}): fp32 {errs[0]:.2e}  1xTF32 {errs[1]:.2e}  3xTF32 {errs[2]:.2e} (mean signed rel {bias:+.2e})  4xTF32 {errs[3]:.2e}", flush=True)
