"""ncu target: the norm kernels at an HBM-resident size (2048 x 256 x 128), three launches each of fwd / fwd+skip / bwd."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree

dev = torch.device("cuda:0")
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to(dev))
B, T, C = 2048, 256, 128
g = torch.Generator(device="cpu").manual_seed(0)
y = torch.randn(B, T, C, generator=g).to(dev)
cond = torch.randn(B, 2 * C, generator=g).to(dev)
res = torch.randn(B, T, C, generator=g).to(dev)
for _ in range(3): o, stats = eng.instnorm_adain_act_fwd(y, cond, None, 1, 0.0)
for _ in range(3): eng.instnorm_adain_act_fwd(y, cond, res, 1, 0.0)
for _ in range(3): eng.instnorm_adain_act_bwd(res, y, stats, cond, 0.0)
torch.cuda.synchronize()
