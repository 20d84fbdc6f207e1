"""Timing of avc_conv1d_wgrad (opt-in third conv kernel) at a batched size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to("cuda:0"))
for B, T, ci, co, k in ((128, 512, 128, 128, 5), (128, 512, 80, 128, 8), (64, 256, 128, 256, 5)):
    x = torch.randn(B, T, ci, device="cuda"); dy = torch.randn(B, T, co, device="cuda")
    for _ in range(3): eng.conv1d_wgrad(x, dy, k)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10): eng.conv1d_wgrad(x, dy, k)
    torch.cuda.synchronize(); dt = (time.time() - t) / 10
    print(f"wgrad B{B} T{T} k{k} {ci}->{co}: {dt*1e3:.3f} ms, {2.0*B*T*ci*co*k/dt/1e12:.2f} TFLOP/s (fp32 CUDA cores)")
