"""Small end-to-end run of every code path for compute-sanitizer (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs
from attack_vc_b200.predictive import PredictiveEngine
from attack_vc_b200.synthetic import pm_make_state_dict
dev = torch.device("cuda:0")
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to(dev))
for kind, B, T in (("e2e", 1, 64), ("fb", 1, 40), ("emb", 9, 256), ("fb", 10, 200)):   # small-M kernels, then tcgen05 plans
    inp = {k: v.to(dev) for k, v in make_inputs(kind, B, T, seed=3).items()}
    out = eng.attack(kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, 2, vc_src=inp.get("vc_src"), w0=inp["w0"])
    torch.cuda.synchronize()
    print(kind, B, T, "ok", bool(torch.isfinite(out).all()), flush=True)
pm = PredictiveEngine({k: v.to(dev) for k, v in pm_make_state_dict(0).items()})
r = pm.train_step(torch.randn(2, 1, 80, 100, device=dev), want_grad_x=True)
torch.cuda.synchronize()
print("pm ok", float(r["loss"]))
x = torch.randn(3, 70, 128, device=dev); dy = torch.randn(3, 35, 80, device=dev)
dw, db = eng.conv1d_wgrad(x, dy, 5, stride=2)
torch.cuda.synchronize()
print("wgrad ok", bool(torch.isfinite(dw).all()))

