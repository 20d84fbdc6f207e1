"""Small fixed workload for ncu captures: `python scripts/ncu_target.py emb 64 512 3` runs 3 iterations eagerly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs
kind, B, T, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dev = torch.device("cuda:0")
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to(dev))
inp = {k: v.to(dev) for k, v in make_inputs(kind, B, T, seed=9).items()}
out = eng.attack(kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, n, vc_src=inp.get("vc_src"), w0=inp["w0"], use_graph=False)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
