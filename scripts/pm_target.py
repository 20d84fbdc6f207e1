import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attack_vc_b200.synthetic import pm_make_state_dict
from attack_vc_b200.predictive import PredictiveEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = PredictiveEngine({k: v.cuda() for k, v in pm_make_state_dict(0).items()})
x = torch.randn(B, 1, 80, 100, device="cuda")
for _ in range(2): r = eng.train_step(x)
torch.cuda.synchronize(); print("loss", float(r["loss"]))
