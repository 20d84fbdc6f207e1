"""Single-GPU accuracy of the trainer's gradients at the batch sizes of scripts/nccl_check.py part 3, against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.predictive import PredictiveEngine, PredictiveTrainer
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, pm_make_state_dict
from oracle import adainvc_oracle as O
from oracle import vsmask_train_oracle as V
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to("cuda"))
cpu = O.OracleAdaInVC(O.SYNTH_CONFIG, seed=0)
sd = pm_make_state_dict(0)
eps = (0.01, 0.005, 0.008)
def rel(a, b): return float((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
for B in (3, 12):
    g = torch.Generator().manual_seed(int(os.environ.get('SEED', '31')))
    s = 0.5 * torch.randn(12, 1, 80, 100, generator=g)[:B]; t = 0.5 * torch.randn(12, 1, 80, 100, generator=g)[:B]
    ref = V.train_steps(sd, cpu.speaker_encoder, [(s, t)], lr=1e-3, eps=eps, record_grads=True)
    pm = PredictiveEngine({k: v.cuda() for k, v in sd.items()})
    tr = PredictiveTrainer(pm, eng, batch_size=B, epsilon1=eps[0], epsilon2=eps[1], epsilon3=eps[2])
    l = float(tr.step(s.cuda(), t.cuda())); gg = tr.grads()
    errs = {k: rel(gg[k], ref["grads"][0][k]) for k in gg if not k.endswith("conv.1.bias")}
    w = max(errs, key=errs.get)
    import statistics
    print("   median", f"{statistics.median(errs.values()):.2e}", "env", {k: os.environ.get(k) for k in ("AVC_PM_NO_KSPLIT", "AVC_PM_NO_TC", "AVC_PM_TC_MASK")})
    print(f"B={B}: loss rel {abs(l - float(ref['losses'][0])) / abs(float(ref['losses'][0])):.2e}, within 1e-3: {sum(v < 1e-3 for v in errs.values())}/{len(errs)}, worst {errs[w]:.2e} {w}")
    tr.close(); pm.close()
