"""Debug: data-parallel trainer (real NCCL, different shards) vs one GPU vs the CPU oracle, ONE step, per-tensor errors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from attack_vc_b200 import Engine
from attack_vc_b200.predictive import PredictiveEngine, PredictiveTrainer
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, pm_make_state_dict

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
out = os.dup(1); os.dup2(2, 1); dist.init_process_group("nccl"); os.dup2(out, 1)
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to("cuda"))
Bl = 3
g = torch.Generator().manual_seed(31)
s = 0.5 * torch.randn(Bl * world, 1, 80, 100, generator=g); t = 0.5 * torch.randn(Bl * world, 1, 80, 100, generator=g)
sd = pm_make_state_dict(0)
eps = dict(epsilon1=0.01, epsilon2=0.005, epsilon3=0.008)
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))
pm1 = PredictiveEngine({k: v.cuda() for k, v in sd.items()}); tr1 = PredictiveTrainer(pm1, eng, batch_size=Bl * world, **eps)
pm2 = PredictiveEngine({k: v.cuda() for k, v in sd.items()}); pm2.set_process_group(None, world)
tr2 = PredictiveTrainer(pm2, eng, batch_size=Bl, inv_norm=1.0 / (Bl * world * 128), **eps)
l1 = tr1.step(s.cuda(), t.cuda()); g1 = tr1.grads()
l2 = tr2.step(s[rank * Bl:(rank + 1) * Bl].contiguous().cuda(), t[rank * Bl:(rank + 1) * Bl].contiguous().cuda()).clone(); g2 = tr2.grads()
dist.all_reduce(l2)
if rank == 0:
    from oracle import adainvc_oracle as O
    from oracle import vsmask_train_oracle as V
    cpu = O.OracleAdaInVC(O.SYNTH_CONFIG, seed=0)
    ref = V.train_steps(sd, cpu.speaker_encoder, [(s, t)], lr=1e-3, eps=(0.01, 0.005, 0.008), record_grads=True)["grads"][0]
    print(f"loss single {float(l1):.8e} dp {float(l2):.8e}")
    for k in g1:
        if k.endswith("conv.1.bias"): continue
        print(f"{k:42s} dp-vs-single {rel(g2[k], g1[k]):.2e}  single-vs-oracle {rel(g1[k], ref[k]):.2e}  dp-vs-oracle {rel(g2[k], ref[k]):.2e}")
dist.barrier(); tr1.close(); tr2.close(); pm1.close(); pm2.close(); dist.destroy_process_group()
