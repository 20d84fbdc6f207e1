"""Host-side phase times of attack_utils.e2e_attack at the driver's 20 iterations (AVC_TIMING=1 prints the C side)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TQDM_DISABLE", "1")
os.environ["AVC_TIMING"] = "1"
import torch
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs
from attack_vc_b200 import engine as E
import attack_utils as AU
dev = torch.device("cuda:0")
model = ParamTree(SYNTH_CONFIG, seed=0).to(dev)
host = {k: v.pin_memory() for k, v in make_inputs("e2e", 1, 256, seed=1).items()}
def call(n):
    t0 = time.perf_counter()
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items() if k != "w0"}
    t1 = time.perf_counter()
    eng = E.engine_for(model)
    t2 = time.perf_counter()
    out = AU.e2e_attack(model, d["vc_src"], d["vc_tgt"], d["adv_tgt"], 0.1, n)
    t3 = time.perf_counter()
    r = out.cpu()
    t4 = time.perf_counter()
    return [1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))]
call(3)
for i in range(5):
    torch.cuda.synchronize(); t = time.perf_counter(); ph = call(20); torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t)
    print(i, f"total {dt:.2f} ms: h2d {ph[0]:.2f}, engine_for {ph[1]:.2f}, e2e_attack {ph[2]:.2f}, d2h {ph[3]:.2f}", flush=True)
