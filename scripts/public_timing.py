import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TQDM_DISABLE", "1")
os.environ["AVC_TIMING"] = "1"
import torch
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs
import attack_utils as AU
dev = torch.device("cuda:0")
model = ParamTree(SYNTH_CONFIG, seed=0).to(dev)
host = {k: v.pin_memory() for k, v in make_inputs("e2e", 1, 256, seed=1).items()}
def call(n):
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items() if k != "w0"}
    return AU.e2e_attack(model, d["vc_src"], d["vc_tgt"], d["adv_tgt"], 0.1, n).cpu()
call(3)
q = "clocks.sm,clocks.mem,power.draw,pstate"
for i in range(8):
    torch.cuda.synchronize(); t = time.perf_counter(); call(1500); torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t)
    print(i, f"{dt:.1f} ms", subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip(), flush=True)
    if i == 3: time.sleep(2)
