import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs
dev = torch.device("cuda:0")
model = ParamTree(SYNTH_CONFIG, seed=0).to(dev)
eng = Engine(model)
inp = {k: v.to(dev) for k, v in make_inputs("e2e", 1, 256, seed=1).items()}
K = int(sys.argv[1]) if len(sys.argv) > 1 else 600
def run(w0, tag, graph=True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = eng.begin("e2e", inp["vc_tgt"], inp["adv_tgt"], 0.1, K, vc_src=inp["vc_src"], w0=w0, use_graph=graph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t1 = time.perf_counter(); s.step(K); t2 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    out, _ = s.end()
    print(f"{tag}: gpu {e0.elapsed_time(e1)/K:.4f} ms/iter, enqueue {1e3*(t2-t1)/K:.4f} ms/iter, finite={bool(torch.isfinite(out).all())}", flush=True)
for i in range(3): run(inp["w0"], f"fixed w0 #{i}")
torch.manual_seed(0)
for i in range(4): run(torch.zeros_like(inp["vc_tgt"]).normal_(0, 1), f"random w0 #{i}")
for i in range(2): run(inp["w0"], f"fixed w0 again #{i}")
run(inp["w0"], "eager (no graph)", graph=False)
