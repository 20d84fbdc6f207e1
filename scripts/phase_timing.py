"""Where does the wall-clock of one public-API attack go?  (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TQDM_DISABLE", "1")
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs

kind = sys.argv[1] if len(sys.argv) > 1 else "e2e"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T = int(sys.argv[3]) if len(sys.argv) > 3 else 256
K = int(sys.argv[4]) if len(sys.argv) > 4 else 1500
dev = torch.device("cuda:0")
model = ParamTree(SYNTH_CONFIG, seed=0).to(dev)
t = time.perf_counter(); eng = Engine(model); torch.cuda.synchronize(); print(f"Engine(): {1e3*(time.perf_counter()-t):.1f} ms")
inp = {k: v.to(dev) for k, v in make_inputs(kind, B, T, seed=1).items()}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = eng.begin(kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, K, vc_src=inp.get("vc_src"), w0=inp["w0"])
    t1 = time.perf_counter(); torch.cuda.synchronize(); t1s = time.perf_counter()
    s.step(K)
    t2 = time.perf_counter(); torch.cuda.synchronize(); t2s = time.perf_counter()
    out, _ = s.end()
    t3 = time.perf_counter()
    print(f"rep {rep}: begin {1e3*(t1-t0):.1f} ms (+sync {1e3*(t1s-t1):.1f}), step enqueue {1e3*(t2-t1s):.1f} ms, drain {1e3*(t2s-t2):.1f} ms, end {1e3*(t3-t2s):.1f} ms; "
          f"{1e3*(t2s-t1s)/K:.4f} ms/iter")
s = eng.begin(kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, 8, vc_src=inp.get("vc_src"), w0=inp["w0"])
s.step(3)
prof = s.profile(); s.end()
names = {0: "conv", 1: "norm", 2: "tail", 3: "affine", 4: "loss", 5: "update", 6: "layout", 7: "copy"}
for i, (k, ms, fl, by) in enumerate(prof):
    print(f"{i:3d} {names[k]:7s} {1e3*ms:8.1f} us  {fl/1e6:9.2f} MFLOP {by/1e3:9.1f} KB" + (f"  {fl/ms/1e9:7.2f} TF/s" if fl else f"  {by/ms/1e6:7.1f} GB/s"))
print("sum", sum(p[1] for p in prof))
