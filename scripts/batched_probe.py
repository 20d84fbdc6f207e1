import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs
dev = torch.device("cuda:0")
eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to(dev))
names = {0: "conv", 1: "norm", 2: "tail", 3: "affine", 4: "loss", 5: "update", 6: "layout", 7: "copy"}
cases = [("emb", 128, 512), ("fb", 64, 256)] if len(sys.argv) < 2 else [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))]
verbose = len(sys.argv) > 4
for kind, B, T in cases:
    inp = {k: v.to(dev) for k, v in make_inputs(kind, B, T, seed=9).items()}
    n = 6
    s = eng.begin(kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, 3 + n + 1, vc_src=inp.get("vc_src"), w0=inp["w0"], want_loss=True)
    s.step(3)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); s.step(n); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    prof = s.profile(); out, info = s.end()
    agg = {}
    for k, m, f, by in prof:
        d = agg.setdefault(names[k], [0, 0.0, 0.0, 0.0]); d[0] += 1; d[1] += m; d[2] += f; d[3] += by
    print(f"{kind} B{B} T{T}: {ms:.3f} ms/iter, {B*1e3/ms:.0f} utt-it/s, loss {float(info['losses'][0]):.4e} -> {float(info['losses'][n+3]):.4e}")
    for k, d in agg.items():
        print(f"   {k:7s} x{d[0]:3d} {d[1]:8.3f} ms  {d[2]/d[1]/1e9 if d[2] else 0:8.2f} TF/s  {d[3]/d[1]/1e6 if d[3] else 0:8.1f} GB/s")
    if verbose:
        for i, (k, m, f, by) in enumerate(prof):
            print(f"{i:3d} {names[k]:7s} {1e3*m:9.1f} us {f/m/1e9 if f else 0:8.2f} TF/s {by/m/1e6:8.1f} GB/s")
