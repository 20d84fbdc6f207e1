import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs
from oracle import adainvc_oracle as O
dev = torch.device("cuda:0")
inp = make_inputs("emb", 2, 128, seed=1)
m64 = O.OracleAdaInVC(O.SYNTH_CONFIG, seed=0, dtype=torch.float64)
o64 = O.run_attack("emb", m64, inp["vc_tgt"].double(), inp["adv_tgt"].double(), 0.1, 1, inp["w0"].double(), record_grads=[0])
g64 = o64["grads"][0]
res = {}
for impl in ("1", "2", "5"):
    os.environ["AVC_CONV_IMPL"] = impl
    eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to(dev))
    adv, info = eng.attack("emb", inp["vc_tgt"].to(dev), inp["adv_tgt"].to(dev), 0.1, 1, w0=inp["w0"].to(dev), want_grad=True, want_loss=True)
    g = info["grad"].cpu().double()
    res[impl] = g
    d = g - g64
    print(f"impl {impl}: loss {float(info['losses'][0]):.9e} (fp64 {float(o64['losses'][0]):.9e}) grad rel err {float(d.norm()/g64.norm()):.3e}")
    per_t = d.norm(dim=1)[0] / g64.norm(dim=1)[0]
    print("   per-frame rel err: first 6", [f"{v:.1e}" for v in per_t[:6].tolist()], "mid", [f"{v:.1e}" for v in per_t[60:64].tolist()], "last 6", [f"{v:.1e}" for v in per_t[-6:].tolist()])
    # is the error a global scale?  project d on g64
    a = float((d * g64).sum() / (g64 * g64).sum())
    print(f"   component along g: {a:+.3e}; residual rel {float((d - a * g64).norm() / g64.norm()):.3e}")
    eng.close()
