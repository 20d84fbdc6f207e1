import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from attack_vc_b200 import Engine
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree
dev = torch.device("cuda:0")
for name in ("emb_T128_it100", "fb_T64_it20", "e2e_T64_it20"):
    g = dict(np.load(f"tests/golden/{name}.npz"))
    kind = name.split("_")[0]
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    for impl in ("1", "2"):
        os.environ["AVC_CONV_IMPL"] = impl
        eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to(dev))
        for key in sorted(k for k in g if k.startswith("grad_")):
            i = int(key.split("_")[1])
            _, info = eng.attack(kind, t("vc_tgt"), t("adv_tgt"), 0.1, 1, vc_src=t("vc_src") if "vc_src" in g else None, w0=t(f"w_{i}"), want_grad=True, want_loss=True)
            gg = info["grad"].cpu().double(); ref = torch.from_numpy(g[key]).double()
            d = gg - ref
            a = float((d * ref).sum() / (ref * ref).sum())
            print(f"{name} impl {impl} {key}: rel err {float(d.norm()/ref.norm()):.3e}  (along g {a:+.2e}, residual {float((d - a*ref).norm()/ref.norm()):.2e})  loss rel {abs(float(info['losses'][0]) - g['losses'][i]) / abs(g['losses'][i]):.2e}")
        eng.close()
