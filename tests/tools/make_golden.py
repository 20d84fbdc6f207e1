"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/tools/make_golden.py
Weights: oracle.make_state_dict(seed=0) loaded into the reference's AdaInVC with strict=True.
Inputs:  oracle.make_inputs(...).  The reference's own attack functions draw w0 from the global
RNG (attack_utils.py:30,68,112); we seed it, draw the same tensor first, and store it.
Per-iteration losses / gradients are captured by wrapping torch.optim.Adam.step (the reference
functions do not expose them); the reference source is not modified.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import attack_utils as RA  # noqa: E402  (reference)
import models as RM  # noqa: E402  (reference)
from oracle import adainvc_oracle as O  # noqa: E402

torch.set_num_threads(8)
OUT = os.path.join(ROOT, "tests", "golden")


class Spy:
    """Record w.grad and the loss the reference's loop produced at every Adam.step()."""

    def __init__(self, record):
        self.record = set(record)
        self.grads = {}
        self.i = 0

    def __enter__(self):
        self._orig = torch.optim.Adam.step
        spy = self

        def step(opt, *a, **k):
            p = opt.param_groups[0]["params"][0]
            if spy.i in spy.record:
                spy.grads[spy.i] = p.grad.detach().clone()
            spy.i += 1
            return spy._orig(opt, *a, **k)

        torch.optim.Adam.step = step
        return self

    def __exit__(self, *exc):
        torch.optim.Adam.step = self._orig


def case(name, kind, B, T, n_iters, record, T_src=None, T_adv=None, cli_layout=False, skip_edges=False):
    ref = RM.AdaInVC(O.SYNTH_CONFIG)
    ref.load_state_dict(O.make_state_dict(seed=0), strict=True)
    inp = O.make_inputs(kind, B, T, seed=1, T_src=T_src, T_adv=T_adv)
    if cli_layout:   # attack.py:49-50: torch.from_numpy(mel).T.unsqueeze(0) -> strides (80*T, 1, 80)
        for k in list(inp):
            inp[k] = inp[k].transpose(1, 2).contiguous().transpose(1, 2)
    torch.manual_seed(1234)
    w0 = torch.zeros_like(inp["vc_tgt"]).normal_(0, 1)
    torch.manual_seed(1234)
    with Spy(record) as spy:
        if kind == "emb":
            adv = RA.emb_attack(ref, inp["vc_tgt"], inp["adv_tgt"], 0.1, n_iters)
        elif kind == "e2e":
            adv = RA.e2e_attack(ref, inp["vc_src"], inp["vc_tgt"], inp["adv_tgt"], 0.1, n_iters)
        else:
            adv = RA.fb_attack(ref, inp["vc_src"], inp["vc_tgt"], inp["adv_tgt"], 0.1, n_iters)
    # losses: re-run through the oracle loop on the *reference* model (same arithmetic, same order)
    o = O.run_attack(kind, ref, inp["vc_tgt"], inp["adv_tgt"], 0.1, n_iters, w0, vc_src=inp.get("vc_src"), record_grads=record,
                     record_w=True)
    assert torch.equal(o["adv"], adv.detach()), "oracle loop != reference loop on the reference model"
    for i in record:
        assert torch.equal(o["grads"][i], spy.grads[i])
    with torch.no_grad():
        emb_final = ref.speaker_encoder(adv.detach())
    d = {"adv": adv.detach().numpy(), "w0": w0.numpy(), "losses": o["losses"].numpy(), "emb_final": emb_final.numpy(),
         "eps": np.float32(0.1), "n_iters": np.int32(n_iters)}
    for k, v in inp.items():
        if k != "w0":
            d[k] = v.contiguous().numpy()
    # Teacher-forced gradient vectors: (w_i, grad_i) pairs.  The loop is only piecewise smooth (ReLU
    # units of the 128-wide dense tail cross zero now and then); at such an iteration two correct fp32
    # implementations disagree by ~1e-3..1e-2 (the reference itself, 8 threads vs 1 thread: 6.8e-3 at
    # iteration 30 of config 1).  Record only iterations where the reference is NOT on such an edge:
    # its fp32 gradient at w_i must agree with an fp64 evaluation at the same w_i to 1e-5.
    m64 = O.OracleAdaInVC(O.SYNTH_CONFIG, seed=0, dtype=torch.float64)
    for i in record:
        wi = o["ws"][i]
        o64 = O.run_attack(kind, m64, inp["vc_tgt"].double(), inp["adv_tgt"].double(), 0.1, 1, wi.double(),
                           vc_src=inp["vc_src"].double() if "vc_src" in inp else None, record_grads=[0])
        g64 = o64["grads"][0]
        edge = float((spy.grads[i].double() - g64).norm() / g64.norm())
        print(f"  {name}: iteration {i}: reference fp32 vs fp64 at the same w: {edge:.2e}")
        if edge >= 1e-4:
            assert skip_edges, f"{name}: iteration {i} sits on a ReLU edge ({edge:.2e}); record another one"
            print(f"  {name}: iteration {i} sits on a ReLU edge: not recorded")
            continue
        d[f"grad_{i}"] = spy.grads[i].numpy()
        d[f"w_{i}"] = wi.numpy()
    assert sum(k.startswith("grad_") for k in d) >= 2, f"{name}: fewer than two usable teacher-forcing points"
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "loss0", float(o["losses"][0]), "lossN", float(o["losses"][-1]), "max|adv-x|", float((adv.detach() - inp["vc_tgt"]).abs().max()))


def model_vectors():
    """Forward-only vectors: SE / CE / decoder / inference outputs of the reference."""
    ref = RM.AdaInVC(O.SYNTH_CONFIG)
    ref.load_state_dict(O.make_state_dict(seed=0), strict=True)
    inp = O.make_inputs("e2e", 2, 96, seed=3, T_src=77, T_adv=50)
    with torch.no_grad():
        emb = ref.speaker_encoder(inp["vc_tgt"])
        mu, ls = ref.content_encoder(inp["vc_src"])
        out = ref.inference(inp["vc_src"], inp["vc_tgt"])
    np.savez_compressed(os.path.join(OUT, "model_fwd.npz"), vc_tgt=inp["vc_tgt"].numpy(), vc_src=inp["vc_src"].numpy(),
                        emb=emb.numpy(), mu=mu.numpy(), log_sigma=ls.numpy(), out=out.numpy())
    print("model_fwd", emb.shape, mu.shape, out.shape)


CASES = {
    "emb_T128_it100": lambda: case("emb_T128_it100", "emb", 1, 128, 100, (0, 1, 10, 99)),                      # BASELINE config 1
    "e2e_T64_it20": lambda: case("e2e_T64_it20", "e2e", 1, 64, 20, (0, 1, 19)),
    "fb_T64_it20": lambda: case("fb_T64_it20", "fb", 1, 64, 20, (0, 1, 19)),
    "emb_B2_ragged_cli": lambda: case("emb_B2_ragged_cli", "emb", 2, 75, 6, (0, 5), T_adv=131, cli_layout=True),   # odd T, T_adv != T, CLI strides
    "e2e_B2_ragged": lambda: case("e2e_B2_ragged", "e2e", 2, 64, 4, (0, 3), T_src=43, T_adv=90),
    "fb_B2_ragged": lambda: case("fb_B2_ragged", "fb", 2, 50, 4, (0, 2), T_src=61, T_adv=33),
    # BASELINE config 2 at its own size and length: e2e, 1 utterance of 80x256, 1500 iterations (~7 min of CPU here).
    # The reference's own 8-thread vs 1-thread runs of this case end 9.3e-5 apart (max |adv|) with per-iteration
    # losses within 1.5e-5 relative: the trajectory is stable, so the GPU test compares all 1500 losses and the result.
    # Late in the run the optimiser parks ReLU units near zero (iteration 700: fp32 vs fp64 gradient 1.7e-2 apart), so
    # several candidate iterations are recorded and those on an edge are dropped.
    "e2e_T256_it1500": lambda: case("e2e_T256_it1500", "e2e", 1, 256, 1500, (0, 100, 300, 700, 1100, 1400, 1499), skip_edges=True),
}

if __name__ == "__main__":
    want = sys.argv[1:] or ["model_fwd"] + list(CASES)
    for name in want:
        if name == "model_fwd":
            model_vectors()
        else:
            CASES[name]()
