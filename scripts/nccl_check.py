"""2-GPU check of the sharded paths over NCCL (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/nccl_check.py
1. sharded_attack (emb, e2e): batch sliced over the ranks, global MSE normaliser, all_gather of the perturbed
   utterances + all_reduce of the loss curves == the unsharded call on one GPU.
2. sharded_header_optimize: per-iteration all_reduce of the 80*T header gradient == the unsharded optimisation.
3. data-parallel PredictiveModel (BASELINE config 5): the VSMask trainer on DIFFERENT shards per rank, BatchNorm over
   the global batch + gradient all-reduce through the library's callback == one GPU on the concatenated batch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attack_vc_b200 import Engine  # noqa: E402
from attack_vc_b200.distributed import (global_inv_norm, shard_bounds, sharded_attack, sharded_attack_shards,  # noqa: E402
                                        sharded_header_optimize)
from attack_vc_b200.synthetic import SYNTH_CONFIG, ParamTree, make_inputs  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    out = os.dup(1)
    os.dup2(2, 1)                       # NCCL prints its banner on stdout
    dist.init_process_group("nccl")
    os.dup2(out, 1)
    eng = Engine(ParamTree(SYNTH_CONFIG, seed=0).to("cuda"))
    ok = True
    for kind, B, T, n in (("emb", 7, 128, 6), ("e2e", 4, 128, 4)):
        inp = {k: v.cuda() for k, v in make_inputs(kind, B, T, seed=17).items()}
        T_out = T if kind != "e2e" else int(eng._lib.avc_decoder_frames(eng._h, T))
        inv = global_inv_norm(kind, B, 128, 80, T_out)
        adv, losses = sharded_attack(eng.attack, kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, n, inv, vc_src=inp.get("vc_src"), w0=inp["w0"])
        full, info = eng.attack(kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, n, vc_src=inp.get("vc_src"), w0=inp["w0"], want_loss=True)
        lo, hi = shard_bounds(B, world, rank)                    # the same with every rank holding only its slice
        adv2, losses2 = sharded_attack_shards(eng.attack, kind, inp["vc_tgt"][lo:hi], inp["adv_tgt"][lo:hi], 0.1, n, B, inv,
                                              vc_src_local=None if inp.get("vc_src") is None else inp["vc_src"][lo:hi], w0_local=inp["w0"][lo:hi])
        err = float((adv - full).abs().max())
        lerr = float(((losses - info["losses"]).abs() / info["losses"].abs()).max())
        good = err < 5e-6 and lerr < 1e-4 and torch.equal(adv, adv2) and torch.equal(losses, losses2)
        ok &= good
        if rank == 0:
            print(f"sharded_attack {kind} B={B} over {world} GPUs: max |adv - unsharded| {err:.2e}, loss rel err {lerr:.2e} -> {'PASS' if good else 'FAIL'}")
    B, T, n = 10, 100, 8
    inp = make_inputs("emb", B, T, seed=23)
    src, tgt = (inp["vc_tgt"] * 0.6).cuda(), (inp["adv_tgt"] * 0.6).cuda()
    hdr, losses = sharded_header_optimize(eng, src, tgt, n, epsilon=0.004)
    full, info = eng.header_optimize(src, tgt, n, epsilon=0.004, want_loss=True)
    err = (hdr - full).abs()
    lerr = float(((losses - info["losses"]).abs() / info["losses"].abs()).max())
    good = float(err.median()) < 1e-6 and float((err > 1e-4).float().mean()) < 0.02 and lerr < 1e-3
    ok &= good
    same = [torch.empty_like(hdr) for _ in range(world)]
    dist.all_gather(same, hdr.contiguous())
    ident = all(torch.equal(same[0], s) for s in same)
    ok &= ident
    if rank == 0:
        print(f"sharded_header_optimize B={B} over {world} GPUs: median |h - unsharded| {float(err.median()):.2e}, "
              f"frac > 1e-4 {float((err > 1e-4).float().mean()):.4f}, loss rel err {lerr:.2e}, identical on all ranks {ident} -> {'PASS' if good and ident else 'FAIL'}")
    # ---- 3. data-parallel VSMask trainer ----
    from attack_vc_b200.predictive import PredictiveEngine, PredictiveTrainer
    from attack_vc_b200.synthetic import pm_make_state_dict
    Bl, steps = 3, 2
    g = torch.Generator().manual_seed(31)
    data = [(0.5 * torch.randn(Bl * world, 1, 80, 100, generator=g).cuda(), 0.5 * torch.randn(Bl * world, 1, 80, 100, generator=g).cuda()) for _ in range(steps)]
    sd = {k: v.cuda() for k, v in pm_make_state_dict(0).items()}
    eps = dict(epsilon1=0.01, epsilon2=0.005, epsilon3=0.008)
    pm_one = PredictiveEngine(sd)
    tr_one = PredictiveTrainer(pm_one, eng, batch_size=Bl * world, **eps)
    pm_dp = PredictiveEngine(sd)
    pm_dp.set_process_group(None, world)
    tr_dp = PredictiveTrainer(pm_dp, eng, batch_size=Bl, inv_norm=1.0 / (Bl * world * 128), **eps)
    lerr = 0.0
    for s_, t_ in data:
        l1 = tr_one.step(s_, t_)
        l2 = tr_dp.step(s_[rank * Bl:(rank + 1) * Bl].contiguous(), t_[rank * Bl:(rank + 1) * Bl].contiguous()).clone()
        dist.all_reduce(l2)
        lerr = max(lerr, abs(float(l2) - float(l1)) / abs(float(l1)))
    a, b = tr_one.state_dict(), tr_dp.state_dict()
    serr = max(float((b[k] - a[k]).norm() / a[k].norm().clamp_min(1e-30)) for k in a if "running" in k)
    perr = max(float((b[k] - a[k]).abs().mean()) for k in a if "running" not in k)
    tr_one.close(); tr_dp.close(); pm_one.close(); pm_dp.close()
    # Gradients of ONE step on three different batches, per tensor.  The two runs sum the BatchNorm statistics in different
    # orders (one device vs partial sums per rank + all-reduce): they differ by ~1e-7 before the activations, and a PReLU /
    # LeakyReLU unit within that distance of zero takes the other branch in one of them.  With 12 windows a single such unit
    # in a deep layer moves EVERY upstream gradient by 1e-3 .. 1e-2 (seen with seed 31, tests/tools/dp_probe.py: 53 % of the tensors beyond 1e-3, one
    # device against the CPU oracle 5e-6 on one side of the kink).  A missing or wrong coupling is O(1) on every batch.  So:
    # at least two of the three batches with every tensor within 1e-3, none beyond 5e-2.
    clean, gworst = 0, 0.0
    for seed in (41, 42, 43):
        gg = torch.Generator().manual_seed(seed)
        s_ = 0.5 * torch.randn(Bl * world, 1, 80, 100, generator=gg).cuda(); t_ = 0.5 * torch.randn(Bl * world, 1, 80, 100, generator=gg).cuda()
        pm_a = PredictiveEngine(sd); tr_a = PredictiveTrainer(pm_a, eng, batch_size=Bl * world, **eps)
        pm_b = PredictiveEngine(sd); pm_b.set_process_group(None, world)
        tr_b = PredictiveTrainer(pm_b, eng, batch_size=Bl, inv_norm=1.0 / (Bl * world * 128), **eps)
        tr_a.step(s_, t_); tr_b.step(s_[rank * Bl:(rank + 1) * Bl].contiguous(), t_[rank * Bl:(rank + 1) * Bl].contiguous())
        g1, g2 = tr_a.grads(), tr_b.grads()
        errs = [float((g2[k] - g1[k]).norm() / g1[k].norm().clamp_min(1e-30)) for k in g1 if not k.endswith("conv.1.bias")]
        clean += int(max(errs) < 1e-3); gworst = max(gworst, max(errs))
        tr_a.close(); tr_b.close(); pm_a.close(); pm_b.close()
    good = lerr < 1e-4 and clean >= 2 and gworst < 5e-2 and serr < 1e-5 and perr < 0.05 * steps * 1e-3
    ok &= good
    if rank == 0:
        print(f"data-parallel VSMask trainer, {world} x {Bl} windows: loss rel err {lerr:.2e}, batches with every gradient within 1e-3: {clean}/3 (worst tensor {gworst:.2e}), "
              f"running-statistics rel err {serr:.2e}, mean |param - single GPU| {perr:.2e} -> {'PASS' if good else 'FAIL'}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
