"""Host logic of the multi-GPU path under gloo, world_size 2, on CPU: batch sharding with the global
MSE normaliser + all_gather of outputs + all_reduce of loss curves == the unsharded call.  The attack
itself is the oracle here (tests may use it); on GPUs the same plumbing wraps Engine.attack."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from attack_vc_b200.distributed import (global_inv_norm, shard_bounds, sharded_attack, sharded_attack_shards,
                                        sharded_header_optimize)


def test_shard_bounds_cover_batch():
    for n in (0, 1, 3, 8, 64, 4096, 7):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_global_inv_norm():
    assert global_inv_norm("emb", 4096, 128, 80, 512) == 1.0 / (4096 * 128)
    assert global_inv_norm("fb", 64, 128, 80, 256) == 1.0 / (64 * 128)
    assert global_inv_norm("e2e", 2, 128, 80, 256) == 1.0 / (2 * 80 * 256)


def _oracle_attack(kind, vc_tgt, adv_tgt, eps, n_iters, vc_src=None, w0=None, inv_norm=None, want_loss=True):
    from oracle import adainvc_oracle as O
    model = O.OracleAdaInVC(O.SYNTH_CONFIG, seed=0)
    o = O.run_attack(kind, model, vc_tgt, adv_tgt, eps, n_iters, w0, vc_src=vc_src, inv_norm=inv_norm)
    return o["adv"], {"losses": o["losses"].float()}


def _worker(rank, world, port, kind, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import adainvc_oracle as O
        B, T, n = 3, 40, 3
        inp = O.make_inputs(kind, B, T, seed=13)          # replicated on every rank
        T_out = T if kind != "e2e" else 40
        inv = global_inv_norm(kind, B, 128, 80, T_out)
        adv, losses = sharded_attack(_oracle_attack, kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, n, inv,
                                     vc_src=inp.get("vc_src"), w0=inp["w0"])
        # the pre-sliced form (every rank holds only its slice -- what bench.py's sharded leg uses) must agree bit for bit
        lo, hi = shard_bounds(B, world, rank)
        loc = {k: v[lo:hi] for k, v in inp.items()}
        adv2, losses2 = sharded_attack_shards(_oracle_attack, kind, loc["vc_tgt"], loc["adv_tgt"], 0.1, n, B, inv,
                                              vc_src_local=loc.get("vc_src"), w0_local=loc["w0"])
        assert torch.equal(adv, adv2) and torch.equal(losses, losses2)
        if rank == 0:
            q.put((adv.numpy().copy(), losses.numpy().copy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind", ["emb", "e2e"])
def test_sharded_equals_unsharded_gloo(kind):
    from oracle import adainvc_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    adv, losses = (torch.from_numpy(a) for a in q.get(timeout=300))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B, T, n = 3, 40, 3
    inp = O.make_inputs(kind, B, T, seed=13)
    full_adv, info = _oracle_attack(kind, inp["vc_tgt"], inp["adv_tgt"], 0.1, n, vc_src=inp.get("vc_src"), w0=inp["w0"])
    # unsharded call uses nn.MSELoss (mean over the whole batch) == global normaliser
    assert adv.shape == full_adv.shape
    assert torch.allclose(adv, full_adv, rtol=0, atol=2e-6)
    assert torch.allclose(losses, info["losses"], rtol=1e-4, atol=0)


class _CpuHeaderSession:
    """Stand-in for HeaderSession (grad_half / all-reduce on .grad / apply_half) built from torch autograd, to
    exercise the per-iteration all-reduce plumbing under gloo."""

    def __init__(self, model, src, tgt, n, eps, lam, lr, header0, inv_norm):
        T = src.shape[-1]
        self.model, self.src, self.tgt, self.eps, self.lam, self.inv = model, src, tgt, eps, lam, inv_norm
        self.h = (torch.zeros(1, 1, 80, T) if header0 is None else header0.clone().reshape(1, 1, 80, T)).requires_grad_(True)
        self.opt = torch.optim.Adam([self.h], lr=lr)
        self.grad = torch.zeros(T * 80)
        self.losses = []

    def grad_half(self):
        enc = lambda m: self.model.speaker_encoder(m.squeeze(1))
        with torch.no_grad():
            e_s, e_t = enc(self.src), enc(self.tgt)
        e = enc(torch.clamp(self.src + self.h, -1.0, 1.0))
        loss = ((e - e_t).square().sum() - self.lam * (e - e_s).square().sum()) * self.inv
        self.opt.zero_grad()
        loss.backward()
        self.losses.append(float(loss))
        self.grad.copy_(self.h.grad[0, 0].t().reshape(-1))        # time-major like the device buffer

    def apply_half(self):
        T = self.h.shape[-1]
        self.h.grad = self.grad.reshape(T, 80).t().reshape(1, 1, 80, T).clone()
        self.opt.step()
        with torch.no_grad():
            self.h.data = torch.clamp(self.h.data, -self.eps, self.eps)

    def end(self):
        return self.h.detach().clone(), {"losses": torch.tensor(self.losses)}


class _CpuHeaderEngine:
    def __init__(self, model):
        self.model = model

    def header_begin(self, src, tgt, n, eps, lam, lr, header0=None, inv_norm=None, want_loss=True):
        return _CpuHeaderSession(self.model, src, tgt, n, eps, lam, lr, header0, inv_norm)


def _header_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import adainvc_oracle as O
        inp = O.make_inputs("emb", 3, 40, seed=21)
        src, tgt = inp["vc_tgt"].unsqueeze(1) * 0.6, inp["adv_tgt"].unsqueeze(1) * 0.6
        eng = _CpuHeaderEngine(O.OracleAdaInVC(O.SYNTH_CONFIG, seed=0))
        hdr, losses = sharded_header_optimize(eng, src, tgt, 3, epsilon=0.002, lambda_param=0.5, lr=1e-3)
        if rank == 0:
            q.put((hdr.numpy().copy(), losses.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_sharded_header_equals_unsharded_gloo():
    """Header gradients summed over 2 ranks + the same Adam step on every rank == the reference's single-process
    optimize over the whole batch (models/header_model.py:25-68)."""
    from oracle import adainvc_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_header_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    hdr, losses = (torch.from_numpy(a) for a in q.get(timeout=300))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    inp = O.make_inputs("emb", 3, 40, seed=21)
    src, tgt = inp["vc_tgt"].unsqueeze(1) * 0.6, inp["adv_tgt"].unsqueeze(1) * 0.6
    o = O.run_header(O.OracleAdaInVC(O.SYNTH_CONFIG, seed=0), src, tgt, 3, epsilon=0.002)
    assert torch.allclose(hdr, o["header"], rtol=0, atol=2e-6)
    assert torch.allclose(losses.double(), o["losses"], rtol=1e-4, atol=1e-9)


def _pm_cb_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from attack_vc_b200 import _lib
        from attack_vc_b200.predictive import allreduce_callback
        comm = torch.arange(16, dtype=torch.float32) * (rank + 1)
        cb = _lib.ALLREDUCE_FN(allreduce_callback(comm))        # exactly what PredictiveEngine.set_process_group registers
        # called through the C function-pointer type, as libavc_b200 calls it: (ctx, comm, n_floats, stream)
        rc = cb(None, C.c_void_p(comm.data_ptr()), 10, None)
        if rank == 0:
            q.put((rc, comm.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_pm_allreduce_callback_gloo():
    """The avc_allreduce_fn of the data-parallel PredictiveModel path (BatchNorm statistics, backward sums, parameter
    gradients): sums the first n floats of the registered buffer over the ranks in place and leaves the rest alone."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pm_cb_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rc, comm = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.arange(16, dtype=torch.float32)
    want[:10] *= 3.0
    assert rc == 0 and torch.equal(torch.from_numpy(comm), want)
