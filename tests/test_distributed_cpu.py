"""Host logic of the multi-GPU path under gloo, world_size 2, on CPU: batch sharding with the global
MSE normaliser + all_gather of outputs + all_reduce of loss curves == the unsharded call.  The attack
itself is the oracle here (tests may use it); on GPUs the same plumbing wraps Engine.attack."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from attack_vc_b200.distributed import global_inv_norm, shard_bounds, sharded_attack


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
