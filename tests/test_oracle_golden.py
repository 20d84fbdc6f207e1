"""Oracle vs the committed golden vectors (generated from the reference by tests/tools/make_golden.py).
Runs everywhere (CPU), so the pin travels to the GPU box."""
import numpy as np
import pytest
import torch

CASES = ["emb_T128_it100", "e2e_T64_it20", "fb_T64_it20", "emb_B2_ragged_cli", "e2e_B2_ragged", "fb_B2_ragged"]


def test_model_forward_golden(oracle, cpu_model, golden):
    g = golden("model_fwd")
    with torch.no_grad():
        emb = cpu_model.speaker_encoder(torch.from_numpy(g["vc_tgt"]))
        mu, ls = cpu_model.content_encoder(torch.from_numpy(g["vc_src"]))
        out = cpu_model.inference(torch.from_numpy(g["vc_src"]), torch.from_numpy(g["vc_tgt"]))
    # same ATen ops, but thread count / ISA may differ from the generating run: allow fp32 noise
    np.testing.assert_allclose(emb.numpy(), g["emb"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(mu.numpy(), g["mu"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(ls.numpy(), g["log_sigma"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name", CASES)
def test_attack_golden(oracle, cpu_model, golden, name):
    g = golden(name)
    kind = name.split("_")[0]
    n = int(g["n_iters"])
    n_run = min(n, 12)            # keep the CPU suite short; prefix of the trajectory is enough to pin it
    record = [int(k.split("_")[1]) for k in g if k.startswith("grad_") and int(k.split("_")[1]) < n_run]
    t = lambda k: torch.from_numpy(g[k])
    o = oracle.run_attack(kind, cpu_model, t("vc_tgt"), t("adv_tgt"), float(g["eps"]), n_run, t("w0"),
                          vc_src=t("vc_src") if "vc_src" in g else None, record_grads=record)
    np.testing.assert_allclose(o["losses"].numpy(), g["losses"][:n_run], rtol=2e-4)
    for i in record:
        ref = g[f"grad_{i}"]
        err = np.linalg.norm(o["grads"][i].numpy() - ref) / np.linalg.norm(ref)
        assert err < 1e-3, (i, err)
    if n_run == n:
        np.testing.assert_allclose(o["adv"].numpy(), g["adv"], rtol=0, atol=2e-5)
    # perturbation bound (attack_utils.py:40): |adv - x| <= eps up to one rounding of the sum
    assert np.abs(g["adv"] - g["vc_tgt"]).max() <= 0.1 * (1 + 1e-6)
