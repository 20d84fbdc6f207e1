"""Forward parity of the B200 path (speaker encoder, full conversion) against the committed golden
vectors of the reference and against the fp64 oracle on fresh inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_golden_forward(engine, golden):
    g = golden("model_fwd")
    tgt = torch.from_numpy(g["vc_tgt"]).cuda()
    src = torch.from_numpy(g["vc_src"]).cuda()
    emb = engine.speaker_encoder(tgt)
    assert rel(emb.cpu(), torch.from_numpy(g["emb"])) < 1e-5
    out = engine.inference(src, tgt)
    assert out.shape == g["out"].shape
    assert rel(out.cpu(), torch.from_numpy(g["out"])) < 1e-4


@pytest.mark.parametrize("B,T,T_src", [(1, 256, 256), (2, 75, 43), (4, 512, 130), (1, 17, 33)])
def test_forward_vs_fp64_oracle(engine, oracle, B, T, T_src):
    m64 = oracle.OracleAdaInVC(oracle.SYNTH_CONFIG, seed=0, dtype=torch.float64)
    inp = oracle.make_inputs("e2e", B, T, seed=11, T_src=T_src)
    with torch.no_grad():
        emb_ref = m64.speaker_encoder(inp["vc_tgt"].double())
        out_ref = m64.inference(inp["vc_src"].double(), inp["vc_tgt"].double())
    emb = engine.speaker_encoder(inp["vc_tgt"].cuda())
    out = engine.inference(inp["vc_src"].cuda(), inp["vc_tgt"].cuda())
    assert rel(emb.cpu(), emb_ref) < 1e-5
    assert rel(out.cpu(), out_ref) < 1e-4


def test_noncontiguous_cli_layout(engine, oracle):
    """attack.py:49-50 hands over [1,80,T] views with strides (80*T, 1, 80)."""
    inp = oracle.make_inputs("emb", 2, 96, seed=4)
    x = inp["vc_tgt"].cuda()
    x_cli = x.transpose(1, 2).contiguous().transpose(1, 2)
    assert not x_cli.is_contiguous()
    assert torch.equal(engine.speaker_encoder(x), engine.speaker_encoder(x_cli))


def test_errors(engine, oracle):
    from attack_vc_b200 import AvcError
    inp = oracle.make_inputs("emb", 1, 64)
    with pytest.raises(AvcError):
        engine.speaker_encoder(inp["vc_tgt"])                      # CPU tensor: no fallback
    with pytest.raises(ValueError):
        engine.speaker_encoder(inp["vc_tgt"].cuda().double())
    with pytest.raises(ValueError):
        engine.speaker_encoder(inp["vc_tgt"].cuda()[:, :40])
    with pytest.raises(ValueError):
        engine.speaker_encoder(torch.randn(1, 80, 4, device="cuda"))   # too short for reflect padding
    with pytest.raises(ValueError):
        engine.attack("e2e", inp["vc_tgt"].cuda(), inp["adv_tgt"].cuda(), 0.1, 1)   # vc_src missing
