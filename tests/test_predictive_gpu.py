"""VSMask PredictiveModel on the B200 path vs the oracle (bit-identical to the reference
module, tests/test_oracle_vs_reference.py) on the same seeded weights and inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-3     # north_star tolerance


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def pm():
    from oracle import predictive_oracle as P
    from attack_vc_b200.predictive import PredictiveEngine
    sd = P.pm_make_state_dict(0)
    eng = PredictiveEngine({k: v.cuda() for k, v in sd.items()})
    return P, sd, eng


@pytest.mark.parametrize("B,F,T", [(2, 80, 100), (1, 80, 100), (3, 64, 77), (5, 80, 131)])
def test_eval_forward(pm, B, F, T):
    P, sd, eng = pm
    x = torch.randn(B, 1, F, T, generator=torch.Generator().manual_seed(B * 100 + T))
    ref = P.pm_forward(sd, x, training=False)
    out = eng.forward(x.cuda(), training=False)
    assert out.shape == ref.shape == (B, 1) + P.pm_out_shape(F, T)
    assert rel(out, ref) < 2e-5
    assert float(out.abs().max()) <= 1.0


@pytest.mark.parametrize("B,F,T", [(4, 80, 100), (2, 64, 77)])
def test_training_forward_uses_batch_statistics(pm, B, F, T):
    P, sd, eng = pm
    x = torch.randn(B, 1, F, T, generator=torch.Generator().manual_seed(7))
    ref = P.pm_forward(sd, x, training=True)
    out = eng.forward(x.cuda(), training=True)
    assert rel(out, ref) < 5e-5
    assert rel(out, P.pm_forward(sd, x, training=False)) > 1e-2      # the two modes really differ


@pytest.mark.parametrize("B,F,T", [(4, 80, 100), (3, 64, 77)])
def test_train_step_gradients(pm, B, F, T):
    """loss = out.square().mean(): every parameter gradient (wgrad, biases, BatchNorm, PReLU), d loss / d x and
    the running-statistics update, against autograd through the oracle."""
    P, sd, eng = pm
    x = torch.randn(B, 1, F, T, generator=torch.Generator().manual_seed(11))
    ref = P.pm_train_step(sd, x)
    got = eng.train_step(x.cuda(), want_grad_x=True)
    assert abs(float(got["loss"]) - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    assert rel(got["out"], ref["out"]) < 5e-5
    # north_star: gradients within 1e-3 relative.  Where the fp32 oracle is missed, an fp64 evaluation of the same step
    # arbitrates (PReLU / LeakyReLU kinks and the batch statistics make the fp32 ORACLE itself ~1e-3 from the truth for a
    # few tensors): every gradient must be within 1e-3 of one of the two.
    ref64 = None

    def check(name, g_gpu, g32, pick):
        nonlocal ref64
        e = rel(g_gpu, g32)
        tol = RTOL
        if e >= RTOL:
            if ref64 is None:
                ref64 = P.pm_train_step({k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}, x.double())
            e = min(e, rel(g_gpu, pick(ref64)))
            if g32.numel() == 1:
                # a PReLU slope's gradient is ONE number, a sum with cancellation over every negative unit of the layer: its
                # relative error is the summands' error times the condition number.  The fp32 reference implementation is
                # itself up to ~6e-4 from the truth there; the kernel may be at most 3x as far as the reference is.
                tol = max(RTOL, 3.0 * rel(g32, pick(ref64)))
        assert e < tol, (name, e, rel(g32, pick(ref64)) if ref64 is not None else None)
    check("grad_x", got["grad_x"], ref["grad_x"], lambda r: r["grad_x"])
    for k, g in ref["grads"].items():
        if k.endswith("conv.1.bias"):
            # Conv2d bias under training-mode BatchNorm: the true gradient is 0, both sides hold rounding noise
            assert float(got["grads"][k].abs().max()) < 1e-6 + 1e-3 * float(ref["grads"]["down_blocks.0.conv.2.bias"].abs().max())
            continue
        check(k, got["grads"][k], g, lambda r, k=k: r["grads"][k])
    for k, v in ref["new_stats"].items():
        assert rel(got["new_stats"][k], v) < 1e-5, k


def test_step_buffers_reused_between_calls_do_not_leak(pm):
    """The handle keeps its step buffers between calls of the same shape without a new zero fill (Arena::rewind): a step must
    not see anything of the step before it.  x1, then x2, then x1 again on the same handle: first and third answers are equal
    bit for bit (every reduction is fixed-order); likewise for the two forward modes, and across a change of shape."""
    P, sd, eng = pm
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.randn(4, 1, 80, 100, generator=g).cuda(), (3.0 * torch.randn(4, 1, 80, 100, generator=g)).cuda()
    xs = torch.randn(2, 1, 64, 77, generator=g).cuda()
    a = eng.train_step(x1, want_grad_x=True)
    eng.train_step(x2, want_grad_x=True)
    b = eng.train_step(x1, want_grad_x=True)
    eng.train_step(xs, want_grad_x=True)                     # another shape in between: the recorded sequence diverges
    c = eng.train_step(x1, want_grad_x=True)
    for r in (b, c):
        assert torch.equal(a["out"], r["out"]) and torch.equal(a["loss"], r["loss"]) and torch.equal(a["grad_x"], r["grad_x"])
        for k, v in a["grads"].items():
            assert torch.equal(v, r["grads"][k]), k
    for training in (False, True):
        o1 = eng.forward(x1, training=training).clone()
        eng.forward(x2, training=training)
        assert torch.equal(o1, eng.forward(x1, training=training))


def test_errors(pm):
    P, sd, eng = pm
    from attack_vc_b200 import AvcError
    with pytest.raises(AvcError):
        eng.forward(torch.randn(1, 1, 80, 100))                  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        eng.forward(torch.randn(1, 2, 80, 100, device="cuda"))
    with pytest.raises(ValueError):
        eng.forward(torch.randn(1, 1, 80, 1, device="cuda"))     # too small for ReflectionPad2d(1)
