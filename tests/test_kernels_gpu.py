"""Per-kernel parity through the C-ABI unit-test entry points, against PyTorch ops evaluated in
fp64 on the same GPU (kernels are fp32: tolerance 2e-5 relative to the output scale)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def ref_conv(x_tl, w, b, stride):
    k = w.shape[-1]
    pl, pr = k // 2, (k // 2 if k % 2 else k // 2 - 1)
    x = x_tl.transpose(1, 2)
    if pl or pr:
        x = F.pad(x, (pl, pr), mode="reflect")
    return F.conv1d(x, w, b, stride=stride).transpose(1, 2)


CONV_CASES = [
    # B, T, c_in, c_out, k, stride
    (1, 256, 128, 128, 5, 1),
    (1, 256, 128, 128, 5, 2),
    (2, 131, 128, 128, 5, 2),      # odd T, ceil-mode output length
    (1, 64, 1104, 128, 1, 1),      # in-conv over the bank concat
    (1, 32, 128, 256, 5, 1),       # decoder second conv (up=2)
    (3, 77, 128, 80, 1, 1),        # decoder out-conv, N=80
    (1, 5, 128, 128, 5, 1),        # shortest legal T for pad 2... (T > pad)
    (2, 3, 128, 128, 5, 1),
    (2, 300, 128, 128, 5, 1),      # several time tiles
    (1, 40, 128, 128, 5, 3),       # stride 3
] + [(2, 50, 80, 128, k, 1) for k in range(1, 9)]   # conv bank, even kernels pad asymmetrically


@pytest.mark.parametrize("impl", [1, 3])
@pytest.mark.parametrize("B,T,ci,co,k,s", CONV_CASES)
def test_conv1d_fwd(engine, B, T, ci, co, k, s, impl):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T + k)
    x = torch.randn(B, T, ci, device="cuda", generator=g)
    w = torch.randn(co, ci, k, device="cuda", generator=g) / (ci * k) ** 0.5
    b = torch.randn(co, device="cuda", generator=g)
    y = engine.conv1d_fwd(x, w, b, stride=s, impl=impl)
    ref = ref_conv(x.double(), w.double(), b.double(), s)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 2e-6


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("B,T,ci,co,k,s", CONV_CASES + [(16, 512, 128, 128, 5, 1), (8, 256, 80, 128, 8, 1), (64, 512, 128, 128, 5, 2)])
def test_conv1d_wgrad(engine, B, T, ci, co, k, s, impl):
    """d/dW, d/db of pad_layer + Conv1d against fp64 autograd (the param.grad the reference's loss.backward() fills).
    impl 1: exact fp32 on the CUDA cores (2e-6); impl 2: tcgen05, TMA-fed MN-major operands, 3xTF32, fp32 accumulation in
    TMEM (stated tolerance 2e-5, csrc/wgrad_tc.cuh)."""
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T + k + 13)
    x = torch.randn(B, T, ci, device="cuda", generator=g)
    w = (torch.randn(co, ci, k, device="cuda", generator=g) / (ci * k) ** 0.5).double().requires_grad_(True)
    b = torch.randn(co, device="cuda", generator=g).double().requires_grad_(True)
    y = ref_conv(x.double(), w, b, s)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    dw_ref, db_ref = torch.autograd.grad(y, (w, b), dy.double())
    dw, db = engine.conv1d_wgrad(x, dy, k, stride=s, impl=impl)
    assert dw.shape == dw_ref.shape and db.shape == db_ref.shape
    assert rel_err(dw, dw_ref) < (2e-6 if impl == 1 else 2e-5)
    assert rel_err(db, db_ref) < 2e-6
    dw2, _ = engine.conv1d_wgrad(x, dy, k, stride=s, bias=False, impl=impl)
    assert torch.equal(dw, dw2)          # fixed summation order: bit-reproducible
    if impl == 2:
        dw1, _ = engine.conv1d_wgrad(x, dy, k, stride=s, bias=False, impl=1)
        assert not torch.equal(dw, dw1)  # really the tensor-core kernel, not a silent fallback


@pytest.mark.parametrize("impl", [1, 3])
@pytest.mark.parametrize("B,T,ci,co,k,s", CONV_CASES)
def test_conv1d_dgrad(engine, B, T, ci, co, k, s, impl):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T + k + 7)
    x = torch.randn(B, T, ci, device="cuda", generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(co, ci, k, device="cuda", generator=g) / (ci * k) ** 0.5
    y = ref_conv(x, w.double(), None, s)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    (dx_ref,) = torch.autograd.grad(y, x, dy.double())
    dx = engine.conv1d_dgrad(dy, w, T, stride=s, impl=impl)
    assert dx.shape == dx_ref.shape
    assert rel_err(dx, dx_ref) < 2e-6


def ref_norm(y, cond, res, up, slope):
    x = F.instance_norm(y.transpose(1, 2), eps=1e-5)
    if cond is not None:
        C = y.shape[2]
        x = x * cond[:, C:, None] + cond[:, :C, None]
    x = F.leaky_relu(x, slope) if slope else F.relu(x)
    if res is not None:
        r = res.transpose(1, 2)
        x = x + (r.repeat_interleave(up, dim=2) if up > 1 else r)
    return x.transpose(1, 2)


@pytest.mark.parametrize("B,T,up,with_cond,with_res,slope", [
    (1, 256, 1, True, True, 0.0), (2, 64, 2, True, True, 0.0), (3, 33, 1, False, False, 0.0),
    (2, 100, 2, True, True, 0.01), (1, 7, 1, True, False, 0.0), (4, 512, 1, True, True, 0.0)])
def test_instnorm_adain_act(engine, B, T, up, with_cond, with_res, slope):
    C = 128
    g = torch.Generator(device="cuda").manual_seed(T * 10 + up)
    y = torch.randn(B, T, C, device="cuda", generator=g) * 2 + 0.5
    cond = torch.randn(B, 2 * C, device="cuda", generator=g) if with_cond else None
    res = torch.randn(B, T // up, C, device="cuda", generator=g) if with_res else None
    if with_res and T % up:
        pytest.skip("T not a multiple of up")
    out, stats = engine.instnorm_adain_act_fwd(y, cond, res, up, slope)
    yd = y.double().requires_grad_(True)
    cd = cond.double().requires_grad_(True) if with_cond else None
    ref = ref_norm(yd, cd, res.double() if with_res else None, up, slope)
    assert rel_err(out, ref) < 5e-6
    mu = y.double().mean(1)
    rstd = 1 / torch.sqrt(y.double().var(1, unbiased=False) + 1e-5)
    assert rel_err(stats[..., 0], mu) < 5e-6 and rel_err(stats[..., 1], rstd) < 5e-6
    # backward
    gup = torch.randn(B, T, C, device="cuda", generator=g)
    grads = torch.autograd.grad(ref, [yd] + ([cd] if with_cond else []), gup.double())
    gy, gcond = engine.instnorm_adain_act_bwd(gup, y, stats, cond, slope)
    assert rel_err(gy, grads[0]) < 2e-5
    if with_cond:
        assert rel_err(gcond, grads[1]) < 2e-5


@pytest.mark.parametrize("step", [1, 2, 10, 1500])
def test_adam_tanh_step_matches_torch_adam(engine, step):
    """One fused update == autograd through x + eps*tanh(w) + torch.optim.Adam.step on CPU fp32."""
    n = 80 * 128
    g = torch.Generator().manual_seed(step)
    x = torch.randn(n, generator=g)
    w = torch.randn(n, generator=g)
    m = torch.randn(n, generator=g) * 1e-8
    v = torch.rand(n, generator=g) * 1e-16
    g_adv = torch.randn(n, generator=g) * 3e-8       # same order as Adam's eps (SURVEY §5 hazard)
    eps = 0.1
    wp = w.clone().requires_grad_(True)
    opt = torch.optim.Adam([wp])
    adv = x + eps * wp.tanh()
    adv.backward(g_adv)
    opt.state[wp] = {"step": torch.tensor(float(step - 1)), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
    opt.step()
    ref_adv = (x + eps * wp.detach().tanh())
    wd, md, vd = w.cuda(), m.cuda(), v.cuda()
    adv_d = engine.adam_tanh_step(g_adv.cuda(), x.cuda(), wd, md, vd, eps, step)
    st = opt.state[wp]
    # one fp32 ulp of |w| (tanhf on CPU and GPU may differ in the last place)
    assert torch.allclose(wd.cpu(), wp.detach(), rtol=3e-7, atol=1e-7)
    # m mixes two ~1e-8 terms of either sign: absolute tolerance at 1e-6 of that scale
    assert torch.allclose(md.cpu(), st["exp_avg"], rtol=1e-5, atol=1e-14)
    assert torch.allclose(vd.cpu(), st["exp_avg_sq"], rtol=1e-5, atol=1e-22)
    assert torch.allclose(adv_d.cpu(), ref_adv, rtol=3e-7, atol=1e-7)
    # the bound is exact on the perturbation term
    assert float((eps * wd.tanh()).abs().max()) <= eps


def test_unit_timing_repeats_do_not_change_results(engine):
    """avc_unit_timing (bench.py's HBM roofline legs): the repeated launches are out of place, so the outputs equal a single
    run bit for bit, and a positive device time per launch comes back."""
    g = torch.Generator().manual_seed(3)
    y = torch.randn(4, 64, 128, generator=g).cuda(); cond = torch.randn(4, 256, generator=g).cuda()
    gup = torch.randn(4, 64, 128, generator=g).cuda()
    o1, s1 = engine.instnorm_adain_act_fwd(y, cond, None, 1, 0.2)
    gy1, gc1 = engine.instnorm_adain_act_bwd(gup, y, s1, cond, 0.2)
    engine.unit_timing(4)
    try:
        o2, s2 = engine.instnorm_adain_act_fwd(y, cond, None, 1, 0.2)
        ms_f = engine.unit_last_ms()
        gy2, gc2 = engine.instnorm_adain_act_bwd(gup, y, s2, cond, 0.2)
        ms_b = engine.unit_last_ms()
    finally:
        engine.unit_timing(0)
    assert torch.equal(o1, o2) and torch.equal(s1, s2) and torch.equal(gy1, gy2) and torch.equal(gc1, gc2)
    assert 0.0 < ms_f < 5.0 and 0.0 < ms_b < 5.0
    with pytest.raises(Exception):
        engine.unit_timing(-1)


TC_CASES = CONV_CASES + [
    (64, 256, 128, 128, 5, 1),     # 130 tiles
    (7, 100, 128, 128, 5, 1),      # tiles straddle utterances
    (16, 32, 128, 256, 5, 1),      # short utterances, N = 256 (two column passes)
    (9, 131, 128, 128, 5, 2),      # stride 2, odd T
    (4, 64, 128, 1104, 1, 1),      # dgrad of this is K=1104; forward N = 1104 -> 9 column passes
]


@pytest.mark.parametrize("impl,tol", [(2, 3e-6), (4, 3e-3), (6, 3e-6), (7, 3e-6)])
@pytest.mark.parametrize("B,T,ci,co,k,s", TC_CASES)
def test_conv1d_fwd_tensor_core(engine, B, T, ci, co, k, s, impl, tol):
    """tcgen05 path: TF32 hi*hi + BF16 correction MMA with chunked accumulation must be fp32-grade -- impl 2 (kernel picked by
    size), 6 (role-swapped N = 256 kernel, conv_tc2.cuh) and 7 (N = 128 kernel, conv_tc.cuh); a single TF32 pass (impl 4,
    measurement only) is ~2e-4."""
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T + k)
    x = torch.randn(B, T, ci, device="cuda", generator=g)
    w = torch.randn(co, ci, k, device="cuda", generator=g) / (ci * k) ** 0.5
    b = torch.randn(co, device="cuda", generator=g)
    y = engine.conv1d_fwd(x, w, b, stride=s, impl=impl)
    ref = ref_conv(x.double(), w.double(), b.double(), s)
    assert y.shape == ref.shape
    err = rel_err(y, ref)
    assert err < tol, err
    y1 = engine.conv1d_fwd(x, w, b, stride=s, impl=1)
    assert not torch.equal(y, y1), "tensor-core route silently fell back to the CUDA-core kernel"


@pytest.mark.parametrize("impl,tol", [(2, 3e-6), (4, 3e-3), (6, 3e-6), (7, 3e-6)])
@pytest.mark.parametrize("B,T,ci,co,k,s", TC_CASES)
def test_conv1d_dgrad_tensor_core(engine, B, T, ci, co, k, s, impl, tol):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T + k + 7)
    x = torch.randn(B, T, ci, device="cuda", generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(co, ci, k, device="cuda", generator=g) / (ci * k) ** 0.5
    y = ref_conv(x, w.double(), None, s)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    (dx_ref,) = torch.autograd.grad(y, x, dy.double())
    dx = engine.conv1d_dgrad(dy, w, T, stride=s, impl=impl)
    err = rel_err(dx, dx_ref)
    assert err < tol, err
    dx1 = engine.conv1d_dgrad(dy, w, T, stride=s, impl=1)
    assert not torch.equal(dx, dx1), "tensor-core route silently fell back to the CUDA-core kernel"
