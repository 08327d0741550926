"""GPU: the torch custom-op layer (poissonvae_b200/ops.py, `torch.ops.pvae.*`): torch.library.opcheck (schema, fake
implementations, autograd registration, AOT dispatch) and agreement with the direct C-ABI calls."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from poissonvae_b200 import ops as o
    return o


def test_opcheck_gemm_and_linear(ops):
    g = torch.Generator("cuda").manual_seed(1)
    x = torch.randn(64, 128, device="cuda", generator=g)
    w = torch.randn(96, 128, device="cuda", generator=g)
    b = torch.randn(96, device="cuda", generator=g)
    torch.library.opcheck(torch.ops.pvae.gemm.default, (x, w, True, True))
    torch.library.opcheck(torch.ops.pvae.gemm.default, (x.t().contiguous(), w.t().contiguous(), False, False))
    torch.library.opcheck(torch.ops.pvae.linear.default, (x.requires_grad_(), w.requires_grad_(), b.requires_grad_()))
    torch.library.opcheck(torch.ops.pvae.linear.default, (x, w, None))
    y = torch.ops.pvae.linear(x, w, b)
    ref = x.double() @ w.double().t() + b.double()
    assert ((y.double() - ref).abs().max() / ref.abs().max()).item() < 2e-6
    gy = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    assert ((gx.double() - gy.double() @ w.double()).abs().max() / gx.abs().max()).item() < 5e-6
    assert ((gw.double() - gy.double().t() @ x.double()).abs().max() / gw.abs().max()).item() < 5e-6
    assert torch.allclose(gb, gy.sum(0), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("indicator", ["sigmoid", "cubic"])
def test_opcheck_poisson_rsample(ops, indicator):
    g = torch.Generator("cuda").manual_seed(2)
    log_rate = (torch.randn(16, 128, device="cuda", generator=g) * 1.5 - 1.0).requires_grad_()
    args = (log_rate, 0.5, 24, indicator, 5.0, 0.0, 7, 12)  # explicit (seed, offset): a pure function of its arguments
    torch.library.opcheck(torch.ops.pvae.poisson_rsample.default, args)
    z, dz, philox = torch.ops.pvae.poisson_rsample(*args)
    assert philox.tolist() == [7, 12] and philox.device.type == "cpu"
    # the op is what Poisson.rsample runs: same draws, same bits, same gradient
    from poissonvae_b200.distributions import Poisson
    torch.cuda.manual_seed(7)
    torch.cuda.default_generators[0].set_offset(12)
    lr2 = log_rate.detach().clone().requires_grad_()
    z2 = Poisson(lr2, temp=0.5, clamp=5.0, indicator_approx=indicator, n_exp=24).rsample()
    assert torch.equal(z, z2)
    assert torch.cuda.default_generators[0].get_offset() == 16  # the op advanced the generator
    w = torch.randn_like(z)
    (g1,) = torch.autograd.grad(z, log_rate, w)
    (g2,) = torch.autograd.grad(z2, lr2, w)
    assert torch.equal(g1, g2) and torch.isfinite(g1).all() and g1.abs().max() > 0


def test_generator_state_is_consumed_inside_the_op(ops):
    log_rate = torch.zeros(8, 64, device="cuda")
    torch.cuda.manual_seed(123)
    gen = torch.cuda.default_generators[0]
    z1, _, p1 = torch.ops.pvae.poisson_rsample(log_rate, 1.0, 16)
    z2, _, p2 = torch.ops.pvae.poisson_rsample(log_rate, 1.0, 16)
    assert p1.tolist() == [123, 0] and p2.tolist() == [123, 4] and gen.get_offset() == 8
    assert not torch.equal(z1, z2)
    torch.cuda.manual_seed(123)  # torch.manual_seed governs the stream (base/config_base.py:51-63)
    z3, _, _ = torch.ops.pvae.poisson_rsample(log_rate, 1.0, 16)
    assert torch.equal(z1, z3)
    zh, dzh, _ = torch.ops.pvae.poisson_rsample(log_rate, 0.0, 16)  # temp == 0: hard counts
    assert torch.equal(zh, zh.round()) and float(dzh.abs().max()) == 0.0


@pytest.mark.parametrize("bias", [False, True])
def test_opcheck_encode_sample_kl(ops, bias):
    g = torch.Generator("cuda").manual_seed(3)
    h = (torch.randn(24, 128, device="cuda", generator=g) * 0.8).requires_grad_()
    log_r = (torch.rand(1, 128, device="cuda", generator=g) * 2 - 3).requires_grad_()
    b = (torch.randn(128, device="cuda", generator=g) * 0.1).requires_grad_() if bias else None
    args = (h, log_r, b, 0.7, 20, "sigmoid", False, 0.0, 9, 20, 0)
    torch.library.opcheck(torch.ops.pvae.encode_sample_kl.default, args)
    z, log_dr, klrow, dz, philox, rm = torch.ops.pvae.encode_sample_kl(*args)
    # against the reference's formulas in torch (same rate -> the op's z equals poisson_rsample of that log-rate)
    sp = torch.nn.functional.softplus
    hb = h if b is None else h + b
    ld = 10.0 - sp(10.0 - hb)
    lam = 5.0 - sp(5.0 - (log_r + ld))
    kl = torch.exp(log_r) * (1 + torch.exp(ld) * (ld - 1))
    assert torch.allclose(log_dr, ld, rtol=1e-5, atol=2e-6)
    assert torch.allclose(klrow, kl.sum(1), rtol=2e-5, atol=1e-6)
    z_ref, _, _ = torch.ops.pvae.poisson_rsample(lam, 0.7, 20, "sigmoid", -1.0, 0.0, 9, 20)
    assert torch.allclose(z, z_ref, rtol=1e-4, atol=1e-5)
    # gradients of a loss that uses all three differentiable outputs, against autograd through the torch formulas
    wz, wl, wk = torch.randn_like(z), torch.randn_like(ld) * 0.1, torch.randn_like(klrow)
    ins = (h, log_r) + ((b,) if bias else ())
    got = torch.autograd.grad((z * wz).sum() + (log_dr * wl).sum() + (klrow * wk).sum(), ins)
    want = torch.autograd.grad((z_ref * wz).sum() + (ld * wl).sum() + (kl.sum(1) * wk).sum(), ins)
    for a, c in zip(got, want):
        scale = c.abs().max().item()
        assert (a - c).abs().max().item() <= 2e-5 * scale, ((a - c).abs().max().item(), scale)
    # the largest rate of the call is an output (update_n's input), not a mutated argument
    assert rm.shape == (1,) and abs(float(rm) - float(torch.exp(lam.detach()).max())) <= 1e-5 * float(rm)
