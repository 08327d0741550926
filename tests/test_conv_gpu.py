"""GPU: the conv encoder's kernels (csrc/conv_tc.cu, csrc/conv_misc.cu) against plain
torch float64 CPU references of the same ops: implicit-GEMM convolution forward / data gradient / weight gradient
on tcgen05 (3x3 stride 1 and 2, the 2x2 FactorizedReduce window, 1x1 stride 2, odd and non-power-of-two maps),
weight normalisation, and every block of the encoder with its hand-derived backward against autograd of
oracle/conv_oracle.py.

Tolerance: the north_star's 1e-5 relative for single ops; for whole blocks (several GEMMs deep, 3xTF32 products)
5e-5 - stated per test."""
import numpy as np
import pytest
import torch
from torch.nn import functional as F

from tests.helpers import assert_rel

pytestmark = pytest.mark.gpu


def nhwc(t):  # NCHW cpu -> NHWC cuda fp32
    return t.permute(0, 2, 3, 1).contiguous().float().cuda()


def nchw(t):  # NHWC cuda -> NCHW cpu f64
    return t.detach().permute(0, 3, 1, 2).double().cpu()


def normalized(weight, lognorm):
    from oracle.conv_oracle import normalize
    return normalize(weight, lognorm)


CONV_CASES = [
    # N, H, W, Ci, Co, k, stride, pad
    (8, 16, 16, 32, 32, 3, 1, 1),
    (8, 16, 16, 32, 64, 3, 2, 1),
    (6, 8, 8, 64, 64, 3, 1, 1),
    (12, 4, 4, 128, 128, 3, 1, 1),
    (4, 8, 8, 64, 128, 3, 2, 1),
    (3, 26, 26, 32, 32, 3, 1, 1),   # map wider than any power-of-two tile: boxes overhang the tensor
    (5, 13, 13, 64, 128, 3, 2, 1),  # odd map, stride 2 -> 7x7
    (4, 7, 7, 128, 256, 3, 2, 1),   # -> 4x4, two column tiles of 128 channels
    (4, 16, 16, 32, 64, 1, 2, 0),   # StrideReduce
    (2, 2, 2, 256, 256, 3, 1, 1),   # decoder-sized map
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_forward_dgrad_wgrad(case):
    from poissonvae_b200 import conv
    N, H, W, Ci, Co, k, stride, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, Ci, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    weight = (torch.randn(Co, Ci, k, k, generator=g, dtype=torch.float64) * 0.2).requires_grad_()
    lognorm = (torch.randn(Co, generator=g, dtype=torch.float64) * 0.3).requires_grad_()
    bias = torch.randn(Co, generator=g, dtype=torch.float64, requires_grad=True)
    y_ref = F.conv2d(x, normalized(weight, lognorm), bias, stride=stride, padding=pad)
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    y_ref.backward(gy)

    wf, wd, bp = conv.weight_prep(weight.detach().float().cuda(), lognorm.detach().float().cuda(), bias.detach().float().cuda(), k * k)
    y, y_act = conv.conv_fwd(nhwc(x.detach()), wf, bp, k, stride, pad, Co, want_act=True)
    assert_rel(nchw(y), y_ref.detach(), what="conv forward")
    assert_rel(nchw(y_act), F.silu(y_ref.detach()), what="fused silu output")

    gy_d = nhwc(gy)
    gx = conv.conv_dgrad(gy_d, wd, (N, H, W, Ci), k, stride, pad)
    assert_rel(nchw(gx), x.grad, what="conv dgrad")
    # fused epilogue: (dgrad + add_pre) * silu'(src) + add_post
    a_pre, src, a_post = (torch.randn(N, Ci, H, W, generator=g, dtype=torch.float64) for _ in range(3))
    gx2 = conv.conv_dgrad(gy_d, wd, (N, H, W, Ci), k, stride, pad, add_pre=nhwc(a_pre), dsilu_src=nhwc(src), add_post=nhwc(a_post))
    sg = torch.sigmoid(src)
    assert_rel(nchw(gx2), (x.grad + a_pre) * (sg * (1 + src * (1 - sg))) + a_post, what="dgrad epilogue")

    gw, gln, gb = conv.conv_wgrad(nhwc(x.detach()), gy_d, weight.detach().float().cuda(), lognorm.detach().float().cuda(), k, stride, pad)
    assert_rel(gw.double().cpu(), weight.grad, rtol=2e-5, what="weight gradient")
    assert_rel(gln.double().cpu(), lognorm.grad, rtol=2e-5, what="lognorm gradient")
    assert_rel(gb.double().cpu(), bias.grad, what="bias gradient (ones slab)")
    assert_rel(conv.colsum(gy_d.view(-1, Co)).double().cpu(), bias.grad, what="bias gradient (column sums)")


def test_factorized_reduce_window():
    """FactorizedReduce base/common.py:98-126 as ONE 2x2 stride-2 implicit GEMM with a block-sparse packed filter."""
    from poissonvae_b200 import conv
    N, H, W, Ci, Co = 6, 16, 16, 32, 64
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, Ci, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    ws = [(torch.randn(Co // 4, Ci, 1, 1, generator=g, dtype=torch.float64) * 0.3).requires_grad_() for _ in range(4)]
    lns = [(torch.randn(Co // 4, generator=g, dtype=torch.float64) * 0.3).requires_grad_() for _ in range(4)]
    bs = [torch.randn(Co // 4, generator=g, dtype=torch.float64, requires_grad=True) for _ in range(4)]
    outs = []
    for idx in range(4):
        i, j = idx // 2, idx % 2
        outs.append(F.conv2d(x[..., i:, j:], normalized(ws[idx], lns[idx]), bs[idx], stride=2))
    y_ref = torch.cat(outs, 1)
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    y_ref.backward(gy)
    w_all = torch.cat([w.detach().flatten(1) for w in ws]).float().cuda()
    ln_all = torch.cat([l.detach() for l in lns]).float().cuda()
    b_all = torch.cat([b.detach() for b in bs]).float().cuda()
    wf, wd, bp = conv.weight_prep(w_all, ln_all, b_all, 4, groups=4)
    y, _ = conv.conv_fwd(nhwc(x.detach()), wf, bp, 2, 2, 0, Co)
    assert_rel(nchw(y), y_ref.detach(), what="factorized reduce forward")
    gx = conv.conv_dgrad(nhwc(gy), wd, (N, H, W, Ci), 2, 2, 0)
    assert_rel(nchw(gx), x.grad, what="factorized reduce dgrad")
    gw, gln, gb = conv.conv_wgrad(nhwc(x.detach()), nhwc(gy), w_all, ln_all, 2, 2, 0, groups=4)
    assert_rel(gw.double().cpu(), torch.cat([w.grad.flatten(1) for w in ws]), rtol=2e-5, what="factorized weight gradient")
    assert_rel(gln.double().cpu(), torch.cat([l.grad for l in lns]), rtol=2e-5, what="factorized lognorm gradient")
    assert_rel(gb.double().cpu(), torch.cat([b.grad for b in bs]), what="factorized bias gradient")


def test_conv_single_pass_tf32_is_coarser():
    from poissonvae_b200 import conv
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 32, 16, 16, generator=g, dtype=torch.float64)
    weight = torch.randn(32, 32, 3, 3, generator=g, dtype=torch.float64) * 0.2
    ref = F.conv2d(x, weight, None, padding=1)
    wf, _, _ = conv.weight_prep(weight.float().cuda(), None, None, 9, need_dgrad=False)
    wf_plain, _, _ = conv.weight_prep(weight.float().cuda(), None, None, 9, need_dgrad=False, with_lo=False)
    y_lo, _ = conv.conv_fwd(nhwc(x), wf, None, 3, 1, 1, 32)        # filter low parts precomputed by the prep kernel
    y_in, _ = conv.conv_fwd(nhwc(x), wf_plain, None, 3, 1, 1, 32)  # ... split in shared memory inside the kernel
    assert torch.equal(y_lo, y_in)
    errs = {}
    try:
        for mode in (True, False):
            conv.PRECISE = mode
            y, _ = conv.conv_fwd(nhwc(x), wf, None, 3, 1, 1, 32)
            errs[mode] = float((nchw(y) - ref).abs().max() / ref.abs().max())
    finally:
        conv.PRECISE = True
    assert errs[True] < 3e-6 and errs[False] > 20 * errs[True], errs


# ---- blocks ---------------------------------------------------------------------------------------------
def _module_state(mod, prefix, dtype=torch.float64):
    return {f"{prefix}.{k}" if prefix else k: v.detach().cpu().to(dtype).requires_grad_() for k, v in mod.state_dict().items()}


def _perturb(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if name.endswith("lognorm"):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def _compare_grads(mod, sd, prefix, rtol):
    for name, p in mod.named_parameters():
        ref = sd[f"{prefix}.{name}" if prefix else name].grad
        assert p.grad is not None, name
        assert_rel(p.grad.double().cpu(), ref, rtol=rtol, what=f"grad {name}")


@pytest.mark.parametrize("kind", ["normal", "down_factorized", "down_stride"])
def test_cell_forward_backward(kind):
    from oracle import conv_oracle
    from poissonvae_b200 import conv
    torch.manual_seed(11)
    ci = 32
    co = ci if kind == "normal" else 64
    cell = conv.Cell(ci, co, "normal_pre" if kind == "normal" else "down_enc", factorized=(kind != "down_stride"))
    _perturb(cell, 3)
    sd = _module_state(cell, "c")
    g = torch.Generator().manual_seed(2)
    N, H = 8, 16
    x = torch.randn(N, ci, H, H, generator=g, dtype=torch.float64, requires_grad=True)
    y_ref = conv_oracle.cell(x, sd, "c", 1 if kind == "normal" else 2)
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    y_ref.backward(gy)
    cell.cuda()
    xd = nhwc(x.detach()).requires_grad_()
    y, y_act = cell(xd)
    assert_rel(nchw(y), y_ref.detach(), rtol=2e-5, what="cell forward")
    assert_rel(nchw(y_act), F.silu(y_ref.detach()), rtol=2e-5, what="cell forward (activated copy)")
    y.backward(nhwc(gy))
    assert_rel(nchw(xd.grad), x.grad, rtol=5e-5, what="cell input gradient")
    _compare_grads(cell, sd, "c", 5e-5)


@pytest.mark.parametrize("kind", ["normal_dec", "up_dec", "up_dec_16", "normal_dec_16"])
def test_decoder_cell_forward_backward(kind):
    """Decoder cells incl. the 16-channel 28x28 layers of the MNIST decoder (stored zero-padded to 32 channels)."""
    from oracle import conv_oracle
    from poissonvae_b200 import conv
    torch.manual_seed(21)
    up = kind.startswith("up")
    ci = {"normal_dec": 64, "up_dec": 64, "up_dec_16": 32, "normal_dec_16": 16}[kind]
    co = ci // 2 if up else ci
    H = 14 if kind == "up_dec_16" else (28 if kind == "normal_dec_16" else 8)
    cell = conv.Cell(ci, co, "up_dec" if up else "normal_dec", n_nodes=1)
    _perturb(cell, 7)
    sd = _module_state(cell, "c")
    g = torch.Generator().manual_seed(3)
    N = 4
    x = torch.randn(N, ci, H, H, generator=g, dtype=torch.float64, requires_grad=True)
    y_ref = conv_oracle.dec_cell(x, sd, "c", up)
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    y_ref.backward(gy)
    cell.cuda()

    def padded(t):  # NCHW f64 -> NHWC cuda fp32 with the channel count padded to a multiple of 32
        t = nhwc(t)
        cp = conv.pad32(t.shape[-1])
        return t if cp == t.shape[-1] else torch.nn.functional.pad(t, (0, cp - t.shape[-1])).contiguous()

    xd = padded(x.detach()).requires_grad_()
    y, y_act = cell(xd)
    assert_rel(nchw(y)[:, :co], y_ref.detach(), rtol=2e-5, what="decoder cell forward")
    assert_rel(nchw(y_act)[:, :co], F.silu(y_ref.detach()), rtol=2e-5, what="decoder cell forward (activated copy)")
    assert y.shape[-1] == co or float(y.detach()[..., co:].abs().max()) == 0.0
    y.backward(padded(gy))
    assert_rel(nchw(xd.grad)[:, :ci], x.grad, rtol=5e-5, what="decoder cell input gradient")
    _compare_grads(cell, sd, "c", 5e-5)


def test_resample_input_permute_and_head():
    from oracle import conv_oracle
    from poissonvae_b200 import conv
    g = torch.Generator().manual_seed(8)
    # nn.Upsample(size=14) from 16x16, and 2x
    for (ih, oh) in ((16, 14), (7, 14), (4, 8)):
        x = torch.randn(3, 32, ih, ih, generator=g, dtype=torch.float64, requires_grad=True)
        y_ref = F.interpolate(x, size=oh, mode="nearest")
        gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
        y_ref.backward(gy)
        y = conv.upsample(nhwc(x.detach()), oh, oh)
        assert torch.equal(nchw(y), y_ref.detach().float().double())
        gx = conv.upsample_bwd(nhwc(gy), ih, ih)
        assert_rel(nchw(gx), x.grad, rtol=1e-6, what=f"resample backward {ih}->{oh}")
    # fc_dec output (N, C*4) NCHW-flat + bias -> NHWC
    h = torch.randn(6, 256 * 4, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn(256 * 4, generator=g, dtype=torch.float64, requires_grad=True)
    ref = (h + b).view(6, 256, 2, 2)
    gy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    ref.backward(gy)
    hd, bd = h.detach().float().cuda().requires_grad_(), b.detach().float().cuda().requires_grad_()
    y, y_act = conv.decoder_input(hd, bd, 256, 2)
    assert_rel(nchw(y), ref.detach(), rtol=1e-6, what="decoder input")
    assert_rel(nchw(y_act), F.silu(ref.detach()), rtol=1e-6, what="decoder input (activated copy)")
    y.backward(nhwc(gy))
    assert_rel(hd.grad.double().cpu(), h.grad, rtol=1e-6, what="decoder input gradient")
    assert_rel(bd.grad.double().cpu(), b.grad, rtol=1e-5, what="fc_dec bias gradient")
    # head: 1x1 convolution to one channel from a 16-channel map stored with 32
    torch.manual_seed(5)
    head = conv.Head(16, 1, 1, padding='valid')
    _perturb(head, 9)
    sd = _module_state(head, "h")
    x = torch.randn(5, 16, 28, 28, generator=g, dtype=torch.float64, requires_grad=True)
    y_ref = conv_oracle.conv2d(x, sd, "h")
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    y_ref.backward(gy)
    head.cuda()
    xd = torch.nn.functional.pad(nhwc(x.detach()), (0, 16)).contiguous().requires_grad_()
    y = head(xd)
    assert_rel(y.detach().double().cpu(), y_ref.detach(), what="head forward")
    y.backward(gy.float().cuda())
    assert_rel(nchw(xd.grad)[:, :16], x.grad, what="head input gradient")
    assert float(xd.grad[..., 16:].abs().max()) == 0.0
    _compare_grads(head, sd, "h", 2e-5)


def test_stem_res_conv_pool_res_dense():
    from oracle import conv_oracle
    from poissonvae_b200 import conv
    torch.manual_seed(12)
    g = torch.Generator().manual_seed(4)
    # stem
    for pad, hw in ((1, 16), (0, 28)):
        stem = conv.Stem(1, 32, 3, padding=pad if pad else 'valid')
        _perturb(stem, 5)
        sd = _module_state(stem, "stem")
        x = torch.randn(6, 1, hw, hw, generator=g, dtype=torch.float64)
        y_ref = conv_oracle.conv2d(x, sd, "stem", padding=pad)
        gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
        y_ref.backward(gy)
        stem.cuda()
        y, y_act = stem(x.float().cuda())
        assert_rel(nchw(y), y_ref.detach(), what="stem forward")
        assert_rel(nchw(y_act), F.silu(y_ref.detach()), what="stem forward (activated copy)")
        y.backward(nhwc(gy))
        _compare_grads(stem, sd, "stem", 2e-5)
    # ResConvPool
    pool = conv.ResConvPool(128)
    _perturb(pool, 6)
    sd = _module_state(pool, "p")
    x = torch.randn(12, 128, 4, 4, generator=g, dtype=torch.float64, requires_grad=True)
    y_ref = conv_oracle.res_conv_pool(x, sd, "p").flatten(1)
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    y_ref.backward(gy)
    pool.cuda()
    xd = nhwc(x.detach()).requires_grad_()
    y = pool(xd)
    assert_rel(y.detach().double().cpu(), y_ref.detach(), rtol=2e-5, what="ResConvPool forward")
    y.backward(gy.float().cuda())
    assert_rel(nchw(xd.grad), x.grad, rtol=2e-5, what="ResConvPool input gradient")
    _compare_grads(pool, sd, "p", 2e-5)
    # ResDenseLayer (eval mode: dropout off)
    dense = conv.ResDenseLayer(128).eval()
    sd = _module_state(dense, "d")
    x = torch.randn(16, 128, generator=g, dtype=torch.float64, requires_grad=True)
    y_ref = conv_oracle.res_dense(x, sd, "d")
    gy = torch.randn(y_ref.shape, generator=g, dtype=torch.float64)
    y_ref.backward(gy)
    dense.cuda()
    xd = x.detach().float().cuda().requires_grad_()
    y = dense(xd)
    assert_rel(y.detach().double().cpu(), y_ref.detach(), rtol=2e-5, what="ResDenseLayer forward")
    y.backward(gy.float().cuda())
    assert_rel(xd.grad.double().cpu(), x.grad, rtol=2e-5, what="ResDenseLayer input gradient")
    _compare_grads(dense, sd, "d", 2e-5)


def test_res_dense_dropout_statistics():
    """Training mode: the Philox mask keeps ~90 % of the hidden units, scales them by 1/0.9, and the backward uses
    the same mask (finite-difference check of one coordinate is not possible under a random mask, so check
    that a second forward with the generator rewound reproduces the output bit for bit)."""
    from poissonvae_b200 import conv
    torch.manual_seed(13)
    dense = conv.ResDenseLayer(128).cuda().train()
    x = torch.randn(64, 128, device="cuda")
    state = torch.cuda.get_rng_state()
    y1 = dense(x)
    torch.cuda.set_rng_state(state)
    y2 = dense(x)
    assert torch.equal(y1, y2)
    y3 = dense(x)
    assert not torch.equal(y1, y3)


# ---- whole model against the unmodified reference ---------------------------------------------------------
import glob
import os

GOLDEN_CONV = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "conv_enc_*.npz")) +
                     glob.glob(os.path.join(os.path.dirname(__file__), "golden", "conv_vae_*.npz")))


@pytest.mark.parametrize("path", GOLDEN_CONV)
def test_conv_vae_step_matches_reference(path):
    """conv+b|lin (BASELINE config 4, MNIST-shaped encoder) and conv+b|conv+b (BASELINE config 5: MNIST-shaped 28x28; and
    its 16x16 variant): forward, loss and every parameter gradient of one step against the fixture the unmodified reference produced with the same parameters, input and exponential draws.
    Tolerance 5e-5 relative: ~25 chained 3xTF32 GEMMs deep (each ~2e-6), stated here as the floating-point bound."""
    from tests.helpers import assert_grad_digest, conv_fixture_model
    g = np.load(path)
    model = conv_fixture_model(g).cuda()
    model.n_exp.fill_(int(g["n_exp"]))
    model._n_exp_host = int(g["n_exp"])
    model.update_t(float(g["temp"]))
    x = torch.tensor(g["x"], device="cuda")
    with torch.no_grad():
        assert_rel(model.encode(x).cpu().numpy(), g["enc_out"], rtol=2e-5, what="encoder output")
    torch.cuda.manual_seed(int(g["philox_seed"]))
    torch.cuda.default_generators[torch.cuda.current_device()].set_offset(int(g["philox_offset"]))
    dist, log_dr, z, y = model(x)
    recon = model.loss_recon(y, x)
    kl = model.loss_kl(log_dr)
    loss = torch.mean(recon + float(g["beta"]) * torch.sum(kl, dim=1))
    loss.backward()
    for name, got in (("log_dr", log_dr), ("z", z), ("y", y), ("recon", recon), ("kl", kl), ("loss", loss), ("rate", dist.rate)):
        assert_rel(got.detach().cpu().numpy(), g[name], rtol=5e-5, what=name)
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        assert_grad_digest(p.grad.cpu().numpy(), g, name, 5e-5)


def test_conv_trainer_steps():
    """TrainerVAE on conv+b|lin: the flat gradient buffer the gather kernel fills equals the model's autograd
    gradients for the same Philox state, parameters move, the loss goes down over a few steps, launches are counted."""
    from poissonvae_b200 import _lib
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    cfg_vae, cfg_tr = default_configs("CIFAR16", "poisson", "conv+b|lin")
    model = PoissonVAE(ConfigPoisVAE(seed=1, **cfg_vae)).cuda()
    model.update_n(8.0)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(64, 1, 16, 16, generator=g)
    x = ((x - x.mean((1, 2, 3), keepdim=True)) / x.std((1, 2, 3), keepdim=True)).cuda()
    tr = TrainerVAE(model, ConfigTrainVAE(lr=0.002, batch_size=64, epochs=50, warmup_portion=0.0, temp_anneal_portion=0.0,
                                          temp_stop=1.0, kl_anneal_portion=0.0, kl_const_portion=0.0, kl_beta=1.0),
                    n_batches_per_epoch=1)
    assert not tr.fused and len(tr.views) == len(list(model.parameters()))
    # reference gradients through the model's public API, same Philox state, eval-mode dropout on both sides
    model.eval()
    philox = (1234, 8)
    model.philox_override = philox
    dist, log_dr, z, y = model(x)
    loss = torch.mean(model.loss_recon(y, x) + torch.sum(model.loss_kl(log_dr), dim=1))
    loss.backward()
    model.philox_override = None
    want = torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()]).clone()
    before = tr.flat_p.clone()
    c0 = _lib.load().pvae_launch_count()
    out = tr.train_step(x, 0, philox=philox)
    launches = _lib.load().pvae_launch_count() - c0
    assert_rel(tr.flat_g.cpu().numpy(), want.cpu().numpy(), rtol=1e-6, what="flat gradient buffer")
    assert_rel(out[0].item(), loss.item(), rtol=1e-6, what="loss")
    assert not torch.equal(before, tr.flat_p) and torch.isfinite(tr.flat_p).all()
    assert 100 < launches < 400, launches
    model.train()
    losses = [tr.train_step(x, i)[0].item() for i in range(1, 40)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.995 * np.mean(losses[:5]), losses


def test_config5_temperature_annealed_training_and_hard_count_eval():
    """BASELINE config 5: poisson conv+b|conv+b on MNIST-shaped 28x28 inputs, K = 512, linear temperature annealing
    1 -> 0.05 over the first half (main/config_vae.py:278-281), gradient clipping at 1000, then the evaluation pass at
    temp = 0 (hard spike counts, base/helper.py:15-17)."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    cfg_vae, cfg_tr = default_configs("MNIST", "poisson", "conv+b|conv+b")
    cfg_vae.update(n_latents=512, seed=2)
    model = PoissonVAE(ConfigPoisVAE(**cfg_vae)).cuda()
    model.update_n(12.0)
    g = torch.Generator().manual_seed(5)
    data = torch.rand(64, 1, 28, 28, generator=g).cuda()
    cfg_tr.update(epochs=20)
    tr = TrainerVAE(model, ConfigTrainVAE(**cfg_tr), n_batches_per_epoch=2)
    assert tr.temperatures[0] == 1.0 and abs(tr.temperatures[-1] - 0.05) < 1e-12
    first = last = None
    for ep in range(20):
        for i in range(2):
            out = tr.train_step(data[32 * i:32 * (i + 1)], 2 * ep + i)
            if first is None:
                first = out.clone()
            last = out
    assert abs(model._temp_host - 0.05) < 1e-6 and torch.isfinite(tr.flat_p).all()
    assert last[1].item() < first[1].item(), (first.tolist(), last.tolist())  # reconstruction error went down
    model.eval()
    data_out, loss, etc = tr.validate(data, temp=0.0, batch_size=32)
    z = data_out["z"]
    assert z.shape == (64, 512) and np.array_equal(z, np.round(z)) and z.min() >= 0
    assert loss["mse"].shape == (64,) and np.isfinite(loss["mse"]).all() and np.isfinite(loss["kl"]).all()


@pytest.mark.parametrize("kind", ["normal_pre", "down_enc", "down_enc_stride", "normal_dec", "up_dec", "up_dec_16"])
def test_native_cell_call_equals_per_kernel_path(kind):
    """pvae_cell_fwd / pvae_cell_bwd (one C call per cell, csrc/conv_net.cu) enqueue the same kernels in the same order
    as the per-kernel Python path: outputs and every gradient are bit-identical."""
    from poissonvae_b200 import conv
    torch.manual_seed(31)
    dec = kind.endswith("dec") or kind == "up_dec_16"
    ci = 32 if kind != "normal_dec" else 64
    co = {"normal_pre": 32, "down_enc": 64, "down_enc_stride": 64, "normal_dec": 64, "up_dec": 32, "up_dec_16": 16}[kind]
    if kind == "up_dec":
        ci, co = 64, 32
    cell_type = {"normal_pre": "normal_pre", "down_enc": "down_enc", "down_enc_stride": "down_enc", "normal_dec": "normal_dec",
                 "up_dec": "up_dec", "up_dec_16": "up_dec"}[kind]
    cell = conv.Cell(ci, co, cell_type, n_nodes=1 if dec else 2, factorized=(kind != "down_enc_stride"))
    _perturb(cell, 3)
    cell.cuda()
    H = 13 if kind == "down_enc_stride" else 14
    if kind == "down_enc":
        H = 16
    x = torch.randn(6, H, H, conv.pad32(ci), device="cuda")
    gy = None
    results = {}
    try:
        for native in (True, False):
            conv.NATIVE_CELLS = native
            for p in cell.parameters():
                p.grad = None
            xd = x.clone().requires_grad_()
            y, y_act = cell(xd)
            if gy is None:
                gy = torch.randn_like(y)
            y.backward(gy)
            results[native] = [y.detach(), y_act.detach(), xd.grad] + [p.grad.clone() for p in cell.parameters()]
    finally:
        conv.NATIVE_CELLS = True
    for a, b in zip(results[True], results[False]):
        assert torch.equal(a, b)


def test_conv_public_api_paths():
    """The remaining public entry points on a conv model: prior sampling through the conv decoder, a state_dict round
    trip, an epoch over a device-resident dataset (`iteration`, with the end-of-epoch n_exp update) and host-fed training
    (`fit_host`: pinned batches in, losses out)."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    cfg_vae, cfg_tr = default_configs("CIFAR16", "poisson", "conv+b|conv+b")
    cfg_vae.update(n_latents=64, seed=5)
    model = PoissonVAE(ConfigPoisVAE(**cfg_vae)).cuda()
    model.update_n(10.0)
    xs, spks = model.sample(8, t=0.0)  # main/vae.py:431-444
    assert xs.shape == (8, 1, 16, 16) and spks.shape == (8, 64) and torch.equal(spks, spks.round())
    assert torch.isfinite(xs).all()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    twin = PoissonVAE(ConfigPoisVAE(**{**cfg_vae, "seed": 6})).cuda()
    twin.load_state_dict(sd)
    twin.eval(), model.eval()
    x = torch.randn(8, 1, 16, 16, device="cuda")
    torch.cuda.manual_seed(3)
    y1 = model(x)[3]
    torch.cuda.manual_seed(3)
    y2 = twin(x)[3]
    assert torch.equal(y1, y2)
    model.train()
    cfg_tr.update(epochs=3, batch_size=16)
    tr = TrainerVAE(model, ConfigTrainVAE(**cfg_tr), n_batches_per_epoch=4)
    data = torch.randn(64, 1, 16, 16, device="cuda")
    n0 = model._n_exp_host
    nelbo = tr.iteration(data, epoch=0)
    assert np.isfinite(nelbo) and model._n_exp_host != n0  # r_max of the epoch -> new n_exp (main/train_vae.py:456-457)
    host = torch.randn(6, 16, 256).pin_memory()
    losses = tr.fit_host(host, first_gstep=4)
    torch.cuda.synchronize()
    assert losses.shape == (6, 3) and torch.isfinite(losses).all()


def test_validate_runs_the_conv_encoder_in_eval_mode():
    """the reference evaluates model.eval() (base/train_base.py:197-198).  validate() on a conv+b|lin model
    that is in train mode must not apply the ResDenseLayer dropout: two passes agree bit for bit (same Philox state),
    equal the pass of a model put in eval mode by hand, leave the training flag alone, and differ from a pass that
    really drops activations."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    cfg_vae, _ = default_configs("CIFAR16", "poisson", "conv+b|lin")
    model = PoissonVAE(ConfigPoisVAE(seed=1, **cfg_vae)).cuda()
    with torch.no_grad():
        model.log_rate.add_(3.0)
    model.update_n(20.0)
    tr = TrainerVAE(model, ConfigTrainVAE(batch_size=32, epochs=1, grad_clip=None))
    data = torch.randn(64, 1, 16, 16, generator=torch.Generator().manual_seed(5)).cuda()
    gen = torch.cuda.default_generators[0]

    def run(fn):
        torch.cuda.manual_seed(21)
        out = fn()
        return out, gen.get_offset()

    model.train()
    (d1, l1, e1), off1 = run(lambda: tr.validate(data))
    assert model.training  # restored
    (d2, l2, e2), off2 = run(lambda: tr.validate(data))
    model.eval()
    (d3, l3, e3), off3 = run(lambda: tr.validate(data))
    for a, b in ((d1, d2), (d1, d3)):
        assert np.array_equal(a['z'], b['z'])
    assert np.array_equal(e1['log_dr'], e2['log_dr']) and np.array_equal(e1['log_dr'], e3['log_dr'])
    assert np.array_equal(l1['mse'], l3['mse']) and off1 == off2 == off3  # no dropout draws consumed
    # a pass with dropout really on moves log_dr
    model.train()
    torch.cuda.manual_seed(21)
    with torch.no_grad():
        _, log_dr_drop = model.infer(data[:32])
    assert not np.array_equal(log_dr_drop.cpu().numpy(), e1['log_dr'][:32])
