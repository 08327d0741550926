"""GPU: the host-side mirror (Poisson, PoissonVAE) driven like the reference drives its own classes,
against fixtures the unmodified reference produced (oracle/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.helpers import assert_rel

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "vae_step_*.npz")))


def set_philox(seed, offset):
    torch.cuda.manual_seed(seed)
    torch.cuda.default_generators[torch.cuda.current_device()].set_offset(offset)


def test_poisson_class_matches_reference_fixture(golden_dir):
    from poissonvae_b200 import Poisson
    g = np.load(os.path.join(golden_dir, "sampler_sigmoid.npz"))
    for temp in (1.0, 0.3, 0.05):
        log_rate = torch.tensor(g["log_rate"], device="cuda", requires_grad=True)
        set_philox(int(g["seed"]), int(g["offset"]))
        dist = Poisson(log_rate=log_rate, temp=temp, indicator_approx="sigmoid", n_exp=int(g["n_exp"]))
        z = dist.rsample()
        (grad,) = torch.autograd.grad((z * torch.tensor(g["w"], device="cuda")).sum(), log_rate)
        assert_rel(z.detach().cpu().numpy(), g[f"z_t{temp:g}"], what="z")
        assert_rel(grad.cpu().numpy(), g[f"g_t{temp:g}"], what="grad")
        assert_rel(dist.rate.detach().cpu().numpy(), g["rate"], rtol=1e-6, what="rate")
        assert dist.mean is dist.rate and dist.variance is dist.rate


def test_poisson_class_api():
    from poissonvae_b200 import Poisson
    lr = torch.full((8, 64), 1.0, device="cuda")
    with pytest.raises(AssertionError):
        Poisson(lr, temp=-1.0)
    with pytest.raises(AssertionError):
        Poisson(lr, temp=1.0, indicator_approx="step")
    with pytest.raises(RuntimeError):
        Poisson(lr.cpu(), temp=1.0)
    d = Poisson(lr, temp=0.0, n_exp="infer", n_exp_p=1e-3)  # base/distributions.py:28-29,40-44
    from scipy import stats
    assert d.n_exp == int(stats.poisson(float(np.exp(1.0)) + 1.19e-7).ppf(1 - 1e-3))
    s = d.rsample()  # temp == 0 -> sample()
    assert torch.equal(s, s.round()) and not s.requires_grad
    lp = d.log_prob(s)
    want = -d.rate - torch.lgamma(s + 1) + s * torch.log(d.rate)
    assert torch.equal(lp, want)
    assert "n_exp" in repr(d)


@pytest.mark.parametrize("path", GOLDEN)
def test_vae_forward_backward_matches_reference(path):
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    g = np.load(path)
    K = g["p0_log_rate"].shape[1]
    cfg = ConfigPoisVAE(n_latents=K, prior_clamp=-4, exc_only=bool(g["exc_only"]), seed=int(g["cfg_seed"]))
    model = PoissonVAE(cfg)
    # same seed -> the reference's own initial parameters
    sd = model.state_dict()
    assert list(sd) == ["log_rate", "temp", "n_exp", "fc_enc.weight", "fc_dec.weight"]
    for k in ("log_rate", "fc_enc.weight", "fc_dec.weight"):
        assert np.array_equal(sd[k].numpy(), g["init_" + k]), k
    assert int(sd["n_exp"]) == 263
    model.load_state_dict({k: torch.tensor(g["p0_" + k]) for k in sd})
    model = model.cuda()
    model.n_exp.fill_(int(g["n_exp"]))
    model._n_exp_host = int(g["n_exp"])
    model.update_t(float(g["temp"]))
    x = torch.tensor(g["x"], device="cuda")
    set_philox(int(g["seed"]), int(g["offset"]))
    dist, log_dr, z, y = model(x)
    recon = model.loss_recon(y, x)
    kl = model.loss_kl(log_dr)
    loss = torch.mean(recon + float(g["beta"]) * torch.sum(kl, dim=1))  # main/train_vae.py:347-351
    loss.backward()
    # GEMM-dependent quantities inherit the GEMM's precision; see tests/helpers.py for the tolerance
    for name, got in (("log_dr", log_dr), ("z", z), ("y", y), ("recon", recon), ("kl", kl), ("loss", loss),
                      ("rate", dist.rate)):
        assert_rel(got.detach().cpu().numpy(), g[name], rtol=2e-5, what=name)
    for n, p in model.named_parameters():
        assert_rel(p.grad.cpu().numpy(), g["g_" + n], rtol=2e-5, what="grad " + n)
    assert dist.temp.item() == pytest.approx(float(g["temp"]))
    assert dist.n_exp == int(g["n_exp"])


def test_vae_eval_and_prior_sampling():
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    model = PoissonVAE(ConfigPoisVAE(n_latents=128, prior_clamp=-4, seed=1)).cuda()
    with torch.no_grad():
        model.log_rate.add_(4.0)
    x = torch.randn(32, 1, 16, 16, device="cuda")
    dist, log_dr, spks, y = model.xtract_ftr(x, t=0.0)  # hard counts, main/vae.py:417-429
    assert torch.equal(spks, spks.round()) and y.shape == x.shape
    dist, log_dr, spks2, _ = model.xtract_ftr(x, t=0.0, ablate_enc=[0, 1], ablate_spks=[2])
    assert (log_dr[:, :2] == 0).all() and (spks2[:, 2] == 0).all()
    xs, s = model.sample(16, t=0.0)  # main/vae.py:431-444
    assert xs.shape == (16, 1, 16, 16) and s.shape == (16, 128) and torch.equal(s, s.round())
    xs, s = model.sample(16, t=1.0)
    assert (s >= 0).all()
    model.update_n(20.0)
    assert int(model.n_exp) == 42


def test_exact_method_matches_reference(golden_dir):
    """method='exact' (closed-form reconstruction loss, main/vae.py:74-93) against the reference's values and autograd."""
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    g = np.load(os.path.join(golden_dir, "vae_exact_k64.npz"))
    model = PoissonVAE(ConfigPoisVAE(n_latents=64, prior_clamp=-4, seed=int(g["cfg_seed"])))
    model.load_state_dict({k: torch.tensor(g["p0_" + k]) for k in model.state_dict()})
    model = model.cuda()
    x = torch.tensor(g["x"], device="cuda")
    recon, dist, log_dr = model.loss_recon_exact(x)
    kl = model.loss_kl(log_dr)
    loss = torch.mean(recon + float(g["beta"]) * torch.sum(kl, dim=1))
    loss.backward()
    for name, got in (("recon", recon), ("kl", kl), ("loss", loss), ("rate", dist.rate), ("log_dr", log_dr)):
        assert_rel(got.detach().cpu().numpy(), g[name], rtol=2e-5, what=name)
    assert dist.mean is dist.rate and dist.variance is dist.rate
    for n, p in model.named_parameters():
        assert_rel(p.grad.cpu().numpy(), g["g_" + n], rtol=2e-5, what="grad " + n)
