"""CPU: the torch-CPU port used as the timed CPU baseline (oracle/ref_port.py) is the reference's step:
against fixtures from the unmodified reference, and bit-for-bit against the reference itself when
its checkout is present (build container only)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_import, ref_port
from tests.helpers import assert_rel

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "vae_step_*.npz")))


@pytest.mark.parametrize("path", GOLDEN)
def test_port_matches_fixture(path):
    g = np.load(path)
    params = dict(log_rate=torch.tensor(g["p0_log_rate"], requires_grad=True),
                  w_enc=torch.tensor(g["p0_fc_enc.weight"], requires_grad=True),
                  w_dec=torch.tensor(g["p0_fc_dec.weight"], requires_grad=True))
    x = torch.tensor(g["x"])
    opt = ref_port.Adamax(params, lr=float(g["lr"]))
    with ref_import.inject_exponentials([g["e"]]):
        loss, aux = ref_port.forward_loss(params, x, int(g["n_exp"]), float(g["temp"]), float(g["beta"]),
                                          exc_only=bool(g["exc_only"]))
    loss.backward()
    opt.t, _ = 0, None
    assert np.array_equal(loss.detach().numpy(), g["loss"])  # same ops, same machine arithmetic
    assert np.array_equal(aux["z"].detach().numpy(), g["z"])
    for k, gk in (("log_rate", "g_log_rate"), ("w_enc", "g_fc_enc.weight"), ("w_dec", "g_fc_dec.weight")):
        assert_rel(params[k].grad.numpy(), g[gk], rtol=1e-6, what=gk)
    opt.step()
    for k, pk in (("log_rate", "p1_log_rate"), ("w_enc", "p1_fc_enc.weight"), ("w_dec", "p1_fc_dec.weight")):
        assert_rel(params[k].detach().numpy(), g[pk], rtol=1e-6, what=pk)


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout only exists in the build container")
def test_port_is_bit_identical_to_reference_under_same_seed():
    ref = ref_import.load()
    cfg_vae, cfg_tr = ref.cfg.default_configs("vH16", "poisson", "lin|lin")
    cfg_vae.update(n_latents=64, seed=7)
    model = ref.vae.PoissonVAE(ref.cfg.ConfigPoisVAE(save=False, **cfg_vae))
    model.n_exp.fill_(12)
    params = dict(log_rate=model.log_rate.detach().clone().requires_grad_(),
                  w_enc=model.fc_enc.weight.detach().clone().requires_grad_(),
                  w_dec=model.fc_dec.weight.detach().clone().requires_grad_())
    x = torch.randn(10, 1, 16, 16, generator=torch.Generator().manual_seed(99))
    opt_ref = ref.adamax.Adamax(model.parameters(), lr=0.005)
    opt = ref_port.Adamax(params, lr=0.005)
    for step in range(3):
        torch.manual_seed(1234 + step)
        opt_ref.zero_grad(set_to_none=True)
        dist, log_dr, z, y = model(x)
        loss_ref = torch.mean(model.loss_recon(y, x) + 2.5 * torch.sum(model.loss_kl(log_dr), dim=1))
        loss_ref.backward()
        opt_ref.step()
        torch.manual_seed(1234 + step)
        loss = ref_port.train_step(params, opt, x, 12, temp=1.0, beta=2.5)
        assert torch.equal(loss, loss_ref.detach())
    assert torch.equal(params["w_enc"].detach(), model.fc_enc.weight.detach())
    assert torch.equal(params["w_dec"].detach(), model.fc_dec.weight.detach())
    assert torch.equal(params["log_rate"].detach(), model.log_rate.detach())
