"""CPU: the conv encoder's oracle (oracle/conv_oracle.py) and the mirror's parameter construction against fixtures of
the UNMODIFIED reference (tests/golden/conv_enc_*.npz, oracle/make_golden.py:gen_conv): CIFAR16-shaped conv+b|lin
(BASELINE config 4), the MNIST-shaped 28x28 encoder in front of a linear decoder, and conv+b|conv+b on MNIST-shaped 28x28
(BASELINE config 5) and CIFAR16-shaped inputs."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import conv_oracle, pvae_oracle as po
from tests.helpers import assert_grad_digest, assert_rel, conv_fixture_model

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "conv_enc_*.npz")) +
                glob.glob(os.path.join(os.path.dirname(__file__), "golden", "conv_vae_*.npz")))


@pytest.mark.parametrize("path", GOLDEN)
def test_conv_oracle_matches_reference(path):
    g = np.load(path)
    model = conv_fixture_model(g)  # also checks the mirror's parameters against the reference's checksums
    assert list(g["param_names"]) == [n for n, _ in model.named_parameters()]
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in model.state_dict().items()}
    x = torch.tensor(g["x"])
    torch.set_num_threads(4)
    h = conv_oracle.conv_encoder(x, sd, str(g["dataset"]))
    assert_rel(h.detach().numpy(), g["enc_out"], rtol=2e-6, what="encoder output")
    # the rest of the step with the numpy oracle's sampler arithmetic restated in torch (autograd for the gradients)
    B, K = h.shape
    idx = np.arange(B * K, dtype=np.uint32).reshape(B, K)
    e = torch.tensor(po.exponential_from_uniform(po.draw_uniforms(int(g["philox_seed"]), int(g["philox_offset"]), idx, int(g["n_exp"]))))
    sc = lambda v, c: c - torch.nn.functional.softplus(c - v)
    log_dr = sc(h, 10.0)
    rate = torch.exp(sc(sd["log_rate"] + log_dr, 5.0)) + torch.finfo(torch.float32).eps
    z = torch.sigmoid((1 - torch.cumsum(e / rate, 0)) / float(g["temp"])).sum(0)
    if "conv+b|conv" in str(g["archi"]):
        y = conv_oracle.conv_decoder(z, sd, str(g["dataset"]))
    else:
        y = (z @ sd["fc_dec.weight"].T).view(x.shape)
    recon = ((y - x) ** 2).sum(dim=[1, 2, 3])
    kl = torch.exp(sd["log_rate"]) * (1 + torch.exp(log_dr) * (log_dr - 1))
    loss = torch.mean(recon + float(g["beta"]) * kl.sum(1))
    loss.backward()
    for name, got in (("log_dr", log_dr), ("z", z), ("y", y), ("recon", recon), ("kl", kl), ("loss", loss), ("rate", rate)):
        assert_rel(got.detach().numpy(), g[name], rtol=1e-5, what=name)
    for name in g["param_names"]:
        assert_grad_digest(sd[str(name)].grad.numpy(), g, str(name), 1e-5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/base"), reason="reference checkout not present")
@pytest.mark.parametrize("dataset,archi", [("CIFAR16", "conv+b|lin"), ("MNIST", "conv+b|lin"), ("MNIST", "conv+b|conv+b"),
                                           ("CIFAR16", "conv+b|conv+b"), ("vH16", "conv|conv")])
def test_mirror_parameters_equal_reference(dataset, archi):
    from oracle import ref_import
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    ref = ref_import.load()
    kws, tr = ref.cfg.default_configs(dataset, "poisson", archi)
    kws2, tr2 = default_configs(dataset, "poisson", archi)
    assert {k: kws2[k] for k in kws} == kws and tr == tr2
    a = ref.vae.PoissonVAE(ref.cfg.ConfigPoisVAE(save=False, seed=3, **kws)).state_dict()
    b = PoissonVAE(ConfigPoisVAE(seed=3, **kws2)).state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
