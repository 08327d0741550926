"""CPU: the numpy oracle against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py) and against the published Philox4x32-10 known answers."""
import glob
import os

import numpy as np
import pytest

from oracle import pvae_oracle as po
from tests.helpers import assert_rel


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kats = [
        ([0, 0, 0, 0], (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
        ([0xFFFFFFFF] * 4, (0xFFFFFFFF, 0xFFFFFFFF), "408f276d 41c83b0e a20bc7c6 6d5451fd"),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], (0xA4093822, 0x299F31D0),
         "d16cfe09 94fdcceb 5001e420 24126ea1"),
    ]
    for ctr, key, want in kats:
        got = po.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert " ".join(f"{v:08x}" for v in got) == want


def test_uniform_range_exact():
    bits = np.array([0, 1, 511, 512, 0xFFFFFFFF, 0x80000000], dtype=np.uint32)
    u = po.uniform_from_bits(bits)
    assert u.dtype == np.float32
    assert u.min() == np.float32(2.0 ** -24) and u.max() == np.float32(1 - 2.0 ** -24)
    assert np.all(po.exponential_from_uniform(u) > 0)


@pytest.mark.parametrize("kind", po.INDICATORS)
def test_sampler_matches_reference(golden_dir, kind):
    g = np.load(os.path.join(golden_dir, f"sampler_{kind}.npz"))
    # the fixture's draws are the ones the oracle's generator produces
    idx = np.arange(g["log_rate"].size, dtype=np.uint32).reshape(g["log_rate"].shape)
    u = po.draw_uniforms(int(g["seed"]), int(g["offset"]), idx, int(g["n_exp"]))
    assert np.array_equal(u, g["u"])
    rate = po.rate_from_log_rate(g["log_rate"])
    assert_rel(rate, g["rate"], rtol=3e-7, what="rate")
    for temp in (1.0, 0.3, 0.05):
        z = po.rsample(g["e"], g["rate"], temp, kind)
        assert_rel(z, g[f"z_t{temp:g}"], what=f"z {kind} {temp}")
        dz = po.rsample_grad_log_rate(g["e"], g["rate"], temp, kind)
        assert_rel(dz * g["w"], g[f"g_t{temp:g}"], what=f"grad {kind} {temp}")


def test_hard_counts_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "hard_counts.npz"))
    # the reference ran on CPU here: torch's CPU cumsum keeps a double running sum
    t = po.arrival_times(g["e"], g["rate"], acc=np.float64)
    assert np.array_equal(t, g["times"])  # IEEE div + sequential add: bit-exact
    c = po.hard_count(g["e"], g["rate"], acc=np.float64)
    assert np.array_equal(c, g["counts"])
    assert c.max() > 10 and c.min() == 0
    # the CUDA-semantics (fp32 running sum) variant differs only in the last bits of t
    t32 = po.arrival_times(g["e"], g["rate"])
    assert np.max(np.abs(t32 - t) / t) < 2e-6


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "vae_step_*.npz"))))
def test_vae_step_matches_reference(path):
    g = np.load(path)
    x = g["x"].reshape(g["x"].shape[0], -1)
    out = po.vae_step(x, g["p0_fc_enc.weight"], g["p0_fc_dec.weight"], g["p0_log_rate"][0], g["e"],
                      float(g["temp"]), float(g["beta"]), exc_only=bool(g["exc_only"]))
    for k in ("log_dr", "z", "recon", "kl", "loss"):
        assert_rel(out[k], g[k], what=k)
    assert_rel(out["y"], g["y"].reshape(x.shape), what="y")
    assert_rel(out["rate"], g["rate"], what="rate")
    assert_rel(out["g_w_enc"], g["g_fc_enc.weight"], what="g_w_enc")
    assert_rel(out["g_w_dec"], g["g_fc_dec.weight"], what="g_w_dec")
    assert_rel(out["g_log_r"], g["g_log_rate"][0], what="g_log_r")
    # one Adamax step (base/adamax.py:34-100), step=1
    for name, gname in (("fc_enc.weight", "g_w_enc"), ("fc_dec.weight", "g_w_dec"), ("log_rate", "g_log_r")):
        p0 = g["p0_" + name]
        grad = out[gname].reshape(p0.shape)
        p1, _, _ = po.adamax_step(p0, grad, np.zeros_like(p0), np.zeros_like(p0), 1, float(g["lr"]))
        # first step moves every weight by lr*sign(g); compare where the gradient is not ~0
        assert_rel(p1, g["p1_" + name], rtol=1e-6, what="adamax " + name)


def test_scalars(golden_dir):
    g = np.load(os.path.join(golden_dir, "scalars.npz"))
    for r, n5, n3 in zip(g["n_exp_rates"], g["n_exp_p1e-5"], g["n_exp_p1e-3"]):
        assert po.compute_n_exp(float(r), 1e-5) == n5
        assert po.compute_n_exp(float(r), 1e-3) == n3
    assert po.compute_n_exp(200.0) == 263  # main/vae.py:382
    x = g["clamp_x"]
    assert_rel(po.softclamp_upper(x, 10.0), g["softclamp_upper_10"], rtol=1e-6, what="sc10")
    assert_rel(po.softclamp_upper(x, 5.0), g["softclamp_upper_5"], rtol=1e-6, what="sc5")
    assert_rel(po.softclamp(x, 10.0, 0.0), g["softclamp_10_0"], rtol=1e-6, what="sc10_0")
    for kind in po.INDICATORS:
        assert_rel(po.indicator(g["ind_l"], kind), g[f"ind_{kind}"], rtol=1e-6, what=kind)
    assert np.array_equal(po.temp_anneal_linear(1000, 1.0, 0.05, 0.5), g["temp_lin_1000"])
    assert np.array_equal(po.beta_anneal_linear(1000, 2.5, 0.5, 0.0, 1e-4), g["beta_lin_1000"])


def test_indicator_grads_numeric():
    l = np.linspace(-2.5, 2.5, 1001, dtype=np.float64)
    h = 1e-6
    for kind in po.INDICATORS:
        num = (po.indicator(l + h, kind) - po.indicator(l - h, kind)) / (2 * h)
        ana = po.indicator_grad(l, kind)
        mask = (np.abs(np.abs(l) - 1) > 1e-3)  # kinks of the finite-support indicators
        assert np.max(np.abs(num - ana)[mask]) < 1e-6, kind


def test_lr_schedule_matches_reference_optimizer(golden_dir):
    """trainer.lr_schedule vs the reference Adamax under torch's CosineAnnealingLR driven like the reference loop
    (oracle/make_golden.py:gen_lr): warm-up, the hand-over step and the recursive cosine."""
    from poissonvae_b200.trainer import lr_schedule
    g = np.load(os.path.join(golden_dir, "lr_schedule.npz"))
    for tag in "abcde":
        lr, n_iters, portion, eta_min = g[f"cfg_{tag}"]
        want = g[f"lr_{tag}"]
        got = np.array([lr_schedule(i, int(n_iters), lr, portion, 'cosine', eta_min) for i in range(int(n_iters))])
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-18, err_msg=tag)
