"""Shared tolerances for parity tests.

FP32_RTOL is the tolerance BASELINE.json's north_star states for floating point: relaxed
samples, KL, ELBO and gradients within 1e-5 relative.  "Relative" is taken per tensor:
|a - b| <= FP32_RTOL * |b| + FP32_RTOL * max|b|  (the second term is the absolute floor for
entries that are many orders of magnitude below the tensor's scale, e.g. relaxed counts of
latents whose first arrival is far beyond t = 1).
"""
import numpy as np

FP32_RTOL = 1e-5


def assert_rel(a, b, rtol=FP32_RTOL, floor=None, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = np.max(np.abs(b)) if b.size else 0.0
    atol = rtol * scale if floor is None else floor
    err = np.abs(a - b) - (rtol * np.abs(b) + atol)
    worst = np.max(err) if err.size else -1.0
    assert worst <= 0, f"{what}: max excess {worst:.3e}, max abs diff {np.max(np.abs(a - b)):.3e}, scale {scale:.3e}"


def assert_ulp(a, b, max_ulp=1, max_frac=1.0, what=""):
    """fp32 arrays equal up to `max_ulp` units in the last place; at most `max_frac` of the entries may differ at all."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    d = np.abs(ia - ib)
    assert d.max(initial=0) <= max_ulp, f"{what}: {d.max()} ulp apart at {np.unravel_index(d.argmax(), d.shape)}"
    frac = float((d > 0).mean()) if d.size else 0.0
    assert frac <= max_frac, f"{what}: {frac:.2e} of the entries differ"


# ---- conv encoder fixtures (tests/golden/conv_enc_*.npz, oracle/make_golden.py:gen_conv) -----------------
def digest_stride(size: int) -> int:
    return 1 if size <= 4096 else max(8, -(-size // 16384))


def conv_fixture_model(g):
    """The mirror model with the parameters the fixture's reference model had: same cfg.seed (-> bit-identical
    initial parameters), then the same seeded perturbation of lognorm / biases / LayerNorm affine and log_rate + 3."""
    import torch
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    cfg_vae, _ = default_configs(str(g["dataset"]), "poisson", str(g["archi"]) if "archi" in g else "conv+b|lin")
    cfg_vae.update(seed=int(g["cfg_seed"]), n_latents=512)
    model = PoissonVAE(ConfigPoisVAE(**cfg_vae)).eval()
    gen = torch.Generator().manual_seed(int(g["perturb_seed"]))
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("lognorm") or name.endswith("bias") or "layer_norm" in name:
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
        model.log_rate.add_(3.0)
    for name, p in model.named_parameters():  # the reference's parameters, by checksum
        ck = np.array([p.detach().double().sum().item(), p.detach().double().abs().sum().item()])
        assert np.allclose(ck, g["ck_" + name], rtol=1e-12, atol=0), name
    return model


def assert_grad_digest(grad, g, name, rtol):
    flat = np.asarray(grad, dtype=np.float64).reshape(-1)
    assert_rel(flat[::digest_stride(flat.size)], g["g_" + name], rtol=rtol, what="grad " + name)
    nrm = float(g["gn_" + name])
    assert abs(np.linalg.norm(flat) - nrm) <= 10 * rtol * nrm + 1e-30, ("grad norm " + name, np.linalg.norm(flat), nrm)
