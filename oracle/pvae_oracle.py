"""CPU oracle for the P-VAE training hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain numpy restatement of the reference algorithm
(hadivafaii/PoissonVAE) for the path named in BASELINE.json.  It is imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, and
only as the checker.  Nothing under poissonvae_b200/ may import it.

Pinning status: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the
UNMODIFIED reference itself, imported in the build container by
oracle/make_golden.py and committed as tests/golden/*.npz; the Philox
generator is pinned against the published Random123 known-answer vectors.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  Arithmetic notes that matter for bit-exactness:

* inter-event time  x = e / rate      true IEEE division   (base/distributions.py:60,
                                      torch Exponential.rsample = exponential_() / rate)
* arrival time      t_j = t_{j-1}+x_j sequential fp32 adds (base/distributions.py:63, cumsum dim 0)
* hard count        #{j < n_exp : t_j <= 1}                (restatement of steps (1)-(2); the
                                      reference's temp==0 branch is torch.poisson, a different
                                      algorithm, base/distributions.py:56-57,83-85)
"""
from __future__ import annotations

import numpy as np

F32_EPS = np.float32(np.finfo(np.float32).eps)  # base/distributions.py:24

INDICATORS = ("sigmoid", "cosine", "linear", "cubic", "quintic")  # base/distributions.py:294-300

# ----------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 philox.h).  The CUDA
# kernels key it as  key = (seed_lo ^ DOMAIN, seed_hi),
# ctr = (latent_index >> 2, event, offset_lo, offset_hi); word latent_index & 3 of
# the block is that latent's draw for that event.  See include/pvae_b200.h "RNG contract".
# ----------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
PVAE_KEY_DOMAIN = 0x50564145  # "PVAE"


def philox4x32_10(ctr: np.ndarray, key) -> np.ndarray:
    """ctr: (..., 4) uint32, key: 2 ints. Returns (..., 4) uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    sh = np.uint64(32)
    for _ in range(10):
        p0 = PHILOX_M0 * c[0]
        p1 = PHILOX_M1 * c[2]
        hi0, lo0 = p0 >> sh, p0 & mask
        hi1, lo1 = p1 >> sh, p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def uniform_from_bits(bits: np.ndarray) -> np.ndarray:
    """u = ((bits >> 9) + 0.5) * 2^-23, exact in fp32, in [2^-24, 1 - 2^-24]."""
    return ((bits >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def draw_uniforms(seed: int, offset: int, latent_index: np.ndarray, n_exp: int) -> np.ndarray:
    """The (n_exp, *latent_index.shape) uniforms the kernels consume (include/pvae_b200.h "RNG
    contract"): block (latent_index >> 2, event) word latent_index & 3."""
    idx = np.asarray(latent_index, dtype=np.uint64)
    ctr = np.zeros((n_exp,) + idx.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = (idx >> np.uint64(2)).astype(np.uint32)[None]
    ctr[..., 1] = np.arange(n_exp, dtype=np.uint32).reshape((n_exp,) + (1,) * idx.ndim)
    ctr[..., 2] = np.uint32(offset & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32((offset >> 32) & 0xFFFFFFFF)
    key = ((seed & 0xFFFFFFFF) ^ PVAE_KEY_DOMAIN, (seed >> 32) & 0xFFFFFFFF)
    bits = philox4x32_10(ctr, key)  # (n_exp, ..., 4)
    word = np.broadcast_to((idx & np.uint64(3)).astype(np.int64)[None, ..., None], bits.shape[:-1] + (1,))
    return uniform_from_bits(np.take_along_axis(bits, word, axis=-1)[..., 0])


def exponential_from_uniform(u: np.ndarray) -> np.ndarray:
    """e = -log(u): torch's CUDA exponential transform (ATen TransformationHelper.h:129-146)."""
    return (-np.log(u.astype(np.float32))).astype(np.float32)


# ----------------------------------------------------------------------------
# soft clamps (base/distributions.py:233-242); F.softplus beta=1, threshold=20
# ----------------------------------------------------------------------------
def softplus(x):
    x = np.asarray(x)
    with np.errstate(over="ignore"):
        return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 30)))).astype(x.dtype)


def softplus_grad(x):
    x = np.asarray(x)
    with np.errstate(over="ignore"):
        z = np.exp(np.minimum(x, 30))
    return np.where(x > 20, np.ones_like(x), z / (z + 1)).astype(x.dtype)


def softclamp_upper(x, c):  # base/distributions.py:241-242
    x = np.asarray(x)
    c = x.dtype.type(c)
    return c - softplus(c - x)


def softclamp_upper_grad(x, c):
    x = np.asarray(x)
    return softplus_grad(x.dtype.type(c) - x)


def softclamp(x, upper, lower=0.0):  # base/distributions.py:233-234
    x = np.asarray(x)
    up, lo = x.dtype.type(upper), x.dtype.type(lower)
    return lo + softplus(x - lo) - softplus(x - up)


def softclamp_grad(x, upper, lower=0.0):
    x = np.asarray(x)
    return softplus_grad(x - x.dtype.type(lower)) - softplus_grad(x - x.dtype.type(upper))


# ----------------------------------------------------------------------------
# soft indicators (base/distributions.py:245-300) and their derivatives
# ----------------------------------------------------------------------------
def _unit(l):
    return np.clip(l.dtype.type(0.5) * l + l.dtype.type(0.5), 0, 1)


def indicator(l: np.ndarray, kind: str) -> np.ndarray:
    if kind == "sigmoid":
        with np.errstate(over="ignore"):
            return (1 / (1 + np.exp(-l))).astype(l.dtype)
    u = _unit(l)
    if kind == "linear":
        return u
    if kind == "cubic":
        return 3 * u ** 2 - 2 * u ** 3
    if kind == "quintic":
        return 6 * u ** 5 - 15 * u ** 4 + 10 * u ** 3
    if kind == "cosine":
        return (0.5 * (1.0 - np.cos(np.pi * u))).astype(l.dtype)
    raise ValueError(kind)


def indicator_grad(l: np.ndarray, kind: str) -> np.ndarray:
    """d indicator / d logit."""
    if kind == "sigmoid":
        s = indicator(l, kind)
        return s * (1 - s)
    u = _unit(l)
    inside = ((l > -1) & (l < 1)).astype(l.dtype)
    half = l.dtype.type(0.5)
    if kind == "linear":
        return half * inside
    if kind == "cubic":
        return half * inside * (6 * u - 6 * u ** 2)
    if kind == "quintic":
        return half * inside * (30 * u ** 4 - 60 * u ** 3 + 30 * u ** 2)
    if kind == "cosine":
        return (half * inside * (0.5 * np.pi * np.sin(np.pi * u))).astype(l.dtype)
    raise ValueError(kind)


# ----------------------------------------------------------------------------
# Poisson sampler (base/distributions.py:5-92)
# ----------------------------------------------------------------------------
def rate_from_log_rate(log_rate):  # base/distributions.py:24-25
    log_rate = np.asarray(log_rate)
    return (np.exp(log_rate) + log_rate.dtype.type(F32_EPS)).astype(log_rate.dtype)


def arrival_times(e: np.ndarray, rate: np.ndarray, acc=None) -> np.ndarray:
    """steps (1)-(2), base/distributions.py:60-63: cumsum over events of e/rate.

    acc: accumulator dtype of the running sum.  torch.cumsum keeps an fp32 running sum on CUDA
    (ATen ScanUtils.cuh, acc_type<float, true> = float) - that is the canonical arithmetic of the
    kernels and the default here - but a DOUBLE running sum rounded to fp32 at every step on CPU
    (ATen ReduceOpsKernel.cpp cumsum_cpu_kernel, acc_type<float, false> = double).  Fixtures made
    by running the reference on CPU are therefore checked with acc=np.float64.
    """
    x = e / rate[None]
    return np.cumsum(x, axis=0, dtype=e.dtype if acc is None else acc).astype(e.dtype)


def rsample(e: np.ndarray, rate: np.ndarray, temp: float, kind: str = "sigmoid", acc=None) -> np.ndarray:
    """Relaxed spike count, base/distributions.py:55-81. e: (n_exp, ...) Exp(1) draws."""
    t = arrival_times(e, rate, acc)
    logits = (1 - t) / e.dtype.type(temp)  # :70
    return indicator(logits, kind).sum(0, dtype=e.dtype)  # :73-79


def rsample_grad_log_rate(e, rate, temp, kind="sigmoid"):
    """d z / d log_rate, SURVEY.md section 8 row a13 (verified there against reference autograd)."""
    t = arrival_times(e, rate)
    logits = (1 - t) / e.dtype.type(temp)
    acc = (indicator_grad(logits, kind) * t).sum(0, dtype=e.dtype)
    return acc / e.dtype.type(temp) * (rate - e.dtype.type(F32_EPS)) / rate


def hard_count(e: np.ndarray, rate: np.ndarray, acc=None) -> np.ndarray:
    """#arrivals in [0, 1]: hard-threshold restatement of base/distributions.py:60-63."""
    return (arrival_times(e, rate, acc) <= 1).sum(0).astype(e.dtype)


def log_prob(rate, samples):  # base/distributions.py:87-92
    from scipy.special import gammaln
    return -rate - gammaln(samples + 1) + samples * np.log(rate)


def compute_n_exp(rate: float, p: float = 1e-5) -> int:
    """base/distributions.py:226-230; PoissonVAE.update_n uses p=1e-5 (main/vae.py:452-459)."""
    from scipy import stats
    assert rate > 0.0
    return int(stats.poisson(rate).ppf(1.0 - p))


# ----------------------------------------------------------------------------
# P-VAE lin|lin forward / loss / closed-form backward
#   main/vae.py:384-415 (forward, infer), :54-72 (decode, loss_recon),
#   :446-450 (loss_kl); main/train_vae.py:346-359 (loss assembly)
# ----------------------------------------------------------------------------
def infer(h, log_r, exc_only=False):
    """h = fc_enc(x) (B,K); log_r (K,). Returns log_dr, u, lam, rate  (main/vae.py:393-415)."""
    log_dr = softclamp(h, 10.0, 0.0) if exc_only else softclamp_upper(h, 10.0)
    u = log_r[None] + log_dr
    lam = softclamp_upper(u, 5.0)
    return log_dr, u, lam, rate_from_log_rate(lam)


def loss_kl(log_r, log_dr):  # main/vae.py:446-450
    f = 1 + np.exp(log_dr) * (log_dr - 1)
    return np.exp(log_r)[None] * f


def vae_step(x, w_enc, w_dec, log_r, e, temp, beta, kind="sigmoid", exc_only=False,
             b_enc=None, need_grads=True, global_batch=None):
    """One forward (+ closed-form backward) of the lin|lin P-VAE given Exp(1) draws e (n_exp,B,K).

    x: (B, D) flattened patches. Returns a dict with h, log_dr, rate, z, y, recon, kl, loss and,
    if need_grads, g_w_enc, g_w_dec, g_log_r (and g_b_enc).  global_batch: the batch size the
    mean in main/train_vae.py:351 runs over (data-parallel shards pass the global B).
    """
    dt = x.dtype
    B = x.shape[0]
    Bg = B if global_batch is None else global_batch
    h = x @ w_enc.T  # main/vae.py:51 via base/common.py:409-414
    if b_enc is not None:
        h = h + b_enc[None]
    log_dr, u, lam, rate = infer(h, log_r, exc_only)
    z = rsample(e, rate, temp, kind)
    y = z @ w_dec.T  # main/vae.py:59-60
    resid = y - x
    recon = (resid ** 2).sum(1)  # main/vae.py:68-72
    kl = loss_kl(log_r, log_dr)
    loss = (recon + dt.type(beta) * kl.sum(1)).sum() / dt.type(Bg)  # main/train_vae.py:347-351
    out = dict(h=h, log_dr=log_dr, lam=lam, rate=rate, z=z, y=y, recon=recon, kl=kl, loss=loss)
    if not need_grads:
        return out
    g_y = 2 * resid / dt.type(Bg)
    g_w_dec = g_y.T @ z
    g_z = g_y @ w_dec
    dz = rsample_grad_log_rate(e, rate, temp, kind)
    g_u = g_z * dz * softclamp_upper_grad(u, 5.0)
    scale = dt.type(beta) / dt.type(Bg)
    g_ldr = g_u + scale * np.exp(log_r)[None] * log_dr * np.exp(log_dr)
    dclamp = softclamp_grad(h, 10.0, 0.0) if exc_only else softclamp_upper_grad(h, 10.0)
    g_h = g_ldr * dclamp
    out.update(g_z=g_z, dz_dlam=dz, g_h=g_h, g_w_enc=g_h.T @ x, g_w_dec=g_w_dec,
               g_log_r=(g_u + scale * kl).sum(0))
    if b_enc is not None:
        out["g_b_enc"] = g_h.sum(0)
    return out


# ----------------------------------------------------------------------------
# Adamax "fast" (base/adamax.py:34-100) on a flat parameter vector
# ----------------------------------------------------------------------------
def adamax_step(p, g, exp_avg, exp_inf, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    dt = p.dtype
    if weight_decay != 0:
        g = g + dt.type(weight_decay) * p
    exp_avg = exp_avg * dt.type(beta1) + dt.type(1 - beta1) * g
    exp_inf = np.maximum(exp_inf * dt.type(beta2), np.abs(g) + dt.type(eps))
    clr = lr / (1 - beta1 ** step)
    p = p - dt.type(clr) * exp_avg / exp_inf
    return p, exp_avg, exp_inf


# ----------------------------------------------------------------------------
# schedules (base/utils_model.py:6-14, 57-68) - inputs to the step
# ----------------------------------------------------------------------------
def temp_anneal_linear(n_iters, t0=1.0, t1=0.1, portion=0.7):
    temps = np.ones(n_iters) * t1
    i = int(np.ceil(portion * n_iters))
    temps[:i] = np.linspace(t0, t1, i)
    return temps


def beta_anneal_linear(n_iters, beta=1.0, anneal_portion=0.3, constant_portion=0.0, min_beta=1e-4):
    betas = np.ones(n_iters) * beta
    a = int(np.ceil(constant_portion * n_iters))
    b = int(np.ceil((constant_portion + anneal_portion) * n_iters))
    betas[:a] = min_beta
    betas[a:b] = np.linspace(min_beta, beta, b - a)
    return betas
