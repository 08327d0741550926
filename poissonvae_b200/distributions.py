"""Host-side mirror of the reference `Poisson` distribution (base/distributions.py:5-92).

Same constructor, attributes and methods as the reference class; `rsample()` and `sample()` run the
hand-written sm_100a kernels through the C-ABI (include/pvae_b200.h) instead of materialising the
(n_exp, B, K) tensors.  CUDA fp32 only - there is no CPU path.
"""
from __future__ import annotations

import torch
from torch.nn import functional as F

from . import _lib
from . import rng
from .rng import next_philox

INDICATOR_CODES = {"sigmoid": 0, "cosine": 1, "linear": 2, "cubic": 3, "quintic": 4}  # base/distributions.py:294-300
_INDICATOR_FNS = INDICATOR_CODES  # same name as the reference's registry, for `in` checks

EXIT_LOSSLESS = 0.0
EXIT_NEVER = -1.0


def compute_n_exp(rate: float, p: float = 1e-6) -> int:
    """base/distributions.py:226-230 (host-side, scipy)."""
    from scipy import stats as sp_stats
    assert rate > 0.0, f"must be positive, got: {rate}"
    return int(sp_stats.poisson(rate).ppf(1.0 - p))


def softclamp(x: torch.Tensor, upper: float, lower: float = 0.0):
    """base/distributions.py:233-234."""
    return lower + F.softplus(x - lower) - F.softplus(x - upper)


def softclamp_upper(x: torch.Tensor, c: float):
    """base/distributions.py:241-242."""
    return c - F.softplus(c - x)


def _ptr(t):
    return None if t is None else t.data_ptr()


_stream = rng.launch_stream


def _as_rows(t: torch.Tensor):
    """View a contiguous (..., K) tensor as (B, K) for the kernels.  The Philox stream is keyed by
    the flat element index, so any factorisation gives the same draws; prefer 512-wide rows."""
    n = t.numel()
    if n % 512 == 0:
        return n // 512, 512
    k = t.shape[-1] if t.dim() > 0 else 1
    return n // max(k, 1), k


def _require_cuda_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: poissonvae_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")


class Poisson:
    """Drop-in for the reference class: same signature (base/distributions.py:6-14).

    Extra keyword `exit_logit` (not in the reference) selects the early-exit policy of the chain:
    EXIT_LOSSLESS (default) leaves a latent's chain only where the reference's fp32 terms are
    exactly zero; EXIT_NEVER walks all n_exp events; c > 0 leaves once every logit < -c.
    """

    def __init__(
            self,
            log_rate: torch.Tensor,
            temp: float = 0.0,
            clamp: float | None = None,
            indicator_approx: str = 'sigmoid',
            n_exp: int | str = 'infer',
            n_exp_p: float = 1e-3,
            exit_logit: float = EXIT_LOSSLESS,
    ):
        assert temp >= 0.0, f"must be non-neg: {temp}"
        assert indicator_approx in _INDICATOR_FNS
        _require_cuda_f32(log_rate, "log_rate")
        self.indicator_approx = indicator_approx
        self.temp = temp
        self.clamp = clamp
        self.exit_logit = float(exit_logit)
        self._log_rate_in = log_rate
        self._rate = None
        if n_exp == 'infer':
            n_exp = self._infer_n_exp(n_exp_p)
        self.n_exp = int(n_exp)

    def __repr__(self):
        parts = [f"rate: {tuple(self._log_rate_in.shape)}", f"temp: {self.temp}", f"n_exp: {self.n_exp}"]
        return f"Poisson({', '.join(parts)})"

    # rate = exp(log_rate) + eps, base/distributions.py:21-25 (computed on first use)
    @property
    def rate(self):
        if self._rate is None:
            lr = self._log_rate_in
            if self.clamp is not None:
                lr = softclamp_upper(lr, self.clamp)
            self._rate = torch.exp(lr) + torch.finfo(torch.float32).eps
        return self._rate

    @torch.no_grad()
    def _infer_n_exp(self, n_exp_p):
        return int(compute_n_exp(self.rate.max().item(), n_exp_p))

    @property
    def mean(self):
        return self.rate

    @property
    def variance(self):
        return self.rate

    def _launch_args(self):
        clamp = -1.0 if self.clamp is None else float(self.clamp)
        return int(self.n_exp), INDICATOR_CODES[self.indicator_approx], clamp

    def rsample(self):
        temp = float(self.temp)
        if temp == 0.0:
            return self.sample()
        n_exp, _, clamp = self._launch_args()
        # the registered custom op (poissonvae_b200/ops.py): draws its Philox (seed, offset) from torch's CUDA generator
        # inside the op, closed-form backward registered with torch.library
        from . import ops  # noqa: F401  (registers torch.ops.pvae.*)
        z, _, _ = torch.ops.pvae.poisson_rsample(self._log_rate_in, temp, n_exp, self.indicator_approx, clamp, self.exit_logit)
        return z

    @torch.no_grad()
    def sample(self):
        """Hard spike counts: #arrivals of the same exponential chain in [0, 1], at most n_exp
        (the reference draws torch.poisson here, base/distributions.py:83-85: same law up to the
        n_exp truncation, different stream)."""
        n_exp, ind, clamp = self._launch_args()
        seed, offset = next_philox(self._log_rate_in.device)
        lr = self._log_rate_in.detach().contiguous()
        B, K = _as_rows(lr)
        z = torch.empty_like(lr)
        with torch.cuda.device(lr.device):
            _lib.check(_lib.load().pvae_rsample_fwd(_ptr(lr), _ptr(z), None, None, B, K, 0, n_exp, 0.0, ind, clamp,
                                                    self.exit_logit, seed, offset, None, _stream()))
        return z

    def log_prob(self, samples: torch.Tensor):
        """base/distributions.py:87-92."""
        return -self.rate - torch.lgamma(samples + 1) + samples * torch.log(self.rate)
