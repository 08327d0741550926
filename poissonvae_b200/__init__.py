"""poissonvae_b200 - B200-native (sm_100a) training hot path of the Poisson VAE.

Host-side mirror of the reference interface (hadivafaii/PoissonVAE: base/distributions.py Poisson,
main/vae.py PoissonVAE) over the C-ABI in include/pvae_b200.h.  CUDA only; no CPU fallback.
"""
from . import _lib  # noqa: F401
from .distributions import Poisson, compute_n_exp, softclamp, softclamp_upper  # noqa: F401

__all__ = ["Poisson", "compute_n_exp", "softclamp", "softclamp_upper"]
