"""Import the UNMODIFIED reference (hadivafaii/PoissonVAE) in the build container.

TEST INFRASTRUCTURE ONLY - used by oracle/make_golden.py (fixture generation) and
tests that are skipped when /root/reference is absent (it never exists on the GPU box).

The reference wires everything through `from x import *`, so importing the sampler
pulls the plotting stack (SURVEY.md section 1 "import chain gotcha").  Plotting-only
modules missing from this image are replaced by MagicMock; nothing on the hot path
touches them.  `wandb` must be imported before the fake IPython is installed.
"""
from __future__ import annotations

import contextlib
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("PVAE_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.backends",
    "matplotlib.backends.backend_pdf", "matplotlib.gridspec", "matplotlib.patches",
    "mpl_toolkits", "mpl_toolkits.axes_grid1", "seaborn", "IPython", "IPython.display",
    "statsmodels",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "base"))


def load():
    """Returns a namespace with the reference symbols the hot path needs."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    try:
        import wandb  # noqa: F401  (before the IPython stub, see module docstring)
    except Exception:
        sys.modules.setdefault("wandb", MagicMock())
    for m in _STUBS:
        try:
            __import__(m)
        except Exception:
            sys.modules[m] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import types
    ns = types.SimpleNamespace()
    from base import distributions as ref_dist
    from base import utils_model as ref_utils
    from base import adamax as ref_adamax
    from main import vae as ref_vae
    from main import config_vae as ref_cfg
    ns.dist, ns.utils, ns.adamax, ns.vae, ns.cfg = ref_dist, ref_utils, ref_adamax, ref_vae, ref_cfg
    return ns


@contextlib.contextmanager
def inject_exponentials(draws):
    """Make every `Tensor.exponential_()` inside the block return the next array of `draws`.

    torch's Exponential.rsample is `rate.new(shape).exponential_() / rate`
    (torch/distributions/exponential.py), so the unmodified reference then consumes exactly
    the Exp(1) draws we hand it (SURVEY.md section 8c "uniform injection").
    """
    import torch
    queue = list(draws)
    orig = torch.Tensor.exponential_

    def fake(self, lambd=1, *, generator=None):
        e = queue.pop(0)
        e = torch.as_tensor(e, dtype=self.dtype)
        assert tuple(e.shape) == tuple(self.shape), (e.shape, self.shape)
        return self.copy_(e)

    torch.Tensor.exponential_ = fake
    try:
        yield
    finally:
        torch.Tensor.exponential_ = orig
