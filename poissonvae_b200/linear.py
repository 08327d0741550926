"""Dense contractions of the lin|lin P-VAE (fc_enc main/vae.py:51, fc_dec main/vae.py:59-60 via
base/common.py:409-414)."""
from __future__ import annotations

import torch
from torch.nn import functional as F


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    # TODO(round 1): route through the tcgen05 GEMM in csrc/gemm_tc.cu
    return F.linear(x, weight, bias)
