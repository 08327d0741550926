"""Dense contractions of the lin|lin P-VAE (fc_enc main/vae.py:51, fc_dec main/vae.py:59-60 via
base/common.py:409-414) and of their backward (dgrad / wgrad)."""
from __future__ import annotations

import torch
from torch.nn import functional as F

# TODO(round 1): route these through the tcgen05 GEMM in csrc/gemm_tc.cu


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    return F.linear(x, weight, bias)


def gemm_nt(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (M,N) = a (M,K) @ b (N,K)^T"""
    return torch.matmul(a, b.t(), out=out)


def gemm_nn(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (M,N) = a (M,K) @ b (K,N)"""
    return torch.matmul(a, b, out=out)


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out (M,N) = a (K,M)^T @ b (K,N)"""
    return torch.matmul(a.t(), b, out=out)
