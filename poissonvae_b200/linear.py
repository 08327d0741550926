"""Dense contractions of the lin|lin P-VAE (fc_enc main/vae.py:51, fc_dec main/vae.py:59-60 via
base/common.py:409-414) and of their backward (dgrad / wgrad), on the tcgen05 / TMEM / TMA GEMM
(csrc/gemm_tc.cu, pvae_gemm_tf32 in include/pvae_b200.h).  fp32 tensors in HBM, no transposed copies."""
from __future__ import annotations

import torch

from . import _lib, rng

# 3xTF32 operand splitting (fp32-grade products).  False = single-pass TF32 (10-bit operands).
PRECISE = True

_ws_cache = {}


def _splits_for(m: int, n: int, kd: int) -> int:
    """Split the reduction when the output has too few 128x128 tiles to occupy the 148 SMs (wgrad)."""
    tiles = ((m + 127) // 128) * ((n + 127) // 128)
    if tiles >= 64 or kd < 512:
        return 1
    return max(1, min(128 // tiles, kd // 256))


def _gemm(a, b, out, m, n, kd, a_kmajor, b_kmajor):
    for t, name in ((a, "a"), (b, "b"), (out, "out")):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"gemm operand {name} must be a contiguous CUDA float32 tensor (no CPU path)")
    lib = _lib.load()
    splits = _splits_for(m, n, kd)
    ws = None
    if splits > 1:
        key = (a.device, m, n, splits)
        ws = _ws_cache.get(key)
        if ws is None:
            ws = _ws_cache[key] = torch.empty(lib.pvae_gemm_ws_floats(m, n, splits), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        _lib.check(lib.pvae_gemm_tf32(a.data_ptr(), b.data_ptr(), out.data_ptr(), m, n, kd, a_kmajor, b_kmajor,
                                      int(PRECISE), splits, None if ws is None else ws.data_ptr(),
                                      rng.launch_stream()))
    return out


def gemm_nt(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """out (M,N) = a (M,K) @ b (N,K)^T  - linear forward"""
    m, kd = a.shape
    n = b.shape[0]
    out = torch.empty(m, n, device=a.device, dtype=torch.float32) if out is None else out
    return _gemm(a, b, out, m, n, kd, 1, 1)


def gemm_nn(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """out (M,N) = a (M,K) @ b (K,N)  - dgrad"""
    m, kd = a.shape
    n = b.shape[1]
    out = torch.empty(m, n, device=a.device, dtype=torch.float32) if out is None else out
    return _gemm(a, b, out, m, n, kd, 1, 0)


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """out (M,N) = a (K,M)^T @ b (K,N)  - wgrad"""
    kd, m = a.shape
    n = b.shape[1]
    out = torch.empty(m, n, device=a.device, dtype=torch.float32) if out is None else out
    return _gemm(a, b, out, m, n, kd, 0, 0)


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """F.linear for 2-D inputs (base/common.py:409-414), through the registered custom op torch.ops.pvae.linear."""
    from . import ops  # noqa: F401  (registers torch.ops.pvae.*)
    return torch.ops.pvae.linear(x, weight, bias)
