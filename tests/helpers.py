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
