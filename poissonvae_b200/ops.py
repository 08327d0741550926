"""torch custom-op layer (`torch.ops.pvae.*`) over the C-ABI of include/pvae_b200.h - the drop-in boundary SURVEY.md
section 8(b) names: schema, fake (meta) implementations and registered autograd formulas, so that the ops are visible to
`torch.library.opcheck`, fake tensors and `torch.compile` tracing.  The kernels are reached through the same ctypes
binding as everything else (no second native layer); the closed-form backward of SURVEY 8 row a13 is registered with
`register_autograd` and is itself made of ops of this library.

    pvae::poisson_rsample      Poisson(log_rate).rsample()            base/distributions.py:55-81
    pvae::encode_sample_kl     softclamps + rate + rsample (+ KL sums)  main/vae.py:393-415, 446-450
    pvae::linear               F.linear in Linear.forward              base/common.py:409-414
    pvae::gemm                 the tcgen05 GEMM the backward formulas are written in

Random draws: `seed < 0` (the default) takes (seed, offset) from torch's CUDA generator INSIDE the op and advances its
offset, so `torch.manual_seed` (called by every reference Config constructor, base/config_base.py:51-63) governs the
stream; the pair used is returned as the CPU int64 tensor `philox`, which the backward reads to address the same
stream.  An explicit (seed, offset) makes the op a pure function of its arguments (tests, opcheck, replay).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib, rng

_INDICATORS = {"sigmoid": 0, "cosine": 1, "linear": 2, "cubic": 3, "quintic": 4}


def _ptr(t):
    return None if t is None else t.data_ptr()


def _check_f32(t: Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: poissonvae_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")


def _rows(t: Tensor):
    """(B, K) view of a contiguous (..., K) tensor for the kernels (the Philox stream is keyed by the flat index)."""
    n = t.numel()
    if n % 512 == 0:
        return n // 512, 512
    k = t.shape[-1] if t.dim() > 0 else 1
    return n // max(k, 1), k


def _philox(device, seed: int, offset: int):
    if seed < 0:
        seed, offset = rng.next_philox(device)
    return int(seed), int(offset)


# ---- pvae::gemm ------------------------------------------------------------------------------------------------
@torch.library.custom_op("pvae::gemm", mutates_args=(), device_types="cuda")
def gemm(a: Tensor, b: Tensor, a_kmajor: bool, b_kmajor: bool) -> Tensor:
    """C (M, N) = A' B'^T; a is (M, Kd) if a_kmajor else (Kd, M); b is (N, Kd) if b_kmajor else (Kd, N)."""
    from . import linear as L
    _check_f32(a, "a"), _check_f32(b, "b")
    a, b = a.contiguous(), b.contiguous()
    m, kd = (a.shape[0], a.shape[1]) if a_kmajor else (a.shape[1], a.shape[0])
    n = b.shape[0] if b_kmajor else b.shape[1]
    out = torch.empty(m, n, device=a.device, dtype=torch.float32)
    return L._gemm(a, b, out, m, n, kd, int(a_kmajor), int(b_kmajor))


@gemm.register_fake
def _(a, b, a_kmajor, b_kmajor):
    m = a.shape[0] if a_kmajor else a.shape[1]
    n = b.shape[0] if b_kmajor else b.shape[1]
    return a.new_empty(m, n)


# ---- pvae::linear ----------------------------------------------------------------------------------------------
@torch.library.custom_op("pvae::linear", mutates_args=(), device_types="cuda")
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """F.linear for 2-D inputs, base/common.py:409-414: y = x W^T (+ b)."""
    y = gemm(x, weight, True, True)
    return y if bias is None else y.add_(bias)


@linear.register_fake
def _(x, weight, bias=None):
    return x.new_empty(x.shape[0], weight.shape[0])


def _linear_setup(ctx, inputs, output):
    x, weight, bias = inputs
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None


def _linear_backward(ctx, g):
    x, weight = ctx.saved_tensors
    g = g.contiguous()
    gx = gemm(g, weight, True, False) if ctx.needs_input_grad[0] else None   # dgrad  g W
    gw = gemm(g, x, False, False) if ctx.needs_input_grad[1] else None        # wgrad  g^T x
    gb = g.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
    return gx, gw, gb


linear.register_autograd(_linear_backward, setup_context=_linear_setup)


# ---- pvae::poisson_rsample ---------------------------------------------------------------------------------------
@torch.library.custom_op("pvae::poisson_rsample", mutates_args=(), device_types="cuda")
def poisson_rsample(log_rate: Tensor, temp: float, n_exp: int, indicator: str = "sigmoid", clamp: float = -1.0,
                    exit_logit: float = 0.0, seed: int = -1, offset: int = -1) -> Tuple[Tensor, Tensor, Tensor]:
    """Relaxed (temp > 0) or hard (temp == 0) Poisson counts of rate exp(softclamp_upper(log_rate, clamp)) + eps
    (clamp < 0: no clamp).  Returns (z, dz_acc, philox): dz_acc = sum_n f'(l_n) t_n feeds the backward (zeros at
    temp == 0), philox = the (seed, offset) the draws came from."""
    _check_f32(log_rate, "log_rate")
    assert temp >= 0.0, f"must be non-neg: {temp}"  # base/distributions.py:15
    assert indicator in _INDICATORS  # base/distributions.py:16
    lr = log_rate.contiguous()
    B, K = _rows(lr)
    z, dz = torch.empty_like(lr), torch.zeros_like(lr) if temp == 0.0 else torch.empty_like(lr)
    seed, offset = _philox(lr.device, seed, offset)
    with torch.cuda.device(lr.device):
        _lib.check(_lib.load().pvae_rsample_fwd(_ptr(lr), _ptr(z), None, None if temp == 0.0 else _ptr(dz), B, K, 0, n_exp, temp,
                                                _INDICATORS[indicator], clamp, exit_logit, seed, offset, None,
                                                rng.launch_stream()))
    return z, dz, torch.tensor([seed, offset], dtype=torch.int64)


@poisson_rsample.register_fake
def _(log_rate, temp, n_exp, indicator="sigmoid", clamp=-1.0, exit_logit=0.0, seed=-1, offset=-1):
    return torch.empty_like(log_rate), torch.empty_like(log_rate), torch.empty(2, dtype=torch.int64, device="cpu")


@torch.library.custom_op("pvae::poisson_rsample_backward", mutates_args=(), device_types="cuda")
def poisson_rsample_backward(g_z: Tensor, log_rate: Tensor, dz_acc: Tensor, philox: Tensor, temp: float, n_exp: int,
                             indicator: str, clamp: float, exit_logit: float) -> Tensor:
    """d loss / d log_rate from g_z and the forward's dz_acc (closed form)."""
    lr, g_z, dz = log_rate.contiguous(), g_z.contiguous(), dz_acc.contiguous()
    B, K = _rows(lr)
    g = torch.empty_like(lr)
    seed, offset = (int(v) for v in philox.tolist())
    with torch.cuda.device(lr.device):
        _lib.check(_lib.load().pvae_rsample_bwd(_ptr(lr), _ptr(g_z), _ptr(dz), _ptr(g), B, K, 0, n_exp, temp,
                                                _INDICATORS[indicator], clamp, exit_logit, seed, offset, None,
                                                rng.launch_stream()))
    return g


@poisson_rsample_backward.register_fake
def _(g_z, log_rate, dz_acc, philox, temp, n_exp, indicator, clamp, exit_logit):
    return torch.empty_like(log_rate)


def _rsample_setup(ctx, inputs, output):
    log_rate, temp, n_exp, indicator, clamp, exit_logit, _, _ = inputs
    _, dz, philox = output
    ctx.save_for_backward(log_rate, dz, philox)
    ctx.args = (temp, n_exp, indicator, clamp, exit_logit)


def _rsample_backward(ctx, g_z, g_dz, g_philox):
    log_rate, dz, philox = ctx.saved_tensors
    temp, n_exp, indicator, clamp, exit_logit = ctx.args
    if temp == 0.0:
        raise RuntimeError("hard counts (temp == 0) are not differentiable")
    g = poisson_rsample_backward(g_z, log_rate, dz, philox, temp, n_exp, indicator, clamp, exit_logit)
    return g, None, None, None, None, None, None, None


poisson_rsample.register_autograd(_rsample_backward, setup_context=_rsample_setup)


# ---- pvae::encode_sample_kl --------------------------------------------------------------------------------------
@torch.library.custom_op("pvae::encode_sample_kl", mutates_args=(), device_types="cuda")
def encode_sample_kl(h: Tensor, log_r: Tensor, b_enc: Optional[Tensor], temp: float, n_exp: int, indicator: str = "sigmoid",
                     exc_only: bool = False, exit_logit: float = 0.0, seed: int = -1, offset: int = -1,
                     row_offset: int = 0) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """PoissonVAE.infer + rsample + the KL's per-sample sums in one kernel (main/vae.py:393-415, 446-450):
    log_dr = softclamp_upper(h + b_enc, 10) (softclamp(., 10, 0) if exc_only), rate = exp(softclamp_upper(log_r + log_dr, 5))
    + eps, z = rsample(rate), kl_rowsum[b] = sum_k exp(log_r)(1 + e^{log_dr}(log_dr - 1)).
    Returns (z, log_dr, kl_rowsum, dz_acc, philox, rate_max).  row_offset: global index of the first row (data parallel:
    the Philox stream is keyed by the global sample index); rate_max: 1 float, the largest rate of this call (what
    `update_n` needs, main/vae.py:452-459) - an output, not a mutated argument: the op is functional."""
    _check_f32(h, "h")
    assert temp >= 0.0, f"must be non-neg: {temp}"
    assert indicator in _INDICATORS
    h = h.contiguous()
    B, K = h.shape
    lr = log_r.reshape(-1).contiguous()
    z, log_dr = torch.empty_like(h), torch.empty_like(h)
    dz = torch.zeros_like(h) if temp == 0.0 else torch.empty_like(h)
    klrow = torch.empty(B, device=h.device, dtype=torch.float32)
    rate_max = torch.zeros(1, device=h.device, dtype=torch.float32)
    seed, offset = _philox(h.device, seed, offset)
    with torch.cuda.device(h.device):
        _lib.check(_lib.load().pvae_sampler_fwd(
            _ptr(h), _ptr(lr), _ptr(b_enc), _ptr(z), _ptr(log_dr), None, None, _ptr(klrow), _ptr(rate_max),
            None if temp == 0.0 else _ptr(dz), B, K, row_offset, n_exp, temp, _INDICATORS[indicator], int(exc_only), exit_logit,
            seed, offset, None, rng.launch_stream()))
    return z, log_dr, klrow, dz, torch.tensor([seed, offset], dtype=torch.int64), rate_max


@encode_sample_kl.register_fake
def _(h, log_r, b_enc, temp, n_exp, indicator="sigmoid", exc_only=False, exit_logit=0.0, seed=-1, offset=-1, row_offset=0):
    return (torch.empty_like(h), torch.empty_like(h), h.new_empty(h.shape[0]), torch.empty_like(h),
            torch.empty(2, dtype=torch.int64, device="cpu"), h.new_empty(1))


@torch.library.custom_op("pvae::encode_sample_kl_backward", mutates_args=(), device_types="cuda")
def encode_sample_kl_backward(g_z: Tensor, g_log_dr: Optional[Tensor], g_kl_rowsum: Optional[Tensor], h: Tensor, log_r: Tensor,
                              b_enc: Optional[Tensor], log_dr: Tensor, dz_acc: Tensor, philox: Tensor, temp: float, n_exp: int,
                              indicator: str, exc_only: bool, exit_logit: float, row_offset: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(g_h, g_log_r, g_b_enc) - g_b_enc is an empty tensor when there is no bias.  The KL's upstream gradient (one value
    per sample) goes through the elementwise KL backward kernel into g_log_dr / g_log_r first."""
    lib = _lib.load()
    h, g_z = h.contiguous(), g_z.contiguous()
    B, K = h.shape
    lr = log_r.reshape(-1).contiguous()
    s = rng.launch_stream()
    g_ld = None if g_log_dr is None else g_log_dr.contiguous()
    g_lr_kl = None
    with torch.cuda.device(h.device):
        ws = torch.empty(2 * lib.pvae_num_partials(B) * K, device=h.device, dtype=torch.float32)
        if g_kl_rowsum is not None:
            g_kl = g_kl_rowsum.reshape(B, 1).expand(B, K).contiguous()
            g_ld_kl, g_lr_kl = torch.empty_like(h), torch.empty_like(lr)
            _lib.check(lib.pvae_kl_bwd(_ptr(log_dr.contiguous()), _ptr(lr), _ptr(g_kl), _ptr(g_ld_kl), _ptr(g_lr_kl), _ptr(ws), 0, B, K, s))
            g_ld = g_ld_kl if g_ld is None else g_ld + g_ld_kl
        g_h, g_lr = torch.empty_like(h), torch.empty_like(lr)
        g_b = torch.empty_like(b_enc) if b_enc is not None else h.new_empty(0)
        seed, offset = (int(v) for v in philox.tolist())
        _lib.check(lib.pvae_sampler_bwd(
            _ptr(h), _ptr(lr), _ptr(b_enc), _ptr(g_z), _ptr(g_ld), _ptr(dz_acc.contiguous()), 0.0, _ptr(g_h), _ptr(g_lr),
            _ptr(g_b) if b_enc is not None else None, _ptr(ws), 0, B, K, row_offset, n_exp, temp, _INDICATORS[indicator],
            int(exc_only), exit_logit, seed, offset, None, s))
    if g_lr_kl is not None:
        g_lr = g_lr + g_lr_kl
    return g_h, g_lr.reshape(log_r.shape), g_b


@encode_sample_kl_backward.register_fake
def _(g_z, g_log_dr, g_kl_rowsum, h, log_r, b_enc, log_dr, dz_acc, philox, temp, n_exp, indicator, exc_only, exit_logit, row_offset):
    return torch.empty_like(h), torch.empty_like(log_r), (torch.empty_like(b_enc) if b_enc is not None else h.new_empty(0))


def _encode_setup(ctx, inputs, output):
    h, log_r, b_enc, temp, n_exp, indicator, exc_only, exit_logit, _, _, row_offset = inputs
    _, log_dr, _, dz, philox, _ = output
    ctx.has_bias = b_enc is not None
    ctx.save_for_backward(h, log_r, b_enc, log_dr, dz, philox)
    ctx.args = (temp, n_exp, indicator, exc_only, exit_logit, row_offset)


def _encode_backward(ctx, g_z, g_log_dr, g_kl_rowsum, g_dz, g_philox, g_rate_max):
    h, log_r, b_enc, log_dr, dz, philox = ctx.saved_tensors
    temp, n_exp, indicator, exc_only, exit_logit, row_offset = ctx.args
    if temp == 0.0:
        raise RuntimeError("hard counts (temp == 0) are not differentiable")
    if g_z is None:
        g_z = torch.zeros_like(h)
    g_h, g_lr, g_b = encode_sample_kl_backward(g_z, g_log_dr, g_kl_rowsum, h, log_r, b_enc, log_dr, dz, philox, temp, n_exp,
                                               indicator, exc_only, exit_logit, row_offset)
    return g_h, g_lr, (g_b if ctx.has_bias else None), None, None, None, None, None, None, None, None


encode_sample_kl.register_autograd(_encode_backward, setup_context=_encode_setup)
