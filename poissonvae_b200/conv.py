"""Host-side mirror of the reference's conv encoder / decoder blocks:
`Conv2D`, `ConvLayer`, `SELayer`, `FactorizedReduce`, `StrideReduce`, `Cell` (base/common.py:75-340),
`ResConvPool`, `ResDenseLayer`, `_build_conv_enc` (main/vae.py:657-756) and the stem (main/vae.py:211-218).

Same module tree, constructor order (-> same initial parameters under the same seed) and state_dict keys
as the reference.  Every block is ONE autograd Function whose forward and hand-derived backward enqueue the
sm_100a kernels of csrc/conv_tc.cu (implicit-GEMM convolutions on tcgen05) and csrc/conv_misc.cu through the
C-ABI; activations are NHWC fp32 between blocks.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from . import _lib
from .linear import gemm_nn, gemm_nt, gemm_tn
from . import rng
from .rng import next_philox

MULT = 2
PRECISE = True  # 3xTF32 operand splitting in the convolutions (fp32-grade); False = single-pass TF32
# True: a cell's forward / backward is ONE C call that enqueues all of its kernels (csrc/conv_net.cu).  False: the same
# kernels in the same order, one ctypes call each, from the Functions below (bit-identical results; kept as the readable
# statement of the algorithm and as the cross-check of the native path).
NATIVE_CELLS = True


def _p(t):
    return None if t is None else t.data_ptr()


_s = rng.launch_stream


def _ints(v):
    return (C.c_int * max(1, len(v)))(*v)


_ws = {}


def _workspace(dev, floats: int) -> torch.Tensor:
    """Scratch for the fixed-order reductions (stream-ordered reuse: every consumer runs before the next producer)."""
    t = _ws.get(dev)
    if t is None or t.numel() < floats:
        t = _ws[dev] = torch.empty(max(floats, 1 << 20), device=dev, dtype=torch.float32)
    return t


# ------------------------------------------------------------------------------------------------------
# thin wrappers over the C entry points
# ------------------------------------------------------------------------------------------------------
def out_size(i: int, k: int, stride: int, pad: int) -> int:
    return (i + 2 * pad - k) // stride + 1


def pad32(c: int) -> int:
    """Channel count in memory: the tensor-core kernels work on multiples of 32 (16-channel decoder layers are zero-padded)."""
    return (c + 31) // 32 * 32


def weight_prep(weight, lognorm, bias, ntaps: int, groups: int = 1, need_dgrad: bool = True, with_lo: bool = True):
    """-> (w_fwd [co][tap][ci], w_dgrad [ci][tap][co], bias) of the weight-normalised filter (base/common.py:426-431),
    channel counts zero-padded to multiples of 32.  with_lo: each packed matrix has twice the rows - the second half holds
    the TF32 low parts x - trunc_tf32(x) of the first, which the convolution kernel then loads instead of computing."""
    co, ci = weight.shape[0], weight.shape[1]
    cop, cip = pad32(co), pad32(ci)
    rep = 2 if with_lo else 1
    w_fwd = torch.empty(rep * cop, ntaps * cip, device=weight.device, dtype=torch.float32)
    w_dg = torch.empty(rep * cip, ntaps * cop, device=weight.device, dtype=torch.float32) if need_dgrad else None
    b_pad = torch.empty(cop, device=weight.device, dtype=torch.float32) if bias is not None else None
    with torch.cuda.device(weight.device):
        _lib.check(_lib.load().pvae_conv_weight_prep(_p(weight), _p(lognorm), _p(bias), _p(w_fwd), _p(w_dg), _p(b_pad), co, ci, cop,
                                                     cip, ntaps, groups, int(with_lo), _s()))
    return w_fwd, w_dg, b_pad


def weight_grad(g_wp, weight, lognorm, ntaps: int, groups: int = 1):
    co, ci = weight.shape[0], weight.shape[1]
    g_w = torch.empty_like(weight)
    g_ln = torch.empty_like(lognorm) if lognorm is not None else None
    with torch.cuda.device(weight.device):
        _lib.check(_lib.load().pvae_conv_weight_grad(_p(g_wp), _p(weight), _p(lognorm), _p(g_w), _p(g_ln), co, ci, ntaps, groups,
                                                     _s()))
    return g_w, g_ln


def _igemm(inp, w, out, out_act, bias, add_pre, dsilu_src, add_post, oh, ow, out_step, off_h, off_w, stride, taps, wtaps):
    n, ih, iw, ci = inp.shape
    _, ohf, owf, co = out.shape
    dy, dx, wi = _ints([t[0] for t in taps]), _ints([t[1] for t in taps]), _ints([t[2] for t in taps])
    with torch.cuda.device(inp.device):
        _lib.check(_lib.load().pvae_conv_igemm(_p(inp), _p(w), _p(out), _p(out_act), _p(bias), _p(add_pre), _p(dsilu_src),
                                               _p(add_post), n, ih, iw, ci, co, oh, ow, ohf, owf, out_step, off_h, off_w, stride,
                                               len(taps), dy, dx, wi, wtaps, int(w.shape[0] == 2 * co), int(PRECISE), _s()))


def conv_fwd(x_act, w_fwd, bias, k: int, stride: int, pad: int, co: int, want_act: bool = False):
    """x_act (N,IH,IW,Ci) NHWC -> conv + bias (N,OH,OW,Co) [, silu of it]"""
    n, ih, iw, ci = x_act.shape
    oh, ow = out_size(ih, k, stride, pad), out_size(iw, k, stride, pad)
    out = torch.empty(n, oh, ow, co, device=x_act.device, dtype=torch.float32)
    act = torch.empty_like(out) if want_act else None
    taps = [(r - pad, s - pad, r * k + s) for r in range(k) for s in range(k)]
    _igemm(x_act, w_fwd, out, act, bias, None, None, None, oh, ow, 1, 0, 0, stride, taps, k * k)
    return out, act


def conv_dgrad(g_out, w_dgrad, in_shape, k: int, stride: int, pad: int, add_pre=None, dsilu_src=None, add_post=None):
    """gradient wrt the convolution's (activated) input, with the fused epilogue
    g_in = (conv^T(g_out) + add_pre) * silu'(dsilu_src) + add_post"""
    n, ih, iw, ci = in_shape
    _, oh, ow, co = g_out.shape
    g_in = torch.empty(n, ih, iw, ci, device=g_out.device, dtype=torch.float32)
    if stride == 1:
        taps = [(pad - r, pad - s, r * k + s) for r in range(k) for s in range(k)]
        _igemm(g_out, w_dgrad, g_in, None, None, add_pre, dsilu_src, add_post, ih, iw, 1, 0, 0, 1, taps, k * k)
        return g_in
    assert stride == 2
    for ph in range(2):  # one launch per parity class of the input grid
        for pw in range(2):
            qh, qw = (ih - ph + 1) // 2, (iw - pw + 1) // 2
            if qh <= 0 or qw <= 0:
                continue
            taps = [((ph + pad - r) // 2, (pw + pad - s) // 2, r * k + s) for r in range(k) for s in range(k)
                    if (ph + pad - r) % 2 == 0 and (pw + pad - s) % 2 == 0]
            _igemm(g_out, w_dgrad, g_in, None, None, add_pre, dsilu_src, add_post, qh, qw, 2, ph, pw, 1, taps, k * k)
    return g_in


def conv_wgrad(x_act, g_out, weight, lognorm, k: int, stride: int, pad: int, groups: int = 1):
    """-> (g_weight, g_lognorm, g_bias): weight gradient on the tensor cores, reduced and pushed through the weight
    normalisation; the bias gradient rides along as an all-ones operand slab."""
    n, ih, iw, ci = x_act.shape
    _, oh, ow, co = g_out.shape
    lib = _lib.load()
    g_w = torch.empty_like(weight)
    g_ln = torch.empty_like(lognorm) if lognorm is not None else None
    g_b = torch.empty(co, device=x_act.device, dtype=torch.float32)
    ws = _workspace(x_act.device, lib.pvae_conv_wgrad_ws_floats(ci, co, k * k))
    taps = [(r - pad, s - pad, r * k + s) for r in range(k) for s in range(k)]
    dy, dx, wi = _ints([t[0] for t in taps]), _ints([t[1] for t in taps]), _ints([t[2] for t in taps])
    with torch.cuda.device(x_act.device):
        _lib.check(lib.pvae_conv_wgrad(_p(x_act), _p(g_out), _p(weight), _p(lognorm), _p(g_w), _p(g_ln), _p(g_b), _p(ws), n, ih, iw,
                                       ci, co, oh, ow, stride, len(taps), dy, dx, wi, groups, weight.shape[1], weight.shape[0],
                                       int(PRECISE), _s()))
    return g_w, g_ln, g_b[:weight.shape[0]]


def colsum(x2d):
    r, c = x2d.shape
    lib = _lib.load()
    out = torch.empty(c, device=x2d.device, dtype=torch.float32)
    ws = _workspace(x2d.device, lib.pvae_reduce_ws_floats(c))
    with torch.cuda.device(x2d.device):
        _lib.check(lib.pvae_colsum(_p(x2d), _p(out), _p(ws), r, c, _s()))
    return out


def small_tn(a, b):
    """(R,Ca)^T (R,Cb) -> (Ca,Cb) for tiny Ca*Cb (squeeze-excitation weights)"""
    r, ca = a.shape
    cb = b.shape[1]
    lib = _lib.load()
    out = torch.empty(ca, cb, device=a.device, dtype=torch.float32)
    ws = _workspace(a.device, lib.pvae_reduce_ws_floats(ca * cb))
    with torch.cuda.device(a.device):
        _lib.check(lib.pvae_small_tn(_p(a), _p(b), _p(out), _p(ws), r, ca, cb, _s()))
    return out


def silu(x):
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pvae_silu_fwd(_p(x), _p(out), x.numel(), _s()))
    return out


def upsample(x, oh: int, ow: int):
    """nearest-neighbour resampling of an NHWC map (nn.Upsample(mode='nearest'))"""
    n, ih, iw, c = x.shape
    out = torch.empty(n, oh, ow, c, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pvae_upsample_fwd(_p(x), _p(out), n, ih, iw, oh, ow, c, _s()))
    return out


def upsample_bwd(g_out, ih: int, iw: int, dsilu_src=None):
    n, oh, ow, c = g_out.shape
    g_in = torch.empty(n, ih, iw, c, device=g_out.device, dtype=torch.float32)
    with torch.cuda.device(g_out.device):
        _lib.check(_lib.load().pvae_upsample_bwd(_p(g_out), _p(dsilu_src), _p(g_in), n, ih, iw, oh, ow, c, _s()))
    return g_in


def _check_input(x, name="x"):
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a CUDA float32 tensor (poissonvae_b200 has no CPU path)")
    return x.contiguous()


# ------------------------------------------------------------------------------------------------------
# autograd Functions: one per block
# ------------------------------------------------------------------------------------------------------
class _StemFn(torch.autograd.Function):
    """x (N,1,H,W) -> (y, silu(y)) NHWC, main/vae.py:211-218 (x carries no gradient)."""

    @staticmethod
    def forward(ctx, x, weight, lognorm, bias, pad):
        x = _check_input(x)
        n, _, ih, iw = x.shape
        co = weight.shape[0]
        oh, ow = ih + 2 * pad - 2, iw + 2 * pad - 2
        y = torch.empty(n, oh, ow, co, device=x.device, dtype=torch.float32)
        y_act = torch.empty_like(y)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pvae_stem_fwd(_p(x), _p(weight), _p(lognorm), _p(bias), _p(y), _p(y_act), n, ih, iw, co, pad, _s()))
        ctx.save_for_backward(x, weight, lognorm)
        ctx.pad = pad
        ctx.mark_non_differentiable(y_act)
        return y, y_act

    @staticmethod
    def backward(ctx, g_y, _g_act):
        x, weight, lognorm = ctx.saved_tensors
        n, _, ih, iw = x.shape
        co = weight.shape[0]
        lib = _lib.load()
        g_y = g_y.contiguous()
        g_wpb = torch.empty(10, co, device=x.device, dtype=torch.float32)
        ws = _workspace(x.device, lib.pvae_reduce_ws_floats(10 * co))
        with torch.cuda.device(x.device):
            _lib.check(lib.pvae_stem_wgrad(_p(x), _p(g_y), _p(g_wpb), _p(ws), n, ih, iw, co, ctx.pad, _s()))
        g_w, g_ln = weight_grad(g_wpb, weight, lognorm, 9)
        return None, g_w, g_ln, g_wpb[9], None


def _se_forward(c2, skip, se_w1, se_b1, se_w2, se_b2, eps_res, skip_up=False):
    n, h, w, ch = c2.shape
    hd, cr = se_w1.shape[0], se_w2.shape[0]
    y = torch.empty_like(c2)
    y_act = torch.empty_like(c2)
    pool = torch.empty(n, cr, device=c2.device, dtype=torch.float32)
    hid = torch.empty(n, hd, device=c2.device, dtype=torch.float32)
    gate = torch.empty(n, cr, device=c2.device, dtype=torch.float32)
    with torch.cuda.device(c2.device):
        _lib.check(_lib.load().pvae_se_fwd(_p(c2), _p(skip), _p(se_w1), _p(se_b1), _p(se_w2), _p(se_b2), eps_res, _p(y), _p(y_act),
                                           _p(pool), _p(hid), _p(gate), n, h, w, ch, cr, hd, int(skip_up), _s()))
    return y, y_act, pool, hid, gate


def _se_backward(g_y, c2, gate, hid, pool, se_w1, se_w2, eps_res):
    n, h, w, ch = c2.shape
    hd, cr = se_w1.shape[0], se_w2.shape[0]
    g_c2 = torch.empty_like(c2)
    d2 = torch.empty(n, cr, device=c2.device, dtype=torch.float32)
    d1 = torch.empty(n, hd, device=c2.device, dtype=torch.float32)
    with torch.cuda.device(c2.device):
        _lib.check(_lib.load().pvae_se_bwd(_p(g_y), _p(c2), _p(gate), _p(hid), _p(se_w1), _p(se_w2), eps_res, _p(g_c2), _p(d2),
                                           _p(d1), n, h * w, ch, cr, hd, _s()))
    lib = _lib.load()
    g_w2, g_b2 = torch.empty(cr, hd, device=c2.device, dtype=torch.float32), torch.empty(cr, device=c2.device, dtype=torch.float32)
    g_w1, g_b1 = torch.empty(hd, cr, device=c2.device, dtype=torch.float32), torch.empty(hd, device=c2.device, dtype=torch.float32)
    ws = _workspace(c2.device, lib.pvae_reduce_ws_floats(2 * cr * hd + cr + hd))
    with torch.cuda.device(c2.device):
        _lib.check(lib.pvae_se_param_grads(_p(d2), _p(hid), _p(d1), _p(pool), _p(g_w2), _p(g_b2), _p(g_w1), _p(g_b1), _p(ws), n, cr, hd,
                                           _s()))
    return g_c2, g_w1, g_b1, g_w2, g_b2


# A trainer that keeps all gradients in one flat buffer sets this to {parameter data_ptr: gradient view} around its step:
# the cells' backward then writes those gradients in place and hands autograd None for them (no per-parameter
# allocation, no AccumulateGrad node, nothing to gather).  None: gradients are returned to autograd as usual.
GRAD_SINK = None

_CELL_FIELDS = ("w1", "ln1", "b1", "w2", "ln2", "b2", "se_w1", "se_b1", "se_w2", "se_b2", "skip_w", "skip_ln", "skip_b")


def _native_cell_forward(ctx, kind, x, x_act, params, eps_res):
    """params: the 13 parameter tensors of pvae_cell in _CELL_FIELDS order (None where the cell kind has none)."""
    lib = _lib.load()
    n, h, w, _ = x.shape
    w1 = params[0]
    d = _lib.Cell()
    d.kind, d.N, d.H, d.W, d.Ci, d.Co, d.eps_res, d.precise = kind, n, h, w, w1.shape[1], w1.shape[0], eps_res, int(PRECISE)
    for name, t in zip(_CELL_FIELDS, params):
        setattr(d, name, _p(t))
    own_act = x_act is None
    up, down = kind == 4, kind in (1, 2)
    oh, ow = (2 * h, 2 * w) if up else ((out_size(h, 3, 2, 1), out_size(w, 3, 2, 1)) if down else (h, w))
    saved = torch.empty(lib.pvae_cell_saved_floats(C.byref(d), int(own_act)), device=x.device, dtype=torch.float32)
    y = torch.empty(n, oh, ow, pad32(w1.shape[0]), device=x.device, dtype=torch.float32)
    y_act = torch.empty_like(y)
    with torch.cuda.device(x.device):
        _lib.check(lib.pvae_cell_fwd(C.byref(d), _p(x), _p(x_act), _p(y), _p(y_act), _p(saved), _s()))
    ctx.save_for_backward(x, x_act, saved, *params)
    ctx.native = (kind, eps_res)
    ctx.mark_non_differentiable(y_act)
    return y, y_act


def _native_cell_backward(ctx, g_y):
    lib = _lib.load()
    x, x_act, saved, *params = ctx.saved_tensors
    kind, eps_res = ctx.native
    n, h, w, _ = x.shape
    w1 = params[0]
    d = _lib.Cell()
    d.kind, d.N, d.H, d.W, d.Ci, d.Co, d.eps_res, d.precise = kind, n, h, w, w1.shape[1], w1.shape[0], eps_res, int(PRECISE)
    grads, ret = [], []  # where the kernels write / what autograd gets
    for t in params:
        sink = None if (t is None or GRAD_SINK is None) else GRAD_SINK.get(t.data_ptr())
        g = None if t is None else (sink if sink is not None else torch.empty_like(t))
        grads.append(g)
        ret.append(None if sink is not None else g)
    for name, t, g in zip(_CELL_FIELDS, params, grads):
        setattr(d, name, _p(t))
        setattr(d, "g_" + name, _p(g))
    g_y = g_y.contiguous()
    g_x = torch.empty_like(x)
    scratch = _workspace(x.device, lib.pvae_cell_scratch_floats(C.byref(d)))
    with torch.cuda.device(x.device):
        _lib.check(lib.pvae_cell_bwd(C.byref(d), _p(x), _p(x_act), _p(g_y), _p(g_x), _p(saved), _p(scratch), _s()))
    return g_x, ret


class _CellFn(torch.autograd.Function):
    """Cell.forward base/common.py:187-192 for the encoder cell types 'normal_pre' (stride 1, identity skip) and
    'down_enc' (stride 2, FactorizedReduce or StrideReduce skip), n_nodes = 2, SiLU, squeeze-excitation on.
    Inputs: x and the hint x_act = silu(x) (produced by the previous block's epilogue; carries no gradient)."""

    @staticmethod
    def forward(ctx, x, x_act, w1, ln1, b1, w2, ln2, b2, se_w1, se_b1, se_w2, se_b2, skip_w, skip_ln, skip_b, stride, skip_groups,
                eps_res):
        x = _check_input(x)
        co = w1.shape[0]
        if NATIVE_CELLS:
            kind = 0 if stride == 1 else (1 if skip_groups > 1 else 2)
            return _native_cell_forward(ctx, kind, x, x_act, (w1, ln1, b1, w2, ln2, b2, se_w1, se_b1, se_w2, se_b2, skip_w, skip_ln,
                                                              skip_b), eps_res)
        ctx.native = None
        if x_act is None:
            x_act = silu(x)
        wf1, wd1, bp1 = weight_prep(w1, ln1, b1, 9)
        wf2, wd2, bp2 = weight_prep(w2, ln2, b2, 9)
        c1, a2 = conv_fwd(x_act, wf1, bp1, 3, stride, 1, co, want_act=True)
        c2, _ = conv_fwd(a2, wf2, bp2, 3, 1, 1, co)
        if stride == 1:
            skip, wfs, wds = x, None, None
        else:
            ks = 2 if skip_groups > 1 else 1
            wfs, wds, bps = weight_prep(skip_w, skip_ln, skip_b, ks * ks, skip_groups)
            skip, _ = conv_fwd(x_act, wfs, bps, ks, 2, 0, co)
        y, y_act, pool, hid, gate = _se_forward(c2, skip, se_w1, se_b1, se_w2, se_b2, eps_res)
        ctx.save_for_backward(x, x_act, c1, a2, c2, pool, hid, gate, w1, ln1, w2, ln2, se_w1, se_w2, skip_w, skip_ln, wd1, wd2, wds)
        ctx.cfg = (stride, skip_groups, eps_res)
        ctx.mark_non_differentiable(y_act)
        return y, y_act

    @staticmethod
    def backward(ctx, g_y, _g_act):
        if ctx.native is not None:
            g_x, g = _native_cell_backward(ctx, g_y)
            return (g_x, None, *g, None, None, None)
        x, x_act, c1, a2, c2, pool, hid, gate, w1, ln1, w2, ln2, se_w1, se_w2, skip_w, skip_ln, wd1, wd2, wds = ctx.saved_tensors
        stride, skip_groups, eps_res = ctx.cfg
        g_y = g_y.contiguous()
        n, oh, ow, co = c2.shape
        g_c2, g_sw1, g_sb1, g_sw2, g_sb2 = _se_backward(g_y, c2, gate, hid, pool, se_w1, se_w2, eps_res)
        # second convolution
        g_w2, g_ln2, g_b2 = conv_wgrad(a2, g_c2, w2, ln2, 3, 1, 1)
        g_c1 = conv_dgrad(g_c2, wd2, c1.shape, 3, 1, 1, dsilu_src=c1)
        # first convolution (+ skip branch)
        g_w1, g_ln1, g_b1 = conv_wgrad(x_act, g_c1, w1, ln1, 3, stride, 1)
        if stride == 1:
            g_x = conv_dgrad(g_c1, wd1, x.shape, 3, 1, 1, dsilu_src=x, add_post=g_y)
            g_sw = g_sln = g_sb = None
        else:
            ks = 2 if skip_groups > 1 else 1
            g_sw, g_sln, g_sb = conv_wgrad(x_act, g_y, skip_w, skip_ln, ks, 2, 0, skip_groups)
            g_a = conv_dgrad(g_c1, wd1, x.shape, 3, 2, 1)
            g_x = conv_dgrad(g_y, wds, x.shape, ks, 2, 0, add_pre=g_a, dsilu_src=x)
        return (g_x, None, g_w1, g_ln1, g_b1, g_w2, g_ln2, g_b2, g_sw1, g_sb1, g_sw2, g_sb2, g_sw, g_sln, g_sb, None, None, None)


class _DecCellFn(torch.autograd.Function):
    """Cell.forward base/common.py:187-192 for the decoder cell types, n_nodes = 1: 'normal_dec' (SiLU -> 3x3 conv, identity
    skip) and 'up_dec' (SiLU -> nearest 2x up-sampling -> 3x3 conv; skip = up-sampling -> 1x1 conv, base/common.py:35-47,
    211-217).  The 1x1 skip convolution commutes with the nearest up-sampling, so it runs on the small map and the
    squeeze-excitation kernel reads it through the up-sampling index map."""

    @staticmethod
    def forward(ctx, x, x_act, w, ln, b, se_w1, se_b1, se_w2, se_b2, skip_w, skip_ln, skip_b, up, eps_res):
        x = _check_input(x)
        n, h, wd_, _ = x.shape
        cop = pad32(w.shape[0])
        if NATIVE_CELLS:
            return _native_cell_forward(ctx, 4 if up else 3, x, x_act, (w, ln, b, None, None, None, se_w1, se_b1, se_w2, se_b2, skip_w,
                                                                          skip_ln, skip_b), eps_res)
        ctx.native = None
        if x_act is None:
            x_act = silu(x)
        wf, wd, bp = weight_prep(w, ln, b, 9)
        if up:
            au = upsample(x_act, 2 * h, 2 * wd_)
            c, _ = conv_fwd(au, wf, bp, 3, 1, 1, cop)
            wfs, wds, bps = weight_prep(skip_w, skip_ln, skip_b, 1)
            skip, _ = conv_fwd(x, wfs, bps, 1, 1, 0, cop)
        else:
            au, wds = None, None
            c, _ = conv_fwd(x_act, wf, bp, 3, 1, 1, cop)
            skip = x
        y, y_act, pool, hid, gate = _se_forward(c, skip, se_w1, se_b1, se_w2, se_b2, eps_res, skip_up=bool(up))
        ctx.save_for_backward(x, x_act, au, c, pool, hid, gate, w, ln, se_w1, se_w2, skip_w, skip_ln, wd, wds)
        ctx.cfg = (bool(up), eps_res)
        ctx.mark_non_differentiable(y_act)
        return y, y_act

    @staticmethod
    def backward(ctx, g_y, _g_act):
        if ctx.native is not None:
            g_x, g = _native_cell_backward(ctx, g_y)
            return (g_x, None, g[0], g[1], g[2], g[6], g[7], g[8], g[9], g[10], g[11], g[12], None, None)
        x, x_act, au, c, pool, hid, gate, w, ln, se_w1, se_w2, skip_w, skip_ln, wd, wds = ctx.saved_tensors
        up, eps_res = ctx.cfg
        g_y = g_y.contiguous()
        n, h, wd_, _ = x.shape
        g_c, g_sw1, g_sb1, g_sw2, g_sb2 = _se_backward(g_y, c, gate, hid, pool, se_w1, se_w2, eps_res)
        if up:
            g_w, g_ln, g_b = conv_wgrad(au, g_c, w, ln, 3, 1, 1)
            g_au = conv_dgrad(g_c, wd, au.shape, 3, 1, 1)
            t = upsample_bwd(g_au, h, wd_, dsilu_src=x)  # gradient through SiLU -> up-sampling
            g_s = upsample_bwd(g_y, h, wd_)
            g_kw, g_kln, g_kb = conv_wgrad(x, g_s, skip_w, skip_ln, 1, 1, 0)
            g_x = conv_dgrad(g_s, wds, x.shape, 1, 1, 0, add_post=t)
        else:
            g_w, g_ln, g_b = conv_wgrad(x_act, g_c, w, ln, 3, 1, 1)
            g_x = conv_dgrad(g_c, wd, x.shape, 3, 1, 1, dsilu_src=x, add_post=g_y)
            g_kw = g_kln = g_kb = None
        return g_x, None, g_w, g_ln, g_b, g_sw1, g_sb1, g_sw2, g_sb2, g_kw, g_kln, g_kb, None, None


class _UpsampleFn(torch.autograd.Function):
    """nn.Upsample(size=...) between decoder cells (main/vae.py:779-780): resamples the map and its SiLU hint."""

    @staticmethod
    def forward(ctx, x, x_act, oh, ow):
        x = _check_input(x)
        ctx.in_hw = (x.shape[1], x.shape[2])
        out = upsample(x, oh, ow)
        out_act = upsample(x_act, oh, ow) if x_act is not None else silu(out)
        ctx.mark_non_differentiable(out_act)
        return out, out_act

    @staticmethod
    def backward(ctx, g, _g_act):
        return upsample_bwd(g.contiguous(), *ctx.in_hw), None, None, None


class _DecInputFn(torch.autograd.Function):
    """fc_dec's output (N, C * P) + bias, viewed as (N, C, 2, 2) NCHW by the reference (main/vae.py:56-57), into the NHWC
    map the decoder cells work on (+ its SiLU hint)."""

    @staticmethod
    def forward(ctx, h, bias, ch, p):
        h = _check_input(h)
        n = h.shape[0]
        side = int(round(math.sqrt(p)))
        y = torch.empty(n, side, side, ch, device=h.device, dtype=torch.float32)
        y_act = torch.empty_like(y)
        with torch.cuda.device(h.device):
            _lib.check(_lib.load().pvae_permute_cp(_p(h), _p(bias), _p(y), _p(y_act), n, ch, p, 1, _s()))
        ctx.dims = (n, ch, p, bias is not None)
        ctx.mark_non_differentiable(y_act)
        return y, y_act

    @staticmethod
    def backward(ctx, g_y, _g_act):
        n, ch, p, has_bias = ctx.dims
        g_y = g_y.contiguous()
        g_h = torch.empty(n, ch * p, device=g_y.device, dtype=torch.float32)
        with torch.cuda.device(g_y.device):
            _lib.check(_lib.load().pvae_permute_cp(_p(g_y), None, _p(g_h), None, n, ch, p, 0, _s()))
        return g_h, (colsum(g_h) if has_bias else None), None, None


class _HeadFn(torch.autograd.Function):
    """The decoder's last layer: weight-normalised 1x1 convolution to one channel (main/vae.py:782-789) -> (N, 1, H, W)."""

    @staticmethod
    def forward(ctx, x, weight, lognorm, bias):
        x = _check_input(x)
        n, h, w, cm = x.shape
        y = torch.empty(n, 1, h, w, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pvae_head_fwd(_p(x), _p(weight), _p(lognorm), _p(bias), _p(y), n * h * w, cm, weight.shape[1], _s()))
        ctx.save_for_backward(x, weight, lognorm)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, weight, lognorm = ctx.saved_tensors
        n, h, w, cm = x.shape
        cr = weight.shape[1]
        lib = _lib.load()
        g_y = g_y.contiguous()
        g_x = torch.empty_like(x)
        g_wb = torch.empty(cm + 1, device=x.device, dtype=torch.float32)
        ws = _workspace(x.device, lib.pvae_reduce_ws_floats(cm + 1))
        with torch.cuda.device(x.device):
            _lib.check(lib.pvae_head_bwd(_p(x), _p(g_y), _p(weight), _p(lognorm), _p(g_x), _p(g_wb), _p(ws), n * h * w, cm, cr, _s()))
        g_w, g_ln = weight_grad(g_wb, weight, lognorm, 1)
        return g_x, g_w, g_ln, g_wb[cm:cm + 1]


class _ResConvPoolFn(torch.autograd.Function):
    """ResConvPool main/vae.py:682-706: mean over the (k x k) map + eps * 'valid' k x k convolution of silu(x) -> (N, C)."""

    @staticmethod
    def forward(ctx, x, x_act, weight, lognorm, bias, eps_res):
        x = _check_input(x)
        n, h, w, ch = x.shape
        co = weight.shape[0]
        if x_act is None:
            x_act = silu(x)
        assert weight.shape[2] == h and weight.shape[3] == w, "ResConvPool expects the 'valid' convolution to cover the whole map"
        assert eps_res == 1.0
        wf, _, _ = weight_prep(weight, lognorm, None, h * w, need_dgrad=False, with_lo=False)
        pool = torch.empty(n, ch, device=x.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.pvae_avgpool_fwd(_p(x), _p(pool), n, h * w, ch, _s()))
        y = gemm_nt(x_act.view(n, h * w * ch), wf)
        with torch.cuda.device(x.device):
            _lib.check(lib.pvae_bias_act(_p(y), _p(bias), _p(pool), n, co, 0, 0.0, 0, 0, _s()))
        ctx.save_for_backward(x, x_act, weight, lognorm, wf)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, x_act, weight, lognorm, wf = ctx.saved_tensors
        n, h, w, ch = x.shape
        co = weight.shape[0]
        g_y = g_y.contiguous()
        g_wp = gemm_tn(x_act.view(n, h * w * ch), g_y)  # (taps * Ci, Co)
        g_w, g_ln = weight_grad(g_wp, weight, lognorm, h * w)
        g_b = colsum(g_y)
        g_a = gemm_nn(g_y, wf)  # (N, taps * Ci)
        g_x = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pvae_pool_silu_bwd(_p(g_a), _p(x), _p(g_y), _p(g_x), n, h * w, ch, _s()))
        return g_x, None, g_w, g_ln, g_b, None


class _ResDenseFn(torch.autograd.Function):
    """ResDenseLayer main/vae.py:657-679: LayerNorm(x + fc2(dropout(relu(fc1(x)))))."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, drop_p, seed, offset):
        x = _check_input(x)
        n, ch = x.shape
        lib = _lib.load()
        f1 = gemm_nt(x, w1.contiguous())
        f2 = torch.empty(n, ch, device=x.device, dtype=torch.float32)
        out, xhat = torch.empty_like(x), torch.empty_like(x)
        rstd = torch.empty(n, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(lib.pvae_bias_act(_p(f1), _p(b1), None, n, f1.shape[1], 1, drop_p, seed, offset, _s()))
            gemm_nt(f1, w2.contiguous(), out=f2)
            _lib.check(lib.pvae_bias_act(_p(f2), _p(b2), None, n, ch, 0, 0.0, 0, 0, _s()))
            _lib.check(lib.pvae_layernorm_fwd(_p(f2), _p(x), _p(gamma), _p(beta), _p(out), _p(xhat), _p(rstd), n, ch, 1e-5, _s()))
        ctx.save_for_backward(x, f1, xhat, rstd, w1, w2, gamma)
        ctx.drop_p = drop_p
        return out

    @staticmethod
    def backward(ctx, g):
        x, f1, xhat, rstd, w1, w2, gamma = ctx.saved_tensors
        n, ch = x.shape
        lib = _lib.load()
        g = g.contiguous()
        g_s = torch.empty_like(x)
        g_gb = torch.empty(2 * ch, device=x.device, dtype=torch.float32)
        ws = _workspace(x.device, lib.pvae_reduce_ws_floats(2 * ch))
        with torch.cuda.device(x.device):
            _lib.check(lib.pvae_layernorm_bwd(_p(g), _p(xhat), _p(rstd), _p(gamma), _p(g_s), _p(g_gb), _p(ws), n, ch, _s()))
        g_w2 = gemm_tn(g_s, f1)
        g_b2 = colsum(g_s)
        g_f1 = gemm_nn(g_s, w2.contiguous())
        with torch.cuda.device(x.device):
            _lib.check(lib.pvae_relu_bwd(_p(g_f1), _p(f1), g_f1.numel(), ctx.drop_p, _s()))
        g_w1 = gemm_tn(g_f1, x)
        g_b1 = colsum(g_f1)
        g_x = gemm_nn(g_f1, w1.contiguous())
        with torch.cuda.device(x.device):
            _lib.check(lib.pvae_add(_p(g_x), _p(g_s), _p(g_x), g_x.numel(), _s()))
        return g_x, g_w1, g_b1, g_w2, g_b2, g_gb[:ch], g_gb[ch:], None, None, None


# ------------------------------------------------------------------------------------------------------
# modules: parameter holders with the reference's names, construction order and initialisation
# ------------------------------------------------------------------------------------------------------
class Conv2D(nn.Conv2d):
    """base/common.py:295-340: nn.Conv2d parameters + `lognorm`; the filter the kernels use is
    exp(lognorm) * weight / (||weight|| + eps) per output channel."""

    def __init__(self, in_channels, out_channels, kernel_size, reg_lognorm: bool = True, init_scale: float = 1.0, **kwargs):
        kwargs = {k: v for k, v in kwargs.items() if k in ('stride', 'padding', 'dilation', 'groups', 'bias', 'padding_mode')}
        super().__init__(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, **kwargs)
        assert init_scale > 0
        self.lognorm = nn.Parameter(torch.log(torch.ones(out_channels).mul(init_scale)), requires_grad=reg_lognorm)

    def forward(self, x):
        raise RuntimeError("Conv2D is a parameter holder here; the enclosing block runs the fused kernels")


class ConvLayer(nn.Module):
    """base/common.py:195-242 (use_bn off): activation -> [upsample] -> Conv2D."""

    def __init__(self, ci, co, stride, act_fn='swish', init_scale=1.0, reg_lognorm=True):
        super().__init__()
        assert act_fn in ('swish', 'silu')
        self.stride = stride
        self.conv = Conv2D(ci, co, 3, reg_lognorm=reg_lognorm, init_scale=init_scale, stride=abs(stride), padding=1, bias=True)


class SELayer(nn.Module):
    """base/common.py:129-142."""

    def __init__(self, ci, reduc=16):
        super().__init__()
        self.hdim = max(ci // reduc, 4)
        self.fc = nn.Sequential(nn.Linear(ci, self.hdim), nn.ReLU(inplace=True), nn.Linear(self.hdim, ci), nn.Sigmoid())


class FactorizedReduce(nn.Module):
    """base/common.py:98-126: four 1x1 stride-2 convolutions on the four pixel phases, concatenated."""

    def __init__(self, ci, co, reg_lognorm=True):
        super().__init__()
        assert co % 2 == 0 and co > 4
        co_each = co // 4
        self.ops = nn.ModuleList()
        for _ in range(3):
            self.ops.append(Conv2D(ci, co_each, 1, reg_lognorm=reg_lognorm, stride=co // ci, padding=0, bias=True))
        self.ops.append(Conv2D(ci, co - 3 * co_each, 1, reg_lognorm=reg_lognorm, stride=co // ci, padding=0, bias=True))

    def packed(self):
        return (torch.cat([op.weight.flatten(1) for op in self.ops]), torch.cat([op.lognorm for op in self.ops]),
                torch.cat([op.bias for op in self.ops]))


class StrideReduce(nn.Module):
    """base/common.py:75-95: SiLU -> 1x1 stride-2 convolution."""

    def __init__(self, ci, co, reg_lognorm=True):
        super().__init__()
        self.conv = Conv2D(ci, co, 1, reg_lognorm=reg_lognorm, stride=co // ci, padding=0, bias=True)


class Cell(nn.Module):
    """base/common.py:145-192: 'normal_pre' / 'down_enc' (encoder, n_nodes = 2) and 'normal_dec' / 'up_dec' (decoder, n_nodes = 1)."""

    def __init__(self, ci, co, cell_type, n_nodes=2, act_fn='swish', use_bn=False, use_se=True, scale=1.0, eps=1.0,
                 reg_lognorm=True, factorized=True):
        super().__init__()
        if cell_type not in ('normal_pre', 'down_enc', 'normal_dec', 'up_dec'):
            raise NotImplementedError(f"cell type {cell_type}")
        self.decoder = cell_type.endswith('_dec')
        if use_bn or not use_se or n_nodes != (1 if self.decoder else 2):
            raise NotImplementedError("Cell: use_se=True, use_bn=False, n_nodes 2 (encoder) / 1 (decoder): the reference defaults")
        self.stride = {'normal': 1, 'down': MULT, 'up': -1}[cell_type.split('_')[0]]
        if self.stride == 1:
            assert ci == co
            self.skip = nn.Identity()
        elif self.stride == -1:  # get_skip_connection base/common.py:35-47
            self.skip = nn.Sequential(nn.Upsample(scale_factor=MULT, mode='nearest'),
                                      Conv2D(ci, int(ci / MULT), 1, reg_lognorm=reg_lognorm))
        else:
            self.skip = FactorizedReduce(ci, co, reg_lognorm) if factorized else StrideReduce(ci, co, reg_lognorm)
        self.ops = nn.ModuleList()
        for i in range(n_nodes):
            self.ops.append(ConvLayer(ci if i == 0 else co, co, self.stride if i == 0 else 1, act_fn,
                                      init_scale=scale if i + 1 == n_nodes else 1.0, reg_lognorm=reg_lognorm))
        self.se = SELayer(co)
        self.eps = eps

    def forward(self, xs):
        x, x_act = xs if isinstance(xs, tuple) else (xs, None)
        fc = self.se.fc
        c1 = self.ops[0].conv
        if self.decoder:
            up = self.stride == -1
            sk = self.skip[1] if up else None
            return _DecCellFn.apply(x, x_act, c1.weight, c1.lognorm, c1.bias, fc[0].weight, fc[0].bias, fc[2].weight, fc[2].bias,
                                    sk.weight if up else None, sk.lognorm if up else None, sk.bias if up else None, up,
                                    float(self.eps))
        c2 = self.ops[1].conv
        if self.stride == 1:
            sw = sln = sb = None
            groups = 1
        elif isinstance(self.skip, FactorizedReduce):
            sw, sln, sb = self.skip.packed()
            groups = 4
        else:
            sw, sln, sb = self.skip.conv.weight, self.skip.conv.lognorm, self.skip.conv.bias
            groups = 1
        return _CellFn.apply(x, x_act, c1.weight, c1.lognorm, c1.bias, c2.weight, c2.lognorm, c2.bias, fc[0].weight, fc[0].bias,
                             fc[2].weight, fc[2].bias, sw, sln, sb, self.stride, groups, float(self.eps))


class ResConvPool(nn.Module):
    """main/vae.py:682-706."""

    def __init__(self, dim, act_fn='swish', eps=1.0, kernel_size=4, reg_lognorm=True):
        super().__init__()
        self.conv = Conv2D(dim, dim, kernel_size, reg_lognorm=reg_lognorm, padding='valid')
        self.eps = eps

    def forward(self, xs):
        x, x_act = xs if isinstance(xs, tuple) else (xs, None)
        return _ResConvPoolFn.apply(x, x_act, self.conv.weight, self.conv.lognorm, self.conv.bias, float(self.eps))


class ResDenseLayer(nn.Module):
    """main/vae.py:657-679.  In training mode the dropout mask comes from this library's Philox stream (seeded by torch's
    CUDA generator), not from torch's dropout kernel: same law, different draws."""

    def __init__(self, dim, expand=8, drop=0.1):
        super().__init__()
        self.fc1 = nn.Linear(dim, dim * expand)
        self.fc2 = nn.Linear(dim * expand, dim)
        self.layer_norm = nn.LayerNorm(dim)
        self.drop = nn.Dropout(drop)
        self.relu = nn.ReLU()
        self.dim = dim

    def forward(self, x):
        p = float(self.drop.p) if self.training else 0.0
        seed, offset = next_philox(x.device) if p > 0.0 else (0, 0)
        return _ResDenseFn.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.layer_norm.weight,
                                 self.layer_norm.bias, p, seed, offset)


class Stem(Conv2D):
    """main/vae.py:211-218: 3x3 convolution of the single-channel image (padding 1, or 'valid' for MNIST)."""

    def forward(self, x):
        pad = 0 if self.padding in ('valid', (0, 0)) else 1
        return _StemFn.apply(x, self.weight, self.lognorm, self.bias, pad)


class _Flatten(nn.Module):
    """nn.Flatten(start_dim=1) placeholder (keeps the reference's nn.Sequential indices); ResConvPool already returns (N, C)."""

    def forward(self, x):
        return x.flatten(1)


def build_conv_enc(nch: int, dataset: str, act_fn='swish', eps=1.0, reg_lognorm=True) -> nn.Sequential:
    """main/vae.py:710-756 (`_build_conv_enc`, add_fc off)."""
    factorized = True
    if dataset.endswith('MNIST'):
        factorized, n_layers = False, 3
    elif dataset in ('vH16', 'CIFAR16'):
        n_layers = 2
    elif dataset.startswith('BALLS'):
        n_layers = 4
    else:
        raise ValueError(dataset)
    kws = dict(n_nodes=2, act_fn=act_fn, use_bn=False, use_se=True, scale=1.0, eps=eps, reg_lognorm=reg_lognorm,
               factorized=factorized)
    layers = [Cell(nch, nch, 'normal_pre', **kws)]
    for _ in range(n_layers):
        layers.extend([Cell(nch, nch * MULT, 'down_enc', **kws), Cell(nch * MULT, nch * MULT, 'normal_pre', **kws)])
        nch *= MULT
    layers.extend([ResConvPool(nch, act_fn=act_fn, eps=eps, kernel_size=4, reg_lognorm=reg_lognorm), _Flatten()])
    return nn.Sequential(*layers, ResDenseLayer(nch))


class Resample(nn.Upsample):
    """nn.Upsample(size=...) inside the decoder's nn.Sequential (main/vae.py:779-780), on (map, SiLU hint) pairs."""

    def forward(self, xs):
        x, x_act = xs if isinstance(xs, tuple) else (xs, None)
        size = (self.size, self.size) if isinstance(self.size, int) else tuple(self.size)
        return _UpsampleFn.apply(x, x_act, size[0], size[1])


class Head(Conv2D):
    """The final Conv2D(nch, 1, kernel_size=1) of the decoder (main/vae.py:782-789)."""

    def forward(self, xs):
        x = xs[0] if isinstance(xs, tuple) else xs
        return _HeadFn.apply(x, self.weight, self.lognorm, self.bias)


def build_conv_dec(nch: int, dataset: str, act_fn='swish', eps=1.0, reg_lognorm=True) -> nn.Sequential:
    """main/vae.py:760-789 (`_build_conv_dec`)."""
    if dataset.endswith('MNIST'):
        n_layers = 4
    elif dataset in ('vH16', 'CIFAR16', 'BALLS'):
        n_layers = 3
    else:
        raise ValueError(dataset)
    kws = dict(n_nodes=1, act_fn=act_fn, use_bn=False, use_se=True, scale=1.0, eps=eps, reg_lognorm=reg_lognorm)
    layers = [Cell(nch, nch, 'normal_dec', **kws)]
    for i in range(n_layers):
        layers.extend([Cell(nch, nch // MULT, 'up_dec', **kws), Cell(nch // MULT, nch // MULT, 'normal_dec', **kws)])
        nch //= MULT
        if dataset.endswith('MNIST') and i == 2:
            layers.append(Resample(size=14))
    return nn.Sequential(*layers, Head(nch, 1, 1, reg_lognorm=True, padding='valid'))


def decoder_input(h, bias, nch: int, spat: int = 2):
    """fc_dec output (N, nch * spat^2) [+ bias] -> ((N, spat, spat, nch) NHWC map, its SiLU hint)"""
    return _DecInputFn.apply(h, bias, nch, spat * spat)
