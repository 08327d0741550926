"""Host-side mirror of the reference `PoissonVAE` (main/vae.py:374-492, BaseVAE main/vae.py:9-72) for the
lin|lin architecture and the conv+b|lin one (conv encoder of main/vae.py:211-259, 710-756 - see conv.py), and of
the slice of its config the hot path reads (main/config_vae.py:14-110, 172-195, 373-465).

Same method names, argument meaning, return tuples, parameter/buffer names and state_dict keys
(`log_rate, temp, n_exp, fc_enc.weight, fc_dec.weight`), so a reference trainer can drive it.
The sampler, KL and their backward run as sm_100a kernels through include/pvae_b200.h.
"""
from __future__ import annotations

import os
import random
from typing import Sequence

import numpy as np
import torch
from torch import nn

from . import _lib
from .distributions import (EXIT_LOSSLESS, INDICATOR_CODES, Poisson, _ptr, _require_cuda_f32, _stream,
                            softclamp, softclamp_upper)
from .linear import linear


class ConfigPoisVAE:
    """The fields of the reference ConfigPoisVAE/ConfigVAE that the lin|lin hot path reads.
    Like the reference (base/config_base.py:51-63) constructing it seeds torch, numpy and random."""

    def __init__(self, n_latents: int = 512, input_sz: int = None, dataset: str = 'vH16',
                 enc_type: str = 'lin', dec_type: str = 'lin', enc_bias: bool = False, dec_bias: bool = False,
                 prior_clamp: float = -2.0, prior_log_dist: str = 'uniform', indicator_approx: str = 'sigmoid',
                 exc_only: bool = False, fit_prior: bool = True, init_dist: str = 'Normal',
                 init_scale: float | None = 0.05, seed: int = 0, n_ch: int = 32, activation_fn: str = 'swish',
                 res_eps: float = 1.0, use_bn: bool = False, use_se: bool = True, enc_norm: bool = False,
                 dec_norm: bool = False, **unused):
        assert prior_log_dist in ('cte', 'uniform', 'normal')
        assert indicator_approx in INDICATOR_CODES
        if enc_type not in ('lin', 'conv') or dec_type not in ('lin', 'conv'):
            raise NotImplementedError("poissonvae_b200 covers the lin and conv encoders / decoders (SURVEY.md section 8); "
                                      f"got {enc_type}|{dec_type}")
        if init_dist != 'Normal':
            raise NotImplementedError(init_dist)
        if enc_norm or dec_norm:
            raise NotImplementedError("weight-normalised fc_enc / fc_dec (off by default in the reference)")
        if 'conv' in (enc_type, dec_type) and (use_bn or not use_se or activation_fn not in ('swish', 'silu') or res_eps != 1.0):
            raise NotImplementedError("conv encoder: the reference defaults only (swish, SE on, BatchNorm off, res_eps 1)")
        if input_sz is None:  # ConfigVAE._init_input_sz main/config_vae.py:97-110
            input_sz = 28 if dataset.endswith('MNIST') else 16
        self.n_latents, self.input_sz, self.dataset = n_latents, input_sz, dataset
        self.enc_type, self.dec_type, self.enc_bias, self.dec_bias = enc_type, dec_type, enc_bias, dec_bias
        self.n_ch, self.activation_fn, self.res_eps, self.use_bn, self.use_se = n_ch, activation_fn, res_eps, use_bn, use_se
        self.prior_clamp, self.prior_log_dist = prior_clamp, prior_log_dist
        self.indicator_approx, self.exc_only, self.fit_prior = indicator_approx, exc_only, fit_prior
        self.init_dist, self.init_scale = init_dist, init_scale
        self._set_seeds(seed)

    @property
    def type(self):
        return 'poisson'

    def _set_seeds(self, seed):
        assert 0 <= seed <= np.iinfo(np.uint32).max
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
        os.environ["SEED"] = str(seed)
        np.random.seed(seed)
        random.seed(seed)
        self.seed = seed


def default_configs(dataset: str = 'vH16', model_type: str = 'poisson', archi_type: str = 'lin|lin'):
    """main/config_vae.py:373-465 restricted to poisson with lin / conv encoders and decoders (with or without +b)."""
    archi = archi_type.strip('<>')
    enc, _, dec = archi.partition('|')
    if model_type != 'poisson' or enc.split('+')[0] not in ('lin', 'conv') or dec.split('+')[0] not in ('lin', 'conv'):
        raise NotImplementedError((model_type, archi_type))
    lin_dec = dec.split('+')[0] == 'lin'
    cfg_vae = dict(dataset=dataset, n_ch=32, n_latents=512 if lin_dec else 10, prior_clamp=-4 if lin_dec else -2,
                   fit_prior=True, enc_type=enc.split('+')[0], dec_type=dec.split('+')[0], enc_bias=enc.endswith('+b'),
                   dec_bias=dec.endswith('+b'), init_dist='Normal', init_scale=0.05 if lin_dec else 0.1)
    if dataset in ('vH16', 'CIFAR16'):
        cfg_tr = dict(lr=0.005, batch_size=1000, epochs=3000 if dataset == 'vH16' else 1500, grad_clip=None)
    elif dataset.endswith('MNIST'):
        cfg_tr = dict(lr=0.002, epochs=200 if dataset == 'EMNIST' else 400, batch_size=100, warm_restart=0, grad_clip=1000)
    else:
        raise NotImplementedError(dataset)
    cfg_tr['kl_const_portion'] = 0.0 if lin_dec else 0.01
    return cfg_vae, cfg_tr


class _Linear(nn.Linear):
    """base/common.py:391-423 without the (default-off) weight normalisation."""

    def forward(self, x):
        return linear(x, self.weight, self.bias)

    def get_weight(self):
        return self.weight


class _InferRateFn(torch.autograd.Function):
    """(h, log_r[, b_enc]) -> (rate, log_dr) without sampling (method='exact', main/vae.py:74-93 needs only the
    posterior's mean = variance = rate).  Forward: the fused kernel with zero events.  Backward: the elementwise
    backward kernel fed dz_acc = rate at temp = 1, for which its chain factor dz_acc / temp * exp(lambda) / rate
    is d rate / d lambda = exp(lambda), so `g_z` plays the role of d loss / d rate."""

    @staticmethod
    def forward(ctx, h, log_r, b_enc, exc_only):
        h = h.contiguous()
        B, K = h.shape
        lr = log_r.reshape(-1).contiguous()
        z = torch.empty_like(h)
        log_dr = torch.empty_like(h)
        rate = torch.empty_like(h)
        with torch.cuda.device(h.device):
            _lib.check(_lib.load().pvae_sampler_fwd(
                _ptr(h), _ptr(lr), _ptr(b_enc), _ptr(z), _ptr(log_dr), _ptr(rate), None, None, None, None, B, K, 0, 0,
                1.0, 0, int(exc_only), -1.0, 0, 0, None, _stream()))
        ctx.save_for_backward(h, lr, b_enc, rate)
        ctx.args = (B, K, int(exc_only), log_r.shape)
        return rate, log_dr

    @staticmethod
    def backward(ctx, g_rate, g_log_dr):
        h, lr, b_enc, rate = ctx.saved_tensors
        B, K, exc_only, lr_shape = ctx.args
        lib = _lib.load()
        g_rate = g_rate.contiguous()
        g_log_dr = None if g_log_dr is None else g_log_dr.contiguous()
        g_h, g_lr = torch.empty_like(h), torch.empty_like(lr)
        g_b = None if b_enc is None else torch.empty_like(b_enc)
        ws = torch.empty(2 * lib.pvae_num_partials(B) * K, device=h.device, dtype=torch.float32)
        with torch.cuda.device(h.device):
            _lib.check(lib.pvae_sampler_bwd(
                _ptr(h), _ptr(lr), _ptr(b_enc), _ptr(g_rate), _ptr(g_log_dr), _ptr(rate), 0.0, _ptr(g_h), _ptr(g_lr),
                _ptr(g_b), _ptr(ws), 0, B, K, 0, 0, 1.0, 0, exc_only, -1.0, 0, 0, None, _stream()))
        return g_h, g_lr.reshape(lr_shape), g_b, None


class _KLFn(torch.autograd.Function):
    """PoissonVAE.loss_kl main/vae.py:446-450 as one elementwise kernel (+ backward)."""

    @staticmethod
    def forward(ctx, log_dr, log_r):
        log_dr = log_dr.contiguous()
        B, K = log_dr.shape
        lr = log_r.reshape(-1).contiguous()
        kl = torch.empty_like(log_dr)
        with torch.cuda.device(log_dr.device):
            _lib.check(_lib.load().pvae_kl_fwd(_ptr(log_dr), _ptr(lr), _ptr(kl), B, K, _stream()))
        ctx.save_for_backward(log_dr, lr)
        ctx.lr_shape = log_r.shape
        return kl

    @staticmethod
    def backward(ctx, g_kl):
        log_dr, lr = ctx.saved_tensors
        B, K = log_dr.shape
        lib = _lib.load()
        g_kl = g_kl.contiguous()
        g_ld = torch.empty_like(log_dr)
        g_lr = torch.empty_like(lr)
        ws = torch.empty(lib.pvae_num_partials(B) * K, device=lr.device, dtype=torch.float32)
        with torch.cuda.device(lr.device):
            _lib.check(lib.pvae_kl_bwd(_ptr(log_dr), _ptr(lr), _ptr(g_kl), _ptr(g_ld), _ptr(g_lr), _ptr(ws), 0, B, K,
                                       _stream()))
        return g_ld, g_lr.reshape(ctx.lr_shape)


class _PosteriorPoisson(Poisson):
    """The `dist` object PoissonVAE.infer returns: a Poisson whose log-rate
    softclamp_upper(log_r + log_dr, 5) (main/vae.py:408-414) is only materialised if a caller
    reads `.rate` / `.mean` / `.variance` or samples from it again."""

    def __init__(self, vae, h, log_dr, temp, temp_host, n_exp, sampled):
        self.indicator_approx = vae.cfg.indicator_approx
        self.temp = temp  # the model's `temp` buffer when t is None, like the reference
        self._temp_host = temp_host
        self.clamp = None
        self.exit_logit = vae.exit_logit
        self.n_exp = int(n_exp)
        self._vae, self._h, self._log_dr = vae, h, log_dr
        self._rate = None
        self._lr_cache = None
        self._sampled = sampled  # (z, log_dr) of the fused launch, handed out by the first rsample()

    @property
    def _log_rate_in(self):
        if self._lr_cache is None:
            log_r = self._vae.log_rate.expand(len(self._log_dr), -1)
            self._lr_cache = softclamp_upper(log_r + self._log_dr, 5.0)
        return self._lr_cache

    def rsample(self):
        if self._sampled is not None:
            z, self._sampled = self._sampled, None
            return z
        saved, self.temp = self.temp, self._temp_host  # no device sync for the buffer-valued temp
        try:
            return super().rsample()
        finally:
            self.temp = saved


class PoissonVAE(nn.Module):
    def __init__(self, cfg: ConfigPoisVAE, exit_logit: float = EXIT_LOSSLESS, **kwargs):
        super().__init__()
        self.cfg = cfg
        self.Dist = Poisson
        self.exit_logit = float(exit_logit)
        self.register_buffer('temp', torch.tensor(1.0))  # main/vae.py:13-16
        D, K = cfg.input_sz ** 2, cfg.n_latents
        # same construction order as BaseVAE._init (main/vae.py:176-184) so that a given cfg.seed
        # yields the reference's initial parameters
        if cfg.enc_type == 'conv':  # main/vae.py:209-236
            from . import conv
            self.stem = conv.Stem(1, cfg.n_ch, 3, padding='valid' if cfg.dataset.endswith('MNIST') else 1, reg_lognorm=True)
            self.enc = conv.build_conv_enc(cfg.n_ch, cfg.dataset, cfg.activation_fn, cfg.res_eps)
            self.fc_enc = _Linear(self.enc[-1].dim, K, bias=cfg.enc_bias)
        else:
            self.fc_enc = _Linear(D, K, bias=cfg.enc_bias)  # main/vae.py:253-259
        if cfg.dec_type == 'conv':  # main/vae.py:264-279
            from . import conv
            nch = cfg.n_ch * 8
            self.shape = (-1, nch, 2, 2)
            self.fc_dec = _Linear(K, nch * 4, bias=cfg.dec_bias)
            self.dec = conv.build_conv_dec(nch, cfg.dataset, cfg.activation_fn, cfg.res_eps)
        else:
            self.shape = (-1, 1, cfg.input_sz, cfg.input_sz)
            self.fc_dec = _Linear(K, D, bias=cfg.dec_bias)  # main/vae.py:306-312
        self._init_prior()
        self._init_weight()
        self.register_buffer('n_exp', torch.tensor(0))  # main/vae.py:378-381
        # host mirrors of the two scalar buffers (plain attributes: no device sync to read them)
        object.__setattr__(self, '_temp_host', 1.0)
        object.__setattr__(self, '_n_exp_host', 0)
        # set by a data-parallel trainer around a step: global index of the first local sample (keys the Philox stream),
        # a 1-float device tensor that receives max(rate), and the (seed, offset) to use instead of the generator's next
        object.__setattr__(self, 'row_offset', 0)
        object.__setattr__(self, 'rate_max_out', None)
        object.__setattr__(self, 'philox_override', None)
        self.update_n(200.0)  # main/vae.py:382

    # ---- reference API -------------------------------------------------------------------
    def forward(self, x):
        dist, log_dr = self.infer(x)
        spks = dist.rsample()
        y = self.decode(spks)
        return dist, log_dr, spks, y

    def infer(self, x: torch.Tensor, t: float = None, ablate: Sequence[int] = None):
        """main/vae.py:393-415.  `ablate` forces the unfused path (log_dr is edited in between)."""
        temp = self._temp_host if t is None else float(t)
        assert temp >= 0.0, f"must be non-neg: {temp}"
        _require_cuda_f32(x, "x")
        h = self.encode_linear(x)
        if ablate is not None:
            if self.cfg.exc_only:
                log_dr = softclamp(h + (0 if self.fc_enc.bias is None else self.fc_enc.bias), 10.0, 0)
            else:
                log_dr = softclamp_upper(h + (0 if self.fc_enc.bias is None else self.fc_enc.bias), 10.0)
            log_dr[:, ablate] = 0.0
            log_r = self.log_rate.expand(len(x), -1)
            dist = Poisson(log_rate=softclamp_upper(log_r + log_dr, 5.0), temp=temp,
                           indicator_approx=self.cfg.indicator_approx, n_exp=self._n_exp_host,
                           exit_logit=self.exit_logit)
            return dist, log_dr
        seed, offset = (-1, -1) if self.philox_override is None else self.philox_override  # -1: the op asks torch's generator
        from . import ops  # noqa: F401  (registers torch.ops.pvae.*)
        z, log_dr, _, _, _, r_max = torch.ops.pvae.encode_sample_kl(h, self.log_rate, self.fc_enc.bias, temp, self._n_exp_host,
                                                                    self.cfg.indicator_approx, bool(self.cfg.exc_only),
                                                                    self.exit_logit, seed, offset, self.row_offset)
        if self.rate_max_out is not None:  # running max of the rates for update_n, kept on the device
            with torch.no_grad():
                torch.maximum(self.rate_max_out, r_max, out=self.rate_max_out)
        dist = _PosteriorPoisson(self, h, log_dr, self.temp if t is None else t, temp, self._n_exp_host, z)
        return dist, log_dr

    @torch.no_grad()
    def xtract_ftr(self, x, t: float = 0.0, ablate_enc: Sequence[int] = None, ablate_spks: Sequence[int] = None):
        dist, log_dr = self.infer(x, t, ablate_enc)
        spks = dist.rsample()
        if ablate_spks is not None:
            spks[:, ablate_spks] = 0.0
        y = self.decode(spks)
        return dist, log_dr, spks, y

    @torch.no_grad()
    def sample(self, n: int, t: float = 0.0):
        """main/vae.py:431-444: prior samples (no softclamp on the prior log-rate)."""
        temp = self._temp_host if t is None else float(t)
        log_r = self.log_rate.expand(n, -1).contiguous()
        dist = self.Dist(log_rate=log_r, n_exp=self._n_exp_host, temp=temp,
                         indicator_approx=self.cfg.indicator_approx, exit_logit=self.exit_logit)
        spks = dist.rsample()
        return self.decode(spks), spks

    def _features(self, x):
        """What fc_enc sees, main/vae.py:40-50: the flattened input, or the conv encoder's (N, C) output."""
        if self.cfg.enc_type == 'conv':
            return self.enc(self.stem(x.reshape(x.shape[0], 1, self.cfg.input_sz, self.cfg.input_sz)))
        return x.flatten(start_dim=1)

    def encode_linear(self, x):
        # bias (if any) is added inside the sampler kernel
        return linear(self._features(x), self.fc_enc.weight, None)

    def encode(self, x):  # main/vae.py:40-52
        return self.fc_enc(self._features(x))

    def decode(self, z):  # main/vae.py:54-66
        if self.cfg.dec_type == 'conv':
            from . import conv
            h = linear(z, self.fc_dec.weight, None)  # the bias joins in the NCHW -> NHWC kernel
            return self.dec(conv.decoder_input(h, self.fc_dec.bias, self.shape[1], self.shape[2]))
        return self.fc_dec(z).view(*self.shape)

    def loss_recon(self, y, x):  # main/vae.py:68-72
        return torch.sum((y - x) ** 2, dim=[1, 2, 3])

    def loss_recon_exact(self, x):
        """main/vae.py:74-93 (method='exact', linear decoder): E_q ||x - z Phi^T||^2 in closed form for a Poisson
        posterior, mean = variance = rate:  ||x - rate Phi^T||^2 + rate . diag(Phi^T Phi).  No sampling."""
        _require_cuda_f32(x, "x")
        h = self.encode_linear(x)
        rate, log_dr = _InferRateFn.apply(h, self.log_rate, self.fc_enc.bias, self.cfg.exc_only)
        dist = _PosteriorPoisson(self, h, log_dr, self.temp, self._temp_host, self._n_exp_host, None)
        dist._rate = rate
        phi = self.fc_dec.weight
        gram = phi.pow(2).sum(0)
        mse = (x.flatten(start_dim=1) - linear(rate, phi)).pow(2).sum(1)
        return mse + rate @ gram, dist, log_dr

    def loss_kl(self, log_dr):  # main/vae.py:446-450
        return _KLFn.apply(log_dr, self.log_rate)

    def loss_weight(self):
        return None

    def update_t(self, new_t: float):  # main/vae.py:33-38
        if new_t is None:
            return
        assert new_t >= 0.0, "must be non-neg"
        if float(new_t) != self._temp_host:  # the buffer write is a kernel launch: skip it when nothing changes
            self.temp.fill_(new_t)
        self._temp_host = float(new_t)

    def update_n(self, rate: float):  # main/vae.py:452-459
        if rate is None:
            return
        assert rate > 0.0, "must be positive"
        from scipy import stats as sp_stats
        n_exp = int(sp_stats.poisson(rate).ppf(1.0 - 1e-5))
        self._n_exp_host = n_exp
        self.n_exp.fill_(n_exp)

    def load_state_dict(self, state_dict, *args, **kwargs):
        out = super().load_state_dict(state_dict, *args, **kwargs)
        self._temp_host = float(self.temp.item())  # host mirrors of the two scalar buffers
        self._n_exp_host = int(self.n_exp.item())
        return out

    # ---- init, main/vae.py:334-348, 461-492 ------------------------------------------------
    def _init_prior(self):
        rng = np.random.default_rng(self.cfg.seed)
        size = (1, self.cfg.n_latents)
        if self.cfg.prior_log_dist == 'cte':
            log_rate = np.ones(size) * self.cfg.prior_clamp
        elif self.cfg.prior_log_dist == 'uniform':
            log_rate = rng.uniform(low=-6.0, high=self.cfg.prior_clamp, size=size)
        else:  # 'normal'
            s = np.abs(np.log(np.abs(self.cfg.prior_clamp)))
            log_rate = rng.normal(loc=0.0, scale=s, size=size)
        log_rate = torch.tensor(log_rate, dtype=torch.float)
        log_rate[log_rate > 6.0] = 0.0
        self.log_rate = nn.Parameter(log_rate, requires_grad=True)  # main/vae.py:488-491

    @torch.no_grad()
    def _init_weight(self):
        if self.cfg.init_scale is None:
            return
        dist = torch.distributions.Normal(loc=0, scale=self.cfg.init_scale)
        self.fc_enc.weight.copy_(dist.sample(self.fc_enc.weight.shape))
        if self.cfg.dec_type == 'lin':
            self.fc_dec.weight.copy_(dist.sample(self.fc_dec.weight.shape))
        if self.fc_enc.bias is not None:
            nn.init.zeros_(self.fc_enc.bias)
