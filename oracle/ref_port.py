"""torch-CPU port of the reference training step.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

A restatement of what hadivafaii/PoissonVAE executes for one poisson lin|lin step on CPU, with the
same sequence of torch ops (so it costs what the reference costs on the same cores): the
(n_exp, B, K) exponential draw, division, cumsum, soft threshold, sum (base/distributions.py:55-81),
the two linear layers, softclamps, KL and loss (main/vae.py:384-450, main/train_vae.py:346-362),
autograd backward and the "fast" Adamax (base/adamax.py:34-100).  It is pinned to the unmodified
reference by tests/test_ref_port.py (fixtures from oracle/make_golden.py, and - when
/root/reference is present - a bit-for-bit comparison under the same torch seed).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import it.
"""
from __future__ import annotations

import torch
from torch.nn import functional as F

EPS = torch.finfo(torch.float32).eps


def softclamp_upper(x, c):  # base/distributions.py:241-242
    return c - F.softplus(c - x)


def softclamp(x, upper, lower=0.0):  # base/distributions.py:233-234
    return lower + F.softplus(x - lower) - F.softplus(x - upper)


def rsample(rate, n_exp, temp):
    """base/distributions.py:55-81 with the sigmoid indicator."""
    x = rate.new(torch.Size((n_exp,)) + rate.shape).exponential_() / rate  # Exponential.rsample
    times = torch.cumsum(x, dim=0)
    logits = (1 - times) / temp
    return torch.sigmoid(logits).sum(0)


def forward_loss(params, x, n_exp, temp, beta, exc_only=False):
    """params: dict(log_rate (1,K), w_enc (K,D), w_dec (D,K)); x (B,1,H,W). Returns loss, aux."""
    log_r = params["log_rate"].expand(len(x), -1)
    h = F.linear(x.flatten(start_dim=1), params["w_enc"])
    log_dr = softclamp(h, 10.0, 0) if exc_only else softclamp_upper(h, 10.0)
    rate = torch.exp(softclamp_upper(log_r + log_dr, 5.0)) + EPS
    z = rsample(rate, n_exp, temp)
    y = F.linear(z, params["w_dec"]).view(x.shape)
    recon = torch.sum((y - x) ** 2, dim=[1, 2, 3])
    kl = torch.exp(log_r) * (1 + torch.exp(log_dr) * (log_dr - 1))
    loss = torch.mean(recon + beta * torch.sum(kl, dim=1))
    return loss, dict(z=z, y=y, recon=recon, kl=kl, log_dr=log_dr, rate=rate)


class Adamax:
    """base/adamax.py:34-100 for a dict of dense parameters (no weight decay)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params, self.lr, self.betas, self.eps, self.t = params, lr, betas, eps, 0
        self.exp_avg = {k: torch.zeros_like(v) for k, v in params.items()}
        self.exp_inf = {k: torch.zeros_like(v) for k, v in params.items()}

    @torch.no_grad()
    def step(self):
        self.t += 1
        b1, b2 = self.betas
        clr = self.lr / (1 - b1 ** self.t)
        for k, p in self.params.items():
            g = p.grad
            self.exp_avg[k].mul_(b1).add_(g, alpha=1 - b1)
            torch.max(self.exp_inf[k].mul_(b2), g.abs().add_(self.eps), out=self.exp_inf[k])
            p.addcdiv_(self.exp_avg[k], self.exp_inf[k], value=-clr)


def make_params(K=512, D=256, seed=0, prior_clamp=-4.0, init_scale=0.05):
    """Random-init weights with the reference's distributions (main/vae.py:334-348, 461-492)."""
    g = torch.Generator().manual_seed(seed)
    return dict(
        log_rate=(torch.rand(1, K, generator=g) * (prior_clamp + 6.0) - 6.0).requires_grad_(),
        w_enc=(torch.randn(K, D, generator=g) * init_scale).requires_grad_(),
        w_dec=(torch.randn(D, K, generator=g) * init_scale).requires_grad_(),
    )


def train_step(params, optim, x, n_exp, temp=1.0, beta=1.0):
    for p in params.values():
        p.grad = None
    loss, _ = forward_loss(params, x, n_exp, temp, beta)
    loss.backward()
    optim.step()
    return loss.detach()
