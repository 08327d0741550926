"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference is imported as-is (oracle/ref_import.py); Exp(1) draws are injected through a
patched Tensor.exponential_, so the recorded outputs are what the reference's own code
computes on inputs we can replay through the CUDA path and the numpy oracle.

Fixtures (all fp32 CPU; sizes kept small so the files stay a few hundred KB):
  sampler_<indicator>.npz   Poisson(log_rate, temp, indicator, n_exp).rsample() and its autograd
                            gradient for three temperatures            (base/distributions.py:55-81)
  hard_counts.npz           (cumsum(Exponential(rate).rsample((n,)), 0) <= 1).sum(0)
                                                                       (base/distributions.py:60-63)
  vae_step_<tag>.npz        PoissonVAE lin|lin forward, loss, backward, one Adamax step
                            (main/vae.py:384-450, main/train_vae.py:346-362, base/adamax.py:34-100)
  scalars.npz               compute_n_exp / update_n table, softclamps, indicator grids, schedules
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from oracle import pvae_oracle as po
from oracle import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _draws(seed, offset, shape, n_exp):
    idx = np.arange(int(np.prod(shape)), dtype=np.uint32).reshape(shape)
    u = po.draw_uniforms(seed, offset, idx, n_exp)
    return u, po.exponential_from_uniform(u)


def gen_sampler(ref):
    rng = np.random.default_rng(20241)
    B, K, n_exp = 4, 48, 32
    for kind in po.INDICATORS:
        rec = {}
        log_rate = np.minimum(rng.normal(-1.0, 2.0, size=(B, K)), 5.0).astype(np.float32)
        w = rng.normal(size=(B, K)).astype(np.float32)
        seed, offset = 1234 + len(kind), 8
        u, e = _draws(seed, offset, (B, K), n_exp)
        rec.update(log_rate=log_rate, w=w, u=u, e=e, seed=seed, offset=offset, n_exp=n_exp)
        for temp in (1.0, 0.3, 0.05):
            lr = torch.tensor(log_rate, requires_grad=True)
            with ref_import.inject_exponentials([e]):
                dist = ref.dist.Poisson(log_rate=lr, temp=temp, indicator_approx=kind, n_exp=n_exp)
                z = dist.rsample()
            (g,) = torch.autograd.grad((z * torch.tensor(w)).sum(), lr)
            tag = f"t{temp:g}"
            rec[f"z_{tag}"] = z.detach().numpy()
            rec[f"g_{tag}"] = g.numpy()
            rec["rate"] = dist.rate.detach().numpy()
        np.savez_compressed(os.path.join(OUT, f"sampler_{kind}.npz"), **rec)


def gen_hard(ref):
    rng = np.random.default_rng(77)
    B, K, n_exp = 8, 64, 48
    log_rate = np.minimum(rng.normal(0.0, 2.0, size=(B, K)), 5.0).astype(np.float32)
    u, e = _draws(99, 4, (B, K), n_exp)
    with ref_import.inject_exponentials([e]):
        dist = ref.dist.Poisson(log_rate=torch.tensor(log_rate), temp=1.0, n_exp=n_exp)
        x = dist._exp.rsample((dist.n_exp,))  # step (1)
        times = torch.cumsum(x, dim=0)  # step (2)
    counts = (times <= 1).sum(0).float().numpy()
    np.savez_compressed(os.path.join(OUT, "hard_counts.npz"), log_rate=log_rate, u=u, e=e, seed=99, offset=4,
                        n_exp=n_exp, rate=dist.rate.numpy(), counts=counts, times=times.numpy())


def gen_vae(ref):
    cases = [
        dict(tag="k64_t1_b1", K=64, B=16, n_exp=20, temp=1.0, beta=1.0, exc_only=False, seed=3),
        dict(tag="k128_t01_b25", K=128, B=12, n_exp=24, temp=0.1, beta=2.5, exc_only=False, seed=4),
        dict(tag="k64_exc", K=64, B=8, n_exp=16, temp=0.5, beta=1.0, exc_only=True, seed=5),
    ]
    for c in cases:
        cfg_vae, cfg_tr = ref.cfg.default_configs("vH16", "poisson", "lin|lin")
        cfg_vae.update(n_latents=c["K"], exc_only=c["exc_only"], seed=c["seed"])
        cfg = ref.cfg.ConfigPoisVAE(save=False, **cfg_vae)
        model = ref.vae.PoissonVAE(cfg)
        model.n_exp.fill_(c["n_exp"])
        model.update_t(c["temp"])
        rng = np.random.default_rng(1000 + c["seed"])  # data seed != cfg.seed
        x = rng.standard_normal((c["B"], 1, 16, 16)).astype(np.float32)
        x = (x - x.mean((1, 2, 3), keepdims=True)) / x.std((1, 2, 3), keepdims=True)
        u, e = _draws(2024, 12, (c["B"], c["K"]), c["n_exp"])
        p_init = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
        # make rates non-trivial: shift the prior up so several arrivals land in [0, 1]
        with torch.no_grad():
            model.log_rate.add_(4.0)
        p0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
        optim = ref.adamax.Adamax(model.parameters(), lr=cfg_tr["lr"])
        optim.zero_grad(set_to_none=True)
        xt = torch.tensor(x)
        with ref_import.inject_exponentials([e]):
            dist, log_dr, z, y = model(xt)  # main/vae.py:384-391
        recon = model.loss_recon(y, xt)  # main/vae.py:68-72
        kl = model.loss_kl(log_dr)  # main/vae.py:446-450
        loss = torch.mean(recon + c["beta"] * torch.sum(kl, dim=1))  # main/train_vae.py:347-351
        loss.backward()
        grads = {n: p.grad.detach().clone().numpy() for n, p in model.named_parameters()}
        optim.step()
        p1 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
        rec = dict(x=x, u=u, e=e, seed=2024, offset=12, n_exp=c["n_exp"], temp=c["temp"], beta=c["beta"],
                   exc_only=int(c["exc_only"]), lr=cfg_tr["lr"],
                   log_dr=log_dr.detach().numpy(), rate=dist.rate.detach().numpy(), z=z.detach().numpy(),
                   y=y.detach().numpy(), recon=recon.detach().numpy(), kl=kl.detach().numpy(),
                   loss=loss.detach().numpy())
        rec["cfg_seed"] = c["seed"]
        for k, v in p_init.items():
            rec["init_" + k] = v
        for k, v in p0.items():
            rec["p0_" + k] = v
        for k, v in p1.items():
            rec["p1_" + k] = v
        for k, v in grads.items():
            rec["g_" + k] = v
        np.savez_compressed(os.path.join(OUT, f"vae_step_{c['tag']}.npz"), **rec)


def gen_exact(ref):
    """method='exact' (main/vae.py:74-93, main/train_vae.py:703-705): closed-form reconstruction loss, no sampling."""
    cfg_vae, cfg_tr = ref.cfg.default_configs("vH16", "poisson", "lin|lin")
    cfg_vae.update(n_latents=64, seed=6)
    model = ref.vae.PoissonVAE(ref.cfg.ConfigPoisVAE(save=False, **cfg_vae))
    with torch.no_grad():
        model.log_rate.add_(4.0)
    rng = np.random.default_rng(1006)
    x = rng.standard_normal((16, 1, 16, 16)).astype(np.float32)
    x = (x - x.mean((1, 2, 3), keepdims=True)) / x.std((1, 2, 3), keepdims=True)
    p0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    xt = torch.tensor(x)
    recon, dist, log_dr = model.loss_recon_exact(xt)
    kl = model.loss_kl(log_dr)
    loss = torch.mean(recon + 2.5 * torch.sum(kl, dim=1))
    loss.backward()
    rec = dict(x=x, beta=2.5, cfg_seed=6, recon=recon.detach().numpy(), kl=kl.detach().numpy(), loss=loss.detach().numpy(),
               rate=dist.rate.detach().numpy(), log_dr=log_dr.detach().numpy())
    for k, v in p0.items():
        rec["p0_" + k] = v
    for n, p in model.named_parameters():
        rec["g_" + n] = p.grad.detach().clone().numpy()
    np.savez_compressed(os.path.join(OUT, "vae_exact_k64.npz"), **rec)


def digest_stride(size: int) -> int:
    return 1 if size <= 4096 else max(8, -(-size // 16384))


def grad_digest(g: np.ndarray):
    """Small tensors in full; large ones as every digest_stride-th element plus the L2 norm (keeps fixtures small)."""
    flat = g.reshape(-1)
    return flat[::digest_stride(flat.size)].copy(), np.float64(np.linalg.norm(flat.astype(np.float64)))


def perturb_conv_params(model, seed):
    """Move lognorm / biases / LayerNorm affine off their trivial initial values (CPU generator: reproducible anywhere)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("lognorm") or name.endswith("bias") or "layer_norm" in name:
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def gen_conv(ref):
    """Conv encoder + linear decoder (BASELINE configs 4 and the encoder of 5): one full forward / loss / backward of the
    unmodified reference in eval mode (dropout off), with injected exponentials.  Parameters are NOT stored: the
    mirror model builds bit-identical ones from cfg.seed (tests/test_conv_oracle.py checks that against the reference
    and against the checksums stored here)."""
    cases = [dict(tag="cifar16", dataset="CIFAR16", B=8, n_exp=12, temp=0.7, beta=1.5, seed=7),
             dict(tag="mnist", dataset="MNIST", B=4, n_exp=10, temp=1.0, beta=1.0, seed=8),
             # conv decoder too (BASELINE config 5: conv+b|conv+b, MNIST-shaped 28x28, K = 512; and its 16x16 variant)
             dict(tag="mnist_convdec", dataset="MNIST", B=4, n_exp=10, temp=0.5, beta=1.0, seed=9, archi="conv+b|conv+b"),
             dict(tag="cifar16_convdec", dataset="CIFAR16", B=4, n_exp=10, temp=1.0, beta=2.0, seed=10, archi="conv+b|conv+b")]
    for c in cases:
        archi = c.get("archi", "conv+b|lin")
        cfg_vae, cfg_tr = ref.cfg.default_configs(c["dataset"], "poisson", archi)
        cfg_vae.update(seed=c["seed"], n_latents=512)
        model = ref.vae.PoissonVAE(ref.cfg.ConfigPoisVAE(save=False, **cfg_vae)).eval()
        perturb_conv_params(model, 100 + c["seed"])
        with torch.no_grad():
            model.log_rate.add_(3.0)
        model.n_exp.fill_(c["n_exp"])
        model.update_t(c["temp"])
        sz = model.cfg.input_sz
        rng = np.random.default_rng(2000 + c["seed"])
        x = rng.standard_normal((c["B"], 1, sz, sz)).astype(np.float32)
        if c["dataset"] == "MNIST":
            x = rng.uniform(0, 1, size=x.shape).astype(np.float32)  # ToTensor range, base/dataset.py:97-100
        K = model.cfg.n_latents
        u, e = _draws(2024, 16, (c["B"], K), c["n_exp"])
        xt = torch.tensor(x)
        enc_out = model.encode(xt).detach().numpy()  # main/vae.py:40-52
        with ref_import.inject_exponentials([e]):
            dist, log_dr, z, y = model(xt)
        recon = model.loss_recon(y, xt)
        kl = model.loss_kl(log_dr)
        loss = torch.mean(recon + c["beta"] * torch.sum(kl, dim=1))
        loss.backward()
        rec = dict(x=x, n_exp=c["n_exp"], temp=c["temp"], beta=c["beta"], cfg_seed=c["seed"], perturb_seed=100 + c["seed"],
                   philox_seed=2024, philox_offset=16, dataset=c["dataset"], archi=archi, enc_out=enc_out,
                   log_dr=log_dr.detach().numpy(), z=z.detach().numpy(), y=y.detach().numpy(), recon=recon.detach().numpy(),
                   kl=kl.detach().numpy(), loss=loss.detach().numpy(), rate=dist.rate.detach().numpy())
        names = []
        for n, p in model.named_parameters():
            names.append(n)
            d, nrm = grad_digest(p.grad.detach().numpy())
            rec["g_" + n], rec["gn_" + n] = d, nrm
            rec["ck_" + n] = np.array([p.detach().double().sum().item(), p.detach().double().abs().sum().item()])
        rec["param_names"] = np.array(names)
        np.savez_compressed(os.path.join(OUT, f"conv_{'vae' if 'archi' in c else 'enc'}_{c['tag']}.npz"), **rec)


def gen_scalars(ref):
    rec = {}
    rates = np.array([0.05, 0.1, 0.43, 1, 5, 20, 50, 100, 148.4, 200], dtype=np.float64)
    rec["n_exp_rates"] = rates
    rec["n_exp_p1e-5"] = np.array([ref.dist.compute_n_exp(float(r), 1e-5) for r in rates])
    rec["n_exp_p1e-3"] = np.array([ref.dist.compute_n_exp(float(r), 1e-3) for r in rates])
    x = torch.linspace(-40, 40, 801)
    rec["clamp_x"] = x.numpy()
    rec["softclamp_upper_10"] = ref.dist.softclamp_upper(x, 10.0).numpy()
    rec["softclamp_upper_5"] = ref.dist.softclamp_upper(x, 5.0).numpy()
    rec["softclamp_10_0"] = ref.dist.softclamp(x, 10.0, 0).numpy()
    l = torch.linspace(-3, 3, 601)
    rec["ind_l"] = l.numpy()
    for kind, fn in ref.dist._INDICATOR_FNS.items():
        rec[f"ind_{kind}"] = fn(l).numpy()
    rec["temp_lin_1000"] = ref.utils.temp_anneal_linear(1000, 1.0, 0.05, 0.5)
    rec["beta_lin_1000"] = ref.utils.beta_anneal_linear(1000, 2.5, 0.5, 0.0, 1e-4)
    np.savez_compressed(os.path.join(OUT, "scalars.npz"), **rec)


def gen_lr(ref):
    """Learning rate seen by every optimiser step of the reference loop: the reference's own Adamax
    (base/adamax.py) under torch's CosineAnnealingLR with the T_max override of base/train_base.py:282-287,
    driven exactly like main/train_vae.py:326-337 (warm-up overwrites param_group['lr']) and :391-400 (the
    scheduler only steps outside warm-up).  CosineAnnealingLR is recursive on the group's current lr, so the
    cosine starts from the LAST warm-up value, not from cfg.lr."""
    rec = {}
    cases = [("a", 0.005, 400, 0.1, 1e-5), ("b", 0.002, 90, 0.05, 1e-5), ("c", 0.002, 64, 0.0, 1e-5),
             ("d", 0.005, 26 * 30, 0.01, 1e-5), ("e", 0.005, 40, 0.01, 1e-5)]  # e: one warm-up step at lr 0 - the cosine never leaves ~0
    for tag, lr, n_iters, warmup_portion, eta_min in cases:
        w = torch.nn.Parameter(torch.zeros(3))
        optim = ref.adamax.Adamax(params=[w], lr=lr)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(optim, T_max=n_iters, eta_min=eta_min)
        sched.T_max = n_iters - n_iters * warmup_portion
        lrs = []
        for gstep in range(n_iters):
            progress = gstep / n_iters
            is_warming_up = progress < warmup_portion
            if is_warming_up:
                for group in optim.param_groups:
                    group['lr'] = lr * progress / warmup_portion
            lrs.append(optim.param_groups[0]['lr'])  # the lr optim.step() uses at this gstep
            w.grad = torch.ones(3)
            optim.step()
            if not is_warming_up:
                sched.step()
        rec[f"lr_{tag}"] = np.array(lrs, dtype=np.float64)
        rec[f"cfg_{tag}"] = np.array([lr, n_iters, warmup_portion, eta_min], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "lr_schedule.npz"), **rec)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ref = ref_import.load()
    if len(sys.argv) > 1:  # regenerate selected fixtures only: python oracle/make_golden.py lr scalars ...
        for name in sys.argv[1:]:
            globals()["gen_" + name](ref)
        return
    gen_scalars(ref)
    gen_lr(ref)
    gen_sampler(ref)
    gen_hard(ref)
    gen_vae(ref)
    gen_exact(ref)
    gen_conv(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
