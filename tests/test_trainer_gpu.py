"""GPU: the explicit training step (kernels + closed-form backward + Adamax) against the fixtures the
unmodified reference produced with autograd and its own Adamax (oracle/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pvae_oracle as po
from tests.helpers import assert_rel

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "vae_step_*.npz")))


def make_trainer(g, **over):
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    K = g["p0_log_rate"].shape[1]
    model = PoissonVAE(ConfigPoisVAE(n_latents=K, prior_clamp=-4, exc_only=bool(g["exc_only"]), seed=int(g["cfg_seed"])))
    model.load_state_dict({k: torch.tensor(g["p0_" + k]) for k in model.state_dict()})
    model = model.cuda()
    model.n_exp.fill_(int(g["n_exp"]))
    model._n_exp_host = int(g["n_exp"])
    kw = dict(lr=float(g["lr"]), epochs=1, batch_size=g["x"].shape[0], warmup_portion=0.0, scheduler_type=None,
              grad_clip=None, kl_beta=float(g["beta"]), kl_anneal_portion=0.0, kl_const_portion=0.0,
              temp_anneal_portion=0.0, temp_stop=float(g["temp"]))
    kw.update(over)
    return TrainerVAE(model, ConfigTrainVAE(**kw))


@pytest.mark.parametrize("path", GOLDEN)
def test_train_step_matches_reference(path):
    g = np.load(path)
    tr = make_trainer(g)
    x = torch.tensor(g["x"], device="cuda")
    loss = tr.train_step(x, 0, philox=(int(g["seed"]), int(g["offset"])))
    torch.cuda.synchronize()
    assert_rel(loss[0].item(), float(g["loss"]), rtol=2e-5, what="loss")
    assert_rel(loss[1].item(), float(g["recon"].mean()), rtol=2e-5, what="mse")
    assert_rel(loss[2].item(), float(g["kl"].sum(1).mean()), rtol=2e-5, what="kl")
    for name in ("log_rate", "fc_enc.weight", "fc_dec.weight"):
        assert_rel(tr.gviews[name].cpu().numpy(), g["g_" + name], rtol=2e-5, what="grad " + name)
        # first Adamax step: p -= lr * g / (|g| + eps); the ratio amplifies the GEMMs' ~1e-6 relative gradient
        # error wherever |g| ~ eps, so the parameters carry the GEMM-dependent tolerance too
        assert_rel(tr.views[name].cpu().numpy(), g["p1_" + name], rtol=2e-5, what="adamax " + name)
    # the module's parameters are views of the flat buffer: state_dict sees the update
    sd = tr.model.state_dict()
    assert torch.equal(sd["fc_dec.weight"], tr.views["fc_dec.weight"])


def test_grad_clip_and_second_step_follow_the_oracle():
    g = np.load(GOLDEN[0])
    tr = make_trainer(g, grad_clip=0.05)  # far below the actual norm: the clip must engage
    x = torch.tensor(g["x"], device="cuda")
    ph = (int(g["seed"]), int(g["offset"]))
    flat_p0 = tr.flat_p.clone()
    tr.train_step(x, 0, philox=ph)
    grads = tr.flat_g.cpu().numpy().astype(np.float64)
    norm = float(np.sqrt((grads ** 2).sum()))
    b = tr._buffers(x.shape[0])
    assert_rel(b["norm"][0].item(), norm, rtol=1e-5, what="norm")
    coef = min(1.0, 0.05 / (norm + 1e-6))
    assert_rel(b["norm"][1].item(), coef, rtol=1e-5, what="clip coefficient")
    p1, m1, u1 = po.adamax_step(flat_p0.cpu().numpy(), (grads * coef).astype(np.float32), np.zeros_like(grads, np.float32),
                                np.zeros_like(grads, np.float32), 1, float(g["lr"]))
    assert_rel(tr.flat_p.cpu().numpy(), p1, rtol=1e-6, what="params after step 1")
    # a second step continues the Adamax state (step = 2)
    tr.train_step(x, 0, philox=ph)
    g2 = tr.flat_g.cpu().numpy().astype(np.float64)
    c2 = min(1.0, 0.05 / (float(np.sqrt((g2 ** 2).sum())) + 1e-6))
    p2, _, _ = po.adamax_step(p1, (g2 * c2).astype(np.float32), m1, u1, 2, float(g["lr"]))
    assert_rel(tr.flat_p.cpu().numpy(), p2, rtol=1e-6, what="params after step 2")


def test_schedules_and_epoch_loop():
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    model = PoissonVAE(ConfigPoisVAE(n_latents=64, prior_clamp=-4, seed=0)).cuda()
    # warm-up of 4 steps; the reference's cosine then starts from the LAST warm-up value, 0.75 lr (with the default 1 %
    # of 40 iterations the single warm-up step has lr 0 and the reference's recursive scheduler never leaves ~0)
    cfg = ConfigTrainVAE(lr=0.005, epochs=10, batch_size=48, grad_clip=None, kl_const_portion=0.0, warmup_portion=0.1)
    tr = TrainerVAE(model, cfg, n_batches_per_epoch=4)
    assert tr.n_iters == 40
    assert np.array_equal(tr.temperatures, po.temp_anneal_linear(40, 1.0, 0.05, 0.5))
    assert np.array_equal(tr.betas, po.beta_anneal_linear(40, 1.0, 0.5, 0.0, 1e-4))
    assert tr.lr_at(0) == 0.0 and tr.lr_at(4) == tr.lr_at(3) == 0.005 * 0.75 and tr.lr_at(39) < tr.lr_at(5)
    with torch.no_grad():
        model.log_rate.add_(4.0)  # rates ~0.1-1 so that the decoder sees spikes from step 0
    data = torch.randn(192, 1, 16, 16, device="cuda")
    before = model.fc_dec.weight.detach().clone()
    n0 = int(model.n_exp)
    nelbo = [tr.iteration(data, epoch) for epoch in range(10)]
    assert all(np.isfinite(nelbo)) and np.mean(nelbo[-3:]) < np.mean(nelbo[:3])
    assert not torch.equal(before, model.fc_dec.weight)
    assert int(model.n_exp) < n0  # update_n from the epoch's mean rate.max(), main/train_vae.py:456-457


def test_fit_host_equals_resident_steps():
    """Pinned host batches through the chunked, double-buffered feeder == the same steps on resident batches."""
    g = np.load(GOLDEN[0])
    host = torch.randn(11, 16, 256, generator=torch.Generator().manual_seed(5)).pin_memory()
    results = []
    for use_host in (False, True):
        tr = make_trainer(g)
        torch.cuda.manual_seed(99)
        if use_host:
            losses = tr.fit_host(host, 0, chunk=4)
            torch.cuda.synchronize()
        else:
            losses = torch.stack([tr.train_step(host[i].cuda(), i).clone() for i in range(11)]).cpu()
        results.append((losses.clone(), tr.flat_p.clone()))
    assert torch.equal(results[0][0], results[1][0])
    assert torch.equal(results[0][1], results[1][1])


def test_exact_method_step_matches_reference(golden_dir):
    """One training step with method='exact' reproduces the reference's loss and (through Adamax) its gradients' effect."""
    g = np.load(os.path.join(golden_dir, "vae_exact_k64.npz"))
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    model = PoissonVAE(ConfigPoisVAE(n_latents=64, prior_clamp=-4, seed=int(g["cfg_seed"])))
    model.load_state_dict({k: torch.tensor(g["p0_" + k]) for k in model.state_dict()})
    model = model.cuda()
    cfg = ConfigTrainVAE(method='exact', lr=0.005, epochs=1, batch_size=16, warmup_portion=0.0, scheduler_type=None,
                         grad_clip=None, kl_beta=float(g["beta"]), kl_anneal_portion=0.0, kl_const_portion=0.0,
                         temp_anneal_portion=0.0)
    tr = TrainerVAE(model, cfg)
    p0 = tr.flat_p.clone()
    loss = tr.train_step(torch.tensor(g["x"], device="cuda"), 0)
    assert_rel(loss[0].item(), float(g["loss"]), rtol=2e-5, what="loss")
    for name in ("log_rate", "fc_enc.weight", "fc_dec.weight"):
        assert_rel(tr.gviews[name].cpu().numpy(), g["g_" + name], rtol=2e-5, what="grad " + name)
    assert not torch.equal(p0, tr.flat_p)


def test_validate_pass_hard_counts():
    """Evaluation (main/train_vae.py:461-601): hard counts at temp = 0, r2 / mse / kl per sample, one host transfer."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    model = PoissonVAE(ConfigPoisVAE(n_latents=128, prior_clamp=-4, seed=2)).cuda()
    with torch.no_grad():
        model.log_rate.add_(3.5)
    model.update_n(20.0)
    tr = TrainerVAE(model, ConfigTrainVAE(batch_size=64, epochs=1, grad_clip=None))
    data = torch.randn(200, 1, 16, 16, device="cuda")
    d, loss, etc = tr.validate(data, full_data=True)
    assert d['z'].shape == (200, 128) and np.array_equal(d['z'], np.round(d['z'])) and d['z'].max() >= 1
    assert d['x'].shape == d['y'].shape == (200, 1, 16, 16)
    x, y = d['x'].reshape(200, -1).astype(np.float64), d['y'].reshape(200, -1).astype(np.float64)
    mse = ((y - x) ** 2).sum(1)
    assert_rel(loss['mse'], mse, what="mse")
    r2 = 1 - mse / (((x - x.mean(1, keepdims=True)) ** 2).sum(1) + np.finfo(np.float32).eps)
    assert_rel(loss['r2'], r2, what="r2")
    assert_rel(loss['elbo'], loss['mse'] + loss['kl'], rtol=1e-6, what="elbo")
    assert loss['kl_diag'].shape == (128,) and etc['r*dr'].shape == (200, 128) and etc['log_dr'].shape == (200, 128)
    # the decoder sees the counts: y = z W_dec^T
    w = model.fc_dec.weight.detach().cpu().numpy().astype(np.float64)
    assert_rel(y, d['z'].astype(np.float64) @ w.T, rtol=2e-5, what="decode(hard counts)")


def test_graph_replay_matches_eager_steps():
    """CUDA-graph replay of the fused step (device-side step counter, lr schedule and Philox offset) against the
    eager launches: same parameters and losses step after step, through a change of n_exp (re-capture), a restore of
    the optimiser state (counters rewritten) and the hand-over back to the eager path."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    g = torch.Generator().manual_seed(9)
    data = torch.randn(12, 64, 256, generator=g).cuda()
    out = {}
    for mode in (True, False):
        model = PoissonVAE(ConfigPoisVAE(n_latents=128, prior_clamp=-4, seed=3)).cuda()
        with torch.no_grad():
            model.log_rate.add_(3.0)
        model.update_n(30.0)
        tr = TrainerVAE(model, ConfigTrainVAE(lr=0.01, epochs=40, batch_size=64, temp_anneal_portion=0.0, temp_stop=0.5,
                                              kl_anneal_portion=0.0, kl_const_portion=0.0, grad_clip=50.0, warmup_portion=0.1),
                        n_batches_per_epoch=1)
        tr.graph = mode
        torch.cuda.manual_seed(77)
        losses = []
        for i in range(8):  # crosses the lr warm-up boundary (clip threshold changes -> another graph)
            if i == 5:
                model.update_n(12.0)
            losses.append(tr.train_step(data[i], i).clone())
        snap = (tr.flat_p.clone(), tr.exp_avg.clone(), tr.exp_inf.clone(), tr.opt_step)
        for i in range(8, 10):
            losses.append(tr.train_step(data[i], i).clone())
        tr.flat_p.copy_(snap[0]), tr.exp_avg.copy_(snap[1]), tr.exp_inf.copy_(snap[2])  # rewind two steps
        tr.opt_step = snap[3]
        torch.cuda.default_generators[0].set_offset(torch.cuda.default_generators[0].get_offset() - 8)
        for i in range(8, 10):
            losses.append(tr.train_step(data[i], i).clone())
        tr.graph = False  # hand over to the eager path
        losses.append(tr.train_step(data[10], 10).clone())
        if mode:
            assert len(tr._graphs) >= 3
        out[mode] = (tr.flat_p.clone(), torch.stack(losses))
    assert_rel(out[True][1].cpu().numpy(), out[False][1].cpu().numpy(), rtol=1e-6, what="losses graph vs eager")
    assert_rel(out[True][0].cpu().numpy(), out[False][0].cpu().numpy(), rtol=1e-6, what="parameters graph vs eager")
    l = out[True][1]
    assert torch.equal(l[8:10], l[10:12])  # the rewound steps reproduce themselves bit for bit


@pytest.mark.parametrize("bias", [False, True])
def test_fused_update_is_bit_identical(bias):
    """pvae_lin_step_train (gradient partials reduced inside the Adamax kernel) against the stand-alone reductions +
    Adamax: same parameters, optimiser state, stored gradients and losses, bit for bit."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    g = torch.Generator().manual_seed(19)
    data = torch.randn(5, 1024, 256, generator=g).cuda()
    from poissonvae_b200 import _lib
    lib = _lib.load()
    out = {}
    for fuse in (True, False, "three kernels"):
        lib.pvae_set_fuse_reduce(0 if fuse == "three kernels" else 1)
        model = PoissonVAE(ConfigPoisVAE(n_latents=256, prior_clamp=-4, seed=4, enc_bias=bias)).cuda()
        with torch.no_grad():
            model.log_rate.add_(3.0)
        model.update_n(25.0)
        tr = TrainerVAE(model, ConfigTrainVAE(lr=0.01, epochs=40, batch_size=1024, temp_anneal_portion=0.0, temp_stop=0.7,
                                              kl_anneal_portion=0.0, kl_const_portion=0.0, grad_clip=None), n_batches_per_epoch=1)
        tr.fuse_update = fuse is True
        torch.cuda.manual_seed(5)
        c0 = lib.pvae_launch_count()
        losses = [tr.train_step(data[i], i).clone() for i in range(5)]
        per_step = (lib.pvae_launch_count() - c0) // 5
        out[fuse] = (tr.flat_p.clone(), tr.exp_avg.clone(), tr.exp_inf.clone(), tr.flat_g.clone(), torch.stack(losses))
        # fused update: 9; one reduction kernel + Adamax: 10; two split-K reductions + column sums (+ bias) + Adamax: 12 (13);
        # + 1 with operand planes when the caller does not bring x's low plane (the step splits x itself);
        # + 1 when the stand-alone Adamax kernel does not keep the parameters' low plane (rebuilt at the next step's head)
        extra = (1 if tr.planes else 0) + (1 if (tr.planes and fuse is not True) else 0)
        assert per_step == {True: 9, False: 10, "three kernels": 13 if bias else 12}[fuse] + extra, per_step
    lib.pvae_set_fuse_reduce(1)
    for other in (False, "three kernels"):
        for a, b in zip(out[True], out[other]):
            assert torch.equal(a, b)


def test_graph_replay_fills_the_same_r_max_slots_and_n_exp():
    """a replayed graph must leave every step's rate.max() in that step's slot of the epoch ring (the
    eager path writes slot gstep % n directly), so that iteration()'s mean of the per-step maxima - and the n_exp
    update_n derives from it (main/train_vae.py:388, 456-457) - do not depend on the graph switch."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    data = torch.randn(6 * 64, 1, 16, 16, generator=torch.Generator().manual_seed(3)).cuda()
    out = {}
    for mode in (True, False):
        model = PoissonVAE(ConfigPoisVAE(n_latents=128, prior_clamp=-4, seed=3)).cuda()
        with torch.no_grad():
            model.log_rate.add_(5.0)
        tr = TrainerVAE(model, ConfigTrainVAE(lr=0.02, epochs=4, batch_size=64, temp_anneal_portion=0.0, temp_stop=0.5,
                                              kl_anneal_portion=0.0, kl_const_portion=0.0, grad_clip=None, warmup_portion=0.0),
                        n_batches_per_epoch=6)
        tr.graph = mode
        torch.manual_seed(11)  # iteration() draws its permutation from the global generators
        torch.cuda.manual_seed(11)
        n_exps, rings = [], []
        for epoch in range(2):
            tr.iteration(data, epoch)
            rings.append(tr.r_max.clone())
            n_exps.append(int(model.n_exp))
        if mode:
            assert len(tr._graphs) >= 1
        out[mode] = (n_exps, torch.stack(rings))
    assert out[True][0] == out[False][0], out
    assert (out[True][1] > 0).all()  # every slot of the ring was written, none holds a running maximum of the epoch
    assert torch.equal(out[True][1], out[False][1])


@pytest.mark.parametrize("bias", [False, True])
def test_operand_planes_step_is_bit_identical(bias):
    """The step with TF32 operand planes (low parts of x, W, z, g_y, g_h written once by their producers and loaded by
    TMA) against the step whose GEMMs split every stage in shared memory: same parameters, optimiser state, gradients
    and losses bit for bit - eager with the fused update, eager with stand-alone Adamax, x_lo supplied or not, after a
    parameter edit through torch, and through graph replay."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    g = torch.Generator().manual_seed(23)
    data = torch.randn(6, 1024, 256, generator=g).cuda()
    out = {}
    for mode in ("split", "planes", "planes+x_lo", "planes unfused", "planes graph"):
        model = PoissonVAE(ConfigPoisVAE(n_latents=512, prior_clamp=-4, seed=4, enc_bias=bias)).cuda()
        with torch.no_grad():
            model.log_rate.add_(3.0)
        model.update_n(25.0)
        tr = TrainerVAE(model, ConfigTrainVAE(lr=0.01, epochs=40, batch_size=1024, temp_anneal_portion=0.0, temp_stop=0.7,
                                              kl_anneal_portion=0.0, kl_const_portion=0.0, grad_clip=None), n_batches_per_epoch=1)
        tr.use_planes = mode != "split"
        tr.fuse_update = mode != "planes unfused"
        tr.graph = mode == "planes graph"
        torch.cuda.manual_seed(5)
        losses = []
        for i in range(6):
            if i == 3:
                with torch.no_grad():
                    model.fc_dec.weight.mul_(1.01)  # an edit through torch: the low plane must follow
            x_lo = tr.split_planes(data[i]) if mode == "planes+x_lo" else None
            losses.append(tr.train_step(data[i], i, x_lo=x_lo).clone())
        if mode == "planes graph":
            assert len(tr._graphs) >= 1
        out[mode] = (tr.flat_p.clone(), tr.exp_avg.clone(), tr.exp_inf.clone(), tr.flat_g.clone(), torch.stack(losses))
        if mode != "split":
            lo = tr.split_planes(tr.flat_p)
            if mode != "planes unfused":  # the fused optimiser kernel keeps the low plane current by itself
                assert torch.equal(tr.flat_p_lo, lo)
    for mode in out:
        for a, b in zip(out["split"], out[mode]):
            assert torch.equal(a, b), mode


def test_coefficient_mode_matches_the_recompute_backward():
    """pvae_set_step_coef(1): the forward sampler leaves d z / d u, the clamp derivative, the KL gradient term and per-block
    KL column sums, the backward is a streaming kernel.  Same step as the default (backward re-evaluates the prologue
    from h) up to the rounding of the regrouped products: losses equal bit for bit (the forward's z and KL are the
    same), gradients and parameters within 1e-6 of their scale, fused and unfused updates bit-identical to each other."""
    from poissonvae_b200 import _lib
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    lib = _lib.load()
    data = torch.randn(4, 1024, 256, generator=torch.Generator().manual_seed(29)).cuda()
    out = {}
    try:
        for mode in ("default", "coef", "coef unfused"):
            lib.pvae_set_step_coef(0 if mode == "default" else 1)
            model = PoissonVAE(ConfigPoisVAE(n_latents=512, prior_clamp=-4, seed=4, enc_bias=True)).cuda()
            with torch.no_grad():
                model.log_rate.add_(3.0)
            model.update_n(25.0)
            tr = TrainerVAE(model, ConfigTrainVAE(lr=0.01, epochs=40, batch_size=1024, temp_anneal_portion=0.0, temp_stop=0.7,
                                                  kl_anneal_portion=0.0, kl_const_portion=0.0, kl_beta=2.0, grad_clip=None),
                            n_batches_per_epoch=1)
            tr.fuse_update = mode != "coef unfused"
            torch.cuda.manual_seed(5)
            losses = [tr.train_step(data[i], i).clone() for i in range(4)]
            out[mode] = (torch.stack(losses), tr.flat_g.clone(), tr.flat_p.clone())
    finally:
        lib.pvae_set_step_coef(0)
    assert torch.equal(out["default"][0][0], out["coef"][0][0])  # first step: same parameters, same forward
    for a, b in zip(out["coef"], out["coef unfused"]):
        assert torch.equal(a, b)
    for name, a, b in zip(("losses", "gradient", "parameters"), out["default"], out["coef"]):
        assert_rel(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-6, what=name)


@pytest.mark.parametrize("clip", [None, 0.5])
def test_iteration_statistics_match_a_synchronous_replay(clip):
    """TrainerVAE.iteration fills the reference's per-gstep statistics (main/train_vae.py:373-408) and its logging dictionary
    (:410-445: nelbo / per-dim averages, kl_diag -> n_active, grad norm) from a device ring read once per log_freq steps.
    Replayed here step by step with a host read after every step and kl_diag from the torch formulas (main/vae.py:405,
    446-450) on the parameters the step saw."""
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    B, n_batches, log_freq, K = 256, 7, 3, 128
    data = torch.randn(B * n_batches, 1, 16, 16, generator=torch.Generator().manual_seed(3)).cuda()

    def make():
        model = PoissonVAE(ConfigPoisVAE(n_latents=K, prior_clamp=-4, seed=2)).cuda()
        with torch.no_grad():
            model.log_rate.add_(4.0)
        model.update_n(15.0)
        cfg = ConfigTrainVAE(lr=0.01, epochs=3, batch_size=B, temp_anneal_portion=0.5, kl_anneal_portion=0.3, kl_const_portion=0.0,
                             warmup_portion=0.2, grad_clip=clip, log_freq=log_freq)
        return model, TrainerVAE(model, cfg, n_batches_per_epoch=n_batches)

    model, tr = make()
    torch.manual_seed(21)
    torch.cuda.manual_seed(21)
    for epoch in range(2):
        tr.iteration(data, epoch)
    # replay
    model2, tr2 = make()
    torch.manual_seed(21)
    torch.cuda.manual_seed(21)
    sp = torch.nn.functional.softplus
    want_log, want_rmax, want_grad = [], {}, {}
    for epoch in range(2):
        perm = torch.randperm(data.shape[0], device=data.device)
        tr2.r_max.zero_()
        nelbo_sum, nelbo_cnt, rsum, pk, pm, gsum, gcnt = 0.0, 0, 0.0, 0.0, 0.0, 0.0, 0
        for i in range(n_batches):
            g = epoch * n_batches + i
            x = data[perm[i * B:(i + 1) * B]]
            with torch.no_grad():
                h = x.reshape(B, -1) @ model2.fc_enc.weight.T
                log_dr = 10.0 - sp(10.0 - h)
                kl_diag = (torch.exp(model2.log_rate) * (1 + torch.exp(log_dr) * (log_dr - 1))).mean(0)
            loss, mse, kl = tr2.train_step(x, g).tolist()
            nelbo_sum, nelbo_cnt = nelbo_sum + mse + kl, nelbo_cnt + 1
            rsum += float(tr2.r_max[g % n_batches])
            pk, pm = pk + kl / K, pm + mse / 256
            want_rmax[g] = rsum / (i + 1)
            if clip is not None:
                gn = float(tr2._buffers(B)['norm'][0])
                want_grad[g] = gn
                gsum, gcnt = gsum + gn, gcnt + 1
            if g > 0 and g % log_freq == 0:
                want_log.append(dict(g=g, kl=kl, mse=mse, nelbo=nelbo_sum / nelbo_cnt, pk=pk / (i + 1), pm=pm / (i + 1),
                                     n_active=int((kl_diag > 0.01).sum()), gn=gsum / max(gcnt, 1)))
            if g % 100 == 0:
                nelbo_sum, nelbo_cnt, gsum, gcnt = 0.0, 0, 0.0, 0
        tr2.model.update_n(float(tr2.r_max[:n_batches].mean()))
    assert sorted(tr.stats['lr']) == list(range(2 * n_batches))
    for g in range(2 * n_batches):
        warming = g / tr.n_iters < 0.2
        assert tr.stats['lr'][g] == (tr.lr_at(g) if warming else tr.lr_at(g + 1))
        assert tr.stats['temp'][g] == float(tr.temperatures[g])
        assert abs(tr.stats['r_max'][g] - want_rmax[g]) <= 1e-6 * want_rmax[g]
        if clip is not None:
            assert abs(tr.stats['grad'][g] - want_grad[g]) <= 1e-5 * want_grad[g]
            assert (g in tr.stats['loss']) == (want_grad[g] > clip)
    assert tr.stats['n_exp'][0] == 15 or tr.stats['n_exp'][0] > 0
    assert [e['gstep'] for e in tr.log] == [w['g'] for w in want_log] and len(tr.log) >= 3
    for got, w in zip(tr.log, want_log):
        for key, ref in (('train/loss_kl', w['kl']), ('train/loss_mse', w['mse']), ('train/nelbo_avg', w['nelbo']),
                         ('train/perdim_kl', w['pk']), ('train/perdim_mse', w['pm'])):
            assert abs(got[key] - ref) <= 2e-5 * abs(ref), (key, got[key], ref)
        assert abs(got['train/n_active'] - w['n_active']) <= 1 and got['train/n_active_ratio'] == got['train/n_active'] / K
        assert 0 < got['train/n_active'] <= K
        if clip is not None:
            assert abs(got['train/grad_norm'] - w['gn']) <= 1e-5 * w['gn']
        assert got['coeffs/n_exp'] > 0 and got['coeffs/beta'] > 0
    assert torch.equal(tr.flat_p, tr2.flat_p)  # the ring did not change the training itself


def test_one_wave_sampler_schedule_is_bit_identical_in_the_step():
    """pvae_set_pool(3), the default: the training forward with operand planes runs the one-wave schedule, parking the state of
    long chains in the quad's own output slots (z, z_lo, dz_acc) so that lane groups continue them exactly as the tile-pool
    schedule does.  Rates high enough that many chains are handed over; several phase-A caps; losses, gradients and updated
    parameters equal those of schedule 1 bit for bit."""
    from poissonvae_b200 import _lib
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    lib = _lib.load()
    data = torch.randn(3, 2048, 256, generator=torch.Generator().manual_seed(31)).cuda()
    try:
        for cap, temp, K in ((8, 0.3, 512), (3, 1.0, 512), (8, 0.5, 1024), (5, 1.0, 520)):  # K = 1024: two column passes; 520: ragged
            lib.pvae_set_phase_a_events(cap)
            out = {}
            for pool in (1, 3):
                lib.pvae_set_pool(pool)
                model = PoissonVAE(ConfigPoisVAE(n_latents=K, prior_clamp=-4, seed=6)).cuda()
                with torch.no_grad():
                    model.log_rate.add_(4.5)  # rates up to tens of events per chain
                model.update_n(60.0)
                tr = TrainerVAE(model, ConfigTrainVAE(lr=0.01, epochs=40, batch_size=2048, temp_anneal_portion=0.0, temp_stop=temp,
                                                      kl_anneal_portion=0.0, kl_const_portion=0.0, grad_clip=None),
                                n_batches_per_epoch=1)
                assert tr.planes
                torch.cuda.manual_seed(8)
                losses = [tr.train_step(data[i], i).clone() for i in range(3)]
                out[pool] = (torch.stack(losses), tr.flat_g.clone(), tr.flat_p.clone())
            for name, a, b in zip(("losses", "gradient", "parameters"), out[1], out[3]):
                assert torch.equal(a, b), (name, cap, temp, K)
    finally:
        lib.pvae_set_pool(3)
        lib.pvae_set_phase_a_events(8)
