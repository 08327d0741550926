"""CPU, world_size 2 over gloo: the data-parallel algebra of the step (contiguous batch shards, gradients
scaled by 1 / B_global, one sum all-reduce of the flat buffer, draws keyed by the global sample index)
reproduces the full-batch gradients the unmodified reference computed (tests/golden/vae_step_*.npz).
The per-shard compute stands in for the CUDA kernels with the numpy oracle - this test covers the host
logic in poissonvae_b200/dp.py, which is all that differs between N = 1 and N > 1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "vae_step_k64_t1_b1.npz")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pvae_oracle as po
        from poissonvae_b200 import dp
        g = np.load(GOLDEN)
        x = g["x"].reshape(g["x"].shape[0], -1)
        Bg = x.shape[0]
        assert dp.world_info() == (world, rank)
        lo, n = dp.shard_rows(Bg, world, rank)
        # the shard's draws: the oracle's generator keyed by GLOBAL latent index reproduces the fixture's rows
        K = g["p0_log_rate"].shape[1]
        idx = (np.arange(n * K, dtype=np.uint64) + lo * K).reshape(n, K)
        u = po.draw_uniforms(int(g["seed"]), int(g["offset"]), idx, int(g["n_exp"]))
        assert np.array_equal(u, g["u"][:, lo:lo + n])
        res = po.vae_step(x[lo:lo + n], g["p0_fc_enc.weight"], g["p0_fc_dec.weight"], g["p0_log_rate"][0],
                          g["e"][:, lo:lo + n], float(g["temp"]), float(g["beta"]), global_batch=Bg)
        flat = torch.tensor(np.concatenate([res["g_log_r"].ravel(), res["g_w_enc"].ravel(), res["g_w_dec"].ravel()]))
        stats = torch.tensor([float(res["loss"])])
        dp.allreduce_gradients(flat, stats)
        rmax = torch.tensor([float(res["rate"].max())])
        dp.allreduce_max(rmax)
        if rank == 0:
            np.savez(out, flat=flat.numpy(), loss=stats.numpy(), rmax=rmax.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gradients_match_full_batch_reference(tmp_path):
    out = str(tmp_path / "dp.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    from tests.helpers import assert_rel
    r, g = np.load(out), np.load(GOLDEN)
    want = np.concatenate([g["g_log_rate"].ravel(), g["g_fc_enc.weight"].ravel(), g["g_fc_dec.weight"].ravel()])
    assert_rel(r["flat"], want, what="all-reduced flat gradient")
    assert_rel(r["loss"][0], float(g["loss"]), what="loss")
    assert_rel(r["rmax"][0], float(g["rate"].max()), rtol=1e-6, what="r_max")


def test_shard_rows():
    from poissonvae_b200 import dp
    assert dp.shard_rows(65536, 8, 3) == (3 * 8192, 8192)
    assert dp.world_info() == (1, 0)
    with pytest.raises(ValueError):
        dp.shard_rows(10, 4, 0)


# ---- conv architecture: the same data-parallel algebra on the ~75-tensor flat gradient buffer ----------------------
GOLDEN_CONV = os.path.join(os.path.dirname(__file__), "golden", "conv_enc_cifar16.npz")


def _conv_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import conv_oracle, pvae_oracle as po
        from poissonvae_b200 import dp
        from tests.helpers import conv_fixture_model
        torch.set_num_threads(2)
        g = np.load(GOLDEN_CONV)
        model = conv_fixture_model(g)
        sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in model.state_dict().items()}
        x = torch.tensor(g["x"])
        Bg = x.shape[0]
        lo, n = dp.shard_rows(Bg, world, rank)
        K = model.cfg.n_latents
        idx = (np.arange(n * K, dtype=np.uint64) + lo * K).reshape(n, K)  # draws keyed by the GLOBAL latent index
        e = torch.tensor(po.exponential_from_uniform(po.draw_uniforms(int(g["philox_seed"]), int(g["philox_offset"]), idx, int(g["n_exp"]))))
        xs = x[lo:lo + n]
        sc = lambda v, c: c - torch.nn.functional.softplus(c - v)
        h = conv_oracle.conv_encoder(xs, sd, str(g["dataset"]))
        log_dr = sc(h, 10.0)
        rate = torch.exp(sc(sd["log_rate"] + log_dr, 5.0)) + torch.finfo(torch.float32).eps
        z = torch.sigmoid((1 - torch.cumsum(e / rate, 0)) / float(g["temp"])).sum(0)
        y = (z @ sd["fc_dec.weight"].T).view(xs.shape)
        kl = torch.exp(sd["log_rate"]) * (1 + torch.exp(log_dr) * (log_dr - 1))
        # this rank's share of the GLOBAL mean: gradients carry 1 / B_global, the all-reduce sums the shares
        loss = torch.sum(((y - xs) ** 2).sum(dim=[1, 2, 3]) + float(g["beta"]) * kl.sum(1)) / Bg
        loss.backward()
        names = [str(k) for k in g["param_names"]]
        flat = torch.cat([sd[k].grad.reshape(-1) for k in names])
        stats = torch.tensor([float(loss)])
        dp.allreduce_gradients(flat, stats)
        if rank == 0:
            np.savez(out, flat=flat.numpy(), loss=stats.numpy(), sizes=np.array([sd[k].numel() for k in names]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_conv_gradients_match_full_batch_reference(tmp_path):
    """conv+b|lin (BASELINE config 4): two gloo ranks' shard gradients, summed by one all-reduce of the flat buffer,
    equal the full-batch gradients of the unmodified reference (digest fixture)."""
    from tests.helpers import assert_grad_digest, assert_rel
    out = str(tmp_path / "dp_conv.npz")
    mp.spawn(_conv_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r, g = np.load(out), np.load(GOLDEN_CONV)
    assert_rel(r["loss"][0], float(g["loss"]), what="loss")
    o = 0
    for name, n in zip(g["param_names"], r["sizes"]):
        assert_grad_digest(r["flat"][o:o + n], g, str(name), 1e-5)
        o += int(n)
