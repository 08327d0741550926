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
