"""GPU: the data-parallel product path - `pvae_allreduce_adamax` (csrc/peer_allreduce.cu: gradient sum over NVLink
peer memory fused with Adamax) and the NCCL alternative.

* one GPU: the kernel with world = 1 (its own buffer as the only peer) must reproduce pvae_adamax_step bit for bit,
  on both buffer parities, with the statistics tail passed through;
* two GPUs (skipped on a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py`): two ranks
  stepping on contiguous shards, with alternating skew between the ranks, end with bit-identical parameters on
  both ranks and within 2e-5 of a single GPU stepping on the full batch (tools/dp_check.py, now automated)."""
import ctypes
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_self_peer_exchange_equals_plain_adamax():
    from poissonvae_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    n, tail = 262656, 4
    g = torch.Generator(device="cuda").manual_seed(5)
    s = torch.cuda.current_stream().cuda_stream
    p0 = torch.randn(n, device=dev, generator=g) * 0.05
    state = {k: (p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)) for k in ("peer", "plain")}
    bufs = torch.zeros(2, n + tail, device=dev)
    flags = torch.zeros(64, device=dev, dtype=torch.int32)
    tail_out = torch.zeros(4, device=dev)
    for step in range(1, 5):  # both parities twice
        grad = torch.randn(n + tail, device=dev, generator=g) * (10.0 ** -(step % 3))
        par = step & 1
        bufs[par].copy_(grad)
        peer_g = (ctypes.c_void_p * 1)(bufs[par].data_ptr())
        peer_f = (ctypes.c_void_p * 1)(flags.data_ptr())
        p, m, u = state["peer"]
        _lib.check(lib.pvae_allreduce_adamax(p.data_ptr(), m.data_ptr(), u.data_ptr(), n, n + tail, peer_g, peer_f, 1, 0, step,
                                             tail_out.data_ptr(), None, 0.003, step, 0.9, 0.999, 1e-8, 0.0, s))
        p, m, u = state["plain"]
        _lib.check(lib.pvae_adamax_step(p.data_ptr(), grad.data_ptr(), m.data_ptr(), u.data_ptr(), n, 0.003, step, 0.9, 0.999,
                                        1e-8, 0.0, None, s))
        torch.cuda.synchronize()
        assert torch.equal(tail_out, grad[n:])
        for a, b in zip(state["peer"], state["plain"]):
            assert torch.equal(a, b), step
    assert int(flags[0]) == 4  # the last token published


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, mode, out_dir):
    import torch.distributed as dist

    from tools.dp_check import run
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p, l, peer = run(world, rank, 6, 2048, 512, mode)
        torch.save({"p": p, "l": l, "peer": peer}, os.path.join(out_dir, f"{mode}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_two_rank_exchange_matches_single_gpu(mode, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from tools.dp_check import run
    torch.cuda.set_device(0)
    ref_p, ref_l, _ = run(1, 0, 6, 2048, 512, mode)
    mp.spawn(_dp_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(str(tmp_path), f"{mode}_{k}.pt")) for k in range(2)]
    assert r[0]["peer"] == (mode == "peer")  # the path under test really ran
    assert torch.equal(r[0]["p"], r[1]["p"])  # ranks bit-identical
    assert torch.equal(r[0]["l"], r[1]["l"])
    perr = ((r[0]["p"] - ref_p).abs().max() / ref_p.abs().max()).item()
    lerr = ((r[0]["l"] - ref_l).abs() / ref_l.abs()).max().item()
    print(f"{mode}: 2 ranks vs 1 GPU: parameters {perr:.2e}, losses {lerr:.2e}")
    assert perr < 2e-5 and lerr < 2e-5
