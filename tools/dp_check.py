"""Multi-GPU check: N ranks each stepping on a contiguous shard of the global batch end up with the
parameters a single GPU gets from the full batch (the Philox stream is keyed by the global sample
index; gradients are summed by one NCCL all-reduce).

  python tools/dp_check.py --mode single --out /tmp/ref.pt
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      tools/dp_check.py --mode dp --ref /tmp/ref.pt
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(world, rank, steps, Bg, K, dp_mode='auto', archi='lin|lin', dataset='vH16'):
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    cfg_vae, _ = default_configs(dataset, 'poisson', archi)
    cfg_vae.update(n_latents=K, seed=0)
    model = PoissonVAE(ConfigPoisVAE(**cfg_vae)).cuda().eval()  # eval: dropout off (its mask is not keyed by the sample index)
    with torch.no_grad():
        model.log_rate.add_(3.0 if model.cfg.dec_type == 'lin' else 0.0)
    model.update_n(20.0)
    cfg = ConfigTrainVAE(lr=0.005 if model.cfg.dec_type == 'lin' else 0.0005, epochs=100, batch_size=Bg // world, warmup_portion=0.0, scheduler_type=None,
                         grad_clip=200.0, kl_beta=2.5, kl_anneal_portion=0.0, kl_const_portion=0.0,
                         temp_anneal_portion=0.0, temp_stop=0.3)
    cfg.grad_clip = None if dp_mode == 'peer' else cfg.grad_clip
    tr = TrainerVAE(model, cfg, dp_mode=dp_mode)
    if rank == 0 and world > 1:
        print("gradient exchange:", "peer memory (fused all-reduce + Adamax)" if tr.peer is not None else "NCCL all-reduce")
    g = torch.Generator().manual_seed(4321)
    data = torch.randn(steps, Bg, model.cfg.input_sz ** 2, generator=g)
    losses = []
    for i in range(steps):
        lo = rank * (Bg // world)
        x = data[i, lo:lo + Bg // world].cuda()
        if world > 1 and (i + rank) % 2 == 1:
            torch.cuda._sleep(3_000_000)  # ~1.5 ms: ranks reach the exchange at very different times, alternately
        losses.append(tr.train_step(x, i, philox=(77, 4 * i)).clone())
    torch.cuda.synchronize()
    return tr.flat_p.cpu(), torch.stack(losses).cpu(), tr.peer is not None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["single", "dp"])
    ap.add_argument("--out", default="/tmp/dp_ref.pt")
    ap.add_argument("--ref", default="/tmp/dp_ref.pt")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--latents", type=int, default=512)
    ap.add_argument("--dp_mode", default="auto", choices=["auto", "peer", "nccl"])
    ap.add_argument("--archi", default="lin|lin")
    ap.add_argument("--dataset", default="vH16")
    a = ap.parse_args()
    if a.mode == "single":
        torch.cuda.set_device(0)
        p, l, _ = run(1, 0, a.steps, a.batch, a.latents, a.dp_mode, a.archi, a.dataset)
        torch.save({"p": p, "l": l}, a.out)
        print("single-GPU reference saved:", l[:, 0].tolist())
        return
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p, l, _ = run(world, rank, a.steps, a.batch, a.latents, a.dp_mode, a.archi, a.dataset)
    ref = torch.load(a.ref)
    dp_err = ((p - ref["p"]).abs().max() / ref["p"].abs().max()).item()
    l_err = ((l - ref["l"]).abs() / ref["l"].abs()).max().item()
    # all ranks hold identical parameters
    mine = p.cuda()
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = torch.equal(lo, hi)
    if rank == 0:
        print(f"world={world}: max rel param diff vs single GPU {dp_err:.2e}, loss diff {l_err:.2e}, ranks identical: {same}")
    dist.destroy_process_group()
    tol = 2e-5 if a.archi == 'lin|lin' else 2e-4  # conv: ~25 GEMMs deep, summation order differs between 1 and N shards
    assert dp_err < tol and l_err < tol and same
    if rank == 0:
        print("DP CHECK OK")


if __name__ == "__main__":
    main()
