"""Per-kernel LIVE durations of the fused lin|lin step (warm L2, real predecessor in the stream): CUDA events between the
main-stream launches (pvae_set_step_trace).  The events defeat the programmatic overlap, so the sum is a little above the
step time bench.py reports; the split is what this is for.

  python tools/trace_step.py [--batch 4096 --latents 512 --steps 30] [--no_planes] [--tf32]
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NAMES = ["x split", "enc GEMM", "sampler fwd", "dec GEMM", "residual", "dgrad GEMM", "sampler bwd", "enc wgrad GEMM",
         "join + reduce + Adamax"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--latents", type=int, default=512)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--temp", type=float, default=1.0)
    ap.add_argument("--no_planes", action="store_true")
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--no_overlap", action="store_true")
    ap.add_argument("--no_carveout", action="store_true")
    ap.add_argument("--coef", action="store_true")
    ap.add_argument("--fused", action="store_true", help="fused forward kernel (sampler + decoder GEMM + residual in one launch)")
    ap.add_argument("--fused_debug", type=int, default=0)
    ap.add_argument("--wave", type=float, default=1.0, help="grid of the one-wave sampler schedule / resident wave")
    ap.add_argument("--pool", type=int, default=3, help="forward sampler schedule: 0 row by row, 1 per-row-tile pool, 2 one-wave pool, 3 one wave where bit-identical to 1 (default)")
    a = ap.parse_args()
    from poissonvae_b200 import _lib, linear
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    lib = _lib.load()
    if a.tf32:
        linear.PRECISE = False
    if a.no_overlap:
        lib.pvae_set_step_overlap(0)
    if a.no_carveout:
        lib.pvae_set_carveout(0)
    lib.pvae_set_pool(a.pool)
    lib.pvae_set_step_coef(1 if a.coef else 0)
    lib.pvae_set_step_fused(1 if a.fused else 0)
    lib.pvae_set_fused_debug(a.fused_debug)
    lib.pvae_set_wave_factor(a.wave)
    B, K = a.batch, a.latents
    model = PoissonVAE(ConfigPoisVAE(n_latents=K, prior_clamp=-4, seed=0)).cuda()
    model.n_exp.fill_(263)
    model._n_exp_host = 263
    cfg = ConfigTrainVAE(lr=0.005, epochs=3000, batch_size=B, warmup_portion=0.01, grad_clip=None, kl_anneal_portion=0.0,
                         kl_const_portion=0.0, temp_anneal_portion=0.0, temp_stop=a.temp)
    tr = TrainerVAE(model, cfg, n_batches_per_epoch=26)
    tr.use_planes = not a.no_planes
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(16, B, 256, generator=g)
    x = ((x - x.mean(2, keepdim=True)) / x.std(2, keepdim=True)).cuda()
    x_lo = tr.split_planes(x) if tr.planes else None
    for i in range(5):
        tr.train_step(x[i % 16], i, x_lo=None if x_lo is None else x_lo[i % 16])
    n = 10
    tot = [0.0] * (n - 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(a.steps):
        tr.train_step(x[i % 16], 5 + i, x_lo=None if x_lo is None else x_lo[i % 16])
    e1.record()
    torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) / a.steps * 1e3
    for i in range(a.steps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for e in evs:
            e.record()
        arr = (ctypes.c_void_p * n)(*[e.cuda_event for e in evs])
        lib.pvae_set_step_trace(arr, n)
        tr.train_step(x[i % 16], 5 + a.steps + i, x_lo=None if x_lo is None else x_lo[i % 16])
        lib.pvae_set_step_trace(None, 0)
        torch.cuda.synchronize()
        for j in range(n - 1):
            tot[j] += evs[j].elapsed_time(evs[j + 1]) * 1e3
    print(f"B={B} K={K} planes={tr.planes} tf32={a.tf32} pool={a.pool} carveout={not a.no_carveout} temp={a.temp}: step {plain:.1f} us untraced; traced per kernel (us):")
    for j, name in enumerate(NAMES):
        print(f"  {name:26s} {tot[j] / a.steps:7.2f}")
    print(f"  {'sum':26s} {sum(tot) / a.steps:7.2f}")


if __name__ == "__main__":
    main()
