#!/usr/bin/env python
"""Training-step throughput of the conv P-VAEs (BASELINE config 4: conv+b|lin on CIFAR16-shaped 16x16 patches; config 5:
conv+b|conv+b on MNIST-shaped 28x28, K = 512) on one B200.  GPU side only, with A/B knobs and a per-kernel profile;
`python bench.py --workload config4|config5` runs the same measurement with the CPU restatement of the reference's
step timed beside it.

  python tools/bench_conv.py [--batch 1000] [--steps 20] [--warmup 5] [--dataset CIFAR16|MNIST] [--archi ...] [--tf32]

Prints one JSON line.  Not the headline benchmark (bench.py is): the secondary configurations of SURVEY 8d."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dataset", default="CIFAR16")
    ap.add_argument("--archi", default="conv+b|lin", help="conv+b|lin (BASELINE config 4) or conv+b|conv+b (config 5 with --dataset MNIST)")
    ap.add_argument("--latents", type=int, default=512)
    ap.add_argument("--n_exp", type=int, default=263)
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--no_cpu", action="store_true", help="(accepted for compatibility; the CPU leg lives in bench.py --workload)")
    ap.add_argument("--no_pdl", action="store_true", help="disable programmatic dependent launch (per-kernel times in a "
                    "profile are only meaningful without it: a dependent kernel's span includes its wait)")
    ap.add_argument("--no_overlap", action="store_true", help="keep a cell's weight-gradient kernels on the main stream (A/B)")
    ap.add_argument("--conv_variant", type=int, default=0)
    ap.add_argument("--profile", action="store_true", help="print the per-kernel time table of one step (torch profiler)")
    args = ap.parse_args()
    line = run(args)
    print(json.dumps(line), flush=True)


def run(args):
    """Times `args.steps` training steps; returns the result dict (also used by bench.py --workload)."""
    import torch
    from poissonvae_b200 import _lib, conv, linear
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    if args.tf32:
        conv.PRECISE = False
        linear.PRECISE = False
    cfg_vae, cfg_tr = default_configs(args.dataset, "poisson", args.archi)
    cfg_vae.update(n_latents=args.latents)
    model = PoissonVAE(ConfigPoisVAE(seed=0, **cfg_vae)).cuda()
    model.update_n(200.0)
    model._n_exp_host = args.n_exp
    sz = model.cfg.input_sz
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(args.batch, 1, sz, sz, generator=g)
    x = ((x - x.mean((1, 2, 3), keepdim=True)) / x.std((1, 2, 3), keepdim=True)).cuda()
    tr = TrainerVAE(model, ConfigTrainVAE(batch_size=args.batch, epochs=1000, temp_anneal_portion=0.0, temp_stop=1.0,
                                          kl_anneal_portion=0.0, kl_const_portion=0.0, **{k: v for k, v in cfg_tr.items()
                                                                                         if k in ("lr", "grad_clip")}),
                    n_batches_per_epoch=100)
    lib = _lib.load()
    if args.no_pdl:
        lib.pvae_set_pdl(0)
    if args.no_overlap:
        lib.pvae_set_cell_overlap(0)
    if args.conv_variant:
        lib.pvae_set_conv_variant(args.conv_variant)
    for i in range(args.warmup):
        tr.train_step(x, i)
    torch.cuda.synchronize()
    c0 = lib.pvae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        out = tr.train_step(x, args.warmup + i)
    e1.record()
    enqueue_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    host_s = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / args.steps
    line = {"metric": f"training samples/sec (P-VAE {args.archi} K={args.latents})", "value": round(args.batch / ms * 1e3, 1), "unit": "samples/s",
            "ms_per_step": round(ms, 3), "host_ms_per_step": round(host_s / args.steps * 1e3, 3),
            "enqueue_ms_per_step": round(enqueue_s / args.steps * 1e3, 3), "batch": args.batch,
            "dataset": args.dataset, "n_exp": args.n_exp, "gpu_launches_per_step": (lib.pvae_launch_count() - c0) // args.steps,
            "gemm": "tf32 single pass" if args.tf32 else "3xTF32", "loss": [round(v, 4) for v in out.tolist()],
            "params": tr.flat_p.numel()}
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            tr.train_step(x, args.warmup + args.steps)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
    return line


if __name__ == "__main__":
    main()
