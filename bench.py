#!/usr/bin/env python
"""Benchmark of the P-VAE training hot path (BASELINE.json metric: training samples/s, poisson
lin|lin K=512, batch 4096 per GPU, n_exp=263, fp32; synthetic whitened 16x16 patches, random-init
weights).  Prints ONE JSON line.

  python bench.py [--gpus N --steps K --warmup W]          our arm (torchrun for N > 1)
  python bench.py --impl reference [...]                   the reference's step on the host cores
                                                           (oracle/ref_port.py, pinned bit-for-bit to
                                                           the unmodified reference), bounded sample

A step = forward + loss + backward + Adamax update of one batch.  `value` is timed with inputs
resident in HBM; `e2e` feeds every batch from pinned host memory through the public
TrainerVAE.train_step_host call and reads the loss back every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "samples/s"


def metric_name(args):
    return f"training samples/sec (P-VAE lin|lin K={args.latents})"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="samples per GPU per step")
    ap.add_argument("--latents", type=int, default=512)
    ap.add_argument("--n_exp", type=int, default=263, help="fresh-model value, main/vae.py:382")
    ap.add_argument("--temp", type=float, default=1.0)
    ap.add_argument("--beta", type=float, default=1.0)
    ap.add_argument("--exit", default="lossless", choices=["lossless", "never"])
    ap.add_argument("--cpu_rows", type=int, default=200,
                    help="rows of the batch the CPU baseline steps on: 200 = BASELINE config 1's batch (the reference's own "
                         "CPU-runnable case), stepped in full")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_pdl", action="store_true", help="disable programmatic dependent launch (A/B)")
    ap.add_argument("--dp", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: gradient exchange over NVLink peer memory fused with Adamax, or NCCL all-reduce")
    ap.add_argument("--no_overlap", action="store_true", help="keep the whole step on one stream (A/B)")
    ap.add_argument("--tf32", action="store_true", help="single-pass TF32 GEMMs instead of 3xTF32 (A/B)")
    ap.add_argument("--dec_splits", type=int, default=-1, help="K-splits of the decoder GEMM (A/B; default: the library's choice)")
    ap.add_argument("--pool", type=int, default=3, help="forward sampler schedule (A/B): 0 row by row, 1 per-row-tile pool, 2 one-wave pool, 3 one wave where bit-identical to 1 (default)")
    ap.add_argument("--fused", action="store_true", help="the fused forward kernel (rsample + KL + decoder GEMM on z in shared memory "
                                                        "+ residual, one launch) instead of three kernels (A/B, measured slower; "
                                                        "see pvae_set_step_fused)")
    ap.add_argument("--coef", action="store_true", help="sampler coefficient mode: streaming backward (A/B, see pvae_set_step_coef)")
    ap.add_argument("--no_carveout", action="store_true", help="leave every kernel's L1 / shared-memory split at the driver default (A/B)")
    ap.add_argument("--no_planes", action="store_true", help="GEMMs split their operands per stage in shared memory instead of "
                                                              "loading the producers' TF32 low planes (A/B, same bits)")
    ap.add_argument("--no_fuse_update", action="store_true", help="keep the stand-alone gradient reductions + Adamax (A/B)")
    ap.add_argument("--graph", default="off", choices=["auto", "on", "off"],
                    help="CUDA-graph replay of the fused step (1 GPU; A/B - measured slower than the eager launches with "
                         "programmatic dependent launch, so off by default): auto = small batches and the static staging "
                         "buffers of the end-to-end path; on forces it")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5"],
                    help="config2 (default): the BASELINE metric, lin|lin K=512 beta=1, batch 4096 per GPU.  config3: BASELINE "
                         "configs[2], lin|lin K=1024 beta=2.5, 8192 samples per GPU (65536 on 8 GPUs).  config4 / config5: the "
                         "secondary conv configurations (conv+b|lin CIFAR16-shaped; conv+b|conv+b MNIST-shaped 28x28), 1 GPU, "
                         "with the CPU restatement of the reference's step beside them - not the driver's bench line")
    ap.add_argument("--min_region_ms", type=float, default=200.0,
                    help="the K-step timed region is repeated (from the same training state) until the regions add up to this "
                         "much device time; the MEDIAN region is reported")
    ap.add_argument("--no_ref_cuda", action="store_true", help="skip the reference-on-CUDA leg of the default line")
    ap.add_argument("--no_dp_check", action="store_true", help="N > 1: skip the correctness check of the gradient exchange")
    ap.add_argument("--ref_device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' times the same torch port of the reference's step with its "
                         "tensors on cuda:0 at the full batch (the reference's own CUDA path; north_star's denominator)")
    a = ap.parse_args()
    if a.workload == "config3":
        a.latents, a.beta, a.batch = 1024, 2.5, 8192
    return a


def workload(args, world=None):
    """`config` of the JSON line - the same dict on the product arm and on the reference arm."""
    world = args.gpus if world is None else world
    return {"workload": f"poisson lin|lin K={args.latents} beta={args.beta:g} batch {args.batch}/GPU n_exp={args.n_exp} "
                        f"temp={args.temp:g} fp32, fwd+loss+bwd+Adamax, exit={args.exit}",
            "batch_per_gpu": args.batch, "n_latents": args.latents, "n_exp": args.n_exp, "temp": args.temp, "beta": args.beta,
            "input": "16x16 z-scored N(0,1) patches, seed 1234; weights N(0,0.05^2), log_rate U(-6,-4), cfg.seed 0",
            "gemm": "tf32 single pass" if getattr(args, "tf32", False) else
                    "3xTF32 (fp32-grade, tcgen05), operand low planes written by the producers (x: once at upload)",
            "schedule": "reference vH16 lr warm-up (lr .005, 1 % of 78000 iterations), beta and temp pinned",
            "forward": "h -> ONE kernel (rsample + KL + decoder GEMM on z in shared memory + residual)" if getattr(args, "fused", False)
                       else "h -> three kernels (sampler, decoder GEMM, residual)",
            "graph": getattr(args, "graph", "off"),
            "l2": "inputs rotate over 64 distinct batches (>= 268 MB > 126 MB L2)",
            "parallelism": f"dp{world}"}


def synthetic_patches(n, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 256, generator=g)
    return (x - x.mean(1, keepdim=True)) / x.std(1, keepdim=True)


# ---------------------------------------------------------------------------------------------
def cpu_reference_steps(args, steps, warmup):
    """The reference's step (torch CPU port) on `cpu_rows` rows of the workload; returns samples/s."""
    import torch
    from oracle import ref_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = ref_port.make_params(K=args.latents, D=256, seed=0)
    optim = ref_port.Adamax(params, lr=0.005)
    x = synthetic_patches(args.cpu_rows, 1234).view(-1, 1, 16, 16)
    for _ in range(warmup):
        ref_port.train_step(params, optim, x, args.n_exp, args.temp, args.beta)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_port.train_step(params, optim, x, args.n_exp, args.temp, args.beta)
    dt = time.perf_counter() - t0
    return args.cpu_rows * steps / dt, cores, dt / steps


def cuda_reference_steps(args, steps, warmup):
    """The reference's step as it runs on its own CUDA path: the same torch op sequence (oracle/ref_port.py) with the
    tensors on cuda:0, full batch, materialised (n_exp, B, K) tensors, autograd backward, eager Adamax."""
    import torch
    from oracle import ref_port
    dev = torch.device("cuda", 0)
    params = {k: v.detach().to(dev).requires_grad_() for k, v in ref_port.make_params(K=args.latents, D=256, seed=0).items()}
    optim = ref_port.Adamax(params, lr=0.005)
    x = synthetic_patches(args.batch, 1234).view(-1, 1, 16, 16).to(dev)
    for _ in range(warmup):
        ref_port.train_step(params, optim, x, args.n_exp, args.temp, args.beta)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        ref_port.train_step(params, optim, x, args.n_exp, args.temp, args.beta)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    return args.batch * steps / dt, dt / steps, torch.cuda.max_memory_allocated() / 2**30


def run_reference(args):
    """Reference arm: the reference's own step on the host cores (its torch-CPU port, all threads): exactly K timed
    steps after W warm-up steps (the same K and W as the product arm).  Every step is a bounded sample of the workload:
    `cpu_rows` = 200 rows of the batch, i.e. BASELINE config 1 - the reference's own CPU-runnable case - stepped IN
    FULL (no scaling of a partial batch); the row count only shrinks when K + W steps of 200 rows would not fit a few
    minutes of CPU time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.ref_device == "cuda":
        v, sps, gib = cuda_reference_steps(args, steps, max(warmup, 1))
        line = {"impl": "reference", "device": "cuda", "metric": metric_name(args), "value": round(v, 1), "unit": UNIT, "n_gpus": 1,
                "steps": steps, "warmup": warmup, "ms_per_step": round(sps * 1e3, 3), "higher_is_better": True,
                "dtype": "f32", "data": "synthetic", "config": workload(args, 1), "peak_mem_gib": round(gib, 2),
                "note": "torch port of the reference step (bit-identical to it on CPU) run eagerly on cuda:0, full batch"}
        print(json.dumps(line), flush=True)
        return
    # ~3.3 ms per row and step on 16 cores at n_exp = 263, K = 512 (measured); keep the whole run under ~4 minutes
    budget_rows = int(240.0 / (0.0033 * (steps + warmup) * max(args.n_exp, 1) / 263.0 * args.latents / 512.0))
    args.cpu_rows = max(8, min(args.cpu_rows, budget_rows - budget_rows % 4))
    v, cores, sps = cpu_reference_steps(args, steps, warmup)
    sample = (f"{args.cpu_rows} rows per step (BASELINE config 1 batch = 200) of the {args.batch}-row workload batch, "
              f"{steps} steps after {warmup} warm-up, torch CPU port of the reference step (oracle/ref_port.py), {cores} threads")
    line = {"impl": "reference", "metric": metric_name(args), "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": round(sps * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(args),
            "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def conv_cpu_port(dataset, archi, latents, n_exp, x_rows):
    """The reference's conv step restated in plain torch on the host cores (oracle/conv_oracle.py, pinned to the
    unmodified reference by tests/test_conv_oracle.py): forward + loss + autograd backward.  Returns (samples/s, cores)."""
    import torch
    from oracle import conv_oracle
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs
    torch.set_num_threads(os.cpu_count() or 1)
    cfg_vae, _ = default_configs(dataset, "poisson", archi)
    cfg_vae.update(n_latents=latents)
    model = PoissonVAE(ConfigPoisVAE(seed=0, **cfg_vae))  # parameters only (CPU tensors); no kernel runs here
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in model.state_dict().items()}
    xs = x_rows.cpu()
    sc = lambda v, c: c - torch.nn.functional.softplus(c - v)

    def step():
        h = conv_oracle.conv_encoder(xs, sd, dataset)
        log_dr = sc(h, 10.0)
        rate = torch.exp(sc(sd["log_rate"] + log_dr, 5.0)) + 1.19e-7
        e = rate.new(torch.Size((n_exp,)) + rate.shape).exponential_()
        z = torch.sigmoid((1 - torch.cumsum(e / rate, 0)) / 1.0).sum(0)
        y = conv_oracle.conv_decoder(z, sd, dataset) if model.cfg.dec_type == "conv" else (z @ sd["fc_dec.weight"].T).view(xs.shape)
        kl = torch.exp(sd["log_rate"]) * (1 + torch.exp(log_dr) * (log_dr - 1))
        loss = torch.mean(((y - xs) ** 2).sum(dim=[1, 2, 3]) + kl.sum(1))
        for v in sd.values():
            v.grad = None
        loss.backward()

    step()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        step()
    return len(xs) * n / (time.perf_counter() - t0), os.cpu_count()


def run_conv_workload(args):
    """Secondary configurations 4 and 5: GPU timing by tools/bench_conv.run, CPU leg from the oracle."""
    import types

    import torch
    from tools import bench_conv
    cfg = {"config4": dict(dataset="CIFAR16", archi="conv+b|lin", batch=args.batch),
           "config5": dict(dataset="MNIST", archi="conv+b|conv+b", batch=args.batch if args.batch != 4096 else 100)}[args.workload]
    a = types.SimpleNamespace(batch=cfg["batch"], steps=args.steps, warmup=args.warmup, dataset=cfg["dataset"], archi=cfg["archi"],
                              latents=args.latents, n_exp=args.n_exp, tf32=args.tf32, no_pdl=args.no_pdl, no_overlap=False,
                              conv_variant=0, profile=False)
    line = bench_conv.run(a)
    line.update({"workload": args.workload, "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic"})
    if not args.no_cpu_baseline:
        sz = 28 if cfg["dataset"] == "MNIST" else 16
        rows = max(4, min(args.cpu_rows, 32))
        x = torch.randn(rows, 1, sz, sz, generator=torch.Generator().manual_seed(1234))
        v, cores = conv_cpu_port(cfg["dataset"], cfg["archi"], args.latents, args.n_exp, x)
        line["cpu_baseline"] = {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{rows} rows, 3 steps after 1 warm-up, fwd + loss + autograd bwd, torch CPU "
                                          "restatement (oracle/conv_oracle.py)"}
    print(json.dumps(line), flush=True)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        """Samples taken before this point (model build, warm-up) are dropped from the summary."""
        self.f.flush()
        self.skip = sum(1 for _ in open(self.f.name))

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = [r.split(",") for r in self.f.read().strip().splitlines()[getattr(self, "skip", 0):] if r.count(",") >= 7]
        self.f.close()
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for i, n in enumerate(names) if any("Active" == r[4 + i].strip() for r in rows)]
        out["samples"] = len(rows)
        return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from poissonvae_b200 import _lib
    from poissonvae_b200.distributions import EXIT_LOSSLESS, EXIT_NEVER
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path for the product arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()  # fails loudly if the CUDA extension has not been built
    if args.no_pdl:
        lib.pvae_set_pdl(0)
    if args.no_carveout:
        lib.pvae_set_carveout(0)
    lib.pvae_set_pool(args.pool)
    lib.pvae_set_step_coef(1 if args.coef else 0)
    lib.pvae_set_step_fused(1 if args.fused else 0)
    if args.no_overlap:
        lib.pvae_set_step_overlap(0)
    if args.tf32:
        from poissonvae_b200 import linear as _linear
        _linear.PRECISE = False
    dev = torch.device("cuda", local)
    clocks = ClockSampler(local) if rank == 0 else None  # nvidia-smi needs a moment to start sampling
    B, K, K_steps, W = args.batch, args.latents, args.steps, args.warmup

    model = PoissonVAE(ConfigPoisVAE(n_latents=K, prior_clamp=-4, seed=0),
                       exit_logit=EXIT_LOSSLESS if args.exit == "lossless" else EXIT_NEVER).to(dev)
    model.n_exp.fill_(args.n_exp)
    model._n_exp_host = args.n_exp
    # the reference's own vH16 schedule (main/config_vae.py:433-440: lr .005, 3000 epochs, 1 % linear lr warm-up,
    # no gradient clipping); 26 batches of 4096 per epoch for ~108 k training patches.  beta and the temperature
    # are pinned to the values the BASELINE config names.
    cfg = ConfigTrainVAE(lr=0.005, epochs=3000, batch_size=B, warmup_portion=0.01, grad_clip=None,
                         kl_beta=args.beta, kl_anneal_portion=0.0, kl_const_portion=0.0, temp_anneal_portion=0.0,
                         temp_stop=args.temp)
    tr = TrainerVAE(model, cfg, n_batches_per_epoch=26, dp_mode=args.dp)
    tr.graph = {"auto": "auto", "on": True, "off": False}[args.graph]
    tr.fuse_update = not args.no_fuse_update
    tr.use_planes = not args.no_planes
    if args.dec_splits >= 0:
        tr.lib.pvae_set_dec_splits(args.dec_splits)

    # inputs: 64 distinct batches (268 MB at B=4096 > 126 MB L2) so that no step finds its input in L2
    NB = 64 if B * 256 * 4 * 64 <= (1 << 30) else 32
    host = synthetic_patches(NB * B, 1234 + rank).view(NB, B, 256).pin_memory()
    data = host.to(dev)
    # the resident batches' TF32 low planes, split once at upload time (they are constant data): the encoder GEMM and the
    # encoder wgrad load them by TMA next to x.  The end-to-end pass splits every uploaded chunk inside its timed region.
    data_lo = tr.split_planes(data) if tr.planes else None
    xlo = (lambda j: data_lo[j]) if data_lo is not None else (lambda j: None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------
    for i in range(W):
        tr.train_step(data[i % NB], i, x_lo=xlo(i % NB))
    barrier()
    # every pass below (device-resident regions, instrumented, end-to-end) trains the same K steps from this state: the
    # sampler's work depends on the rates, so the passes must not see different stages of training
    snap = (tr.flat_p.clone(), tr.exp_avg.clone(), tr.exp_inf.clone(), tr.opt_step)

    def restore():
        tr.flat_p.copy_(snap[0]), tr.exp_avg.copy_(snap[1]), tr.exp_inf.copy_(snap[2])
        tr.opt_step = snap[3]
        torch.cuda.manual_seed(4321)
        barrier()

    if clocks is not None:
        clocks.mark()
    # A region = EXACTLY K steps between barrier + synchronize on both sides, timed with CUDA events on the launch
    # stream, max over ranks.  K steps of ~0.1 ms are a thin sample, so the region is repeated (each time from the same
    # training state) until the regions add up to min_region_ms of device time; the MEDIAN region is the reported one.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    regions, launches = [], 0
    while True:
        restore()
        launches0 = tr.lib.pvae_launch_count()
        e0.record()
        for i in range(K_steps):
            tr.train_step(data[(W + i) % NB], W + i, x_lo=xlo((W + i) % NB))
        e1.record()
        launches = tr.lib.pvae_launch_count() - launches0  # counted by the library itself, every launch of the region
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        regions.append(ms.item())  # identical on every rank: all ranks leave the loop together
        if sum(regions) >= args.min_region_ms or len(regions) >= 200:
            break
    ms_total = sorted(regions)[len(regions) // 2]
    # the same K steps again, now with CUDA events around the two sampler launches (roofline numerator).  Kept out
    # of the regions above because an event record between two kernels defeats their programmatic dependent launch.
    restore()
    tr.kernel_events = {}
    for i in range(K_steps):
        tr.train_step(data[(W + i) % NB], W + i, x_lo=xlo((W + i) % NB))
    barrier()
    ktimes = tr.kernel_times_us()
    tr.kernel_events = None

    # ---- end to end: pinned host batches in, loss out, every step --------------------------------
    # public API: TrainerVAE.fit_host(pinned (n_steps, B, D)) - chunked double-buffered uploads on a side stream,
    # every batch crosses PCIe inside the timed region and every step's loss is copied back to pinned memory
    tr.fit_host(host, 0, n_steps=max(W, 1))
    restore()
    barrier()
    t0 = time.perf_counter()
    loss_host = tr.fit_host(host, W, n_steps=K_steps, start=W % NB)
    t_enq = time.perf_counter() - t0  # host-side enqueue time (the GPU may still be working)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clk = clocks.stop() if clocks is not None else None
    final_loss = loss_host[K_steps - 1][0].item()

    # ---- N > 1: correctness of the gradient exchange that just ran, outside every timed region ------
    dp_check = None
    if world > 1 and not args.no_dp_check:
        dp_check = exchange_check(args, world, rank, dev, tr.peer is not None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # ALGORITHMIC bytes per latent, SURVEY.md 8(d): forward read h + write z = 8 B; backward read h, g_z + write g_h =
    # 12 B.  The implementation also moves dz_acc (written by the forward, read by the backward: +4 B each, L2-resident)
    # and the TF32 low planes of z / g_h for the GEMMs; that is reported as implementation traffic, not credited.
    alg = {"sampler_fwd": 8, "sampler_bwd": 12}
    impl_extra = {"sampler_fwd": 4 + (4 if tr.planes else 0), "sampler_bwd": 4 + (4 if tr.planes else 0)}
    # with the fused forward the "sampler_fwd" events bracket ONE kernel that also runs the decoder GEMM and the residual:
    # it is still credited the sampler's 8 B/latent only (its other algorithmic traffic - x, g_y, W_dec - is listed)
    fused_fwd = bool(tr.planes and args.fused and tr.lib.pvae_fused_fwd_supported(B, K, 256))
    dom = "sampler_fwd" if ktimes.get("sampler_fwd", 0.0) >= ktimes.get("sampler_bwd", 0.0) else "sampler_bwd"
    gbs = {k: B * K * alg[k] / (ktimes[k] * 1e-6) / 1e9 for k in alg if ktimes.get(k)}
    value = world * B * K_steps / (ms_total * 1e-3)
    step_us = ms_total / K_steps * 1e3
    line = {
        "metric": metric_name(args), "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K_steps, "warmup": W,
        "ms_per_step": round(ms_total / K_steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload(args, world),
        "final_loss": round(final_loss, 4),
        "gradient_exchange": ("none (1 GPU)" if world == 1 else
                              "NVLink peer-memory all-reduce fused with Adamax" if tr.peer is not None else "NCCL all-reduce"),
        "regions": {"n": len(regions), "steps_each": K_steps, "ms_median": round(ms_total, 4), "ms_min": round(min(regions), 4),
                    "ms_max": round(max(regions), 4), "note": "every region restarts from the same training state"},
        "clocks": clk,
        "e2e": {"value": round(world * B * K_steps / e2e_s.item(), 1), "unit": UNIT,
                "h2d_bytes_per_step": B * 256 * 4, "d2h_bytes_per_step": 12,
                "host_enqueue_us_per_step": round(t_enq / K_steps * 1e6, 1)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(gbs[dom], 1), "peak": peak, "unit": "GB/s",
                     "frac": round(gbs[dom] / peak, 4),
                     "traffic": committed_traffic(dom, B, K),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_latent": alg, "algorithmic_bytes_per_launch": B * K * alg[dom],
                     "implementation_extra_bytes_per_latent": impl_extra,
                     "sampler_fwd_is": ("fused_fwd_kernel: rsample + KL + decoder GEMM + residual (also moves x, g_y: "
                                        f"{8 * B * 256} B and W_dec: {4 * 256 * K} B per launch, not credited)") if fused_fwd
                                       else "sampler_pool_kernel (one-wave schedule, pool 2 / 3) or sampler_kernel (pool 0 / 1): the stand-alone forward sampler",
                     "kernel_us": {k: round(v, 2) for k, v in ktimes.items()},
                     "kernels": {k: {"achieved": round(v, 1), "frac": round(v / peak, 4)} for k, v in gbs.items()},
                     "sampler_fwd_plus_bwd": {"achieved": round(B * K * 20 / ((ktimes["sampler_fwd"] + ktimes["sampler_bwd"]) * 1e-6) / 1e9, 1),
                                              "frac": round(B * K * 20 / ((ktimes["sampler_fwd"] + ktimes["sampler_bwd"]) * 1e-6) / 1e9 / peak, 4)},
                     "whole_step": {"achieved": round(B * K * 20 / (step_us * 1e-6) / 1e9, 1),
                                    "frac": round(B * K * 20 / (step_us * 1e-6) / 1e9 / peak, 4),
                                    "note": "20 B/latent (sampler fwd + bwd) over the whole step's time"},
                     "timed": f"CUDA events on the launch stream around each sampler launch, mean over {K_steps} steps "
                              "run right after the timed regions"},
    }
    if ktimes.get("exchange"):
        # the exchange bucket that stays on the critical path (log_r, W_enc, b_enc): flag barrier + one-shot pull of every
        # peer's slice + Adamax, timed with events on the launch stream; NVLink bytes = what this rank reads from its peers
        nbytes = 4 * (K + K * 256) * (world - 1)
        line["exchange_us"] = round(ktimes["exchange"], 2)
        line["exchange"] = {"kernel": "allreduce_adamax_kernel (bucket 1 of 2; bucket 0 = W_dec runs on the side stream under the backward)",
                            "peer_bytes_per_rank": nbytes, "nvlink_gbs_per_rank": round(nbytes / (ktimes["exchange"] * 1e-6) / 1e9, 1)}
    if dp_check is not None:
        line["dp_check"] = dp_check
    ref_bytes = 4.0 * args.n_exp * B * K  # one materialised (n_exp, B, K) fp32 tensor of the reference path
    if world == 1 and not args.no_ref_cuda and ref_bytes * 16 > 150e9:
        line["ref_cuda"] = {"skipped": f"the reference path materialises ~16 tensors of {ref_bytes / 1e9:.1f} GB at this shape"}
    elif world == 1 and not args.no_ref_cuda:
        # north_star's denominator: the reference's own CUDA path on this same B200 (its op sequence - bit-identical
        # to the unmodified reference on CPU, tests/test_ref_port.py - run eagerly with the tensors on cuda:0, full batch)
        torch.cuda.empty_cache()
        ref_steps = 10
        rv, rsps, rgib = cuda_reference_steps(args, ref_steps, 3)
        line["ref_cuda"] = {"value": round(rv, 1), "unit": UNIT, "ms_per_step": round(rsps * 1e3, 3), "steps": ref_steps, "warmup": 3,
                            "peak_mem_gib": round(rgib, 2), "x_over_ref_cuda": round(value / rv, 1),
                            "what": "oracle/ref_port.py (the reference's torch op sequence: materialised (n_exp,B,K) tensors, "
                                    "autograd backward, eager Adamax) on cuda:0 at the same batch, CUDA-event timed"}
    if world == 1 and not args.no_cpu_baseline:
        steps_cpu, warm_cpu = 10, 3
        args.cpu_rows = min(args.cpu_rows, 200)
        v, cores, _ = cpu_reference_steps(args, steps_cpu, warm_cpu)
        line["cpu_baseline"] = {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"BASELINE config 1 exactly: batch {args.cpu_rows} stepped in full (no row sampling of "
                                          f"a larger batch), K={K}, n_exp={args.n_exp}, {steps_cpu} steps after {warm_cpu} warm-up, "
                                          f"torch CPU port of the reference step (oracle/ref_port.py), {cores} threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def committed_traffic(kernel, B, K):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the named kernel at this shape, from the ncu --set
    full capture committed under profiles/ (profiles/traffic.json names the capture and the commit it was taken at);
    null when no capture of this kernel + shape is committed.  Never a number of this run."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    rec = json.load(open(path)).get(f"{kernel}:B{B}:K{K}")
    return None if rec is None else rec["dram_bytes"]


def exchange_check(args, world, rank, dev, peer):
    """Data-parallel correctness, run once outside the timed regions: `world` ranks stepping 4 times (both gradient
    buffer parities, twice) on contiguous shards of a global batch, through the same exchange path the benchmark used,
    against ONE GPU stepping on the full global batch (every rank computes that reference itself).  Returns
    {"ranks_identical", "rel_vs_1gpu", ...}."""
    import torch
    import torch.distributed as dist
    from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
    from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
    K, Bl, steps = args.latents, 512, 4
    Bg = Bl * world
    data = torch.randn(steps, Bg, 256, generator=torch.Generator().manual_seed(99)).to(dev)

    def run(dp_mode):
        model = PoissonVAE(ConfigPoisVAE(n_latents=K, prior_clamp=-4, seed=0)).to(dev)
        with torch.no_grad():
            model.log_rate.add_(3.0)
        model.update_n(20.0)
        cfg = ConfigTrainVAE(lr=0.005, epochs=100, batch_size=Bl if dp_mode != "off" else Bg, warmup_portion=0.0,
                             scheduler_type=None, grad_clip=None, kl_beta=args.beta, kl_anneal_portion=0.0, kl_const_portion=0.0,
                             temp_anneal_portion=0.0, temp_stop=0.3)
        tr = TrainerVAE(model, cfg, dp_mode=dp_mode)
        losses = []
        for i in range(steps):
            x = data[i] if dp_mode == "off" else data[i, rank * Bl:(rank + 1) * Bl]
            if dp_mode != "off" and (i + rank) % 2 == 1:
                torch.cuda._sleep(2_000_000)  # ~1 ms of skew, alternating between the ranks
            losses.append(tr.train_step(x, i, philox=(77, 4 * i)).clone())
        torch.cuda.synchronize()
        return tr.flat_p.clone(), torch.stack(losses), tr.peer is not None

    p_dp, l_dp, used_peer = run("peer" if peer else "nccl")
    p_1, l_1, _ = run("off")
    lo, hi = p_dp.clone(), p_dp.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    rel = ((p_dp - p_1).abs().max() / p_1.abs().max()).reshape(1)
    lrel = ((l_dp - l_1).abs() / l_1.abs()).max().reshape(1)
    dist.all_reduce(rel, op=dist.ReduceOp.MAX)
    dist.all_reduce(lrel, op=dist.ReduceOp.MAX)
    return {"ranks_identical": bool(torch.equal(lo, hi)), "rel_vs_1gpu": float(rel.item()), "loss_rel_vs_1gpu": float(lrel.item()),
            "steps": steps, "global_batch": Bg, "path": "peer" if used_peer else "nccl",
            "ok": bool(torch.equal(lo, hi)) and rel.item() < 2e-5 and lrel.item() < 2e-5}


if __name__ == "__main__":
    a = parse()
    if a.workload in ("config4", "config5") and a.impl == "ours":
        run_conv_workload(a)
        sys.exit(0)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
