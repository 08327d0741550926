"""Micro-benchmark of the sampler kernels (CUDA events, L2 flushed between launches)."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poissonvae_b200 import _lib  # noqa: E402


def P(t):
    return None if t is None else t.data_ptr()


def time_it(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--K", type=int, default=512)
    ap.add_argument("--n_exp", type=int, default=263)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--single", action="store_true", help="only random_init, temp 1, lossless (for ncu captures)")
    ap.add_argument("--single_exit", default="lossless", choices=["lossless", "never"], help="exit policy of --single")
    ap.add_argument("--phase_a", type=int, default=0, help="override pvae_set_phase_a_events")
    ap.add_argument("--rows", type=int, default=0, help="override pvae_set_rows_per_block")
    args = ap.parse_args()
    lib = _lib.load()
    if args.phase_a > 0:
        assert lib.pvae_set_phase_a_events(args.phase_a) == 0
    if args.rows > 0:
        assert lib.pvae_set_rows_per_block(args.rows) == 0
    B, K, n_exp = args.B, args.K, args.n_exp
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator("cuda").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dists = {
        "random_init": (torch.randn(B, K, device="cuda", generator=g) * 0.82,
                        torch.rand(K, device="cuda", generator=g) * 2 - 6),
        "trained_like": (torch.randn(B, K, device="cuda", generator=g) * 2.0,
                         torch.full((K,), -3.0, device="cuda")),
    }
    z = torch.empty(B, K, device="cuda")
    gz = torch.randn(B, K, device="cuda", generator=g)
    gh = torch.empty(B, K, device="cuda")
    glr = torch.empty(K, device="cuda")
    ws = torch.empty(2 * lib.pvae_num_partials(B) * K, device="cuda")
    klrow = torch.empty(B, device="cuda")
    dz = torch.empty(B, K, device="cuda")
    temps = (1.0,) if args.single else (1.0, 0.05)
    exits = ((args.single_exit, 0.0 if args.single_exit == "lossless" else -1.0),) if args.single else (("lossless", 0.0), ("c30", 30.0), ("never", -1.0))
    if args.single:
        dists = {"random_init": dists["random_init"]}
    for name, (h, log_r) in dists.items():
        for temp in temps:
            for exit_name, ex in exits:
                def fwd():
                    rc = lib.pvae_sampler_fwd(P(h), P(log_r), None, P(z), None, None, None, P(klrow), None, None, B, K, 0,
                                              n_exp, temp, 0, 0, ex, 1, 4, None, s)
                    assert rc == 0

                def fwd_both():
                    rc = lib.pvae_sampler_fwd(P(h), P(log_r), None, P(z), None, None, None, P(klrow), None, P(dz), B, K, 0,
                                              n_exp, temp, 0, 0, ex, 1, 4, None, s)
                    assert rc == 0

                def bwd():
                    rc = lib.pvae_sampler_bwd(P(h), P(log_r), None, P(gz), None, None, 1.0 / B, P(gh), P(glr), None, P(ws), 0,
                                              B, K, 0, n_exp, temp, 0, 0, ex, 1, 4, None, s)
                    assert rc == 0

                def bwd_saved():
                    rc = lib.pvae_sampler_bwd(P(h), P(log_r), None, P(gz), None, P(dz), 1.0 / B, P(gh), P(glr), None, P(ws), 0,
                                              B, K, 0, n_exp, temp, 0, 0, ex, 1, 4, None, s)
                    assert rc == 0
                it = args.iters if ex >= 0 else 3
                tf, tb = time_it(fwd, it, 2, flush), time_it(bwd, it, 2, flush)
                tfb, tbs = time_it(fwd_both, it, 2, flush), time_it(bwd_saved, it, 2, flush)
                print(json.dumps(dict(dist=name, temp=temp, exit=exit_name, fwd_us=round(tf, 1), bwd_recompute_us=round(tb, 1),
                                      fwd_both_us=round(tfb, 1), bwd_saved_us=round(tbs, 1),
                                      fwd_both_gbs=round(B * K * 12 / tfb / 1e3, 1), bwd_saved_gbs=round(B * K * 16 / tbs / 1e3, 1),
                                      mean_z=round(z.mean().item(), 4))), flush=True)


if __name__ == "__main__":
    main()
