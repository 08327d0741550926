"""Micro-benchmark of the tcgen05 GEMM on the five contractions of one lin|lin step (B=4096, K=512, D=256)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poissonvae_b200 import _lib  # noqa: E402


def time_us(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
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
    lib = _lib.load()
    B, K, D = 4096, 512, 256
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator("cuda").manual_seed(0)
    r = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    shapes = {  # name: (A, B, M, N, Kd, a_k, b_k, splits)
        "enc  h=x W^T": (r(B, D), r(K, D), B, K, D, 1, 1, 1),
        "dec  y=z W^T": (r(B, K), r(D, K), B, D, K, 1, 1, 1),
        "dgrad g_z=g_y W": (r(B, D), r(D, K), B, K, D, 1, 0, 1),
        "wgrad_dec g_y^T z": (r(B, D), r(B, K), D, K, B, 0, 0, 16),
        "wgrad_enc g_h^T x": (r(B, K), r(B, D), K, D, B, 0, 0, 16),
    }
    quick = len(sys.argv) > 1 and sys.argv[1] == "--quick"
    if quick:
        shapes = {k: v for k, v in shapes.items() if k.startswith(("enc", "wgrad_dec"))}
    for variant in ((1,) if quick else (1, 2)):
        lib.pvae_set_gemm_variant(variant)
        for precise in (1, 0):
            tot = 0.0
            for name, (A, Bm, M, N, Kd, ak, bk, sp) in shapes.items():
                C = torch.empty(M, N, device="cuda")
                ws = torch.empty(max(1, lib.pvae_gemm_ws_floats(M, N, sp)), device="cuda")
                fn = lambda: lib.pvae_gemm_tf32(A.data_ptr(), Bm.data_ptr(), C.data_ptr(), M, N, Kd, ak, bk, precise, sp,
                                                ws.data_ptr(), s)
                assert fn() == 0, lib.pvae_last_error()
                t = time_us(fn, iters=2 if quick else 20, warm=1 if quick else 3)
                tot += t
                print(json.dumps(dict(variant=variant, precise=precise, gemm=name, us=round(t, 2),
                                      tflops=round(2 * M * N * Kd / t / 1e6, 1))), flush=True)
            print(json.dumps(dict(variant=variant, precise=precise, total_us=round(tot, 1))), flush=True)
    a, b = r(B, D), r(K, D)
    t = time_us(lambda: torch.matmul(a, b.t()))
    print(json.dumps(dict(gemm="torch fp32 enc", us=round(t, 2))))


if __name__ == "__main__":
    main()
