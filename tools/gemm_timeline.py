"""Pipeline timeline of the tcgen05 GEMM (pvae_set_gemm_timeline): for the enc / dgrad / wgrad shapes of config 2, when tile
(0,0,0) passed griddepcontrol.wait, requested and received each K block, committed its MMAs, saw the accumulator and
finished its epilogue; plus the earliest start and the latest finish over all CTAs.  warm L2, planes mode.
python tools/gemm_timeline.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poissonvae_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    B, K, D = 4096, 512, 256
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator("cuda").manual_seed(0)
    r = lambda *sh: torch.randn(*sh, device="cuda", generator=g)

    def lo(t):
        o = torch.empty_like(t)
        lib.pvae_tf32_lo(t.data_ptr(), o.data_ptr(), t.numel(), s)
        return o
    shapes = {  # name: (A, B, M, N, Kd, a_k, b_k, splits)
        "enc  h=x W^T": (r(B, D), r(K, D), B, K, D, 1, 1, 1),
        "dgrad g_z=g_y W": (r(B, D), r(D, K), B, K, D, 1, 0, 1),
        "wgrad_enc g_h^T x": (r(B, K), r(B, D), K, D, B, 0, 0, 16),
    }
    ts = torch.zeros(1024, dtype=torch.int64, device="cuda")
    for name, (A, Bm, M, N, Kd, ak, bk, sp) in shapes.items():
        Al, Bl = lo(A), lo(Bm)
        C = torch.empty(M, N, device="cuda")
        ws = torch.empty(lib.pvae_gemm_ws_floats(M, N, sp), device="cuda") if sp > 1 else None

        def run():
            rc = lib.pvae_gemm_tf32_planes(A.data_ptr(), Al.data_ptr(), Bm.data_ptr(), Bl.data_ptr(), C.data_ptr(), M, N, Kd, ak, bk, sp,
                                           None if ws is None else ws.data_ptr(), s)
            assert rc == 0, lib.pvae_last_error()
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts.zero_()
        ts[700::2] = 2 ** 62
        lib.pvae_set_gemm_timeline(ts.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        lib.pvae_set_gemm_timeline(None)
        t = ts.cpu().tolist()
        t0 = t[700]
        us = lambda v: (v - t0) / 1e3 if 0 < v < 2 ** 61 else float("nan")
        nkb = (Kd // max(sp, 1) + 31) // 32
        print(f"{name}: {e0.elapsed_time(e1) * 1e3:.1f} us by events (includes the split-K reduction if any); us after the first CTA's entry")
        print(f"  all CTAs: last past griddepcontrol.wait {us(t[701]):.2f}, last accumulator complete {us(t[703]):.2f}, last CTA done {us(t[705]):.2f}")
        print(f"  tile 0: entry {us(t[0]):.2f}, past wait {us(t[1]):.2f}, accumulator seen {us(t[70]):.2f}, epilogue done {us(t[71]):.2f}")
        print("  K block: requested / received / MMAs committed")
        print("   " + "  ".join(f"{kb}:{us(t[10 + kb]):.2f}/{us(t[30 + kb]):.2f}/{us(t[50 + kb]):.2f}" for kb in range(min(nkb, 16))))


if __name__ == "__main__":
    main()
