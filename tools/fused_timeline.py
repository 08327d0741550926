"""Phase timeline of CTA 0 of the fused forward kernel (pvae_set_fused_timeline): when each sampler group finished the
prologue / the listed long chains / the pool / the batch, when the issuer acquired each k-block of z and committed its MMAs,
when the accumulators were complete.  python tools/fused_timeline.py [--batch 4096 --latents 512 --temp 1.0 --debug 0]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--latents", type=int, default=512)
    ap.add_argument("--temp", type=float, default=1.0)
    ap.add_argument("--debug", type=int, default=0)
    a = ap.parse_args()
    from poissonvae_b200 import _lib
    lib = _lib.load()
    B, K, D = a.batch, a.latents, 256
    g = torch.Generator("cuda").manual_seed(1)
    x = torch.randn(B, D, device="cuda", generator=g)
    w_enc = torch.randn(K, D, device="cuda", generator=g) * 0.05
    h = x @ w_enc.T
    log_r = torch.rand(K, device="cuda", generator=g) * 2 - 6
    w_dec = torch.randn(D, K, device="cuda", generator=g) * 0.05
    w_lo = torch.empty_like(w_dec)
    s = torch.cuda.current_stream().cuda_stream
    lib.pvae_tf32_lo(w_dec.data_ptr(), w_lo.data_ptr(), w_dec.numel(), s)
    outs = [torch.zeros(B, K, device="cuda") for _ in range(3)] + [torch.zeros(B, D, device="cuda") for _ in range(2)]
    recon, kl, rmax = torch.zeros(B, device="cuda"), torch.zeros(B, device="cuda"), torch.zeros(1, device="cuda")
    ts = torch.zeros(1024, dtype=torch.int64, device="cuda")
    lib.pvae_set_fused_debug(a.debug)

    def run():
        rc = lib.pvae_fused_fwd(h.data_ptr(), log_r.data_ptr(), None, w_dec.data_ptr(), w_lo.data_ptr(), x.data_ptr(), outs[0].data_ptr(),
                                outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), outs[4].data_ptr(), recon.data_ptr(),
                                kl.data_ptr(), rmax.data_ptr(), B, K, D, 0, 263, a.temp, 0, 0, 0.0, 2.0 / B, 5, 0, None, s)
        assert rc == 0, lib.pvae_last_error()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts[700:710:2] = 2 ** 62  # even words 700..708 collect minima
    lib.pvae_set_fused_timeline(ts.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    lib.pvae_set_fused_timeline(None)
    t = ts.cpu().tolist()
    t0 = t[0]
    us = lambda v: (v - t0) / 1e3 if v else float("nan")
    print(f"kernel {e0.elapsed_time(e1) * 1e3:.1f} us (events); stamps in us after CTA 0's sampler start")
    print(f"issuer start {us(t[1]):.1f}")
    nkb = K // 32
    print("k-block: z acquired / MMAs committed")
    print("  " + "  ".join(f"{kb}:{us(t[100 + kb]):.1f}/{us(t[200 + kb]):.1f}" for kb in range(nkb)))
    for bi in range(K // 256):
        print(f"batch {bi}: group: prologue / long chains / pool / end")
        for gi in range(7):
            b = 300 + (gi * 8 + bi) * 4
            print(f"  g{gi}: {us(t[b]):6.1f} {us(t[b + 1]):6.1f} {us(t[b + 2]):6.1f} {us(t[b + 3]):6.1f}")
    print(f"all CTAs: first start {us(t[700]):.1f}, last start {us(t[701]):.1f}, last batch end {us(t[703]):.1f}, "
          f"last accumulators complete {us(t[705]):.1f}, last epilogue done {us(t[707]):.1f}")
    ends = sorted(us(v) for v in t[800:948] if v)
    if ends:
        print("per CTA, end of the sampler part (us): min %.1f  p25 %.1f  median %.1f  p75 %.1f  p90 %.1f  max %.1f" % (
            ends[0], ends[len(ends) // 4], ends[len(ends) // 2], ends[3 * len(ends) // 4], ends[9 * len(ends) // 10], ends[-1]))
    print(f"accumulators complete {us(t[600]):.1f}, epilogue done {us(t[601]):.1f}")


if __name__ == "__main__":
    main()
