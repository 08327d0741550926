"""Pinned host -> device copy bandwidth by transfer size, alone and while kernels run on another stream."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def bw(nbytes, reps, busy=False):
    n = nbytes // 4
    h = torch.empty(n).pin_memory(); d = torch.empty(n, device="cuda")
    cs = torch.cuda.Stream()
    a = torch.randn(4096, 4096, device="cuda")
    with torch.cuda.stream(cs):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(cs):
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
    if busy:
        for _ in range(40):
            a = a @ a * 1e-4
    cs.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return nbytes * reps / dt / 1e9

for mib in (4, 16, 64, 256):
    print(f"{mib:4d} MiB: alone {bw(mib << 20, max(2, 256 // mib)):.1f} GB/s   with kernels running {bw(mib << 20, max(2, 256 // mib), True):.1f} GB/s", flush=True)
