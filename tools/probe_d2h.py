import time, torch
ring = torch.zeros(32, 3, device="cuda")
host = torch.empty(1000, 3).pin_memory()
print("row view pinned:", host[5].is_pinned())
cs = torch.cuda.Stream()
ev = [torch.cuda.Event() for _ in range(32)]
main = torch.cuda.current_stream()
torch.cuda.synchronize()
for variant in ("copy_", "events+copy_", "events only"):
    t0 = time.perf_counter()
    for i in range(1000):
        k = i % 32
        if variant != "copy_":
            ev[k].record(main)
            cs.wait_event(ev[k])
        if variant != "events only":
            with torch.cuda.stream(cs):
                host[i].copy_(ring[k], non_blocking=True)
    dt = (time.perf_counter() - t0) / 1000
    torch.cuda.synchronize()
    print(f"{variant}: {dt*1e6:.1f} us per step (host side)")
