"""How fast do pinned host batches reach the device, and what does the host side of one e2e step cost?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = 4096 * 256
h = [torch.randn(n).pin_memory() for _ in range(8)]
d = torch.empty(n, device="cuda")
for _ in range(5):
    d.copy_(h[0], non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200):
    d.copy_(h[i % 8], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 200
print(f"H2D 4 MiB pinned: {dt*1e6:.1f} us  ({n*4/dt/1e9:.1f} GB/s)")
big = torch.randn(64 * n).pin_memory(); dbig = torch.empty(64 * n, device="cuda")
dbig.copy_(big, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); dbig.copy_(big, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D 256 MiB pinned: {big.numel()*4/dt/1e9:.1f} GB/s")
# host cost of a step: run train_step on a resident batch without waiting for the GPU
from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
m = PoissonVAE(ConfigPoisVAE(n_latents=512, prior_clamp=-4, seed=0)).cuda()
tr = TrainerVAE(m, ConfigTrainVAE(lr=.005, epochs=3000, batch_size=4096, grad_clip=None, temp_anneal_portion=0.0, temp_stop=1.0, kl_anneal_portion=0.0, kl_const_portion=0.0), n_batches_per_epoch=26)
x = torch.randn(4096, 256, device="cuda")
for i in range(10):
    tr.train_step(x, i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200):
    tr.train_step(x, 10 + i)
host = (time.perf_counter() - t0) / 200
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / 200
print(f"train_step: host-side enqueue {host*1e6:.1f} us per step, with GPU drain {tot*1e6:.1f} us per step")
