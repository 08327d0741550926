"""A/B of the host-fed path: chunk size and loss read-back variant, each from the same training state."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
m = PoissonVAE(ConfigPoisVAE(n_latents=512, prior_clamp=-4, seed=0)).cuda()
tr = TrainerVAE(m, ConfigTrainVAE(lr=.005, epochs=3000, batch_size=4096, grad_clip=None, temp_anneal_portion=0.0, temp_stop=1.0, kl_anneal_portion=0.0, kl_const_portion=0.0), n_batches_per_epoch=26)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
g = torch.Generator().manual_seed(0)
x = torch.randn(64, 4096, 256, generator=g)
x = ((x - x.mean(2, keepdim=True)) / x.std(2, keepdim=True))
host = x.repeat((K + 63) // 64, 1, 1)[:K].pin_memory()
dev = x.cuda()
for i in range(10):
    tr.train_step(dev[i], i)
torch.cuda.synchronize()
snap = (tr.flat_p.clone(), tr.exp_avg.clone(), tr.exp_inf.clone(), tr.opt_step)
def restore():
    tr.flat_p.copy_(snap[0]); tr.exp_avg.copy_(snap[1]); tr.exp_inf.copy_(snap[2]); tr.opt_step = snap[3]
    torch.cuda.manual_seed(1); torch.cuda.synchronize()
restore()
t0 = time.perf_counter()
for i in range(K):
    tr.train_step(dev[i % 64], 10 + i)
torch.cuda.synchronize()
print(f"resident: {(time.perf_counter()-t0)/K*1e6:.1f} us/step")
for chunk in (4, 8, 16):
    for direct in (True, False):
        tr._chunk_bufs = None
        tr.fit_host(host[:2 * chunk], 0, chunk=chunk, direct=direct)
        restore()
        t0 = time.perf_counter()
        tr.fit_host(host, 10, chunk=chunk, direct=direct)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"chunk {chunk:2d} direct {direct}: enqueue {(t1-t0)/K*1e6:.1f} wall {(t2-t0)/K*1e6:.1f} us/step")
