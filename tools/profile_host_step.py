"""cProfile of the host side of train_step / fit_host (where do the ~70 us per step go?)."""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE
from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE
m = PoissonVAE(ConfigPoisVAE(n_latents=512, prior_clamp=-4, seed=0)).cuda()
tr = TrainerVAE(m, ConfigTrainVAE(lr=.005, epochs=3000, batch_size=4096, grad_clip=None, temp_anneal_portion=0.0, temp_stop=1.0, kl_anneal_portion=0.0, kl_const_portion=0.0), n_batches_per_epoch=26)
x = torch.randn(4096, 256, device="cuda")
host = torch.randn(64, 4096, 256).pin_memory()
for i in range(20):
    tr.train_step(x, i)
tr.fit_host(host[:32], 20)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(300):
    tr.train_step(x, 100 + i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
t0 = time.perf_counter()
tr.fit_host(host, 1000)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"fit_host 64 steps: host enqueue {(t1-t0)/64*1e6:.1f} us/step, total {(t2-t0)/64*1e6:.1f} us/step")

# where does fit_host spend its wall time?  (a) GPU time of the training stream, (b) end-to-end wall
for n_steps, chunk in ((64, 16), (256, 16), (256, 32), (256, 4)):
    hb = host.repeat((n_steps + 63) // 64, 1, 1)[:n_steps].pin_memory()
    tr._chunk_bufs = None
    tr.fit_host(hb[:chunk * 2], 0, chunk=chunk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    tr.fit_host(hb, 2000, chunk=chunk)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"steps {n_steps} chunk {chunk}: enqueue {(t1-t0)/n_steps*1e6:.1f} us/step, wall {(t2-t0)/n_steps*1e6:.1f} us/step, "
          f"main-stream span {e0.elapsed_time(e1)/n_steps*1e3:.1f} us/step")
t0 = time.perf_counter()
for i in range(256):
    tr.train_step(x, 5000 + i)
torch.cuda.synchronize()
print(f"resident loop: wall {(time.perf_counter()-t0)/256*1e6:.1f} us/step")
