"""Where the host time of one conv training step goes (cProfile over a few steps; GPU box)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from poissonvae_b200.trainer import ConfigTrainVAE, TrainerVAE  # noqa: E402
from poissonvae_b200.vae import ConfigPoisVAE, PoissonVAE, default_configs  # noqa: E402


def main():
    dataset, archi, batch = (sys.argv[1:] + ["MNIST", "conv+b|conv+b", "100"])[:3]
    cfg_vae, cfg_tr = default_configs(dataset, "poisson", archi)
    cfg_vae.update(n_latents=512)
    model = PoissonVAE(ConfigPoisVAE(seed=0, **cfg_vae)).cuda()
    sz = model.cfg.input_sz
    x = torch.rand(int(batch), 1, sz, sz, device="cuda")
    tr = TrainerVAE(model, ConfigTrainVAE(batch_size=int(batch), epochs=1000, temp_anneal_portion=0.0, temp_stop=1.0,
                                          kl_anneal_portion=0.0, kl_const_portion=0.0, lr=cfg_tr["lr"], grad_clip=cfg_tr["grad_clip"]),
                    n_batches_per_epoch=100)
    for i in range(5):
        tr.train_step(x, i)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(20):
        tr.train_step(x, 5 + i)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
