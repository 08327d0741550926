"""Host-side mirror of the reference training step (main/train_vae.py:317-459 `TrainerVAE.iteration`,
schedules main/train_vae.py:17-74, optimiser base/adamax.py, lr schedule base/train_base.py:258-307)
for the poisson lin|lin model, driving the sm_100a kernels through include/pvae_b200.h.

One step = forward + loss + backward + optimiser update on a device-resident batch; the gradient
is the closed form of SURVEY.md section 8 row a13 evaluated by the kernels (no autograd graph).
Data parallel: every rank runs the same step on its contiguous slice of the global batch, the flat
gradient buffer is summed with one NCCL all-reduce, and the Philox stream is keyed by the global
sample index, so the result does not depend on the number of GPUs.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _lib, dp, rng
from .distributions import INDICATOR_CODES, _ptr, _stream
from . import linear as linear_mod
from .rng import next_philox
from .vae import PoissonVAE


def temp_anneal_linear(n_iters, t0=1.0, t1=0.1, portion=0.7):
    """base/utils_model.py:6-14."""
    temps = np.ones(n_iters) * t1
    i = int(np.ceil(portion * n_iters))
    temps[:i] = np.linspace(t0, t1, i)
    return temps


def temp_anneal_exp(n_iters, t0=1.0, t1=0.1, portion=0.7):
    """base/utils_model.py:17-34 with rate='infer'."""
    n = min(int(np.ceil(n_iters * portion)), n_iters)
    rate = -(n / (n - 1)) * np.log(t1 / 100)
    temps = np.ones(n_iters) * t1
    i = np.arange(n)
    temps[:n] = t1 + (t0 - t1) * np.exp(-rate * i / n)
    return temps


def beta_anneal_linear(n_iters, beta=1.0, anneal_portion=0.3, constant_portion=0.0, min_beta=1e-4):
    """base/utils_model.py:57-68."""
    betas = np.ones(n_iters) * beta
    a = int(np.ceil(constant_portion * n_iters))
    b = int(np.ceil((constant_portion + anneal_portion) * n_iters))
    betas[:a] = min_beta
    betas[a:b] = np.linspace(min_beta, beta, b - a)
    return betas


def lr_schedule(gstep, n_iters, lr, warmup_portion=0.01, scheduler_type='cosine', eta_min=1e-5):
    """Learning rate of the optimiser step at `gstep`: linear warm-up (main/train_vae.py:329-337), then torch's
    CosineAnnealingLR with T_max = n_iters - n_iters * warmup_portion (base/train_base.py:282-287), stepped only
    outside warm-up (main/train_vae.py:391-400).  That scheduler is recursive on the group's current lr, so the
    cosine starts from the LAST warm-up value - which is also the lr of the first step after warm-up - not from
    cfg.lr.  Pinned to the reference optimiser + torch scheduler by tests/golden/lr_schedule.npz."""
    progress = gstep / n_iters
    if progress < warmup_portion:
        return lr * progress / warmup_portion
    warm = math.ceil(warmup_portion * n_iters)  # first gstep with progress >= warmup_portion
    while warm > 0 and (warm - 1) / n_iters >= warmup_portion:
        warm -= 1
    while warm / n_iters < warmup_portion:
        warm += 1
    lr_warm = lr * ((warm - 1) / n_iters) / warmup_portion if warm >= 1 else lr
    if scheduler_type is None:
        return lr_warm
    t_max = n_iters - n_iters * warmup_portion
    s = gstep - warm  # scheduler steps taken before this iteration
    return eta_min + (lr_warm - eta_min) * (1 + math.cos(math.pi * s / t_max)) / 2


class ConfigTrainVAE:
    """The fields of the reference ConfigTrainVAE / BaseConfigTrain the step reads
    (main/config_vae.py:267-330, base/config_base.py:66-110)."""

    def __init__(self, method='mc', kl_beta=1.0, kl_beta_min=1e-4, kl_anneal_cycles=0, kl_anneal_portion=0.5,
                 kl_const_portion=1e-2, temp_anneal_portion=0.5, temp_anneal_type='lin', temp_start=1.0,
                 temp_stop=0.05, lr=0.002, epochs=1200, batch_size=200, warmup_portion=0.01,
                 optimizer='adamax_fast', scheduler_type='cosine', grad_clip=500, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, eta_min=1e-5, log_freq=10, **unused):
        if method not in ('mc', 'exact'):
            raise ValueError(method)
        if kl_anneal_cycles != 0:
            raise NotImplementedError("cyclic beta annealing")
        if optimizer not in ('adamax_fast', 'adamax'):
            raise NotImplementedError(optimizer)
        if scheduler_type not in ('cosine', None):
            raise NotImplementedError(scheduler_type)
        assert temp_anneal_type in ('lin', 'exp')
        self.__dict__.update({k: v for k, v in locals().items() if k not in ('self', 'unused')})


class TrainerVAE:
    def __init__(self, model: PoissonVAE, cfg: ConfigTrainVAE, n_batches_per_epoch: int = 1,
                 process_group=None, dp_mode: str = 'auto'):
        self.model, self.cfg = model, cfg
        self.device = model.log_rate.device
        if self.device.type != 'cuda':
            raise RuntimeError("TrainerVAE needs the model on a CUDA device (no CPU path)")
        self.pg = process_group
        # dp_mode='off': a single-GPU trainer inside a distributed process (reference runs of the exchange checks)
        self.world, self.rank = (1, 0) if dp_mode == 'off' else dp.world_info(process_group)
        self.n_iters = cfg.epochs * n_batches_per_epoch
        self.betas = beta_anneal_linear(self.n_iters, cfg.kl_beta, cfg.kl_anneal_portion, cfg.kl_const_portion,
                                        cfg.kl_beta_min)  # main/train_vae.py:27-34
        if cfg.temp_anneal_portion > 0.0:  # main/train_vae.py:59-74
            fn = temp_anneal_linear if cfg.temp_anneal_type == 'lin' else temp_anneal_exp
            self.temperatures = fn(self.n_iters, cfg.temp_start, cfg.temp_stop, cfg.temp_anneal_portion)
        else:
            self.temperatures = np.ones(self.n_iters) * cfg.temp_stop
        self._flatten_parameters()
        self.opt_step = 0
        # per-gstep statistics with the reference's keys (main/train_vae.py:373-408) and, every cfg.log_freq steps, the
        # dictionary the reference hands to wandb.log (:410-445).  The step itself never syncs: loss / mse / kl, the gradient
        # norm and rate.max() of every step land in a small device ring that iteration() reads once per logging period.
        self.stats = {k: {} for k in ('grad', 'loss', 'lr', 'temp', 'r_max', 'n_exp')}
        self.log = []
        self._ring = None
        self._kl_diag = None       # (K) per-latent KL of the last logging step (device)
        self._want_kl_diag = False
        self._bufs = {}
        self.lib = _lib.load()
        # data-parallel exchange: 'peer' = all-reduce over NVLink peer memory fused with Adamax (one kernel),
        # 'nccl' = NCCL all-reduce followed by the Adamax kernel; 'auto' prefers 'peer' when it can be set up
        # (it has no gradient clipping: the norm of the summed gradient would need a second pass)
        self.peer = None
        self._peer_token = 0
        assert dp_mode in ('auto', 'peer', 'nccl', 'off')
        if self.world > 1 and dp_mode != 'nccl' and cfg.grad_clip is None and cfg.method == 'mc' and self.fused:
            err = None
            try:
                self.peer = dp.PeerExchange(self._flat_g_ext.numel(), self.device, process_group)
            except Exception as exc:  # no P2P / symmetric memory on this system
                err = exc
            # every rank must end up in the same mode (the peer kernels wait on one another): agree on the minimum
            ok = torch.tensor([0 if err is not None else 1], device=self.device, dtype=torch.int32)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=process_group)
            if int(ok.item()) == 0:
                self.peer = None
                if dp_mode == 'peer':
                    raise RuntimeError(f"peer-memory gradient exchange unavailable on at least one rank ({err!r})")
                import warnings
                warnings.warn(f"peer-memory gradient exchange unavailable on at least one rank ({err!r}); using NCCL all-reduce")
        elif dp_mode == 'peer' and self.world > 1:
            raise ValueError("dp_mode='peer' does not support grad_clip")
        # TF32 operand planes (csrc/gemm_tc.cu mode 2): every producer of a GEMM operand also writes its low part once, the
        # GEMMs load hi and lo tiles by TMA instead of splitting every stage in shared memory.  Bit-identical results.
        self.use_planes = True
        self.flat_p_lo = None
        self._planes_version = None
        self._descs = {}
        self.r_max = torch.zeros(max(n_batches_per_epoch, 1), device=self.device)  # rate.max() of each step of the epoch
        self._gather_tab = None
        self._sink = None
        # CUDA-graph replay of the fused lin|lin step (single GPU), see _graph_step.  OFF by default: measured slower than
        # the eager launches (73.7 vs 69.6 us per step at batch 200, 104 vs 78 us at batch 1024, end-to-end 31.6 M vs 34.0 M
        # samples/s at batch 4096) although the host enqueue time halves (91 -> 49 us): the step is bound by the GPU, and
        # between the kernel nodes of a replayed graph the programmatic-dependent-launch overlap of the eager stream
        # (every kernel's prologue under its predecessor's tail) is lost.  True forces it; 'auto' = small or static inputs.
        self.graph = False
        # single GPU without clipping: reduce the gradient partials inside the Adamax kernel (pvae_lin_step_train); False
        # keeps the stand-alone reduction kernels (bit-identical results, A/B)
        self.fuse_update = True
        self._graphs = {}
        self._graph_static = set()  # data_ptr() of input buffers that stay put (fit_host's staging slots)
        self._gstate = None
        self.kernel_events = None  # set to {} to record CUDA events around the named launches (bench.py)
        self._stream_handle = None  # fit_host pins the launch stream for the duration of its loop
        self._copy_stream = None

    # -- one flat fp32 buffer for parameters, gradients and Adamax state ------------------------
    def _flatten_parameters(self):
        m = self.model
        # the fused single-call step (csrc/lin_step.cu) covers lin|lin; every other architecture runs its blocks'
        # kernels under autograd and shares the flat-buffer exchange / clip / Adamax kernels
        self.fused = m.cfg.enc_type == 'lin' and m.cfg.dec_type == 'lin'
        if self.fused and m.fc_dec.bias is not None:
            raise NotImplementedError("decoder bias on the fused lin|lin step")
        if self.fused:
            names = ['log_rate', 'fc_enc.weight', 'fc_dec.weight']
            if m.fc_enc.bias is not None:
                names.append('fc_enc.bias')
        else:
            names = [k for k, p in m.named_parameters() if p.requires_grad]
        params = dict(m.named_parameters())
        n = sum(params[k].numel() for k in names)
        self.flat_p = torch.empty(n, device=self.device, dtype=torch.float32)
        # gradient buffer + 4 trailing floats for [loss, mse, kl, -]: data parallel sums both with ONE all-reduce
        self._flat_g_ext = torch.zeros(n + 4, device=self.device, dtype=torch.float32)
        self.flat_g = self._flat_g_ext[:n]

        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_inf = torch.zeros_like(self.flat_p)
        self._names, self._params = names, params
        self._plist = [params[k] for k in names]
        self.views, self.gviews, o = {}, {}, 0
        for k in names:
            p = params[k]
            self.flat_p[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + p.numel()].view_as(p)
            self.views[k] = p.data
            self.gviews[k] = self.flat_g[o:o + p.numel()].view_as(p)
            o += p.numel()

    def _buffers(self, B):
        if B not in self._bufs:
            K, D = self.model.cfg.n_latents, self.model.cfg.input_sz ** 2
            f = lambda *s: torch.empty(*s, device=self.device, dtype=torch.float32)
            self._bufs[B] = dict(ws=f(self.lib.pvae_lin_step_ws_floats(B, K, D)) if self.fused else None, loss=f(4),
                                 nws=f(self.lib.pvae_grad_norm_ws_floats()), norm=f(2))
        return self._bufs[B]

    # -- schedules, main/train_vae.py:326-345, 391-400 -------------------------------------------
    def lr_at(self, gstep: int) -> float:
        cfg = self.cfg
        return lr_schedule(gstep, self.n_iters, cfg.lr, cfg.warmup_portion, cfg.scheduler_type, cfg.eta_min)

    def _timed(self, name, fn, *args):
        """Launch `fn(*args)`; when kernel_events is a dict, bracket it with CUDA events on the launch stream."""
        if self.kernel_events is None:
            return fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*args)
        e1.record()
        self.kernel_events.setdefault(name, []).append((e0, e1))
        return out

    @property
    def planes(self) -> bool:
        """The fused step runs its GEMMs on operand planes (TF32 low parts written once by the operands' producers)."""
        return bool(self.use_planes and self.fused and linear_mod.PRECISE)

    def launches_per_step(self) -> int:
        """Kernels of libpvae_b200.so launched by one train_step (fused lin|lin step; for the other architectures
        difference pvae_launch_count() around a step)."""
        # 5 tcgen05 GEMMs (the decoder one with the residual fused in) + 2 split-K reductions (wgrad),
        # sampler fwd, loss, sampler bwd, column-sum finalize, Adamax; the fused update drops the three reductions
        if self.world == 1 and self.cfg.grad_clip is None and self.fuse_update:
            return 9
        n = 10  # one kernel reduces all gradient partials (pvae_set_fuse_reduce(0): three or four stand-alone reductions)
        if self.cfg.grad_clip is not None:
            n += 2  # sum of squares partials + finalize
        return n

    def kernel_times_us(self):
        """Mean duration per launch of the launches recorded in kernel_events (call after a synchronize)."""
        return {k: 1e3 * sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in (self.kernel_events or {}).items()}

    # -- host-fed step: the public entry point for batches that live in (pinned) host memory ----------
    def train_step_host(self, x_host: torch.Tensor, gstep: int = 0, x_next_host: torch.Tensor = None):
        """One step on a pinned host batch: the host->device copy of the NEXT batch is issued on a
        side stream before this step's kernels so that it overlaps them (double-buffered staging).
        Returns the device tensor [loss, mse, kl]; the caller reads it back (`.to('cpu')`)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [None, None]
            self._stage_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_src = [None, None]
            self._cur = 0
        cur = self._cur
        main = torch.cuda.current_stream()
        if self._stage_src[cur] is not x_host:  # not prefetched: copy now
            self._upload(cur, x_host, main)
        if x_next_host is not None:
            self._upload(1 - cur, x_next_host, main)
        main.wait_event(self._stage_ev[cur])
        out = self.train_step(self._stage[cur], gstep)
        self._stage_free[cur].record(main)
        self._stage_src[cur] = None
        self._cur = 1 - cur
        return out

    def fit_host(self, batches_host: torch.Tensor, first_gstep: int = 0, chunk: int = 8, losses_host: torch.Tensor = None,
                 direct: bool = None, n_steps: int = None, start: int = 0):
        """Train on `batches_host`, a pinned host tensor (n_steps, B, D): the public entry point for data that lives
        in host memory.  Batches are uploaded `chunk` at a time on a side stream into two alternating device
        buffers (one large DMA amortises the per-copy latency that dominates 4 MiB transfers) while the previous
        chunk trains; every step's [loss, mse, kl] is copied back to pinned memory asynchronously.
        `n_steps` / `start`: train n_steps steps cycling through the pool, step i on batch (start + i) % len(pool).
        Returns the pinned (n_steps, 3) tensor of results (valid after torch.cuda.synchronize())."""
        pool = batches_host.shape[0]
        n_steps = pool if n_steps is None else n_steps
        if losses_host is None:  # pinned allocations cost milliseconds: keep one and grow it when needed
            if getattr(self, "_losses_host", None) is None or self._losses_host.shape[0] < n_steps:
                self._losses_host = torch.empty(max(n_steps, 1024), 3).pin_memory()
            losses_host = self._losses_host[:n_steps]
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        shape = (chunk,) + tuple(batches_host.shape[1:])
        if getattr(self, "_chunk_bufs", None) is None or self._chunk_bufs[0].shape != shape:
            self._chunk_bufs = [torch.empty(shape, device=self.device, dtype=torch.float32) for _ in range(2)]
            # TF32 low planes of the staged batches: split once per upload on the copy stream, off the training stream
            self._chunk_lo = [torch.empty(shape, device=self.device, dtype=torch.float32) for _ in range(2)] if self.planes else None
            self._chunk_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._chunk_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._loss_ring = torch.empty(2 * chunk, 4, device=self.device, dtype=torch.float32)
            self._loss_ev = [torch.cuda.Event() for _ in range(2 * chunk)]
            for buf in self._chunk_bufs:  # static input buffers: their steps can be replayed as CUDA graphs
                for j in range(chunk):
                    self._graph_static.add(buf[j].data_ptr())
        use_lo = self.planes and getattr(self, "_chunk_lo", None) is not None
        lo_of = (lambda slot, j: self._chunk_lo[slot][j].reshape(shape[1], -1)) if use_lo else (lambda slot, j: None)
        if self.fused and self.world == 1 and self.graph is not False and self.cfg.method == 'mc' and not direct:
            for slot in range(2):
                for j in range(chunk):
                    self.precapture(self._chunk_bufs[slot][j].reshape(shape[1], -1), first_gstep, self._loss_ring[slot * chunk + j],
                                    x_lo=lo_of(slot, j))
        main = torch.cuda.current_stream()
        cs = self._copy_stream
        # every step's [loss, mse, kl] lands in a small device ring and is copied back on the side stream, so the
        # training stream never waits for PCIe.  direct=True (single process only) lets the loss kernel write its
        # 12 bytes straight into the pinned host row instead (UVA) - measured slightly slower, kept for comparison.
        if direct is None:
            direct = False
        if direct and self.world > 1:
            raise ValueError("direct host writes are not all-reduced: use direct=False for data-parallel runs")
        self._stream_handle = main.cuda_stream

        # chunk boundaries: 1, 2, 4, ... batches up to `chunk`, so that training starts after ~0.1 ms of DMA and
        # each upload is hidden behind the (doubling) compute of the chunk before it
        bounds, lo, size = [], 0, 1
        while lo < n_steps:
            first = (start + lo) % pool
            hi = min(lo + size, n_steps, lo + (pool - first))  # a chunk is one contiguous slice of the pool
            bounds.append((lo, hi))
            lo, size = hi, min(2 * size, chunk)

        def upload(c):  # chunk index -> staging buffer c % 2
            if c >= len(bounds):
                return
            lo, hi = bounds[c]
            slot = c % 2
            cs.wait_event(self._chunk_free[slot])
            with torch.cuda.stream(cs):
                first = (start + lo) % pool
                self._chunk_bufs[slot][:hi - lo].copy_(batches_host[first:first + hi - lo], non_blocking=True)
                if use_lo:
                    _lib.check(self.lib.pvae_tf32_lo(_ptr(self._chunk_bufs[slot]), _ptr(self._chunk_lo[slot]),
                                                     (hi - lo) * self._chunk_bufs[slot][0].numel(), cs.cuda_stream))
                self._chunk_ready[slot].record(cs)

        upload(0)
        for c, (lo, hi) in enumerate(bounds):
            upload(c + 1)  # overlaps this chunk's compute
            slot = c % 2
            main.wait_event(self._chunk_ready[slot])
            for i in range(lo, hi):
                if direct:
                    self.train_step(self._chunk_bufs[slot][i - lo], first_gstep + i, loss_out=losses_host[i], x_lo=lo_of(slot, i - lo))
                    continue
                k = slot * chunk + (i - lo)  # loss slot: reused two chunks later, after its read-back was enqueued
                self.train_step(self._chunk_bufs[slot][i - lo], first_gstep + i, loss_out=self._loss_ring[k], x_lo=lo_of(slot, i - lo))
                # read this step's result back on the side stream: the training stream never waits for PCIe
                self._loss_ev[k].record(main)
                cs.wait_event(self._loss_ev[k])
                with torch.cuda.stream(cs):
                    losses_host[i].copy_(self._loss_ring[k, :3], non_blocking=True)
            self._chunk_free[slot].record(main)
        self._stream_handle = None
        return losses_host

    def _upload(self, slot, x_host, main):
        if self._stage[slot] is None or self._stage[slot].shape != x_host.shape:
            self._stage[slot] = torch.empty(x_host.shape, device=self.device, dtype=torch.float32)
        cs = self._copy_stream
        cs.wait_event(self._stage_free[slot])  # the step that last read this slot has finished
        with torch.cuda.stream(cs):
            self._stage[slot].copy_(x_host, non_blocking=True)
            self._stage_ev[slot].record(cs)
        self._stage_src[slot] = x_host

    # -- the step ------------------------------------------------------------------------------------
    def train_step(self, x: torch.Tensor, gstep: int = 0, philox=None, loss_out: torch.Tensor = None, x_lo: torch.Tensor = None):
        """x: (B_local, 1, H, W) or (B_local, D) device tensor (this rank's slice of the global batch).
        x_lo: optional TF32 low plane of x (split_planes(x)); without it the step computes it itself.
        Returns the device tensor [loss, mse, kl] (global-batch means; no host sync); `loss_out` (3 floats on
        the device) receives it instead of the trainer's own scratch."""
        cfg, m, lib = self.cfg, self.model, self.lib
        x2 = x.reshape(x.shape[0], -1).contiguous()
        B, D = x2.shape
        K = m.cfg.n_latents
        Bg = B * self.world
        row_offset, _ = dp.shard_rows(Bg, self.world, self.rank)
        temp = float(self.temperatures[min(gstep, self.n_iters - 1)])
        beta = float(self.betas[min(gstep, self.n_iters - 1)])
        lr = self.lr_at(gstep)
        if temp != m._temp_host:
            m.update_t(temp)
        n_exp = m._n_exp_host
        ind = INDICATOR_CODES[m.cfg.indicator_approx]
        seed, offset = next_philox(self.device) if philox is None else philox
        b = self._buffers(B)
        n = self.flat_p.numel()
        if cfg.method == 'exact' or not self.fused:
            return self._train_step_autograd(x, beta, lr, gstep, b, loss_out, row_offset, (seed, offset))
        if philox is None and self._graph_wanted(x2, gstep):
            return self._graph_step(x2, gstep, temp, beta, n_exp, ind, seed, offset, loss_out, x_lo)
        # where this step's gradients and its [loss, mse, kl] shares go
        if self.peer is not None:  # symmetric double buffer, alternating with the exchange token's parity
            self._peer_token += 1  # strictly increasing for the life of the exchange (never rewound with opt_step)
            gbuf = self.peer.grads(self._peer_token & 1)
            self.flat_g, step_loss = gbuf[:n], gbuf[n:n + 4]
        elif self.world > 1:  # NCCL: the statistics ride behind the gradients in one all-reduce
            gbuf, step_loss = self._flat_g_ext, self._flat_g_ext[n:n + 4]
        else:
            gbuf, step_loss = self._flat_g_ext, (b['loss'] if loss_out is None else loss_out)
        loss = (b['loss'] if loss_out is None else loss_out) if self.world > 1 else step_loss
        s = self._stream_handle if self._stream_handle is not None else _stream()
        chk = _lib.check
        events = None
        if self.kernel_events is not None:  # CUDA events around the two sampler launches (bench.py roofline)
            n_ev = 6 if self.peer is not None else 4  # peer exchange: 4 / 5 bracket the bucket on the critical path
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
            for e in evs:
                e.record()  # materialises the cudaEvent_t; the library re-records it at the right place
            events = (ctypes.c_void_p * n_ev)(*[e.cuda_event for e in evs])
            self.kernel_events.setdefault('sampler_fwd', []).append((evs[0], evs[1]))
            self.kernel_events.setdefault('sampler_bwd', []).append((evs[2], evs[3]))
            if n_ev == 6:
                self.kernel_events.setdefault('exchange', []).append((evs[4], evs[5]))
        # one C call per step (csrc/lin_step.cu): forward + loss + backward and, single GPU without clipping or over
        # peer memory, the optimiser too.  The descriptor is kept from step to step; only what changes is rewritten.
        d = self._desc(B, Bg, row_offset, K, D)
        planes = self._planes_ready(s)
        d.x, d.x_lo = x2.data_ptr(), (_ptr(x_lo) if planes else None)
        d.params_lo = self.flat_p_lo.data_ptr() if planes else None
        d.grads, d.loss3 = gbuf.data_ptr(), step_loss.data_ptr()
        d.rate_max = self.r_max.data_ptr() + 4 * (gstep % self.r_max.numel())
        d.n_exp, d.temp, d.beta, d.indicator, d.precise = n_exp, temp, beta, ind, int(linear_mod.PRECISE)
        d.seed, d.offset, d.offset_dev = seed, offset, None
        d.events = ctypes.cast(events, ctypes.c_void_p) if events is not None else None
        d.lr, d.step = lr, self.opt_step + 1
        d.state = None
        if self.world == 1 and cfg.grad_clip is None and self.fuse_update:
            d.update = 1  # the gradient reductions happen inside the Adamax kernel, which also keeps params_lo current
            self.opt_step += 1
            chk(lib.pvae_lin_step(ctypes.byref(d), s))
            return loss[:3]
        if self.peer is not None:
            # exchange + Adamax inside the step, in two buckets (W_dec under the rest of the backward), NVLink peer memory
            tail = loss if loss.numel() >= 4 else b['loss']  # the kernel stores the summed statistics as one float4
            d.update, d.loss3 = 2, tail.data_ptr()
            d.world, d.rank, d.token = self.world, self.rank, self._peer_token
            d.peer_grads = ctypes.cast(self.peer.grad_ptrs[self._peer_token & 1], ctypes.c_void_p)
            d.peer_flags = ctypes.cast(self.peer.flag_ptrs, ctypes.c_void_p)
            self.opt_step += 1
            chk(lib.pvae_lin_step(ctypes.byref(d), s))
            if tail is not loss:
                loss.copy_(tail[:3])
            return loss[:3]
        d.update = 0
        chk(lib.pvae_lin_step(ctypes.byref(d), s))
        self.opt_step += 1
        if self.world > 1:
            dp.allreduce_gradients(self._flat_g_ext, None, self.pg)  # gradients already carry 1 / B_global
            loss[:3].copy_(step_loss[:3])
        # clip (main/train_vae.py:365-377) + Adamax (base/adamax.py)
        clip = None
        if cfg.grad_clip is not None:
            warming = gstep / self.n_iters < cfg.warmup_portion
            chk(lib.pvae_grad_norm(_ptr(self.flat_g), self.flat_g.numel(), cfg.grad_clip * (3 if warming else 1),
                                   _ptr(b['nws']), _ptr(b['norm']), s))
            clip = b['norm']
        chk(lib.pvae_adamax_step(_ptr(self.flat_p), _ptr(self.flat_g), _ptr(self.exp_avg), _ptr(self.exp_inf),
                                 self.flat_p.numel(), lr, self.opt_step, cfg.betas[0], cfg.betas[1], cfg.eps,
                                 cfg.weight_decay, _ptr(clip), s))
        self._planes_version = None  # this optimiser kernel does not write the low plane: the next step refreshes it
        return loss[:3]

    # -- TF32 operand planes ----------------------------------------------------------------------------
    def _param_version(self):
        """Changes whenever somebody wrote the parameters through torch (load_state_dict, in-place ops on a parameter or
        on the flat buffer).  Writes through a fresh `.data` alias are invisible to it: call params_changed() after those."""
        v = self.flat_p._version
        for p in self._plist:
            v += p._version
        return v

    def params_changed(self):
        """Tell the trainer that the parameters were edited behind its back: their TF32 low plane is rebuilt next step."""
        self._planes_version = None

    def _planes_ready(self, s) -> bool:
        """True when this step runs its GEMMs on operand planes; makes sure the parameters' low plane is current."""
        if not (self.use_planes and self.fused and linear_mod.PRECISE):
            return False
        if self.flat_p_lo is None:
            self.flat_p_lo = torch.empty_like(self.flat_p)
        v = self._param_version()
        if v != self._planes_version:
            _lib.check(self.lib.pvae_tf32_lo(_ptr(self.flat_p), _ptr(self.flat_p_lo), self.flat_p.numel(), s))
            self._planes_version = v
        return True

    def split_planes(self, x: torch.Tensor) -> torch.Tensor:
        """x - trunc_tf32(x) of a device tensor (numel a multiple of 4): the low plane train_step(x, x_lo=...) takes, so
        that data which stays in HBM is split once at upload time rather than once per step."""
        xc = x.contiguous()
        lo = torch.empty_like(xc)
        _lib.check(self.lib.pvae_tf32_lo(_ptr(xc), _ptr(lo), xc.numel(), _stream()))
        return lo

    def _desc(self, B, Bg, row_offset, K, D):
        key = (B, Bg, row_offset)
        d = self._descs.get(key)
        if d is None:
            cfg, m = self.cfg, self.model
            d = _lib.LinStepDesc()
            d.B, d.B_global, d.row_offset, d.K, d.D = B, Bg, row_offset, K, D
            d.has_bias, d.exc_only, d.exit_logit = int(m.fc_enc.bias is not None), int(m.cfg.exc_only), m.exit_logit
            d.params, d.ws = self.flat_p.data_ptr(), self._buffers(B)['ws'].data_ptr()
            d.exp_avg, d.exp_inf = self.exp_avg.data_ptr(), self.exp_inf.data_ptr()
            d.beta1, d.beta2, d.eps, d.weight_decay = cfg.betas[0], cfg.betas[1], cfg.eps, cfg.weight_decay
            self._descs[key] = d
        return d

    @torch.inference_mode()
    def validate(self, data: torch.Tensor, temp: float = 0.0, batch_size: int = None, full_data: bool = False):
        """Evaluation pass, main/train_vae.py:461-470, 530-601: hard counts (temp = 0, base/helper.py:15-17) ->
        decode -> r2 / mse / kl per sample.  Results are gathered on the device and cross to the host once, not
        eight times per batch.  Returns (data, loss, etc) dicts of numpy arrays with the reference's keys."""
        m = self.model
        was_training = m.training
        m.eval()  # select_model() returns model.eval(), base/train_base.py:197-198: no dropout in the conv encoder's dense layers
        try:
            return self._validate(data, temp, batch_size, full_data)
        finally:
            m.train(was_training)

    def _validate(self, data, temp, batch_size, full_data):
        m = self.model
        bs = batch_size or self.cfg.batch_size
        acc = {k: [] for k in ('x', 'y', 'z', 'r2', 'mse', 'kl', 'kl_diag', 'log_dr', 'rate')}
        eps = torch.finfo(torch.float32).eps
        for lo in range(0, data.shape[0], bs):
            x = data[lo:lo + bs].to(self.device)
            dist_, log_dr, z, y = m.xtract_ftr(x, temp)
            kl = m.loss_kl(log_dr)
            true, pred = x.flatten(start_dim=1), y.flatten(start_dim=1)
            ss_res = (true - pred).pow(2).sum(-1)  # compute_r2, base/utils_model.py:71-80
            ss_tot = (true - true.mean(dim=-1, keepdim=True)).pow(2).sum(-1)
            acc['r2'].append(1.0 - ss_res / (ss_tot + eps))
            acc['mse'].append(m.loss_recon(y, x))
            acc['kl'].append(kl.sum(1))
            acc['kl_diag'].append(kl.mean(0, keepdim=True))
            acc['z'].append(z)
            acc['log_dr'].append(log_dr)
            acc['rate'].append(dist_.rate)
            if full_data:
                acc['x'].append(x)
                acc['y'].append(y)
        out = {k: torch.cat(v).cpu().numpy() if v else None for k, v in acc.items()}
        data_out = {'x': out['x'], 'y': out['y'], 'z': out['z'], 'g': None}
        loss = {'r2': out['r2'], 'mse': out['mse'], 'kl': out['kl'], 'elbo': out['mse'] + out['kl'],
                'kl_diag': out['kl_diag'].mean(0)}
        etc = {'log_dr': out['log_dr'], 'r*dr': out['rate']}
        return data_out, loss, etc

    # -- CUDA-graph replay of the fused step ---------------------------------------------------------
    def _graph_wanted(self, x2, gstep):
        if self.graph is False or self.world > 1 or self.kernel_events is not None or gstep != self.opt_step:
            return False
        nxt = min(gstep + 1, self.n_iters - 1)
        if self.temperatures[min(gstep, self.n_iters - 1)] != self.temperatures[nxt] or self.betas[min(gstep, self.n_iters - 1)] != self.betas[nxt]:
            return False  # annealing: this step's scalars are used once - a graph would be captured for nothing
        if self.graph is True:
            return True
        # 'auto': where the 12-14 launches, not the kernels, set the pace (small batches), or where the caller promised
        # static input buffers (fit_host registers its staging slots)
        return x2.data_ptr() in self._graph_static or x2.numel() <= (1 << 18)

    def _graph_state(self):
        if self._gstate is None:
            self._gstate = torch.zeros(4, device=self.device, dtype=torch.int32)
            self._gstate_host = [-1, -1]  # what the device words hold (step count, Philox offset)
            self._lr_sched = torch.tensor([self.lr_at(i) for i in range(self.n_iters)], device=self.device, dtype=torch.float32)
            self._gstate_pin = torch.zeros(2, dtype=torch.int64).pin_memory()  # [step count | Philox offset] = the 4 device ints
            self._rmax_step = torch.zeros(1, device=self.device)  # the captured sampler's rate.max() word (moved per replay)

    def _graph_key(self, x2, x_lo, temp, beta, n_exp, seed, warming, loss_out):
        static_in = x2.data_ptr() in self._graph_static
        return (x2.data_ptr() if static_in else 0, _ptr(x_lo) if (static_in and x_lo is not None) else 0, x2.shape[0], temp, beta,
                n_exp, seed, warming, 0 if loss_out is None else loss_out.data_ptr(), bool(self.planes)), static_in

    def _graph_step(self, x2, gstep, temp, beta, n_exp, ind, seed, offset, loss_out, x_lo=None):
        """One training step as ONE cudaGraphLaunch.  The graph holds the kernels of the fused step (forward, loss,
        backward, [clip,] Adamax) with the step-dependent scalars on the device: the learning-rate schedule is a device
        array indexed by a device step counter, the Philox offset is a device word read by the sampler kernels
        (`offset_dev`); a one-thread kernel at the end of the graph advances both and files the step's rate.max() into
        its slot of the epoch ring.  A graph is specific to (input buffer, batch, temp, beta, n_exp, clip threshold):
        while the temperature anneals every step differs and the eager path runs instead; inputs that are neither
        registered static buffers nor small are stepped eagerly too."""
        cfg = self.cfg
        B, D = x2.shape
        self._graph_state()
        warming = cfg.grad_clip is not None and gstep / self.n_iters < cfg.warmup_portion
        key, static_in = self._graph_key(x2, x_lo, temp, beta, n_exp, seed, warming, loss_out)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 256:
                self._graphs.clear()
            g = self._graphs[key] = self._capture(x2 if static_in else None, x_lo if static_in else None, B, D, temp, beta, n_exp, ind,
                                                  seed, warming, loss_out)
        # device counters: rewritten only when they are not what the previous replay left behind (first use, restore,
        # a generator that was re-seeded or advanced by somebody else)
        if self._gstate_host != [self.opt_step, offset]:
            self._gstate_pin[0] = self.opt_step
            self._gstate_pin[1] = offset
            self._gstate.view(torch.int64).copy_(self._gstate_pin, non_blocking=True)
        if self.planes:  # parameters edited through torch since the last step: rebuild their low plane before the replay
            self._planes_ready(_stream())
        if not static_in:
            g['x'].copy_(x2, non_blocking=True)
        g['graph'].replay()
        self.opt_step += 1
        self._gstate_host = [self.opt_step, offset + 4]
        if g['clip']:
            self._planes_version = None
        return g['loss'][:3]

    def precapture(self, x2, gstep, loss_out=None, x_lo=None):
        """Build the graph a later train_step(x2, gstep, loss_out=loss_out, x_lo=x_lo) would replay, without running it
        (captures cost about a millisecond each: callers with static buffers pay that before their loop, not inside it)."""
        if not (self.fused and self.cfg.method == 'mc' and self.world == 1 and self.graph is not False):
            return
        m = self.model
        temp = float(self.temperatures[min(gstep, self.n_iters - 1)])
        beta = float(self.betas[min(gstep, self.n_iters - 1)])
        seed = torch.cuda.default_generators[self.device.index].initial_seed()
        self._graph_state()
        warming = self.cfg.grad_clip is not None and gstep / self.n_iters < self.cfg.warmup_portion
        key, static_in = self._graph_key(x2, x_lo, temp, beta, m._n_exp_host, seed, warming, loss_out)
        if key not in self._graphs:
            self._graphs[key] = self._capture(x2 if static_in else None, x_lo if static_in else None, x2.shape[0], x2.shape[1], temp,
                                              beta, m._n_exp_host, INDICATOR_CODES[m.cfg.indicator_approx], seed, warming, loss_out)

    def _capture(self, x_static, x_lo_static, B, D, temp, beta, n_exp, ind, seed, warming, loss_out):
        cfg, m, lib = self.cfg, self.model, self.lib
        K, n = m.cfg.n_latents, self.flat_p.numel()
        b = self._buffers(B)
        x = x_static if x_static is not None else torch.empty(B, D, device=self.device, dtype=torch.float32)
        loss = loss_out if loss_out is not None else torch.zeros(4, device=self.device, dtype=torch.float32)
        planes = self.planes
        if planes and self.flat_p_lo is None:
            self.flat_p_lo = torch.empty_like(self.flat_p)
        d = _lib.LinStepDesc()  # a private copy: the capture reads it once
        ctypes.memmove(ctypes.byref(d), ctypes.byref(self._desc(B, B, 0, K, D)), ctypes.sizeof(d))
        d.x, d.x_lo = x.data_ptr(), (_ptr(x_lo_static) if planes else None)
        d.params_lo = self.flat_p_lo.data_ptr() if planes else None
        d.grads, d.loss3, d.rate_max = self._flat_g_ext.data_ptr(), loss.data_ptr(), self._rmax_step.data_ptr()
        d.n_exp, d.temp, d.beta, d.indicator, d.precise = n_exp, temp, beta, ind, int(linear_mod.PRECISE)
        d.seed, d.offset, d.offset_dev, d.events = seed, 0, self._gstate.data_ptr() + 8, None
        d.lr_sched, d.n_sched, d.state = self._lr_sched.data_ptr(), self._lr_sched.numel(), self._gstate.data_ptr()
        d.philox_increment, d.rmax_ring, d.ring_n = 4, self.r_max.data_ptr(), self.r_max.numel()
        fused_update = cfg.grad_clip is None and self.fuse_update
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        graph = torch.cuda.CUDAGraph()
        pinned, self._stream_handle = self._stream_handle, None
        try:
            with torch.cuda.graph(graph, stream=side):
                s = torch.cuda.current_stream().cuda_stream
                if fused_update:  # the whole step incl. Adamax (which keeps params_lo current) and the counter advance
                    d.update, d.step = 1, 0
                    _lib.check(lib.pvae_lin_step(ctypes.byref(d), s))
                else:
                    if planes:  # the optimiser below does not write the low plane: rebuild it at the head of every replay
                        _lib.check(lib.pvae_tf32_lo(_ptr(self.flat_p), _ptr(self.flat_p_lo), n, s))
                    d.update, d.state = 0, None
                    _lib.check(lib.pvae_lin_step(ctypes.byref(d), s))
                    clip = None
                    if cfg.grad_clip is not None:
                        _lib.check(lib.pvae_grad_norm(_ptr(self._flat_g_ext), n, cfg.grad_clip * (3 if warming else 1), _ptr(b['nws']),
                                                      _ptr(b['norm']), s))
                        clip = b['norm']
                    _lib.check(lib.pvae_adamax_step_dev(_ptr(self.flat_p), _ptr(self._flat_g_ext), _ptr(self.exp_avg), _ptr(self.exp_inf),
                                                        n, _ptr(self._lr_sched), self._lr_sched.numel(), _ptr(self._gstate), cfg.betas[0],
                                                        cfg.betas[1], cfg.eps, cfg.weight_decay, _ptr(clip), 4, _ptr(self._rmax_step),
                                                        _ptr(self.r_max), self.r_max.numel(), s))
        finally:
            self._stream_handle = pinned
        cur.wait_stream(side)
        return dict(graph=graph, x=x, loss=loss, clip=not fused_update)

    def _train_step_autograd(self, x, beta, lr, gstep, b, loss_out, row_offset, philox):
        """The step for everything but fused lin|lin: method='exact' (main/train_vae.py:703-705, closed-form
        reconstruction loss, no sampling) and the conv architectures.  The model's own blocks (our kernels under
        autograd) build the loss; one gather kernel packs the per-parameter gradients into the flat buffer and the
        usual exchange + clip + Adamax kernels follow."""
        with rng.pinned_launch_stream():
            return self._autograd_step(x, beta, lr, gstep, b, loss_out, row_offset, philox)

    def _autograd_step(self, x, beta, lr, gstep, b, loss_out, row_offset, philox):
        cfg, m, lib, s = self.cfg, self.model, self.lib, _stream()
        names = self._names
        params = self._params
        for k in names:
            params[k].grad = None
        xin = x.reshape(x.shape[0], 1, m.cfg.input_sz, m.cfg.input_sz)
        m.row_offset = row_offset  # data parallel: the Philox stream is keyed by the global sample index
        m.rate_max_out = self.r_max[gstep % self.r_max.numel():][:1]
        m.philox_override = philox
        try:
            if cfg.method == 'exact':
                recon, dist_, log_dr = m.loss_recon_exact(xin)
            else:
                dist_, log_dr, z, y = m(xin)
                recon = m.loss_recon(y, xin)
            kl = m.loss_kl(log_dr)
            if self._want_kl_diag:
                self._kl_diag = kl.detach().mean(0).reshape(-1)  # main/train_vae.py:348
        finally:
            m.row_offset, m.rate_max_out, m.philox_override = 0, None, None
        Bg = x.shape[0] * self.world
        from . import conv
        if self._sink is None:  # the conv cells write their parameter gradients straight into the flat buffer
            self._sink = {self.views[k].data_ptr(): self.gviews[k] for k in names}
        conv.GRAD_SINK = self._sink
        try:
            (torch.sum(recon + beta * torch.sum(kl, dim=1)) / Bg).backward()
        finally:
            conv.GRAD_SINK = None
        n = self.flat_p.numel()
        if self._gather_tab is None:
            offs, o = [], 0
            for k in names:
                offs.append(o)
                o += params[k].numel()
            offs.append(o)
            self._gather_tab = (ctypes.c_int64 * len(offs))(*offs)
        # parameters whose gradient already sits in the flat buffer (written in place by a cell) copy onto themselves
        srcs = (ctypes.c_void_p * len(names))(*[self.gviews[k].data_ptr() if params[k].grad is None else params[k].grad.data_ptr()
                                                for k in names])
        _lib.check(lib.pvae_gather(srcs, self._gather_tab, len(names), _ptr(self._flat_g_ext), s))
        stats = self._flat_g_ext[n:n + 4]
        stats[0] = (recon.detach().sum() + beta * kl.detach().sum()) / Bg
        stats[1] = recon.detach().sum() / Bg
        stats[2] = kl.detach().sum() / Bg
        if self.world > 1:
            dp.allreduce_gradients(self._flat_g_ext, None, self.pg)
        loss = b['loss'] if loss_out is None else loss_out
        loss[:3].copy_(stats[:3])
        clip = None
        if cfg.grad_clip is not None:
            warming = gstep / self.n_iters < cfg.warmup_portion
            _lib.check(lib.pvae_grad_norm(_ptr(self._flat_g_ext), n, cfg.grad_clip * (3 if warming else 1), _ptr(b['nws']),
                                          _ptr(b['norm']), s))
            clip = b['norm']
        self.opt_step += 1
        _lib.check(lib.pvae_adamax_step(_ptr(self.flat_p), _ptr(self._flat_g_ext), _ptr(self.exp_avg), _ptr(self.exp_inf), n,
                                        lr, self.opt_step, cfg.betas[0], cfg.betas[1], cfg.eps, cfg.weight_decay, _ptr(clip), s))
        for k in names:  # the gradient tensors may be recycled now
            params[k].grad = None
        return loss[:3]

    def _kl_diag_of_last_step(self, B):
        """kl_diag = mean_b kl (K) of the step that just ran (main/train_vae.py:348), on the device.  Fused path: recomputed
        from the encoder output the step left in its workspace (pvae_kl_diag); autograd path: kept by the step."""
        m, lib = self.model, self.lib
        K, D = m.cfg.n_latents, m.cfg.input_sz ** 2
        if not (self.fused and self.cfg.method == 'mc'):
            return self._kl_diag
        b = self._buffers(B)
        if 'kld_ws' not in b:
            b['kld_ws'] = torch.empty(lib.pvae_kl_diag_ws_floats(B, K), device=self.device)
            b['kld'] = torch.empty(K, device=self.device)
            offs = (ctypes.c_int64 * 7)()
            _lib.check(lib.pvae_lin_step_ws_offsets(B, K, D, offs))
            b['h_off'] = int(offs[0])
        h = b['ws'][b['h_off']:]
        bias = m.fc_enc.bias
        _lib.check(lib.pvae_kl_diag(_ptr(h), _ptr(self.views['log_rate']), None if bias is None else _ptr(self.views['fc_enc.bias']),
                                    _ptr(b['kld']), _ptr(b['kld_ws']), B, K, int(m.cfg.exc_only), _stream()))
        return b['kld']

    def _flush_stats(self, pend, ring, meters, write, kl_diag, beta):
        """One device->host read for the steps in `pend` (ring rows [loss, mse, kl, -, grad_norm]); fills self.stats like
        main/train_vae.py:373-408 and, when `write`, appends the reference's wandb dictionary (:410-445) to self.log."""
        cfg, m = self.cfg, self.model
        n, slots, K = len(pend), self.r_max.numel(), m.cfg.n_latents
        parts = [ring[:n].reshape(-1), self.r_max]
        if write and kl_diag is not None:
            if self.world > 1:
                kl_diag = kl_diag.clone()
                torch.distributed.all_reduce(kl_diag, group=self.pg)
                kl_diag /= self.world
            parts.append(kl_diag)
        host = torch.cat(parts).cpu()  # the only sync
        rows, rmax = host[:n * 8].view(n, 8), host[n * 8:n * 8 + slots]
        for j, g in enumerate(pend):
            loss, mse, kl, gn = (float(rows[j, i]) for i in (0, 1, 2, 4))
            if cfg.grad_clip is not None:  # :365-377
                meters['grads'].update(gn)
                self.stats['grad'][g] = gn
                if gn > cfg.grad_clip:
                    self.stats['loss'][g] = loss
            meters['nelbo'].update(mse + kl)  # :379-388
            meters['perdim_mse'].update(mse / m.cfg.input_sz ** 2)
            meters['perdim_kl'].update(kl / K)
            meters['r_max'].update(float(rmax[g % slots]))
            warming = g / self.n_iters < cfg.warmup_portion
            self.stats['lr'][g] = self.lr_at(g) if warming else self.lr_at(g + 1)  # read after the scheduler stepped, :394-402
            self.stats['temp'][g] = float(self.temperatures[min(g, self.n_iters - 1)])
            self.stats['r_max'][g] = meters['r_max'].avg
            self.stats['n_exp'][g] = m._n_exp_host
            if write and j == n - 1:
                to_write = {'coeffs/beta': beta, 'coeffs/temp': self.stats['temp'][g], 'coeffs/lr': self.stats['lr'][g],
                            'coeffs/r_max': meters['r_max'].avg, 'coeffs/n_exp': m._n_exp_host, 'train/loss_kl': kl,
                            'train/loss_mse': mse, 'train/nelbo_avg': meters['nelbo'].avg, 'train/perdim_kl': meters['perdim_kl'].avg,
                            'train/perdim_mse': meters['perdim_mse'].avg, 'train/weight_reg': 0}
                if cfg.grad_clip is not None:
                    to_write['train/grad_norm'] = meters['grads'].avg
                if kl_diag is not None:
                    n_active = int((host[n * 8 + slots:] > 0.01).sum())  # :442-445
                    to_write['train/n_active'] = n_active
                    to_write['train/n_active_ratio'] = n_active / K
                to_write['gstep'] = g
                self.log.append(to_write)
            if g % 100 == 0:  # :447-449
                meters['grads'].reset()
                meters['nelbo'].reset()

    def iteration(self, data: torch.Tensor, epoch: int = 0):
        """One epoch over a device-resident dataset (N, 1, H, W), like main/train_vae.py:317-459 with a
        batched index-select instead of the DataLoader; returns the epoch's mean nelbo.  Host syncs: one per
        cfg.log_freq steps (the statistics ring) and one at the end of the epoch."""
        cfg = self.cfg
        B = cfg.batch_size
        self.model.train()  # main/train_vae.py:318
        n_batches = data.shape[0] // B  # drop_last
        perm = torch.randperm(data.shape[0], device=data.device)
        self.r_max.zero_()
        total = torch.zeros(3, device=self.device)
        log_freq = max(int(cfg.log_freq), 1)
        if self._ring is None or self._ring.shape[0] != log_freq + 1:
            self._ring = torch.zeros(log_freq + 1, 8, device=self.device)
        ring, pend = self._ring, []
        meters = {k: _AverageMeter() for k in ('nelbo', 'grads', 'r_max', 'perdim_kl', 'perdim_mse')}
        for i in range(n_batches):
            gstep = epoch * n_batches + i
            x = data[perm[i * B:(i + 1) * B]]
            write = gstep > 0 and gstep % log_freq == 0
            slot = ring[len(pend)]
            self._want_kl_diag = write
            self.train_step(x, gstep, loss_out=slot[:4])
            self._want_kl_diag = False
            total += slot[:3]
            if cfg.grad_clip is not None:
                slot[4:5].copy_(self._buffers(x.shape[0])['norm'][:1])
            pend.append(gstep)
            if write or len(pend) == ring.shape[0]:
                kl_diag = self._kl_diag_of_last_step(x.shape[0]) if write else None
                self._flush_stats(pend, ring, meters, write, kl_diag, float(self.betas[min(gstep, self.n_iters - 1)]))
                pend = []
        if pend:
            self._flush_stats(pend, ring, meters, False, None, 0.0)
        nelbo_mse_kl = (total / max(n_batches, 1)).tolist()
        # end-of-epoch n_exp update from the mean of the per-step rate.max(), main/train_vae.py:388,456-457
        dp.allreduce_max(self.r_max, self.pg)
        r_max = self.r_max[:n_batches].mean().item() if n_batches <= self.r_max.numel() else self.r_max.mean().item()
        if r_max > 0:
            self.model.update_n(r_max)
        return nelbo_mse_kl[1] + nelbo_mse_kl[2]


class _AverageMeter:
    """base/utils_model.py AverageMeter: running mean since the last reset."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.sum, self.cnt, self.avg = 0.0, 0, 0.0

    def update(self, v, n=1):
        self.sum += v * n
        self.cnt += n
        self.avg = self.sum / self.cnt
