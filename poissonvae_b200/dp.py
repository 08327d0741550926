"""Data-parallel plumbing of the training step (torch.distributed: NCCL on the GPUs, gloo in CPU tests).

The P-VAE shards only over the batch: every rank owns a contiguous slice of the
global batch, runs the same step on it with gradients already scaled by 1 / B_global, and one sum
all-reduce of the flat gradient buffer gives every rank the global gradient (the reference has no
distributed path; its loss is a mean over the batch, main/train_vae.py:351).  The Philox stream is keyed
by the global sample index (row_offset), so the draws do not depend on the number of ranks.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None):
    """(world_size, rank) of the default / given process group; (1, 0) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_rows(global_batch: int, world: int, rank: int):
    """Contiguous equal slices: returns (row_offset, local_batch).  The global batch must divide evenly
    (the mean over the batch weights every sample by 1 / B_global on every rank)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    local = global_batch // world
    return rank * local, local


def allreduce_gradients(flat_grad: torch.Tensor, loss_stats: torch.Tensor = None, group=None):
    """In-place sum over ranks of the flat gradient buffer (and of the [loss, mse, kl] partial means)."""
    world, _ = world_info(group)
    if world == 1:
        return
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    if loss_stats is not None:
        dist.all_reduce(loss_stats, op=dist.ReduceOp.SUM, group=group)


def allreduce_max(t: torch.Tensor, group=None):
    world, _ = world_info(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)


class PeerExchange:
    """Symmetric (peer-mapped) gradient double buffer + flag array for `pvae_allreduce_adamax`: every rank can read
    every other rank's gradients over NVLink / NVSwitch, so the all-reduce and the optimiser update become one kernel.
    Built on torch.distributed._symmetric_memory (allocation + handle exchange only - the kernel is ours)."""

    def __init__(self, n_ext: int, device, group=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.n_ext = n_ext
        self.buf = symm_mem.empty(2 * n_ext, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)
        self.buf.zero_()
        self.flags.zero_()
        torch.cuda.synchronize(device)
        h_buf = symm_mem.rendezvous(self.buf, grp)
        h_flag = symm_mem.rendezvous(self.flags, grp)
        dist.barrier(grp)  # nobody publishes a token before every rank has zeroed its flags
        self.world, self.rank = h_buf.world_size, h_buf.rank
        ptrs = [int(p) for p in h_buf.buffer_ptrs]
        self.grad_ptrs = [(ctypes.c_void_p * self.world)(*[p + parity * n_ext * 4 for p in ptrs]) for parity in (0, 1)]
        self.flag_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in h_flag.buffer_ptrs])
        self._handles = (h_buf, h_flag)  # keep the mappings alive

    def grads(self, parity: int) -> torch.Tensor:
        """This rank's gradient buffer (n_ext floats) for the given step parity."""
        return self.buf[parity * self.n_ext:(parity + 1) * self.n_ext]
