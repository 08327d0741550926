"""Philox (seed, offset) pairs for the sampler kernels, taken from torch's CUDA generator so that
`torch.manual_seed` - called by every reference Config constructor, base/config_base.py:51-63 -
governs the draws.  Host-side only: no device synchronisation."""
from __future__ import annotations

import torch


def next_philox(device=None, increment: int = 4):
    """Returns (seed, offset) and advances the device generator's offset (torch needs multiples of 4)."""
    if device is None:
        idx = torch.cuda.current_device()
    else:
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[idx]
    seed = gen.initial_seed() & 0xFFFFFFFFFFFFFFFF
    offset = gen.get_offset()
    gen.set_offset(offset + increment)
    return seed, offset
