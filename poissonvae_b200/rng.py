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


# ---- launch stream ---------------------------------------------------------------------------------------
# torch.cuda.current_stream() costs ~10 us of host time per call; a training step issues ~100 launches.  A trainer pins
# the raw stream handle for the duration of a step (the current stream cannot change inside it).
_pinned_stream = None


def launch_stream() -> int:
    """cudaStream_t (as an int) every kernel of this library is launched on: torch's current stream."""
    if _pinned_stream is not None:
        return _pinned_stream
    return torch.cuda.current_stream().cuda_stream


class pinned_launch_stream:
    """Context manager: look the current stream up once and reuse the handle for every launch inside."""

    def __enter__(self):
        global _pinned_stream
        self._prev = _pinned_stream
        if _pinned_stream is None:
            _pinned_stream = torch.cuda.current_stream().cuda_stream
        return _pinned_stream

    def __exit__(self, *exc):
        global _pinned_stream
        _pinned_stream = self._prev
        return False
