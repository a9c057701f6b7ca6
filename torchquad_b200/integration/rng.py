"""Counter-based random numbers (replaces torchquad/integration/rng.py).

The reference seeds torch's global generator and calls `torch.rand` (rng.py:119-125), which makes the
stream depend on call history, device and torch version.  Here every uniform is a pure function
Philox4x32-10(seed; call index, row, column) evaluated by `tq_philox_uniform`, so
  * the same seed reproduces the same numbers, different seeds differ (tests/rng_test.py:10-42),
  * ranks of a multi-GPU run draw disjoint row ranges of the *same* stream (see distributed.py),
  * nothing has to be saved for autograd: samples are regenerated from the counter.
"""
import os

import torch

from .. import ops
from .utils import _default_device, _require_torch_backend

_MASK64 = 0xFFFFFFFFFFFFFFFF


class RNG:
    """Random number generator with the reference's surface: `RNG(backend, seed).uniform(size, dtype)`."""

    def __init__(self, backend="torch", seed=None, torch_save_state=False):
        _require_torch_backend(backend)
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.seed = int(seed) & _MASK64
        # Each uniform() call (and each fused sampling pass) consumes one call index of the stream.
        self._call = 0
        # The stream never touches torch's global generator, so `torch_save_state` has nothing to protect.
        self._save_state = bool(torch_save_state)
        # Optional int32[1] device word added to the call index inside the sampling kernel: set by the CUDA-graph
        # replay of get_jit_compiled_integrate so that every replay draws fresh samples (integration/compiled.py).
        self._call_offset = None

    def next_call(self):
        """Reserve the next call index (used by the fused kernels, which draw inside the kernel)."""
        c = self._call
        if self._call_offset is None:  # device-counted generators (compiled integrate) advance the device word instead
            self._call += 1
        return c

    def uniform(self, size, dtype, device=None, row_begin=0):
        """Uniform numbers in [0, 1) of shape `size` (list) and torch dtype `dtype`."""
        size = [int(s) for s in (size if isinstance(size, (list, tuple, torch.Size)) else [size])]
        if len(size) == 0:
            size = [1]
        cols = size[-1] if len(size) > 1 else 1
        rows = 1
        for s in (size[:-1] if len(size) > 1 else size):
            rows *= s
        device = torch.device(device) if device is not None else _default_device()
        call = self.next_call()
        if rows == 0 or cols == 0:
            return torch.empty(size, dtype=dtype, device=device)
        if cols > 128:  # very wide rows: fold columns into rows so the Philox block index stays small
            flat = ops.philox_uniform(rows * cols, 1, dtype, device, self.seed, call, row_begin * cols)
            return flat.reshape(size)
        out = ops.philox_uniform(rows, cols, dtype, device, self.seed, call, row_begin)
        return out.reshape(size)
