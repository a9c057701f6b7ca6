"""VEGAS+ adaptive stratified sampling on the GPU (replaces torchquad/integration/vegas_stratification.py).

Names follow the reference (tests/vegas_stratification_test.py): `N_strat, N_cubes, V_cubes, JF, JF2, dh,
strat_counts`.  The reference materialises `repeat(arange(C), nh)` twice per iteration (a device->host sync
each time); here rows locate their cube through the exclusive scan of `nh` (csrc/vegas_strat.cu).
"""
import numpy as np
import torch

from .. import ops
from .rng import RNG
from .utils import _default_device, _require_torch_backend


class VEGASStratification:
    """Hypercube stratification of VEGAS Enhanced (arXiv:2009.05112, section III)."""

    def __init__(self, N_increment, dim, rng, backend="torch", dtype=torch.float32, beta=0.75, device=None, shard=None):
        """`shard` = (log2 block, cubes owned) from `torchquad_b200.distributed.cube_shard`: a multi-GPU fused run keeps
        only this rank's cubes in `dh`, `JF`, `JF2`, ... (block-cyclic deal); `N_cubes` / `V_cubes` stay global."""
        _require_torch_backend(backend)
        self.rng = rng
        self.dim = dim
        # stratification steps per dim, EQ 41 (vegas_stratification.py:27-31)
        self.N_strat = int((N_increment / 4.0) ** (1.0 / dim))
        self.N_strat = 1000 if self.N_strat > 1000 else self.N_strat
        self.beta = beta
        self.N_cubes = self.N_strat**self.dim
        self.V_cubes = (1.0 / self.N_strat) ** self.dim
        if self.N_cubes >= 2**31:
            raise ValueError(f"N_cubes = {self.N_cubes} exceeds the 2^31 cubes supported by the CUDA kernels")
        self.dtype = dtype
        self.backend = "torch"
        self.device = torch.device(device) if device is not None else _default_device()
        self.shard = shard
        self.N_cubes_local = self.N_cubes if shard is None else shard[1]
        zeros = torch.zeros((3, self.N_cubes_local), dtype=dtype, device=self.device)  # one fill; JF and JF2 stay adjacent
        self.JF, self.JF2, self._strat_counts = zeros[0], zeros[1], zeros[2]
        self._counts_stale = False  # set by the native loop: strat_counts = float(nh) is made on first access
        # dh = ones * 1.0 / N_cubes (vegas_stratification.py:43): the working-dtype quotient, computed on the host
        one = (np.float32 if dtype == torch.float32 else np.float64)(1.0)
        self.dh = torch.full([self.N_cubes_local], float(one / type(one)(self.N_cubes)), dtype=dtype, device=self.device)
        self._nh = None        # int64 counts of the current iteration
        self._offsets = None   # their exclusive scan, [N_cubes + 1]
        self.last_scalars = None  # fp64 [4]: I_it, sigma2_it, sum d^beta, sum nh of the last update_DH

    @property
    def strat_counts(self):
        """Samples per cube of the last pass as floats (the reference's attribute)."""
        if self._counts_stale:
            self._strat_counts, self._counts_stale = self._nh.to(self.dtype), False
        return self._strat_counts

    @strat_counts.setter
    def strat_counts(self, value):
        self._strat_counts, self._counts_stale = value, False

    # -- sample counts ------------------------------------------------------------------------
    def get_NH(self, nevals_exp):
        """Samples per cube, EQ 44: max(2, floor(dh * nevals_exp)) as int64 (vegas_stratification.py:92-103)."""
        self._nh, self._offsets = ops.strat_nh(self.dh, nevals_exp)
        return self._nh

    def _offsets_for(self, nevals):
        if self._nh is not None and nevals is self._nh:
            return self._offsets
        nevals = nevals.to(device=self.device, dtype=torch.int64)
        self._nh, self._offsets = nevals, ops.strat_offsets(nevals)
        return self._offsets

    # -- sampling -----------------------------------------------------------------------------
    def _uses_native_rng(self):
        return type(self.rng) is RNG

    def get_Y(self, nevals, num_rows=None):
        """Stratified points in [0,1)^dim, rows sorted by cube (vegas_stratification.py:140-165).

        A plain `RNG` draws inside the kernel from the cube-keyed Philox stream; any other object with a
        `.uniform(size, dtype)` method (the reference's injection mechanism, tests/vegas_test.py:143-156)
        is asked for the [M, dim] block exactly like the reference does."""
        offsets = self._offsets_for(nevals)
        M = int(offsets[-1].item()) if num_rows is None else num_rows
        if self._uses_native_rng():
            return ops.strat_sample(offsets, self.N_strat, self.dim, self.dtype, 0, M,
                                    seed=self.rng.seed, call_idx=self.rng.next_call())
        u = self.rng.uniform(size=[M, self.dim], dtype=self.dtype)
        u = u.to(device=self.device)
        return ops.strat_sample(offsets, self.N_strat, self.dim, self.dtype, 0, M, u_in=u)

    # -- accumulation -------------------------------------------------------------------------
    def accumulate_weight(self, nevals, weight_all_cubes):
        """Per-cube sums of jf and jf^2 (vegas_stratification.py:46-70)."""
        offsets = self._offsets_for(nevals)
        self.JF, self.JF2 = ops.strat_accumulate(weight_all_cubes, offsets)
        self.strat_counts = nevals.to(self.dtype)
        return self.JF, self.JF2

    def update_DH(self):
        """Damped variances -> sampling probabilities dh, EQ 42 (vegas_stratification.py:72-90).

        The same pass yields the iteration estimate (vegas.py:293-303), kept in `last_scalars`."""
        nh = self._nh if self._nh is not None else self.strat_counts.to(torch.int64)
        self.dh, self.last_scalars = ops.strat_update(self.JF, self.JF2, nh, self.V_cubes, self.beta)
