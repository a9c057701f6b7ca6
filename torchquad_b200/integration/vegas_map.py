"""VEGAS+ adaptive map on the GPU (replaces torchquad/integration/vegas_map.py).

State and method names follow the reference so its unit tests read the same (tests/vegas_map_test.py):
`x_edges [dim, Ni+1]`, `dx_edges [dim, Ni]`, `weights [dim, Ni]` (float), `counts [dim, Ni]` (int64).
All arithmetic is in libtqb200 (csrc/vegas_map.cu); the Python loops over `dim` of the reference are gone.
"""
import warnings

import numpy as np
import torch

from .. import ops
from ..utils.set_log_level import logger
from .utils import _default_device, _require_torch_backend


class VEGASMap:
    """The piecewise-linear importance-sampling map of VEGAS Enhanced (arXiv:2009.05112, section II)."""

    def __init__(self, N_intervals, dim, backend="torch", dtype=torch.float32, alpha=0.5, device=None):
        _require_torch_backend(backend)
        self.dim = dim
        self.N_intervals = N_intervals
        self.alpha = alpha
        self.backend = "torch"
        self.dtype = dtype
        self.device = torch.device(device) if device is not None else _default_device()
        # Uniform start (vegas_map.py:32-39): dx = 1/Ni (the quotient ones/Ni of the working dtype, computed once
        # on the host), edges from linspace (not a cumsum of dx).
        one = (np.float32 if dtype == torch.float32 else np.float64)(1.0)
        self.dx_edges = torch.full((dim, N_intervals), float(one / type(one)(N_intervals)), dtype=dtype, device=self.device)
        edges = torch.linspace(0.0, 1.0, N_intervals + 1, dtype=dtype, device=self.device)
        self.x_edges = edges.reshape(1, -1).repeat(dim, 1).contiguous()
        self._status_word = None
        self._edges2, self._edges2_stale = None, True
        self._records, self._records_stale = None, True
        self._scratch = None
        self._hist = self._hist_flat = None
        self._reset_weight()

    @property
    def _status(self):
        if self._status_word is None:
            self._status_word = torch.zeros(4, dtype=torch.int32, device=self.device)
        return self._status_word

    # -- bin lookup ---------------------------------------------------------------------------
    def get_X(self, y):
        """Mapped points x(y), EQ 9 (vegas_map.py:44-58)."""
        return ops.map_forward(y, self.x_edges, self.dx_edges, want_jac=False)[0]

    def get_Jac(self, y):
        """Jacobian of the map, EQ 12 (vegas_map.py:60-74)."""
        return ops.map_forward(y, self.x_edges, self.dx_edges, want_x=False)[1]

    def get_X_and_Jac(self, y):
        """Both in a single pass over y (what the integrator uses)."""
        x, jac, _ = ops.map_forward(y, self.x_edges, self.dx_edges)
        return x, jac

    def _get_interval_ID(self, y):
        """floor(y*Ni) as int64, EQ 10 (vegas_map.py:76-85)."""
        ids = ops.map_forward(y, self.x_edges, self.dx_edges, want_x=False, want_jac=False, want_ids=True)[2]
        return ids.to(torch.int64)

    def _get_interval_offset(self, y):
        """y*Ni - floor(y*Ni), EQ 11 (vegas_map.py:87-97)."""
        return ops.map_forward(y, self.x_edges, self.dx_edges, want_x=False, want_jac=False, want_offset=True)[3]

    # -- histogram ----------------------------------------------------------------------------
    def accumulate_weight(self, y, jf_vec2):
        """weights[d, k] += jf^2 and counts[d, k] += 1 for every sample (vegas_map.py:99-111)."""
        ops.map_accumulate(y, jf_vec2, self.weights, self.counts)

    @staticmethod
    def _smooth_map(weights, counts, alpha):
        """Smoothed, compressed weights, EQ 18-22; None if a dimension sums to zero (vegas_map.py:113-172)."""
        smoothed, status = ops.map_smooth(weights, counts, alpha)
        if int(status[0].item()) != 0:
            return None
        return smoothed

    def _reset_weight(self):
        """Zero the histogram (vegas_map.py:174-183)."""
        n = self.dim * self.N_intervals
        elt = 4 if self.dtype == torch.float32 else 8
        raw = torch.zeros(n * (8 + elt), dtype=torch.uint8, device=self.device)  # one fill for both tables
        self.counts = raw[: n * 8].view(torch.int64).view(self.dim, self.N_intervals)
        self.weights = raw[n * 8:].view(self.dtype).view(self.dim, self.N_intervals)

    # -- rebinning ----------------------------------------------------------------------------
    def update_map(self, check=True, status=None):
        """Adapt the edges to the accumulated weights, section II C (vegas_map.py:185-261).

        The kernels never synchronise: problems are reported through a device status word.  With
        `check=True` (default, reference behaviour) the word is read back here and turned into the
        reference's warnings / RuntimeError; the integrator passes `check=False` and calls
        `check_status()` at its own synchronisation points."""
        st = self._status if status is None else status
        keep_packed = self._edges2 is not None  # refresh the packed copy in the same launch once it exists
        if self._scratch is None:
            self._scratch = ops.map_scratch(self.dim, self.N_intervals, self.dtype, self.device)
        ops.map_update(self.x_edges, self.dx_edges, self.weights, self.counts, self.alpha, st,
                       edges_packed=self._edges2 if keep_packed else None, scratch=self._scratch)
        self._edges2_stale = not keep_packed
        self._records_stale = True
        if check:
            self.check_status(st)

    def packed_edges(self):
        """{x_edges, dx_edges} interleaved [dim, Ni, 2]: the gather layout of the fused / packed kernels.
        Cached; `update_map` invalidates it (call `invalidate_packed()` after editing the edges by hand)."""
        if self._edges2 is None or self._edges2_stale:
            self._edges2 = ops.pack_edges(self.x_edges, self.dx_edges, self._edges2)
            self._edges2_stale = False
        return self._edges2

    def hist_pairs(self):
        """fp64 [dim, Ni, 2] = {sum jf^2, count} accumulator of the fused passes (zero between passes), see
        `ops.fused_vegas` / `ops.unpack_hist`."""
        if self._hist is None:
            # 8 spare words behind the table: the multi-GPU loop appends its per-pass scalars so that one all-reduce
            # carries both (tq_vegas_run_fused_sharded)
            n = self.dim * self.N_intervals * 2
            self._hist_flat = torch.zeros(n + 8, dtype=torch.float64, device=self.device)
            self._hist = self._hist_flat[:n].view(self.dim, self.N_intervals, 2)
        return self._hist

    def unpack_hist(self):
        """Fold the pair accumulator into `weights` / `counts` (and zero it)."""
        ops.unpack_hist(self._hist, self.weights, self.counts)

    def invalidate_packed(self):
        self._edges2_stale = True
        self._records_stale = True

    # Large maps (tables beyond L2): one {x_edge, dx_edge, weight, count} record per bin, so that the gather and
    # both histogram updates of a sample fall into one DRAM sector (include/tqb200.h, TQ_EDGES_RECORDS).
    records_min_bytes = 48 << 20  # use records when they would occupy at least this much (None: never)

    def wants_records(self):
        if self.records_min_bytes is None:
            return False
        elt = 4 if self.dtype == torch.float32 else 8
        return self.dim * self.N_intervals * 4 * elt >= self.records_min_bytes  # = tq_vegas_map_records_bytes

    sweep_l2_budget = 64 << 20  # bytes of one histogram band tq_vegas_hist_sweep keeps in flight (L2 is 126 MB)

    def sweep_group(self, n_strat):
        """Dimensions per launch of `tq_vegas_hist_sweep` for this map (0: not applicable).  Large maps only; the bands of a
        group ((Ni / N_strat) bins x 16 bytes per dimension) must stay resident in L2 while the group's cubes are binned."""
        if not self.wants_records() or n_strat < 2:
            return 0
        # ONE dimension per launch: every bin of the band in flight is then hit ~(rows / Ni) times while it is resident.
        # With g dimensions per launch the cubes are ordered by the g digits together and only the slowest digit's band
        # stays resident across the whole group (measured, 8-D Ni=1e7: g=2 19.8 ms per pass, g=1 see profiles/r2).
        band_bytes = (self.N_intervals // n_strat + 2) * 16
        return 1 if band_bytes <= self.sweep_l2_budget else 0

    def records(self):
        """The record table of the current edges with zeroed histogram fields (opaque uint8 tensor), cached."""
        if self._records is None or self._records_stale:
            self._records = ops.pack_records(self.x_edges, self.dx_edges, self._records)
            self._records_stale = False
        return self._records

    def unpack_records(self):
        """Move the histogram a fused pass left in the records into `weights` / `counts`."""
        ops.unpack_records(self._records, self.weights, self.counts)

    def check_status(self, status=None):
        """Raise / warn like vegas_map.py:188-196,240-257 from a status word (device read-back)."""
        st = (self._status if status is None else status)
        st = st.tolist() if isinstance(st, torch.Tensor) else list(st)
        if st[0]:
            msg = ("Cannot update the VEGASMap. This can happen with an integrand "
                   "which evaluates to zero everywhere.")
            logger.warning(msg)
            warnings.warn(msg, RuntimeWarning)
        if st[1]:
            num_edges = self.x_edges.shape[1]
            msg = f"{st[1]} out of {num_edges * self.dim} calculated VEGASMap edges were infinite"
            logger.warning(msg)
            warnings.warn(msg, RuntimeWarning)
        if st[2]:
            raise RuntimeError("Could not replace all infinite edges")
