"""VEGAS Enhanced (VEGAS+) driver for the GPU (replaces torchquad/integration/vegas.py).

Host control flow only: the iteration schedule, warm-up, chi^2 / abort logic and the weighted mean of
vegas.py:88-209,318-362 are reproduced decision by decision (they determine how many evaluations happen),
but every per-sample and per-bin step runs in libtqb200 kernels, and the >= 7+dim device->host syncs per
iteration of the reference collapse to one (the sample count M, which sizes the launch) plus one packed
read-back of (I_k, sigma^2_k) every fifth iteration.

Paths
  * arbitrary integrand: y -> (x, jac) -> fn(x) by torch -> histogram + per-cube sums; autograd flows through
    fn and the per-cube sums exactly like in the reference (tests/gradient_test.py).
  * built-in integrand (torchquad_b200.integrands): one fused kernel per pass, no sample traffic to HBM.
Multi-GPU (torchquad_b200.distributed.enable()): ranks take contiguous row ranges of the same cube-sorted
sample set and all-reduce [weights | counts | JF | JF2] once per iteration.
"""
import numpy as np
import torch

from .. import _lib
from .. import distributed as tqdist
from .. import ops
from ..integrands import BuiltinIntegrand
from ..utils.set_log_level import logger
from .base_integrator import BaseIntegrator
from .rng import RNG
from .utils import _check_integration_domain, _setup_integration_domain
from .vegas_map import VEGASMap
from .vegas_stratification import VEGASStratification


class _NeedsAutograd(Exception):
    """Raised by the callback of the C++-driven loop when the integrand's values require grad."""


class VEGAS(BaseIntegrator):
    """VEGAS Enhanced, arXiv:2009.05112.  Same surface as the reference class.

    Extension attributes (defaults reproduce the reference):
      regenerate_samples callback integrands with the library's generator: every pass is two kernels around the integrand
                         (ops.sample_map, ops.accumulate_regen) and the stratified samples y are never written to HBM.
                         False: the round-1 pipeline (y materialised, read back for the map and the histogram).
      max_map_intervals  cap on the map size per dimension.  The reference uses Ni = (N // (max_it + 5)) // 10
                         (vegas.py:117), i.e. 1e7 .. 4e7 bins per dimension for N = 2.5e9 .. 1e10: tables far
                         larger than L2 that turn every bin lookup and histogram update into a random HBM
                         access (and collapse numerically in fp32).  None keeps the formula.
      initial_adaptation dict returned by `adaptation_state()` of an earlier run with the same dim / N / dtype
                         (so that the table shapes agree): the run starts from that map and those hypercube
                         probabilities and skips the warm-up.  The reference keeps this state on the integrator
                         (`vegas.py:118-133`) but offers no way to carry it over.
      l2_fetch_bytes     L2 fetch granularity hint used while a large map is in flight (None: leave as is;
                         measured on B200: 32 vs the default 64 makes no difference, profiles/README.md).
    """

    max_map_intervals = None
    l2_fetch_bytes = None
    initial_adaptation = None  # adaptation_state() of an earlier run: start from its map and stratification
    regenerate_samples = True  # callback integrands: never materialise y (False: the round-1 pipeline with y in HBM)
    native_loop = True  # single-GPU runs: drive all passes from C++ (tq_vegas_run_fused / tq_vegas_run_unfused)
    native_unfused_max_bytes = 1 << 30  # sample buffer (x) the callback-integrand loop may allocate up front
    _large_map_bytes = 64 << 20
    _pairs_min_rows = 1 << 20  # fused passes at least this long accumulate the histogram as fp64 pairs
    min_rows_per_rank = 1 << 17  # multi-GPU fused runs with fewer samples per pass and rank are replicated, not sharded

    def __init__(self):
        super().__init__()

    def integrate(self, fn, dim, N=10000, integration_domain=None, seed=None, rng=None, use_grid_improve=True,
                  eps_rel=0, eps_abs=0, max_iterations=20, use_warmup=True, backend=None):
        """Integrate `fn` over the domain with at most ~N evaluations (vegas.py:30-159).

        Returns a 0-dim tensor of the domain's dtype on the domain's device."""
        # Input checks as in the reference; the domain's VALUES are checked below from the single host copy this
        # method makes of them instead of through a separate device reduction + read-back.
        self._check_inputs(dim=dim, N=N)
        if integration_domain is not None and _check_integration_domain(integration_domain, check_values=False) != dim:
            raise ValueError("The dimension of the integration domain must match the passed function dimensionality dim.")
        self._dim = dim
        self._nr_of_fevals = 0
        self._max_iterations = max_iterations
        self._eps_rel = eps_rel
        self._eps_abs = eps_abs
        self.use_grid_improve = use_grid_improve
        self.N = N
        # evaluations per iteration (vegas.py:90-91)
        self._starting_N = N // (self._max_iterations + 5)
        self._N_increment = N // (self._max_iterations + 5)
        domain = _setup_integration_domain(dim, integration_domain, backend)
        self.backend = "torch"
        self.dtype = domain.dtype
        self.device = domain.device
        if rng is None:
            rng = RNG(backend="torch", seed=seed)
        elif seed is not None:
            raise ValueError("seed and rng cannot both be passed")
        self.rng = rng
        self._np = np.float32 if self.dtype == torch.float32 else np.float64

        # unit-cube transform of the integrand (vegas.py:104-112)
        self._starts = domain[:, 0]
        self._sizes = domain[:, 1] - self._starts
        self._volume = torch.prod(self._sizes)
        self._user_fn = fn
        self._fn = lambda x: fn(x * self._sizes + self._starts) * self._volume
        # One read-back brings the bounds and the volume to the host (the only synchronisation of the set-up).
        host = torch.cat((domain.detach().reshape(-1), self._volume.detach().reshape(1))).tolist()
        bounds = [(host[2 * i], host[2 * i + 1]) for i in range(dim)]
        if any(hi < lo for lo, hi in bounds):
            raise ValueError("integration_domain has invalid boundary values")
        # Without a gradient through the domain, the unit-cube -> domain transform and the f*volume*jac tail
        # are fused into the map kernels (same mul/add order as the torch expressions above).
        self._domain = domain
        self._fuse_tail = not domain.requires_grad
        self._volume_host = host[-1] if self._fuse_tail else None

        self._fused = (isinstance(fn, BuiltinIntegrand) and type(rng) is RNG and fn.dim == dim
                       and not domain.requires_grad)
        # Callback integrands with the library's own generator: every pass is two kernels around the integrand and the
        # samples y are never materialised (ops.sample_map / ops.accumulate_regen); injected generators and gradients
        # through the domain keep the materialised pipeline.  dim <= 128: wider rows use another Philox keying (RNG.uniform).
        self._regen = (type(rng) is RNG and self._fuse_tail and not self._fused and dim <= 128 and self.regenerate_samples
                       and getattr(rng, "_call_offset", None) is None)
        if self._fused:
            sizes = [float(self._np(hi) - self._np(lo)) for lo, hi in bounds]  # the working-dtype difference
            self._fn_struct = fn.to_struct([lo for lo, _ in bounds], sizes, host[-1])

        N_intervals = max(2, self._N_increment // 10)  # vegas.py:117
        if self.max_map_intervals is not None:
            N_intervals = max(2, min(N_intervals, int(self.max_map_intervals)))
        self.map = VEGASMap(N_intervals, dim, "torch", self.dtype, device=self.device)
        # Multi-GPU, built-in integrand: the hypercubes are dealt to the ranks (block-cyclic), each rank keeps only its
        # share of the stratification state and a pass needs ONE all-reduce (tq_vegas_run_fused_sharded).
        self._shard = None
        self._replicated = False
        if (tqdist.is_enabled() and self._fused and self.native_loop and self.initial_adaptation is None
                and max_iterations + 5 <= _lib.TQ_VEGAS_MAX_PASSES):
            # Passes of a few 1e4 samples are pure launch latency: sharding them only adds a collective per
            # pass.  Every rank then runs the whole problem (identical samples => identical results, no collective).
            self._replicated = self._N_increment < self.min_rows_per_rank * tqdist.rank_and_world()[1]
        if (tqdist.is_enabled() and self._fused and self.native_loop and self.initial_adaptation is None
                and max_iterations + 5 <= _lib.TQ_VEGAS_MAX_PASSES and not self._replicated):
            n_strat = min(1000, int((self._N_increment / 4.0) ** (1.0 / dim)))
            self._shard = tqdist.cube_shard(n_strat**dim)
        self.strat = VEGASStratification(self._N_increment, dim=dim, rng=self.rng, backend="torch", dtype=self.dtype,
                                         device=self.device, shard=self._shard)
        # maps beyond L2, callback integrand, one GPU: the histogram of a stratified pass is binned band by band
        self._regen_sweep = 0 if tqdist.is_enabled() else self.map.sweep_group(self.strat.N_strat)
        # Multi-GPU, arbitrary integrand: the float statistics of a pass live in ONE buffer [weights | JF | JF2] so that
        # a pass needs one all-reduce for them plus one for the int64 counts.
        self._stats = None
        if tqdist.is_enabled() and self._shard is None and not self._replicated:
            n_w = dim * N_intervals
            self._stats = torch.zeros(n_w + 2 * self.strat.N_cubes, dtype=self.dtype, device=self.device)
            self.map.weights = self._stats[:n_w].view(dim, N_intervals)
            self._stats_w = self._stats[:n_w]
            self._stats_jf = self._stats[n_w:].view(2, self.strat.N_cubes)
        if self.initial_adaptation is not None:
            self._load_adaptation(self.initial_adaptation)
            use_warmup = False
        self.results = []  # per-iteration integral estimates (0-dim tensors)
        self.sigma2 = []   # per-iteration variances (0-dim tensors, detached)
        self.it = 0
        self._host_block = None
        self._status_used = 0

        # random-access regime: optionally change the L2 fetch granularity while the big tables are in flight
        restore_l2 = None
        if (self.l2_fetch_bytes and self.device.type == "cuda"
                and dim * N_intervals * (2 * domain.element_size() + 8) > self._large_map_bytes):
            restore_l2 = _lib.l2_fetch_granularity(self.device, int(self.l2_fetch_bytes))
        try:
            if self._shard is not None or self._replicated:
                return self._integrate_native_loop(N, use_warmup)
            native = self.native_loop and not tqdist.is_enabled() and max_iterations + 5 <= _lib.TQ_VEGAS_MAX_PASSES
            if native and self._fused:
                return self._integrate_native_loop(N, use_warmup)
            if native and self._regen:  # callback integrand, library generator: the C++ loop (two kernels per pass)
                # a pass has at most starting_N * sum(dh) + 2 * N_cubes rows and the schedule keeps 5 * starting_N <= N
                cap_rows = N // 5 + N // 500 + 2 * self.strat.N_cubes + 4096
                if cap_rows * dim * domain.element_size() <= self.native_unfused_max_bytes:
                    try:
                        return self._integrate_native_unfused(N, use_warmup, cap_rows)
                    except _NeedsAutograd:
                        # the integrand's values carry a graph: start over with the differentiable Python loop
                        self._nr_of_fevals = 0
                        self.rng._call = self._first_call
                        self.map = VEGASMap(N_intervals, dim, "torch", self.dtype, device=self.device)
                        self.strat = VEGASStratification(self._N_increment, dim=dim, rng=self.rng, backend="torch",
                                                         dtype=self.dtype, device=self.device)
                        if self.initial_adaptation is not None:
                            self._load_adaptation(self.initial_adaptation)
            # one status word per map update, written by the kernels, read back in one go at the sync points
            self._status_buf = torch.zeros((max_iterations + 16, 4), dtype=torch.int32, device=self.device)
            if use_warmup:
                self._warmup_grid(5, self._starting_N // 5)
            while True:
                self.it += 1
                self.results.append(0)
                self.sigma2.append(0)
                self._run_iteration()
                if self._check_abort_conditions():
                    break
        finally:
            if restore_l2:
                _lib.l2_fetch_granularity(self.device, restore_l2)
        self._flush_map_status()
        logger.debug("VEGAS finished")
        return self._get_result()

    # ------------------------------------------------------------------ save / resume of the adaptation
    def adaptation_state(self):
        """The adapted map and stratification of the last run as a dict of CPU tensors (torch.save-able)."""
        vmap, strat = self.map, self.strat
        return {"dim": vmap.dim, "N_intervals": vmap.N_intervals, "N_strat": strat.N_strat, "dtype": str(vmap.dtype),
                "x_edges": vmap.x_edges.detach().cpu().clone(), "dx_edges": vmap.dx_edges.detach().cpu().clone(),
                "dh": strat.dh.detach().cpu().clone()}

    def _load_adaptation(self, state):
        vmap, strat = self.map, self.strat
        want = (vmap.dim, vmap.N_intervals, strat.N_strat, str(vmap.dtype))
        got = (state["dim"], state["N_intervals"], state["N_strat"], state["dtype"])
        if want != got:
            raise ValueError(f"initial_adaptation was made for (dim, N_intervals, N_strat, dtype) = {got}, this run needs {want}")
        vmap.x_edges.copy_(state["x_edges"])
        vmap.dx_edges.copy_(state["dx_edges"])
        vmap.invalidate_packed()
        strat.dh.copy_(state["dh"])

    def _integrate_native_unfused(self, N, use_warmup, cap_rows):
        """Single-GPU run with an arbitrary Python integrand and the pass loop in C++ (tq_vegas_run_unfused): per
        pass Python only evaluates the integrand on a view of the sample buffer.  Same kernels and Philox call
        indices as the Python-driven loop below; gradients need that loop (`_NeedsAutograd`)."""
        self._first_call = self.rng._call

        def evaluate(x):
            f_raw, n = self.evaluate_integrand(self._user_fn, x)
            f_raw = f_raw.reshape(-1) if f_raw.numel() == n else f_raw.squeeze()
            if torch.is_grad_enabled() and f_raw.requires_grad:
                raise _NeedsAutograd()
            return f_raw if f_raw.dtype == self.dtype else f_raw.to(self.dtype)

        res = ops.vegas_run_unfused(evaluate, self.map, self.strat, self._domain, self._volume_host, cap_rows, N,
                                    self._max_iterations, self._eps_rel, self._eps_abs, self.use_grid_improve, use_warmup,
                                    self.rng.seed, self._first_call)
        return self._finish_native(res)

    def _integrate_native_loop(self, N, use_warmup):
        """Fused single-GPU run with the pass loop and schedule in C++ (csrc/vegas_driver.cu): same kernels, same
        Philox call indices and therefore the same samples as the Python-driven loop below."""
        first_call = self.rng._call
        shard = None
        if self._shard is not None:
            rank, world = tqdist.rank_and_world()
            shard = (rank, world, self._shard[0], tqdist.all_reduce_sum_)
        res = ops.vegas_run_fused(self._fn_struct, self.map, self.strat, N, self._max_iterations, self._eps_rel,
                                  self._eps_abs, self.use_grid_improve, use_warmup, self.rng.seed, first_call, shard=shard)
        return self._finish_native(res)

    def _finish_native(self, res):
        """Adopt the outcome of a C++-driven run (tq_vegas_result) as the integrator's state."""
        self.rng._call = res.calls_used
        self.it = res.it
        self._nr_of_fevals = res.fevals
        self._starting_N = res.starting_N
        nb = res.n_block
        block = torch.tensor(list(res.results[:nb]) + list(res.sigma2[:nb]), dtype=torch.float64, device=self.device)
        self.results = list(block[:nb].unbind())  # 0-dim device tensors like the reference's, from one copy
        self.sigma2 = list(block[nb:].unbind())
        self._host_cache = ([self._np(res.results[k]) for k in range(nb)], [self._np(res.sigma2[k]) for k in range(nb)])
        self.map._edges2_stale = False
        for p in range(res.n_passes):
            self.map.check_status(list(res.status[4 * p: 4 * p + 4]))
        with np.errstate(all="ignore"):
            mean = self._weighted_mean(*self._host_cache)
        return torch.tensor(mean, dtype=self.dtype, device=self.device)

    # ------------------------------------------------------------------ passes
    def _rank_rows(self, total):
        return tqdist.shard_range(total)

    def _reduce_stats(self, with_cubes):
        """Sum the pass statistics over the ranks: [weights | JF | JF2] in one collective, counts in another."""
        if self._stats is None:
            return
        if with_cubes:
            tqdist.all_reduce_sum_(self._stats, self.map.counts)
        else:
            tqdist.all_reduce_sum_(self._stats_w, self.map.counts)

    def _update_map(self):
        if self._status_used == self._status_buf.shape[0]:
            self._flush_map_status()
        self.map.update_map(check=False, status=self._status_buf[self._status_used])
        self._status_used += 1

    def _flush_map_status(self):
        """Turn accumulated device status words into the reference's warnings / errors (one read-back)."""
        if not self._status_used:
            return
        words = self._status_buf[: self._status_used].tolist()
        self._status_used = 0
        for w in words:
            self.map.check_status(w)

    def _fused_pass(self, begin, end, hist, **strat_args):
        """One fused pass over rows [begin, end).  Large maps go through the record layout: the histogram lands in
        the records and is moved to `weights` / `counts` (what the all-reduce and `update_map` read) right after."""
        vmap = self.map
        if hist and vmap.wants_records():
            ops.fused_vegas(self._fn_struct, None, None, None, begin, end, self.rng.seed, self.rng.next_call(),
                            records=vmap.records(), dtype=self.dtype, n_intervals=vmap.N_intervals, **strat_args)
            vmap.unpack_records()
        elif hist and end - begin >= self._pairs_min_rows:
            # big passes: {sum jf^2, count} fp64 pairs, one reduction sector per sample and dimension (ops.fused_vegas)
            ops.fused_vegas(self._fn_struct, vmap.packed_edges(), None, None, begin, end, self.rng.seed, self.rng.next_call(),
                            hist_pairs=vmap.hist_pairs(), **strat_args)
            vmap.unpack_hist()
        else:
            ops.fused_vegas(self._fn_struct, vmap.packed_edges(), vmap.weights if hist else None, vmap.counts, begin, end,
                            self.rng.seed, self.rng.next_call(), **strat_args)

    def _warmup_grid(self, warmup_N_it=5, N_samples=1000):
        """Adapt the map with unstratified passes whose results are discarded (vegas.py:211-266)."""
        for _ in range(warmup_N_it):
            begin, end = self._rank_rows(N_samples)
            if self._fused:
                self._fused_pass(begin, end, hist=True)
                self._nr_of_fevals += N_samples
            else:
                regen = self._regen
                if regen:
                    # two kernels around the integrand: x, jac straight from the Philox stream; the bins are regenerated
                    # for the histogram, so the warm-up samples y never exist in HBM (ops.sample_map / accumulate_regen)
                    call_idx = self.rng.next_call()
                    yrnd = None
                    x, jac = self._sample_map(None, begin, end, call_idx)
                    f_raw = self._eval_raw(x)
                    del x
                elif type(self.rng) is RNG:
                    yrnd = self.rng.uniform([end - begin, self._dim], self.dtype, device=self.device, row_begin=begin)
                    yrnd = yrnd * 0.999999
                    f_raw, jac = self._map_and_eval(yrnd)
                else:
                    yrnd = self.rng.uniform(size=[N_samples, self._dim], dtype=self.dtype).to(self.device) * 0.999999
                    yrnd = yrnd[begin:end]
                    f_raw, jac = self._map_and_eval(yrnd)
                if tqdist.is_enabled():
                    self._nr_of_fevals += N_samples - (end - begin)
                if f_raw is not None and not (torch.is_grad_enabled() and f_raw.requires_grad):
                    if regen:
                        self._accumulate_regen(None, begin, end, call_idx, f_raw, jac, hist=True, want_jf=False)
                    else:
                        self._accumulate_tail(yrnd, f_raw, jac, want_jf=False)
                else:
                    if regen:  # the integrand's values carry a graph: the reference expressions need the samples
                        yrnd = ops.philox_uniform(end - begin, self._dim, self.dtype, self.device, self.rng.seed, call_idx, begin) * 0.999999
                        self._last_f_eval = f_raw * self._volume
                    f_eval = self._last_f_eval
                    jf_vec2 = ((f_eval * jac) ** 2).detach()
                    self.map.accumulate_weight(yrnd, jf_vec2)
            self._reduce_stats(with_cubes=False)
            self._update_map()

    def _run_iteration(self):
        """One stratified VEGAS iteration (vegas.py:268-315)."""
        strat, vmap = self.strat, self.map
        neval = strat.get_NH(self._starting_N)
        offsets = strat._offsets
        if tqdist.is_enabled() and not self._fused:
            M, cube_lo, cube_hi, begin, end = self._cube_aligned_shard(offsets)
        else:
            M = int(offsets[-1].item())  # the one sync of the iteration: sample count sizes the launches
            begin, end = self._rank_rows(M)
            cube_lo, cube_hi = 0, strat.N_cubes
        grad_path = False
        if self._fused:
            if self._stats is not None:
                JFs = self._stats_jf.zero_()
            else:
                JFs = torch.zeros((2, strat.N_cubes), dtype=self.dtype, device=self.device)
            self._fused_pass(begin, end, hist=self.use_grid_improve, offsets=offsets, n_strat=strat.N_strat, JF=JFs[0],
                             JF2=JFs[1])
            self._nr_of_fevals += M
            self._reduce_stats(with_cubes=True)
            strat.JF, strat.JF2 = JFs[0], JFs[1]
        else:
            regen = self._regen
            if regen:
                call_idx = self.rng.next_call()
                y = None
                x, jac = self._sample_map(offsets, begin, end, call_idx)
                f_raw = self._eval_raw(x)
                del x
            elif type(self.rng) is RNG:
                y = ops.strat_sample(offsets, strat.N_strat, self._dim, self.dtype, begin, end, seed=self.rng.seed,
                                     call_idx=self.rng.next_call())
                f_raw, jac = self._map_and_eval(y)
            else:
                u = self.rng.uniform(size=[M, self._dim], dtype=self.dtype).to(self.device)
                y = ops.strat_sample(offsets, strat.N_strat, self._dim, self.dtype, begin, end,
                                     u_in=u[begin:end].contiguous())
                f_raw, jac = self._map_and_eval(y)
            if tqdist.is_enabled():
                self._nr_of_fevals += M - (end - begin)
            if f_raw is not None and not (torch.is_grad_enabled() and f_raw.requires_grad):
                if regen:
                    jf_vec = self._accumulate_regen(offsets, begin, end, call_idx, f_raw, jac, hist=self.use_grid_improve,
                                                    want_jf=True, whole_pass=(begin == 0 and end == M))
                elif self.use_grid_improve:
                    jf_vec = self._accumulate_tail(y, f_raw, jac, want_jf=True)
                else:
                    jf_vec = (f_raw * self._volume.detach()) * jac
            else:
                if regen:  # the integrand's values carry a graph: the reference expressions need the samples
                    y = ops.strat_sample(offsets, strat.N_strat, self._dim, self.dtype, begin, end, seed=self.rng.seed,
                                         call_idx=call_idx)
                    self._last_f_eval = f_raw * self._volume
                f_eval = self._last_f_eval
                jf_vec = f_eval * jac
                if self.use_grid_improve:
                    vmap.accumulate_weight(y, (jf_vec**2).detach())
            grad_path = torch.is_grad_enabled() and jf_vec.requires_grad
            JF, JF2 = ops.strat_accumulate(jf_vec, offsets, row_base=begin, cube_begin=cube_lo, cube_end=cube_hi)
            if self._stats is not None:
                both = self._stats_jf
                both[0].copy_(JF.detach())
                both[1].copy_(JF2)
                self._reduce_stats(with_cubes=True)
                JF = JF + (both[0] - JF.detach()) if grad_path else both[0]
                JF2 = both[1]
            strat.JF, strat.JF2 = JF, JF2
            strat._counts_stale = True  # strat_counts = float(nh) is materialised on first access

        # estimator + damped-variance update in one kernel (vegas.py:293-303, vegas_stratification.py:72-90)
        strat.update_DH()
        scal = strat.last_scalars
        # scal[0], scal[1] are fp64 views (no kernel); they are rounded to the working dtype when read back
        if grad_path:
            inv = 1.0 / neval.to(self.dtype)
            self.results[-1] = (strat.JF * (inv * strat.V_cubes)).sum()
        else:
            self.results[-1] = scal[0]
        self.sigma2[-1] = scal[1]
        if self.use_grid_improve:
            self._update_map()

    def _accumulate_tail(self, y, f_raw, jac, want_jf):
        """jf = f*V*jac and the map histogram in one kernel; large maps accumulate into their record table and the
        histogram is moved to `weights` / `counts` (what the all-reduce and `update_map` read) right after."""
        vmap = self.map
        if vmap.wants_records():
            jf = ops.accumulate_fused(y, f_raw, jac, self._volume_host, None, None, want_jf=want_jf, records=vmap.records(),
                                      n_intervals=vmap.N_intervals)
            vmap.unpack_records()
            return jf
        return ops.accumulate_fused(y, f_raw, jac, self._volume_host, vmap.weights, vmap.counts, want_jf=want_jf)

    def _sample_map(self, offsets, begin, end, call_idx):
        """(x in domain coordinates, jac) of rows [begin, end) of a pass without materialising y (ops.sample_map)."""
        vmap = self.map
        n_strat = self.strat.N_strat if offsets is not None else 1
        if vmap.wants_records() and self._regen_sweep == 0:
            return ops.sample_map(offsets, n_strat, self._dim, self.dtype, begin, end, self.rng.seed, call_idx,
                                  self._domain.detach(), records=vmap.records(), n_intervals=vmap.N_intervals)
        return ops.sample_map(offsets, n_strat, self._dim, self.dtype, begin, end, self.rng.seed, call_idx,
                              self._domain.detach(), edges_packed=vmap.packed_edges())

    def _accumulate_regen(self, offsets, begin, end, call_idx, f_raw, jac, hist, want_jf, whole_pass=False):
        """jf and the map histogram of the rows of `_sample_map`, bins regenerated from the Philox stream.  Maps beyond L2,
        whole stratified passes: jf^2 rows + band-ordered sweeps; otherwise their record table; ordinary
        maps: the fp64 pair table.  The histogram is folded into weights / counts (what update_map and the all-reduce read)."""
        vmap = self.map
        n_strat = self.strat.N_strat if offsets is not None else 1
        args = (offsets, n_strat, self._dim, begin, end, vmap.N_intervals, f_raw, jac, self._volume_host, self.rng.seed, call_idx)
        if not hist:
            return ops.accumulate_regen(*args, want_jf=want_jf)[0]
        if vmap.wants_records() and self._regen_sweep == 0:
            jf, _ = ops.accumulate_regen(*args, records=vmap.records(), want_jf=want_jf)
            vmap.unpack_records()
            return jf
        if vmap.wants_records() and whole_pass and offsets is not None:
            jf, jf2 = ops.accumulate_regen(*args, want_jf=want_jf, want_jf2=True)
            ops.hist_sweep(offsets, n_strat, self._dim, jf2, vmap.N_intervals, vmap.hist_pairs(), self._regen_sweep, self.rng.seed,
                           call_idx)
        else:
            jf, _ = ops.accumulate_regen(*args, hist_pairs=vmap.hist_pairs(), want_jf=want_jf)
        vmap.unpack_hist()
        return jf

    def _eval_raw(self, x):
        """The user's integrand on x (already in domain coordinates) as a flat vector of the working dtype."""
        f_raw, n = self.evaluate_integrand(self._user_fn, x)
        self._nr_of_fevals += n
        f_raw = f_raw.reshape(-1) if f_raw.numel() == n else f_raw.squeeze()
        if f_raw.dtype != self.dtype:
            f_raw = f_raw.to(self.dtype)
        return f_raw

    def _map_and_eval(self, y):
        """y -> (raw integrand values, jac).  Fused tail: x comes out of the map kernel already in domain
        coordinates and the raw values are returned for `accumulate_fused`; otherwise (gradient through the
        domain) the reference's torch expressions are used and (None, jac) is returned with the scaled values
        left in `self._last_f_eval`."""
        if self._fuse_tail:
            if self.map.wants_records():  # large map: gather the edge pairs out of the record table
                x, jac, _ = ops.map_forward_packed(y, None, self._domain.detach(), records=self.map.records(),
                                                   n_intervals=self.map.N_intervals)
            else:
                x, jac, _ = ops.map_forward_packed(y, self.map.packed_edges(), self._domain.detach())
            f_raw = self._eval_raw(x)
            if torch.is_grad_enabled() and f_raw.requires_grad:
                self._last_f_eval = f_raw * self._volume
                return f_raw, jac
            return f_raw, jac
        x, jac = self.map.get_X_and_Jac(y)
        f_eval = self._eval(x).squeeze()
        if f_eval.dim() == 0:
            f_eval = f_eval.reshape(1)
        self._last_f_eval = f_eval
        return None, jac

    def _cube_aligned_shard(self, offsets):
        """(M, cube_lo, cube_hi, row_lo, row_hi) of this rank: contiguous cube ranges balanced by rows.

        The unfused path sums each cube's rows in order inside one thread, so its shards end on cube
        boundaries; everything is computed on the device and read back once."""
        rank, world = tqdist.rank_and_world()
        M_dev = offsets[-1]
        targets = torch.stack([(M_dev * rank) // world, (M_dev * (rank + 1)) // world])
        cubes = torch.searchsorted(offsets, targets, right=False).clamp(max=offsets.shape[0] - 1)
        rows = offsets[cubes]
        M, c0, c1, r0, r1 = torch.cat([M_dev.reshape(1), cubes, rows]).tolist()
        return M, c0, c1, r0, r1

    # ------------------------------------------------------------------ schedule (host)
    def _host_values(self):
        """(results, sigma2) of the current block as numpy scalars of the working dtype (one read-back)."""
        packed = torch.stack([r.detach().to(torch.float64) for r in self.results]
                             + [s.detach().to(torch.float64) for s in self.sigma2]).cpu().numpy()
        k = len(self.results)
        return [self._np(v) for v in packed[:k]], [self._np(v) for v in packed[k:]]

    @staticmethod
    def _weighted_mean(results, sigma2):
        """EQ 30 with the reference's zero-variance rule (vegas.py:318-335); works on tensors and numpy scalars."""
        if any(s == 0.0 for s in sigma2):
            return sum(results) / len(results)
        num = sum(r / s for r, s in zip(results, sigma2))
        den = sum(1.0 / s for s in sigma2)
        return num / den

    def _check_abort_conditions(self):
        """Every fifth iteration: stop, or grow the per-iteration budget and reset (vegas.py:161-209)."""
        if self.it % 5 > 0:
            return False
        self._flush_map_status()
        res, sig = self._host_values()
        with np.errstate(all="ignore"):
            mean = self._weighted_mean(res, sig)
            res_abs = abs(mean)
            inv = sum(self._np(1.0) / s for s in sig if s != 0.0)
            err = sig[0] if inv == 0 else self._np(1.0) / np.sqrt(inv)
            chi2 = sum(((r - mean) ** 2 / s for r, s in zip(res, sig) if r != mean), start=res[0] * self._np(0.0))
            if (err <= self._eps_rel * res_abs or err <= self._eps_abs) and chi2 / 5.0 < 1.0:
                return True
            if chi2 / 5.0 < 1.0:
                if res_abs == 0.0:
                    self._starting_N += self._N_increment
                else:
                    acc = err / res_abs
                    self._starting_N = min(
                        self._starting_N + self._N_increment,
                        int(self._starting_N * np.sqrt(acc / self._np(self._eps_rel + 1e-8))),
                    )
            elif chi2 / 5.0 > 1.0:
                self._starting_N += self._N_increment
        if self._nr_of_fevals + self._starting_N * 5 > self.N:
            return True
        if self.it + 5 > self._max_iterations:
            return True
        self.results = []
        self.sigma2 = []
        return False

    # ------------------------------------------------------------------ results
    def _get_result(self):
        """Inverse-variance weighted mean of the current block, EQ 30 (vegas.py:318-335).

        Without autograd the mean is formed from the host copy of the block (same dtype, same operation
        order) and uploaded once; with autograd it is built from the result tensors so gradients flow."""
        if any(isinstance(r, torch.Tensor) and r.requires_grad for r in self.results):
            _, sig = self._host_values()
            if any(s == 0.0 for s in sig):
                return sum(self.results) / len(self.results)
            num = sum(r / float(s) for r, s in zip(self.results, sig))
            den = sum(self._np(1.0) / s for s in sig)
            return num / float(den)
        res, sig = self._host_values()
        with np.errstate(all="ignore"):
            mean = self._weighted_mean(res, sig)
        return torch.tensor(mean, dtype=self.dtype, device=self.device)

    def _get_error(self):
        """Error estimate from the variances, EQ 31 (vegas.py:337-346)."""
        res = sum(1.0 / s for s in self.sigma2 if s != 0.0)
        return self.sigma2[0] if res == 0 else 1.0 / torch.sqrt(res)

    def _get_chisq(self):
        """Chi square of the block, EQ 32 (vegas.py:348-362)."""
        I_final = self._get_result()
        return sum(((r - I_final) ** 2 / s for r, s in zip(self.results, self.sigma2) if r != I_final),
                   start=self.results[0] * 0.0)
