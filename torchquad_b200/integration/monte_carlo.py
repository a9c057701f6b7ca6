"""Plain Monte Carlo integration on the GPU (replaces torchquad/integration/monte_carlo.py)."""
import math

import torch

from .. import distributed as tqdist
from .. import ops
from ..integrands import BuiltinIntegrand
from ..utils.set_log_level import logger
from .base_integrator import BaseIntegrator
from .compiled import GraphedIntegrate
from .rng import RNG
from .utils import (_setup_integration_domain, _split_function_values, _to_working_dtype,
                    expand_func_values_and_squeeze_integral)


class MonteCarlo(BaseIntegrator):
    """Monte Carlo integration: I = V/N * sum f(x_j), x_j uniform in the domain.

    Two paths, both CUDA only:
      * arbitrary Python integrand: points are generated once in HBM by `tq_mc_sample` (chunked when
        N*dim exceeds `max_points_bytes`), the integrand is evaluated by torch, and `tq_sum_columns`
        reduces the values in fp64.  Differentiable wrt the domain and integrand parameters.
      * built-in integrand (`torchquad_b200.integrands`): `tq_fused_mc` generates, evaluates and
        accumulates without writing samples; also yields sum f^2 for `get_error_estimate()`.
    """

    max_points_bytes = 8 << 30  # per evaluation chunk of the unfused path

    def __init__(self):
        super().__init__()
        self._moments = None  # (sum f, sum f^2, N, volume) of the last run when available

    def integrate(self, fn, dim, N=1000, integration_domain=None, seed=None, rng=None, backend=None):
        """Integrate `fn` over `integration_domain` with N uniform samples (monte_carlo.py:20-58)."""
        self._check_inputs(dim=dim, N=N, integration_domain=integration_domain)
        domain = _setup_integration_domain(dim, integration_domain, backend)
        if rng is not None and seed is not None:
            raise ValueError("seed and rng cannot both be passed")
        native_rng = rng is None or type(rng) is RNG
        if rng is None:
            rng = RNG(backend="torch", seed=seed)
        self._moments = None
        if isinstance(fn, BuiltinIntegrand) and native_rng and fn.dim == dim and not domain.requires_grad:
            return self._integrate_fused(fn, N, domain, rng)

        rank, world = tqdist.rank_and_world()
        elt = domain.element_size()
        chunk_rows = max(1, self.max_points_bytes // (dim * elt))
        if native_rng and (world > 1 or N > chunk_rows):
            # Rows [begin, end) of ONE stream call: any chunking / rank split draws the same samples.
            begin, end = tqdist.shard_range(N, rank, world)
            call = rng.next_call()
            total, one_d = None, None
            for r0 in range(begin, end, chunk_rows):
                rows = min(chunk_rows, end - r0)
                pts = ops.mc_sample(domain, rows, rng.seed, call, r0, call_offset=rng._call_offset)
                vals, _ = self.evaluate_integrand(fn, pts)
                if one_d is None:  # the reference's 1-D rule (utils.py:235-277): [N] and [N, 1] values give a 0-dim result
                    _, one_d = _split_function_values(vals)
                part = ops.reduce_sum_f64(vals)
                total = part if total is None else total + part
                del pts, vals
            if total is None:  # a rank without rows still takes part in the all-reduce
                probe = fn(domain[:, 0].detach().reshape(1, -1).clone())
                _, one_d = _split_function_values(probe)
                total = torch.zeros(tuple(probe.shape[1:]), dtype=torch.float64, device=domain.device)
            if one_d:
                total = total.reshape(())
            if world > 1:
                total = ops.all_reduce_sum_autograd(total)
            self._nr_of_fevals = N  # evaluations of the whole job, like the grid rules and VEGAS report them
            volume = torch.prod(domain[:, 1] - domain[:, 0])
            return volume * _to_working_dtype(total, domain.dtype) / N
        sample_points = self.calculate_sample_points(N, domain, rng=rng)
        function_values, self._nr_of_fevals = self.evaluate_integrand(fn, sample_points)
        return self.calculate_result(function_values, domain)

    def _integrate_fused(self, fn, N, domain, rng):
        bounds = domain.detach().tolist()
        starts = [b[0] for b in bounds]
        sizes_t = (domain[:, 1] - domain[:, 0]).detach()
        sizes = sizes_t.tolist()  # differences rounded in the working dtype, like the reference
        begin, end = tqdist.shard_range(N)
        call = rng.next_call()
        if rng._call_offset is not None:  # private generator of a compiled integrate: the call counter lives on the device
            call += int(rng._call_offset.item())
        sums = ops.fused_mc(fn.to_struct(starts, sizes, 1.0), domain.dtype, domain.device, begin, end, rng.seed, call)
        tqdist.all_reduce_sum_(sums)
        self._nr_of_fevals = N
        volume = torch.prod(sizes_t)
        self._moments = (sums, N, volume)
        logger.debug("Computed fused Monte Carlo integral")
        return volume * sums[0].to(domain.dtype) / N

    def get_error_estimate(self):
        """One-sigma error V*sqrt(Var f / N) of the last fused run (extension; the reference has none)."""
        if self._moments is None:
            return None
        sums, N, volume = self._moments
        s, q = sums.tolist()
        var = max(q / N - (s / N) ** 2, 0.0)
        return float(volume) * math.sqrt(var / N)

    @expand_func_values_and_squeeze_integral
    def calculate_result(self, function_values, integration_domain):
        """V/N * sum(function_values, axis 0) (monte_carlo.py:60-82)."""
        scales = integration_domain[:, 1] - integration_domain[:, 0]
        volume = torch.prod(scales)
        N = function_values.shape[0]
        return volume * ops.reduce_sum(function_values) / N

    def calculate_sample_points(self, N, integration_domain, seed=None, rng=None):
        """[N, dim] uniform points in the domain (monte_carlo.py:84-106)."""
        if rng is None:
            rng = RNG(backend="torch", seed=seed)
        elif seed is not None:
            raise ValueError("seed and rng cannot both be passed")
        dim = integration_domain.shape[0]
        if type(rng) is RNG:
            return ops.mc_sample(integration_domain, N, rng.seed, rng.next_call(), 0, call_offset=rng._call_offset)
        # injected generator (reference semantics): scale and translate its numbers
        starts = integration_domain[:, 0]
        sizes = integration_domain[:, 1] - starts
        u = rng.uniform(size=[N, dim], dtype=sizes.dtype)
        return u.to(sizes.device) * sizes + starts

    def get_jit_compiled_integrate(self, dim, N=1000, integration_domain=None, seed=None, backend=None,
                                   capture_integrand=False):
        """`compiled_integrate(fn, integration_domain)` with everything but the two arguments fixed
        (monte_carlo.py:108-225).  As in the reference the integrand is evaluated eagerly on every call (it sees
        current Python state, gradients flow); our own steps are single launches.  `capture_integrand=True` (extension)
        captures the whole call -- sampling kernel, the integrand's torch ops, the fp64 reduction -- once per integrand
        as a CUDA graph and replays it, under the frozen-state rule documented in integration/compiled.py.  Every call
        or replay advances the Philox call index on the device: fresh samples per call."""
        self._check_inputs(dim=dim, N=N, integration_domain=integration_domain)
        domain0 = _setup_integration_domain(dim, integration_domain, backend)
        rng = RNG(backend="torch", seed=seed)

        def run(fn, domain, graph_rng):
            return self.integrate(fn, dim, N, domain, rng=graph_rng)

        return GraphedIntegrate(run, domain0, rng, capture_integrand)
