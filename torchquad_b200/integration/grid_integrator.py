"""Grid-based integrators on the GPU (replaces torchquad/integration/grid_integrator.py).

The reference reshapes the function values to [..., n, ..., n] and applies the composite rule as `dim`
successive stencil-and-sum passes over the last axis (grid_integrator.py:57-91 and the rule files).  The
same number is  sum_p f(p) * prod_d w[i_d(p)] * prod_d (h_d * c)  with the rule's 1-D weight pattern w, and
that is what `tq_nc_contract` / `tq_fused_nc` evaluate in a single fp64-accumulated pass."""
import torch

from .. import distributed as tqdist
from .. import ops
from ..integrands import BuiltinIntegrand
from .base_integrator import BaseIntegrator
from .compiled import GraphedIntegrate
from .integration_grid import IntegrationGrid, grid_nodes
from .utils import (_check_integration_domain, _is_compiling, _linspace_with_grads, _setup_integration_domain,
                    _split_function_values, _to_working_dtype, expand_func_values_and_squeeze_integral)


class GridIntegrator(BaseIntegrator):
    """Base of the composite Newton-Cotes integrators."""

    max_points_bytes = 8 << 30  # evaluation chunk of the unfused path

    # 1-D weight pattern of the rule and the divisor of h: subclasses override.
    _rule_denominator = 1.0

    def __init__(self):
        super().__init__()

    @property
    def _grid_func(self):
        def f(integration_domain, N, requires_grad=False, backend=None):
            return _linspace_with_grads(integration_domain[0], integration_domain[1], N, requires_grad=requires_grad)

        f._equally_spaced = True
        return f

    def _weights(self, N, dim, backend, requires_grad=False):
        return None

    @staticmethod
    def _rule_weights_1d(n, dtype, device):
        raise NotImplementedError

    @staticmethod
    def _adjust_N(dim, N):
        return N

    _tables = {}  # (rule, n, dim, dtype, device) -> [dim, n] weight table; built once, treated as read-only

    def _weight_table(self, n, dim, dtype, device):
        key = (type(self).__name__, n, dim, dtype, str(device))
        table = GridIntegrator._tables.get(key)
        if table is None:
            w = self._rule_weights_1d(n, dtype, device)
            table = w.reshape(1, n).repeat(dim, 1).contiguous()
            if len(GridIntegrator._tables) > 64:
                GridIntegrator._tables.clear()
            GridIntegrator._tables[key] = table
        return table

    def _scale(self, hs, domain=None):
        """prod_d h_d / c, multiplied in the order the reference applies its passes."""
        q = (hs / self._rule_denominator).unbind()  # one division kernel, then views
        s = q[0]
        for d in range(1, len(q)):
            s = s * q[d]
        return s

    def integrate(self, fn, dim, N, integration_domain, backend):
        """Composite Newton-Cotes integration (grid_integrator.py:32-55)."""
        if N is None:
            N = self._get_minimal_N(dim)
        domain = _setup_integration_domain(dim, integration_domain, backend)
        # One read-back of the domain serves the reference's value check (utils.py:196-205) and the node vectors.
        self._check_inputs(dim=dim, N=N)
        if _check_integration_domain(domain, check_values=False) != dim:
            raise ValueError("The dimension of the integration domain must match the passed function dimensionality dim.")
        host_bounds = None
        if not domain.requires_grad and not _is_compiling(domain):
            host_bounds = domain.detach().tolist()
            if any(hi < lo for lo, hi in host_bounds):
                raise ValueError("integration_domain has invalid boundary values")
        N = self._adjust_N(dim=dim, N=N)
        rank, world = tqdist.rank_and_world()
        fused = isinstance(fn, BuiltinIntegrand) and fn.dim == dim and not domain.requires_grad
        n = int(N ** (1.0 / dim) + 1e-8)
        total_points = n**dim
        chunk_rows = max(1, self.max_points_bytes // (dim * domain.element_size()))
        if not fused and world == 1 and total_points <= chunk_rows:
            grid_func = self._grid_func
            if host_bounds is not None and getattr(grid_func, "_equally_spaced", False):
                IntegrationGrid._check_inputs(None, N, domain, True)
                nodes, hs, n_per_dim = grid_nodes(N, domain, grid_func, host_bounds=host_bounds)
                grid_points = ops.nc_grid_points(nodes)
            else:
                grid_points, hs, n_per_dim = self.calculate_grid(N, domain, disable_integration_domain_check=True)
            function_values, num_points = self.evaluate_integrand(fn, grid_points)
            self._nr_of_fevals = num_points
            if hasattr(self, "integrate_values"):  # Gaussian rules: weights go into the contraction kernel
                return self._squeeze_1d(function_values, self.integrate_values, dim, n_per_dim, domain)
            return self.calculate_result(function_values, dim, n_per_dim, hs, domain)

        # sharded / chunked / fused: contiguous point ranges of the same grid
        IntegrationGrid._check_inputs(None, N, domain, True)  # values already validated above
        nodes, hs, n = grid_nodes(N, domain, self._grid_func, host_bounds=host_bounds)
        table = self._weight_table(n, dim, domain.dtype, domain.device)
        begin, end = tqdist.shard_range(total_points, rank, world)
        if fused:
            unit = fn.to_struct([0.0] * dim, [1.0] * dim, 1.0)  # nodes are already in domain coordinates
            total = ops.fused_nc(unit, nodes.detach().contiguous(), table, begin, end)[0]
        else:
            total, one_d = None, None
            for p0 in range(begin, end, chunk_rows):
                p1 = min(end, p0 + chunk_rows)
                pts = ops.nc_grid_points(nodes, p0, p1)
                vals, _ = self.evaluate_integrand(fn, pts)
                if one_d is None:  # the reference's 1-D rule (utils.py:235-277): [N] and [N, 1] values give a 0-dim result
                    _, one_d = _split_function_values(vals)
                part = ops.nc_contract_f64(vals, table, p0, p1)
                total = part if total is None else total + part
                del pts, vals
            if total is None:  # a rank without points still takes part in the all-reduce
                probe = fn(nodes[:, 0].detach().reshape(1, -1).clone())
                _, one_d = _split_function_values(probe)
                total = torch.zeros(probe.shape[1:], dtype=torch.float64, device=domain.device)
            if one_d:
                total = total.reshape(())
        if world > 1:
            total = ops.all_reduce_sum_autograd(total)
        self._nr_of_fevals = total_points
        return _to_working_dtype(total, domain.dtype) * self._scale(hs, domain)

    @staticmethod
    def _squeeze_1d(function_values, fn, *args):
        """The 1-D squeeze rule of the reference's decorator (utils.py:235-277) around `fn(values, *args)`."""
        from .utils import _split_function_values

        values, one_d = _split_function_values(function_values)
        result = fn(values, *args)
        return torch.squeeze(result) if one_d else result

    @expand_func_values_and_squeeze_integral
    def calculate_result(self, function_values, dim, n_per_dim, hs, integration_domain):
        """Apply the composite rule to values on the full grid (grid_integrator.py:57-91)."""
        table = self._weight_table(n_per_dim, dim, integration_domain.dtype, function_values.device)  # real, also for complex values
        return ops.nc_contract(function_values, table) * self._scale(hs, integration_domain)

    def calculate_grid(self, N, integration_domain, disable_integration_domain_check=False):
        """(points [n^dim, dim], h [dim], n) (grid_integrator.py:93-127)."""
        N = self._adjust_N(dim=integration_domain.shape[0], N=N)
        grid = IntegrationGrid(N, integration_domain, self._grid_func, disable_integration_domain_check)
        return grid.points, grid.h, grid._N

    def get_jit_compiled_integrate(self, dim, N=None, integration_domain=None, backend=None, capture_integrand=False):
        """`compiled_integrate(fn, integration_domain)` with everything but the two arguments fixed
        (grid_integrator.py:134-255).  As in the reference the integrand is evaluated eagerly on every call;
        `capture_integrand=True` (extension) captures the whole call -- node/point kernels, the integrand's torch ops,
        the contraction -- once per integrand as a CUDA graph and replays it (frozen-state rule: integration/compiled.py)."""
        if N is None:
            N = self._get_minimal_N(dim)
        domain0 = _setup_integration_domain(dim, integration_domain, backend)
        self._check_inputs(dim=dim, N=N, integration_domain=domain0)

        def run(fn, domain, _rng):
            return self.integrate(fn, dim, N, domain)

        return GraphedIntegrate(run, domain0, None, capture_integrand)
