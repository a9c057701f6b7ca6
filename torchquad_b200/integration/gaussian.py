"""Gaussian quadrature on the GPU (replaces torchquad/integration/gaussian.py).

Same contraction kernels as the Newton-Cotes rules: the rule is its 1-D weight vector (numpy `leggauss` on the
host, cached like the reference does) and its nodes; the result is
    prod_d 0.5*(b_d - a_d) * sum_p f(p) * prod_d w[i_d(p)]
which the reference evaluates by multiplying the values with the meshgrid product of the weights
(`_weights`, gaussian.py:44-69) and summing axis by axis (`_apply_composite_rule`, :131-144).
"""
import numpy
import torch

from .. import ops
from .grid_integrator import GridIntegrator
from .utils import expand_func_values_and_squeeze_integral


class Gaussian(GridIntegrator):
    """Base class of the Gaussian rules; as in the reference the parent integrates on [-1, 1]^dim
    (`_resize_roots` is the identity) and subclasses rescale the roots."""

    def __init__(self):
        super().__init__()
        self.name = "Gauss-Legendre"
        self._root_fn = numpy.polynomial.legendre.leggauss
        self._root_args = ()
        self._cache = {}

    def integrate(self, fn, dim, N=8, integration_domain=None, backend=None):
        return super().integrate(fn, dim, N, integration_domain, backend)

    # ---- rule definition -------------------------------------------------------------------------
    def _cached_points_and_weights(self, N):
        """(roots, weights) of the n-point rule, cached per (n, *root_args) (gaussian.py:110-129)."""
        if not isinstance(N, int):
            if hasattr(N, "item"):
                N = N.item()
            else:
                raise NotImplementedError(f"N {N} is not an int and lacks an `item` method")
        key = (N, *self._root_args)
        if key not in self._cache:
            self._cache[key] = self._root_fn(*key)
        return self._cache[key]

    def _roots(self, N, backend="torch", requires_grad=False, dtype=None, device=None):
        roots = torch.as_tensor(self._cached_points_and_weights(N)[0], dtype=dtype or torch.float64, device=device)
        if requires_grad:
            roots.requires_grad = True
        return roots

    def _weights_1d(self, N, dtype, device):
        return torch.as_tensor(self._cached_points_and_weights(N)[1], dtype=dtype, device=device)

    def _weights(self, N, dim, backend="torch", requires_grad=False, dtype=None, device=None):
        """Product weights [N^dim] in grid order (gaussian.py:44-69); None-free variant of the NC hook."""
        from .utils import _default_device

        device = device if device is not None else _default_device()
        table = self._weights_1d(N, dtype or torch.float64, device).reshape(1, N).repeat(dim, 1).contiguous()
        return ops.nc_point_weights(table, 0, N**dim)

    def _resize_roots(self, integration_domain, roots):
        return roots

    @property
    def _grid_func(self):
        def f(integration_domain, N, requires_grad=False, backend=None):
            roots = self._roots(N, "torch", requires_grad, dtype=integration_domain.dtype, device=integration_domain.device)
            return self._resize_roots(integration_domain, roots)

        return f

    # ---- hooks of the shared contraction ---------------------------------------------------------
    def _weight_table(self, n, dim, dtype, device):
        return self._weights_1d(n, dtype, device).reshape(1, n).repeat(dim, 1).contiguous()

    def _scale(self, hs, domain=None):
        s = 0.5 * (domain[0][1] - domain[0][0])
        for d in range(1, domain.shape[0]):
            s = s * (0.5 * (domain[d][1] - domain[d][0]))
        return s

    def integrate_values(self, function_values, dim, n_per_dim, integration_domain):
        """Integral from raw (unweighted) values on the full grid."""
        table = self._weight_table(n_per_dim, dim, integration_domain.dtype, function_values.device)  # real, also for complex values
        return ops.nc_contract(function_values, table) * self._scale(None, integration_domain)

    @expand_func_values_and_squeeze_integral
    def calculate_result(self, function_values, dim, n_per_dim, hs, integration_domain):
        """Reference semantics (grid_integrator.py:57-91 + gaussian.py:131-144): `function_values` already carry
        the product weights (they are applied by `evaluate_integrand(..., weights=self._weights(...))`), so the
        composite rule is a plain sum times prod 0.5*(b-a)."""
        return ops.reduce_sum(function_values) * self._scale(hs, integration_domain)

    @staticmethod
    def _apply_composite_rule(cur_dim_areas, dim, hs, domain):
        for cur_dim in range(dim):
            cur_dim_areas = 0.5 * (domain[cur_dim][1] - domain[cur_dim][0]) * torch.sum(cur_dim_areas, dim=cur_dim_areas.dim() - 1)
        return cur_dim_areas


class GaussLegendre(Gaussian):
    """Gauss-Legendre quadrature on arbitrary boxes [a, b] (gaussian.py:147-162)."""

    def __init__(self):
        super().__init__()

    def _resize_roots(self, integration_domain, roots):
        a = integration_domain[0]
        b = integration_domain[1]
        return ((b - a) / 2) * roots + ((a + b) / 2)
