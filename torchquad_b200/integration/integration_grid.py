"""Tensor-product integration grid (replaces torchquad/integration/integration_grid.py).

The reference builds the [n^dim, dim] point list with meshgrid + ravel + stack (three full-size
temporaries, integration_grid.py:98-99); here the 1-D node vectors are expanded by `tq_nc_grid_points`
in one coalesced pass, optionally for a sub-range of points (chunking / multi-GPU slabs)."""
from time import perf_counter

import torch

from .. import ops
from .utils import _check_integration_domain, _is_compiling, _linspace_with_grads, _setup_integration_domain


def grid_func(integration_domain, N, requires_grad=False, backend=None):
    """Default 1-D node generator: N equally spaced points from a to b (integration_grid.py:13-16)."""
    a = integration_domain[0]
    b = integration_domain[1]
    return _linspace_with_grads(a, b, N, requires_grad=requires_grad)


grid_func._equally_spaced = True  # lets grid_nodes read all bounds back at once instead of two scalars per dim


def grid_nodes(N, integration_domain, grid_func=grid_func, host_bounds=None):
    """(nodes [dim, n], h [dim], n): per-dimension nodes and mesh widths (integration_grid.py:64-93).
    `host_bounds` = integration_domain.tolist() when the caller has already read it back."""
    dim = integration_domain.shape[0]
    n = int(N ** (1.0 / dim) + 1e-8)
    requires_grad = bool(getattr(integration_domain, "requires_grad", False))
    if getattr(grid_func, "_equally_spaced", False) and not requires_grad and not _is_compiling(integration_domain):
        # torch.linspace reads tensor bounds back one scalar at a time; one copy of the whole domain gives the
        # same numbers with a single synchronisation
        bounds = host_bounds if host_bounds is not None else integration_domain.detach().tolist()
        grid_1d = [torch.linspace(lo, hi, n, dtype=integration_domain.dtype, device=integration_domain.device)
                   for lo, hi in bounds]
    else:
        grid_1d = [grid_func(integration_domain[d], n, requires_grad=requires_grad, backend="torch") for d in range(dim)]
    nodes = torch.stack(grid_1d)
    h = nodes[:, 1] - nodes[:, 0]  # g[1] - g[0] per dimension (integration_grid.py:91), one kernel for all of them
    return nodes, h, n


class IntegrationGrid:
    """Grid of N points (n = floor(N^(1/dim)) per dimension) over a domain; dim 0 varies slowest."""

    points = None
    h = None
    _N = None
    _dim = None
    _runtime = None

    def __init__(self, N, integration_domain, grid_func=grid_func, disable_integration_domain_check=False):
        start = perf_counter()
        self._check_inputs(N, integration_domain, disable_integration_domain_check)
        if not isinstance(integration_domain, torch.Tensor):
            integration_domain = _setup_integration_domain(len(integration_domain), integration_domain, backend="torch")
        elif not integration_domain.is_floating_point():
            integration_domain = integration_domain.to(torch.float64)  # issue #180 of the reference
        self._dim = integration_domain.shape[0]
        nodes, self.h, self._N = grid_nodes(N, integration_domain, grid_func)
        self._nodes = nodes
        self.points = ops.nc_grid_points(nodes)
        self._runtime = perf_counter() - start

    def _check_inputs(self, N, integration_domain, disable_integration_domain_check):
        """ValueErrors of integration_grid.py:105-124."""
        if disable_integration_domain_check:
            dim = len(integration_domain)
        else:
            dim = _check_integration_domain(integration_domain)
        if N < 2:
            raise ValueError("N has to be > 1.")
        if N ** (1.0 / dim) < 2:
            raise ValueError("Cannot create a ", dim, "-dimensional grid with ", N, " points. Too few points per dimension.")
