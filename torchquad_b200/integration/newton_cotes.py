"""Composite Newton-Cotes rules: Trapezoid, Simpson, Boole (replaces newton_cotes.py, trapezoid.py,
simpson.py and boole.py of torchquad/integration).  Each rule is its 1-D weight pattern; the
tensor-product contraction is shared (grid_integrator.py in this package)."""
import warnings

import torch

from .grid_integrator import GridIntegrator


class NewtonCotes(GridIntegrator):
    """Abstract parent of the composite Newton-Cotes integrators."""

    def __init__(self):
        super().__init__()


class Trapezoid(NewtonCotes):
    """Composite trapezoidal rule: h/2 * (1, 2, ..., 2, 1)  (trapezoid.py:28-37)."""

    _rule_denominator = 2.0

    def integrate(self, fn, dim, N=1000, integration_domain=None, backend=None):
        return super().integrate(fn, dim, N, integration_domain, backend)

    @staticmethod
    def _rule_weights_1d(n, dtype, device):
        w = torch.full((n,), 2.0, dtype=dtype, device=device)
        w[:1] = 1.0  # slices: scalar fills on the device (w[0] = ... would be a host-to-device copy)
        w[-1:] = 1.0
        return w

    @staticmethod
    def _apply_composite_rule(cur_dim_areas, dim, hs, domain=None):
        """Values already shaped [..., n, ..., n] -> integral; kept for callers of the reference's hook."""
        return _contract_shaped(Trapezoid, cur_dim_areas, dim, hs)


class Simpson(NewtonCotes):
    """Composite Simpson rule: h/3 * (1, 4, 2, 4, ..., 4, 1), odd n >= 3  (simpson.py:30-46)."""

    _rule_denominator = 3.0

    def integrate(self, fn, dim, N=None, integration_domain=None, backend=None):
        return super().integrate(fn, dim, N, integration_domain, backend)

    @staticmethod
    def _rule_weights_1d(n, dtype, device):
        w = torch.full((n,), 2.0, dtype=dtype, device=device)
        w[1::2] = 4.0
        w[:1] = 1.0
        w[-1:] = 1.0
        return w

    @staticmethod
    def _apply_composite_rule(cur_dim_areas, dim, hs, domain=None):
        return _contract_shaped(Simpson, cur_dim_areas, dim, hs)

    @staticmethod
    def _get_minimal_N(dim):
        return 3**dim

    @staticmethod
    def _adjust_N(dim, N):
        """Round n per dimension down to an odd number >= 3 (simpson.py:54-81)."""
        n = int(N ** (1.0 / dim) + 1e-8)
        if n < 3:
            warnings.warn("N per dimension cannot be lower than 3. N per dim will now be changed to 3.")
            return 3**dim
        if n % 2 != 1:
            warnings.warn(
                "N per dimension cannot be even due to necessary subdivisions. "
                f"N per dim will now be changed to the next lower integer, i.e. {n} -> {n - 1}."
            )
            return (n - 1) ** dim
        return N


class Boole(NewtonCotes):
    """Composite Boole rule: 2h/45 * (7, 32, 12, 32, 14, ..., 32, 7), n = 4k+1 >= 5  (boole.py:30-48)."""

    _rule_denominator = 22.5

    def integrate(self, fn, dim, N=None, integration_domain=None, backend=None):
        return super().integrate(fn, dim, N, integration_domain, backend)

    @staticmethod
    def _rule_weights_1d(n, dtype, device):
        w = torch.full((n,), 32.0, dtype=dtype, device=device)
        w[2::4] = 12.0
        w[0::4] = 14.0
        w[:1] = 7.0
        w[-1:] = 7.0
        return w

    @staticmethod
    def _apply_composite_rule(cur_dim_areas, dim, hs, domain=None):
        return _contract_shaped(Boole, cur_dim_areas, dim, hs)

    @staticmethod
    def _get_minimal_N(dim):
        return 5**dim

    @staticmethod
    def _adjust_N(dim, N):
        """Round n per dimension down to 4k+1 >= 5 (boole.py:56-84)."""
        n = int(N ** (1.0 / dim) + 1e-8)
        if n < 5:
            warnings.warn("N per dimension cannot be lower than 5. N per dim will now be changed to 5.")
            return 5**dim
        if (n - 1) % 4 != 0:
            new_n = n - ((n - 1) % 4)
            warnings.warn(
                "N per dimension must be N = 1 + 4n with n a positive integer due to necessary subdivisions. "
                f"N per dim will now be changed to the next lower N satisfying this, i.e. {n} -> {new_n}."
            )
            return new_n**dim
        return N


def _contract_shaped(rule_cls, shaped, dim, hs):
    """Contract values shaped [*integrand_shape, n, ..., n] (the reference hook's input layout)."""
    from .. import ops

    n = shaped.shape[-1]
    lead = shaped.shape[: shaped.dim() - dim]
    flat = shaped.reshape(*lead, n**dim) if lead else shaped.reshape(n**dim)
    values = flat.movedim(-1, 0).contiguous()
    integ = rule_cls()
    table = integ._weight_table(n, dim, shaped.real.dtype if shaped.is_complex() else shaped.dtype, shaped.device)
    return ops.nc_contract(values, table) * integ._scale(hs)
