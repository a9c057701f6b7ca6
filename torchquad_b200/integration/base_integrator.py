"""Plumbing shared by all integrators (the role of torchquad/integration/base_integrator.py):
argument validation, the call into the user's integrand with the vectorisation check, the feval counter."""
import warnings

import torch

from .utils import _check_integration_domain

_NOT_VECTORISED = (
    "The passed function was given {given} points but only returned {got} value(s)."
    "Please ensure that your function is vectorized, i.e. can be called with multiple evaluation points at once. "
    "It should return a tensor where first dimension matches length of passed elements. "
)
_FOREIGN_RESULT = (
    "The passed function's return value has a different numerical backend than the passed points. Will try to "
    "convert. Note that this may be slow as it results in memory transfers between CPU and GPU, if torchquad "
    "uses the GPU."
)


def _as_device_tensor(values, like):
    """Integrand outputs that are not torch tensors (numpy arrays, lists) are converted with a warning."""
    if isinstance(values, torch.Tensor):
        return values
    warnings.warn(_FOREIGN_RESULT)
    return torch.as_tensor(values, device=like.device)


class BaseIntegrator:
    """Abstract parent of MonteCarlo, VEGAS and the grid integrators."""

    _fn = None                  # integrand
    _dim = None                 # dimensionality of its domain
    _integration_domain = None
    _nr_of_fevals = None        # integrand evaluations of the last `integrate` call

    def __init__(self):
        self._nr_of_fevals = 0

    def integrate(self):
        raise NotImplementedError("This is an abstract base class. Should not be called.")

    @staticmethod
    def evaluate_integrand(fn, points, weights=None, args=None):
        """Evaluate `fn(points, *args)` once for all points -> (values, number of points).

        Raises ValueError when the integrand is not vectorised (base_integrator.py:66-75); `weights`
        (one per point) are broadcast over any trailing integrand dimensions (:77-89)."""
        given = points.shape[0]
        values = _as_device_tensor(fn(points, *(args if args is not None else ())), points)
        got = values.shape[0] if values.dim() > 0 else 1
        if values.dim() == 0 or got != given:
            raise ValueError(_NOT_VECTORISED.format(given=given, got=got))
        if weights is not None:
            values = values * weights.reshape((given,) + (1,) * (values.dim() - 1))
        return values, given

    def _eval(self, points, weights=None, args=None):
        """`evaluate_integrand` on `self._fn`, adding to the feval counter."""
        values, n = self.evaluate_integrand(self._fn, points, weights=weights, args=args)
        self._nr_of_fevals += n
        return values

    @staticmethod
    def _check_inputs(dim=None, N=None, integration_domain=None):
        """ValueError for dim < 1, a non-int or non-positive N, or a domain that disagrees with dim."""
        if dim is not None and dim < 1:
            raise ValueError("Dimension needs to be 1 or larger.")
        if N is not None and not (type(N) is int and N >= 1):
            raise ValueError("N has to be a positive integer.")
        if integration_domain is None:
            return
        if dim is not None and _check_integration_domain(integration_domain) != dim:
            raise ValueError("The dimension of the integration domain must match the passed function dimensionality dim.")
