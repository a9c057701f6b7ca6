"""Common integrator plumbing (mirrors torchquad/integration/base_integrator.py)."""
import warnings

import torch

from .utils import _check_integration_domain


class BaseIntegrator:
    """Abstract base: input checks, integrand evaluation with the vectorisation check, feval counter."""

    _fn = None
    _dim = None
    _integration_domain = None
    _nr_of_fevals = None

    def __init__(self):
        self._nr_of_fevals = 0

    def integrate(self):
        raise NotImplementedError("This is an abstract base class. Should not be called.")

    def _eval(self, points, weights=None, args=None):
        result, num_points = self.evaluate_integrand(self._fn, points, weights=weights, args=args)
        self._nr_of_fevals += num_points
        return result

    @staticmethod
    def evaluate_integrand(fn, points, weights=None, args=None):
        """Call the user's integrand on `points` (base_integrator.py:42-91): returns (values, num_points)."""
        num_points = points.shape[0]
        result = fn(points, *(args or ()))
        if not isinstance(result, torch.Tensor):
            warnings.warn(
                "The passed function's return value has a different numerical backend than the passed points. "
                "Will try to convert. Note that this may be slow as it results in memory transfers between CPU "
                "and GPU, if torchquad uses the GPU."
            )
            result = torch.as_tensor(result, device=points.device)
        if result.dim() == 0 or result.shape[0] != num_points:
            num_results = 1 if result.dim() == 0 else result.shape[0]
            raise ValueError(
                f"The passed function was given {num_points} points but only returned {num_results} value(s)."
                f"Please ensure that your function is vectorized, i.e. can be called with multiple evaluation points at once. It should return a tensor "
                f"where first dimension matches length of passed elements. "
            )
        if weights is not None:
            result = result * weights.reshape([num_points] + [1] * (result.dim() - 1))
        return result, num_points

    @staticmethod
    def _check_inputs(dim=None, N=None, integration_domain=None):
        """ValueError on inconsistent dim / N / domain (base_integrator.py:93-116)."""
        if dim is not None and dim < 1:
            raise ValueError("Dimension needs to be 1 or larger.")
        if N is not None and (type(N) is not int or N < 1):
            raise ValueError("N has to be a positive integer.")
        if integration_domain is not None:
            dim_domain = _check_integration_domain(integration_domain)
            if dim is not None and dim != dim_domain:
                raise ValueError(
                    "The dimension of the integration domain must match the passed function dimensionality dim."
                )
