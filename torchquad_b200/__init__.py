"""torchquad_b200 -- B200-native implementation of torchquad's sampling-and-reduction hot path.

Drop-in for `torchquad.{MonteCarlo, VEGAS, Trapezoid, Simpson, Boole}().integrate(fn, dim, N,
integration_domain, backend="torch")` on CUDA devices: the Python classes keep the reference's surface
(esa/torchquad v0.5.0), every arithmetic step runs in hand-written sm_100a kernels behind the C ABI of
include/tqb200.h (libtqb200.so, loaded through ctypes).  There is no CPU path.
"""
__version__ = "0.1.0"

from . import distributed, integrands  # noqa: F401
from .integration.base_integrator import BaseIntegrator
from .integration.gaussian import Gaussian, GaussLegendre
from .integration.grid_integrator import GridIntegrator
from .integration.integration_grid import IntegrationGrid
from .integration.monte_carlo import MonteCarlo
from .integration.newton_cotes import Boole, NewtonCotes, Simpson, Trapezoid
from .integration.rng import RNG
from .integration.vegas import VEGAS
from .integration.vegas_map import VEGASMap
from .integration.vegas_stratification import VEGASStratification
from .utils.config import enable_cuda, set_precision, set_up_backend
from .utils.deployment_test import _deployment_test
from .utils.set_log_level import set_log_level

import os as _os

set_log_level(_os.environ.get("TORCHQUAD_LOG_LEVEL", "WARNING"))

__all__ = [
    "_deployment_test",
    "__version__", "GridIntegrator", "BaseIntegrator", "IntegrationGrid", "MonteCarlo", "Trapezoid", "Simpson",
    "Boole", "NewtonCotes", "GaussLegendre", "Gaussian", "VEGAS", "VEGASMap", "VEGASStratification", "RNG", "enable_cuda", "set_precision",
    "set_log_level", "set_up_backend", "integrands", "distributed",
]
