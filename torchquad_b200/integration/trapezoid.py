"""Module-path parity with torchquad/integration/trapezoid.py."""
from .newton_cotes import Trapezoid  # noqa: F401
