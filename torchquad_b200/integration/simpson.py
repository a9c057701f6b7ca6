"""Module-path parity with torchquad/integration/simpson.py."""
from .newton_cotes import Simpson  # noqa: F401
