"""Module-path parity with torchquad/integration/boole.py."""
from .newton_cotes import Boole  # noqa: F401
