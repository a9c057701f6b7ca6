"""Module-path parity with torchquad/utils/set_precision.py (implementation in config.py)."""
from .config import _get_precision, set_precision  # noqa: F401
