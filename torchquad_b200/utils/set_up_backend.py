"""Module-path parity with torchquad/utils/set_up_backend.py (implementation in config.py)."""
from .config import _get_default_backend, set_up_backend  # noqa: F401
