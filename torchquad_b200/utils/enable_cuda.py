"""Module-path parity with torchquad/utils/enable_cuda.py (implementation in config.py)."""
from .config import enable_cuda  # noqa: F401
