"""Backend selection (mirrors torchquad/utils/set_up_backend.py:7-43; only "torch" is driven here)."""
import os

from .enable_cuda import enable_cuda
from .set_precision import set_precision


def _get_default_backend():
    return os.environ.get("TORCHQUAD_DEFAULT_BACKEND", "torch")


def set_up_backend(backend, data_type=None, torch_enable_cuda=True):
    """Configure the numerical backend; `backend` must be "torch" for this package."""
    if backend != "torch":
        raise ValueError(f'torchquad_b200 implements the backend="torch" CUDA path only, got backend={backend!r}')
    if torch_enable_cuda:
        if data_type is None:
            enable_cuda()
        else:
            enable_cuda(data_type=None)  # set_precision runs once, below
    if data_type is not None:
        set_precision(data_type, backend=backend)
    os.environ["TORCHQUAD_DEFAULT_BACKEND"] = backend
