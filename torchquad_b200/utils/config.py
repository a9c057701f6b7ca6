"""Global configuration of the drop-in: backend selection, CUDA enablement, default precision.

One module instead of the reference's three (`torchquad/utils/{set_up_backend,enable_cuda,set_precision}.py`);
the per-function modules of the same names re-export from here so `torchquad.utils.<name>` paths keep working.
Only the torch backend exists in this package, so everything that the reference does for numpy / jax /
tensorflow collapses to an error message.
"""
import os
import sys
import warnings

from .set_log_level import logger

_DTYPE_ALIASES = {"float": "float32", "double": "float64"}
_BACKEND_ENV = "TORCHQUAD_DEFAULT_BACKEND"


def _get_default_backend():
    """Backend passed to the latest `set_up_backend` call ("torch" when there was none)."""
    return os.environ.get(_BACKEND_ENV, "torch")


def _get_precision(backend):
    """Per-backend dtype override of the reference (`TORCHQUAD_DTYPE_<BACKEND>`); unused for torch."""
    return os.environ.get("TORCHQUAD_DTYPE_" + backend.upper())


def _complain(message):
    logger.error(message)
    print("ERROR: " + message, file=sys.stderr)


def set_precision(data_type="float32", backend="torch"):
    """Make `data_type` ("float32" / "float64", or the legacy "float" / "double") torch's default dtype.
    Once CUDA has been initialised the default device becomes CUDA as well, as in the reference."""
    import torch

    name = _DTYPE_ALIASES.get(str(data_type).lower(), data_type)
    if name not in ("float32", "float64"):
        _complain(f'Invalid data type "{name}". Only float32 and float64 are supported. Setting the data type to float32.')
        name = "float32"
    if backend != "torch":
        _complain(f"Changing the data type is not supported for backend {backend}: torchquad_b200 drives torch only")
        return
    torch.set_default_dtype(getattr(torch, name))
    on_gpu = torch.cuda.is_initialized()
    if on_gpu:
        torch.set_default_device("cuda")
    logger.info(f"Torch default dtype is now {name}" + (" on CUDA." if on_gpu else " (CPU default device)."))


def enable_cuda(data_type="float32"):
    """Initialise CUDA (so later `set_precision` calls move the default device) and optionally set the dtype."""
    import torch

    if not torch.cuda.is_available():
        message = "Error enabling CUDA. cuda.is_available() returned False. torchquad_b200 has no CPU path."
        logger.warning(message)
        warnings.warn(message, RuntimeWarning)
        return
    torch.cuda.init()
    logger.info(f"torch {torch.__version__}: {torch.cuda.device_count()} CUDA device(s), current GPU{torch.cuda.current_device()}")
    if data_type is not None:
        set_precision(data_type)


def set_up_backend(backend, data_type=None, torch_enable_cuda=True):
    """`set_up_backend("torch", "float64")`: enable CUDA, then set the precision once, then remember the backend."""
    if backend != "torch":
        raise ValueError(f'torchquad_b200 implements the backend="torch" CUDA path only, got backend={backend!r}')
    if torch_enable_cuda:
        # the reference passes data_type=None down when it will call set_precision itself right after
        enable_cuda(data_type="float32" if data_type is None else None)
    if data_type is not None:
        set_precision(data_type, backend=backend)
    os.environ[_BACKEND_ENV] = backend
