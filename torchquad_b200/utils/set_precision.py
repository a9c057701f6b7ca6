"""Global precision switch (mirrors torchquad/utils/set_precision.py:18-89 for the torch backend)."""
import os
import sys

from .set_log_level import logger


def _get_precision(backend):
    return os.environ.get(f"TORCHQUAD_DTYPE_{backend.upper()}", None)


def set_precision(data_type="float32", backend="torch"):
    """Set torch's default floating point dtype; moves the default device to CUDA once CUDA is initialised."""
    data_type = {"float": "float32", "double": "float64"}.get(str(data_type).lower(), data_type)
    if data_type not in ("float32", "float64"):
        msg = f'Invalid data type "{data_type}". Only float32 and float64 are supported. Setting the data type to float32.'
        logger.error(msg)
        print(f"ERROR: {msg}", file=sys.stderr)
        data_type = "float32"
    if backend != "torch":
        msg = f"torchquad_b200 only drives the torch backend; cannot set the data type for backend {backend}"
        logger.error(msg)
        print(f"ERROR: {msg}", file=sys.stderr)
        return
    import torch

    torch.set_default_dtype(torch.float32 if data_type == "float32" else torch.float64)
    if torch.cuda.is_initialized():
        torch.set_default_device("cuda")
        logger.info(f"Setting Torch's default dtype to {data_type} and device to CUDA.")
    else:
        logger.info(f"Setting Torch's default dtype to {data_type} (CPU).")
