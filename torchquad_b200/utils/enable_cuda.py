"""CUDA switch (mirrors torchquad/utils/enable_cuda.py:7-26)."""
import warnings

from .set_log_level import logger
from .set_precision import set_precision


def enable_cuda(data_type="float32"):
    """Initialise CUDA and make it torch's default device; optionally set the default precision."""
    import torch

    if torch.cuda.is_available():
        torch.cuda.init()
        logger.info("PyTorch VERSION: " + str(torch.__version__))
        logger.info("Number of CUDA Devices: " + str(torch.cuda.device_count()))
        logger.info("Active CUDA Device: GPU" + str(torch.cuda.current_device()))
        if data_type is not None:
            set_precision(data_type)
    else:
        msg = "Error enabling CUDA. cuda.is_available() returned False. torchquad_b200 has no CPU path."
        logger.warning(msg)
        warnings.warn(msg, RuntimeWarning)
