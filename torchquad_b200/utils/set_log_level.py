"""Log-level control (mirrors torchquad/utils/set_log_level.py:5-18; loguru when present, stdlib otherwise)."""
import logging
import sys

try:  # loguru is what the reference uses; keep its sink format when available
    from loguru import logger as _loguru
except Exception:  # pragma: no cover - loguru is installed in the target image
    _loguru = None

_std = logging.getLogger("torchquad_b200")


class _Logger:
    """Tiny facade so the package logs the same way with or without loguru."""

    def __getattr__(self, name):
        target = _loguru if _loguru is not None else _std
        return getattr(target, name)


logger = _Logger()


def set_log_level(log_level: str):
    """Set the log level ('TRACE','DEBUG','INFO','SUCCESS','WARNING','ERROR','CRITICAL')."""
    if _loguru is not None:
        _loguru.remove()
        _loguru.add(sys.stderr, level=log_level,
                    format="<green>{time:HH:mm:ss}</green>|TQ-<blue>{level}</blue>| <level>{message}</level>")
    else:
        _std.setLevel(getattr(logging, log_level, logging.WARNING))
