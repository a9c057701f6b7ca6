"""Import the UNMODIFIED reference (esa/torchquad 0.5.0) from the staged copy baseline/_ref/.

Used only by bench.py's reference arm / cpu_baseline leg and by tests; the product package never imports it.
The reference's one un-vendored dependency, autoray, is absent from this image: oracle/autoray_standin provides the
subset of it that the reference calls (with it the reference's own test-suite passes unchanged).
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
STANDIN = os.path.join(os.path.dirname(HERE), "oracle", "autoray_standin")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "torchquad", "__init__.py"))


def why_unavailable():
    return (f"{REF_ROOT}/torchquad is missing: run baseline/stage_ref.sh in the build container "
            "(python __graft_entry__.py does it when /root/reference exists)")


def import_reference(quiet=True):
    """Return the reference's top-level module `torchquad` (imported from baseline/_ref, never from site-packages)."""
    if not available():
        raise RuntimeError(why_unavailable())
    for p in (STANDIN, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    mod = sys.modules.get("torchquad")
    if mod is not None and not os.path.abspath(getattr(mod, "__file__", "")).startswith(REF_ROOT):
        raise RuntimeError(f"a different `torchquad` is already imported from {mod.__file__}")
    mod = importlib.import_module("torchquad")
    if quiet:
        mod.set_log_level("WARNING")
    return mod
