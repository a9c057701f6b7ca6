import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    warnings.filterwarnings("ignore", message="DEPRECATION WARNING: In future versions of torchquad")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def cuda():
    import torch

    return torch.device("cuda", 0)


_DELTAS = os.path.join(ROOT, "gpurun_out", "parity_deltas.txt")


@pytest.fixture(scope="session")
def record_delta():
    """record_delta(name, achieved, bound): print the ACHIEVED parity delta next to the bound the test asserts and append it
    to gpurun_out/parity_deltas.txt (copied to profiles/ per round), so tolerances are reported as measured values."""
    lines = []

    def rec(name, achieved, bound):
        line = f"{name:78s} achieved {achieved:.3e}   bound {bound:.1e}"
        print("PARITY " + line)
        lines.append(line)
        return achieved

    yield rec
    try:
        os.makedirs(os.path.dirname(_DELTAS), exist_ok=True)
        with open(_DELTAS, "a") as f:
            f.write("\n".join(lines) + ("\n" if lines else ""))
    except OSError:
        pass
