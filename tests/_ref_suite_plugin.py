"""pytest plugin for tests/test_gpu_reference_suite.py::test_reference_own_suite: makes `import torchquad` (and every
`torchquad.<sub>` module the reference's tests import) resolve to torchquad_b200, provides the autoray stand-in the
reference's tests import themselves, and makes CUDA the default device -- then the reference's UNMODIFIED test files
(baseline/_ref/tests) run against this package.  Loaded with `-p _ref_suite_plugin`; never imported by the product."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle", "autoray_standin")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import torchquad_b200  # noqa: E402

SUBMODULES = [
    "integration", "integration.base_integrator", "integration.boole", "integration.gaussian", "integration.grid_integrator",
    "integration.integration_grid", "integration.monte_carlo", "integration.newton_cotes", "integration.rng",
    "integration.simpson", "integration.trapezoid", "integration.utils", "integration.vegas", "integration.vegas_map",
    "integration.vegas_stratification", "utils", "utils.deployment_test", "utils.enable_cuda", "utils.set_log_level",
    "utils.set_precision", "utils.set_up_backend",
]
sys.modules["torchquad"] = torchquad_b200
for name in SUBMODULES:
    sys.modules["torchquad." + name] = importlib.import_module("torchquad_b200." + name)


def pytest_configure(config):
    if torch.cuda.is_available():
        torch.set_default_device("cuda")  # what torchquad.enable_cuda() does for a user of the torch backend
    for w in ("ignore::UserWarning", "ignore::RuntimeWarning", "ignore::DeprecationWarning"):
        config.addinivalue_line("filterwarnings", w)
