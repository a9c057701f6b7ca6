"""Host-side helpers of the integrators (mirrors the torch branch of torchquad/integration/utils.py).

Only input handling lives here; arithmetic on sample data goes through the CUDA library."""
import warnings

import torch

from ..utils.set_log_level import logger
from ..utils.config import _get_default_backend, _get_precision  # noqa: F401  (_get_precision: API parity)


def _infer_backend(x):
    if isinstance(x, torch.Tensor):
        return "torch"
    mod = type(x).__module__.split(".")[0]
    return mod if mod in ("numpy", "jax", "jaxlib", "tensorflow") else "builtins"


def _require_torch_backend(backend):
    if backend not in (None, "torch"):
        raise ValueError(f'Unsupported numerical backend: {backend} (torchquad_b200 implements backend="torch" on CUDA)')


def _default_device():
    """Device for domains given as lists: torch's default device when it is CUDA, else the current GPU."""
    dev = torch.get_default_device() if hasattr(torch, "get_default_device") else torch.empty(0).device
    if dev.type == "cuda":
        return dev
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return dev  # CPU: host-side checks still work, the compute ops raise (no CPU fallback)


def _linspace_with_grads(start, stop, N, requires_grad):
    """Equally spaced 1-D grid that keeps gradients wrt start/stop (utils.py:24-58)."""
    if requires_grad or _is_compiling(start):
        # affine form: differentiable, and free of the host read-back of tensor bounds that torch.linspace makes
        # (which a CUDA-graph capture cannot contain)
        grid = torch.linspace(0.0, 1.0, N, dtype=start.dtype, device=start.device)
        return grid * (stop - start) + start
    return torch.linspace(start, stop, N, dtype=start.dtype, device=start.device)


def _add_at_indices(target, indices, source, is_sorted=False):
    """target[indices[i]] += source[i] in place (utils.py:61-107).

    Kept for API parity only: the VEGAS map and stratification of this package do their scatter work in
    tqb200 kernels (vegas_map.cu / vegas_strat.cu), not through this function."""
    if not isinstance(target, torch.Tensor):
        raise NotImplementedError(f"Unsupported numerical backend: {_infer_backend(target)}")
    target.scatter_add_(dim=0, index=indices, src=source)


def _setup_integration_domain(dim, integration_domain, backend):
    """Turn the user's domain into a [dim, 2] tensor (utils.py:110-159)."""
    _require_torch_backend(backend)
    if integration_domain is None:
        integration_domain = [[-1.0, 1.0]] * dim
    domain_backend = _infer_backend(integration_domain)
    convert = domain_backend == "builtins"
    if not convert and domain_backend != "torch":
        if backend is None:
            raise ValueError(f'Unsupported numerical backend: {domain_backend} (torchquad_b200 implements backend="torch" on CUDA)')
        msg = "integration_domain should be a list when the backend argument is set."
        logger.warning(msg)
        warnings.warn(msg, RuntimeWarning)
        integration_domain = [[float(b) for b in bounds] for bounds in integration_domain]
        convert = True
    if convert:
        if backend is None and _get_default_backend() != "torch":
            raise ValueError(f"Unsupported numerical backend: {_get_default_backend()}")
        rows = [[float(b) for b in bounds] for bounds in integration_domain]
        integration_domain = torch.tensor(rows, dtype=torch.get_default_dtype(), device=_default_device())
    if tuple(integration_domain.shape) != (dim, 2):
        raise ValueError(
            "The integration domain has an unexpected shape. "
            f"Expected {(dim, 2)}, got {tuple(integration_domain.shape)}"
        )
    return integration_domain


def _check_integration_domain(integration_domain, check_values=True):
    """Validate the domain and return its dimensionality (utils.py:162-206).  `check_values=False` skips the
    device reduction + read-back of the bounds check for callers that inspect a host copy themselves."""
    if _infer_backend(integration_domain) == "builtins":
        dim = len(integration_domain)
        if dim < 1:
            raise ValueError("len(integration_domain) needs to be 1 or larger.")
        for bounds in integration_domain:
            if len(bounds) != 2:
                raise ValueError(bounds, " in ", integration_domain, " does not specify a valid integration bound.")
            if bounds[0] > bounds[1]:
                raise ValueError(bounds, " in ", integration_domain, " does not specify a valid integration bound.")
        return dim
    if len(integration_domain.shape) != 2:
        raise ValueError("The integration_domain tensor has an invalid shape")
    dim, num_bounds = integration_domain.shape
    if dim < 1:
        raise ValueError("integration_domain.shape[0] needs to be 1 or larger.")
    if num_bounds != 2:
        raise ValueError("integration_domain must have 2 values per boundary")
    if not check_values or _is_compiling(integration_domain):
        return dim
    if bool((integration_domain[:, 1] - integration_domain[:, 0]).min() < 0.0):
        raise ValueError("integration_domain has invalid boundary values")
    return dim


def _split_function_values(function_values):
    """The 1-D squeeze rule of `expand_func_values_and_squeeze_integral` (utils.py:235-277):
    returns (values with an explicit integrand axis, squeeze_result)."""
    one_d = function_values.dim() == 1 or (function_values.dim() == 2 and function_values.shape[1] == 1)
    if one_d:
        warnings.warn("DEPRECATION WARNING: In future versions of torchquad, an array-like object will be returned.")
        if function_values.dim() == 1:
            function_values = function_values.unsqueeze(1)
    return function_values, one_d


def expand_func_values_and_squeeze_integral(f):
    """Decorator form of the rule above, for signature parity with the reference."""

    def wrap(*args, **kwargs):
        if len(args) > 1:
            values = args[1]
        elif "function_values" in kwargs:
            values = kwargs["function_values"]
        else:
            raise ValueError(
                "function_values argument not found in either positional or keyword arguments. "
                "Please provide function_values as the second positional argument or as a keyword argument."
            )
        values, one_d = _split_function_values(values)
        if len(args) > 1:
            args = (args[0], values, *args[2:])
        else:
            kwargs["function_values"] = values
        result = f(*args, **kwargs)
        return torch.squeeze(result) if one_d else result

    return wrap


_capture_probe = False  # set by integration/compiled.py while it dry-runs a call before capturing it


def _is_compiling(x):
    """True while torch.jit is tracing (utils.py:280-308) or a CUDA graph is being captured: value-dependent
    host checks (which would need a read-back) are skipped, as the reference does under tracing."""
    if not isinstance(x, torch.Tensor):
        return False
    if torch.jit.is_tracing() or _capture_probe:
        return True
    return x.is_cuda and torch.cuda.is_current_stream_capturing()


def _to_working_dtype(total, dtype):
    """fp64-accumulated sums -> the domain's precision (complex sums keep their imaginary part)."""
    if total.is_complex():
        return total.to(torch.complex64 if dtype == torch.float32 else torch.complex128)
    return total.to(dtype)
