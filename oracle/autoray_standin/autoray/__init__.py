"""Minimal torch+numpy stand-in for the third-party `autoray` package.

TEST INFRASTRUCTURE ONLY.  The reference (esa/torchquad, /root/reference) reaches
all of its arithmetic through `autoray` (pyproject.toml:28, autoray>=0.5.0, not
vendored, not installed in this image, no network).  This module implements just
the dispatch surface the reference uses so that the UNMODIFIED reference can be
imported in the build container to (a) pin the oracle (`oracle/ref_oracle.py`)
and (b) generate the golden fixtures under `tests/golden/`
(`oracle/make_golden.py`).  It is never imported by the product package
`torchquad_b200`, and it does not travel in any code path on the GPU box.

Semantics follow autoray's documented behaviour: `autoray.numpy.<fn>(*args,
like=...)` dispatches on `like` (a backend name or an array) or else on the
first array argument, translating numpy-style names to the backend's.
"""
import importlib
import sys
import types

import numpy as _np

_registered = {}  # (backend, name) -> callable


def infer_backend(x):
    if isinstance(x, str):
        return x
    mod = type(x).__module__.split(".")[0]
    if mod in ("torch", "numpy"):
        return mod
    return "builtins"


def register_function(backend, name, fn, wrap=False):
    _registered[(backend, name)] = fn


def _torch():
    return importlib.import_module("torch")


def _torch_dtype(d):
    torch = _torch()
    if d is None or isinstance(d, torch.dtype):
        return d
    if isinstance(d, str):
        return getattr(torch, d)
    return getattr(torch, _np.dtype(d).name)


def to_backend_dtype(dtype_name, like):
    backend = infer_backend(like)
    if backend == "torch":
        return _torch_dtype(dtype_name)
    return _np.dtype(dtype_name).type if backend == "numpy" else dtype_name


def get_dtype_name(x):
    d = x.dtype if hasattr(x, "dtype") else x
    s = str(d)
    if s.startswith("torch."):
        return s[len("torch."):]
    try:
        return _np.dtype(d).name
    except TypeError:
        return s


def astype(x, dtype_name, **kwargs):
    backend = infer_backend(x)
    if backend == "torch":
        return x.to(_torch_dtype(dtype_name))
    return _np.asarray(x).astype(dtype_name if isinstance(dtype_name, str) else dtype_name)


def to_numpy(x):
    if infer_backend(x) == "torch":
        return x.detach().cpu().numpy()
    return _np.asarray(x)


# ---- torch name translation -------------------------------------------------
def _t_array(x, dtype=None, **kw):
    torch = _torch()
    if isinstance(x, torch.Tensor):
        y = x.clone()
        return y.to(_torch_dtype(dtype)) if dtype is not None else y
    return torch.tensor(x, dtype=_torch_dtype(dtype), **kw)


def _t_asarray(x, dtype=None, **kw):
    return _torch().as_tensor(x, dtype=_torch_dtype(dtype), **kw)


def _t_axis_to_dim(name):
    def fn(xs, axis=0, **kw):
        if "dim" in kw:
            axis = kw.pop("dim")
        return getattr(_torch(), name)(xs, dim=axis, **kw)

    return fn


def _t_creation(name):
    def fn(*a, dtype=None, **kw):
        if dtype is not None:
            kw["dtype"] = _torch_dtype(dtype)
        return getattr(_torch(), name)(*a, **kw)

    return fn


def _t_clip(x, a_min=None, a_max=None):
    return _torch().clamp(x, min=a_min, max=a_max)


def _t_meshgrid(*xs, indexing="xy"):
    return _torch().meshgrid(*xs, indexing=indexing)


def _t_array_equal(a, b):
    return _torch().equal(a, b)


_TORCH_TABLE = {
    "array": _t_array,
    "asarray": _t_asarray,
    "max": lambda x, *a, **k: _torch().amax(x, *a, **k) if (a or k) else _torch().max(x),
    "min": lambda x, *a, **k: _torch().amin(x, *a, **k) if (a or k) else _torch().min(x),
    "clip": _t_clip,
    "concatenate": _t_axis_to_dim("cat"),
    "stack": _t_axis_to_dim("stack"),
    "meshgrid": _t_meshgrid,
    "array_equal": _t_array_equal,
    "zeros": _t_creation("zeros"),
    "ones": _t_creation("ones"),
    "arange": _t_creation("arange"),
    "linspace": _t_creation("linspace"),
    "ndarray": None,
}


def _np_creation(name):
    def fn(*a, dtype=None, **kw):
        if dtype is not None:
            kw["dtype"] = dtype if not str(dtype).startswith("torch.") else str(dtype)[6:]
        return getattr(_np, name)(*a, **kw)

    return fn


def _dispatch_backend(name, args, like):
    if like is not None:
        b = infer_backend(like)
        return "numpy" if b == "builtins" else b
    probe = args[0] if args else None
    if name in ("stack", "concatenate") and isinstance(probe, (list, tuple)) and probe:
        probe = probe[0]
    if name == "einsum" and len(args) > 1:
        probe = args[1]
    b = infer_backend(probe)
    return "numpy" if b == "builtins" else b


def do(name, *args, like=None, **kwargs):
    backend = _dispatch_backend(name, args, like)
    if (backend, name) in _registered:
        return _registered[(backend, name)](*args, **kwargs)
    if backend == "torch":
        torch = _torch()
        fn = _TORCH_TABLE.get(name)
        if fn is None:
            fn = getattr(torch, name)
        return fn(*args, **kwargs)
    if backend == "numpy":
        if name in ("zeros", "ones", "arange", "linspace", "array", "asarray"):
            return _np_creation(name)(*args, **kwargs)
        if name == "div":
            kwargs.pop("rounding_mode", None)
            return _np.floor_divide(*args, **kwargs)
        return getattr(_np, name)(*args, **kwargs)
    raise NotImplementedError(f"autoray stand-in: backend {backend!r} unsupported")


class _NumpyMimic(types.ModuleType):
    """`from autoray import numpy as anp` -> attribute access builds dispatchers."""

    def __init__(self, name, prefix=""):
        super().__init__(name)
        self._prefix = prefix

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        if item == "ndarray":
            return _np.ndarray
        if item == "random":
            return _RandomMimic()
        full = self._prefix + item

        def fn(*args, **kwargs):
            return do(full, *args, **kwargs)

        fn.__name__ = full
        return fn


class _RandomMimic:
    def seed(self, seed, like=None):
        if infer_backend(like) == "torch":
            _torch().manual_seed(seed)
        else:
            _np.random.seed(seed)

    def uniform(self, low=0.0, high=1.0, size=None, dtype=None, like=None):
        if infer_backend(like) == "torch":
            return _torch().rand(size, dtype=_torch_dtype(dtype)) * (high - low) + low
        return _np.random.uniform(low, high, size).astype(dtype or "float64")


numpy = _NumpyMimic("autoray.numpy")
sys.modules["autoray.numpy"] = numpy
