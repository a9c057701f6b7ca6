"""ctypes binding of libtqb200.so (the C ABI declared in include/tqb200.h).

PyTorch is used for device memory, streams and autograd plumbing only; every arithmetic step of the hot
path is a kernel behind this boundary.  There is no CPU fallback: loading fails loudly when the shared
library is missing, and every op raises when handed a non-CUDA tensor.
"""
import ctypes
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TQB200_LIB") or os.path.join(_PKG, "libtqb200.so")  # TQB200_LIB: kernel-variant experiments

TQ_F32, TQ_F64 = 0, 1
TQ_MAX_DIM = 32

c_i32, c_i64, c_u32, c_u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
c_f64, c_sz, c_p = ctypes.c_double, ctypes.c_size_t, ctypes.c_void_p


class tq_integrand(ctypes.Structure):
    """Mirror of `struct tq_integrand` (include/tqb200.h)."""

    _fields_ = [
        ("family", c_i32), ("dim", c_i32), ("ncoeff", c_i32), ("_pad", c_i32),
        ("a", c_f64 * TQ_MAX_DIM), ("u", c_f64 * TQ_MAX_DIM), ("coeff", c_f64 * 8),
        ("start", c_f64 * TQ_MAX_DIM), ("size", c_f64 * TQ_MAX_DIM), ("scale", c_f64),
    ]


_P_INTEGRAND = ctypes.POINTER(tq_integrand)
TQ_VEGAS_MAX_PASSES = 128
TQ_EDGES_PAIRS, TQ_EDGES_RECORDS = 0, 1


class tq_vegas_state(ctypes.Structure):
    """Mirror of `struct tq_vegas_state`: caller-owned device buffers of a fused VEGAS run."""

    _fields_ = [
        ("x_edges", c_p), ("dx_edges", c_p), ("edges_packed", c_p), ("weights", c_p), ("counts", c_p), ("hist_pairs", c_p),
        ("jf2_rows", c_p), ("jf2_rows_cap", c_i64), ("sweep_dims_per_group", c_i32), ("_pad0", c_i32),
        ("dh", c_p), ("nh", c_p), ("offsets", c_p), ("JF", c_p), ("JF2", c_p), ("records", c_p), ("status", c_p),
        ("map_ws", c_p), ("map_ws_bytes", c_sz), ("ws", c_p), ("ws_bytes", c_sz), ("edges_layout", c_i32),
    ]


class tq_vegas_unfused_buffers(ctypes.Structure):
    """Mirror of `struct tq_vegas_unfused_buffers`: sample buffers of a callback-integrand VEGAS run."""

    _fields_ = [
        ("y", c_p), ("x", c_p), ("jac", c_p), ("jf", c_p), ("domain", c_p), ("warm_domain", c_p), ("cap_rows", c_i64),
        ("volume", c_f64),
    ]


# int allreduce(void* user, int64_t offset, int64_t count): sum comm[offset:offset+count] over the ranks, in place
tq_allreduce_callback = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64)


class tq_vegas_shard(ctypes.Structure):
    """Mirror of `struct tq_vegas_shard`: this rank's share of a multi-GPU fused VEGAS run."""

    _fields_ = [
        ("rank", c_i32), ("world", c_i32), ("cube_block_log2", c_i32), ("_pad", c_i32), ("n_cubes_local", c_i64),
        ("comm", c_p), ("allreduce", tq_allreduce_callback), ("user", c_p),
    ]


# int eval(void* user, int64_t rows, const void** f)
tq_eval_callback = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p))


class tq_vegas_result(ctypes.Structure):
    """Mirror of `struct tq_vegas_result` (host memory)."""

    _fields_ = [
        ("it", c_i32), ("n_block", c_i32), ("fevals", c_i64), ("starting_N", c_i64), ("calls_used", c_i32),
        ("n_passes", c_i32), ("results", c_f64 * 8), ("sigma2", c_f64 * 8), ("status", c_i32 * (TQ_VEGAS_MAX_PASSES * 4)),
    ]

# name -> (restype, argtypes); must list every symbol of include/tqb200.h (checked by tests/test_abi.py)
PROTOTYPES = {
    "tq_last_error": (ctypes.c_char_p, []),
    "tq_version": (ctypes.c_int, []),
    "tq_kernel_launches": (ctypes.c_uint64, []),
    "tq_workspace_bytes": (c_sz, []),
    "tq_device_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)] * 3),
    "tq_philox_uniform": (ctypes.c_int, [c_p, c_i64, c_i64, c_i32, c_i32, c_u64, c_u32, c_p]),
    "tq_mc_sample": (ctypes.c_int, [c_p, c_p, c_i64, c_i64, c_i32, c_i32, c_u64, c_u32, c_p]),
    "tq_mc_sample_replayable": (ctypes.c_int, [c_p, c_p, c_i64, c_i64, c_i32, c_i32, c_u64, c_u32, c_p, c_p]),
    "tq_mc_sample_backward": (ctypes.c_int, [c_p, c_i64, c_i64, c_i32, c_i32, c_u64, c_u32, c_p, c_p, c_sz, c_p]),
    "tq_sum_columns": (ctypes.c_int, [c_p, c_i64, c_i64, c_i32, c_p, c_p, c_p, c_sz, c_p]),
    "tq_vegas_map_forward": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_i64, c_i32, c_p]),
    "tq_vegas_map_forward_packed": (ctypes.c_int, [c_p, c_p, c_i32, c_p, c_p, c_p, c_p, c_i64, c_i32, c_i64, c_i32, c_p]),
    "tq_vegas_accumulate_fused": (ctypes.c_int, [c_p, c_p, c_p, c_f64, c_p, c_p, c_p, c_p, c_i64, c_i32, c_i64, c_i32, c_p]),
    "tq_vegas_sample_map": (ctypes.c_int, [c_p, c_i64, c_i32, c_i32, c_i32, c_i64, c_i64, c_p, c_i32, c_i64, c_p, c_u64, c_u32, c_p, c_p, c_p]),
    "tq_vegas_accumulate_regen": (ctypes.c_int, [c_p, c_i64, c_i32, c_i32, c_i32, c_i64, c_i64, c_i64, c_p, c_p, c_f64, c_p, c_p, c_p, c_p,
                                                 c_p, c_p, c_u64, c_u32, c_p]),
    "tq_vegas_map_accumulate": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i32, c_i64, c_i32, c_p]),
    "tq_vegas_map_workspace_bytes": (c_sz, [c_i32, c_i64, c_i32]),
    "tq_vegas_map_smooth": (ctypes.c_int, [c_p, c_p, c_p, c_i32, c_i64, c_f64, c_i32, c_p, c_p, c_sz, c_p]),
    "tq_vegas_map_update": (ctypes.c_int, [c_p, c_p, c_p, c_p, c_p, c_i32, c_i64, c_f64, c_i32, c_p, c_p, c_sz, c_p]),
    "tq_vegas_strat_nh": (ctypes.c_int, [c_p, c_i64, c_f64, c_i32, c_p, c_p, c_p, c_sz, c_p]),
    "tq_vegas_strat_offsets": (ctypes.c_int, [c_p, c_i64, c_p, c_p, c_sz, c_p]),
    "tq_vegas_strat_sample": (ctypes.c_int, [c_p, c_i64, c_i32, c_i32, c_i32, c_p, c_u64, c_u32, c_i64, c_i64, c_p, c_p]),
    "tq_vegas_strat_accumulate": (ctypes.c_int, [c_p, c_i64, c_p, c_i64, c_i64, c_p, c_p, c_i32, c_p]),
    "tq_vegas_strat_accumulate_backward": (ctypes.c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_p, c_i32, c_p]),
    "tq_vegas_strat_update": (ctypes.c_int, [c_p, c_p, c_p, c_i64, c_f64, c_f64, c_i32, c_p, c_p, c_p, c_sz, c_p]),
    "tq_nc_grid_points": (ctypes.c_int, [c_p, c_i32, c_i32, c_i64, c_i64, c_p, c_i32, c_p]),
    "tq_nc_grid_points_backward": (ctypes.c_int, [c_p, c_i32, c_i32, c_i64, c_i64, c_p, c_i32, c_p]),
    "tq_nc_contract": (ctypes.c_int, [c_p, c_p, c_i32, c_i32, c_i64, c_i64, c_i64, c_i32, c_p, c_p, c_sz, c_p]),
    "tq_nc_point_weights": (ctypes.c_int, [c_p, c_i32, c_i32, c_i64, c_i64, c_p, c_i32, c_p]),
    "tq_fused_mc": (ctypes.c_int, [_P_INTEGRAND, c_i32, c_i64, c_i64, c_u64, c_u32, c_p, c_p, c_sz, c_p]),
    "tq_fused_nc": (ctypes.c_int, [_P_INTEGRAND, c_p, c_p, c_i32, c_i32, c_i64, c_i64, c_p, c_p, c_sz, c_p]),
    "tq_vegas_map_pack_edges": (ctypes.c_int, [c_p, c_p, c_p, c_i32, c_i64, c_i32, c_p]),
    "tq_fused_vegas": (ctypes.c_int, [_P_INTEGRAND, c_i32, c_p, c_i64, c_i32, c_i64, c_i64, c_p, c_i32, c_i64, c_p, c_p, c_p,
                                      c_p, c_p, c_u64, c_u32, c_p, c_p, c_sz, c_p]),
    "tq_fused_vegas_sharded": (ctypes.c_int, [_P_INTEGRAND, c_i32, c_p, c_i64, c_i32, c_i64, c_i64, c_p, c_i32, c_i64, c_p, c_p,
                                              c_p, c_p, c_p, c_u64, c_u32, c_i32, c_i32, c_i32, c_p, c_p, c_sz, c_p]),
    "tq_vegas_map_unpack_hist": (ctypes.c_int, [c_p, c_p, c_p, c_i32, c_i64, c_i32, c_p]),
    "tq_fused_vegas_deferred": (ctypes.c_int, [_P_INTEGRAND, c_i32, c_p, c_i64, c_i32, c_i64, c_i64, c_p, c_i64, c_p, c_p, c_p, c_u64,
                                               c_u32, c_p, c_sz, c_p]),
    "tq_vegas_hist_sweep": (ctypes.c_int, [c_p, c_i64, c_i32, c_i32, c_i32, c_p, c_i64, c_p, c_i32, c_u64, c_u32, c_p, c_sz, c_p]),
    "tq_vegas_run_fused_sharded": (ctypes.c_int, [_P_INTEGRAND, c_i32, c_i64, c_i32, c_f64, c_f64, c_i32, c_i32, c_i64, c_i32, c_i64,
                                                  c_f64, c_f64, c_f64, c_u64, c_u32, ctypes.POINTER(tq_vegas_state),
                                                  ctypes.POINTER(tq_vegas_shard), ctypes.POINTER(tq_vegas_result), c_p]),
    "tq_vegas_map_records_bytes": (c_sz, [c_i32, c_i64, c_i32]),
    "tq_vegas_map_pack_records": (ctypes.c_int, [c_p, c_p, c_p, c_i32, c_i64, c_i32, c_p]),
    "tq_vegas_map_unpack_records": (ctypes.c_int, [c_p, c_p, c_p, c_i32, c_i64, c_i32, c_p]),
    "tq_vegas_run_unfused": (ctypes.c_int, [tq_eval_callback, c_p, c_i32, c_i32, c_i64, c_i32, c_f64, c_f64, c_i32, c_i32, c_i64,
                                            c_i32, c_i64, c_f64, c_f64, c_f64, c_u64, c_u32, ctypes.POINTER(tq_vegas_state),
                                            ctypes.POINTER(tq_vegas_unfused_buffers), ctypes.POINTER(tq_vegas_result), c_p]),
    "tq_vegas_schedule": (ctypes.c_int, [c_p, c_p, c_i32, c_i32, c_f64, c_f64, c_i64, c_i64, c_i32, c_i32, c_i64, c_p, c_p, c_p]),
    "tq_vegas_run_fused": (ctypes.c_int, [_P_INTEGRAND, c_i32, c_i64, c_i32, c_f64, c_f64, c_i32, c_i32, c_i64, c_i32, c_i64,
                                          c_f64, c_f64, c_f64, c_u64, c_u32, ctypes.POINTER(tq_vegas_state),
                                          ctypes.POINTER(tq_vegas_result), c_p]),
    "tq_l2_fetch_granularity": (ctypes.c_int, [c_i32, ctypes.POINTER(c_i32)]),
    "tq_peak_microbench": (ctypes.c_int, [c_i32, c_i64, c_p, ctypes.POINTER(c_f64), c_p]),
    "tq_red_microbench": (ctypes.c_int, [c_p, c_i64, c_i64, ctypes.POINTER(c_f64), c_p]),
}

_cdll = None
_fns = {}
_lock = threading.Lock()
_workspaces = {}
_MAX_WORKSPACES = 8
launch_count = 0  # kernels-launching C-ABI calls issued by this process (bench.py reports it)


def load():
    """Load the shared library (once).  Raises RuntimeError with build instructions when it is missing."""
    global _cdll
    if _cdll is not None:
        return _cdll
    with _lock:
        if _cdll is not None:
            return _cdll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"torchquad_b200: CUDA library {LIB_PATH} not found. Build it with "
                "`python -m torchquad_b200.build` (needs nvcc); there is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
            _fns[name] = fn
        _cdll = lib
    return _cdll


def dtype_code(dtype):
    if dtype == torch.float32:
        return TQ_F32
    if dtype == torch.float64:
        return TQ_F64
    raise ValueError(f"torchquad_b200 supports float32 and float64 only, got {dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "torchquad_b200 runs on CUDA only (no CPU fallback): got a tensor on "
                f"{t.device}. Create the integration domain on a CUDA device or call "
                "torchquad_b200.set_up_backend('torch') on a GPU machine."
            )


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise RuntimeError("torchquad_b200: internal error, non-contiguous tensor passed to the C ABI")
    return t.data_ptr()


def stream_ptr(device=None):
    """Raw cudaStream_t of torch's current stream on `device` (the C ABI launches on the caller's stream)."""
    idx = device.index if device is not None and device.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class on_device:
    """`with on_device(dev):` -- make `dev` the current CUDA device for the C-ABI call; free when it already is."""

    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def workspace(device):
    """Zero-initialised scratch buffer handed to every call that needs temporaries: one per (device, stream).

    The buffer holds the per-CTA partials and the ticket of the grid reductions, so two integrations that overlap on
    different streams (or a CUDA-graph replay overlapping eager work) must not share it; work on ONE stream is ordered
    and reuses its buffer.  Streams come and go: only the most recently used few buffers are kept."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch._C._cuda_getCurrentRawStream(idx))
    ws = _workspaces.pop(key, None)
    if ws is None:
        nbytes = load().tq_workspace_bytes()
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=torch.device("cuda", idx))
        while len(_workspaces) >= _MAX_WORKSPACES:
            _workspaces.pop(next(iter(_workspaces)))  # least recently used (dicts keep insertion order)
    _workspaces[key] = ws
    return ws


def call(name, *args):
    """Invoke a C-ABI entry point and turn a non-zero status into RuntimeError."""
    global launch_count
    fn = _fns.get(name)
    if fn is None:
        load()
        fn = _fns[name]
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {_cdll.tq_last_error().decode()}")
    launch_count += 1
    return rc


def kernel_launches():
    """Kernels launched by libtqb200 in this process so far (counted at every launch site of the library)."""
    return int(load().tq_kernel_launches())


def l2_fetch_granularity(device, nbytes=0):
    """Return the current L2 fetch granularity hint of `device`; set it first when nbytes > 0."""
    prev = c_i32()
    with on_device(device):
        call("tq_l2_fetch_granularity", nbytes, ctypes.byref(prev))
    return prev.value


def device_info():
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    call("tq_device_info", ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor))
    return sm.value, major.value, minor.value
