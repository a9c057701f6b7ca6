"""`get_jit_compiled_integrate` for the GPU (monte_carlo.py:108-225, grid_integrator.py:134-255 of the reference).

The reference traces its OWN steps (sample / grid creation, the result step) with torch.jit and evaluates the user's
integrand eagerly on every call.  The returned callable keeps exactly that contract by default:

  compiled = integrator.get_jit_compiled_integrate(dim, N, integration_domain, ...)
  result = compiled(fn, integration_domain)       # fn runs eagerly: it sees current Python state, autograd flows

All argument checking is done once, and the sampling / grid / reduction steps are single kernel launches already, so
there is nothing left to trace.  Monte Carlo draws fresh samples per call from a private generator.

`capture_integrand=True` (opt-in, an extension) additionally captures the WHOLE call -- our kernels AND the integrand's
torch ops -- into a CUDA graph per integrand object and replays it: a call then costs one domain copy, one graph launch
and one result copy, which is what small-N repeated quadrature (parameter scans, inner loops of a fit) is bound by
(MonteCarlo 3-D N=1e4: 177 -> 29 us per call on B200).  The price is the frozen-state rule of any graph capture:
  * everything `fn` reads from Python (globals, attributes, closure floats) is baked in at capture time -- pass changing
    values as CUDA tensors that are updated IN PLACE, or use the default mode;
  * a replay has no autograd graph.  Calls whose domain requires grad, and integrands whose values require grad (they
    close over parameters with requires_grad=True) are detected and run eagerly instead;
  * integrands that synchronise with the host cannot be captured and run eagerly (detected by a dry run);
  * graphs are cached per integrand object (bound methods: per instance + function) in a small LRU; code that passes
    a fresh lambda on every call stops capturing after a few misses and runs eagerly.
"""
import collections

import torch

from .. import _lib
from ..integrands import BuiltinIntegrand
from ..utils.set_log_level import logger
from . import utils
from .utils import _setup_integration_domain


class _Entry:
    __slots__ = ("graph", "domain", "out", "fn", "workspace")


def _fn_key(fn):
    """Cache key of an integrand: bound methods are re-created on every attribute access, so they are identified by
    (instance, function); everything else by object identity (the cache entry keeps `fn` alive, so ids are not reused)."""
    self_obj, func = getattr(fn, "__self__", None), getattr(fn, "__func__", None)
    if self_obj is not None and func is not None:
        return ("method", id(self_obj), id(func))
    return ("object", id(fn))


class GraphedIntegrate:
    """Callable `compiled_integrate(fn, integration_domain=None)` returned by get_jit_compiled_integrate.

    `run(fn, domain, rng)` performs one eager integrate() of the owning integrator with everything else fixed;
    `rng` (None for the deterministic grid rules) is the private generator of this compiled function."""

    max_graphs = 8          # LRU capacity (each graph pins its sample / value buffers)
    max_consecutive_misses = 4  # after this many captures in a row without a single reuse, stop capturing

    def __init__(self, run, domain0, rng, capture_integrand=False):
        self._run = run
        self._domain0 = domain0
        self._rng = rng
        self._capture_integrand = bool(capture_integrand)
        self._entries = collections.OrderedDict()  # key -> _Entry, least recently used first
        self._eager = collections.OrderedDict()    # key -> fn (kept alive so that the id stays valid)
        self._misses = 0
        self.replays = 0     # statistics: how many calls were served by a graph replay
        if rng is not None and domain0.is_cuda:
            # device-side call counter of the private generator: every step (eager or replayed) uses Philox call
            # `counter` and increments it, so eager calls and replays can be mixed freely without ever reusing a call
            rng._call_offset = torch.zeros(1, dtype=torch.int32, device=domain0.device)

    # ------------------------------------------------------------------------------------------------------
    def _capture(self, fn, domain):
        """Returns an _Entry, or None when the integrand's values carry an autograd graph (must stay eager)."""
        dev = domain.device
        entry = _Entry()
        entry.fn = fn  # keeps the key's ids alive
        entry.domain = domain.detach().clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            # allocator warm-up + lazy initialisation outside the capture, WITH autograd on: an integrand that closes
            # over parameters requiring grad shows up here and keeps its gradients (eager path)
            with torch.enable_grad():
                probe = self._step(fn, entry.domain)
            if isinstance(probe, torch.Tensor) and probe.requires_grad:
                torch.cuda.current_stream(dev).wait_stream(side)
                return None
            # Dry run with torch's synchronisation checker armed: an integrand that reads values back to the host
            # (or copies pageable host memory) is found HERE, where failing is harmless -- a capture that dies
            # half-way leaves torch's CUDA generator registered with a dead graph.
            previous = torch.cuda.get_sync_debug_mode()
            utils._capture_probe = True
            torch.cuda.set_sync_debug_mode("error")
            try:
                with torch.no_grad():
                    self._step(fn, entry.domain)
            finally:
                torch.cuda.set_sync_debug_mode(previous)
                utils._capture_probe = False
            # the library's scratch buffer is per stream: capture on the stream the warm-up ran on, and keep the buffer
            # alive with the graph (its address is baked into the captured launches)
            entry.workspace = _lib.workspace(dev)
        torch.cuda.current_stream(dev).wait_stream(side)
        entry.graph = torch.cuda.CUDAGraph()
        try:
            with torch.no_grad(), torch.cuda.graph(entry.graph, stream=side):
                entry.out = self._step(fn, entry.domain)
        except Exception:
            self._repair_after_failed_capture(dev)
            raise
        return entry

    @staticmethod
    def _repair_after_failed_capture(dev):
        """A capture that raised never reached the generator's epilogue; an empty capture runs it."""
        import warnings

        try:
            torch.cuda.synchronize(dev)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                with torch.cuda.graph(torch.cuda.CUDAGraph()):
                    pass
        except Exception:  # nothing more to do here; the caller falls back to eager execution
            pass

    def _step(self, fn, domain):
        out = self._run(fn, domain, self._rng)
        if self._rng is not None and self._rng._call_offset is not None:
            self._rng._call_offset.add_(1)  # next call (or replay) takes the next Philox call index
        return out

    def _remember_eager(self, key, fn):
        self._eager[key] = fn
        while len(self._eager) > 4 * self.max_graphs:
            self._eager.popitem(last=False)

    # ------------------------------------------------------------------------------------------------------
    def __call__(self, fn, integration_domain=None):
        domain = self._domain0 if integration_domain is None else integration_domain
        if not isinstance(domain, torch.Tensor):
            domain = _setup_integration_domain(self._domain0.shape[0], domain, "torch")
        if tuple(domain.shape) != tuple(self._domain0.shape):
            raise ValueError("The integration domain has an unexpected shape. "
                             f"Expected {tuple(self._domain0.shape)}, got {tuple(domain.shape)}")
        if (not self._capture_integrand or not domain.is_cuda or isinstance(fn, BuiltinIntegrand)
                or (torch.is_grad_enabled() and domain.requires_grad)):
            return self._step(fn, domain)
        key = _fn_key(fn)
        if key in self._eager:
            return self._step(fn, domain)
        entry = self._entries.get(key)
        if entry is None:
            if self._misses >= self.max_consecutive_misses:
                return self._step(fn, domain)  # a new integrand object on every call: capturing only costs time
            self._misses += 1
            try:
                entry = self._capture(fn, domain.to(self._domain0.dtype))
            except Exception as exc:  # the integrand is not capturable: keep working, eagerly
                logger.warning(f"get_jit_compiled_integrate: CUDA-graph capture failed ({exc}); running eagerly")
                torch.cuda.synchronize(domain.device)
                self._remember_eager(key, fn)
                return self._step(fn, domain)
            if entry is None:  # values require grad: eager, so that gradients reach the integrand's parameters
                self._remember_eager(key, fn)
                return self._step(fn, domain)
            self._entries[key] = entry  # capturing does not execute: the first result comes from a replay too
            while len(self._entries) > self.max_graphs:
                self._entries.popitem(last=False)  # frees the evicted graph and its private memory pool
        else:
            self._misses = 0
            self._entries.move_to_end(key)
        entry.domain.copy_(domain, non_blocking=True)
        entry.graph.replay()
        self.replays += 1
        return entry.out.clone()
