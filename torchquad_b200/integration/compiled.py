"""CUDA-graph replay of whole integrate() calls: the B200 counterpart of the reference's
`get_jit_compiled_integrate` (monte_carlo.py:108-225, grid_integrator.py:134-255).

The reference traces sample/grid creation and the result step with torch.jit and evaluates the integrand
eagerly.  Here the complete call -- our sampling / grid kernels, the integrand's own torch ops, the reduction
or contraction -- is captured ONCE per integrand into a `torch.cuda.CUDAGraph` on static buffers and replayed:
a call then costs one domain copy, one graph launch and one result copy instead of a dozen launches plus the
Python between them.  This is what small-N repeated quadrature (parameter scans, inner loops of a fit) is
bound by.

Rules (the reference's tracing has the same ones): N, dim and the integrand object are fixed per graph; the
integrand must be capturable (torch ops on the GPU, no host read-backs); gradients do not flow through a
replay -- with a domain that requires grad, an integrand that cannot be captured, or a built-in fused
integrand (already a single launch) the call runs eagerly instead.
"""
import torch

from ..integrands import BuiltinIntegrand
from ..utils.set_log_level import logger
from . import utils
from .utils import _setup_integration_domain


class _Entry:
    __slots__ = ("graph", "domain", "out", "fn")


class GraphedIntegrate:
    """Callable `compiled_integrate(fn, integration_domain=None)` returned by get_jit_compiled_integrate.

    `run(fn, domain, rng)` performs one eager integrate() of the owning integrator with everything else fixed;
    `rng` (None for the deterministic grid rules) is the private generator of this compiled function."""

    def __init__(self, run, domain0, rng):
        self._run = run
        self._domain0 = domain0
        self._rng = rng
        self._entries = {}
        self._eager = set()  # ids of integrands that could not be captured
        self.replays = 0     # statistics: how many calls were served by a graph replay
        if rng is not None and domain0.is_cuda:
            rng._call_offset = torch.zeros(1, dtype=torch.int32, device=domain0.device)

    # ------------------------------------------------------------------------------------------------------
    def _capture(self, fn, domain):
        dev = domain.device
        entry = _Entry()
        entry.fn = fn  # keeps id(fn) alive for the cache key
        entry.domain = domain.detach().clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            self._step(fn, entry.domain)  # allocator warm-up + lazy initialisation outside the capture
            # Dry run with torch's synchronisation checker armed: an integrand that reads values back to the host
            # (or copies pageable host memory) is found HERE, where failing is harmless -- a capture that dies
            # half-way leaves torch's CUDA generator registered with a dead graph.
            previous = torch.cuda.get_sync_debug_mode()
            utils._capture_probe = True
            torch.cuda.set_sync_debug_mode("error")
            try:
                self._step(fn, entry.domain)
            finally:
                torch.cuda.set_sync_debug_mode(previous)
                utils._capture_probe = False
        torch.cuda.current_stream(dev).wait_stream(side)
        entry.graph = torch.cuda.CUDAGraph()
        try:
            with torch.no_grad(), torch.cuda.graph(entry.graph):
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

    # ------------------------------------------------------------------------------------------------------
    def __call__(self, fn, integration_domain=None):
        domain = self._domain0 if integration_domain is None else integration_domain
        if not isinstance(domain, torch.Tensor):
            domain = _setup_integration_domain(self._domain0.shape[0], domain, "torch")
        if tuple(domain.shape) != tuple(self._domain0.shape):
            raise ValueError("The integration domain has an unexpected shape. "
                             f"Expected {tuple(self._domain0.shape)}, got {tuple(domain.shape)}")
        key = id(fn)
        eager = (not domain.is_cuda or isinstance(fn, BuiltinIntegrand) or key in self._eager
                 or (torch.is_grad_enabled() and domain.requires_grad))
        if eager:
            return self._step(fn, domain)
        entry = self._entries.get(key)
        if entry is None:
            try:
                entry = self._capture(fn, domain.to(self._domain0.dtype))
            except Exception as exc:  # the integrand is not capturable: keep working, eagerly
                logger.warning(f"get_jit_compiled_integrate: CUDA-graph capture failed ({exc}); running eagerly")
                torch.cuda.synchronize(domain.device)
                self._eager.add(key)
                return self._step(fn, domain)
            self._entries[key] = entry  # capturing does not execute: the first result comes from a replay too
        entry.domain.copy_(domain, non_blocking=True)
        entry.graph.replay()
        self.replays += 1
        return entry.out.clone()
