"""`torchquad._deployment_test()` for this package: a quick self-check of an installation on a CUDA device.

Own implementation (the reference's utils/deployment_test.py checks a PyPI install across backends): every
integrator of the drop-in surface integrates functions with known integrals on the GPU, on both paths
(torch-callable and built-in fused integrands), in both precisions; the library's symbols are looked up first.
Returns True when everything passed; failures are logged, not raised."""
import math

import torch

from .set_log_level import logger


def _deployment_test():
    from .. import VEGAS, Boole, GaussLegendre, MonteCarlo, Simpson, Trapezoid, _lib, integrands

    ok = True

    def check(label, got, want, tol):
        nonlocal ok
        good = math.isfinite(got) and abs(got - want) <= tol * max(1.0, abs(want))
        ok &= good
        (logger.info if good else logger.error)(f"{'ok  ' if good else 'FAIL'} {label}: {got:.8g} (expected {want:.8g})")

    try:
        lib = _lib.load()
        logger.info(f"libtqb200 version {lib.tq_version()}, {len(_lib.PROTOTYPES)} entry points")
    except Exception as exc:  # noqa: BLE001
        logger.error(f"libtqb200.so could not be loaded: {exc}")
        return False
    if not torch.cuda.is_available():
        logger.error("no CUDA device: torchquad_b200 has no CPU path")
        return False
    dev = torch.device("cuda", torch.cuda.current_device())
    for dt in (torch.float32, torch.float64):
        tol = 2e-3 if dt == torch.float32 else 1e-3
        dom = torch.tensor([[0.0, 2.0], [-1.0, 1.0]], dtype=dt, device=dev)
        exact = 2.0 * (1.0 - math.cos(2.0)) + 0.0  # int sin(x) dx dy + int y dx dy over the domain
        fn = lambda x: torch.sin(x[:, 0]) + x[:, 1]  # noqa: E731
        for cls, kw in ((Trapezoid, dict(N=101**2)), (Simpson, dict(N=51**2)), (Boole, dict(N=49**2)),
                        (GaussLegendre, dict(N=12**2))):
            check(f"{cls.__name__} {dt}", float(cls().integrate(fn, 2, integration_domain=dom, **kw)), exact, tol)
        check(f"MonteCarlo {dt}", float(MonteCarlo().integrate(fn, 2, N=400_000, integration_domain=dom, seed=0)), exact, 2e-2)
        check(f"VEGAS {dt}", float(VEGAS().integrate(fn, 2, N=200_000, integration_domain=dom, seed=0)), exact, 2e-2)
        g = integrands.GenzGaussian(3, a=3.0, u=0.5)
        unit = torch.tensor([[0.0, 1.0]] * 3, dtype=dt, device=dev)
        check(f"fused VEGAS {dt}", float(VEGAS().integrate(g, 3, N=300_000, integration_domain=unit, seed=0)), g.exact(), 1e-2)
        check(f"fused MonteCarlo {dt}", float(MonteCarlo().integrate(g, 3, N=1_000_000, integration_domain=unit, seed=0)),
              g.exact(), 1e-2)
        check(f"fused Boole {dt}", float(Boole().integrate(g, 3, N=33**3, integration_domain=unit)), g.exact(), 1e-4)
    logger.info("deployment test passed" if ok else "deployment test FAILED")
    return ok
