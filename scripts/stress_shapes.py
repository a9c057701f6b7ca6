"""Unusual-shape sweep of the public API (prints FAIL lines; exit code 1 when any)."""
import itertools, sys, traceback, warnings, math
sys.path.insert(0, ".")
import torch
import torchquad_b200 as tq
from torchquad_b200 import integrands as F
warnings.simplefilter("ignore")
dev = torch.device("cuda")
fails = 0
def check(label, fn):
    global fails
    try:
        ok, msg = fn()
        if not ok:
            fails += 1; print("FAIL", label, msg)
    except Exception as e:
        fails += 1; print("FAIL", label, type(e).__name__, str(e)[:200])
for dt in (torch.float64, torch.float32):
    tol = 1e-2 if dt == torch.float32 else 1e-2
    for dim in (1, 2, 3, 5, 11, 20, 32):
        dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
        g = F.GenzGaussian(dim, a=1.5, u=0.4)
        ex = g.exact()
        for N in (1000, 50_000, 1_000_000):
            for fused in (True, False):
                fn = g if fused else (lambda x: g(x))
                def run_mc():
                    m = tq.MonteCarlo(); r = float(m.integrate(fn, dim, N=N, integration_domain=dom, seed=1))
                    return abs(r - ex) <= 8 * ex / math.sqrt(N) + 1e-12, f"{r} vs {ex}"
                check(f"MC {dt} dim={dim} N={N} fused={fused}", run_mc)
                for kw in ({}, {"use_warmup": False}, {"use_grid_improve": False}, {"max_iterations": 1}, {"max_iterations": 7, "eps_rel": 1e-3}):
                    def run_v():
                        v = tq.VEGAS(); r = float(v.integrate(fn, dim, N=N, integration_domain=dom, seed=1, **kw))
                        err = float(v._get_error())
                        return math.isfinite(r) and abs(r - ex) <= 8 * err + 0.3 * ex * (N <= 1000) + 0.02 * ex, f"{r} vs {ex} err {err} it {v.it}"
                    check(f"VEGAS {dt} dim={dim} N={N} fused={fused} {kw}", run_v)
    for dim in (1, 2, 3, 5):
        dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
        g = F.GenzGaussian(dim, a=1.5, u=0.4)
        ex = g.exact()
        for cls in (tq.Trapezoid, tq.Simpson, tq.Boole, tq.GaussLegendre):
            for n in (5, 9, 33):
                for fused in (True, False):
                    fn = g if fused else (lambda x: g(x))
                    def run_nc():
                        r = float(cls().integrate(fn, dim, N=n**dim, integration_domain=dom))
                        loose = cls is tq.Trapezoid or n == 5
                        return abs(r - ex) <= (1e-1 if loose else 5e-3) * ex, f"{r} vs {ex}"
                    check(f"{cls.__name__} {dt} dim={dim} n={n} fused={fused}", run_nc)
print("stress done, fails:", fails)
sys.exit(1 if fails else 0)
