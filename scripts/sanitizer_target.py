"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, warnings
sys.path.insert(0, ".")
import torch
import torchquad_b200 as tq
from torchquad_b200 import integrands as F, ops
warnings.simplefilter("ignore")
dev = torch.device("cuda")
for dt in (torch.float32, torch.float64):
    dom = torch.tensor([[0.0, 1.0], [-1.0, 2.0], [0.5, 1.5]], dtype=dt, device=dev)
    fn = lambda x: torch.exp(-torch.sum((x - 0.4) ** 2, dim=1))
    g = F.GenzGaussian(3, a=[2.0, 1.0, 3.0], u=[0.4, 0.5, 0.6])
    print(dt, "MC", float(tq.MonteCarlo().integrate(fn, 3, N=5001, integration_domain=dom, seed=1)),
          float(tq.MonteCarlo().integrate(g, 3, N=5001, integration_domain=dom, seed=1)))
    print(dt, "MC vec", tq.MonteCarlo().integrate(lambda x: torch.stack([fn(x), fn(x) * 2, fn(x) ** 2], 1), 3, N=3333, integration_domain=dom, seed=1).tolist())
    for N in (2000, 30000):
        print(dt, "VEGAS", N, float(tq.VEGAS().integrate(fn, 3, N=N, integration_domain=dom, seed=2)),
              float(tq.VEGAS().integrate(g, 3, N=N, integration_domain=dom, seed=2)))
    v = tq.VEGAS(); v.max_map_intervals = 7
    print(dt, "VEGAS tiny map", float(v.integrate(g, 3, N=40000, integration_domain=dom, seed=3)))
    for cls, N in ((tq.Trapezoid, 1000), (tq.Simpson, 11**3), (tq.Boole, 9**3), (tq.GaussLegendre, 6**3)):
        print(dt, cls.__name__, float(cls().integrate(fn, 3, N, dom)), float(cls().integrate(g, 3, N, dom)))
    # record layout (large-map tables) forced on a small problem, both loops, fused and callback integrands
    from torchquad_b200.integration.vegas_map import VEGASMap
    keep = VEGASMap.records_min_bytes
    VEGASMap.records_min_bytes = 0
    for native in (True, False):
        v = tq.VEGAS(); v.native_loop = native
        print(dt, "VEGAS records", native, float(v.integrate(g, 3, N=30000, integration_domain=dom, seed=4)),
              float(v.integrate(fn, 3, N=30000, integration_domain=dom, seed=4)))
    # maps "beyond L2" with >= 2^20 rows per pass: deferred histogram + band sweeps (fused), jf^2 rows + sweeps (callback)
    if "--no-sweeps" not in sys.argv:
        v = tq.VEGAS(); v.native_loop = False
        print(dt, "VEGAS band sweeps", float(tq.VEGAS().integrate(F.GenzGaussian(2, a=3.0, u=0.5), 2, N=27_000_000, integration_domain=dom[:2], seed=4)),
              float(v.integrate(lambda x: torch.exp(-torch.sum((x - 0.4) ** 2, dim=1)), 2, N=27_000_000, integration_domain=dom[:2], seed=4)),
              v._regen_sweep)
    VEGASMap.records_min_bytes = keep
    v = tq.VEGAS(); v.native_loop = False
    print(dt, "VEGAS python loop", float(v.integrate(g, 3, N=30000, integration_domain=dom, seed=5)),
          float(v.integrate(fn, 3, N=30000, integration_domain=dom, seed=5)))
    # a map / stratification above the single-launch limits (multi-kernel update pipeline)
    v = tq.VEGAS()
    print(dt, "VEGAS 2-D large tables", float(v.integrate(F.GenzGaussian(2, a=3.0, u=0.5), 2, N=12_000_000,
                                                         integration_domain=dom[:2], seed=6)), v.map.N_intervals, v.strat.N_cubes)
    # vector-valued contraction, narrow and wide
    for cols in (3, 300):
        print(dt, "Simpson vec", cols, float(tq.Simpson().integrate(lambda x: fn(x)[:, None].expand(-1, cols) * 1.0, 3, N=9**3,
                                                                     integration_domain=dom).sum()))
    c = tq.Simpson().get_jit_compiled_integrate(dim=3, N=9**3, integration_domain=dom, capture_integrand=True)
    print(dt, "compiled", float(c(fn, dom)), float(c(fn, dom)))
    d1 = torch.tensor([[0.0, 1.0]], dtype=dt, device=dev, requires_grad=True)
    r = tq.VEGAS().integrate(lambda x: x[:, 0] ** 2, 1, N=5000, integration_domain=d1, seed=1); r.backward()
    r = tq.MonteCarlo().integrate(lambda x: x[:, 0] ** 2, 1, N=5000, integration_domain=d1, seed=1); r.backward()
    r = tq.Simpson().integrate(lambda x: x[:, 0] ** 2, 1, N=101, integration_domain=d1); r.backward()
    print(dt, "grads", d1.grad.tolist())
torch.cuda.synchronize()
print("SANITIZER-TARGET-DONE")
