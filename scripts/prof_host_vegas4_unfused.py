"""cProfile of the host side of the BASELINE configs[0] run with a torch-callable integrand (unfused path)."""
import cProfile, pstats, sys, warnings, time
sys.path.insert(0, ".")
import torch
import torchquad_b200 as tq
from torchquad_b200 import integrands as F
warnings.simplefilter("ignore")
dev = torch.device("cuda")
dom = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=dev)
g = F.GenzGaussian(4, a=5.0, u=0.5)
fn = lambda x: g(x)
v = tq.VEGAS()
for s in range(5):
    v.integrate(fn, 4, N=10**6, integration_domain=dom, seed=s)
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(20):
    float(v.integrate(fn, 4, N=10**6, integration_domain=dom, seed=s))
print("ms/run", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for s in range(20):
    float(v.integrate(fn, 4, N=10**6, integration_domain=dom, seed=s))
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
