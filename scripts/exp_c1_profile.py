import sys, time, warnings, cProfile, pstats
sys.path.insert(0, ".")
import torch
import torchquad_b200 as tq
from torchquad_b200 import integrands as F
warnings.simplefilter("ignore")
dev = torch.device("cuda")
g4 = F.GenzGaussian(4, a=5.0, u=0.5)
dom = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=dev)
v = tq.VEGAS()
for s in range(5):
    v.integrate(g4, 4, N=10**6, integration_domain=dom, seed=s)
torch.cuda.synchronize()
t = time.perf_counter()
for s in range(20):
    r = v.integrate(g4, 4, N=10**6, integration_domain=dom, seed=s)
torch.cuda.synchronize()
print("fused ms/run", (time.perf_counter() - t) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for s in range(20):
    r = v.integrate(g4, 4, N=10**6, integration_domain=dom, seed=s)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
