"""Wall-clock time of the BASELINE configs[0] run (VEGAS 4-D Genz Gaussian, N=1e6, fp64), fused path."""
import sys
import time
import warnings

sys.path.insert(0, ".")
import torch

import torchquad_b200 as tq
from torchquad_b200 import integrands as F

warnings.simplefilter("ignore")
dev = torch.device("cuda")
dom = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=dev)
fn = F.GenzGaussian(4, a=5.0, u=0.5)
for native in (True, False):
    v = tq.VEGAS()
    v.native_loop = native
    for s in range(5):
        r = v.integrate(fn, 4, N=10**6, integration_domain=dom, seed=s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    for s in range(n):
        r = v.integrate(fn, 4, N=10**6, integration_domain=dom, seed=s)
        float(r)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"native_loop={native}: {ms:.3f} ms/run, it={v.it}, fevals={v._nr_of_fevals}, result={float(r):.9f}")
