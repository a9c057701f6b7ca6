"""Small drivers for ncu captures: python scripts/profile_target.py {vegas8|vegas16cap|mc_sample|vegas4|unfused8}"""
import sys
import warnings

sys.path.insert(0, ".")
import torch

import torchquad_b200 as tq
from torchquad_b200 import integrands as F
from torchquad_b200 import ops

warnings.simplefilter("ignore")
dev = torch.device("cuda")
what = sys.argv[1]
if what == "vegas8":
    dom = torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=dev)
    r = tq.VEGAS().integrate(F.GenzOscillatory(8, a=0.5, u=0.3), 8, N=2_500_000_000, integration_domain=dom, seed=1)
    print(float(r))
elif what == "unfused8":
    dom = torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=dev)
    fn = F.GenzOscillatory(8, a=0.5, u=0.3)
    r = tq.VEGAS().integrate(lambda x: fn(x), 8, N=500_000_000, integration_domain=dom, seed=1)
    print(float(r))
elif what == "vegas4":
    dom = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=dev)
    for s in range(3):
        r = tq.VEGAS().integrate(F.GenzGaussian(4, a=5.0, u=0.5), 4, N=10**6, integration_domain=dom, seed=s)
    print(float(r))
elif what == "mc_sample":
    dom = torch.tensor([[0.0, 1.0]] * 10, dtype=torch.float32, device=dev)
    for _ in range(3):
        p = ops.mc_sample(dom, 2 * 10**8, 1, 0, 0)
    dom = torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=dev)
    for _ in range(3):
        p = ops.mc_sample(dom, 10**8, 1, 0, 0)
    torch.cuda.synchronize()
elif what == "vegas8cap":
    dom = torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=dev)
    v = tq.VEGAS(); v.max_map_intervals = 4096
    r = v.integrate(F.GenzOscillatory(8, a=0.5, u=0.3), 8, N=2_500_000_000, integration_domain=dom, seed=1)
    print(float(r))
elif what == "vegas16cap":
    dom = torch.tensor([[0.0, 1.0]] * 16, dtype=torch.float32, device=dev)
    v = tq.VEGAS(); v.max_map_intervals = 4096
    r = v.integrate(F.GenzProductPeak(16, a=2.0, u=0.5), 16, N=10**10, integration_domain=dom, seed=1)
    print(float(r), v._nr_of_fevals)
