import sys, warnings
sys.path.insert(0, ".")
import torch
import torchquad_b200 as tq
from torchquad_b200 import integrands as F
warnings.simplefilter("ignore")
dev = torch.device("cuda")
for dt in (torch.float32, torch.float64):
    dom = torch.tensor([[0.0, 1.0]] * 4, dtype=dt, device=dev)
    fn = F.GenzGaussian(4, a=5.0, u=0.5)
    for native in (True, True, False, False):
        v = tq.VEGAS(); v.native_loop = native
        r = v.integrate(fn, 4, N=10**6, integration_domain=dom, seed=3)
        print(dt, native, v.it, v._nr_of_fevals, "%.9e" % float(r), ["%.7e" % float(x) for x in v.results], "err %.3e" % float(v._get_error()),
              "edges", float(v.map.x_edges.double().sum()), "dh", float(v.strat.dh.double().sum()))
