"""VEGAS 8-D fp64 at the reference's map size (Ni = 1e7 per dimension): record layout vs pair layout."""
import sys
import time
import warnings

sys.path.insert(0, ".")
import torch

import torchquad_b200 as tq
from torchquad_b200 import integrands as F
from torchquad_b200.integration.vegas_map import VEGASMap

warnings.simplefilter("ignore")
dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "vegas8"
if which == "vegas8":
    dim, N, dt, fn = 8, 2_500_000_000, torch.float64, F.GenzOscillatory(8, a=0.5, u=0.3)
else:
    dim, N, dt, fn = 16, 10**10, torch.float32, F.GenzProductPeak(16, a=2.0, u=0.5)
dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
default = VEGASMap.records_min_bytes
for label, thr in (("pairs", None), ("records", default), ("pairs", None), ("records", default)):
    VEGASMap.records_min_bytes = thr
    v = tq.VEGAS()
    if which != "vegas8":
        v.max_map_intervals = 1 << 22
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = v.integrate(fn, dim, N=N, integration_domain=dom, seed=1)
    float(r)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    print(f"{which} {label}: {el:.3f} s, {v._nr_of_fevals / el:.3e} evals/s, result {float(r):.9f} (exact {fn.exact():.9f}), it {v.it}, Ni {v.map.N_intervals}")
    del v
    torch.cuda.empty_cache()
