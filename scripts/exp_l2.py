import sys, time, warnings
sys.path.insert(0, ".")
import torch
import torchquad_b200 as tq
from torchquad_b200 import integrands as F, _lib
warnings.simplefilter("ignore")
dev = torch.device("cuda")
print("default L2 fetch granularity:", _lib.l2_fetch_granularity(dev))
dom = torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=dev)
fn = F.GenzOscillatory(8, a=0.5, u=0.3)
def run(label, **attrs):
    v = tq.VEGAS()
    for k, a in attrs.items():
        setattr(v, k, a)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = v.integrate(fn, 8, N=2_500_000_000, integration_domain=dom, seed=rep)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{label:40s} {dt*1e3:9.1f} ms  {v._nr_of_fevals/dt:.3e} evals/s  result {float(r):.6f} +- {float(v._get_error()):.1e} (exact {fn.exact():.6f}) Ni={v.map.N_intervals}")
run("reference map size, L2 fetch default", l2_fetch_bytes=None)
run("reference map size, L2 fetch 64", l2_fetch_bytes=64)
run("reference map size, L2 fetch 32", l2_fetch_bytes=32)
for cap in (1 << 20, 1 << 16, 4096, 1024):
    run(f"map capped at {cap}", max_map_intervals=cap)
print("L2 fetch granularity after:", _lib.l2_fetch_granularity(dev))
