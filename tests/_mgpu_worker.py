"""Worker of tests/test_gpu_multi.py (one rank per GPU, NCCL): every integrator run sharded over the ranks
must reproduce the single-GPU run on the same seed, because the sample set is identical by construction."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import torchquad_b200 as tq
from torchquad_b200 import integrands as F

warnings.simplefilter("ignore")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)


def both(make, fn, dim, kw, dt):
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
    tq.distributed.disable()
    single = make().integrate(fn, dim, integration_domain=dom, **kw)
    tq.distributed.enable()
    integ = make()
    sharded = integ.integrate(fn, dim, integration_domain=dom, **kw)
    tq.distributed.disable()
    return float(single), float(sharded), integ


checks = []
g = F.GenzGaussian(4, a=4.0, u=0.45)
for dt, tol in [(torch.float64, 1e-9), (torch.float32, 2e-4)]:
    for label, fn in [("fused", g), ("unfused", lambda x: g(x))]:
        a, b, integ = both(tq.MonteCarlo, fn, 4, dict(N=1_000_003, seed=3), dt)
        checks.append((f"MC {label} {dt}", a, b, tol))
        a, b, integ = both(tq.VEGAS, fn, 4, dict(N=400_000, seed=3), dt)
        checks.append((f"VEGAS {label} {dt}", a, b, tol * (1 if dt == torch.float64 else 20)))
        a, b, integ = both(tq.Boole, fn, 4, dict(N=21**4), dt)
        checks.append((f"Boole {label} {dt}", a, b, tol))
# large-map record layout ({x, dx, weight, count} per bin), forced on this small problem: the histogram is
# unpacked into weights/counts before the all-reduce, so sharded == single still holds
from torchquad_b200.integration.vegas_map import VEGASMap

default_threshold = VEGASMap.records_min_bytes
VEGASMap.records_min_bytes = 0
a, b, integ = both(tq.VEGAS, g, 4, dict(N=400_000, seed=3), torch.float64)
checks.append(("VEGAS fused records float64", a, b, 1e-9))
assert integ.map._records is not None
VEGASMap.records_min_bytes = default_threshold
ok = True
for name, a, b, tol in checks:
    good = abs(a - b) <= tol * abs(a)
    ok &= good
    if rank == 0:
        print(f"{'OK ' if good else 'BAD'} {name:45s} single={a:.12e} sharded={b:.12e} rel={abs(a-b)/abs(a):.2e}")
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("MGPU-OK" if float(flag) == 1.0 else "MGPU-FAIL")
sys.exit(0 if float(flag) == 1.0 else 1)
