"""Strong-scaling timing of the fused VEGAS run (launch with torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node N scripts/time_vegas_mgpu.py [N_total] [map_cap] [dim] [dtype]"""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import torchquad_b200 as tq
from torchquad_b200 import integrands as F

warnings.simplefilter("ignore")
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    tq.distributed.enable()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_500_000_000
cap = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "none" else None
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dt = getattr(torch, sys.argv[4]) if len(sys.argv) > 4 else torch.float64
fn = F.GenzOscillatory(dim, a=0.5, u=0.3) if dim == 8 else F.GenzProductPeak(dim, a=2.0, u=0.5)
dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
v = tq.VEGAS()
v.max_map_intervals = cap
for rep in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = v.integrate(fn, dim, N=N, integration_domain=dom, seed=rep)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} N={N:.2e} cap={cap} dim={dim} {dt}: {float(t)*1e3:9.2f} ms  {v._nr_of_fevals/float(t):.3e} evals/s  "
              f"it={v.it} fevals={v._nr_of_fevals} result={float(r):.9e} (exact {fn.exact():.9e}) err={float(v._get_error()):.2e}", flush=True)
if world > 1:
    dist.destroy_process_group()
