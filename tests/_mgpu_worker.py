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
# small fused VEGAS runs are replicated by default (a pass of 1e4 samples is launch latency, not work): identical results
a, b, integ = both(tq.VEGAS, g, 4, dict(N=400_000, seed=3), torch.float64)
assert integ._replicated and integ._shard is None
checks.append(("VEGAS fused replicated (small N) float64", a, b, 1e-12))  # fp atomics: runs agree to rounding
tq.VEGAS.min_rows_per_rank = 0  # ... but the sharded path must also be right at this size
for dt, tol in [(torch.float64, 1e-9), (torch.float32, 2e-4)]:
    for label, fn in [("fused", g), ("unfused", lambda x: g(x))]:
        a, b, integ = both(tq.MonteCarlo, fn, 4, dict(N=1_000_003, seed=3), dt)
        checks.append((f"MC {label} {dt}", a, b, tol))
        a, b, integ = both(tq.VEGAS, fn, 4, dict(N=400_000, seed=3), dt)
        checks.append((f"VEGAS {label} {dt}", a, b, tol * (1 if dt == torch.float64 else 20)))
        a, b, integ = both(tq.Boole, fn, 4, dict(N=21**4), dt)
        checks.append((f"Boole {label} {dt}", a, b, tol))
# bigger fused VEGAS runs: the block-cyclic cube shards of tq_vegas_run_fused_sharded (65536 / 6561 cubes), several
# checkpoints of the schedule, both dtypes; the evaluation counts must agree too (same nh on every world size up to a
# last-ulp flip of one floor())
g8 = F.GenzOscillatory(8, a=0.5, u=0.3)
a, b, integ = both(tq.VEGAS, g8, 8, dict(N=20_000_000, seed=5), torch.float64)
checks.append(("VEGAS fused 8-D float64 N=2e7", a, b, 1e-9))
assert integ._shard is not None and integ.strat.dh.shape[0] == integ._shard[1] < integ.strat.N_cubes
tq.distributed.disable()
ref = tq.VEGAS()
ref.integrate(g8, 8, N=20_000_000, integration_domain=torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=dev), seed=5)
checks.append(("VEGAS fused 8-D fevals", float(ref._nr_of_fevals), float(integ._nr_of_fevals), 1e-6))
checks.append(("VEGAS fused 8-D iterations", float(ref.it), float(integ.it), 0.0))
g6 = F.GenzProductPeak(6, a=2.0, u=0.5)
a, b, integ = both(tq.VEGAS, g6, 6, dict(N=3_000_000, seed=2, max_iterations=10), torch.float32)
assert integ._shard is not None
checks.append(("VEGAS fused 6-D float32", a, b, 5e-3))
# large-map record layout ({x, dx, weight, count} per bin), forced on this small problem: the histogram is
# unpacked into weights/counts before the all-reduce, so sharded == single still holds
from torchquad_b200.integration.vegas_map import VEGASMap

default_threshold = VEGASMap.records_min_bytes
VEGASMap.records_min_bytes = 0
a, b, integ = both(tq.VEGAS, g, 4, dict(N=400_000, seed=3), torch.float64)
checks.append(("VEGAS fused records float64", a, b, 1e-9))
assert integ.map._records is not None
# maps beyond L2 with >= 2^20 rows per pass: deferred histogram + band sweeps, every rank sweeping its own cubes
# (hist_sweep_launch with a CubeShard); the single-GPU run takes the same path with all cubes
a, b, integ = both(tq.VEGAS, g, 4, dict(N=30_000_000, seed=4), torch.float64)
checks.append(("VEGAS fused band sweeps float64", a, b, 1e-9))
assert integ._shard is not None and integ.map.sweep_group(integ.strat.N_strat) >= 1
a, b, integ = both(tq.VEGAS, g6, 6, dict(N=40_000_000, seed=2, max_iterations=10), torch.float32)
checks.append(("VEGAS fused band sweeps 6-D float32", a, b, 5e-3))
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
