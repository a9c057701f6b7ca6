"""Vector-valued integrands (tutorial.rst:741-771 of the reference): achieved HBM rate of the reductions."""
import sys, statistics
sys.path.insert(0, ".")
import torch
from torchquad_b200 import ops
dev = torch.device("cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e-3)
    return statistics.mean(ts)
for dt in (torch.float32, torch.float64):
    es = 4 if dt == torch.float32 else 8
    for rows, cols in ((100_000_000, 1), (50_000_000, 2), (20_000_000, 16), (2_000_000, 200), (200_000, 2048)):
        f = torch.rand(rows, cols, dtype=dt, device=dev)
        s = t(lambda: ops.sum_columns(f))
        print(f"sum_columns {dt} [{rows}, {cols}]: {rows * cols * es / s / 1e9:8.1f} GB/s  {s * 1e3:.3f} ms")
        del f
    n, dim = 41, 4
    w = torch.rand(dim, n, dtype=dt, device=dev)
    for cols in (1, 3, 16, 200):
        f = torch.rand(n**dim, cols, dtype=dt, device=dev)
        s = t(lambda: ops.nc_contract(f, w))
        print(f"nc_contract {dt} n={n} dim={dim} cols={cols}: {f.numel() * es / s / 1e9:8.1f} GB/s  {s * 1e3:.3f} ms")
        del f
