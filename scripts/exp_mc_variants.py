"""Fused Monte Carlo kernel variants (TQB200_LIB selects the library): configs[1] rate, CUDA events."""
import statistics
import sys

sys.path.insert(0, ".")
import torch

import torchquad_b200 as tq
from torchquad_b200 import integrands as F

dev = torch.device("cuda")
for name, fn, dim, dt, N in [("mc10 sum_sin f32", F.SumOfSines(10), 10, torch.float32, 10**9),
                             ("mc8 osc f64", F.GenzOscillatory(8, a=0.5, u=0.3), 8, torch.float64, 5 * 10**8),
                             ("mc5 gauss f32", F.GenzGaussian(5, a=3.0, u=0.5), 5, torch.float32, 10**9)]:
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
    m = tq.MonteCarlo()
    ts = []
    for s in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = m.integrate(fn, dim, N=N, integration_domain=dom, seed=s)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    t = statistics.median(ts[2:])
    print(f"{name:20s} {t*1e3:8.3f} ms  {N/t:.4e} evals/s  result {float(r):.6f} (exact {fn.exact():.6f})", flush=True)
