"""VEGAS with a torch-callable integrand (the path every user of the reference is on): whole-run evals/s on one GPU.
    python scripts/time_vegas_unfused.py [N ...]"""
import sys
import time

sys.path.insert(0, ".")
import torch

import torchquad_b200 as tq

dev = torch.device("cuda")
CASES = {
    "osc8_f64": (8, torch.float64, lambda x: torch.cos(2.0 * 3.141592653589793 * 0.3 + torch.sum(0.5 * x, dim=1)), -0.6772157012),
    "peak6_f32": (6, torch.float32, lambda x: torch.prod(1.0 / (0.25 + (x - 0.5) ** 2), dim=1), None),
}
Ns = [int(float(a)) for a in sys.argv[1:]] or [25_000_000, 250_000_000, 2_500_000_000]
for name, (dim, dt, fn, exact) in CASES.items():
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
    for N in Ns:
        for cap in (4096, None):
            if cap is None and N < 10**9:
                continue
            v = tq.VEGAS()
            v.max_map_intervals = cap
            r = None
            ts = []
            for rep in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = float(v.integrate(fn, dim, N=N, integration_domain=dom, seed=rep))
                ts.append(time.perf_counter() - t0)
            t = min(ts[1:])
            print(f"{name} N={N:.1e} cap={cap}: {t*1e3:9.2f} ms  {v._nr_of_fevals/t:.3e} evals/s  it={v.it} fevals={v._nr_of_fevals} "
                  f"result={r:.8e}" + (f" exact={exact}" if exact else "") + f" peak_mem={torch.cuda.max_memory_allocated()/2**30:.1f} GiB",
                  flush=True)
            del v
            torch.cuda.empty_cache()
