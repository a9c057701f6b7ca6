"""Latency of small-N repeated quadrature: eager integrate() vs the CUDA-graph replay of get_jit_compiled_integrate."""
import sys
import time
import warnings

sys.path.insert(0, ".")
import torch

import torchquad_b200 as tq

warnings.simplefilter("ignore")
dev = torch.device("cuda")
torch.set_default_dtype(torch.float32)


def fn(x):
    return torch.sin(x).sum(dim=1)


def bench(label, call, n=300):
    for _ in range(10):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = call()
    float(r)
    torch.cuda.synchronize()
    print(f"{label}: {(time.perf_counter() - t0) / n * 1e6:.1f} us/call")


for dim, N in ((3, 10_000), (5, 100_000), (10, 1_000_000)):
    dom = torch.tensor([[0.0, 1.0]] * dim, device=dev)
    mc = tq.MonteCarlo()
    comp = mc.get_jit_compiled_integrate(dim=dim, N=N, integration_domain=dom, seed=1, capture_integrand=True)
    bench(f"MC dim={dim} N={N} eager   ", lambda: mc.integrate(fn, dim, N=N, integration_domain=dom, seed=1))
    bench(f"MC dim={dim} N={N} compiled", lambda: comp(fn, dom))
for dim, n in ((3, 21), (4, 17), (6, 9)):
    dom = torch.tensor([[0.0, 1.0]] * dim, device=dev)
    sp = tq.Simpson()
    comp = sp.get_jit_compiled_integrate(dim=dim, N=n**dim, integration_domain=dom, capture_integrand=True)
    bench(f"Simpson dim={dim} n={n} eager   ", lambda: sp.integrate(fn, dim, N=n**dim, integration_domain=dom))
    bench(f"Simpson dim={dim} n={n} compiled", lambda: comp(fn, dom))
