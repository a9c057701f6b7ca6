"""contract1_kernel / contractk_kernel alone: achieved HBM GB/s (CUDA events) + agreement with a torch fp64 einsum."""
import statistics
import sys

sys.path.insert(0, ".")
import torch

from torchquad_b200 import ops

dev = torch.device("cuda")


def t(fn, reps=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return statistics.median(ts)


for dt in (torch.float64, torch.float32):
    for n, dim, P in [(33, 6, 400_000_000), (101, 4, 101**4), (5, 10, 5**10), (2, 24, 2**24), (3, 17, 3**17)]:
        w = (torch.rand(dim, n, dtype=torch.float64, device=dev) + 0.5).to(dt)
        P = min(P, n**dim)
        f = ops.philox_uniform(P, 1, dt, dev, 1, 0).reshape(-1)
        s = t(lambda: ops.nc_contract(f, w, 0, P))
        got = float(ops.nc_contract(f, w, 0, P))
        if n**dim <= 2 * 10**8:
            W = w[0].double()
            for d in range(1, dim):
                W = (W[:, None] * w[d].double()[None, :]).reshape(-1)
            want = float((f.double() * W[:P]).sum())
            err = abs(got - want) / abs(want)
        else:
            err = float("nan")
        es = f.element_size()
        print(f"{str(dt):14s} n={n:3d} dim={dim:2d} P={P:.2e}: {s*1e3:7.3f} ms  {P*es/s/1e9:7.1f} GB/s  rel err vs torch fp64 {err:.1e}", flush=True)
        del f
