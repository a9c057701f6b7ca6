"""Achieved HBM GB/s of every kernel on the unfused path (CUDA events, kernel alone, inputs >> L2)."""
import sys, torch, statistics
sys.path.insert(0, ".")
from torchquad_b200 import ops
from oracle import ref_oracle as O
dev = torch.device("cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e-3)
    return statistics.mean(ts)
def report(name, bytes_alg, s):
    print(f"{name:58s} {bytes_alg/s/1e9:8.1f} GB/s  {s*1e3:8.3f} ms  ({bytes_alg/1e9:.2f} GB algorithmic)", flush=True)
for dt, dim in [(torch.float64, 8), (torch.float32, 8), (torch.float32, 16), (torch.float64, 4)]:
    es = 8 if dt == torch.float64 else 4
    M = 100_000_000 if dim <= 8 else 50_000_000
    print(f"--- {dt} dim={dim} M={M:.0e}")
    for ni in (4096, 2_000_000):
        xe, dxe, w, c = (x.to(dev) for x in O.map_init(ni, dim, dt))
        y = ops.philox_uniform(M, dim, dt, dev, 1, 0) * 0.999999
        s = t(lambda: ops.map_forward(y, xe, dxe))
        report(f"map_forward Ni={ni}", M * (2 * dim + 1) * es, s)
        jf2 = ops.philox_uniform(M, 1, dt, dev, 2, 0).reshape(-1)
        s = t(lambda: ops.map_accumulate(y, jf2, w, c))
        report(f"map_accumulate Ni={ni} (+{2*dim} atomics/row)", M * (dim + 1) * es, s)
        del y, jf2
    ns = 8
    C = ns**dim if ns**dim < 2**24 else 2**24
    dh = torch.full((C,), 1.0 / C, dtype=dt, device=dev)
    nh, offsets = ops.strat_nh(dh, M)
    Mr = int(offsets[-1])
    s = t(lambda: ops.strat_nh(dh, M))
    report(f"strat_nh C={C}", C * (2 * es + 16), s)
    s = t(lambda: ops.strat_sample(offsets, ns, dim, dt, 0, Mr, seed=1, call_idx=0))
    report(f"strat_sample M={Mr}", Mr * dim * es, s)
    jf = ops.philox_uniform(Mr, 1, dt, dev, 2, 0).reshape(-1)
    s = t(lambda: ops.strat_accumulate(jf, offsets))
    report(f"strat_accumulate", Mr * es + C * (2 * es + 8), s)
    JF, JF2 = ops.strat_accumulate(jf, offsets)
    s = t(lambda: ops.strat_update(JF, JF2, nh, 1.0 / C, 0.75))
    report(f"strat_update", C * (2 * es + 8 + 3 * es), s)
    del jf
for n, dim in [(33, 6), (101, 4)]:
    dt = torch.float64
    nodes = torch.linspace(0, 1, n, dtype=dt, device=dev).repeat(dim, 1).contiguous()
    P = min(n**dim, 400_000_000)
    s = t(lambda: ops.nc_grid_points(nodes, 0, P))
    report(f"nc_grid_points n={n} dim={dim}", P * dim * 8, s)
    f = ops.philox_uniform(P, 1, dt, dev, 1, 0).reshape(-1)
    s = t(lambda: ops.nc_contract(f, nodes, 0, P))
    report(f"nc_contract n={n} dim={dim}", P * 8, s)
    del f
