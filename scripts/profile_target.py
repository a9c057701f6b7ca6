"""Small drivers for ncu captures (scripts/capture_ncu.py):  python scripts/profile_target.py <target>
Every target prints `UNITS <n>` = the work units (evaluations / samples / values) of the launch that is captured, which
is always the LAST launch of the named kernel."""
import sys
import warnings

sys.path.insert(0, ".")
import torch

import torchquad_b200 as tq
from torchquad_b200 import integrands as F
from torchquad_b200 import ops

warnings.simplefilter("ignore")
dev = torch.device("cuda")
what = sys.argv[1]


def vegas_final_pass(fn, dim, N, dt, cap):
    """Run the workload, then launch ONE stratified pass over its final map and sample allocation (the launch to capture)."""
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
    v = tq.VEGAS()
    v.max_map_intervals = cap
    r = v.integrate(fn, dim, N=N, integration_domain=dom, seed=1)
    vmap, strat = v.map, v.strat
    offsets = strat._offsets
    rows = int(offsets[-1].item())
    JF = torch.zeros((2, strat.N_cubes), dtype=dt, device=dev)
    g = vmap.sweep_group(strat.N_strat)
    if vmap.wants_records() and g >= 1:  # deferred pass + band sweeps: the LAST fused_vegas / hist_sweep launches are captured
        jf2 = torch.empty(rows, dtype=dt, device=dev)
        h = vmap.hist_pairs()
        ops.fused_vegas_deferred(v._fn_struct, vmap.packed_edges(), 0, rows, 1, 7, offsets, strat.N_strat, JF[0], JF[1], jf2)
        ops.hist_sweep(offsets, strat.N_strat, dim, jf2, vmap.N_intervals, h, g, 1, 7)
    elif vmap.wants_records():
        ops.fused_vegas(v._fn_struct, None, None, None, 0, rows, 1, 7, offsets=offsets, n_strat=strat.N_strat, JF=JF[0], JF2=JF[1],
                        records=vmap.records(), dtype=dt, n_intervals=vmap.N_intervals)
    else:
        h = vmap.hist_pairs()
        h.zero_()
        ops.fused_vegas(v._fn_struct, vmap.packed_edges(), None, None, 0, -rows, 1, 7, offsets=offsets, n_strat=strat.N_strat,
                        JF=JF[0], JF2=JF[1], hist_pairs=h)  # device-side row count, like the passes of the run
    torch.cuda.synchronize()
    print("RESULT", float(r), "fevals", v._nr_of_fevals, "Ni", vmap.N_intervals)
    print("UNITS", rows)


if what == "mc10":
    m = tq.MonteCarlo()
    for s in range(2):
        r = m.integrate(F.SumOfSines(10), 10, N=10**9, integration_domain=torch.tensor([[0.0, 1.0]] * 10, device=dev), seed=s)
    print("RESULT", float(r))
    print("UNITS", 10**9)
elif what == "boole6":
    dom = torch.tensor([[0.0, 1.0]] * 6, dtype=torch.float64, device=dev)
    b = tq.Boole()
    for _ in range(2):
        r = b.integrate(F.ProductOfCosines(6), 6, N=33**6, integration_domain=dom)
    print("RESULT", float(r))
    print("UNITS", b._nr_of_fevals)
elif what == "vegas8_cap4096":
    vegas_final_pass(F.GenzOscillatory(8, a=0.5, u=0.3), 8, 2_500_000_000, torch.float64, 4096)
elif what == "vegas8":
    vegas_final_pass(F.GenzOscillatory(8, a=0.5, u=0.3), 8, 2_500_000_000, torch.float64, None)
elif what == "vegas16_cap4096":
    vegas_final_pass(F.GenzProductPeak(16, a=2.0, u=0.5), 16, 10**10, torch.float32, 4096)
elif what == "vegas4":
    dom = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=dev)
    for s in range(3):
        v = tq.VEGAS()
        r = v.integrate(F.GenzGaussian(4, a=5.0, u=0.5), 4, N=10**6, integration_domain=dom, seed=s)
    print("RESULT", float(r))
    print("UNITS", v._nr_of_fevals)
elif what == "uniform_f32_d10":
    dom = torch.tensor([[0.0, 1.0]] * 10, dtype=torch.float32, device=dev)
    for _ in range(2):
        p = ops.mc_sample(dom, 2 * 10**8, 1, 0, 0)
    torch.cuda.synchronize()
    print("UNITS", 2 * 10**8 * 10)
elif what == "sum1_f32":
    f = torch.rand(2 * 10**8, dtype=torch.float32, device=dev)
    for _ in range(2):
        ops.sum_columns(f)
    torch.cuda.synchronize()
    print("UNITS", 2 * 10**8)
elif what == "contract1_f64":
    n, dim = 33, 6
    nodes = torch.linspace(0, 1, n, dtype=torch.float64, device=dev).repeat(dim, 1).contiguous()
    P = 400_000_000
    f = ops.philox_uniform(P, 1, torch.float64, dev, 1, 0).reshape(-1)
    for _ in range(2):
        ops.nc_contract(f, nodes, 0, P)
    torch.cuda.synchronize()
    print("UNITS", P)
elif what == "unfused8_cap":  # callback integrand, capped map: the two-kernel pass (5 warm-up + 10 stratified passes)
    dom = torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=dev)
    v = tq.VEGAS()
    v.max_map_intervals = 4096
    r = v.integrate(lambda x: torch.cos(2.0 * 3.141592653589793 * 0.3 + torch.sum(0.5 * x, dim=1)), 8, N=250_000_000,
                    integration_domain=dom, seed=1)
    print("RESULT", float(r))
    print("UNITS", int(v.strat._offsets[-1].item()))
elif what == "unfused8":
    dom = torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=dev)
    fn = F.GenzOscillatory(8, a=0.5, u=0.3)
    r = tq.VEGAS().integrate(lambda x: fn(x), 8, N=500_000_000, integration_domain=dom, seed=1)
    print("RESULT", float(r))
else:
    raise SystemExit(f"unknown target {what}")
