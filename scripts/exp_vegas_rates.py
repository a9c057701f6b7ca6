import sys, time, warnings
sys.path.insert(0, ".")
import torch
import torchquad_b200 as tq
from torchquad_b200 import integrands as F
warnings.simplefilter("ignore")
dev = torch.device("cuda")
def run(label, fn, dim, N, dt, cap=None, reps=2):
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
    v = tq.VEGAS(); v.max_map_intervals = cap
    for rep in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = v.integrate(fn, dim, N=N, integration_domain=dom, seed=rep)
        torch.cuda.synchronize(); dt_s = time.perf_counter() - t
    print(f"{label:52s} {dt_s*1e3:9.1f} ms {v._nr_of_fevals/dt_s:.3e} evals/s  res {float(r):.6e} +- {float(v._get_error()):.1e} Ni={v.map.N_intervals} C={v.strat.N_cubes}", flush=True)
g8 = F.GenzOscillatory(8, a=0.5, u=0.3)
run("fused 8D f64 N=2.5e9 cap4096", g8, 8, 2_500_000_000, torch.float64, 4096)
run("fused 8D f32 N=2.5e9 cap4096", g8, 8, 2_500_000_000, torch.float32, 4096)
run("unfused 8D f64 N=5e8 refmap", lambda x: g8(x), 8, 500_000_000, torch.float64)
run("unfused 8D f64 N=5e8 cap4096", lambda x: g8(x), 8, 500_000_000, torch.float64, 4096)
run("unfused 8D f32 N=5e8 cap4096", lambda x: g8(x), 8, 500_000_000, torch.float32, 4096)
g16 = F.GenzProductPeak(16, a=2.0, u=0.5)
run("fused 16D f32 N=1e10 cap4096", g16, 16, 10**10, torch.float32, 4096, reps=2)
run("fused 16D f64 N=1e10 cap4096", g16, 16, 10**10, torch.float64, 4096, reps=1)
run("fused 16D f64 N=1e10 refmap", g16, 16, 10**10, torch.float64, None, reps=1)
g4 = F.GenzGaussian(4, a=5.0, u=0.5)
run("fused 4D f64 N=1e6 (C1)", g4, 4, 10**6, torch.float64, None, reps=3)
run("unfused 4D f64 N=1e6 (C1)", lambda x: g4(x), 4, 10**6, torch.float64, None, reps=3)
run("fused 4D f64 N=1e8", g4, 4, 10**8, torch.float64, None, reps=2)
run("fused 4D f64 N=1e9", g4, 4, 10**9, torch.float64, None, reps=2)
