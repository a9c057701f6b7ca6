"""cProfile of small eager integrate() calls (host overhead per call)."""
import cProfile, pstats, sys, warnings, time
sys.path.insert(0, ".")
import torch
import torchquad_b200 as tq
warnings.simplefilter("ignore")
dev = torch.device("cuda")
torch.set_default_dtype(torch.float32)
fn = lambda x: torch.sin(x).sum(dim=1)
dom = torch.tensor([[0.0, 1.0]] * 3, device=dev)
for name, call in (("Simpson", lambda: tq.Simpson().integrate(fn, 3, N=21**3, integration_domain=dom)),
                   ("MonteCarlo", lambda: tq.MonteCarlo().integrate(fn, 3, N=10_000, integration_domain=dom, seed=1))):
    for _ in range(20): call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): r = call()
    float(r); print(name, "us/call", (time.perf_counter() - t0) / 200 * 1e6)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): call()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
