import sys, torch, warnings
sys.path.insert(0, ".")
import torchquad_b200 as tq
from torchquad_b200 import integrands as F, ops
warnings.simplefilter("ignore")
dev = torch.device("cuda")
fn = F.GenzProductPeak(16, a=2.0, u=0.5)
for N in [10**7, 10**8, 10**9]:
    for dt in [torch.float32, torch.float64]:
        dom = torch.tensor([[0.0, 1.0]] * 16, dtype=dt, device=dev)
        v = tq.VEGAS()
        # monkeypatch update to check
        orig = v._update_map if hasattr(v, "_update_map") else None
        def upd(self=v):
            m = self.map
            w, c = m.weights.clone(), m.counts.clone()
            bad_w = int((~torch.isfinite(w)).sum())
            zc = int((c == 0).sum())
            m.update_map(check=False)
            st = m._status.tolist()
            nf = int((~torch.isfinite(m.x_edges)).sum())
            neg = int((m.dx_edges <= 0).sum())
            print(f"  N={N:.0e} {dt} it={self.it} Ni={m.N_intervals} wsum={float(w.sum()):.3e} wmax={float(w.max()):.3e} nonfinite_w={bad_w} zero_counts={zc} status={st} nonfinite_edges={nf} dx<=0:{neg}")
            if st[2]:
                sm, s2 = ops.map_smooth(w, c, 0.5)
                print("   smooth: nonfinite", int((~torch.isfinite(sm)).sum()), "zeros", int((sm == 0).sum()), "min", float(sm.min()), "max", float(sm.max()))
                raise SystemExit
        v._update_map = upd
        try:
            r = v.integrate(fn, 16, N=N, integration_domain=dom, seed=1)
            print(f"N={N:.0e} {dt}: result {float(r):.6e} exact {fn.exact():.6e} fevals {v._nr_of_fevals}")
        except Exception as e:
            print("ERR", type(e).__name__, e)
