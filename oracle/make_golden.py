"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Run:  python oracle/make_golden.py
Needs /root/reference (read-only) and the autoray stand-in in
oracle/autoray_standin.  Every case feeds identical seeded inputs to the reference
classes and records inputs + outputs; while doing so it asserts that
oracle/ref_oracle.py reproduces the reference BITWISE, which is what pins the
oracle.  The GPU box has no /root/reference: there the fixtures written here are
the pin (tests/test_oracle_pinning.py) and the parity targets for the CUDA path
(tests/test_gpu_*.py).
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("TQ_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path[:0] = [os.path.join(HERE, "autoray_standin"), REF, ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import torchquad  # noqa: E402  (the reference)
from torchquad.integration.vegas_map import VEGASMap  # noqa: E402
from torchquad.integration.vegas_stratification import VEGASStratification  # noqa: E402
from oracle import ref_oracle as O  # noqa: E402

torchquad.set_log_level("ERROR")
warnings.filterwarnings("ignore")
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
DT = {"f32": torch.float32, "f64": torch.float64}


def same(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape, what
    assert torch.equal(a, b) or bool(((a == b) | (a.isnan() & b.isnan())).all()), f"oracle != reference: {what}"


def npy(t):
    return t.detach().cpu().numpy().copy()  # copy: the reference mutates some tensors in place later


class Injected:
    """rng object whose uniform() replays the oracle's Philox stream (tests/vegas_test.py:143-156 pattern)."""

    def __init__(self, seed):
        self.seed, self.call = seed, 0

    def uniform(self, size, dtype):
        u = O.philox_uniform(self.seed, self.call, 0, size[0], size[1], dtype)
        self.call += 1
        return u


def peak(x):
    return torch.exp(-torch.sum(25.0 * (x - 0.3) ** 2, dim=1)) + 0.01


def case_vegas_map(tag, dtype):
    g = torch.Generator().manual_seed(11)
    out = {}
    # the reference's own golden vector (tests/vegas_map_test.py:28-48)
    m = VEGASMap(20, 3, "torch", dtype)
    y0 = torch.tensor([[0.8121, 0.4319, 0.1612], [0.4746, 0.6501, 0.9241], [0.6143, 0.0724, 0.5818]], dtype=dtype)
    ids = m._get_interval_ID(y0)
    assert ids.tolist() == [[16, 8, 3], [9, 13, 18], [12, 1, 11]]
    same(ids, O.interval_id(y0, 20), "ids")
    same(m._get_interval_offset(y0), O.interval_offset(y0, 20), "offset")
    out.update(y0=npy(y0), ids0=npy(ids), off0=npy(m._get_interval_offset(y0)), x0=npy(m.get_X(y0)))
    # three adaptive updates on a peaked integrand, dim 4, Ni 50
    dim, ni, M = 4, 50, 6000
    m = VEGASMap(ni, dim, "torch", dtype)
    xe, dxe, w, c = O.map_init(ni, dim, dtype)
    same(m.x_edges, xe, "init x")
    same(m.dx_edges, dxe, "init dx")
    for it in range(3):
        y = torch.rand(M, dim, generator=g, dtype=torch.float64).to(dtype) * 0.999999
        x, jac = m.get_X(y), m.get_Jac(y)
        same(x, O.map_get_x(y, xe, dxe), "x")
        same(jac, O.map_get_jac(y, dxe), "jac")
        jf2 = (peak(x) * jac) ** 2
        m.accumulate_weight(y, jf2)
        O.map_accumulate(w, c, y, jf2)
        same(m.weights, w, "weights")
        same(m.counts, c, "counts")
        sm = VEGASMap._smooth_map(m.weights.clone(), m.counts.clone(), 0.5)
        same(sm, O.smooth_map(w, c, 0.5), "smooth")
        out.update({f"y{it+1}": npy(y), f"x{it+1}": npy(x), f"jac{it+1}": npy(jac), f"jf2_{it+1}": npy(jf2),
                    f"w{it+1}": npy(m.weights), f"c{it+1}": npy(m.counts), f"sm{it+1}": npy(sm)})
        m.update_map()
        xe, dxe, st = O.map_update(xe, dxe, w, c, 0.5)
        w, c = O.map_reset(ni, dim, dtype)
        assert st == "ok"
        same(m.x_edges, xe, "x_edges")
        same(m.dx_edges, dxe, "dx_edges")
        out.update({f"xe{it+1}": npy(xe), f"dxe{it+1}": npy(dxe)})
    # smoothing with zero-count runs incl. the reference's 2x6 golden (tests/vegas_map_test.py:84-100)
    w6 = torch.tensor([[0, 0, 0, 1, 1, 1], [0, 0, 0, 1, 0, 0]], dtype=dtype)
    c6 = torch.ones(w6.shape, dtype=torch.int64)
    s6 = VEGASMap._smooth_map(w6.clone(), c6, 0.5)
    same(s6, O.smooth_map(w6, c6, 0.5), "smooth 2x6")
    wz = torch.rand(3, 80, generator=g, dtype=torch.float64).to(dtype)
    cz = torch.randint(1, 9, (3, 80), generator=g)
    for lo, hi in [(0, 4), (10, 25), (40, 52), (70, 80)]:
        cz[:, lo:hi] = 0
        wz[:, lo:hi] = 0
    cz[1, 30:33] = 0
    wz[1, 30:33] = 0
    sz = VEGASMap._smooth_map(wz.clone(), cz.clone(), 0.5)
    same(sz, O.smooth_map(wz, cz, 0.5), "smooth zero runs")
    out.update(w6=npy(w6), c6=npy(c6), s6=npy(s6), wz=npy(wz), cz=npy(cz), sz=npy(sz))
    # update with a zero-count map region (uses the fill), and the all-zero skip
    m = VEGASMap(80, 3, "torch", dtype)
    m.weights, m.counts = wz.clone(), cz.clone()
    m.update_map()
    xe, dxe, _, _ = O.map_init(80, 3, dtype)
    xe, dxe, st = O.map_update(xe, dxe, wz, cz, 0.5)
    same(m.x_edges, xe, "zero-run x_edges")
    out.update(xez=npy(xe), dxez=npy(dxe))
    np.savez_compressed(os.path.join(OUT, f"vegas_map_{tag}.npz"), **out)


def case_strat(tag, dtype):
    g = torch.Generator().manual_seed(5)
    out = {}
    n_inc, dim = 1000, 3
    rng = Injected(7)
    s = VEGASStratification(n_inc, dim, rng, "torch", dtype)
    ns, nc, vc = O.strat_config(n_inc, dim)
    assert (s.N_strat, s.N_cubes, s.V_cubes) == (ns, nc, vc)
    dh = O.strat_init(nc, dtype)
    same(s.dh, dh, "dh0")
    for it in range(3):
        nev = 4000 + 1500 * it
        nh = s.get_NH(nev)
        same(nh, O.strat_get_nh(dh, nev), "nh")
        call = rng.call
        y = s.get_Y(nh)
        u = O.philox_uniform(7, call, 0, int(nh.sum()), dim, dtype)
        same(y, O.strat_get_y(nh, ns, dim, u), "y")
        jf = torch.prod(torch.exp(3.0 * y), dim=1)
        JF, JF2 = s.accumulate_weight(nh, jf)
        oJF, oJF2 = O.strat_accumulate(nh, jf)
        same(JF, oJF, "JF")
        same(JF2, oJF2, "JF2")
        I, s2 = O.vegas_iteration_estimate(oJF, oJF2, nh, vc)
        s.update_DH()
        dh = O.strat_update_dh(oJF, oJF2, nh.to(dtype), vc, 0.75)
        same(s.dh, dh, "dh")
        out.update({f"nev{it}": np.int64(nev), f"nh{it}": npy(nh), f"y{it}": npy(y), f"u{it}": npy(u),
                    f"jf{it}": npy(jf), f"JF{it}": npy(JF), f"JF2{it}": npy(JF2), f"dh{it}": npy(dh),
                    f"I{it}": npy(I), f"s2{it}": npy(s2)})
    # get_NH on an injected, strongly peaked dh (bit-exact target)
    dhp = torch.rand(nc, generator=g, dtype=torch.float64) ** 8
    dhp = (dhp / dhp.sum()).to(dtype)
    s.dh = dhp
    nhp = s.get_NH(54321)
    same(nhp, O.strat_get_nh(dhp, 54321), "nh peaked")
    out.update(dhp=npy(dhp), nhp=npy(nhp), n_strat=np.int64(ns), n_cubes=np.int64(nc), v_cubes=np.float64(vc))
    np.savez_compressed(os.path.join(OUT, f"vegas_strat_{tag}.npz"), **out)


def case_vegas_run(tag, dtype):
    """Whole VEGAS.integrate on injected uniforms (reference vs oracle driver), small N."""
    out = {}
    for name, dim, N, fn, dom in [
        ("peak3", 3, 20000, peak, [[0.0, 1.0]] * 3),
        ("sin2", 2, 10000, lambda x: torch.sum(torch.sin(x), dim=1), [[0.0, 2.0], [-1.0, 1.0]]),
    ]:
        domain = torch.tensor(dom, dtype=dtype)
        v = torchquad.VEGAS()
        ref = v.integrate(fn, dim, N=N, integration_domain=domain, rng=Injected(3))
        run = O.VegasRun(fn, dim, N, domain, Injected(3).uniform)
        res = run.run()
        same(ref, res, f"vegas {name}")
        assert v._nr_of_fevals == run.fevals and v.it == run.it
        same(v.map.x_edges, run.x_edges, "final edges")
        same(v.strat.dh, run.dh, "final dh")
        out.update({f"{name}_result": npy(ref), f"{name}_fevals": np.int64(run.fevals), f"{name}_it": np.int64(run.it),
                    f"{name}_trace": np.array(run.trace, dtype=np.float64), f"{name}_x_edges": npy(run.x_edges),
                    f"{name}_dh": npy(run.dh), f"{name}_sigma2": np.array([float(s) for s in run.sigma2]),
                    f"{name}_results": np.array([float(s) for s in run.results]), f"{name}_domain": npy(domain)})
    np.savez_compressed(os.path.join(OUT, f"vegas_run_{tag}.npz"), **out)


def case_mc(tag, dtype):
    out = {}
    dom = torch.tensor([[0.0, 2.0], [-1.0, 1.5], [3.0, 3.5]], dtype=dtype)
    u = O.philox_uniform(1, 0, 0, 5000, 3, dtype)

    class R:
        def uniform(self, size, dtype):
            return u

    mc = torchquad.MonteCarlo()
    pts = mc.calculate_sample_points(5000, dom, rng=R())
    same(pts, O.mc_sample_points(u, dom), "mc points")
    f = torch.sum(torch.sin(pts), dim=1)
    res = mc.calculate_result(f, dom)
    same(res, O.mc_result(f, dom), "mc result")
    fv = torch.stack([f, 2 * f, f * f], dim=1)
    resv = mc.calculate_result(fv, dom)
    same(resv, O.mc_result(fv, dom), "mc vector result")
    out.update(domain=npy(dom), u=npy(u), points=npy(pts), f=npy(f), result=npy(res), fv=npy(fv), resultv=npy(resv))
    np.savez_compressed(os.path.join(OUT, f"monte_carlo_{tag}.npz"), **out)


def case_nc(tag, dtype):
    out = {}
    for rule, cls in [("trapezoid", torchquad.Trapezoid), ("simpson", torchquad.Simpson), ("boole", torchquad.Boole)]:
        for dim, N, dom in [(1, 401, [[-1.0, 2.0]]), (2, 1000, [[0.0, 1.0], [-2.0, 0.5]]), (3, 9 ** 3 + 5, [[0.0, 1.0], [1.0, 3.0], [-1.0, 1.0]]),
                            (4, 2, [[0.0, 1.0]] * 4) if rule != "trapezoid" else (4, 16, [[0.0, 1.0]] * 4)]:
            domain = torch.tensor(dom, dtype=dtype)
            integ = cls()
            pts, hs, n = integ.calculate_grid(N, domain)
            opts, ohs, on = O.nc_grid(rule, N, domain)
            same(pts, opts, f"{rule} grid")
            same(hs, ohs, f"{rule} h")
            assert n == on
            f = torch.prod(torch.cos(pts), dim=1) + torch.sum(pts**3, dim=1)
            res = integ.calculate_result(f, dim, n, hs, domain)
            same(res, O.nc_result(rule, f, dim, n, hs), f"{rule} result")
            fv = torch.stack([f, torch.sum(torch.exp(pts), dim=1)], dim=1)
            resv = integ.calculate_result(fv, dim, n, hs, domain)
            same(resv, O.nc_result(rule, fv, dim, n, hs), f"{rule} vec result")
            k = f"{rule}_d{dim}"
            out.update({f"{k}_domain": npy(domain), f"{k}_N": np.int64(N), f"{k}_n": np.int64(n), f"{k}_h": npy(hs),
                        f"{k}_points": npy(pts), f"{k}_f": npy(f), f"{k}_result": npy(res), f"{k}_fv": npy(fv),
                        f"{k}_resultv": npy(resv)})
    np.savez_compressed(os.path.join(OUT, f"newton_cotes_{tag}.npz"), **out)


def case_gauss():
    """GaussLegendre (fp64, the only dtype the reference supports here: its weights are float64 numpy arrays)."""
    out = {}
    dtype = torch.float64
    for dim, N, dom in [(1, 60, [[0.0, 5.0]]), (2, 8**2, [[0.0, 1.0], [-2.0, 0.5]]), (3, 5**3 + 3, [[0.0, 1.0], [1.0, 3.0], [-1.0, 1.0]])]:
        domain = torch.tensor(dom, dtype=dtype)
        gl = torchquad.GaussLegendre()
        pts, hs, n = gl.calculate_grid(N, domain)
        W = gl._weights(n, dim, "torch")
        opts, oW, on = O.gauss_grid(N, domain)
        same(pts, opts, "gauss points")
        same(W, oW, "gauss weights")
        assert n == on
        f = torch.prod(torch.cos(pts), dim=1) + torch.sum(pts**3, dim=1)
        ref = gl.integrate(lambda x: torch.prod(torch.cos(x), dim=1) + torch.sum(x**3, dim=1), dim, N, domain)
        same(ref, O.gauss_result(f, oW, dim, n, domain), "gauss result")
        fv = torch.stack([f, torch.sum(torch.exp(pts), dim=1)], dim=1)
        refv = gl.integrate(lambda x: torch.stack([torch.prod(torch.cos(x), dim=1) + torch.sum(x**3, dim=1),
                                                   torch.sum(torch.exp(x), dim=1)], dim=1), dim, N, domain)
        same(refv, O.gauss_result(fv, oW, dim, n, domain), "gauss vector result")
        k = f"d{dim}"
        out.update({f"{k}_domain": npy(domain), f"{k}_N": np.int64(N), f"{k}_n": np.int64(n), f"{k}_points": npy(pts),
                    f"{k}_W": npy(W), f"{k}_f": npy(f), f"{k}_result": npy(ref), f"{k}_fv": npy(fv), f"{k}_resultv": npy(refv)})
    np.savez_compressed(os.path.join(OUT, "gauss_legendre_f64.npz"), **out)


def case_reference_records():
    """Scalar records of the reference's own end-to-end runs with ITS RNG (torch CPU mt19937)."""
    torchquad.set_up_backend("torch", "float64", torch_enable_cuda=False)
    a, u = 5.0, 0.5

    def gauss(x):
        return torch.exp(-torch.sum(a * a * (x - u) ** 2, dim=1))

    v = torchquad.VEGAS()
    r = v.integrate(gauss, dim=4, N=10**6, integration_domain=[[0.0, 1.0]] * 4, seed=0)
    err = v._get_error()
    rec = dict(c1_result=np.float64(r), c1_error=np.float64(err), c1_fevals=np.int64(v._nr_of_fevals), c1_it=np.int64(v.it),
               c1_exact=np.float64(O.genz_exact("gaussian", [a] * 4, [u] * 4)))
    mc = torchquad.MonteCarlo()
    torchquad.set_precision("float32")
    r2 = mc.integrate(lambda x: torch.sum(torch.sin(x), dim=1), dim=10, N=10**6, integration_domain=[[0.0, 1.0]] * 10, seed=0)
    rec.update(c2_result_1e6=np.float32(r2), c2_exact=np.float64(20 * np.sin(0.5) ** 2))
    torchquad.set_precision("float64")
    b = torchquad.Boole()
    r3 = b.integrate(lambda x: torch.prod(torch.cos(x), dim=1), dim=6, N=13**6, integration_domain=[[0.0, 1.0]] * 6)
    rec.update(c3_boole_n13=np.float64(r3), c3_exact=np.float64(np.sin(1.0) ** 6))
    np.savez_compressed(os.path.join(OUT, "reference_records.npz"), **rec)
    print({k: float(v) for k, v in rec.items()})


if __name__ == "__main__":
    for tag, dt in DT.items():
        case_vegas_map(tag, dt)
        case_strat(tag, dt)
        case_vegas_run(tag, dt)
        case_mc(tag, dt)
        case_nc(tag, dt)
        print("golden", tag, "ok (oracle == reference bitwise)")
    case_gauss()
    print("golden gauss ok (oracle == reference bitwise)")
    case_reference_records()
    print("wrote", sorted(os.listdir(OUT)))
