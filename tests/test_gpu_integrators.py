"""End-to-end parity of the integrator classes (the drop-in surface) against the oracle / golden fixtures,
plus the behavioural pins of the reference test-suite (/root/reference/tests, cited per test)."""
import math
import warnings

import numpy as np
import pytest
import torch

import torchquad_b200 as tq
from oracle import ref_oracle as O
from torchquad_b200 import integrands as F
from torchquad_b200 import ops

pytestmark = pytest.mark.gpu
DT = {"f32": torch.float32, "f64": torch.float64}


class Injected:
    """rng whose uniform() replays the oracle's Philox stream (the reference's injection seam,
    /root/reference/tests/vegas_test.py:143-156)."""

    def __init__(self, seed):
        self.seed, self.call = seed, 0

    def uniform(self, size, dtype):
        u = O.philox_uniform(self.seed, self.call, 0, size[0], size[1], dtype)
        self.call += 1
        return u


def peak(x):
    return torch.exp(-torch.sum(25.0 * (x - 0.3) ** 2, dim=1)) + 0.01


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


# ---------------------------------------------------------------- VEGAS on identical injected samples
@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("name,dim,N,fn", [
    ("peak3", 3, 20000, peak),
    ("sin2", 2, 10000, lambda x: torch.sum(torch.sin(x), dim=1)),
])
def test_vegas_identical_samples_match_reference(cuda, golden, tag, name, dim, N, fn, record_delta):
    g = golden(f"vegas_run_{tag}")
    domain = torch.from_numpy(g[f"{name}_domain"]).to(cuda)
    v = tq.VEGAS()
    res = v.integrate(fn, dim, N=N, integration_domain=domain, rng=Injected(3))
    want = float(g[f"{name}_result"])
    assert res.dtype == DT[tag] and res.device.type == "cuda" and res.dim() == 0
    assert v.it == int(g[f"{name}_it"])
    if tag == "f64":
        # schedule, sample counts and per-iteration estimates follow the reference run exactly
        assert v._nr_of_fevals == int(g[f"{name}_fevals"])
        # whole run on identical injected uniforms: 10 iterations of compounding state (map, dh).  The per-iteration
        # estimate is a sum over cubes of JF_c * V / n_c whose terms agree to the last bits; what is reported is the
        # end-to-end drift after all iterations.
        e_res = record_delta(f"VEGAS whole run {name} f64: |I - ref| / |ref|", abs(float(res) - want) / abs(want), 1e-12)
        e_it = record_delta(f"VEGAS whole run {name} f64: max rel err of per-iteration estimates",
                            float(np.max(np.abs(np.array([float(r) for r in v.results]) - g[f"{name}_results"])
                                         / np.abs(g[f"{name}_results"]))), 1e-12)
        e_s2 = record_delta(f"VEGAS whole run {name} f64: max rel err of per-iteration sigma^2",
                            float(np.max(np.abs(np.array([float(s_) for s_ in v.sigma2]) - g[f"{name}_sigma2"])
                                         / np.abs(g[f"{name}_sigma2"]))), 1e-8)
        e_map = record_delta(f"VEGAS whole run {name} f64: max |x_edges - ref| after the last update",
                             float((v.map.x_edges.cpu() - torch.from_numpy(g[f"{name}_x_edges"])).abs().max()), 1e-12)
        assert e_res <= 1e-12 and e_it <= 1e-12  # north_star: estimates within 1e-12 relative in fp64
        assert e_s2 <= 1e-8   # sigma^2 = |JF2 V^2/n - ih^2| is a difference of close numbers
        assert e_map <= 1e-12
        assert torch.allclose(v.strat.dh.cpu(), torch.from_numpy(g[f"{name}_dh"]), rtol=1e-7, atol=1e-12)  # d = difference of close numbers
    else:
        # fp32: a floor() in get_NH may flip on a last-ulp difference of pow(); compare statistically
        sig = math.sqrt(float(sum(g[f"{name}_sigma2"])) / len(g[f"{name}_sigma2"]))
        assert abs(float(res) - want) <= max(1e-4 * abs(want), 3 * sig)
        assert abs(v._nr_of_fevals - int(g[f"{name}_fevals"])) <= 0.001 * int(g[f"{name}_fevals"])


@pytest.mark.parametrize("name,dim,N,fn", [
    ("peak3", 3, 20000, peak),
    ("sin2", 2, 10000, lambda x: torch.sum(torch.sin(x), dim=1)),
    ("peak5", 5, 200000, lambda x: torch.exp(-torch.sum(9.0 * (x - 0.4) ** 2, dim=1)) + 0.05),
])
def test_vegas_fp32_stepwise_pinned_to_oracle(cuda, name, dim, N, fn, record_delta):
    """fp32 whole runs can only be compared statistically with the reference (one last-ulp flip of a floor() in get_NH
    changes the sample set from there on).  This test pins every step instead: the oracle (= the reference's arithmetic,
    tests/test_oracle_pinning.py) runs the whole schedule on injected uniforms; for EVERY pass the CUDA path starts from
    the oracle's state (map edges, dh) and must reproduce that pass: nh and the samples bit-exactly, the estimate, the
    new map and the new dh within 1e-5 (north_star's fp32 bar).  A real regression cannot hide inside 3 sigma here."""
    from torchquad_b200.integration.vegas_map import VEGASMap
    from torchquad_b200.integration.vegas_stratification import VEGASStratification

    dt = torch.float32
    domain = torch.tensor([[0.0, 1.0]] * dim, dtype=dt)
    inj = Injected(11)
    run = O.VegasRun(fn, dim, N, domain, inj.uniform)
    vmap = VEGASMap(run.n_intervals, dim, "torch", dt, device=cuda)
    strat = VEGASStratification(run.n_increment, dim=dim, rng=None, backend="torch", dtype=dt, device=cuda)
    assert strat.N_strat == run.n_strat and strat.N_cubes == run.n_cubes
    worst = {"I": 0.0, "s2": 0.0, "x": 0.0, "dx": 0.0, "dh": 0.0}

    def load_map():
        vmap.x_edges.copy_(run.x_edges)
        vmap.dx_edges.copy_(run.dx_edges)
        vmap.invalidate_packed()
        vmap._reset_weight()

    def check_map(tag):
        worst["x"] = max(worst["x"], float((vmap.x_edges.cpu() - run.x_edges).abs().max()))
        worst["dx"] = max(worst["dx"], float((vmap.dx_edges.cpu() - run.dx_edges).abs().max()))

    # warm-up passes (vegas.py:211-266): same uniforms, state reloaded from the oracle before each pass
    ns = run.starting_N // 5
    for _ in range(5):
        load_map()
        call0 = inj.call
        u = O.philox_uniform(inj.seed, call0, 0, ns, dim, dt)
        y = (u * 0.999999).to(cuda)
        x, jac = vmap.get_X_and_Jac(y)
        f = fn(x.cpu()).to(cuda)  # integrand values from the CPU, like the oracle's: torch's CUDA sin/exp differ in the last ulp
        vmap.accumulate_weight(y, ((f * jac) ** 2))
        vmap.update_map()
        run.warmup(1)  # consumes the same Philox call
        assert inj.call == call0 + 1
        check_map("warm-up")
    n_it = 0
    while True:
        run.it += 1
        run.results.append(0)
        run.sigma2.append(0)
        # ---- our pass from the oracle's state
        load_map()
        strat.dh = run.dh.to(cuda)
        nh = strat.get_NH(run.starting_N)
        nh_ref = O.strat_get_nh(run.dh, run.starting_N)
        assert torch.equal(nh.cpu(), nh_ref)  # get_NH bit-exact on identical dh
        M = int(nh_ref.sum())
        u = O.philox_uniform(inj.seed, inj.call, 0, M, dim, dt)
        y = ops.strat_sample(strat._offsets, strat.N_strat, dim, dt, 0, M, u_in=u.to(cuda))
        assert torch.equal(y.cpu(), O.strat_get_y(nh_ref, run.n_strat, dim, u))  # samples bit-exact
        x, jac = vmap.get_X_and_Jac(y)
        assert torch.equal(x.cpu(), O.map_get_x(y.cpu(), run.x_edges, run.dx_edges))
        assert torch.equal(jac.cpu(), O.map_get_jac(y.cpu(), run.dx_edges))
        jf = fn(x.cpu()).to(cuda) * jac  # same integrand values as the oracle (x is bit-identical): what is compared is OUR path
        vmap.accumulate_weight(y, jf**2)
        strat.accumulate_weight(nh, jf)
        strat.update_DH()
        I, s2 = float(strat.last_scalars[0]), float(strat.last_scalars[1])
        vmap.update_map()
        # ---- the oracle's pass on the same uniforms
        run.iteration()
        I_ref, s2_ref, M_ref = run.trace[-1]
        assert M_ref == M
        worst["I"] = max(worst["I"], abs(I - I_ref) / abs(I_ref))
        worst["s2"] = max(worst["s2"], abs(s2 - s2_ref) / abs(s2_ref))
        check_map("iteration")
        dh_ref = run.dh
        worst["dh"] = max(worst["dh"], float((strat.dh.cpu() - dh_ref).abs().max() / dh_ref.abs().max()))
        n_it += 1
        if run.check_abort():
            break
    assert n_it >= 5
    # x_edges live in [0, 1] (absolute = relative to the map, a few fp32 ulps); dx_edges = diff(x_edges) inherit at most
    # twice that absolute error -- the reference forms them by an fp32 cumsum and a subtraction (vegas_map.py:233-259)
    bounds = {"I": 1e-5, "s2": 1e-5, "x": 1e-6, "dx": 2e-6, "dh": 1e-5}
    for k, v_ in worst.items():
        record_delta(f"VEGAS fp32 step-wise {name} ({n_it} iterations): worst {k}", v_, bounds[k])
        assert v_ <= bounds[k], (k, v_)


def test_vegas_c1_gaussian_matches_reference_record(cuda, golden):
    """BASELINE configs[0]: 4-D Genz Gaussian, N=1e6, fp64.  Fresh RNG => agree within 3 sigma of the
    reference's reported error; same schedule (10 iterations, ~645k evaluations)."""
    rec = golden("reference_records")
    fn = F.GenzGaussian(4, a=5.0, u=0.5)
    exact = fn.exact()
    assert abs(exact - float(rec["c1_exact"])) < 1e-15
    for fused in (True, False):
        v = tq.VEGAS()
        f = fn if fused else (lambda x: fn(x))
        res = v.integrate(f, 4, N=10**6, integration_domain=torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=cuda), seed=0)
        err = float(v._get_error())
        assert v.it == int(rec["c1_it"])
        assert abs(v._nr_of_fevals - int(rec["c1_fevals"])) < 0.02 * int(rec["c1_fevals"])
        assert abs(float(res) - float(rec["c1_result"])) <= 3 * math.hypot(err, float(rec["c1_error"]))
        assert abs(float(res) - exact) <= 4 * err
        assert err < 3 * float(rec["c1_error"])


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_vegas_fused_equals_unfused(cuda, tag):
    """Fused and unfused paths draw the same cube-keyed Philox samples => same estimate up to summation order."""
    dt = DT[tag]
    dom = torch.tensor([[0.0, 1.0], [0.0, 2.0], [-1.0, 1.0]], dtype=dt, device=cuda)
    for fn in (F.GenzGaussian(3, a=[3.0, 2.0, 4.0], u=[0.3, 0.6, 0.5]), F.GenzProductPeak(3, a=2.0, u=0.4), F.SumOfSines(3)):
        a = tq.VEGAS()
        ra = a.integrate(fn, 3, N=60000, integration_domain=dom, seed=11)
        b = tq.VEGAS()
        rb = b.integrate(lambda x: fn(x), 3, N=60000, integration_domain=dom, seed=11)
        tol = 1e-9 if tag == "f64" else 2e-3
        assert a._nr_of_fevals == b._nr_of_fevals or tag == "f32"
        assert abs(float(ra) - float(rb)) <= tol * abs(float(rb))


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_vegas_native_loop_equals_python_loop(cuda, tag):
    """tq_vegas_run_fused (pass loop + schedule in C++) must reproduce the Python-driven loop exactly: same kernels,
    same Philox call indices, same schedule decisions in the working precision."""
    dt = DT[tag]
    for fn, dim, N, kw in [
        (F.GenzGaussian(4, a=5.0, u=0.5), 4, 10**6, {}),
        (F.GenzProductPeak(3, a=3.0, u=0.4), 3, 60_000, dict(max_iterations=12)),
        (F.GenzOscillatory(5, a=0.7, u=0.2), 5, 200_000, dict(eps_abs=1e-3)),
        (F.SumOfSines(2), 2, 10_000, dict(use_warmup=False)),
        (F.GenzC0(3, a=2.0, u=0.5), 3, 50_000, dict(use_grid_improve=False)),
    ]:
        dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=cuda)
        a, b = tq.VEGAS(), tq.VEGAS()
        b.native_loop = False
        ra = a.integrate(fn, dim, N=N, integration_domain=dom, seed=5, **kw)
        rb = b.integrate(fn, dim, N=N, integration_domain=dom, seed=5, **kw)
        # Float atomics (histogram weights, per-cube sums) make two fused runs agree to rounding, not bitwise: a
        # last-ulp difference in dh can flip a floor() in get_NH, so sample counts may differ by a few.
        assert a.it == b.it and a._starting_N == b._starting_N, (type(fn).__name__, tag)
        assert abs(a._nr_of_fevals - b._nr_of_fevals) <= 2e-3 * b._nr_of_fevals
        sigma = float(b._get_error())
        assert ra.dtype == rb.dtype == dt
        # fp64 runs agree to rounding.  In fp32 a flipped count changes a cube's samples, which feeds back through
        # the histogram into later maps: two runs of the SAME loop differ by a fraction of sigma (measured 0.1-0.9).
        tol = 1e-9 * abs(float(rb)) if tag == "f64" else 2.5 * sigma
        assert abs(float(ra) - float(rb)) <= tol
        assert len(a.results) == len(b.results) and a.rng._call == b.rng._call
        assert float((a.map.x_edges - b.map.x_edges).abs().max()) <= (1e-9 if tag == "f64" else 1e-3)
        assert abs(float(a._get_error()) - sigma) <= 0.05 * sigma


def test_vegas_special_cases(cuda):
    """/root/reference/tests/vegas_test.py:159-199."""
    torch.set_default_dtype(torch.float64)
    try:
        integ = tq.VEGAS()
        dom = torch.tensor([[0.0, 3.0]] * 2, device=cuda)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            zero = integ.integrate(lambda x: x[:, 0] * 0.0, 2, N=10000, integration_domain=dom, seed=0)
        assert float(zero.abs()) == 0.0
        const = integ.integrate(lambda x: x[:, 0] * 0.0 + 10.0, 2, N=10000, integration_domain=dom, seed=0)
        assert abs(float(const) - 90.0) < 1e-13

        class ModifiedRNG(tq.RNG):
            """Half of the uniforms replaced by exact 0.0 / 1.0 (vegas_test.py:143-156)."""

            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                base = self.uniform
                self.uniform = lambda *a, **k: self.modify(base(*a, **k))

            @staticmethod
            def modify(n):
                return torch.where(n < 0.5, n * 2.0, torch.where(n < 0.75, torch.zeros_like(n), torch.ones_like(n)))

        r = integ.integrate(lambda x: torch.sum(x, dim=1), 2, N=10000,
                            integration_domain=torch.tensor([[0.0, 1.0]] * 2, device=cuda), rng=ModifiedRNG(seed=0))
        assert isinstance(integ.rng, ModifiedRNG)
        assert abs(float(r) - 1.0) < 0.1
    finally:
        torch.set_default_dtype(torch.float32)


def test_vegas_peak_accuracy(cuda):
    """/root/reference/tests/vegas_test.py:71-140 (hypercube peak and diagonal peaks, 5 seeds)."""
    dt = torch.float64
    dom = torch.tensor([[1.0, 5.0], [-4.0, 4.0], [2.0, 6.0]], dtype=dt, device=cuda)

    def cube_peak(x):
        return torch.prod((x >= 3.0) * (x < 4.0), dim=1).to(dt) + 0.001

    ref = float(torch.prod(dom[:, 1] - dom[:, 0])) * 0.001 + 1.0
    for seed in [0, 1, 2, 3, 41317]:
        r = tq.VEGAS().integrate(cube_peak, 3, N=30000, integration_domain=dom, seed=seed)
        assert abs(float(r) - ref) < 0.03
    c = 100.0
    dom2 = torch.tensor([[1.0, 1.0 + c], [-4.0, -4.0 + c]], dtype=dt, device=cuda)

    def diag(x):
        return torch.exp(torch.sum(dom2[:, 0] - x, dim=1)) + torch.exp(torch.sum(x - dom2[:, 1], dim=1))

    ref2 = 2.0 - 4.0 * math.exp(-c) + 2.0 * math.exp(-2 * c)
    for seed in [0, 1, 2, 3, 41317]:
        r = tq.VEGAS().integrate(diag, 2, N=30000, integration_domain=dom2, seed=seed)
        assert abs(float(r) - ref2) < 0.03


# ---------------------------------------------------------------- Monte Carlo
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_monte_carlo_matches_oracle_on_same_samples(cuda, tag):
    dt = DT[tag]
    dom = torch.tensor([[0.0, 2.0], [-1.0, 1.5], [3.0, 3.5]], dtype=dt, device=cuda)
    mc = tq.MonteCarlo()
    fn = lambda x: torch.sum(torch.sin(x), dim=1)  # noqa: E731
    res = mc.integrate(fn, 3, N=200_000, integration_domain=dom, seed=42)
    u = O.philox_uniform(42, 0, 0, 200_000, 3, dt)
    pts = O.mc_sample_points(u, dom.cpu())
    want = O.mc_result(fn(pts), dom.cpu())
    assert res.dtype == dt and res.dim() == 0 and mc._nr_of_fevals == 200_000
    assert abs(float(res) - float(want)) <= (1e-12 if tag == "f64" else 1e-5) * abs(float(want))
    # fused functor on the same stream
    fres = mc.integrate(F.SumOfSines(3), 3, N=200_000, integration_domain=dom, seed=42)
    assert abs(float(fres) - float(want)) <= (1e-12 if tag == "f64" else 1e-5) * abs(float(want))
    assert mc.get_error_estimate() > 0
    # vector-valued integrand
    vres = mc.integrate(lambda x: torch.stack([fn(x), 2 * fn(x)], dim=1), 3, N=200_000, integration_domain=dom, seed=42)
    assert vres.shape == (2,) and abs(float(vres[1]) - 2 * float(want)) <= 1e-5 * abs(float(want))


def test_monte_carlo_constant_is_exact_and_chunking_is_invisible(cuda):
    """Order-0 polynomials integrate with zero error (/root/reference/tests/monte_carlo_test.py:29-30)."""
    dom = torch.tensor([[0.0, 2.0]], dtype=torch.float32, device=cuda)
    mc = tq.MonteCarlo()
    r = mc.integrate(lambda x: x[:, 0] * 0 + 2.0, 1, N=100_000, integration_domain=dom, seed=0)
    assert float(r) == 4.0
    fn = lambda x: torch.sum(torch.exp(x), dim=1)  # noqa: E731
    dom3 = torch.tensor([[0.0, 1.0]] * 3, dtype=torch.float64, device=cuda)
    whole = tq.MonteCarlo().integrate(fn, 3, N=300_001, integration_domain=dom3, seed=5)
    chunked = tq.MonteCarlo()
    chunked.max_points_bytes = 100_000 * 3 * 8
    part = chunked.integrate(fn, 3, N=300_001, integration_domain=dom3, seed=5)
    assert abs(float(whole) - float(part)) <= 1e-13 * abs(float(whole))
    assert chunked._nr_of_fevals == 300_001


def test_monte_carlo_list_domain_and_types(cuda):
    """Result dtype/device follow the domain (/root/reference/tests/integrator_types_test.py:19-113)."""
    tq.set_up_backend("torch", "float64")
    try:
        for integ, kw in [(tq.MonteCarlo(), dict(N=1000, seed=0)), (tq.VEGAS(), dict(N=2000, seed=0)),
                          (tq.Trapezoid(), dict(N=100)), (tq.Simpson(), dict(N=121)), (tq.Boole(), dict(N=169))]:
            seen = {}

            def fn(x):
                seen["dtype"], seen["shape"] = x.dtype, x.shape
                return x[:, 0] * 0.0 - 1.0

            r = integ.integrate(fn, 2, integration_domain=[[0, 2], [0, 2]], backend="torch", **kw)
            assert r.dtype == torch.float64 and r.is_cuda and seen["dtype"] == torch.float64 and seen["shape"][1] == 2
            assert abs(float(r) + 4.0) < (0.03 if isinstance(integ, tq.VEGAS) else 1e-5)
            r32 = integ.integrate(fn, 2, integration_domain=torch.tensor([[0, 2], [0, 2]], dtype=torch.float32, device=cuda), **kw)
            assert r32.dtype == torch.float32 and seen["dtype"] == torch.float32
    finally:
        torch.set_default_dtype(torch.float32)
        torch.set_default_device("cpu")


# ---------------------------------------------------------------- Newton-Cotes exactness pins
def test_newton_cotes_polynomial_exactness(cuda):
    """/root/reference/tests/{trapezoid,simpson,boole}_test.py: rules are exact up to degree 1/3/5."""
    dt = torch.float64
    dom1 = torch.tensor([[0.0, 2.0]], dtype=dt, device=cuda)
    lin = lambda x: 3.0 * x[:, 0] + 1.0  # noqa: E731
    assert abs(float(tq.Trapezoid().integrate(lin, 1, N=2, integration_domain=dom1)) - 8.0) < 1e-15
    cub = lambda x: x[:, 0] ** 3 - x[:, 0] + 2.0  # noqa: E731
    assert abs(float(tq.Simpson().integrate(cub, 1, N=3, integration_domain=dom1)) - 6.0) < 1e-14
    quint = lambda x: x[:, 0] ** 5 + x[:, 0] ** 2  # noqa: E731
    assert abs(float(tq.Boole().integrate(quint, 1, N=401, integration_domain=dom1)) - (64 / 6 + 8 / 3)) < 6.33e-11
    dom3 = torch.tensor([[0.0, 1.0], [-1.0, 1.0], [0.0, 2.0]], dtype=dt, device=cuda)
    f3 = lambda x: torch.sum(x**5, dim=1) + torch.prod(x, dim=1)  # noqa: E731
    exact = (1 / 6) * 4 + 0 + (64 / 6) * 2 + 0.0
    r = tq.Boole().integrate(f3, 3, N=1_076_890, integration_domain=dom3)  # adjusts 102 -> 101 per dim
    assert abs(float(r) - exact) < 2e-11
    with pytest.warns(UserWarning):
        tq.Simpson().integrate(cub, 1, N=4, integration_domain=dom1)
    # 10-D Simpson on 3^10 points (simpson_test.py:78-88)
    dom10 = torch.tensor([[0.0, 1.0]] * 10, dtype=dt, device=cuda)
    r10 = tq.Simpson().integrate(lambda x: torch.sum(x**2, dim=1), 10, N=3**10, integration_domain=dom10)
    assert abs(float(r10) - 10 / 3) < 5e-9


def test_gauss_legendre_matches_reference_fixture_and_exactness(cuda, golden):
    """GaussLegendre against the reference fixture; exactness up to degree 2n-1
    (/root/reference/tests/gauss_test.py:9-49)."""
    g = golden("gauss_legendre_f64")
    fn = lambda x: torch.prod(torch.cos(x), dim=1) + torch.sum(x**3, dim=1)  # noqa: E731
    for dim in (1, 2, 3):
        k = f"d{dim}"
        dom = torch.from_numpy(g[f"{k}_domain"]).to(cuda)
        gl = tq.GaussLegendre()
        pts, hs, n = gl.calculate_grid(int(g[f"{k}_N"]), dom)
        assert n == int(g[f"{k}_n"])
        assert torch.allclose(pts.cpu(), torch.from_numpy(g[f"{k}_points"]), rtol=0, atol=1e-15)
        assert torch.allclose(gl._weights(n, dim, device=cuda).cpu(), torch.from_numpy(g[f"{k}_W"]), rtol=1e-14)
        r = gl.integrate(fn, dim, int(g[f"{k}_N"]), dom)
        assert r.dtype == torch.float64 and abs(float(r) - float(g[f"{k}_result"])) <= 1e-13 * max(1, abs(float(g[f"{k}_result"])))
        rv = gl.integrate(lambda x: torch.stack([fn(x), torch.sum(torch.exp(x), dim=1)], dim=1), dim, int(g[f"{k}_N"]), dom)
        assert torch.allclose(rv.cpu(), torch.from_numpy(g[f"{k}_resultv"]), rtol=1e-13)
        # reference-style split-phase use: weights applied to the values, then calculate_result
        vals, _ = gl.evaluate_integrand(fn, pts, weights=gl._weights(n, dim, device=cuda))
        r2 = gl.calculate_result(vals, dim, n, hs, dom)
        assert abs(float(r2) - float(r)) <= 1e-13 * max(1, abs(float(r)))
    dom1 = torch.tensor([[0.0, 2.0]], dtype=torch.float64, device=cuda)
    assert abs(float(tq.GaussLegendre().integrate(lambda x: x[:, 0] ** 3 - x[:, 0] + 2.0, 1, 2, dom1)) - 6.0) < 8e-15
    assert abs(float(tq.GaussLegendre().integrate(lambda x: torch.sin(x[:, 0]), 1, 60, torch.tensor([[0.0, 5.0]], dtype=torch.float64, device=cuda)))
               - (1 - math.cos(5.0))) < 7e-11
    # fused functor and multi-chunk paths
    dom4 = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=cuda)
    pc = F.ProductOfCosines(4)
    a = tq.GaussLegendre().integrate(pc, 4, 12**4, dom4)
    b = tq.GaussLegendre().integrate(lambda x: pc(x), 4, 12**4, dom4)
    assert abs(float(a) - pc.exact()) < 1e-14 and abs(float(b) - pc.exact()) < 1e-14
    base = tq.Gaussian().integrate(lambda x: x[:, 0] ** 2, 1, 8, None)  # parent class integrates on [-1, 1]
    assert abs(float(base) - 2.0 / 3.0) < 1e-6


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_newton_cotes_fused_and_chunked_equal_plain(cuda, tag):
    dt = DT[tag]
    dom = torch.tensor([[0.0, 1.0], [0.5, 2.0], [-1.0, 0.0], [0.0, 1.0]], dtype=dt, device=cuda)
    fn = F.ProductOfCosines(4)
    for cls, N in [(tq.Trapezoid, 20**4), (tq.Simpson, 21**4), (tq.Boole, 21**4)]:
        plain = cls().integrate(lambda x: fn(x), 4, N=N, integration_domain=dom)
        fused = cls().integrate(fn, 4, N=N, integration_domain=dom)
        ch = cls()
        ch.max_points_bytes = 50_000 * 4 * dom.element_size()
        chunked = ch.integrate(lambda x: fn(x), 4, N=N, integration_domain=dom)
        tol = 1e-12 if tag == "f64" else 2e-5
        assert abs(float(fused) - float(plain)) <= tol * abs(float(plain))
        assert abs(float(chunked) - float(plain)) <= tol * abs(float(plain))
        exact = math.sin(1.0) * (math.sin(2.0) - math.sin(0.5)) * math.sin(1.0) * math.sin(1.0)
        assert abs(float(plain) - exact) < 2e-3


# ---------------------------------------------------------------- fused integrand families
@pytest.mark.parametrize("cls,kw", [
    (F.GenzOscillatory, dict(a=[0.5, 0.7, 0.2, 0.9, 0.4], u=[0.3] * 5)),
    (F.GenzProductPeak, dict(a=[2.0, 1.5, 3.0, 1.0, 2.5], u=[0.5, 0.4, 0.6, 0.3, 0.7])),
    (F.GenzCornerPeak, dict(a=[0.5, 0.3, 0.2, 0.6, 0.1])),
    (F.GenzGaussian, dict(a=[2.0, 3.0, 1.0, 2.5, 1.5], u=[0.5, 0.4, 0.6, 0.3, 0.7])),
    (F.GenzC0, dict(a=[2.0, 1.0, 3.0, 1.5, 0.5], u=[0.5, 0.4, 0.6, 0.3, 0.7])),
    (F.GenzDiscontinuous, dict(a=[0.5, 0.3, 0.2, 0.6, 0.1], u=[0.6, 0.7, 0.5, 0.5, 0.5])),
    (F.SumOfSines, {}), (F.SumOfExp, {}), (F.ProductOfCosines, {}),
])
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_fused_integrands_match_torch_formulation_and_exact(cuda, cls, kw, tag):
    dt = DT[tag]
    fn = cls(5, **kw)
    dom = torch.tensor([[0.0, 1.0]] * 5, dtype=dt, device=cuda)
    N = 400_000
    fused = tq.MonteCarlo()
    rf = fused.integrate(fn, 5, N=N, integration_domain=dom, seed=9)
    ru = tq.MonteCarlo().integrate(lambda x: fn(x), 5, N=N, integration_domain=dom, seed=9)
    tol = 1e-11 if tag == "f64" else 2e-5
    assert abs(float(rf) - float(ru)) <= tol * abs(float(ru)), "fused functor != torch formulation on identical samples"
    err = fused.get_error_estimate()
    assert abs(float(rf) - fn.exact()) <= 5 * err + 1e-6 * abs(fn.exact())
    # oracle formulation agrees with the package's torch formulation
    x = torch.rand(100, 5, dtype=torch.float64)
    if cls.family.startswith("genz"):
        assert torch.allclose(fn(x), O.genz(cls.family[5:], x, fn.a, fn.u), rtol=1e-13)
        assert abs(fn.exact() - O.genz_exact(cls.family[5:], fn.a, fn.u)) <= 1e-13 * abs(fn.exact())


def test_fast_math_variants_stay_within_fp32_tolerance(cuda):
    """Opt-in SFU evaluation of sin/cos/exp (fp32): the estimate moves by far less than the 1e-5 bar."""
    dom = torch.tensor([[0.0, 1.0]] * 6, dtype=torch.float32, device=cuda)
    for cls in (F.SumOfSines, F.SumOfExp, F.ProductOfCosines):
        a = float(tq.MonteCarlo().integrate(cls(6), 6, N=2_000_000, integration_domain=dom, seed=4))
        b = float(tq.MonteCarlo().integrate(cls(6, fast_math=True), 6, N=2_000_000, integration_domain=dom, seed=4))
        assert abs(a - b) <= 2e-6 * abs(a), (cls.__name__, a, b)
    d64 = dom.double()
    a = float(tq.MonteCarlo().integrate(F.SumOfSines(6), 6, N=100_000, integration_domain=d64, seed=4))
    b = float(tq.MonteCarlo().integrate(F.SumOfSines(6, fast_math=True), 6, N=100_000, integration_domain=d64, seed=4))
    assert a == b  # fp64 has no fast variant


def test_fused_polynomial(cuda):
    fn = F.Polynomial(3, [1.0, -2.0, 0.5, 3.0])
    dom = torch.tensor([[0.0, 1.0]] * 3, dtype=torch.float64, device=cuda)
    r = tq.Boole().integrate(fn, 3, N=9**3, integration_domain=dom)
    assert abs(float(r) - fn.exact()) < 1e-13
    x = torch.rand(50, 3, dtype=torch.float64)
    assert torch.allclose(fn(x), O.test_integrand("polynomial", x, fn.coeffs), rtol=1e-13)


# ---------------------------------------------------------------- autograd (reference tests/gradient_test.py)
@pytest.mark.parametrize("make,kw,tol", [
    (tq.Trapezoid, dict(N=100001), 2e-2), (tq.Simpson, dict(N=100001), 2e-2), (tq.Boole, dict(N=100001), 2e-2),
    (tq.MonteCarlo, dict(N=1_000_000, seed=0), 2e-2), (tq.VEGAS, dict(N=200_000, seed=0), 2e-2),
])
def test_gradient_wrt_domain(cuda, make, kw, tol):
    """d/d(domain) of int 2|x| over [-1,1] is [-2, 2] (gradient_test.py:187-215)."""
    dom = torch.tensor([[-1.0, 1.0]], dtype=torch.float64, device=cuda, requires_grad=True)
    res = make().integrate(lambda x: 2.0 * torch.abs(x[:, 0]), 1, integration_domain=dom, **kw)
    assert res.grad_fn is not None and res.dtype == torch.float64
    res.backward()
    g = dom.grad.cpu().numpy().ravel()
    assert abs(g[0] + 2.0) < tol and abs(g[1] - 2.0) < tol
    assert abs(float(res) - 2.0) < 1e-2


@pytest.mark.parametrize("make,kw,tol", [
    (tq.Trapezoid, dict(N=10201), 0.1), (tq.Simpson, dict(N=10201), 0.1), (tq.Boole, dict(N=10201), 0.1),
    (tq.MonteCarlo, dict(N=200_000, seed=0), 0.1), (tq.VEGAS, dict(N=100_000, seed=0), 0.1),
])
def test_gradient_wrt_integrand_parameters(cuda, make, kw, tol):
    """2-D polynomial with learnable coefficients (gradient_test.py:217-259): dI/dc_k = 2*int x^k over [0,1]^2... """
    c = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device=cuda, requires_grad=True)
    dom = torch.tensor([[0.0, 1.0], [0.0, 1.0]], dtype=torch.float64, device=cuda)

    def fn(x):
        return torch.sum(c[0] + c[1] * x + c[2] * x**2, dim=1)

    res = make().integrate(fn, 2, integration_domain=dom, **kw)
    res.backward()
    want = np.array([2.0, 1.0, 2.0 / 3.0])
    assert np.all(np.abs(c.grad.cpu().numpy() - want) < tol)
    assert abs(float(res) - float((c.detach().cpu().numpy() * want).sum())) < 5e-2


# ---------------------------------------------------------------- RNG surface (reference tests/rng_test.py)
def test_rng_surface(cuda):
    a = tq.RNG(backend="torch", seed=547).uniform(size=[3, 9], dtype=torch.float32, device=cuda)
    b = tq.RNG(backend="torch", seed=547).uniform(size=[3, 9], dtype=torch.float32, device=cuda)
    c = tq.RNG(backend="torch", seed=548).uniform(size=[3, 9], dtype=torch.float32, device=cuda)
    d = tq.RNG(backend="torch").uniform(size=[3, 9], dtype=torch.float32, device=cuda)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)
    r = tq.RNG(backend="torch", seed=1)
    x1, x2 = r.uniform([5], torch.float64, device=cuda), r.uniform([5], torch.float64, device=cuda)
    assert x1.shape == (5,) and not torch.equal(x1, x2)
    assert r.uniform([0], torch.float64, device=cuda).shape == (0,)
    big = r.uniform([100000], torch.float32, device=cuda)
    assert 0.0 <= float(big.min()) and float(big.max()) < 1.0 and abs(float(big.mean()) - 0.5) < 0.01
    with pytest.raises(ValueError):
        tq.RNG(backend="numpy")


def test_compiled_integrate_replays_a_cuda_graph(cuda):
    """get_jit_compiled_integrate (monte_carlo.py:108-225, grid_integrator.py:134-255): the whole call is captured
    once per integrand and replayed; replays draw fresh Philox calls and follow the domain passed in."""
    from torchquad_b200.integration.rng import RNG

    torch.set_default_dtype(torch.float64)
    dom = torch.tensor([[0.0, 1.0], [0.0, 2.0], [-1.0, 1.0]], dtype=torch.float64, device=cuda)

    def fn(x):
        return torch.exp(-(x * x).sum(dim=1)) + x[:, 1]

    def exact(d):
        import math
        d = d.tolist()
        vol = 1.0
        g = 1.0
        for a, b in d:
            vol *= b - a
            g *= math.sqrt(math.pi) / 2 * (math.erf(b) - math.erf(a))
        lin = vol * (d[1][0] + d[1][1]) / 2
        return g + lin

    mc = tq.MonteCarlo()
    compiled = mc.get_jit_compiled_integrate(dim=3, N=200_000, integration_domain=dom, seed=5, capture_integrand=True)
    r = [compiled(fn, dom) for _ in range(3)]
    assert compiled.replays == 3
    assert r[0].dtype == torch.float64 and r[0].is_cuda and r[0].dim() == 0
    vals = [float(x) for x in r]
    assert len(set(vals)) == 3  # fresh samples per replay
    for v in vals:
        assert abs(v - exact(dom)) < 0.02 * exact(dom)
    # every step of a compiled function uses Philox call = its device counter: warm-up 0, dry run 1, first replay 2
    rng = RNG(seed=5)
    rng._call = 2
    assert float(tq.MonteCarlo().integrate(fn, 3, N=200_000, integration_domain=dom, rng=rng)) == vals[0]
    dom2 = torch.tensor([[0.0, 0.5], [1.0, 2.0], [0.0, 1.0]], dtype=torch.float64, device=cuda)
    assert abs(float(compiled(fn, dom2)) - exact(dom2)) < 0.02 * exact(dom2)
    assert abs(float(compiled(fn, dom2.tolist())) - exact(dom2)) < 0.02 * exact(dom2)  # list domains work too

    for rule, N in ((tq.Simpson(), 41**3), (tq.Boole(), 41**3), (tq.Trapezoid(), 101**3)):
        c = rule.get_jit_compiled_integrate(dim=3, N=N, integration_domain=dom, capture_integrand=True)
        for d in (dom, dom2, dom):
            got, want = float(c(fn, d)), float(rule.integrate(fn, 3, N=N, integration_domain=d))
            assert abs(got - want) <= 1e-12 * abs(want)
        assert c.replays == 3

    # an integrand with a host read-back cannot be captured: the call still works (eagerly)
    def syncing(x):
        return torch.exp(-(x * x).sum(dim=1)) * float(x[0, 0] * 0 + 1)

    c = tq.Simpson().get_jit_compiled_integrate(dim=3, N=21**3, integration_domain=dom, capture_integrand=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = float(c(syncing, dom))
        b = float(c(syncing, dom))
    assert c.replays == 0 and a == b
    torch.rand(4, device=cuda)  # torch's generator is usable after the refused capture
    assert abs(a - float(tq.Simpson().integrate(lambda x: torch.exp(-(x * x).sum(dim=1)), 3, N=21**3, integration_domain=dom))) < 1e-12
    # a built-in (fused) integrand is a single launch already and runs eagerly
    unit = torch.tensor([[0.0, 1.0]] * 3, dtype=torch.float64, device=cuda)
    c = tq.MonteCarlo().get_jit_compiled_integrate(dim=3, N=100_000, integration_domain=unit, seed=1, capture_integrand=True)
    assert abs(float(c(F.SumOfSines(3), unit)) - F.SumOfSines(3).exact()) < 0.02 and c.replays == 0


def test_compiled_integrate_default_is_eager_and_differentiable(cuda):
    """The reference evaluates the integrand eagerly inside compiled_integrate (monte_carlo.py:197-212): current Python
    state is seen on every call and gradients flow to the domain and to integrand parameters; with capture_integrand=True
    integrands whose values require grad are kept eager, bound methods share one graph, the graph cache is bounded."""
    torch.set_default_dtype(torch.float64)
    dom = torch.tensor([[0.0, 1.0], [0.0, 2.0]], dtype=torch.float64, device=cuda)
    scale = {"a": 1.0}

    def fn(x):
        return scale["a"] * (x[:, 0] + x[:, 1])

    for integ, kw in ((tq.MonteCarlo(), dict(N=100_000, seed=3)), (tq.Simpson(), dict(N=21**2))):
        c = integ.get_jit_compiled_integrate(dim=2, integration_domain=dom, **kw)
        scale["a"] = 1.0
        one = float(c(fn, dom))
        scale["a"] = 3.0
        three = float(c(fn, dom))
        assert c.replays == 0 and abs(three / one - 3.0) < 0.05  # the closure value is read on every call
        # gradient wrt the domain (gradient_test.py:162-259) through the compiled callable
        d = dom.clone().requires_grad_(True)
        scale["a"] = 1.0
        c(fn, d).backward()
        assert d.grad is not None and torch.isfinite(d.grad).all() and float(d.grad.abs().sum()) > 0
    # gradient wrt integrand parameters, default mode and capture mode (must stay eager)
    for capture in (False, True):
        p = torch.tensor(2.0, dtype=torch.float64, device=cuda, requires_grad=True)
        c = tq.Simpson().get_jit_compiled_integrate(dim=2, N=21**2, integration_domain=dom, capture_integrand=capture)
        r = c(lambda x: p * x[:, 0] * x[:, 1], dom)
        assert r.requires_grad and c.replays == 0
        r.backward()
        assert abs(float(p.grad) - 1.0) < 1e-9  # int x y over [0,1]x[0,2] = 1
    # Monte Carlo with a grad domain after plain calls: the device-side call counter is read back for the backward
    c = tq.MonteCarlo().get_jit_compiled_integrate(dim=2, N=50_000, integration_domain=dom, seed=1, capture_integrand=True)
    plain = lambda x: x[:, 0] + x[:, 1]  # noqa: E731
    assert abs(float(c(plain, dom)) - 3.0) < 0.1 and c.replays == 1
    d = dom.clone().requires_grad_(True)
    c(plain, d).backward()
    assert torch.isfinite(d.grad).all()

    class Model:
        def __init__(self, k):
            self.k = k

        def f(self, x):
            return self.k * x[:, 0]

    m = Model(2.0)
    c = tq.Trapezoid().get_jit_compiled_integrate(dim=2, N=11**2, integration_domain=dom, capture_integrand=True)
    vals = [float(c(m.f, dom)) for _ in range(3)]  # m.f is a new bound-method object each time: one graph
    assert c.replays == 3 and len(c._entries) == 1 and abs(vals[0] - 2.0) < 1e-9
    for k in range(20):  # a fresh lambda per call: capturing stops after a few misses, memory stays bounded
        c(lambda x, k=k: x[:, 0] * (k + 1), dom)
    assert len(c._entries) <= c.max_graphs and c._misses >= c.max_consecutive_misses


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("dim,ns,ni,g", [(3, 6, 512, 1), (4, 2, 100, 2), (5, 3, 257, 1), (7, 4, 64, 2), (2, 9, 1000, 1), (6, 3, 50, 4)])
def test_deferred_pass_and_band_sweep_match_direct_histogram(cuda, tag, dim, ns, ni, g):
    """Maps beyond L2 (tq_fused_vegas_deferred + tq_vegas_hist_sweep): the pass stores jf^2 per row, the sweep regenerates the
    uniforms and bins the rows band by band.  Must reproduce the direct histogram of the same pass: counts exactly, weights
    to rounding, per-cube sums bit for bit."""
    from torchquad_b200.integration.vegas_map import VEGASMap

    dt = DT[tag]
    if g > (2 if tag == "f64" else 4):
        pytest.skip("dims_per_group beyond one Philox block")
    torch.manual_seed(dim * 100 + ns)
    vm = VEGASMap(ni, dim, "torch", dt, device=cuda)
    vm.x_edges[:, 1:-1] += (torch.rand(dim, ni - 1, device=cuda, dtype=dt) - 0.5) * (0.5 / ni)
    vm.dx_edges = (vm.x_edges[:, 1:] - vm.x_edges[:, :-1]).contiguous()
    fn = F.GenzOscillatory(dim, a=0.7, u=0.2)
    s = fn.to_struct([0.0] * dim, [1.0] * dim, 1.0)
    C = ns**dim
    nh = torch.randint(2, 9, (C,), device=cuda, dtype=torch.int64)
    nh[C // 3] = 700  # one heavy cube (spans several warps of the pass, many steps of the sweep)
    offsets = ops.strat_offsets(nh)
    M = int(offsets[-1])
    edges = ops.pack_edges(vm.x_edges, vm.dx_edges)
    w1, c1 = torch.zeros_like(vm.weights), torch.zeros_like(vm.counts)
    JF1 = torch.zeros((2, C), dtype=dt, device=cuda)
    ops.fused_vegas(s, edges, w1, c1, 0, M, 1, 7, offsets=offsets, n_strat=ns, JF=JF1[0], JF2=JF1[1])
    JF2 = torch.zeros((2, C), dtype=dt, device=cuda)
    jf2_rows = torch.full((M,), float("nan"), dtype=dt, device=cuda)
    ops.fused_vegas_deferred(s, edges, 0, M, 1, 7, offsets, ns, JF2[0], JF2[1], jf2_rows)
    hist = torch.zeros((dim, ni, 2), dtype=torch.float64, device=cuda)
    ops.hist_sweep(offsets, ns, dim, jf2_rows, ni, hist, g, 1, 7)
    assert torch.isfinite(jf2_rows).all()
    assert torch.equal(hist[..., 1].to(torch.int64), c1) and int(c1.sum()) == dim * M
    assert rel_err_t(hist[..., 0].to(dt), w1) <= (1e-12 if tag == "f64" else 2e-5)
    assert torch.allclose(JF1, JF2, rtol=1e-12 if tag == "f64" else 1e-5)


def rel_err_t(a, b):
    a, b = a.double(), b.double()
    return float(((a - b).abs() / b.abs().clamp_min(1e-300)).max())


def test_vegas_band_sweep_run_matches_default_run(cuda, monkeypatch):
    """A whole fused run with the large-map path forced (pair tables + deferred passes + band sweeps, csrc/vegas_driver.cu)
    must reproduce the default run: same schedule, same evaluation count, same map counts, same result."""
    from torchquad_b200.integration.vegas_map import VEGASMap

    fn = F.GenzGaussian(4, a=5.0, u=0.5)
    dom = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=cuda)

    def run():
        v = tq.VEGAS()
        return v, v.integrate(fn, 4, N=30_000_000, integration_domain=dom, seed=3)

    a, ra = run()
    monkeypatch.setattr(VEGASMap, "records_min_bytes", 0)
    b, rb = run()
    assert b.map.sweep_group(b.strat.N_strat) == 1 and b.map._records is None and b.map._hist is not None
    assert a.it == b.it and a._nr_of_fevals == b._nr_of_fevals
    assert abs(float(ra) - float(rb)) <= 1e-9 * abs(float(ra))
    assert float((a.map.x_edges - b.map.x_edges).abs().max()) <= 1e-9
    assert torch.equal(a.map.counts, b.map.counts)


@pytest.mark.parametrize("native", [True, False])
def test_vegas_record_layout_matches_pair_layout(cuda, monkeypatch, native):
    """Large maps keep {x, dx, weight, count} records (TQ_EDGES_RECORDS); forcing that layout on a small problem
    must reproduce the default layout's run: same samples, same histogram, same map."""
    from torchquad_b200.integration.vegas_map import VEGASMap

    fn = F.GenzGaussian(4, a=5.0, u=0.5)
    dom = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=cuda)

    def run():
        v = tq.VEGAS()
        v.native_loop = native
        return v, v.integrate(fn, 4, N=400_000, integration_domain=dom, seed=3)

    a, ra = run()
    monkeypatch.setattr(VEGASMap, "records_min_bytes", 0)
    b, rb = run()
    assert b.map._records is not None and a.map._records is None
    assert a.it == b.it and a._nr_of_fevals == b._nr_of_fevals
    assert abs(float(ra) - float(rb)) <= 1e-9 * abs(float(ra))
    assert float((a.map.x_edges - b.map.x_edges).abs().max()) <= 1e-9
    assert torch.equal(a.map.counts, b.map.counts)
    # kernel level: one stratified pass into records, unpacked, against the same pass into the arrays
    from torchquad_b200 import ops

    for dt in (torch.float64, torch.float32):
        vm = VEGASMap(512, 3, "torch", dt, device=cuda)
        vm.x_edges[:, 1:-1] += (torch.rand(3, 511, device=cuda, dtype=dt) - 0.5) * 1e-3
        vm.dx_edges = (vm.x_edges[:, 1:] - vm.x_edges[:, :-1]).contiguous()
        f3 = F.GenzOscillatory(3, a=0.7, u=0.2)
        s = f3.to_struct([0.0] * 3, [1.0] * 3, 1.0)
        nh = torch.randint(2, 9, (6**3,), device=cuda, dtype=torch.int64)
        offsets = ops.strat_offsets(nh)
        M = int(offsets[-1])
        w1, c1 = torch.zeros_like(vm.weights), torch.zeros_like(vm.counts)
        JF1 = torch.zeros((2, 6**3), dtype=dt, device=cuda)
        ops.fused_vegas(s, ops.pack_edges(vm.x_edges, vm.dx_edges), w1, c1, 0, M, 1, 7, offsets=offsets, n_strat=6,
                        JF=JF1[0], JF2=JF1[1])
        rec = ops.pack_records(vm.x_edges, vm.dx_edges)
        JF2 = torch.zeros((2, 6**3), dtype=dt, device=cuda)
        ops.fused_vegas(s, None, None, None, 0, M, 1, 7, offsets=offsets, n_strat=6, JF=JF2[0], JF2=JF2[1], records=rec,
                        dtype=dt, n_intervals=512)
        w2, c2 = torch.zeros_like(vm.weights), torch.zeros_like(vm.counts)
        ops.unpack_records(rec, w2, c2)
        assert torch.equal(c1, c2) and int(c1.sum()) == 3 * M
        assert torch.allclose(w1, w2, rtol=1e-5 if dt == torch.float32 else 1e-12, atol=0)
        assert torch.allclose(JF1, JF2, rtol=1e-5 if dt == torch.float32 else 1e-12, atol=0)
        w3, c3 = torch.zeros_like(vm.weights), torch.zeros_like(vm.counts)
        ops.unpack_records(rec, w3, c3)  # the records were zeroed by the first unpack
        assert int(c3.sum()) == 0 and float(w3.abs().sum()) == 0.0


@pytest.mark.parametrize("fused", [True, False])
def test_vegas_adaptation_state_round_trip(cuda, fused, tmp_path):
    """`adaptation_state()` / `initial_adaptation`: a run that starts from a saved map and stratification skips the
    warm-up, begins with exactly the saved tables and is at least as accurate as a cold run."""
    g = F.GenzGaussian(4, a=6.0, u=0.4)
    fn = g if fused else (lambda x: g(x))
    dom = torch.tensor([[0.0, 1.0]] * 4, dtype=torch.float64, device=cuda)
    cold = tq.VEGAS()
    r0 = cold.integrate(fn, 4, N=400_000, integration_domain=dom, seed=1)
    state = cold.adaptation_state()
    torch.save(state, tmp_path / "vegas_state.pt")
    state = torch.load(tmp_path / "vegas_state.pt")
    assert torch.equal(state["x_edges"], cold.map.x_edges.cpu()) and torch.equal(state["dh"], cold.strat.dh.cpu())

    warm = tq.VEGAS()
    warm.initial_adaptation = state
    r1 = warm.integrate(fn, 4, N=400_000, integration_domain=dom, seed=2, use_grid_improve=False)
    # no warm-up evaluations, and without grid improvement the map is still the saved one
    assert torch.equal(warm.map.x_edges.cpu(), state["x_edges"])
    assert abs(float(r1) - g.exact()) <= 5 * float(warm._get_error())
    fresh = tq.VEGAS()
    fresh.integrate(fn, 4, N=400_000, integration_domain=dom, seed=2, use_grid_improve=False, use_warmup=False)
    assert float(warm._get_error()) < float(fresh._get_error())  # the saved adaptation pays from the first iteration
    assert abs(float(r0) - g.exact()) <= 5 * float(cold._get_error())

    other = tq.VEGAS()
    other.initial_adaptation = state
    with pytest.raises(ValueError):
        other.integrate(fn, 4, N=100_000, integration_domain=dom, seed=2)  # different table shapes


def test_vegas_unfused_record_layout_matches_pair_layout(cuda, monkeypatch):
    """Torch-callable integrand on a (forced) large-map record table: the gather reads the pairs out of the records,
    the fused tail accumulates into them; the run must reproduce the default layout."""
    from torchquad_b200.integration.vegas_map import VEGASMap

    g = F.GenzGaussian(4, a=5.0, u=0.5)
    for dt, tol in ((torch.float64, 1e-9), (torch.float32, None)):
        dom = torch.tensor([[0.0, 1.0]] * 4, dtype=dt, device=cuda)

        def run():
            v = tq.VEGAS()
            return v, v.integrate(lambda x: g(x), 4, N=300_000, integration_domain=dom, seed=3)

        monkeypatch.setattr(VEGASMap, "records_min_bytes", 48 << 20)
        a, ra = run()
        monkeypatch.setattr(VEGASMap, "records_min_bytes", 0)
        b, rb = run()
        assert b.map._records is not None and a.map._records is None
        assert a.it == b.it
        if tol is not None:
            assert a._nr_of_fevals == b._nr_of_fevals
            assert abs(float(ra) - float(rb)) <= tol * abs(float(ra))
            assert float((a.map.x_edges - b.map.x_edges).abs().max()) <= 1e-9
        else:  # fp32 runs agree statistically (see test_vegas_native_loop_equals_python_loop)
            assert abs(float(ra) - float(rb)) <= 2.5 * float(a._get_error())


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_vegas_native_unfused_loop_equals_python_loop(cuda, tag):
    """Python-callable integrand: the C++-driven loop (tq_vegas_run_unfused, integrand evaluated through a callback)
    against the Python-driven loop on the same seed -- same samples, same schedule."""
    dt = DT[tag]
    g = F.GenzGaussian(4, a=5.0, u=0.5)
    calls = []

    def fn(x):
        calls.append(x.shape[0])
        return g(x)

    dom = torch.tensor([[0.0, 1.0], [0.0, 1.0], [-0.5, 1.0], [0.0, 2.0]], dtype=dt, device=cuda)
    runs = {}
    for native in (True, False):
        calls.clear()
        v = tq.VEGAS()
        v.native_loop = native
        r = v.integrate(fn, 4, N=300_000, integration_domain=dom, seed=11)
        runs[native] = (v, r, list(calls))
    (a, ra, ca), (b, rb, cb_) = runs[True], runs[False]
    assert a.it == b.it and a.rng._call == b.rng._call and len(ca) == len(cb_) == 5 + a.it
    assert ra.dtype == rb.dtype == dt and ra.is_cuda
    if tag == "f64":
        assert ca == cb_ and a._nr_of_fevals == b._nr_of_fevals == sum(ca)
        assert abs(float(ra) - float(rb)) <= 1e-9 * abs(float(rb))
        assert float((a.map.x_edges - b.map.x_edges).abs().max()) <= 1e-9
        assert torch.equal(a.strat._nh, b.strat._nh)
    else:  # fp32: see test_vegas_native_loop_equals_python_loop
        assert abs(a._nr_of_fevals - b._nr_of_fevals) <= 2e-3 * b._nr_of_fevals
        assert abs(float(ra) - float(rb)) <= 2.5 * float(b._get_error())

    # exceptions raised by the integrand propagate out of the C++ loop
    class Boom(Exception):
        pass

    def bad(x):
        if len(calls) >= 3:
            raise Boom("third pass")
        calls.append(0)
        return g(x)

    calls.clear()
    with pytest.raises(Boom):
        tq.VEGAS().integrate(bad, 4, N=300_000, integration_domain=dom, seed=1)
    # wrong output shape -> the reference's ValueError
    with pytest.raises(ValueError):
        tq.VEGAS().integrate(lambda x: g(x)[:-1], 4, N=300_000, integration_domain=dom, seed=1)
    # an integrand whose values require grad falls back to the differentiable loop
    w = torch.tensor(1.5, dtype=dt, device=cuda, requires_grad=True)
    r = tq.VEGAS().integrate(lambda x: w * g(x), 4, N=300_000, integration_domain=dom, seed=11)
    r.backward()
    assert r.requires_grad and abs(float(w.grad) - float(r) / 1.5) <= 1e-3 * abs(float(r))
