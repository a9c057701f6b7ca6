"""Edge cases of the drop-in surface: minimal sizes, odd shapes, large dimension counts, ragged tails."""
import math
import warnings

import pytest
import torch

import torchquad_b200 as tq
from oracle import ref_oracle as O
from torchquad_b200 import integrands as F
from torchquad_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_minimal_sizes(cuda, dt):
    dom = torch.tensor([[0.0, 2.0]], dtype=dt, device=cuda)
    one = tq.MonteCarlo().integrate(lambda x: x[:, 0] * 0 + 3.0, 1, N=1, integration_domain=dom, seed=0)
    assert float(one) == 6.0
    assert float(tq.MonteCarlo().integrate(F.Polynomial(1, [3.0]), 1, N=1, integration_domain=dom, seed=0)) == 6.0
    assert abs(float(tq.Trapezoid().integrate(lambda x: x[:, 0], 1, N=2, integration_domain=dom)) - 2.0) < 1e-6
    assert abs(float(tq.Simpson().integrate(lambda x: x[:, 0] ** 2, 1, N=None, integration_domain=dom)) - 8 / 3) < 1e-5
    assert abs(float(tq.Boole().integrate(lambda x: x[:, 0] ** 4, 1, N=None, integration_domain=dom)) - 32 / 5) < 1e-4
    # the smallest VEGAS run the reference can do: N=125 -> 5 evaluations per iteration, 2 map intervals
    v = tq.VEGAS()
    r = v.integrate(lambda x: x[:, 0] * 0 + 1.0, 1, N=125, integration_domain=dom, seed=0)
    assert v.map.N_intervals == 2 and v.strat.N_cubes == 1 and abs(float(r) - 2.0) < 1e-5
    with pytest.raises(ZeroDivisionError):  # N_strat = 0, exactly like the reference (vegas_stratification.py:27-31)
        tq.VEGAS().integrate(lambda x: x[:, 0], 1, N=26, integration_domain=dom, seed=0)


def test_many_dimensions_unfused(cuda):
    """Unfused kernels are not limited to TQ_MAX_DIM (only the fused functors are)."""
    dim = 40
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=torch.float64, device=cuda)
    fn = lambda x: torch.sum(x, dim=1)  # noqa: E731
    r = tq.MonteCarlo().integrate(fn, dim, N=200_000, integration_domain=dom, seed=0)
    assert abs(float(r) - dim / 2) < 0.05
    pts = ops.mc_sample(dom, 1000, 5, 0, 0)
    want = O.mc_sample_points(O.philox_uniform(5, 0, 0, 1000, dim, torch.float64), dom.cpu())
    assert torch.equal(pts.cpu(), want)
    v = tq.VEGAS()
    rv = v.integrate(fn, dim, N=100_000, integration_domain=dom, seed=0)
    assert abs(float(rv) - dim / 2) < 0.2 and v.strat.N_strat == 1
    with pytest.raises(ValueError):
        F.SumOfSines(dim)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("dim", [1, 2, 3, 5, 7, 9, 31])
def test_ragged_dims_sample_and_map(cuda, dt, dim):
    """Every (dim, dtype) store-width path of the row-major producers against the oracle, odd row counts."""
    rows = 1237
    u = ops.philox_uniform(rows, dim, dt, cuda, 9, 2, 1000)
    assert torch.equal(u.cpu(), O.philox_uniform(9, 2, 1000, rows, dim, dt))
    ns = 3
    n_cubes = ns ** min(dim, 6)
    g = torch.Generator().manual_seed(dim)
    nh = torch.randint(2, 6, (n_cubes,), generator=g)
    offsets = ops.strat_offsets(nh.to(cuda))
    M = int(offsets[-1])
    y = ops.strat_sample(offsets, ns, dim, dt, 0, M, seed=4, call_idx=1)
    assert y.shape == (M, dim) and float(y.min()) >= 0 and float(y.max()) < 1
    # same rows from an arbitrary sub-range; digits follow the cube index (dim 0 fastest)
    a, b = M // 5, min(M, M // 5 + 333)
    assert torch.equal(ops.strat_sample(offsets, ns, dim, dt, a, b, seed=4, call_idx=1), y[a:b])
    cube = torch.repeat_interleave(torch.arange(n_cubes), nh)
    digits = O.strat_cube_digits(n_cubes, ns, dim)[cube]
    assert torch.equal(torch.floor(y.cpu().double() * ns).long().clamp(max=ns - 1)[:, : min(dim, 6)], digits[:, : min(dim, 6)])
    xe, dxe, w, c = (t.to(cuda) for t in O.map_init(11, dim, dt))
    x, jac, ids = ops.map_forward(y, xe, dxe, want_ids=True)
    assert torch.equal(x.cpu(), O.map_get_x(y.cpu(), xe.cpu(), dxe.cpu()))
    assert torch.equal(ids.cpu().long(), O.interval_id(y.cpu(), 11))
    nodes = torch.linspace(0, 1, 4, dtype=dt, device=cuda).repeat(min(dim, 6), 1).contiguous()
    total = 4 ** min(dim, 6)
    lo, hi = min(2, total - 1), total - 1
    pts = ops.nc_grid_points(nodes, lo, hi)
    mesh = torch.stack([m.ravel() for m in torch.meshgrid(*[nodes[0].cpu()] * min(dim, 6), indexing="ij")], dim=1)
    assert torch.equal(pts.cpu(), mesh[lo:hi])


def test_results_do_not_depend_on_launch_geometry(cuda):
    """Row sub-ranges, chunk sizes and grid sizes must not change what is computed (fp64: to rounding)."""
    fn = F.GenzC0(3, a=[1.0, 2.0, 3.0], u=[0.2, 0.5, 0.8])
    s = fn.to_struct([0.0] * 3, [1.0] * 3, 1.0)
    full = ops.fused_mc(s, torch.float64, cuda, 0, 1_000_003, 3, 0)
    parts = sum(ops.fused_mc(s, torch.float64, cuda, a, b, 3, 0) for a, b in [(0, 1), (1, 17), (17, 999_999), (999_999, 1_000_003)])
    assert torch.allclose(full, parts, rtol=1e-13)
    nodes = torch.linspace(0, 1, 13, dtype=torch.float64, device=cuda).repeat(3, 1).contiguous()
    table = tq.Simpson()._weight_table(13, 3, torch.float64, cuda)
    whole = ops.fused_nc(s, nodes, table, 0, 13**3)
    split = ops.fused_nc(s, nodes, table, 0, 1000) + ops.fused_nc(s, nodes, table, 1000, 13**3)
    assert torch.allclose(whole, split, rtol=1e-13)
    f = torch.rand(13**3, dtype=torch.float64, device=cuda)
    both = float(ops.nc_contract_f64(f[:777].contiguous(), table, 0, 777) + ops.nc_contract_f64(f[777:].contiguous(), table, 777, 13**3))
    assert abs(float(ops.nc_contract(f, table)) - both) <= 1e-14 * abs(both)


@pytest.mark.parametrize("cols", [2, 3, 16, 200, 256, 257, 1000, 2048])
def test_vector_valued_reductions_every_column_count(cuda, cols):
    """Vector-valued integrands (base_integrator.py:77-89, grid_integrator.py:70-82): column sums and the weighted
    contraction for narrow (a thread keeps one column) and wide (a thread keeps several) outputs, tile remainders."""
    g = torch.Generator().manual_seed(cols)
    for dt, rtol in ((torch.float64, 1e-12), (torch.float32, 2e-6)):
        rows = 3000 if cols > 256 else 70_001
        f = (torch.rand(rows, cols, generator=g, dtype=torch.float64) - 0.3).to(dt)
        s, q = ops.sum_columns(f.to(cuda), want_sumsq=True)
        assert torch.allclose(s.cpu(), f.double().sum(0), rtol=1e-11, atol=1e-9)
        assert torch.allclose(q.cpu(), (f.double() ** 2).sum(0), rtol=1e-11, atol=1e-9)
        n, dim = 11, 3
        w = (torch.rand(dim, n, generator=g, dtype=torch.float64) + 0.5).to(dt)
        fv = (torch.rand(n**dim, cols, generator=g, dtype=torch.float64) - 0.3).to(dt)
        W = torch.einsum("i,j,k->ijk", w[0].double(), w[1].double(), w[2].double()).reshape(-1, 1)
        want = (fv.double() * W).sum(0)
        got = ops.nc_contract(fv.to(cuda), w.to(cuda))
        assert got.shape == (cols,) and got.dtype == dt
        assert torch.allclose(got.double().cpu(), want, rtol=rtol * 50, atol=rtol * 50 * float(want.abs().max()))
        # a sub-range of points, as the chunked / multi-GPU paths use it
        part = ops.nc_contract_f64(fv[100:900].contiguous().to(cuda), w.to(cuda), 100, 900)
        assert torch.allclose(part.cpu(), (fv[100:900].double() * W[100:900]).sum(0), rtol=rtol * 50,
                              atol=rtol * 50 * float(want.abs().max()))
        # deterministic: the same launch twice gives the same bits
        assert torch.equal(ops.nc_contract(fv.to(cuda), w.to(cuda)), got)


def test_complex_integrand_values(cuda):
    """The reference's tests integrate complex-valued integrands (tests/helper_functions.py:98-110); sums and weighted
    contractions are linear, so complex values run through the real kernels as [..., 2] views."""
    import torchquad_b200 as tq

    torch.set_default_dtype(torch.float64)
    dom = torch.tensor([[0.0, 2.0], [-1.0, 1.0]], dtype=torch.float64, device=cuda)
    fn = lambda x: (x[:, 0] + 1j * x[:, 1] ** 2) * (2.0 - 0.5j)  # noqa: E731
    exact = (4.0 + 1j * (2.0 * 2.0 / 3.0)) * (2.0 - 0.5j)  # int x0 = 2*2 = 4; int x1^2 = 2 * 2/3
    for integ, kw, tol in ((tq.Simpson(), dict(N=41**2), 1e-12), (tq.Boole(), dict(N=41**2), 1e-12), (tq.Trapezoid(), dict(N=201**2), 1e-3),
                           (tq.MonteCarlo(), dict(N=400_000, seed=1), 2e-2), (tq.GaussLegendre(), dict(N=8**2), 1e-12)):
        r = integ.integrate(fn, 2, integration_domain=dom, **kw)
        assert r.is_complex() and r.dtype == torch.complex128 and r.dim() == 0
        assert abs(complex(r) - exact) <= tol * abs(exact), (type(integ).__name__, complex(r), exact)
    vec = lambda x: torch.stack([fn(x), 2.0 * fn(x)], dim=1)  # noqa: E731
    r = tq.Simpson().integrate(vec, 2, N=41**2, integration_domain=dom)
    assert r.shape == (2,) and abs(complex(r[1]) - 2 * exact) <= 1e-12 * abs(exact)
    f32 = tq.Simpson().integrate(fn, 2, N=41**2, integration_domain=dom.float())
    assert f32.dtype == torch.complex64 and abs(complex(f32) - exact) <= 1e-5 * abs(exact)
