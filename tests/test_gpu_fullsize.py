"""BASELINE.json configurations at FULL size on the GPU, checked through size-independent properties
(the oracle cannot run these sizes): closed-form integrals within the reported error, exact integer
invariants of the histograms / sample counts, composition over row ranges, run-to-run determinism of the
integer state.  A few seconds each on a B200."""
import math
import warnings

import pytest
import torch

import torchquad_b200 as tq
from torchquad_b200 import integrands as F
from torchquad_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def test_c2_monte_carlo_10d_1e9_fp32(cuda):
    """configs[1]: MonteCarlo 10-D sum of sines, N=1e9, fp32."""
    fn = F.SumOfSines(10)
    dom = torch.tensor([[0.0, 1.0]] * 10, dtype=torch.float32, device=cuda)
    mc = tq.MonteCarlo()
    res = mc.integrate(fn, 10, N=10**9, integration_domain=dom, seed=0)
    sigma = mc.get_error_estimate()
    assert res.dtype == torch.float32 and mc._nr_of_fevals == 10**9
    assert abs(float(res) - fn.exact()) <= 5 * sigma + 1e-6 * fn.exact()  # fp32 rounding of the final scalar
    # composition: the sums over disjoint row ranges of the same stream add up to the full-range sums
    starts, sizes = [0.0] * 10, [1.0] * 10
    s = fn.to_struct(starts, sizes, 1.0)
    full = ops.fused_mc(s, torch.float32, cuda, 0, 10**9, 0, 0)
    cuts = [0, 123_456_789, 500_000_001, 999_999_999, 10**9]
    parts = sum(ops.fused_mc(s, torch.float32, cuda, a, b, 0, 0) for a, b in zip(cuts, cuts[1:]))
    assert torch.allclose(full, parts, rtol=1e-12, atol=0)
    # unfused chunked path on 1e8 rows of the same stream agrees with the fused kernel on those rows
    small = tq.MonteCarlo()
    small.max_points_bytes = 1 << 30
    u = small.integrate(lambda x: torch.sum(torch.sin(x), dim=1), 10, N=10**8, integration_domain=dom, seed=0)
    f = ops.fused_mc(s, torch.float32, cuda, 0, 10**8, 0, 0)
    assert abs(float(u) - float(f[0]) / 1e8) <= 2e-6 * float(u)


def test_c3_boole_simpson_6d_fp64(cuda):
    """configs[2]: Boole / Simpson 6-D tensor-product grids of ~1e9 points, fp64."""
    fn = F.ProductOfCosines(6)
    dom = torch.tensor([[0.0, 1.0]] * 6, dtype=torch.float64, device=cuda)
    b = tq.Boole()
    rb = b.integrate(fn, 6, N=33**6, integration_domain=dom)
    assert b._nr_of_fevals == 33**6 and abs(float(rb) - fn.exact()) < 1e-11
    s = tq.Simpson()
    rs = s.integrate(fn, 6, N=31**6, integration_domain=dom)
    assert s._nr_of_fevals == 31**6 and abs(float(rs) - fn.exact()) < 5e-8
    bb = tq.Boole()
    rbb = bb.integrate(fn, 6, N=31**6, integration_domain=dom)  # adjusts 31 -> 29 per dim (boole.py:56-84)
    assert bb._nr_of_fevals == 29**6 and abs(float(rbb) - fn.exact()) < 1e-10
    # linearity in the integrand: weights contract a sum as the sum of contractions (unfused kernel, n=21)
    nodes = torch.linspace(0, 1, 21, dtype=torch.float64, device=cuda).repeat(6, 1).contiguous()
    table = tq.Boole()._weight_table(21, 6, torch.float64, cuda)
    pts = ops.nc_grid_points(nodes)
    f1, f2 = torch.prod(torch.cos(pts), dim=1), torch.sum(pts**2, dim=1)
    lhs = ops.nc_contract(f1 + 2.0 * f2, table)
    rhs = ops.nc_contract(f1, table) + 2.0 * ops.nc_contract(f2, table)
    assert abs(float(lhs) - float(rhs)) <= 1e-12 * abs(float(rhs))


def test_c4_vegas_8d_iteration_invariants(cuda):
    """configs[3] scale: one warm-up-free stratified pass of ~8.4e7 samples with the reference's map size
    (Ni = 1e7 per dim): integer state must satisfy its invariants exactly and be reproducible."""
    dim, N = 8, 2_500_000_000
    fn = F.GenzOscillatory(dim, a=0.5, u=0.3)
    inc = N // 25
    vmap = tq.VEGASMap(max(2, inc // 10), dim, "torch", torch.float64, device=cuda)
    strat = tq.VEGASStratification(inc, dim, tq.RNG(seed=1), "torch", torch.float64, device=cuda)
    assert vmap.N_intervals == 10**7 and strat.N_strat == 8 and strat.N_cubes == 8**8
    nh = strat.get_NH(inc)
    offsets = strat._offsets
    M = int(offsets[-1])
    assert M == int(nh.sum()) == 5 * 8**8 and int(nh.min()) >= 2
    s = fn.to_struct([0.0] * dim, [1.0] * dim, 1.0)
    snap = []
    for rep in range(2):
        vmap._reset_weight()
        JF = torch.zeros((2, strat.N_cubes), dtype=torch.float64, device=cuda)
        ops.fused_vegas(s, vmap.packed_edges(), vmap.weights, vmap.counts, 0, M, 1, 7, offsets=offsets,
                        n_strat=strat.N_strat, JF=JF[0], JF2=JF[1])
        assert vmap.counts.sum(dim=1).tolist() == [M] * dim  # every sample lands in exactly one bin per dimension
        assert int(vmap.counts.min()) >= 0 and float(vmap.weights.min()) >= 0.0 and float(JF[1].min()) >= 0.0
        snap.append((vmap.counts.clone(), JF.clone(), vmap.weights.clone()))
    assert torch.equal(snap[0][0], snap[1][0])  # integer state is run-to-run deterministic
    assert torch.allclose(snap[0][1], snap[1][1], rtol=1e-12) and torch.allclose(snap[0][2], snap[1][2], rtol=1e-12)
    # estimator on the fresh (identity) map == plain stratified MC estimate of the integral
    strat.JF, strat.JF2 = snap[1][1][0], snap[1][1][1]
    strat.update_DH()
    I, s2 = strat.last_scalars[:2].tolist()
    assert abs(I - fn.exact()) <= 5 * math.sqrt(s2)
    assert abs(float(strat.dh.sum()) - 1.0) < 1e-9
    # the row-range split used by the multi-GPU path reproduces the single-range counts exactly
    vmap._reset_weight()
    JF = torch.zeros((2, strat.N_cubes), dtype=torch.float64, device=cuda)
    for a, b in [(0, M // 3), (M // 3, M // 3 + 12345), (M // 3 + 12345, M)]:
        ops.fused_vegas(s, vmap.packed_edges(), vmap.weights, vmap.counts, a, b, 1, 7, offsets=offsets,
                        n_strat=strat.N_strat, JF=JF[0], JF2=JF[1])
    assert torch.equal(vmap.counts, snap[0][0]) and torch.allclose(JF, snap[0][1], rtol=1e-12)


def test_c4_c5_vegas_full_runs_hit_the_closed_form(cuda):
    """configs[3]/[4]: complete VEGAS runs at full N (fused path); the map is capped for the fp32 run because
    the reference's Ni = 4e7 exceeds fp32 resolution."""
    dom8 = torch.tensor([[0.0, 1.0]] * 8, dtype=torch.float64, device=cuda)
    for fn in (F.GenzOscillatory(8, a=0.5, u=0.3), F.GenzCornerPeak(8, a=0.25)):
        v = tq.VEGAS()
        r = v.integrate(fn, 8, N=2_500_000_000, integration_domain=dom8, seed=3)
        err = float(v._get_error())
        assert v.it == 10 and v.map.N_intervals == 10**7
        assert abs(float(r) - fn.exact()) <= 5 * err, (float(r), fn.exact(), err)
        assert 0.55 * 2.5e9 < v._nr_of_fevals <= 2.5e9
    fn16 = F.GenzProductPeak(16, a=2.0, u=0.5)
    v = tq.VEGAS()
    v.max_map_intervals = 4096
    dom16 = torch.tensor([[0.0, 1.0]] * 16, dtype=torch.float32, device=cuda)
    r = v.integrate(fn16, 16, N=10**10, integration_domain=dom16, seed=3)
    err = float(v._get_error())
    assert abs(float(r) - fn16.exact()) <= 5 * err + 2e-6 * fn16.exact(), (float(r), fn16.exact(), err)
    assert v.strat.N_strat == 3 and v.strat.N_cubes == 3**16
