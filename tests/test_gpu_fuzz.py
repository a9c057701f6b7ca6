"""Randomised parity sweep: every stateless kernel against the oracle on random shapes / dtypes / values
(seeded, ~60 configurations, a few seconds).  Integer outputs exact, floating outputs bit-identical where the
arithmetic is order-free, tolerance otherwise."""
import random

import pytest
import torch

from oracle import ref_oracle as O
from torchquad_b200 import ops

pytestmark = pytest.mark.gpu


def _cases(n, seed):
    rnd = random.Random(seed)
    for i in range(n):
        yield i, rnd.choice([torch.float32, torch.float64]), rnd


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fuzz_map_and_strat_kernels(cuda, seed):
    for i, dt, rnd in _cases(10, seed):
        g = torch.Generator().manual_seed(1000 * seed + i)
        dim = rnd.choice([1, 2, 3, 4, 5, 6, 8, 11, 16])
        ni = rnd.choice([2, 3, 17, 64, 333, 4096, 40_000])
        rows = rnd.choice([1, 31, 256, 257, 1000, 5003, 20_000])
        tol = 1e-12 if dt == torch.float64 else 1e-5
        # an adapted, strictly monotone map
        xe = torch.sort(torch.rand(dim, ni + 1, generator=g, dtype=torch.float64), dim=1).values
        xe[:, 0], xe[:, -1] = 0.0, 1.0
        xe = xe.to(dt)
        dxe = xe[:, 1:] - xe[:, :-1]
        y = (torch.rand(rows, dim, generator=g, dtype=torch.float64) * 0.999999).to(dt)
        x, jac, ids = ops.map_forward(y.to(cuda), xe.to(cuda), dxe.to(cuda), want_ids=True)
        assert torch.equal(ids.cpu().long(), O.interval_id(y, ni)), (dim, ni, rows, dt)
        assert torch.equal(x.cpu(), O.map_get_x(y, xe, dxe)) and torch.equal(jac.cpu(), O.map_get_jac(y, dxe))
        jf2 = torch.rand(rows, generator=g, dtype=torch.float64).to(dt)
        w, c = (t.to(cuda) for t in O.map_reset(ni, dim, dt))
        ops.map_accumulate(y.to(cuda), jf2.to(cuda), w, c)
        wr, cr = O.map_reset(ni, dim, dt)
        O.map_accumulate(wr, cr, y, jf2)
        assert torch.equal(c.cpu(), cr) and int(c.sum()) == rows * dim
        assert torch.allclose(w.cpu(), wr, rtol=1e-12 if dt == torch.float64 else 2e-4, atol=0)
        # update from the accumulated state (zero-count bins are common when rows << ni)
        ref = O.smooth_map(wr, cr, 0.5)
        sm, st = ops.map_smooth(w, c, 0.5)
        if ref is None:
            assert int(st[0]) == 1
        else:
            assert int(st[0]) == 0
            assert torch.allclose(sm.cpu(), ref, rtol=max(tol, 1e-6 if dt == torch.float32 else 0), atol=1e-30)
        # stratification with a random (peaked) dh
        ns = rnd.choice([1, 2, 3, 5])
        sdim = min(dim, 6)
        n_cubes = ns**sdim
        dh = torch.rand(n_cubes, generator=g, dtype=torch.float64) ** rnd.choice([1, 4, 12])
        dh = (dh / dh.sum()).to(dt)
        nev = rnd.choice([10, 1000, 50_000])
        nh, offsets = ops.strat_nh(dh.to(cuda), nev)
        nh_ref = O.strat_get_nh(dh, nev)
        assert torch.equal(nh.cpu(), nh_ref)
        M = int(offsets[-1])
        assert M == int(nh_ref.sum())
        u = O.philox_uniform(seed, i, 0, M, sdim, dt)
        ys = ops.strat_sample(offsets, ns, sdim, dt, 0, M, u_in=u.to(cuda))
        assert torch.equal(ys.cpu(), O.strat_get_y(nh_ref, ns, sdim, u))
        jf = torch.rand(M, generator=g, dtype=torch.float64).to(dt) - 0.3
        JF, JF2 = ops.strat_accumulate(jf.to(cuda), offsets)
        oJF, oJF2 = O.strat_accumulate(nh_ref, jf)
        light = nh_ref <= 256
        assert torch.equal(JF.cpu()[light], oJF[light]) and torch.equal(JF2.cpu()[light], oJF2[light])
        assert torch.allclose(JF.cpu(), oJF, rtol=tol * 10, atol=tol) and torch.allclose(JF2.cpu(), oJF2, rtol=tol * 10, atol=tol)
        dh2, scal = ops.strat_update(oJF.to(cuda), oJF2.to(cuda), nh, (1.0 / ns) ** sdim, 0.75)
        want = O.strat_update_dh(oJF, oJF2, nh_ref.to(dt), (1.0 / ns) ** sdim, 0.75)
        assert torch.allclose(dh2.cpu(), want, rtol=1e-11 if dt == torch.float64 else 2e-5, atol=1e-30)
        I, s2 = O.vegas_iteration_estimate(oJF, oJF2, nh_ref, (1.0 / ns) ** sdim)
        assert abs(float(scal[0]) - float(I)) <= tol * 10 * max(abs(float(I)), 1e-30) + 1e-30


@pytest.mark.parametrize("seed", [0, 1])
def test_fuzz_sampling_and_grid_kernels(cuda, seed):
    import torchquad_b200 as tq

    for i, dt, rnd in _cases(12, seed + 10):
        g = torch.Generator().manual_seed(77 * seed + i)
        dim = rnd.choice([1, 2, 3, 4, 6, 7, 10, 13])
        rows = rnd.choice([1, 2, 255, 256, 1025, 30_001])
        row0 = rnd.choice([0, 1, 2**32 - 7, 10**12])
        lo = torch.rand(dim, generator=g, dtype=torch.float64) * 4 - 2
        dom = torch.stack([lo, lo + torch.rand(dim, generator=g, dtype=torch.float64) * 3], dim=1).to(dt)
        pts = ops.mc_sample(dom.to(cuda), rows, 99 + i, 3, row0)
        want = O.mc_sample_points(O.philox_uniform(99 + i, 3, row0, rows, dim, dt), dom)
        assert torch.equal(pts.cpu(), want), (dim, rows, row0, dt)
        f = torch.rand(rows, rnd.choice([1, 2, 5]), generator=g, dtype=torch.float64).to(dt)
        s, q = ops.sum_columns(f.to(cuda), want_sumsq=True)
        assert torch.allclose(s.cpu(), f.double().sum(0), rtol=1e-12) and torch.allclose(q.cpu(), (f.double() ** 2).sum(0), rtol=1e-12)
        # Newton-Cotes / Gauss on a small grid
        gdim = min(dim, 4)
        rule, cls, n = rnd.choice([("trapezoid", tq.Trapezoid, 6), ("simpson", tq.Simpson, 7), ("boole", tq.Boole, 9)])
        gdom = dom[:gdim]
        pts_ref, hs_ref, n_ref = O.nc_grid(rule, n**gdim, gdom)
        integ = cls()
        gp, hs, nn = integ.calculate_grid(n**gdim, gdom.to(cuda))
        assert nn == n_ref and torch.equal(gp.cpu(), pts_ref) and torch.equal(hs.cpu(), hs_ref)
        fv = torch.rand(n**gdim, generator=g, dtype=torch.float64).to(dt)
        got = integ.calculate_result(fv.to(cuda), gdim, nn, hs, gdom.to(cuda))
        ref = O.nc_result(rule, fv, gdim, n_ref, hs_ref)
        assert abs(float(got) - float(ref)) <= (1e-12 if dt == torch.float64 else 2e-5) * max(abs(float(ref)), 1e-12)
