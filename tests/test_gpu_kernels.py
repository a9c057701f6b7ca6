"""Parity of every C-ABI kernel against the CPU oracle / golden fixtures on identical injected inputs.

Bars (BASELINE.json north_star): integer results (bin ids, counts, nh, offsets) bit-exact; maps, Jacobians
and estimates within 1e-12 relative in fp64 and 1e-5 in fp32.  Where the arithmetic is order-free the CUDA
result is in fact bit-identical to the CPU reference and the test says so.
"""
import numpy as np
import pytest
import torch

from oracle import ref_oracle as O
from torchquad_b200 import ops

pytestmark = pytest.mark.gpu
DT = {"f32": torch.float32, "f64": torch.float64}
RTOL = {"f32": 1e-5, "f64": 1e-12}


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1e-300)).max()) if a.numel() else 0.0


# ---------------------------------------------------------------- RNG / Monte Carlo
@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("rows,dim,row0", [(1, 1, 0), (1000, 3, 0), (4097, 4, 123456789012), (513, 10, 7), (100, 16, 2**32 - 50), (33, 31, 0)])
def test_philox_uniform_bit_exact(cuda, tag, rows, dim, row0):
    got = ops.philox_uniform(rows, dim, DT[tag], cuda, seed=0xDEADBEEFCAFE, call_idx=5, row_begin=row0)
    want = O.philox_uniform(0xDEADBEEFCAFE, 5, row0, rows, dim, DT[tag])
    assert got.dtype == DT[tag] and got.shape == (rows, dim)
    assert torch.equal(got.cpu(), want)
    assert float(got.min()) >= 0.0 and float(got.max()) < 1.0


def test_philox_row_ranges_compose(cuda):
    full = ops.philox_uniform(1000, 5, torch.float32, cuda, 3, 1, 0)
    parts = torch.cat([ops.philox_uniform(b - a, 5, torch.float32, cuda, 3, 1, a) for a, b in [(0, 333), (333, 334), (334, 1000)]])
    assert torch.equal(full, parts)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_mc_sample_matches_golden(cuda, golden, tag):
    g = golden(f"monte_carlo_{tag}")
    dom = dev(g["domain"], cuda)
    pts = ops.mc_sample(dom, 5000, 1, 0, 0)
    assert torch.equal(pts.cpu(), torch.from_numpy(g["points"]))  # bit-identical to u*(b-a)+a on CPU


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_sum_columns(cuda, golden, tag):
    g = golden(f"monte_carlo_{tag}")
    f = dev(g["f"], cuda)
    s, q = ops.sum_columns(f, want_sumsq=True)
    assert abs(float(s[0]) - float(g["f"].astype(np.float64).sum())) <= 1e-12 * abs(float(s[0]))
    assert abs(float(q[0]) - float((g["f"].astype(np.float64) ** 2).sum())) <= 1e-12 * abs(float(q[0]))
    fv = dev(g["fv"], cuda)
    s, q = ops.sum_columns(fv, want_sumsq=True)
    want = g["fv"].astype(np.float64).sum(axis=0)
    assert np.allclose(s.cpu().numpy(), want, rtol=1e-12)
    assert np.allclose(q.cpu().numpy(), (g["fv"].astype(np.float64) ** 2).sum(axis=0), rtol=1e-12)
    # odd sizes / unaligned views / empty
    for n in [0, 1, 3, 255, 1025]:
        t = torch.arange(n, dtype=DT[tag], device=cuda) + 1
        s, _ = ops.sum_columns(t)
        assert float(s[0]) == n * (n + 1) / 2
    big = torch.ones(3_000_001, dtype=DT[tag], device=cuda)[1:]
    assert float(ops.sum_columns(big)[0][0]) == 3_000_000.0


# ---------------------------------------------------------------- VEGAS map
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_map_golden_ids_and_fresh_map(cuda, golden, tag):
    g = golden(f"vegas_map_{tag}")
    xe, dxe, _, _ = O.map_init(20, 3, DT[tag])
    y0 = dev(g["y0"], cuda)
    x, jac, ids, off = ops.map_forward(y0, xe.to(cuda), dxe.to(cuda), want_ids=True, want_offset=True)
    assert ids.cpu().tolist() == [[16, 8, 3], [9, 13, 18], [12, 1, 11]]  # reference golden, tests/vegas_map_test.py:33-37
    assert torch.equal(off.cpu(), torch.from_numpy(g["off0"]))
    assert torch.equal(x.cpu(), torch.from_numpy(g["x0"]))
    assert float((x - y0).abs().max()) <= 3e-7 and float((jac - 1).abs().max()) < 1e-14 + (1e-6 if tag == "f32" else 0)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_map_forward_accumulate_update_sequence(cuda, golden, tag, record_delta):
    """Three adaptive updates on identical samples: x/jac bit-identical, counts exact, edges within tolerance."""
    g = golden(f"vegas_map_{tag}")
    dt = DT[tag]
    xe, dxe, w, c = (t.to(cuda) for t in O.map_init(50, 4, dt))
    status = torch.zeros(4, dtype=torch.int32, device=cuda)
    for it in (1, 2, 3):
        y = dev(g[f"y{it}"], cuda)
        # feed the golden edges so each step is compared on identical state
        x, jac, _ = ops.map_forward(y, xe, dxe)
        if it == 1:
            assert torch.equal(x.cpu(), torch.from_numpy(g["x1"])) and torch.equal(jac.cpu(), torch.from_numpy(g["jac1"]))
        else:
            assert rel_err(x, torch.from_numpy(g[f"x{it}"])) <= RTOL[tag]
            assert rel_err(jac, torch.from_numpy(g[f"jac{it}"])) <= RTOL[tag] * 10
        ops.map_accumulate(y, dev(g[f"jf2_{it}"], cuda), w, c)
        if it == 1:
            assert torch.equal(c.cpu(), torch.from_numpy(g["c1"]))  # counts bit-exact
            assert rel_err(w, torch.from_numpy(g["w1"])) <= (1e-13 if tag == "f64" else 2e-6)
        # smoothing of the golden histogram
        sm, st = ops.map_smooth(dev(g[f"w{it}"], cuda), dev(g[f"c{it}"], cuda), 0.5)
        assert int(st[0]) == 0
        assert rel_err(sm, torch.from_numpy(g[f"sm{it}"])) <= (5e-14 if tag == "f64" else 2e-6)
        # update from golden state (isolates the rebinning kernel)
        gxe = dev(g[f"xe{it-1}"], cuda) if it > 1 else O.map_init(50, 4, dt)[0].to(cuda)
        gdxe = dev(g[f"dxe{it-1}"], cuda) if it > 1 else O.map_init(50, 4, dt)[1].to(cuda)
        gw, gc = dev(g[f"w{it}"], cuda), dev(g[f"c{it}"], cuda)
        ops.map_update(gxe, gdxe, gw, gc, 0.5, status)
        assert status.tolist() == [0, 0, 0, 0]
        want_xe, want_dxe = torch.from_numpy(g[f"xe{it}"]), torch.from_numpy(g[f"dxe{it}"])
        # north_star: maps within 1e-12 (fp64) / 1e-5 (fp32).  x_edges are O(1) numbers: absolute = relative to the map;
        # dx_edges = diff(x_edges) are differences of close numbers, so their error is judged against the map's scale too
        # (max dx); the element-wise relative error of a tiny dx is reported, not asserted below its conditioning
        # (the reference itself forms dx with an fp32 cumsum + subtraction, vegas_map.py:233-259).
        dx_scale = float(want_dxe.abs().max())
        e_x = record_delta(f"map_update {tag} it{it}: max |x_edges - ref|", float((gxe.cpu() - want_xe).abs().max()),
                           1e-14 if tag == "f64" else 1e-6)
        e_dx = record_delta(f"map_update {tag} it{it}: max |dx_edges - ref| / max dx",
                            float((gdxe.cpu() - want_dxe).abs().max()) / dx_scale, 1e-12 if tag == "f64" else 1e-5)
        record_delta(f"map_update {tag} it{it}: max elementwise rel err of dx_edges (reported)", rel_err(gdxe, want_dxe),
                     1e-11 if tag == "f64" else 5e-4)
        assert e_x <= (1e-14 if tag == "f64" else 1e-6)
        assert e_dx <= (1e-12 if tag == "f64" else 1e-5)
        assert rel_err(gdxe, want_dxe) <= (1e-11 if tag == "f64" else 5e-4)
        assert torch.equal(gxe[:, [0, -1]].cpu(), want_xe[:, [0, -1]])  # outer edges exactly 0 and 1
        assert int(gc.abs().sum()) == 0 and float(gw.abs().sum()) == 0.0  # reset
        # continue the chain on our own state
        ops.map_update(xe, dxe, w, c, 0.5, status)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_map_smooth_zero_count_fill(cuda, golden, tag):
    g = golden(f"vegas_map_{tag}")
    s6, st = ops.map_smooth(dev(g["w6"], cuda), dev(g["c6"], cuda), 0.5)
    want = torch.tensor([[0.0, 0.0, 0.54913316, 0.75820765, 0.77899047, 0.77899047],
                         [-0.0, -0.0, 0.64868024, 0.93220967, 0.64868024, -0.0]], dtype=torch.float64)
    assert float((s6.double().cpu() - want).abs().max()) <= 3e-7  # reference golden, tests/vegas_map_test.py:84-100
    assert rel_err(s6, torch.from_numpy(g["s6"])) <= (1e-14 if tag == "f64" else 1e-6) or float((s6.cpu() - torch.from_numpy(g["s6"])).abs().max()) < 1e-7
    sz, st = ops.map_smooth(dev(g["wz"], cuda), dev(g["cz"], cuda), 0.5)
    assert int(st[0]) == 0
    assert float((sz.cpu() - torch.from_numpy(g["sz"])).abs().max()) <= (1e-14 if tag == "f64" else 1e-6)
    xe, dxe, _, _ = (t.to(cuda) for t in O.map_init(80, 3, DT[tag]))
    status = torch.zeros(4, dtype=torch.int32, device=cuda)
    w, c = dev(g["wz"], cuda), dev(g["cz"], cuda)
    ops.map_update(xe, dxe, w, c, 0.5, status)
    assert float((xe.cpu() - torch.from_numpy(g["xez"])).abs().max()) <= (1e-14 if tag == "f64" else 3e-7)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("rows,dim", [(1, 1), (1000, 3), (4099, 8), (777, 10)])
def test_map_forward_packed_and_fused_accumulate(cuda, tag, rows, dim):
    """Packed-edge forward with the domain transform and the fused tail == the separate reference steps."""
    dt = DT[tag]
    g = torch.Generator().manual_seed(rows + dim)
    ni = 37
    xe = torch.sort(torch.rand(dim, ni + 1, generator=g, dtype=torch.float64), dim=1).values
    xe[:, 0], xe[:, -1] = 0.0, 1.0
    xe = xe.to(dt)
    dxe = xe[:, 1:] - xe[:, :-1]
    y = (torch.rand(rows, dim, generator=g, dtype=torch.float64) * 0.999999).to(dt)
    dom = torch.stack([torch.linspace(-1, 1, dim), torch.linspace(2, 5, dim)], dim=1).to(dt)
    packed = ops.pack_edges(xe.to(cuda), dxe.to(cuda))
    x, jac, ids = ops.map_forward_packed(y.to(cuda), packed, dom.to(cuda), want_ids=True)
    want_x = O.map_get_x(y, xe, dxe) * (dom[:, 1] - dom[:, 0]) + dom[:, 0]
    assert torch.equal(x.cpu(), want_x) and torch.equal(jac.cpu(), O.map_get_jac(y, dxe))
    assert torch.equal(ids.cpu().long(), O.interval_id(y, ni))
    x2, jac2, _ = ops.map_forward(y[1:].to(cuda), xe.to(cuda), dxe.to(cuda))  # unaligned view -> scalar path
    assert torch.equal(x2.cpu(), O.map_get_x(y[1:], xe, dxe)) and torch.equal(jac2, jac[1:])
    f = torch.rand(rows, generator=g, dtype=torch.float64).to(dt)
    vol = float(torch.prod(dom[:, 1] - dom[:, 0]))
    w, c = (t.to(cuda) for t in O.map_reset(ni, dim, dt))
    jf = ops.accumulate_fused(y.to(cuda), f.to(cuda), jac, vol, w, c)
    want_jf = (f * torch.tensor(vol, dtype=dt)) * jac.cpu()
    assert torch.equal(jf.cpu(), want_jf)
    w_ref, c_ref = O.map_reset(ni, dim, dt)
    O.map_accumulate(w_ref, c_ref, y, want_jf**2)
    assert torch.equal(c.cpu(), c_ref) and torch.allclose(w.cpu(), w_ref, rtol=1e-12 if tag == "f64" else 1e-4)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("ni", [33_000, 150_001])
def test_map_update_large_map_pipeline_matches_oracle(cuda, tag, ni):
    """Maps above 32768 intervals take the multi-kernel pipeline (tile scans across CTAs); maps below take the
    single-launch kernel (covered by the golden fixtures).  Same oracle, same tolerances, zero-count runs included."""
    dt = DT[tag]
    g = torch.Generator().manual_seed(ni)
    dim = 3
    pos = torch.linspace(0, 1, ni, dtype=torch.float64)
    w = (torch.rand(dim, ni, generator=g, dtype=torch.float64) + 0.05) * torch.exp(-((pos - 0.3) ** 2) * 40.0)
    c = torch.randint(1, 20, (dim, ni), generator=g)
    for lo, hi in [(0, 7), (1000, 1015), (ni // 2, ni // 2 + 9), (ni - 12, ni)]:
        c[:, lo:hi] = 0
        w[:, lo:hi] = 0
    w = (w * c).to(dt)
    xe, dxe, _, _ = O.map_init(ni, dim, dt)
    sm_ref = O.smooth_map(w, c, 0.5)
    sm, st = ops.map_smooth(w.to(cuda), c.to(cuda), 0.5)
    assert int(st[0]) == 0
    assert rel_err(sm, sm_ref) <= (1e-12 if tag == "f64" else 5e-6)
    xe_ref, dxe_ref, status = O.map_update(xe, dxe, w, c, 0.5)
    assert status == "ok"
    gx, gdx, gw, gc = xe.to(cuda), dxe.to(cuda), w.to(cuda), c.to(cuda)
    packed = torch.empty((dim, ni, 2), dtype=dt, device=cuda)
    stat = torch.zeros(4, dtype=torch.int32, device=cuda)
    ops.map_update(gx, gdx, gw, gc, 0.5, stat, edges_packed=packed)
    assert stat.tolist() == [0, 0, 0, 0]
    assert float((gx.cpu() - xe_ref).abs().max()) <= (1e-13 if tag == "f64" else 1e-6)
    assert torch.equal(gx[:, [0, -1]].cpu(), xe_ref[:, [0, -1]])
    assert float((gdx.cpu().double() - dxe_ref.double()).abs().max()) <= (1e-15 if tag == "f64" else 2e-7) * 10
    assert torch.equal(packed[..., 0], gx[:, :-1]) and torch.equal(packed[..., 1], gdx)
    assert int(gc.abs().sum()) == 0 and float(gw.abs().sum()) == 0.0
    # and a second update chained on the result keeps the edges monotone
    assert float(gdx.min()) > 0


def test_map_update_skips_on_zero_dimension(cuda):
    xe, dxe, w, c = (t.to(cuda) for t in O.map_init(16, 2, torch.float64))
    w[0] = 1.0
    c[0] = 3
    c[1] = 2  # second dimension: counts but zero weights -> row sum 0
    xe0, dxe0 = xe.clone(), dxe.clone()
    status = torch.zeros(4, dtype=torch.int32, device=cuda)
    ops.map_update(xe, dxe, w, c, 0.5, status)
    assert status.tolist()[0] == 1
    assert torch.equal(xe, xe0) and torch.equal(dxe, dxe0)
    assert int(c.sum()) == 0 and float(w.sum()) == 0.0
    sm, st = ops.map_smooth(torch.zeros(2, 16, device=cuda), torch.ones(2, 16, dtype=torch.int64, device=cuda), 0.5)
    assert int(st[0]) == 1


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("ni,rows", [(7, 50_000), (5000, 300_000), (40_000, 20_000)])
def test_map_accumulate_counts_exact_both_paths(cuda, tag, ni, rows):
    """Shared-memory privatised and global-atomic histograms against the CPU scatter_add_."""
    dt = DT[tag]
    g = torch.Generator().manual_seed(ni)
    y = (torch.rand(rows, 3, generator=g, dtype=torch.float64) * 0.999999).to(dt)
    jf2 = torch.rand(rows, generator=g, dtype=torch.float64).to(dt)
    w_ref, c_ref = O.map_reset(ni, 3, dt)
    O.map_accumulate(w_ref, c_ref, y, jf2)
    w, c = (t.to(cuda) for t in O.map_reset(ni, 3, dt))
    ops.map_accumulate(y.to(cuda), jf2.to(cuda), w, c)
    assert torch.equal(c.cpu(), c_ref)
    assert torch.allclose(w.cpu(), w_ref, rtol=1e-12 if tag == "f64" else 2e-4, atol=0)


# ---------------------------------------------------------------- stratification
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_strat_sequence_matches_golden(cuda, golden, tag):
    g = golden(f"vegas_strat_{tag}")
    dt = DT[tag]
    ns, nc, vc = int(g["n_strat"]), int(g["n_cubes"]), float(g["v_cubes"])
    dh = O.strat_init(nc, dt).to(cuda)
    for it in range(3):
        nev = int(g[f"nev{it}"])
        nh, offsets = ops.strat_nh(dh, nev)
        want_nh = torch.from_numpy(g[f"nh{it}"])
        assert torch.equal(nh.cpu(), want_nh)  # bit-exact
        assert torch.equal(offsets.cpu(), torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(want_nh, 0)]))
        M = int(offsets[-1])
        y = ops.strat_sample(offsets, ns, 3, dt, 0, M, u_in=dev(g[f"u{it}"], cuda))
        assert torch.equal(y.cpu(), torch.from_numpy(g[f"y{it}"]))  # bit-identical to the CPU reference
        # sub-range sampling reproduces the same rows
        a, b = M // 3, M // 3 + 1000
        part = ops.strat_sample(offsets, ns, 3, dt, a, b, u_in=dev(g[f"u{it}"][a:b], cuda))
        assert torch.equal(part, y[a:b])
        JF, JF2 = ops.strat_accumulate(dev(g[f"jf{it}"], cuda), offsets)
        light = want_nh <= 256  # one thread per cube, rows in order: bit-identical to the CPU scatter_add_
        wJF, wJF2 = torch.from_numpy(g[f"JF{it}"]), torch.from_numpy(g[f"JF2{it}"])
        assert torch.equal(JF.cpu()[light], wJF[light]) and torch.equal(JF2.cpu()[light], wJF2[light])
        # heavier cubes are summed by a warp with fp64 accumulation (more accurate than the fp32 sequential sum)
        assert torch.allclose(JF.cpu(), wJF, rtol=RTOL[tag] * 2) and torch.allclose(JF2.cpu(), wJF2, rtol=RTOL[tag] * 2)
        JF, JF2 = wJF.to(cuda), wJF2.to(cuda)  # continue from the golden sums
        dh, scal = ops.strat_update(JF, JF2, nh, vc, 0.75)
        want_dh = torch.from_numpy(g[f"dh{it}"])
        assert rel_err(dh, want_dh) <= (1e-13 if tag == "f64" else 3e-6)
        assert abs(float(scal[0]) - float(g[f"I{it}"])) <= RTOL[tag] * abs(float(g[f"I{it}"]))
        assert float(scal[3]) == float(nh.sum())
        assert abs(float(scal[1]) - float(g[f"s2{it}"])) <= RTOL[tag] * 10 * abs(float(g[f"s2{it}"]))
        dh = want_dh.to(cuda)  # continue from the golden state so nh stays bit-comparable
    nhp, _ = ops.strat_nh(dev(g["dhp"], cuda), 54321)
    assert torch.equal(nhp.cpu(), torch.from_numpy(g["nhp"]))


def test_strat_offsets_and_backward(cuda):
    g = torch.Generator().manual_seed(1)
    nh = torch.randint(2, 40, (5000,), generator=g)
    offsets = ops.strat_offsets(nh.to(cuda))
    want = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(nh, 0)])
    assert torch.equal(offsets.cpu(), want)
    M = int(want[-1])
    jf = torch.rand(M, generator=g, dtype=torch.float64).to(cuda).requires_grad_(True)
    JF, JF2 = ops.strat_accumulate(jf, offsets)
    wts = torch.rand(5000, generator=g, dtype=torch.float64).to(cuda)
    (JF * wts).sum().backward()
    want_grad = torch.repeat_interleave(wts.cpu(), nh)
    assert torch.equal(jf.grad.cpu(), want_grad)
    oJF, _ = O.strat_accumulate(nh, jf.detach().cpu())
    assert torch.equal(JF.detach().cpu(), oJF)


def test_strat_heavy_cube(cuda):
    """One cube holding most samples (peaked dh): long sequential segment, tile-spanning lookups."""
    nh = torch.full((300,), 2, dtype=torch.int64)
    nh[137] = 50_000
    offsets = ops.strat_offsets(nh.to(cuda))
    M = int(offsets[-1])
    u = O.philox_uniform(5, 0, 0, M, 2, torch.float64)
    y = ops.strat_sample(offsets, 20, 2, torch.float64, 0, M, u_in=u.to(cuda))
    want = O.strat_get_y(nh, 20, 2, u)[:, :]
    # n_cubes here is 300 < 20^2, digits still follow c -> (c % 20, c // 20)
    assert torch.equal(y.cpu(), want)
    jf = torch.rand(M, dtype=torch.float64)
    JF, JF2 = ops.strat_accumulate(jf.to(cuda), offsets)
    oJF, oJF2 = O.strat_accumulate(nh, jf)
    light = nh <= 256  # thread-per-cube sequential sums: bit-identical; the heavy cube is summed by a warp in fp64
    assert torch.equal(JF.cpu()[light], oJF[light]) and torch.equal(JF2.cpu()[light], oJF2[light])
    assert torch.allclose(JF.cpu(), oJF, rtol=1e-13) and torch.allclose(JF2.cpu(), oJF2, rtol=1e-13)


# ---------------------------------------------------------------- Newton-Cotes
@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("rule", ["trapezoid", "simpson", "boole"])
def test_newton_cotes_grid_and_contract(cuda, golden, tag, rule):
    import torchquad_b200 as tq

    g = golden(f"newton_cotes_{tag}")
    integ = {"trapezoid": tq.Trapezoid, "simpson": tq.Simpson, "boole": tq.Boole}[rule]()
    for dim in (1, 2, 3, 4):
        k = f"{rule}_d{dim}"
        dom = dev(g[f"{k}_domain"], cuda)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pts, hs, n = integ.calculate_grid(int(g[f"{k}_N"]), dom)
        assert n == int(g[f"{k}_n"])
        assert torch.equal(hs.cpu(), torch.from_numpy(g[f"{k}_h"]))
        assert torch.equal(pts.cpu(), torch.from_numpy(g[f"{k}_points"]))  # same torch.linspace nodes, same order
        res = integ.calculate_result(dev(g[f"{k}_f"], cuda), dim, n, hs, dom)
        want = float(g[f"{k}_result"])
        assert res.dtype == DT[tag] and res.dim() == 0
        assert abs(float(res) - want) <= (2e-13 if tag == "f64" else 2e-5) * max(1.0, abs(want))
        resv = integ.calculate_result(dev(g[f"{k}_fv"], cuda), dim, n, hs, dom)
        assert resv.shape == (2,)
        assert np.allclose(resv.cpu().numpy(), g[f"{k}_resultv"], rtol=(2e-13 if tag == "f64" else 2e-5))
