"""The two-kernel VEGAS pass for callback integrands (csrc/vegas_unfused.cu): tq_vegas_sample_map and
tq_vegas_accumulate_regen against the materialised pipeline they replace (tq_vegas_strat_sample ->
tq_vegas_map_forward_packed -> tq_vegas_accumulate_fused), which the other GPU tests pin to the oracle / the reference's
fixtures.  x, jac, jf must be BIT-identical, histogram counts exact, weights equal up to the order of the atomics.
References: vegas_stratification.py:140-165, vegas_map.py:44-111, vegas.py:104-112,230-290."""
import pytest
import torch

import torchquad_b200 as tq
from torchquad_b200 import ops
from torchquad_b200.integration.vegas_map import VEGASMap

pytestmark = pytest.mark.gpu
DT = {"f32": torch.float32, "f64": torch.float64}


def _adapted_map(ni, dim, dt, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    vm = VEGASMap(ni, dim, "torch", dt, device=dev)
    vm.weights.copy_((torch.rand(vm.weights.shape, generator=g, dtype=torch.float64) ** 3 + 0.01).to(dt))
    vm.counts.fill_(1)
    vm.update_map()
    return vm


def _offsets(n_strat, dim, nh_mean, dt, dev, seed=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    C = n_strat**dim
    dh = (torch.rand(C, generator=g, dtype=torch.float64) ** 4 + 1e-3)
    dh = (dh / dh.sum()).to(dt).to(dev)
    _nh, offsets = ops.strat_nh(dh, nh_mean * C)
    return offsets


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("dim,n_strat,ni,nh_mean", [(1, 50, 64, 3.0), (3, 6, 333, 4.5), (8, 3, 1000, 6.0), (10, 2, 97, 9.0), (17, 2, 50, 2.5)])
def test_sample_map_is_bit_identical_to_the_materialised_pipeline(cuda, tag, dim, n_strat, ni, nh_mean):
    dt = DT[tag]
    vm = _adapted_map(ni, dim, dt, cuda)
    offsets = _offsets(n_strat, dim, nh_mean, dt, cuda)
    M = int(offsets[-1])
    dom = torch.stack([torch.linspace(-1.0, 0.5, dim), torch.linspace(1.0, 3.0, dim)], dim=1).to(dt).to(cuda).contiguous()
    seed, call = 1234567, 5
    for begin, end in [(0, M), (M // 3 + 1, M - 7), (M // 2, M // 2)]:
        y = ops.strat_sample(offsets, n_strat, dim, dt, begin, end, seed=seed, call_idx=call)
        x_ref, jac_ref, _ = ops.map_forward_packed(y, vm.packed_edges(), dom)
        x, jac = ops.sample_map(offsets, n_strat, dim, dt, begin, end, seed, call, dom, edges_packed=vm.packed_edges())
        assert torch.equal(x, x_ref) and torch.equal(jac, jac_ref)
        # record layout of large maps: same values
        x2, jac2 = ops.sample_map(offsets, n_strat, dim, dt, begin, end, seed, call, dom, records=vm.records(), n_intervals=ni)
        assert torch.equal(x2, x_ref) and torch.equal(jac2, jac_ref)
    # warm-up pass: y = u * 0.999999 of the row-keyed stream (vegas.py:230-233)
    rows, row0 = 5003, 77
    yw = ops.philox_uniform(rows, dim, dt, cuda, seed, call + 1, row0) * 0.999999
    xw_ref, jw_ref, _ = ops.map_forward_packed(yw, vm.packed_edges(), dom)
    xw, jw = ops.sample_map(None, 1, dim, dt, row0, row0 + rows, seed, call + 1, dom, edges_packed=vm.packed_edges())
    assert torch.equal(xw, xw_ref) and torch.equal(jw, jw_ref)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("dim,n_strat,ni,nh_mean", [(2, 20, 200, 3.0), (8, 3, 1000, 6.0), (19, 2, 40, 2.5)])
def test_accumulate_regen_matches_accumulate_fused(cuda, tag, dim, n_strat, ni, nh_mean):
    dt = DT[tag]
    vm = _adapted_map(ni, dim, dt, cuda)
    offsets = _offsets(n_strat, dim, nh_mean, dt, cuda)
    M = int(offsets[-1])
    seed, call, volume = 99, 3, 1.75
    begin, end = 11, M - 5
    rows = end - begin
    g = torch.Generator(device="cpu").manual_seed(4)
    f = torch.randn(rows, generator=g, dtype=torch.float64).to(dt).to(cuda)
    jac = (torch.rand(rows, generator=g, dtype=torch.float64) + 0.5).to(dt).to(cuda)
    y = ops.strat_sample(offsets, n_strat, dim, dt, begin, end, seed=seed, call_idx=call)
    w_ref = torch.zeros_like(vm.weights)
    c_ref = torch.zeros_like(vm.counts)
    jf_ref = ops.accumulate_fused(y, f, jac, volume, w_ref, c_ref, want_jf=True)
    tol = 1e-12 if dt == torch.float64 else 2e-5
    # pair table
    h = vm.hist_pairs()
    h.zero_()
    jf, jf2 = ops.accumulate_regen(offsets, n_strat, dim, begin, end, ni, f, jac, volume, seed, call, hist_pairs=h, want_jf2=True)
    assert torch.equal(jf, jf_ref) and torch.equal(jf2, jf_ref * jf_ref)
    assert torch.equal(h[..., 1].to(torch.int64), c_ref) and int(c_ref.sum()) == rows * dim
    assert float(((h[..., 0] - w_ref.double()).abs() / w_ref.double().abs().clamp_min(1e-300)).max()) <= tol
    h.zero_()
    # weights / counts arrays
    w = torch.zeros_like(vm.weights)
    c = torch.zeros_like(vm.counts)
    jf_b, _ = ops.accumulate_regen(offsets, n_strat, dim, begin, end, ni, f, jac, volume, seed, call, weights=w, counts=c)
    assert torch.equal(jf_b, jf_ref) and torch.equal(c, c_ref)
    assert float(((w - w_ref).abs() / w_ref.abs().clamp_min(1e-30)).max()) <= tol
    # records of a large map
    rec = vm.records()
    jf_c, _ = ops.accumulate_regen(offsets, n_strat, dim, begin, end, ni, f, jac, volume, seed, call, records=rec)
    vm._reset_weight()
    vm.unpack_records()
    assert torch.equal(jf_c, jf_ref) and torch.equal(vm.counts, c_ref)
    assert float(((vm.weights - w_ref).abs() / w_ref.abs().clamp_min(1e-30)).max()) <= tol
    # no target: jf only
    jf_d, _ = ops.accumulate_regen(offsets, n_strat, dim, begin, end, ni, f, jac, volume, seed, call)
    assert torch.equal(jf_d, jf_ref)
    # warm-up rows (row-keyed stream)
    rows_w, row0 = min(4001, rows), 13
    yw = ops.philox_uniform(rows_w, dim, dt, cuda, seed, call + 1, row0) * 0.999999
    w_ref.zero_()
    c_ref.zero_()
    ops.accumulate_fused(yw, f[:rows_w].contiguous(), jac[:rows_w].contiguous(), volume, w_ref, c_ref, want_jf=False)
    ops.accumulate_regen(None, 1, dim, row0, row0 + rows_w, ni, f[:rows_w].contiguous(), jac[:rows_w].contiguous(), volume, seed,
                         call + 1, hist_pairs=h, want_jf=False)
    assert torch.equal(h[..., 1].to(torch.int64), c_ref)
    assert float(((h[..., 0] - w_ref.double()).abs() / w_ref.double().abs().clamp_min(1e-300)).max()) <= tol
    h.zero_()


def _run(fn, dim, N, dt, dev, regen, seed=3, native=True, **kw):
    v = tq.VEGAS()
    v.regenerate_samples = regen
    v.native_loop = native
    for k, val in kw.items():
        setattr(v, k, val)
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=dev)
    r = float(v.integrate(fn, dim, N=N, integration_domain=dom, seed=seed))
    return r, v


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("native", [True, False])
def test_whole_run_with_and_without_materialised_samples(cuda, tag, native):
    """Same seed => same samples, same schedule; results agree to the rounding of the histogram atomics."""
    dt = DT[tag]
    fn = lambda x: torch.exp(-torch.sum(16.0 * (x - 0.4) ** 2, dim=1)) + 0.05  # noqa: E731
    for dim, N in [(3, 300_000), (5, 2_000_000)]:
        a, va = _run(fn, dim, N, dt, cuda, regen=False, native=native)
        b, vb = _run(fn, dim, N, dt, cuda, regen=True, native=native)
        assert vb._regen and not va._regen
        assert va.it == vb.it
        tol = 1e-9 if dt == torch.float64 else 2e-3
        assert abs(a - b) <= tol * abs(a), (a, b)
        if dt == torch.float64:
            assert va._nr_of_fevals == vb._nr_of_fevals


def test_large_map_uses_band_sweeps_and_matches(cuda):
    """Maps beyond L2 (forced here by the threshold): x from the pair table, jf^2 rows + tq_vegas_hist_sweep."""
    fn = lambda x: torch.cos(0.9 + torch.sum(0.5 * x, dim=1)) + 1.5  # noqa: E731
    default = VEGASMap.records_min_bytes
    VEGASMap.records_min_bytes = 0
    try:
        a, va = _run(fn, 4, 3_000_000, torch.float64, cuda, regen=False, native=False)
        b, vb = _run(fn, 4, 3_000_000, torch.float64, cuda, regen=True, native=False)
        assert vb._regen_sweep >= 1 and vb.map.wants_records()
        assert va.it == vb.it and va._nr_of_fevals == vb._nr_of_fevals
        assert abs(a - b) <= 1e-9 * abs(a), (a, b)
        # and through the record table (sweeps disabled)
        c, vc = _run(fn, 4, 3_000_000, torch.float64, cuda, regen=True, native=True)
        assert abs(a - c) <= 1e-9 * abs(a), (a, c)
    finally:
        VEGASMap.records_min_bytes = default


def test_gradient_through_integrand_still_uses_the_samples(cuda):
    """An integrand whose values carry a graph falls back to the reference's expressions (gradient_test.py:162-259)."""
    p = torch.tensor(2.0, dtype=torch.float64, device=cuda, requires_grad=True)
    dom = torch.tensor([[0.0, 1.0]] * 2, dtype=torch.float64, device=cuda)
    v = tq.VEGAS()
    r = v.integrate(lambda x: p * torch.sum(x * x, dim=1), 2, N=50_000, integration_domain=dom, seed=1)
    (g,) = torch.autograd.grad(r, p)
    assert abs(float(r) - 2.0 * 2.0 / 3.0) < 2e-2 and abs(float(g) - 2.0 / 3.0) < 1e-2


def _cube_keyed_uniforms(seed, call, nh, dim, dt):
    """The cube-keyed Philox stream restated with the oracle's generator: counter = (cube, index in cube, block, call)."""
    import numpy as np

    from oracle import ref_oracle as O

    lanes = 4 if dt == torch.float32 else 2
    nblk = (dim + lanes - 1) // lanes
    cube = np.repeat(np.arange(nh.shape[0], dtype=np.uint32), nh.numpy())
    idx = np.concatenate([np.arange(int(n), dtype=np.uint32) for n in nh.tolist()])
    ctr = np.empty((cube.shape[0], nblk, 4), dtype=np.uint32)
    ctr[..., 0], ctr[..., 1] = cube[:, None], idx[:, None]
    ctr[..., 2], ctr[..., 3] = np.arange(nblk, dtype=np.uint32)[None, :], np.uint32(call)
    out = O.philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32))
    if dt == torch.float32:
        u = ((out >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)).reshape(cube.shape[0], nblk * 4)[:, :dim]
    else:
        lo, hi = out[..., 0::2].astype(np.uint64), out[..., 1::2].astype(np.uint64)
        u = ((((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64) * 2.0**-53).reshape(cube.shape[0], nblk * 2)[:, :dim]
    return torch.from_numpy(np.ascontiguousarray(u))


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_two_kernel_pass_against_the_oracle_directly(cuda, tag):
    """tq_vegas_sample_map / tq_vegas_accumulate_regen against oracle/ref_oracle.py (the reference's ATen expressions on the
    CPU) on the same Philox uniforms: x, jac, jf bit-identical, counts exact, weights to the order of the atomics."""
    from oracle import ref_oracle as O

    dt = DT[tag]
    dim, n_strat, ni, seed, call, volume = 5, 4, 300, 424242, 9, 2.5
    vm = _adapted_map(ni, dim, dt, cuda, seed=3)
    xe, dxe = vm.x_edges.cpu(), vm.dx_edges.cpu()
    dom = torch.tensor([[-1.0, 1.0], [0.0, 2.0], [0.5, 0.75], [-3.0, -1.0], [0.0, 1.0]], dtype=dt)
    starts, sizes = dom[:, 0], dom[:, 1] - dom[:, 0]
    offsets = _offsets(n_strat, dim, 3.7, dt, cuda, seed=5)
    nh = (offsets[1:] - offsets[:-1]).cpu()
    M = int(offsets[-1])
    # stratified pass
    y = O.strat_get_y(nh, n_strat, dim, _cube_keyed_uniforms(seed, call, nh, dim, dt))
    x_ref = O.map_get_x(y, xe, dxe) * sizes + starts
    jac_ref = O.map_get_jac(y, dxe)
    x, jac = ops.sample_map(offsets, n_strat, dim, dt, 0, M, seed, call, dom.to(cuda), edges_packed=vm.packed_edges())
    assert torch.equal(x.cpu(), x_ref) and torch.equal(jac.cpu(), jac_ref)
    f = torch.sin(x_ref.sum(dim=1)) + 1.25
    jf_ref = (f * volume) * jac_ref
    w_ref, c_ref = O.map_reset(ni, dim, dt)
    O.map_accumulate(w_ref, c_ref, y, jf_ref * jf_ref)
    h = vm.hist_pairs()
    h.zero_()
    jf, _ = ops.accumulate_regen(offsets, n_strat, dim, 0, M, ni, f.to(cuda), jac, volume, seed, call, hist_pairs=h)
    assert torch.equal(jf.cpu(), jf_ref)
    assert torch.equal(h[..., 1].cpu().to(torch.int64), c_ref)
    tol = 1e-12 if dt == torch.float64 else 2e-5
    assert float(((h[..., 0].cpu() - w_ref.double()).abs() / w_ref.double().abs().clamp_min(1e-300)).max()) <= tol
    h.zero_()
    # warm-up pass: y = u * 0.999999 from the row-keyed stream (vegas.py:230-233)
    rows, row0 = 3001, 11
    yw = O.philox_uniform(seed, call + 1, row0, rows, dim, dt) * 0.999999
    xw, jw = ops.sample_map(None, 1, dim, dt, row0, row0 + rows, seed, call + 1, dom.to(cuda), edges_packed=vm.packed_edges())
    assert torch.equal(xw.cpu(), O.map_get_x(yw, xe, dxe) * sizes + starts) and torch.equal(jw.cpu(), O.map_get_jac(yw, dxe))
