"""Host-side logic of the drop-in classes that needs no GPU: input validation, N adjustment, domain set-up,
rule weight patterns, the VEGAS schedule, sharding arithmetic and the built-in integrands' closed forms.
Mirrors the host-side assertions of /root/reference/tests (cited per test)."""
import math
import warnings

import numpy as np
import pytest
import torch

import torchquad_b200 as tq
from oracle import ref_oracle as O
from torchquad_b200 import distributed as tqdist
from torchquad_b200 import integrands as F
from torchquad_b200.integration.utils import (_check_integration_domain, _linspace_with_grads,
                                              _setup_integration_domain)


def test_check_inputs_and_domain_validation():
    """/root/reference/torchquad/integration/base_integrator.py:93-116, utils.py:162-206."""
    chk = tq.BaseIntegrator._check_inputs
    with pytest.raises(ValueError):
        chk(dim=0)
    with pytest.raises(ValueError):
        chk(N=0)
    with pytest.raises(ValueError):
        chk(N=10.0)
    with pytest.raises(ValueError):
        chk(dim=3, integration_domain=[[0, 1]] * 2)
    with pytest.raises(ValueError):
        _check_integration_domain([[1.0, 0.0]])
    with pytest.raises(ValueError):
        _check_integration_domain([[0.0, 1.0, 2.0]])
    with pytest.raises(ValueError):
        _check_integration_domain(torch.tensor([[1.0, 0.0]]))
    with pytest.raises(ValueError):
        _check_integration_domain(torch.zeros(3))
    assert _check_integration_domain([[0, 1], [2, 3]]) == 2
    assert _check_integration_domain(torch.tensor([[0.0, 1.0]] * 4)) == 4


def test_setup_integration_domain_dtype_rules():
    """/root/reference/tests/utils_integration_test.py:123-180 (torch backend rows)."""
    torch.set_default_dtype(torch.float64)
    try:
        d = _setup_integration_domain(2, None, None)
        assert d.dtype == torch.float64 and d.tolist() == [[-1.0, 1.0]] * 2
        d = _setup_integration_domain(1, [[0, 3]], "torch")
        assert d.dtype == torch.float64 and d.shape == (1, 2)
        given = torch.tensor([[0.0, 1.0]], dtype=torch.float32)
        assert _setup_integration_domain(1, given, None) is given
        with pytest.raises(ValueError):
            _setup_integration_domain(2, [[0, 1]], None)
        with pytest.raises(ValueError):
            _setup_integration_domain(1, [[0, 1]], "numpy")
    finally:
        torch.set_default_dtype(torch.float32)


def test_linspace_with_grads_endpoints_exact():
    """/root/reference/tests/utils_integration_test.py:47-72."""
    for rg in (False, True):
        a = torch.tensor(-2.0, dtype=torch.float64, requires_grad=rg)
        b = torch.tensor(7.0, dtype=torch.float64, requires_grad=rg)
        g = _linspace_with_grads(a, b, 11, rg)
        assert g.shape == (11,) and float(g[0]) == -2.0 and float(g[-1]) == 7.0
        assert g.requires_grad == rg


@pytest.mark.parametrize("rule,cls", [("trapezoid", tq.Trapezoid), ("simpson", tq.Simpson), ("boole", tq.Boole)])
def test_adjust_n_and_rule_weights(rule, cls):
    """_adjust_N (simpson.py:54-81, boole.py:56-84) and the 1-D weight patterns against the oracle's
    axis-by-axis stencil (trapezoid.py:28-37, simpson.py:30-46, boole.py:30-48)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for dim, N in [(1, 2), (1, 401), (2, 16), (2, 1000), (3, 1_076_890), (6, 31**6), (6, 33**6), (10, 3**10)]:
            assert cls._adjust_N(dim, N) == O.nc_adjust_n(rule, dim, N)
    n = {"trapezoid": 8, "simpson": 9, "boole": 9}[rule]
    w = cls._rule_weights_1d(n, torch.float64, "cpu")
    f = torch.rand(n, dtype=torch.float64)
    hs = torch.tensor([0.37], dtype=torch.float64)
    want = O.nc_result(rule, f, 1, n, hs)
    got = (f * w).sum() * hs[0] / cls._rule_denominator
    assert abs(float(got) - float(want)) < 1e-15 * max(1.0, abs(float(want)))
    # tensor product: sum_p f(p) prod_d w[i_d] prod_d h_d/c == oracle in 3-D
    f3 = torch.rand(n**3, dtype=torch.float64)
    hs3 = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64)
    W = torch.einsum("i,j,k->ijk", w, w, w).reshape(-1)
    got3 = (f3 * W).sum() * torch.prod(hs3 / cls._rule_denominator)
    assert abs(float(got3) - float(O.nc_result(rule, f3, 3, n, hs3))) < 1e-14
    if rule != "trapezoid":
        assert cls._get_minimal_N(3) == {"simpson": 27, "boole": 125}[rule]


def test_vegas_schedule_matches_oracle_decisions():
    """_check_abort_conditions / weighted mean (vegas.py:161-209,318-362) on synthetic iteration records:
    the host-side numpy arithmetic must take the same decisions as the reference's tensor arithmetic."""
    rng = np.random.default_rng(0)
    for dt, npdt in [(torch.float64, np.float64), (torch.float32, np.float32)]:
        for trial in range(20):
            res = [1.0 + 0.01 * rng.standard_normal() for _ in range(5)]
            sig = [abs(1e-4 * (1 + rng.standard_normal())) for _ in range(5)]
            if trial % 7 == 0:
                sig[2] = 0.0
            run = O.VegasRun(lambda x: x[:, 0], 2, 100_000, torch.tensor([[0.0, 1.0]] * 2, dtype=dt), None)
            run.results = [torch.tensor(r, dtype=dt) for r in res]
            run.sigma2 = [torch.tensor(s, dtype=dt) for s in sig]
            run.it, run.fevals = 5, 20_000 + trial * 1000
            v = tq.VEGAS()
            v.dtype, v._np, v.device = dt, npdt, torch.device("cpu")
            v.results = [torch.tensor(r, dtype=dt) for r in res]
            v.sigma2 = [torch.tensor(s, dtype=dt) for s in sig]
            v.it, v._nr_of_fevals, v.N = 5, run.fevals, 100_000
            v._max_iterations, v._eps_rel, v._eps_abs = 20, 0, 0
            v._starting_N = v._N_increment = 100_000 // 25
            v._status_used, v._status_buf = 0, torch.zeros((4, 4), dtype=torch.int32)
            want_mean = run.result()
            assert torch.equal(v._get_result(), want_mean)
            stop_ref = run.check_abort()
            stop = v._check_abort_conditions()
            assert stop == stop_ref and v._starting_N == run.starting_N, (trial, dt)
            # the same checkpoint as taken by the C++ loop of tq_vegas_run_fused (host-only entry point)
            import ctypes

            from torchquad_b200 import _lib

            arr = ctypes.c_double * 5
            start, mean, flag = ctypes.c_int64(100_000 // 25), ctypes.c_double(), ctypes.c_int32()
            _lib.call("tq_vegas_schedule", arr(*[float(npdt(r)) for r in res]), arr(*[float(npdt(s)) for s in sig]), 5,
                      _lib.dtype_code(dt), 0.0, 0.0, 100_000, run.fevals, 5, 20, 100_000 // 25, ctypes.byref(start),
                      ctypes.byref(mean), ctypes.byref(flag))
            assert bool(flag.value) == stop_ref and (stop_ref or start.value == run.starting_N), (trial, dt)
            assert mean.value == float(want_mean)


def test_shard_ranges_partition_exactly():
    for total in [0, 1, 7, 1000, 10**9 + 7]:
        for world in [1, 2, 3, 8]:
            spans = [tqdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_builtin_integrands_closed_forms_and_struct():
    a, u = [1.3, 0.7, 2.1], [0.3, 0.6, 0.45]
    for cls, fam in [(F.GenzOscillatory, "oscillatory"), (F.GenzProductPeak, "product_peak"), (F.GenzCornerPeak, "corner_peak"),
                     (F.GenzGaussian, "gaussian"), (F.GenzC0, "c0"), (F.GenzDiscontinuous, "discontinuous")]:
        fn = cls(3, a=a, u=u)
        assert abs(fn.exact() - O.genz_exact(fam, a, u)) <= 1e-14 * abs(fn.exact())
        x = torch.rand(64, 3, dtype=torch.float64)
        assert torch.allclose(fn(x), O.genz(fam, x, a, u), rtol=1e-14, atol=0)
        s = fn.to_struct([0.0, 1.0, 2.0], [1.0, 2.0, 3.0], 6.0)
        assert s.dim == 3 and s.family == F.FAMILY[fn.family] and list(s.a)[:3] == a and s.scale == 6.0
        assert list(s.start)[:3] == [0.0, 1.0, 2.0] and list(s.size)[:3] == [1.0, 2.0, 3.0]
    assert abs(F.SumOfSines(10).exact() - 20 * math.sin(0.5) ** 2) < 1e-15
    assert abs(F.Polynomial(2, [1.0, 2.0, 3.0]).exact() - 2 * (1 + 1 + 1)) < 1e-15
    x = torch.rand(10, 4, dtype=torch.float64)
    assert torch.allclose(F.ProductOfCosines(4)(x), O.test_integrand("product_cos", x))
    assert torch.allclose(F.SumOfExp(4)(x), O.test_integrand("exponential", x))
    with pytest.raises(ValueError):
        F.GenzGaussian(40)
    with pytest.raises(ValueError):
        F.GenzGaussian(3, a=[1.0, 2.0])


def test_stratification_configuration_matches_reference_formula():
    """N_strat / N_cubes / V_cubes (vegas_stratification.py:27-31) and map size (vegas.py:117) for BASELINE configs."""
    for N, dim, ns, ni in [(10**6, 4, 10, 4000), (2_500_000_000, 8, 8, 10**7), (10**10, 16, 3, 4 * 10**7)]:
        inc = N // 25
        assert O.strat_config(inc, dim)[0] == ns and max(2, inc // 10) == ni
        s = tq.VEGASStratification.__new__(tq.VEGASStratification)
        n_strat = int((inc / 4.0) ** (1.0 / dim))
        assert n_strat == ns


def test_record_layout_selection_is_by_table_size():
    """Large maps switch to {x, dx, weight, count} records (include/tqb200.h, TQ_EDGES_RECORDS): the choice is a pure
    function of the table size, 32 B (fp64) / 16 B (fp32) per bin against `records_min_bytes`."""
    from torchquad_b200.integration.vegas_map import VEGASMap

    class Probe(VEGASMap):
        def __init__(self, dim, ni, dtype):  # no device tensors: only the fields wants_records reads
            self.dim, self.N_intervals, self.dtype = dim, ni, dtype

    assert VEGASMap.records_min_bytes == 48 << 20
    assert not Probe(4, 4000, torch.float64).wants_records()           # configs[0]: 512 KB
    assert not Probe(8, 4096, torch.float64).wants_records()           # capped maps stay on pair tables
    assert Probe(8, 10_000_000, torch.float64).wants_records()         # configs[3] at the reference's size: 2.56 GB
    assert Probe(16, 4_194_304, torch.float32).wants_records()         # 1 GiB
    assert Probe(8, 196_608, torch.float64).wants_records() and not Probe(8, 196_607, torch.float64).wants_records()
    p = Probe(8, 10_000_000, torch.float64)
    p.records_min_bytes = None
    assert not p.wants_records()


def test_reciprocal_division_is_correctly_rounded(tmp_path):
    """csrc/common.cuh `div_by_const`: q0 = RN(a * RN(1/b)); q = fma(fma(-q0, b, a), RN(1/b), q0) must equal the IEEE quotient
    a / b bit for bit for the operands of the stratified sampling (a = digit + u, u a multiple of 2^-24 / 2^-53 in [0, 1),
    b = N_strat <= 1000).  The same three operations compiled for the host (gcc, hardware FMA, contraction off): every fp32
    uniform for a few N_strat, 2e5 random operands per N_strat in both precisions."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "div.c"
    src.write_text(r"""
#include <math.h>
#include <stdio.h>
#include <stdint.h>
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main(void) {
    long bad = 0, n = 0;
    int exhaustive[] = {3, 7, 10, 1000};
    for (int bi = 0; bi < 4; ++bi) {
        const int b = exhaustive[bi];
        const float bf = (float)b, y = 1.0f / bf;
        const int p = b - 1;
        for (uint32_t k = 0; k < (1u << 24); ++k) {
            const float a = (float)p + (float)k * 5.9604644775390625e-08f, q0 = a * y;
            bad += fmaf(fmaf(-q0, bf, a), y, q0) != a / bf;
            ++n;
        }
    }
    for (int b = 1; b <= 1000; ++b) {
        const double bd = (double)b, y = 1.0 / bd;
        const float bf = (float)b, yf = 1.0f / bf;
        for (int i = 0; i < 200000; ++i) {
            const int p = (int)(rnd() % (uint64_t)b);
            const double a = (double)p + (double)(rnd() >> 11) * 1.1102230246251565e-16, q0 = a * y;
            bad += fma(fma(-q0, bd, a), y, q0) != a / bd;
            const float af = (float)p + (float)(rnd() >> 40) * 5.9604644775390625e-08f, q0f = af * yf;
            bad += fmaf(fmaf(-q0f, bf, af), yf, q0f) != af / bf;
            n += 2;
        }
    }
    printf("%ld %ld\n", bad, n);
    return 0;
}
""")
    exe = tmp_path / "div"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", str(exe), str(src), "-lm"], check=True)
    bad, n = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert bad == 0 and n > 4 * 10**8


def test_cube_shard_inverse_matches_the_deal():
    """The band sweeps of a sharded run walk GLOBAL cube ids and keep the cubes this rank owns (CubeShard::local_cube,
    csrc/fused.cu): owner and local id must invert distributed.global_cube_ids for every rank, including the incomplete
    last round and the trailing partial block."""
    from torchquad_b200 import distributed as tqdist

    def local_cube(c, lb, rank, world):  # the device formula, restated
        blk = c >> lb
        rnd, pos = divmod(blk, world)
        return (rank + tqdist._skew(rnd)) % world == pos, (rnd << lb) + (c & ((1 << lb) - 1))

    for n_cubes, world in [(8**8, 8), (3**16 // 9, 4), (10007, 2), (65536 + 17, 8), (6561, 3), (8**5, 5)]:
        seen = 0
        for rank in range(world):
            shard = tqdist.cube_shard(n_cubes, rank, world)
            assert shard is not None
            lb, n_local = shard
            ids = tqdist.global_cube_ids(n_local, lb, rank, world).tolist()
            step = max(1, n_local // 5000)
            for l in list(range(0, n_local, step)) + [n_local - 1]:
                mine, l2 = local_cube(ids[l], lb, rank, world)
                assert mine and l2 == l, (n_cubes, world, rank, l)
            other = (rank + 1) % world
            for l in range(0, n_local, max(1, n_local // 200)):
                assert not local_cube(ids[l], lb, other, world)[0]
            seen += n_local
        assert seen == n_cubes


def test_tile_kernel_band_contains_every_reachable_bin():
    """fused_vegas_tile_kernel stages, per slow dimension, the bins [blo, blo + band_w) with blo = max(0, p*Ni//Ns - 1) and
    band_w = ceil(Ni/Ns) + 3 (csrc/fused.cu).  Every bin a sample of digit p can reach -- k = floor(fl(fl((p + u)/Ns) * Ni)),
    the arithmetic of vegas_stratification.py:140-165 + vegas_map.py:76-85 in the working precision -- must lie inside, so
    the kernel's out-of-band fallback is never taken.  Checked in fp32 for the extremes of u and random u."""
    import numpy as np

    rng = np.random.default_rng(0)
    u_ext = np.array([0.0, 2.0**-24, 0.5, 1.0 - 2.0**-24], dtype=np.float32)
    for ns in (2, 3, 5, 7, 8, 13, 56, 1000):
        for ni in (2, 97, 777, 1000, 4096, 65536, 1 << 20):
            band_w = (ni + ns - 1) // ns + 3
            p = np.arange(ns, dtype=np.int64)
            u = np.concatenate([u_ext, rng.integers(0, 1 << 24, 64).astype(np.float32) * np.float32(2.0**-24)])
            y = ((p[:, None].astype(np.float32) + u[None, :]) / np.float32(ns)).astype(np.float32)
            y[y >= 1.0] = np.float32(0.999999)
            k = np.floor(y * np.float32(ni)).astype(np.int64)
            k = np.clip(k, 0, ni - 1)
            blo = np.maximum(0, (p * ni) // ns - 1)[:, None]
            assert np.all(k >= blo) and np.all(k < blo + band_w), (ns, ni)
