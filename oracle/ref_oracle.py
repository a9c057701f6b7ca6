"""CPU oracle for the torchquad sampling-and-reduction hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this
module; `torchquad_b200` never does (its ops raise when the CUDA library is
missing -- there is no CPU fallback).

What this is: a functional restatement of the reference algorithm (esa/torchquad
v0.5.0, `/root/reference`, pure Python over ATen via `autoray`) written against
plain CPU torch tensors.  The reference's arithmetic *is* ATen's CPU kernels
(`torch.floor`, `scatter_add_`, `cumsum`, `linspace`, ...), so using the same
kernels makes the oracle bit-identical to the reference on CPU instead of merely
close; the state is explicit (no classes holding hidden tensors) so the CUDA
kernels can be compared step by step on identical injected samples.

Parity pinning (see tests/test_oracle_pinning.py, oracle/make_golden.py):
  * every function below is compared bitwise with the UNMODIFIED reference
    imported from /root/reference when that directory exists (build container);
  * the same comparisons are frozen as fixtures under tests/golden/*.npz so the
    pin also holds on the GPU box where /root/reference does not exist;
  * the reference's own golden vectors (tests/vegas_map_test.py:28-48,84-100 etc.)
    are re-checked against this oracle.

Citations are `file:line` relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
# SC'11; Random123 v1.14 `philox.h`).  The reference has no counter-based
# generator (rng.py:119-125 is `torch.rand` on the global generator); this is the
# published algorithm the product's `philox` kernels implement, restated in numpy
# and pinned by the Random123 known-answer vectors (tests/test_philox.py).
# --------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(ctr, key):
    """ctr: uint32[...,4], key: uint32[...,2] -> uint32[...,4] (10 rounds)."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint64) for i in range(4))
    k0 = np.broadcast_to(key[..., 0], c0.shape).astype(np.uint32)
    k1 = np.broadcast_to(key[..., 1], c0.shape).astype(np.uint32)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0
            p1 = PHILOX_M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & mask
            hi1, lo1 = p1 >> np.uint64(32), p1 & mask
            n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
            n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def philox_uniform(seed, call_idx, row0, rows, dim, dtype):
    """The product's uniform stream u[row, d], restated (DESIGN.md "Philox layout").

    key = (seed_lo, seed_hi); counter = (row_lo, row_hi, block, call_idx) where
    block = d // 4 (float32: 4 lanes per block, u = (x >> 8) * 2^-24) or d // 2
    (float64: 2 lanes per block, u = ((hi<<32|lo) >> 11) * 2^-53, lo = word 2j,
    hi = word 2j+1).  Every value is a pure function of (seed, call, global row,
    d), so any rank/row partition of the same call draws identical numbers.
    """
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)
    lanes = 4 if dtype == torch.float32 else 2
    nblk = (dim + lanes - 1) // lanes
    r = np.arange(row0, row0 + rows, dtype=np.uint64)
    ctr = np.empty((rows, nblk, 4), dtype=np.uint32)
    ctr[..., 0] = (r & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    ctr[..., 1] = (r >> np.uint64(32)).astype(np.uint32)[:, None]
    ctr[..., 2] = np.arange(nblk, dtype=np.uint32)[None, :]
    ctr[..., 3] = np.uint32(call_idx & 0xFFFFFFFF)
    out = philox4x32_10(ctr, key)
    if dtype == torch.float32:
        u = (out >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)
        u = u.reshape(rows, nblk * 4)[:, :dim]
    else:
        lo = out[..., 0::2].astype(np.uint64)
        hi = out[..., 1::2].astype(np.uint64)
        u = (((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64) * 2.0**-53
        u = u.reshape(rows, nblk * 2)[:, :dim]
    return torch.from_numpy(np.ascontiguousarray(u))


# --------------------------------------------------------------------------
# VEGAS map  (torchquad/integration/vegas_map.py)
# --------------------------------------------------------------------------
def map_init(n_intervals, dim, dtype):
    """Fresh map: uniform edges, dx = 1/Ni (vegas_map.py:32-42)."""
    dx = torch.ones((dim, n_intervals), dtype=dtype) / n_intervals
    x1 = torch.linspace(0.0, 1.0, n_intervals + 1, dtype=dtype)
    x = torch.repeat_interleave(x1.reshape(1, -1), dim, dim=0)
    w, c = map_reset(n_intervals, dim, dtype)
    return x, dx, w, c


def map_reset(n_intervals, dim, dtype):
    """vegas_map.py:174-183."""
    return (
        torch.zeros((dim, n_intervals), dtype=dtype),
        torch.zeros((dim, n_intervals), dtype=torch.int64),
    )


def interval_id(y, n_intervals):
    """k = int64(floor(y * float(Ni)))  (vegas_map.py:76-85)."""
    return torch.floor(y * float(n_intervals)).to(torch.int64)


def interval_offset(y, n_intervals):
    """o = y*Ni - floor(y*Ni)  (vegas_map.py:87-97)."""
    t = y * float(n_intervals)
    return t - torch.floor(t)


def map_get_x(y, x_edges, dx_edges):
    """x_i = xe[i,k_i] + dxe[i,k_i]*o_i  (vegas_map.py:44-58)."""
    ni = dx_edges.shape[1]
    k, o = interval_id(y, ni), interval_offset(y, ni)
    cols = [x_edges[i, k[:, i]] + dx_edges[i, k[:, i]] * o[:, i] for i in range(y.shape[1])]
    return torch.stack(cols, dim=1)


def map_get_jac(y, dx_edges):
    """jac = prod_i Ni*dxe[i,k_i], multiplied left to right from 1 (vegas_map.py:60-74)."""
    ni = dx_edges.shape[1]
    k = interval_id(y, ni)
    jac = torch.ones([y.shape[0]], dtype=y.dtype)
    for i in range(y.shape[1]):
        jac = jac * (ni * dx_edges[i][k[:, i]])
    return jac


def map_accumulate(weights, counts, y, jf2):
    """weights[i,k]+=jf2, counts[i,k]+=1 in place (vegas_map.py:99-111)."""
    k = interval_id(y, weights.shape[1])
    ones = torch.ones(jf2.shape, dtype=counts.dtype)
    for i in range(y.shape[1]):
        weights[i].scatter_add_(0, k[:, i], jf2)
        counts[i].scatter_add_(0, k[:, i], ones)


def smooth_map(weights, counts, alpha):
    """Average, zero-count fill (<=10 hops), [1,6,1]/8 smoothing, compression
    (vegas_map.py:113-172).  Returns None when a dimension sums to zero."""
    w = weights.clone()
    z = counts == 0
    if bool(z.any()):
        nz = ~z
        w[nz] = w[nz] / counts[nz]
        z = z.clone()
        for _ in range(10):
            w[:, :-1] = torch.where(z[:, :-1], w[:, 1:], w[:, :-1])
            z[:, :-1] = z[:, :-1] & z[:, 1:]
            w[:, 1:] = torch.where(z[:, 1:], w[:, :-1], w[:, 1:])
            z[:, 1:] = z[:, 1:] & z[:, :-1]
            if not bool(z.any()):
                break
    else:
        w = w / counts
    dim, ni = w.shape
    sums = torch.sum(w, dim=1).reshape(dim, 1)
    if bool((sums == 0.0).any()):
        return None
    d = torch.cat(
        [
            7.0 * w[:, 0:1] + w[:, 1:2],
            w[:, :-2] + 6.0 * w[:, 1:-1] + w[:, 2:],
            w[:, ni - 2 : ni - 1] + 7.0 * w[:, ni - 1 : ni],
        ],
        dim=1,
    )
    d = d / (8.0 * sums)
    nzd = d != 0
    d[nzd] = ((d[nzd] - 1.0) / torch.log(d[nzd])) ** alpha
    return d


def map_update(x_edges, dx_edges, weights, counts, alpha=0.5):
    """Equal-mass rebinning (vegas_map.py:185-261).

    Returns (x_edges, dx_edges, status) with status in {"ok", "skipped", "repaired"};
    the caller resets weights/counts afterwards like the reference (:196,:261).
    """
    sm = smooth_map(weights, counts, alpha)
    if sm is None:
        return x_edges, dx_edges, "skipped"
    x_edges, dx_edges = x_edges.clone(), dx_edges.clone()
    dim, ni = sm.shape
    delta = torch.sum(sm, dim=1) / ni
    status = "ok"
    for i in range(dim):
        dd = delta[i]
        mult = (torch.cumsum(sm[i, :-1].to(torch.float64), dim=0) / dd).to(torch.int64)
        num = torch.zeros([ni + 1], dtype=torch.int64)
        num.scatter_add_(0, mult, torch.ones(mult.shape, dtype=torch.int64))
        val = torch.zeros([ni + 1], dtype=sm.dtype)
        # The reference passes the full-length row as source; scatter_add_ only
        # consumes index.numel() = Ni-1 entries (vegas_map.py:225-226).
        val.scatter_add_(0, mult, sm[i])
        idx = torch.cumsum(num[:-2], dim=0)
        acc = torch.cumsum(dd - val[:-2], dim=0)
        x_edges[i][1:-1] = x_edges[i][idx] + acc / sm[i][idx] * dx_edges[i][idx]
        fin = torch.isfinite(x_edges[i])
        if not bool(fin.all()):
            status = "repaired"
            mid = 0.5 * (x_edges[i][:-2] + x_edges[i][2:])
            x_edges[i][1:-1] = torch.where(fin[1:-1], x_edges[i][1:-1], mid)
            if not bool(torch.isfinite(x_edges[i]).all()):
                raise RuntimeError("Could not replace all infinite edges")
        dx_edges[i] = x_edges[i][1:] - x_edges[i][:-1]
    return x_edges, dx_edges, status


# --------------------------------------------------------------------------
# VEGAS stratification  (torchquad/integration/vegas_stratification.py)
# --------------------------------------------------------------------------
def strat_config(n_increment, dim):
    """N_strat, N_cubes, V_cubes (vegas_stratification.py:27-31)."""
    ns = int((n_increment / 4.0) ** (1.0 / dim))
    ns = 1000 if ns > 1000 else ns
    return ns, ns**dim, (1.0 / ns) ** dim


def strat_init(n_cubes, dtype):
    """dh = 1/C (vegas_stratification.py:41)."""
    return torch.ones([n_cubes], dtype=dtype) * 1.0 / n_cubes


def strat_get_nh(dh, nevals_exp):
    """nh = int64(max(2, floor(dh*nevals_exp)))  (vegas_stratification.py:92-103)."""
    return torch.clamp(torch.floor(dh * nevals_exp), min=2).to(torch.int64)


def strat_cube_digits(n_cubes, n_strat, dim):
    """digit_d(c) = (c // Ns^d) % Ns, dim 0 fastest (vegas_stratification.py:105-138)."""
    c = torch.arange(n_cubes, dtype=torch.int64).reshape(-1, 1)
    strides = n_strat ** torch.arange(dim, dtype=torch.int64)
    p = torch.div(c, strides, rounding_mode="floor")
    p[:, :-1] = p[:, :-1] - n_strat * p[:, 1:]
    return p


def strat_get_y(nh, n_strat, dim, u):
    """y = (digits + u)/Ns, rows cube-sorted, y>=1 -> 0.999999 (vegas_stratification.py:140-165).

    `u` is the [sum(nh), dim] uniform block the reference would draw from its RNG."""
    c = torch.arange(nh.shape[0], dtype=torch.int64)
    pos = strat_cube_digits(nh.shape[0], n_strat, dim)[torch.repeat_interleave(c, nh), :]
    y = (pos.to(u.dtype) + u) / n_strat
    y[y >= 1.0] = 0.999999
    return y


def strat_accumulate(nh, jf):
    """JF[c] = sum jf, JF2[c] = sum jf^2 over the cube's rows (vegas_stratification.py:46-70)."""
    c = torch.arange(nh.shape[0], dtype=torch.int64)
    idx = torch.repeat_interleave(c, nh)
    JF = torch.zeros([nh.shape[0]], dtype=jf.dtype)
    JF2 = torch.zeros([nh.shape[0]], dtype=jf.dtype)
    JF.scatter_add_(0, idx, jf)
    JF2.scatter_add_(0, idx, jf**2.0)
    return JF, JF2


def strat_update_dh(JF, JF2, strat_counts, v_cubes, beta=0.75):
    """EQ 42 damped variances (vegas_stratification.py:72-90)."""
    v2 = v_cubes * v_cubes
    d = v2 * JF2 / strat_counts - (v_cubes * JF / strat_counts) ** 2
    d[d < 0.0] = 0.0
    dh = d**beta
    s = torch.sum(dh)
    if s != 0:
        dh = dh / s
    return dh


def vegas_iteration_estimate(JF, JF2, nh, v_cubes):
    """Per-iteration integral and variance (vegas.py:293-303)."""
    inv = 1.0 / nh.to(JF.dtype)
    ih = JF * (inv * v_cubes)
    sig2 = torch.abs(JF2 * inv * (v_cubes**2) - ih**2)
    return ih.sum(), (sig2 * inv).sum()


# --------------------------------------------------------------------------
# VEGAS driver  (torchquad/integration/vegas.py)
# --------------------------------------------------------------------------
class VegasRun:
    """Restates VEGAS.integrate (vegas.py:30-362) around the functions above.

    `uniform(size, dtype)` supplies the random numbers (the reference's
    `rng.uniform`); injecting recorded numbers gives identical-sample parity.
    """

    def __init__(self, fn, dim, N, domain, uniform, max_iterations=20, eps_rel=0.0,
                 eps_abs=0.0, use_grid_improve=True, use_warmup=True, alpha=0.5, beta=0.75):
        self.fn, self.dim, self.N = fn, dim, N
        self.dtype = domain.dtype
        self.uniform = uniform
        self.max_iterations, self.eps_rel, self.eps_abs = max_iterations, eps_rel, eps_abs
        self.use_grid_improve, self.use_warmup = use_grid_improve, use_warmup
        self.alpha, self.beta = alpha, beta
        self.starts = domain[:, 0]
        self.sizes = domain[:, 1] - self.starts
        self.volume = torch.prod(self.sizes)
        self.starting_N = self.n_increment = N // (max_iterations + 5)  # vegas.py:90-91
        self.n_intervals = max(2, self.n_increment // 10)  # vegas.py:117
        self.x_edges, self.dx_edges, self.weights, self.counts = map_init(
            self.n_intervals, dim, self.dtype)
        self.n_strat, self.n_cubes, self.v_cubes = strat_config(self.n_increment, dim)
        self.dh = strat_init(self.n_cubes, self.dtype)
        self.fevals = 0
        self.results, self.sigma2, self.it = [], [], 0
        self.trace = []  # per-iteration (I, sigma2, M) for tests

    def _eval(self, x):
        self.fevals += x.shape[0]
        return (self.fn(x * self.sizes + self.starts) * self.volume).squeeze()

    def _update_map(self):
        self.x_edges, self.dx_edges, status = map_update(
            self.x_edges, self.dx_edges, self.weights, self.counts, self.alpha)
        self.weights, self.counts = map_reset(self.n_intervals, self.dim, self.dtype)
        return status

    def warmup(self, n_it=5):
        """vegas.py:211-266 (results discarded)."""
        ns = self.starting_N // 5
        for _ in range(n_it):
            y = self.uniform([ns, self.dim], self.dtype) * 0.999999
            x = map_get_x(y, self.x_edges, self.dx_edges)
            f = self._eval(x)
            jac = map_get_jac(y, self.dx_edges)
            jf2 = ((f * jac) ** 2).detach()
            map_accumulate(self.weights, self.counts, y, jf2)
            self._update_map()

    def iteration(self):
        """vegas.py:268-315."""
        nh = strat_get_nh(self.dh, self.starting_N)
        u = self.uniform([int(nh.sum()), self.dim], self.dtype)
        y = strat_get_y(nh, self.n_strat, self.dim, u)
        x = map_get_x(y, self.x_edges, self.dx_edges)
        f = self._eval(x)
        jac = map_get_jac(y, self.dx_edges)
        jf = f * jac
        jf2 = (jf**2).detach()
        if self.use_grid_improve:
            map_accumulate(self.weights, self.counts, y, jf2)
        JF, JF2 = strat_accumulate(nh, jf)
        I, s2 = vegas_iteration_estimate(JF, JF2.detach(), nh, self.v_cubes)
        self.results[-1], self.sigma2[-1] = I, s2.detach()
        if self.use_grid_improve:
            self._update_map()
        self.dh = strat_update_dh(JF.detach(), JF2.detach(), nh.to(self.dtype), self.v_cubes, self.beta)
        self.trace.append((float(I), float(s2), int(nh.sum())))

    def result(self):
        """vegas.py:318-335."""
        if any(s == 0.0 for s in self.sigma2):
            return sum(self.results) / len(self.results)
        num = sum(r / s for r, s in zip(self.results, self.sigma2))
        den = sum(1.0 / s for s in self.sigma2)
        return num / den

    def error(self):
        """vegas.py:337-346."""
        res = sum(1.0 / s for s in self.sigma2 if s != 0.0)
        return self.sigma2[0] if res == 0 else 1.0 / torch.sqrt(res)

    def chisq(self):
        """vegas.py:348-362."""
        Ifin = self.result()
        return sum(((r - Ifin) ** 2 / s for r, s in zip(self.results, self.sigma2) if r != Ifin),
                   start=self.results[0] * 0.0)

    def check_abort(self):
        """vegas.py:161-209."""
        if self.it % 5 > 0:
            return False
        res_abs = torch.abs(self.result())
        err, chi2 = self.error(), self.chisq()
        if (err <= self.eps_rel * res_abs or err <= self.eps_abs) and chi2 / 5.0 < 1.0:
            return True
        if chi2 / 5.0 < 1.0:
            if res_abs == 0.0:
                self.starting_N += self.n_increment
            else:
                acc = err / res_abs
                self.starting_N = min(
                    self.starting_N + self.n_increment,
                    int(self.starting_N * torch.sqrt(acc / (self.eps_rel + 1e-8))),
                )
        elif chi2 / 5.0 > 1.0:
            self.starting_N += self.n_increment
        if self.fevals + self.starting_N * 5 > self.N:
            return True
        if self.it + 5 > self.max_iterations:
            return True
        self.results, self.sigma2 = [], []
        return False

    def run(self):
        if self.use_warmup:
            self.warmup()
        while True:
            self.it += 1
            self.results.append(0)
            self.sigma2.append(0)
            self.iteration()
            if self.check_abort():
                break
        return self.result()


# --------------------------------------------------------------------------
# Monte Carlo  (torchquad/integration/monte_carlo.py)
# --------------------------------------------------------------------------
def mc_sample_points(u, domain):
    """x = u*(b-a) + a (monte_carlo.py:102-106)."""
    starts = domain[:, 0]
    sizes = domain[:, 1] - starts
    return u * sizes + starts


def mc_result(f, domain):
    """I = V*sum(f, axis 0)/N (monte_carlo.py:60-82 incl. the 1-D squeeze decorator, utils.py:235-277)."""
    one_d = f.dim() == 1 or (f.dim() == 2 and f.shape[1] == 1)
    if f.dim() == 1:
        f = f.unsqueeze(1)
    vol = torch.prod(domain[:, 1] - domain[:, 0])
    res = vol * torch.sum(f, dim=0) / f.shape[0]
    return res.squeeze() if one_d else res


def mc_integrate(fn, dim, N, domain, seed=None):
    """MonteCarlo.integrate with the reference's own RNG (monte_carlo.py:20-58, rng.py:119-125):
    torch.manual_seed + torch.rand on the CPU.  Used as the CPU baseline of bench.py."""
    if seed is not None:
        torch.random.manual_seed(seed)
    u = torch.rand(size=[N, dim], dtype=domain.dtype)
    return mc_result(fn(mc_sample_points(u, domain)), domain)


# --------------------------------------------------------------------------
# Newton-Cotes grids  (integration_grid.py, grid_integrator.py, trapezoid/simpson/boole.py)
# --------------------------------------------------------------------------
def nc_adjust_n(rule, dim, N):
    """_adjust_N (simpson.py:54-81, boole.py:56-84; trapezoid: identity)."""
    n = int(N ** (1.0 / dim) + 1e-8)
    if rule == "simpson":
        if n < 3:
            return 3**dim
        if n % 2 != 1:
            return (n - 1) ** dim
    elif rule == "boole":
        if n < 5:
            return 5**dim
        if (n - 1) % 4 != 0:
            return (n - ((n - 1) % 4)) ** dim
    return N


def nc_grid(rule, N, domain):
    """points [n^dim, dim] (dim 0 slowest), h [dim], n (integration_grid.py:64-99)."""
    dim = domain.shape[0]
    N = nc_adjust_n(rule, dim, N)
    n = int(N ** (1.0 / dim) + 1e-8)
    g = [torch.linspace(domain[d][0], domain[d][1], n, dtype=domain.dtype) for d in range(dim)]
    h = torch.stack([g[d][1] - g[d][0] for d in range(dim)])
    mesh = torch.meshgrid(*g, indexing="ij")
    pts = torch.stack([m.ravel() for m in mesh], dim=1)
    return pts, h, n


def nc_result(rule, f, dim, n, hs):
    """Composite rule applied axis by axis (grid_integrator.py:57-91; trapezoid.py:28-37;
    simpson.py:30-46; boole.py:30-48)."""
    one_d = f.dim() == 1 or (f.dim() == 2 and f.shape[1] == 1)
    if f.dim() == 1:
        f = f.unsqueeze(1)
    shape = list(f.shape[1:])
    a = f.movedim(0, -1).reshape(shape + [n] * dim)
    for cur in range(dim):
        if rule == "trapezoid":
            a = hs[cur] / 2.0 * (a[..., 0:-1] + a[..., 1:])
        elif rule == "simpson":
            a = hs[cur] / 3.0 * (a[..., 0:-2][..., ::2] + 4 * a[..., 1:-1][..., ::2] + a[..., 2:][..., ::2])
        elif rule == "boole":
            a = hs[cur] / 22.5 * (
                7 * a[..., 0:-4][..., ::4] + 32 * a[..., 1:-3][..., ::4] + 12 * a[..., 2:-2][..., ::4]
                + 32 * a[..., 3:-1][..., ::4] + 7 * a[..., 4:][..., ::4])
        else:
            raise ValueError(rule)
        a = torch.sum(a, dim=a.dim() - 1)
    return a.squeeze() if one_d else a


# --------------------------------------------------------------------------
# Gauss-Legendre  (torchquad/integration/gaussian.py)
# --------------------------------------------------------------------------
def gauss_grid(N, domain):
    """points [n^dim, dim], product weights [n^dim], n (gaussian.py:44-69,110-129,159-162;
    integration_grid.py:64-99 with the Gauss `_grid_func`)."""
    dim = domain.shape[0]
    n = int(N ** (1.0 / dim) + 1e-8)
    roots, weights = np.polynomial.legendre.leggauss(n)
    roots, weights = torch.tensor(roots), torch.tensor(weights)
    g = [((domain[d][1] - domain[d][0]) / 2) * roots + ((domain[d][0] + domain[d][1]) / 2) for d in range(dim)]
    pts = torch.stack([m.ravel() for m in torch.meshgrid(*g, indexing="ij")], dim=1)
    W = torch.prod(torch.stack(list(torch.meshgrid(*([weights] * dim), indexing="ij")), dim=0), dim=0).ravel()
    return pts, W, n


def gauss_result(f, W, dim, n, domain):
    """evaluate_integrand's `result *= weights` (base_integrator.py:77-89) followed by the Gaussian composite
    rule 0.5*(b-a)*sum per axis (gaussian.py:131-144, grid_integrator.py:70-88)."""
    one_d = f.dim() == 1 or (f.dim() == 2 and f.shape[1] == 1)
    if f.dim() == 1:
        f = f.unsqueeze(1)
    f = f * W.reshape([-1] + [1] * (f.dim() - 1))
    a = f.movedim(0, -1).reshape(list(f.shape[1:]) + [n] * dim)
    for cur in range(dim):
        a = 0.5 * (domain[cur][1] - domain[cur][0]) * torch.sum(a, dim=a.dim() - 1)
    return a.squeeze() if one_d else a


# --------------------------------------------------------------------------
# Built-in integrands: Genz families (Genz 1984/1987; not in the reference)
# and the reference's test integrands (tests/integration_test_functions.py:146-325).
# Each returns (f(x), exact integral over [0,1]^d or the given domain when known).
# --------------------------------------------------------------------------
def genz(family, x, a, u):
    a = torch.as_tensor(a, dtype=x.dtype)
    u = torch.as_tensor(u, dtype=x.dtype)
    if family == "oscillatory":
        return torch.cos(2.0 * math.pi * u[0] + torch.sum(a * x, dim=1))
    if family == "product_peak":
        return torch.prod(1.0 / (a**-2.0 + (x - u) ** 2), dim=1)
    if family == "corner_peak":
        return (1.0 + torch.sum(a * x, dim=1)) ** (-(x.shape[1] + 1.0))
    if family == "gaussian":
        return torch.exp(-torch.sum(a * a * (x - u) ** 2, dim=1))
    if family == "c0":
        return torch.exp(-torch.sum(a * torch.abs(x - u), dim=1))
    if family == "discontinuous":
        inside = (x[:, 0] <= u[0])
        if x.shape[1] > 1:
            inside = inside & (x[:, 1] <= u[1])
        return torch.where(inside, torch.exp(torch.sum(a * x, dim=1)), torch.zeros_like(x[:, 0]))
    raise ValueError(family)


def genz_exact(family, a, u):
    """Closed forms on [0,1]^d in float64."""
    a = np.asarray(a, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    d = a.shape[0]
    if family == "oscillatory":
        return float(np.cos(2 * np.pi * u[0] + a.sum() / 2) * np.prod(2 * np.sin(a / 2) / a))
    if family == "product_peak":
        return float(np.prod(a * (np.arctan(a * (1 - u)) + np.arctan(a * u))))
    if family == "corner_peak":
        tot = 0.0
        for m in range(1 << d):
            v = np.array([(m >> i) & 1 for i in range(d)], dtype=np.float64)
            tot += (-1.0) ** v.sum() / (1.0 + (v * a).sum())
        return float(tot / (math.factorial(d) * np.prod(a)))
    if family == "gaussian":
        erf = np.vectorize(math.erf)
        return float(np.prod(np.sqrt(np.pi) / (2 * a) * (erf(a * (1 - u)) + erf(a * u))))
    if family == "c0":
        return float(np.prod((2 - np.exp(-a * u) - np.exp(-a * (1 - u))) / a))
    if family == "discontinuous":
        r = 1.0
        for i in range(d):
            r *= (np.exp(a[i] * u[i]) - 1) / a[i] if i < 2 else (np.exp(a[i]) - 1) / a[i]
        return float(r)
    raise ValueError(family)


def test_integrand(name, x, coeffs=None):
    """sum sin / sum exp / prod cos / polynomial (tests/integration_test_functions.py:190-325)."""
    if name == "sinusoid":
        return torch.sum(torch.sin(x), dim=1)
    if name == "exponential":
        return torch.sum(torch.exp(x), dim=1)
    if name == "product_cos":
        return torch.prod(torch.cos(x), dim=1)
    if name == "polynomial":
        c = torch.as_tensor(coeffs, dtype=x.dtype)
        k = torch.linspace(0, len(coeffs) - 1, len(coeffs), dtype=x.dtype)
        e = x.reshape(x.shape + (1,)) ** k
        return torch.sum(torch.sum(e * c, dim=2), dim=1)
    raise ValueError(name)
