"""Built-in integrands with device functors (fused path) and an equivalent torch formulation.

Genz test families (A. Genz, "Testing multidimensional integration routines", 1984; definitions and closed
forms on [0,1]^d as in SURVEY 8d) and the reference's test integrands
(/root/reference/tests/integration_test_functions.py:146-325).  Passing one of these objects as `fn` to
MonteCarlo / VEGAS / Trapezoid / Simpson / Boole selects the fused kernels of csrc/fused.cu
(generate -> map -> evaluate -> accumulate without writing samples to HBM).  Calling the object on a tensor
evaluates the same function with torch ops, which is what the unfused path (and autograd) uses.
"""
import math

import torch

from ._lib import TQ_MAX_DIM, tq_integrand

FAMILY = {
    "genz_oscillatory": 0, "genz_product_peak": 1, "genz_corner_peak": 2, "genz_gaussian": 3, "genz_c0": 4,
    "genz_discontinuous": 5, "sum_sin": 6, "sum_exp": 7, "prod_cos": 8, "polynomial": 9,
    "sum_sin_fast": 10, "sum_exp_fast": 11, "prod_cos_fast": 12,
}


def _vec(v, dim, name):
    if isinstance(v, (int, float)):
        v = [float(v)] * dim
    v = [float(t) for t in (v.tolist() if isinstance(v, torch.Tensor) else v)]
    if len(v) != dim:
        raise ValueError(f"{name} must have {dim} entries, got {len(v)}")
    return v


class BuiltinIntegrand:
    """Base class: family id + parameters; subclasses give the torch formulation and the exact integral."""

    family = None
    fast_math = False  # fp32 only: evaluate sin/cos/exp with the SFU intrinsics (abs error ~5e-7)

    def __init__(self, dim, a=None, u=None, coeffs=None):
        if not 1 <= dim <= TQ_MAX_DIM:
            raise ValueError(f"built-in integrands support 1 <= dim <= {TQ_MAX_DIM}")
        self.dim = dim
        self.a = _vec(1.0 if a is None else a, dim, "a")
        self.u = _vec(0.5 if u is None else u, dim, "u")
        self.coeffs = [float(c) for c in (coeffs or [])]
        if len(self.coeffs) > 8:
            raise ValueError("at most 8 polynomial coefficients")

    def _params(self, x):
        """(a, u) as tensors like x; cached per (dtype, device): two host-to-device copies per CALL otherwise."""
        key = (x.dtype, x.device, tuple(self.a), tuple(self.u))
        cached = getattr(self, "_param_cache", None)
        if cached is None or cached[0] != key:
            a = torch.tensor(self.a, dtype=x.dtype, device=x.device)
            u = torch.tensor(self.u, dtype=x.dtype, device=x.device)
            self._param_cache = cached = (key, a, u)
        return cached[1], cached[2]

    def __call__(self, x):
        raise NotImplementedError

    def exact(self):
        """Closed-form integral over [0,1]^dim (float64), None when not implemented."""
        return None

    def to_struct(self, starts, sizes, scale=1.0):
        """Fill the C struct `tq_integrand` for a domain given as host floats."""
        s = tq_integrand()
        s.family = FAMILY[self.family + "_fast"] if self.fast_math and self.family + "_fast" in FAMILY else FAMILY[self.family]
        s.dim = self.dim
        s.ncoeff = len(self.coeffs)
        for i in range(self.dim):
            s.a[i], s.u[i] = self.a[i], self.u[i]
            s.start[i], s.size[i] = float(starts[i]), float(sizes[i])
        for i, c in enumerate(self.coeffs):
            s.coeff[i] = c
        s.scale = float(scale)
        return s


class GenzOscillatory(BuiltinIntegrand):
    """cos(2 pi u_1 + sum a_i x_i)"""

    family = "genz_oscillatory"

    def __call__(self, x):
        a, u = self._params(x)
        return torch.cos(2.0 * math.pi * u[0] + torch.sum(a * x, dim=1))

    def exact(self):
        r = math.cos(2 * math.pi * self.u[0] + sum(self.a) / 2)
        for a in self.a:
            r *= 2 * math.sin(a / 2) / a
        return r


class GenzProductPeak(BuiltinIntegrand):
    """prod 1 / (a_i^-2 + (x_i - u_i)^2)"""

    family = "genz_product_peak"

    def __call__(self, x):
        a, u = self._params(x)
        return torch.prod(1.0 / (a**-2.0 + (x - u) ** 2), dim=1)

    def exact(self):
        r = 1.0
        for a, u in zip(self.a, self.u):
            r *= a * (math.atan(a * (1 - u)) + math.atan(a * u))
        return r


class GenzCornerPeak(BuiltinIntegrand):
    """(1 + sum a_i x_i)^-(d+1)"""

    family = "genz_corner_peak"

    def __call__(self, x):
        a, _ = self._params(x)
        return (1.0 + torch.sum(a * x, dim=1)) ** (-(self.dim + 1.0))

    def exact(self):
        if self.dim > 20:
            return None
        tot = 0.0
        for m in range(1 << self.dim):
            s, bits = 1.0, 0
            for i in range(self.dim):
                if (m >> i) & 1:
                    s += self.a[i]
                    bits += 1
            tot += (-1.0) ** bits / s
        return tot / (math.factorial(self.dim) * math.prod(self.a))


class GenzGaussian(BuiltinIntegrand):
    """exp(-sum a_i^2 (x_i - u_i)^2)"""

    family = "genz_gaussian"

    def __call__(self, x):
        a, u = self._params(x)
        return torch.exp(-torch.sum(a * a * (x - u) ** 2, dim=1))

    def exact(self):
        r = 1.0
        for a, u in zip(self.a, self.u):
            r *= math.sqrt(math.pi) / (2 * a) * (math.erf(a * (1 - u)) + math.erf(a * u))
        return r


class GenzC0(BuiltinIntegrand):
    """exp(-sum a_i |x_i - u_i|)"""

    family = "genz_c0"

    def __call__(self, x):
        a, u = self._params(x)
        return torch.exp(-torch.sum(a * torch.abs(x - u), dim=1))

    def exact(self):
        r = 1.0
        for a, u in zip(self.a, self.u):
            r *= (2 - math.exp(-a * u) - math.exp(-a * (1 - u))) / a
        return r


class GenzDiscontinuous(BuiltinIntegrand):
    """exp(sum a_i x_i) if x_1 <= u_1 and x_2 <= u_2 else 0"""

    family = "genz_discontinuous"

    def __call__(self, x):
        a, u = self._params(x)
        inside = x[:, 0] <= u[0]
        if self.dim > 1:
            inside = inside & (x[:, 1] <= u[1])
        return torch.where(inside, torch.exp(torch.sum(a * x, dim=1)), torch.zeros_like(x[:, 0]))

    def exact(self):
        r = 1.0
        for i, (a, u) in enumerate(zip(self.a, self.u)):
            r *= (math.exp(a * u) - 1) / a if i < 2 else (math.exp(a) - 1) / a
        return r


class SumOfSines(BuiltinIntegrand):
    """sum sin(x_i)  (tests/integration_test_functions.py:286-287, docs tutorial integrand)"""

    family = "sum_sin"

    def __init__(self, dim, fast_math=False):
        super().__init__(dim)
        self.fast_math = bool(fast_math)

    def __call__(self, x):
        return torch.sum(torch.sin(x), dim=1)

    def exact(self):
        return 2.0 * self.dim * math.sin(0.5) ** 2


class SumOfExp(BuiltinIntegrand):
    """sum exp(x_i)  (tests/integration_test_functions.py:247-249)"""

    family = "sum_exp"

    def __init__(self, dim, fast_math=False):
        super().__init__(dim)
        self.fast_math = bool(fast_math)

    def __call__(self, x):
        return torch.sum(torch.exp(x), dim=1)

    def exact(self):
        return self.dim * (math.e - 1.0)


class ProductOfCosines(BuiltinIntegrand):
    """prod cos(x_i)  (tests/integration_test_functions.py:323-325)"""

    family = "prod_cos"

    def __init__(self, dim, fast_math=False):
        super().__init__(dim)
        self.fast_math = bool(fast_math)

    def __call__(self, x):
        return torch.prod(torch.cos(x), dim=1)

    def exact(self):
        return math.sin(1.0) ** self.dim


class Polynomial(BuiltinIntegrand):
    """sum_i sum_k c_k x_i^k  (tests/integration_test_functions.py:146-212)"""

    family = "polynomial"

    def __init__(self, dim, coeffs):
        super().__init__(dim, coeffs=coeffs)
        if not self.coeffs:
            raise ValueError("Polynomial needs at least one coefficient")

    def __call__(self, x):
        h = torch.full_like(x, self.coeffs[-1])
        for c in reversed(self.coeffs[:-1]):
            h = h * x + c
        return torch.sum(h, dim=1)

    def exact(self):
        return self.dim * sum(c / (k + 1) for k, c in enumerate(self.coeffs))
