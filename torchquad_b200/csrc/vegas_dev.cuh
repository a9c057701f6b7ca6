// Device helpers shared by the VEGAS map / stratification kernels and their small-problem cluster versions.
#pragma once
#include "common.cuh"

namespace tq {

template <typename T> struct EdgePair;
template <> struct EdgePair<float> { using type = float2; };
template <> struct EdgePair<double> { using type = double2; };

// Record layout of a LARGE map (tables beyond L2): {x_edge, dx_edge, weight, count} of one bin in one 32-byte (fp64)
// or 16-byte (fp32) record, so that the edge gather and both histogram updates of a sample touch ONE DRAM sector
// instead of three scattered ones.
template <typename T> struct MapRecord;
// The fp64 record keeps its count as an fp64 (exact below 2^53) so that ONE reduction instruction whose lane pairs
// address {w, c} updates both words of the sector (fused.cu, "sector-paired reductions").
template <> struct __align__(32) MapRecord<double> { double x, dx, w, c; };
template <> struct __align__(16) MapRecord<float> { float x, dx, w; unsigned int c; };

// get_NH for one cube: max(2, floor(dh * nevals_exp)) (vegas_stratification.py:92-103)
template <typename T>
__device__ __forceinline__ long long nh_of(T dh, T nev) {
    T v = floor(mul_rn(dh, nev));
    v = v < (T)2 ? (T)2 : v;  // clamp(min=2)
    return (long long)v;
}

// Average with the zero-count fill of vegas_map.py:118-144 in closed form.  After t fill rounds a
// zero-count bin at distance t from the nearest counted bin has taken that bin's average; the right
// neighbour wins ties (its copy happens first in a round); bins farther than 10 keep their raw weight
// (always 0 in practice because weights and counts are accumulated together).
template <typename T>
__device__ __forceinline__ T filled_average(const T* __restrict__ w, const long long* __restrict__ c, long long j,
                                            long long ni) {
    const long long cj = c[j];
    if (cj != 0) return div_rn(w[j], (T)cj);
    int dr = 0, dl = 0;
    for (int t = 1; t <= 10; ++t)
        if (j + t < ni && c[j + t] != 0) { dr = t; break; }
    for (int t = 1; t <= 10; ++t)
        if (j - t >= 0 && c[j - t] != 0) { dl = t; break; }
    if (dr && (!dl || dr <= dl)) return div_rn(w[j + dr], (T)c[j + dr]);
    if (dl) return div_rn(w[j - dl], (T)c[j - dl]);
    return w[j];
}

// Rebinning predicate of update_map (vegas_map.py:214-229): trunc(S_j / delta) > m.  With q = m + 1 this is
// RN(S_j / delta) >= q, and because the rounded quotient is monotone in S_j it equals S_j >= t for the smallest
// double t with RN(t / delta) >= q.  division_threshold finds t with two or three divisions ONCE per new edge,
// so every probe of the search is a load and a compare instead of an fp64 division and a 64-bit conversion
// (those made map_edges_kernel issue-bound: 1190 instructions per edge at Ni = 1e7).  Returns a negative value
// when q * delta is not a normal positive number; callers then keep the division.
__device__ __forceinline__ double division_threshold(double q, double delta) {
    double x = __dmul_rn(q, delta);
    if (!(x >= 2.2250738585072014e-308) || isinf(x)) return -1.0;
    for (int i = 0; i < 64; ++i) {  // walk down while the predecessor still reaches q
        const double p = __longlong_as_double(__double_as_longlong(x) - 1);
        if (p >= 2.2250738585072014e-308 && __ddiv_rn(p, delta) >= q) x = p; else break;
    }
    for (int i = 0; i < 64; ++i) {  // walk up until x reaches q
        if (__ddiv_rn(x, delta) >= q) return x;
        x = __longlong_as_double(__double_as_longlong(x) + 1);
    }
    return -1.0;
}

// Non-finite repair of one new edge (vegas_map.py:240-257): the mean of its two neighbours.
template <typename T>
__device__ __forceinline__ T repaired_edge(const T* __restrict__ xn, long long e, long long ni, bool& was_bad,
                                           bool& still_bad) {
    T v = xn[e];
    was_bad = false;
    still_bad = false;
    if (!isfinite(v)) {
        was_bad = true;
        if (e > 0 && e < ni) v = mul_rn((T)0.5, add_rn(xn[e - 1], xn[e + 1]));
        still_bad = !isfinite(v);
    }
    return v;
}

}  // namespace tq
