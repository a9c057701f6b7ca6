// VEGAS adaptive map: bin lookup + Jacobian, f^2 histogram, smoothing and equal-mass rebinning.
// Replaces torchquad/integration/vegas_map.py (get_X :44-58, get_Jac :60-74, accumulate_weight :99-111,
// _smooth_map :113-172, update_map :185-261).  The reference runs these as Python loops over dim with
// gather/scatter/cumsum ATen launches; here each step is one pass over its data.
#include "common.cuh"
#include "internal.cuh"
#include "vegas_dev.cuh"

namespace tq {

// ------------------------------------------------------------------ forward (get_X + get_Jac + ids)
template <typename T>
__device__ __forceinline__ long long bin_of(T y, T nif, long long ni, T& offset) {
    const T t = mul_rn(y, nif);
    const T fl = floor(t);
    long long k = (long long)fl;
    offset = sub_rn(t, fl);
    // The reference indexes out of range (IndexError) when y*Ni rounds up to Ni; clamp for memory safety.
    k = k < 0 ? 0 : (k >= ni ? ni - 1 : k);
    return k;
}

// Tile kernel.  Phase 1: one thread per 16-byte vector of the row-major tile (coalesced 128-bit loads of y and
// stores of x / ids / offsets); every element looks up its bin, gathers the two edge values (L1/L2 resident
// for ordinary map sizes) and parks its Jacobian factor Ni*dx in shared memory (rows padded to an odd
// stride).  Phase 2: one thread per row multiplies the factors left to right (the reference's order,
// vegas_map.py:70-73) and writes jac coalesced.  HBM traffic = the algorithmic 2*dim*s + s bytes per row.
constexpr int MF_TILE_ELEMS = 2048;

// PACKED: edges come as {x_edge, dx_edge} pairs [dim, Ni] (one gather per element; tq_vegas_map_pack_edges) and
// `domain` (nullable, [dim, 2]) applies the integrator's unit-cube -> domain transform x*size + start of
// vegas.py:109-110 in the same pass (mul then add, like the torch expression it replaces).
template <typename T, int V, bool PACKED>
__global__ void __launch_bounds__(256)
map_forward_kernel(const T* __restrict__ y, const T* __restrict__ xe, const T* __restrict__ dxe,
                   const T* __restrict__ domain, T* __restrict__ x, T* __restrict__ jac, int32_t* __restrict__ ids,
                   T* __restrict__ off, int64_t rows, int dim, long long ni, int tile_rows, uint32_t magic, int ep_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_fac = reinterpret_cast<T*>(smem_raw);  // [tile_rows][dimp]
    const int dimp = dim | 1;
    T* s_dom = s_fac + (size_t)tile_rows * dimp;  // [2][dim] start, size (only with a domain)
    if (domain) {
        for (int d = threadIdx.x; d < dim; d += 256) {
            const T a = domain[2 * d], b = domain[2 * d + 1];
            s_dom[d] = a;
            s_dom[dim + d] = sub_rn(b, a);
        }
        __syncthreads();
    }
    const typename EdgePair<T>::type* ep = reinterpret_cast<const typename EdgePair<T>::type*>(xe);
    const T nif = (T)ni;
    const int64_t ntiles = (rows + tile_rows - 1) / tile_rows;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t r0 = t * tile_rows;
        const int rows_here = (int)(rows - r0 < tile_rows ? rows - r0 : tile_rows);
        const int n_el = rows_here * dim;
        const int64_t e0 = r0 * dim;
        for (int v = threadIdx.x * V; v < n_el; v += 256 * V) {
            alignas(16) T yv[V], xv[V], ov[V];
            alignas(16) int kv[V];
            if (V > 1 && v + V <= n_el) {
                *reinterpret_cast<uint4*>(yv) = __ldcs(reinterpret_cast<const uint4*>(y + e0 + v));
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) yv[j] = v + j < n_el ? y[e0 + v + j] : (T)0;
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const uint32_t e = (uint32_t)(v + j);
                const uint32_t r = (uint32_t)(((uint64_t)e * magic) >> 24);  // e / dim, exact for e < 2^16
                const int d = (int)(e - r * dim);
                T o;
                const long long k = bin_of<T>(yv[j], nif, ni, o);
                if (v + j < n_el) {
                    T xev, dxv;
                    if (PACKED) {
                        // ep_stride 1: pair table; 2: {x, dx, weight, count} records (the pair is a record's first half)
                        const typename EdgePair<T>::type e2 = __ldg(&ep[((int64_t)d * ni + k) * ep_stride]);
                        xev = e2.x;
                        dxv = e2.y;
                    } else {
                        dxv = __ldg(&dxe[(int64_t)d * ni + k]);
                        xev = x ? __ldg(&xe[(int64_t)d * (ni + 1) + k]) : (T)0;
                    }
                    if (x) {
                        T xx = add_rn(xev, mul_rn(dxv, o));
                        if (domain) xx = add_rn(mul_rn(xx, s_dom[dim + d]), s_dom[d]);
                        xv[j] = xx;
                    }
                    s_fac[r * dimp + d] = mul_rn(nif, dxv);
                }
                kv[j] = (int)k;
                ov[j] = o;
            }
            if (V > 1 && v + V <= n_el) {
                if (x) __stcs(reinterpret_cast<uint4*>(x + e0 + v), *reinterpret_cast<uint4*>(xv));
                if (off) __stcs(reinterpret_cast<uint4*>(off + e0 + v), *reinterpret_cast<uint4*>(ov));
                if (ids) {
                    if (V == 4) __stcs(reinterpret_cast<int4*>(ids + e0 + v), *reinterpret_cast<int4*>(kv));
                    else __stcs(reinterpret_cast<int2*>(ids + e0 + v), *reinterpret_cast<int2*>(kv));
                }
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (v + j < n_el) {
                        if (x) x[e0 + v + j] = xv[j];
                        if (off) off[e0 + v + j] = ov[j];
                        if (ids) ids[e0 + v + j] = kv[j];
                    }
                }
            }
        }
        __syncthreads();
        if (jac) {
            for (int r = threadIdx.x; r < rows_here; r += 256) {
                T j = (T)1;
                for (int d = 0; d < dim; ++d) j = mul_rn(j, s_fac[r * dimp + d]);
                jac[r0 + r] = j;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ accumulate (weights += jf2, counts += 1)
// L2 reductions (RED.ADD.F32/F64 + RED.ADD.U64), one thread per 16-byte vector of the row-major y tile
// (coalesced 128-bit loads; (row, dim) of an element from a multiply-shift, no division).
// FUSED: the weight is jf^2 with jf = (f*volume)*jac formed here (vegas.py:104-112,284-287), and jf is
// written out for the per-cube sums -- this replaces three torch elementwise passes of the unfused path.
template <typename T, int V, bool FUSED>
__global__ void __launch_bounds__(256)
map_accumulate_global_kernel(const T* __restrict__ y, const T* __restrict__ jf2_or_f, const T* __restrict__ jacp,
                             T volume, T* __restrict__ jf_out, T* __restrict__ weights,
                             unsigned long long* __restrict__ counts, MapRecord<T>* __restrict__ recs, int64_t rows, int dim,
                             long long ni, int tile_rows, uint32_t magic) {
    const T nif = (T)ni;
    const int64_t ntiles = (rows + tile_rows - 1) / tile_rows;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t r0 = t * tile_rows;
        const int rows_here = (int)(rows - r0 < tile_rows ? rows - r0 : tile_rows);
        const int n_el = rows_here * dim;
        const int64_t e0 = r0 * dim;
        for (int v = threadIdx.x * V; v < n_el; v += 256 * V) {
            alignas(16) T yv[V];
            if (V > 1 && v + V <= n_el) {
                *reinterpret_cast<uint4*>(yv) = __ldcs(reinterpret_cast<const uint4*>(y + e0 + v));
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) yv[j] = v + j < n_el ? y[e0 + v + j] : (T)0;
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (v + j < n_el) {
                    const uint32_t e = (uint32_t)(v + j);
                    const uint32_t r = (uint32_t)(((uint64_t)e * magic) >> 24);
                    const int d = (int)(e - r * dim);
                    T o;
                    const long long k = bin_of<T>(yv[j], nif, ni, o);
                    T wgt;
                    if (FUSED) {
                        const T jf = mul_rn(mul_rn(__ldg(&jf2_or_f[r0 + r]), volume), __ldg(&jacp[r0 + r]));
                        wgt = mul_rn(jf, jf);
                        if (d == 0 && jf_out) jf_out[r0 + r] = jf;
                    } else {
                        wgt = __ldg(&jf2_or_f[r0 + r]);
                    }
                    if (!recs && !weights) continue;  // jf only
                    if (recs) {  // large maps: both reductions in the bin's record (one DRAM sector)
                        MapRecord<T>* rc = &recs[(int64_t)d * ni + k];
                        atomicAdd(&rc->w, wgt);
                        atomicAdd(&rc->c, (decltype(rc->c))1);
                    } else {
                        atomicAdd(&weights[(int64_t)d * ni + k], wgt);
                        atomicAdd(&counts[(int64_t)d * ni + k], 1ull);
                    }
                }
            }
        }
    }
}

// Small maps: per-CTA privatised histogram in shared memory (weights in T, counts u32), flushed once.
template <typename T>
__global__ void __launch_bounds__(512)
map_accumulate_smem_kernel(const T* __restrict__ y, const T* __restrict__ jf2, T* __restrict__ weights,
                           unsigned long long* __restrict__ counts, int64_t rows, int dim, long long ni,
                           int64_t rows_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int bins = dim * (int)ni;
    T* sw = reinterpret_cast<T*>(smem_raw);
    unsigned int* sc = reinterpret_cast<unsigned int*>(sw + bins);
    for (int i = threadIdx.x; i < bins; i += blockDim.x) { sw[i] = (T)0; sc[i] = 0u; }
    __syncthreads();
    const T nif = (T)ni;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    const int64_t e1 = r1 * dim;
    for (int64_t e = r0 * dim + threadIdx.x; e < e1; e += blockDim.x) {
        const int64_t r = e / dim;
        const int d = (int)(e - r * dim);
        T o;
        const int k = (int)bin_of<T>(y[e], nif, ni, o);
        atomicAdd(&sw[d * (int)ni + k], jf2[r]);
        atomicAdd(&sc[d * (int)ni + k], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x) {
        const unsigned int c = sc[i];
        if (c) {
            atomicAdd(&weights[i], sw[i]);
            atomicAdd(&counts[i], (unsigned long long)c);
        }
    }
}

// ------------------------------------------------------------------ smoothing pipeline
constexpr int MAP_TILE = 1024;  // bins per CTA tile (256 threads x 4)

template <typename T>
__global__ void __launch_bounds__(256)
map_average_kernel(const T* __restrict__ weights, const long long* __restrict__ counts, T* __restrict__ avg,
                   double* __restrict__ tile_sums, long long ni, int ntiles, int32_t* status) {
    __shared__ double sh[32];
    const int d = blockIdx.y;
    const T* w = weights + (int64_t)d * ni;
    const long long* c = counts + (int64_t)d * ni;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 4) status[threadIdx.x] = 0;
    double s[1] = {0.0};
    const long long j0 = (long long)blockIdx.x * MAP_TILE + threadIdx.x * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long j = j0 + i;
        if (j < ni) {
            const T a = filled_average<T>(w, c, j, ni);
            avg[(int64_t)d * ni + j] = a;
            s[0] += (double)a;
        }
    }
    block_sum<1>(s, sh);
    if (threadIdx.x == 0) tile_sums[(int64_t)d * ntiles + blockIdx.x] = s[0];
}

// K1r / K2r: one CTA per dimension turns the tile sums into exclusive tile offsets (fp64, fixed order) and
// the row total.
__global__ void __launch_bounds__(256)
map_tile_scan_kernel(double* __restrict__ tile_sums, double* __restrict__ totals, int ntiles) {
    __shared__ double sh[33];
    const int d = blockIdx.x;
    double* t = tile_sums + (int64_t)d * ntiles;
    double carry = 0.0;
    for (int base = 0; base < ntiles; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const double v = i < ntiles ? t[i] : 0.0;
        double total;
        const double ex = block_excl_scan<double>(v, sh, total);
        if (i < ntiles) t[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) totals[d] = carry;
}

// K2: [1,6,1]/8 smoothing with 7/1 borders, normalisation by the row sum, compression ((d-1)/ln d)^alpha
// (vegas_map.py:146-170).  Sets status[0] when any dimension sums to zero (reference returns None).
template <typename T>
__global__ void __launch_bounds__(256)
map_smooth_kernel(const T* __restrict__ avg, const double* __restrict__ row_totals, T* __restrict__ smoothed,
                  double* __restrict__ tile_sums, long long ni, int ntiles, int dim, T alpha, int32_t* status) {
    __shared__ double sh[32];
    __shared__ int any_zero;
    if (threadIdx.x == 0) {
        int z = 0;
        for (int i = 0; i < dim; ++i) z |= ((T)row_totals[i] == (T)0);
        any_zero = z;
        if (z && blockIdx.x == 0 && blockIdx.y == 0) status[0] = 1;
    }
    __syncthreads();
    if (any_zero) return;
    const int d = blockIdx.y;
    const T* a = avg + (int64_t)d * ni;
    const T denom = mul_rn((T)8, (T)row_totals[d]);
    double s[1] = {0.0};
    const long long j0 = (long long)blockIdx.x * MAP_TILE + threadIdx.x * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long j = j0 + i;
        if (j < ni) {
            T v;
            if (j == 0) v = add_rn(mul_rn((T)7, a[0]), a[1]);
            else if (j == ni - 1) v = add_rn(a[ni - 2], mul_rn((T)7, a[ni - 1]));
            else v = add_rn(add_rn(a[j - 1], mul_rn((T)6, a[j])), a[j + 1]);
            v = div_rn(v, denom);
            if (v != (T)0) {
                const T base = div_rn(sub_rn(v, (T)1), log(v));
                v = (alpha == (T)0.5) ? sqrt(base) : pow(base, alpha);  // ATen evaluates x**0.5 as sqrt
            }
            smoothed[(int64_t)d * ni + j] = v;
            s[0] += (double)v;
        }
    }
    block_sum<1>(s, sh);
    if (threadIdx.x == 0) tile_sums[(int64_t)d * ntiles + blockIdx.x] = s[0];
}

// K3: fp64 inclusive prefix sums S[d, j] of the smoothed weights (vegas_map.py:207-213 casts to float64).
template <typename T>
__global__ void __launch_bounds__(256)
map_prefix_kernel(const T* __restrict__ smoothed, const double* __restrict__ tile_offsets, double* __restrict__ S,
                  long long ni, int ntiles, const int32_t* status) {
    __shared__ double sh[33];
    if (status[0]) return;
    const int d = blockIdx.y;
    const T* sm = smoothed + (int64_t)d * ni;
    const long long j0 = (long long)blockIdx.x * MAP_TILE + threadIdx.x * 4;
    double v[4], run = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = (j0 + i < ni) ? (double)sm[j0 + i] : 0.0;
        run += v[i];
    }
    double total;
    double ex = block_excl_scan<double>(run, sh, total) + tile_offsets[(int64_t)d * ntiles + blockIdx.x];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ex += v[i];
        if (j0 + i < ni) S[(int64_t)d * ni + j0 + i] = ex;
    }
}

// K4: new inner edges (vegas_map.py:214-239).  For m = 0..Ni-2:
//   idx = #{j <= Ni-2 : trunc(S_j/delta) <= m}   (the reference builds it as histogram + cumsum)
//   acc = (m+1)*delta - S_{idx-1}                (reference: cumsum of delta - val_per_multiple)
//   x_new[m+1] = xe[idx] + acc/sm[idx]*dxe[idx]
// idx is monotone in m, so the search is windowed: map_edge_bounds_kernel finds idx of the FIRST edge of every block of
// ME_BLOCK edges with a full bisection (one thread per block), and map_edges_kernel searches each edge only inside its
// block's window [bounds[b], bounds[b+1]], staged in shared memory with coalesced loads.  (One full bisection per edge --
// 23 dependent probes at Ni = 1e7 -- made this kernel 8.7 % of a VEGAS run at the reference's map size.)
constexpr int ME_BLOCK = 1024;   // edges per CTA of map_edges_kernel (4 per thread)
constexpr int ME_WINDOW = 3072;  // prefix sums staged per CTA (24 KB); wider windows are searched in global memory

// smallest j in [lo, hi] with trunc(S_j/delta) > m, hi when there is none: bisection on S_j >= thr
// (division_threshold, vegas_dev.cuh; thr < 0: the division itself)
__device__ __forceinline__ long long edge_search(const double* __restrict__ Sd, long long lo, long long hi, double thr, double delta,
                                                 long long m) {
    if (thr >= 0.0) {
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (Sd[mid] >= thr) hi = mid; else lo = mid + 1;
        }
    } else {
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if ((long long)(__ddiv_rn(Sd[mid], delta)) > m) hi = mid; else lo = mid + 1;
        }
    }
    return lo;
}

template <typename T>
__global__ void __launch_bounds__(256)
map_edge_bounds_kernel(const double* __restrict__ S, const double* __restrict__ row_totals, long long* __restrict__ bounds,
                       long long ni, int nblk, const int32_t* status) {
    if (status[0]) return;
    const int d = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const long long m = (long long)b * ME_BLOCK;  // <= Ni - 2 by the choice of nblk
    const double delta = (double)div_rn((T)row_totals[d], (T)ni);
    const double thr = division_threshold((double)(m + 1), delta);
    bounds[(int64_t)d * nblk + b] = edge_search(S + (int64_t)d * ni, 0, ni - 1, thr, delta, m);
}

template <typename T>
__global__ void __launch_bounds__(256)
map_edges_kernel(const T* __restrict__ smoothed, const double* __restrict__ S, const double* __restrict__ row_totals,
                 const long long* __restrict__ bounds, int nblk, const T* __restrict__ xe, const T* __restrict__ dxe,
                 T* __restrict__ x_new, long long ni, const int32_t* status) {
    __shared__ double s_S[ME_WINDOW];
    if (status[0]) return;
    const int d = blockIdx.y;
    const int b = blockIdx.x;
    const T* x_old = xe + (int64_t)d * (ni + 1);
    T* xn = x_new + (int64_t)d * (ni + 1);
    if (b == 0 && threadIdx.x == 0) { xn[0] = x_old[0]; xn[ni] = x_old[ni]; }
    const double* Sd = S + (int64_t)d * ni;
    const double delta = (double)div_rn((T)row_totals[d], (T)ni);
    const long long w_lo = bounds[(int64_t)d * nblk + b];
    const long long w_hi = b + 1 < nblk ? bounds[(int64_t)d * nblk + b + 1] : ni - 1;  // idx of the next block's first edge
    const bool staged = w_hi - w_lo < ME_WINDOW;
    if (staged) {
        for (long long j = w_lo + threadIdx.x; j <= w_hi; j += 256) s_S[j - w_lo] = Sd[j];
        __syncthreads();
    }
    const long long m0 = (long long)b * ME_BLOCK;
#pragma unroll
    for (int i = 0; i < ME_BLOCK / 256; ++i) {
        const long long m = m0 + threadIdx.x + 256 * i;
        if (m > ni - 2) break;
        const double thr = division_threshold((double)(m + 1), delta);
        long long idx;
        double below;
        if (staged) {
            idx = w_lo + edge_search(s_S, 0, w_hi - w_lo, thr, delta, m);
            below = idx > w_lo ? s_S[idx - 1 - w_lo] : (idx > 0 ? Sd[idx - 1] : 0.0);
        } else {
            idx = edge_search(Sd, w_lo, w_hi, thr, delta, m);
            below = idx > 0 ? Sd[idx - 1] : 0.0;
        }
        const T acc = (T)((double)(m + 1) * delta - below);
        const T sm = smoothed[(int64_t)d * ni + idx];
        xn[m + 1] = add_rn(x_old[idx], mul_rn(div_rn(acc, sm), dxe[(int64_t)d * ni + idx]));
    }
}

// K5: non-finite repair (vegas_map.py:240-257, repaired_edge in vegas_dev.cuh), dx = diff(x) (:259) and the
// weight/count reset (:261,:196).
template <typename T>
__global__ void __launch_bounds__(256)
map_finalize_kernel(const T* __restrict__ x_new, T* __restrict__ xe, T* __restrict__ dxe, T* __restrict__ weights,
                    long long* __restrict__ counts, typename EdgePair<T>::type* __restrict__ packed, long long ni,
                    int32_t* status) {
    const int d = blockIdx.y;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < ni) {
        weights[(int64_t)d * ni + e] = (T)0;
        counts[(int64_t)d * ni + e] = 0;
    }
    if (status[0] || e > ni) return;
    const T* xn = x_new + (int64_t)d * (ni + 1);
    bool bad, still;
    const T v = repaired_edge<T>(xn, e, ni, bad, still);
    if (bad) atomicAdd(&status[1], 1);
    if (still) status[2] = 1;
    xe[(int64_t)d * (ni + 1) + e] = v;
    if (e < ni) {
        bool b2, s2;
        const T vn = repaired_edge<T>(xn, e + 1, ni, b2, s2);
        const T dv = sub_rn(vn, v);
        dxe[(int64_t)d * ni + e] = dv;
        if (packed) {
            typename EdgePair<T>::type pr;
            pr.x = v;
            pr.y = dv;
            packed[(int64_t)d * ni + e] = pr;
        }
    }
}

static size_t map_scratch_bytes(int dim, long long ni, size_t elt) {
    const long long ntiles = (ni + MAP_TILE - 1) / MAP_TILE;
    size_t b = 0;
    auto add = [&](size_t n) { b = ((b + 255) & ~(size_t)255) + n; };
    add((size_t)dim * ni * elt);
    add((size_t)dim * ni * elt);
    add((size_t)dim * (ni + 1) * elt);
    add((size_t)dim * ni * sizeof(double));
    add((size_t)dim * ntiles * sizeof(double));
    add((size_t)dim * sizeof(double));
    add((size_t)dim * sizeof(double));
    return b + 256;
}

template <typename T>
static bool carve(Workspace& w, int dim, long long ni, MapScratch& s, bool need_edges) {
    s.ntiles = (int)((ni + MAP_TILE - 1) / MAP_TILE);
    s.avg = w.take<T>((size_t)dim * ni);
    s.smoothed = w.take<T>((size_t)dim * ni);
    s.x_new = need_edges ? (void*)w.take<T>((size_t)dim * (ni + 1)) : nullptr;
    s.S = need_edges ? w.take<double>((size_t)dim * ni) : nullptr;
    s.tile_sums = w.take<double>((size_t)dim * s.ntiles);
    s.totals = w.take<double>(dim);
    s.totals2 = w.take<double>(dim);
    return s.avg && s.smoothed && s.tile_sums && s.totals && s.totals2 && (!need_edges || (s.x_new && s.S));
}

bool map_scratch_carve(void* ws, size_t ws_bytes, int dim, long long ni, int32_t dtype, bool need_edges, MapScratch& s) {
    Workspace w(ws, ws_bytes);
    return dtype == TQ_F64 ? carve<double>(w, dim, ni, s, need_edges) : carve<float>(w, dim, ni, s, need_edges);
}

template <typename T>
static int run_smooth(const T* weights, const long long* counts, T* smoothed_out, MapScratch& s, int dim,
                      long long ni, double alpha, int32_t* status, cudaStream_t st) {
    dim3 grid(s.ntiles, dim);
    map_average_kernel<T><<<TQ_GRID(grid), 256, 0, st>>>(weights, counts, (T*)s.avg, s.tile_sums, ni, s.ntiles, status);
    map_tile_scan_kernel<<<TQ_GRID(dim), 256, 0, st>>>(s.tile_sums, s.totals, s.ntiles);
    map_smooth_kernel<T><<<TQ_GRID(grid), 256, 0, st>>>((const T*)s.avg, s.totals, smoothed_out, s.tile_sums, ni, s.ntiles, dim,
                                               (T)alpha, status);
    map_tile_scan_kernel<<<TQ_GRID(dim), 256, 0, st>>>(s.tile_sums, s.totals2, s.ntiles);
    return check_launch("map smoothing");
}

template <bool PACKED>
static int launch_map_forward(const void* y, const void* xe, const void* dxe, const void* domain, void* x, void* jac,
                              int32_t* ids, void* offset, int64_t rows, int32_t dim, int64_t n_intervals, int32_t dtype,
                              void* stream, int ep_stride = 1) {
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1 && rows >= 0, "tq_vegas_map_forward: bad shape");
    TQ_REQUIRE(ids == nullptr || n_intervals <= 0x7fffffffLL, "tq_vegas_map_forward: ids need Ni < 2^31");
    if (rows == 0) return TQ_OK;
    TQ_REQUIRE(dim <= 255, "tq_vegas_map_forward: dim %d too large (max 255)", dim);
    int tile_rows = MF_TILE_ELEMS / dim;
    if (tile_rows < 4) tile_rows = 4;
    tile_rows &= ~3;  // every tile starts 16-byte aligned
    const uint32_t magic = (uint32_t)((1u << 24) / (uint32_t)dim) + 1u;
    const int64_t ntiles = (rows + tile_rows - 1) / tile_rows;
    const int64_t cap = (int64_t)num_sms() * 8;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    const bool aligned = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(offset)) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(ids) & 7) == 0 && (dtype == TQ_F64 || (reinterpret_cast<uintptr_t>(ids) & 15) == 0);
    TQ_DISPATCH_DTYPE(dtype, {
        const size_t smem = ((size_t)tile_rows * (dim | 1) + 2 * (size_t)dim) * sizeof(T);
        constexpr int VW = 16 / sizeof(T);
        if (aligned)
            map_forward_kernel<T, VW, PACKED><<<TQ_GRID(grid), 256, smem, as_stream(stream)>>>(
                (const T*)y, (const T*)xe, (const T*)dxe, (const T*)domain, (T*)x, (T*)jac, ids, (T*)offset, rows, dim,
                n_intervals, tile_rows, magic, ep_stride);
        else
            map_forward_kernel<T, 1, PACKED><<<TQ_GRID(grid), 256, smem, as_stream(stream)>>>(
                (const T*)y, (const T*)xe, (const T*)dxe, (const T*)domain, (T*)x, (T*)jac, ids, (T*)offset, rows, dim,
                n_intervals, tile_rows, magic, ep_stride);
    });
    return check_launch("map_forward_kernel");
}

template <bool FUSED>
static int launch_accumulate_global(const void* y, const void* a, const void* jac, double volume, void* jf_out,
                                    void* weights, int64_t* counts, int64_t rows, int32_t dim, int64_t n_intervals,
                                    int32_t dtype, cudaStream_t st, void* records = nullptr) {
    TQ_REQUIRE(dim <= 255, "tq_vegas_map_accumulate: dim %d too large (max 255)", dim);
    int tile_rows = MF_TILE_ELEMS / dim;
    if (tile_rows < 4) tile_rows = 4;
    tile_rows &= ~3;
    const uint32_t magic = (uint32_t)((1u << 24) / (uint32_t)dim) + 1u;
    const int64_t ntiles = (rows + tile_rows - 1) / tile_rows;
    const int64_t cap = (int64_t)num_sms() * 8;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    const bool aligned = (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    TQ_DISPATCH_DTYPE(dtype, {
        constexpr int VW = 16 / sizeof(T);
        if (aligned)
            map_accumulate_global_kernel<T, VW, FUSED><<<TQ_GRID(grid), 256, 0, st>>>((const T*)y, (const T*)a, (const T*)jac, (T)volume,
                                                                            (T*)jf_out, (T*)weights, (unsigned long long*)counts,
                                                                            (MapRecord<T>*)records, rows, dim, n_intervals, tile_rows, magic);
        else
            map_accumulate_global_kernel<T, 1, FUSED><<<TQ_GRID(grid), 256, 0, st>>>((const T*)y, (const T*)a, (const T*)jac, (T)volume,
                                                                           (T*)jf_out, (T*)weights, (unsigned long long*)counts,
                                                                           (MapRecord<T>*)records, rows, dim, n_intervals, tile_rows, magic);
    });
    return check_launch("map_accumulate_global_kernel");
}

// update_map; `clear_status` false when the caller has already zeroed status[0..4) (the native loop clears the
// words of all its passes once).
int map_update_launch(void* x_edges, void* dx_edges, void* weights, int64_t* counts, void* edges_packed, int32_t dim,
                      int64_t n_intervals, double alpha, int32_t dtype, int32_t* status, bool clear_status, void* ws,
                      size_t ws_bytes, void* stream) {
    TQ_REQUIRE(dim >= 1 && dim <= 65535 && n_intervals >= 2, "tq_vegas_map_update: need dim >= 1 and Ni >= 2");
    Workspace w(ws, ws_bytes);
    MapScratch s;
    cudaStream_t st = as_stream(stream);
    const long long ni = n_intervals;
    TQ_DISPATCH_DTYPE(dtype, {
        using P2 = typename EdgePair<T>::type;
        if (!carve<T>(w, dim, ni, s, true)) { set_error("tq_vegas_map_update: workspace too small (need %zu bytes)", map_scratch_bytes(dim, ni, sizeof(T))); return TQ_ERR_WORKSPACE; }
        if (small_map_ok(dim, ni)) {
            if (clear_status) cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st);
            SmallMap a = {x_edges, dx_edges, weights, counts, edges_packed, s, dim, ni, alpha, status, true};
            return small_update_launch(nullptr, &a, dtype, stream);
        }
        int rc = run_smooth<T>((const T*)weights, (const long long*)counts, (T*)s.smoothed, s, dim, ni, alpha, status, st);
        if (rc) return rc;
        dim3 grid(s.ntiles, dim);
        map_prefix_kernel<T><<<TQ_GRID(grid), 256, 0, st>>>((const T*)s.smoothed, s.tile_sums, s.S, ni, s.ntiles, status);
        // blocks of ME_BLOCK new edges m = 0 .. Ni-2; their search windows live in tile_sums, free again after the prefix
        // kernel (nblk <= ntiles because ME_BLOCK == MAP_TILE)
        static_assert(ME_BLOCK == MAP_TILE, "the window bounds reuse the tile-sum scratch");
        const int nblk = (int)((ni - 1 + ME_BLOCK - 1) / ME_BLOCK);
        long long* bounds = reinterpret_cast<long long*>(s.tile_sums);
        dim3 grid_b((unsigned)((nblk + 255) / 256), dim);
        map_edge_bounds_kernel<T><<<TQ_GRID(grid_b), 256, 0, st>>>(s.S, s.totals2, bounds, ni, nblk, status);
        dim3 grid_e((unsigned)nblk, dim);
        map_edges_kernel<T><<<TQ_GRID(grid_e), 256, 0, st>>>((const T*)s.smoothed, s.S, s.totals2, bounds, nblk, (const T*)x_edges,
                                                   (const T*)dx_edges, (T*)s.x_new, ni, status);
        dim3 grid_f((unsigned)((ni + 1 + 255) / 256), dim);
        map_finalize_kernel<T><<<TQ_GRID(grid_f), 256, 0, st>>>((const T*)s.x_new, (T*)x_edges, (T*)dx_edges, (T*)weights,
                                                      (long long*)counts, (P2*)edges_packed, ni, status);
    });
    return check_launch("map update");
}


}  // namespace tq

using namespace tq;

extern "C" {

int tq_vegas_map_forward(const void* y, const void* x_edges, const void* dx_edges, void* x, void* jac,
                         int32_t* ids, void* offset, int64_t rows, int32_t dim, int64_t n_intervals,
                         int32_t dtype, void* stream) {
    return launch_map_forward<false>(y, x_edges, dx_edges, nullptr, x, jac, ids, offset, rows, dim, n_intervals, dtype, stream);
}

int tq_vegas_map_forward_packed(const void* y, const void* edges_packed, int32_t edges_layout, const void* domain, void* x,
                                void* jac, int32_t* ids, int64_t rows, int32_t dim, int64_t n_intervals, int32_t dtype,
                                void* stream) {
    TQ_REQUIRE(edges_layout == TQ_EDGES_PAIRS || edges_layout == TQ_EDGES_RECORDS, "tq_vegas_map_forward_packed: unknown edges layout %d", edges_layout);
    return launch_map_forward<true>(y, edges_packed, nullptr, domain, x, jac, ids, nullptr, rows, dim, n_intervals, dtype, stream,
                                    edges_layout == TQ_EDGES_RECORDS ? 2 : 1);
}

int tq_vegas_map_accumulate(const void* y, const void* jf2, void* weights, int64_t* counts, int64_t rows,
                            int32_t dim, int64_t n_intervals, int32_t dtype, void* stream) {
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1 && rows >= 0, "tq_vegas_map_accumulate: bad shape");
    if (rows == 0) return TQ_OK;
    cudaStream_t st = as_stream(stream);
    const size_t elt = dtype == TQ_F64 ? 8 : 4;
    const int64_t bins = (int64_t)dim * n_intervals;
    const size_t smem = (size_t)bins * (elt + 4);
    // Privatise in shared memory only for small maps (<= 4096 bins: same-address contention would serialise
    // the L2 atomics) and when every CTA amortises zero+flush; larger maps stay L2-resident (see fused.cu).
    int64_t ctas = (rows * dim) / (16 * bins);
    if (ctas > (int64_t)num_sms() * 4) ctas = (int64_t)num_sms() * 4;
    if (bins <= 4096 && ctas >= 1) {
        const int64_t rows_per_cta = (rows + ctas - 1) / ctas;
        TQ_DISPATCH_DTYPE(dtype, {
            cudaFuncSetAttribute(map_accumulate_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            map_accumulate_smem_kernel<T><<<TQ_GRID((int)ctas), 512, smem, st>>>((const T*)y, (const T*)jf2, (T*)weights,
                                                                       (unsigned long long*)counts, rows, dim,
                                                                       n_intervals, rows_per_cta);
        });
        return check_launch("map_accumulate_smem_kernel");
    }
    return launch_accumulate_global<false>(y, jf2, nullptr, 1.0, nullptr, weights, counts, rows, dim, n_intervals, dtype, st);
}

int tq_vegas_accumulate_fused(const void* y, const void* f, const void* jac, double volume, void* jf_out, void* weights,
                              int64_t* counts, void* records, int64_t rows, int32_t dim, int64_t n_intervals,
                              int32_t dtype, void* stream) {
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1 && rows >= 0, "tq_vegas_accumulate_fused: bad shape");
    TQ_REQUIRE(!(records != nullptr && (weights != nullptr || counts != nullptr)) && ((weights != nullptr) == (counts != nullptr)),
               "tq_vegas_accumulate_fused: pass weights + counts, or records, or neither (jf only)");
    if (rows == 0) return TQ_OK;
    return launch_accumulate_global<true>(y, f, jac, volume, jf_out, weights, counts, rows, dim, n_intervals, dtype,
                                          as_stream(stream), records);
}

size_t tq_vegas_map_workspace_bytes(int32_t dim, int64_t n_intervals, int32_t dtype) {
    return map_scratch_bytes(dim, n_intervals, dtype == TQ_F64 ? 8 : 4);
}

int tq_vegas_map_smooth(const void* weights, const int64_t* counts, void* smoothed, int32_t dim,
                        int64_t n_intervals, double alpha, int32_t dtype, int32_t* status, void* ws,
                        size_t ws_bytes, void* stream) {
    TQ_REQUIRE(dim >= 1 && dim <= 65535 && n_intervals >= 2, "tq_vegas_map_smooth: need dim >= 1 and Ni >= 2");
    Workspace w(ws, ws_bytes);
    MapScratch s;
    cudaStream_t st = as_stream(stream);
    TQ_DISPATCH_DTYPE(dtype, {
        if (!carve<T>(w, dim, n_intervals, s, false)) { set_error("tq_vegas_map_smooth: workspace too small"); return TQ_ERR_WORKSPACE; }
        if (small_map_ok(dim, n_intervals)) {
            cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st);
            s.smoothed = smoothed;
            SmallMap a = {nullptr, nullptr, const_cast<void*>(weights), const_cast<int64_t*>(counts), nullptr, s, dim, n_intervals,
                          alpha, status, false};
            return small_update_launch(nullptr, &a, dtype, stream);
        }
        return run_smooth<T>((const T*)weights, (const long long*)counts, (T*)smoothed, s, dim, n_intervals, alpha, status, st);
    });
    return TQ_OK;
}

int tq_vegas_map_update(void* x_edges, void* dx_edges, void* weights, int64_t* counts, void* edges_packed,
                        int32_t dim, int64_t n_intervals, double alpha, int32_t dtype, int32_t* status, void* ws,
                        size_t ws_bytes, void* stream) {
    return map_update_launch(x_edges, dx_edges, weights, counts, edges_packed, dim, n_intervals, alpha, dtype, status,
                             true, ws, ws_bytes, stream);
}

}  // extern "C"
