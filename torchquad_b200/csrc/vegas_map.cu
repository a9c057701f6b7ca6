// VEGAS adaptive map: bin lookup + Jacobian, f^2 histogram, smoothing and equal-mass rebinning.
// Replaces torchquad/integration/vegas_map.py (get_X :44-58, get_Jac :60-74, accumulate_weight :99-111,
// _smooth_map :113-172, update_map :185-261).  The reference runs these as Python loops over dim with
// gather/scatter/cumsum ATen launches; here each step is one pass over its data.
#include "common.cuh"

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

template <typename T>
__global__ void __launch_bounds__(256)
map_forward_kernel(const T* __restrict__ y, const T* __restrict__ xe, const T* __restrict__ dxe,
                   T* __restrict__ x, T* __restrict__ jac, int32_t* __restrict__ ids, T* __restrict__ off,
                   int64_t rows, int dim, long long ni) {
    const T nif = (T)ni;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        T j = (T)1;
        for (int d = 0; d < dim; ++d) {
            T o;
            const long long k = bin_of<T>(y[r * dim + d], nif, ni, o);
            const T dxv = __ldg(&dxe[(int64_t)d * ni + k]);
            if (x) {
                const T xv = __ldg(&xe[(int64_t)d * (ni + 1) + k]);
                x[r * dim + d] = add_rn(xv, mul_rn(dxv, o));
            }
            j = mul_rn(j, mul_rn(nif, dxv));
            if (ids) ids[r * dim + d] = (int32_t)k;
            if (off) off[r * dim + d] = o;
        }
        if (jac) jac[r] = j;
    }
}

// ------------------------------------------------------------------ accumulate (weights += jf2, counts += 1)
// Large maps: straight L2 reductions (RED.ADD.F32/F64 + RED.ADD.U64), one thread per (row, dim) element so
// the y reads are coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
map_accumulate_global_kernel(const T* __restrict__ y, const T* __restrict__ jf2, T* __restrict__ weights,
                             unsigned long long* __restrict__ counts, int64_t rows, int dim, long long ni) {
    const T nif = (T)ni;
    const int64_t total = rows * dim;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / dim;
        const int d = (int)(e - r * dim);
        T o;
        const long long k = bin_of<T>(y[e], nif, ni, o);
        atomicAdd(&weights[(int64_t)d * ni + k], jf2[r]);
        atomicAdd(&counts[(int64_t)d * ni + k], 1ull);
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

// K1: average with the zero-count fill of vegas_map.py:118-144 in closed form.  After t fill rounds a
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
template <typename T>
__global__ void __launch_bounds__(256)
map_edges_kernel(const T* __restrict__ smoothed, const double* __restrict__ S, const double* __restrict__ row_totals,
                 const T* __restrict__ xe, const T* __restrict__ dxe, T* __restrict__ x_new, long long ni,
                 const int32_t* status) {
    if (status[0]) return;
    const int d = blockIdx.y;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const T* x_old = xe + (int64_t)d * (ni + 1);
    T* xn = x_new + (int64_t)d * (ni + 1);
    if (m == 0) { xn[0] = x_old[0]; xn[ni] = x_old[ni]; }
    if (m > ni - 2) return;
    const double* Sd = S + (int64_t)d * ni;
    const T delta_t = div_rn((T)row_totals[d], (T)ni);
    const double delta = (double)delta_t;
    // smallest j in [0, Ni-2] with trunc(S_j/delta) > m; Ni-1 when there is none
    long long lo = 0, hi = ni - 1;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const long long k = (long long)(__ddiv_rn(Sd[mid], delta));
        if (k > m) hi = mid; else lo = mid + 1;
    }
    const long long idx = lo;
    const double below = idx > 0 ? Sd[idx - 1] : 0.0;
    const T acc = (T)((double)(m + 1) * delta - below);
    const T sm = smoothed[(int64_t)d * ni + idx];
    xn[m + 1] = add_rn(x_old[idx], mul_rn(div_rn(acc, sm), dxe[(int64_t)d * ni + idx]));
}

// K5: non-finite repair (vegas_map.py:240-257), dx = diff(x) (:259) and the weight/count reset (:261,:196).
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

template <typename T>
__global__ void __launch_bounds__(256)
map_finalize_kernel(const T* __restrict__ x_new, T* __restrict__ xe, T* __restrict__ dxe, T* __restrict__ weights,
                    long long* __restrict__ counts, long long ni, int32_t* status) {
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
        dxe[(int64_t)d * ni + e] = sub_rn(vn, v);
    }
}

struct MapScratch {
    void* avg;       // T[dim*ni]   (reused as x_new: T[dim*(ni+1)] needs its own buffer)
    void* smoothed;  // T[dim*ni]
    void* x_new;     // T[dim*(ni+1)]
    double* S;       // [dim*ni]
    double* tile_sums;  // [dim*ntiles]
    double* totals;     // [dim] (row sums of avg, then of smoothed)
    double* totals2;
    int ntiles;
};

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

template <typename T>
static int run_smooth(const T* weights, const long long* counts, T* smoothed_out, MapScratch& s, int dim,
                      long long ni, double alpha, int32_t* status, cudaStream_t st) {
    dim3 grid(s.ntiles, dim);
    map_average_kernel<T><<<grid, 256, 0, st>>>(weights, counts, (T*)s.avg, s.tile_sums, ni, s.ntiles, status);
    map_tile_scan_kernel<<<dim, 256, 0, st>>>(s.tile_sums, s.totals, s.ntiles);
    map_smooth_kernel<T><<<grid, 256, 0, st>>>((const T*)s.avg, s.totals, smoothed_out, s.tile_sums, ni, s.ntiles, dim,
                                               (T)alpha, status);
    map_tile_scan_kernel<<<dim, 256, 0, st>>>(s.tile_sums, s.totals2, s.ntiles);
    return check_launch("map smoothing");
}

}  // namespace tq

using namespace tq;

extern "C" {

int tq_vegas_map_forward(const void* y, const void* x_edges, const void* dx_edges, void* x, void* jac,
                         int32_t* ids, void* offset, int64_t rows, int32_t dim, int64_t n_intervals,
                         int32_t dtype, void* stream) {
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1 && rows >= 0, "tq_vegas_map_forward: bad shape");
    TQ_REQUIRE(ids == nullptr || n_intervals <= 0x7fffffffLL, "tq_vegas_map_forward: ids need Ni < 2^31");
    if (rows == 0) return TQ_OK;
    const int grid = grid_for(rows, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        map_forward_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)y, (const T*)x_edges, (const T*)dx_edges,
                                                                  (T*)x, (T*)jac, ids, (T*)offset, rows, dim, n_intervals);
    });
    return check_launch("map_forward_kernel");
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
            map_accumulate_smem_kernel<T><<<(int)ctas, 512, smem, st>>>((const T*)y, (const T*)jf2, (T*)weights,
                                                                       (unsigned long long*)counts, rows, dim,
                                                                       n_intervals, rows_per_cta);
        });
        return check_launch("map_accumulate_smem_kernel");
    }
    const int grid = grid_for(rows * dim, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        map_accumulate_global_kernel<T><<<grid, 256, 0, st>>>((const T*)y, (const T*)jf2, (T*)weights,
                                                             (unsigned long long*)counts, rows, dim, n_intervals);
    });
    return check_launch("map_accumulate_global_kernel");
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
    TQ_DISPATCH_DTYPE(dtype, {
        if (!carve<T>(w, dim, n_intervals, s, false)) { set_error("tq_vegas_map_smooth: workspace too small"); return TQ_ERR_WORKSPACE; }
        return run_smooth<T>((const T*)weights, (const long long*)counts, (T*)smoothed, s, dim, n_intervals, alpha, status,
                             as_stream(stream));
    });
    return TQ_OK;
}

int tq_vegas_map_update(void* x_edges, void* dx_edges, void* weights, int64_t* counts, int32_t dim,
                        int64_t n_intervals, double alpha, int32_t dtype, int32_t* status, void* ws,
                        size_t ws_bytes, void* stream) {
    TQ_REQUIRE(dim >= 1 && dim <= 65535 && n_intervals >= 2, "tq_vegas_map_update: need dim >= 1 and Ni >= 2");
    Workspace w(ws, ws_bytes);
    MapScratch s;
    cudaStream_t st = as_stream(stream);
    const long long ni = n_intervals;
    TQ_DISPATCH_DTYPE(dtype, {
        if (!carve<T>(w, dim, ni, s, true)) { set_error("tq_vegas_map_update: workspace too small (need %zu bytes)", map_scratch_bytes(dim, ni, sizeof(T))); return TQ_ERR_WORKSPACE; }
        int rc = run_smooth<T>((const T*)weights, (const long long*)counts, (T*)s.smoothed, s, dim, ni, alpha, status, st);
        if (rc) return rc;
        dim3 grid(s.ntiles, dim);
        map_prefix_kernel<T><<<grid, 256, 0, st>>>((const T*)s.smoothed, s.tile_sums, s.S, ni, s.ntiles, status);
        dim3 grid_e((unsigned)((ni + 255) / 256), dim);
        map_edges_kernel<T><<<grid_e, 256, 0, st>>>((const T*)s.smoothed, s.S, s.totals2, (const T*)x_edges,
                                                   (const T*)dx_edges, (T*)s.x_new, ni, status);
        dim3 grid_f((unsigned)((ni + 1 + 255) / 256), dim);
        map_finalize_kernel<T><<<grid_f, 256, 0, st>>>((const T*)s.x_new, (T*)x_edges, (T*)dx_edges, (T*)weights,
                                                      (long long*)counts, ni, status);
    });
    return check_launch("map update");
}

}  // extern "C"
