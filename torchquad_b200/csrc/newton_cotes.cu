// Newton-Cotes tensor-product grids: point generation from 1-D nodes and the weighted contraction.
// Replaces IntegrationGrid (integration_grid.py:79-99: meshgrid + ravel + stack), the einsum/reshape of
// GridIntegrator.calculate_result (grid_integrator.py:70-82) and the per-axis stencil passes of
// trapezoid.py:28-37 / simpson.py:30-46 / boole.py:30-48, which are one weighted sum
//   sum_p f(p) * prod_d w[d, i_d(p)],   i_d(p) = (p / n^(dim-1-d)) % n   (dim 0 slowest).
#include <stdlib.h>

#include "common.cuh"

namespace tq {

// Multi-index of point p in base n; dim 0 is the most significant digit (integration_grid.py:98-99).
struct GridIndex {
    uint32_t n;
    int dim;
};

// Mixed-radix walker: digit i_d(p) = (p / stride_d) % n of a point index that advances by a constant step.
// Keeps (q = digit, r = p % stride_d); adding the step is a handful of adds/compares, no division.
struct DigitWalker {
    uint64_t stride, r, step_r;
    uint32_t q, step_q, n;
    __device__ __forceinline__ void init(uint64_t p, uint64_t stride_, uint32_t n_, uint64_t step) {
        stride = stride_;
        n = n_;
        const uint64_t t = p / stride_;
        r = p - t * stride_;
        q = (uint32_t)(t % n_);
        const uint64_t ts = step / stride_;
        step_r = step - ts * stride_;
        step_q = (uint32_t)(ts % n_);
    }
    __device__ __forceinline__ void advance() {
        r += step_r;
        uint32_t carry = 0;
        if (r >= stride) { r -= stride; carry = 1; }
        q += step_q + carry;
        if (q >= n) q -= n;
        if (q >= n) q -= n;
    }
};

// Thread t owns the vector of V consecutive dimensions `t % nvb` of rows t / nvb + k*rows_per_pass: the
// columns (hence the digit walkers and the store width) are loop-invariant, a warp writes one contiguous
// span of the row-major output, and no division is executed inside the loop.
template <typename T, int V>
__global__ void __launch_bounds__(256)
grid_points_kernel(const T* __restrict__ nodes, uint32_t n, int dim, int nvb, int64_t p_begin, int64_t p_end,
                   T* __restrict__ out, bool vec_ok) {
    const int rows_per_pass = 256 / nvb;
    const int rloc = threadIdx.x / nvb;
    const int vb = threadIdx.x - rloc * nvb;
    if (rloc >= rows_per_pass) return;
    const int d0 = vb * V;
    const int nvalid = dim - d0 < V ? dim - d0 : V;
    const int64_t nrows = p_end - p_begin;
    const int64_t step = (int64_t)gridDim.x * rows_per_pass;
    int64_t row = (int64_t)blockIdx.x * rows_per_pass + rloc;
    if (row >= nrows) return;
    DigitWalker w[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        uint64_t stride = 1;
        const int d = d0 + j < dim ? d0 + j : dim - 1;
        for (int i = dim - 1; i > d; --i) stride *= n;
        w[j].init((uint64_t)(p_begin + row), stride, n, (uint64_t)step);
    }
    T* p = out + row * dim + d0;
    const int64_t pstep = step * dim;
    const bool vec = vec_ok && nvalid == V;
    for (; row < nrows; row += step, p += pstep) {
        alignas(16) T v[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int d = d0 + j < dim ? d0 + j : dim - 1;
            v[j] = __ldg(&nodes[(int64_t)d * n + w[j].q]);
            w[j].advance();
        }
        if (vec) {
            __stcs(reinterpret_cast<uint4*>(p), *reinterpret_cast<uint4*>(v));
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (j < nvalid) p[j] = v[j];
        }
    }
}

// grad_nodes[d, j] += grad_points[p, d] for i_d(p) = j: shared-memory privatised [dim, n] table per CTA.
template <typename T>
__global__ void __launch_bounds__(256)
grid_points_backward_kernel(const T* __restrict__ g, uint32_t n, int dim, int64_t p_begin, int64_t p_end,
                            double* __restrict__ grad_nodes, bool use_smem) {
    extern __shared__ double s_acc[];  // [dim*n]
    __shared__ uint64_t s_stride[TQ_MAX_DIM];
    if (use_smem)
        for (int i = threadIdx.x; i < dim * (int)n; i += blockDim.x) s_acc[i] = 0.0;
    if (threadIdx.x == 0) {
        uint64_t s = 1;
        for (int d = dim - 1; d >= 0; --d) { s_stride[d] = s; s *= n; }
    }
    __syncthreads();
    const int64_t total = (p_end - p_begin) * dim;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / dim;
        const int d = (int)(e - r * dim);
        const uint64_t p = (uint64_t)(p_begin + r);
        const uint32_t i = (uint32_t)((p / s_stride[d]) % n);
        if (use_smem) atomicAdd(&s_acc[d * n + i], (double)g[e]);
        else atomicAdd(&grad_nodes[(int64_t)d * n + i], (double)g[e]);
    }
    if (!use_smem) return;
    __syncthreads();
    for (int i = threadIdx.x; i < dim * (int)n; i += blockDim.x)
        if (s_acc[i] != 0.0) atomicAdd(&grad_nodes[i], s_acc[i]);
}

// The [dim, n] table of 1-D weights is staged in shared memory when it fits (always, except 1-D grids
// with very many nodes), otherwise read through the read-only path.
constexpr size_t NC_TABLE_SMEM = 96 * 1024;

template <typename T>
__device__ __forceinline__ const T* stage_weights(const T* __restrict__ w, T* smem, int dim, uint32_t n, bool use_smem) {
    if (!use_smem) return w;
    for (int i = threadIdx.x; i < dim * (int)n; i += blockDim.x) smem[i] = w[i];
    __syncthreads();
    return smem;
}

// prod_d w[d, i_d(p)], digits peeled from the least significant (last) dimension; no digit array.
template <typename T>
__device__ __forceinline__ T point_weight(const T* __restrict__ sw, const GridIndex& gi, uint64_t p) {
    T w = (T)1;
    if (p <= 0xffffffffull) {
        uint32_t q = (uint32_t)p;
        for (int d = gi.dim - 1; d >= 0; --d) {
            const uint32_t t = q / gi.n;
            w *= sw[d * gi.n + (q - t * gi.n)];
            q = t;
        }
    } else {
        for (int d = gi.dim - 1; d >= 0; --d) {
            const uint64_t t = p / gi.n;
            w *= sw[d * gi.n + (uint32_t)(p - t * gi.n)];
            p = t;
        }
    }
    return w;
}

template <typename T>
__global__ void __launch_bounds__(256)
point_weights_kernel(const T* __restrict__ w, uint32_t n, int dim, int64_t p_begin, int64_t p_end, T* __restrict__ out,
                     bool use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const T* sw = stage_weights<T>(w, reinterpret_cast<T*>(smem_raw), dim, n, use_smem);
    const GridIndex gi{n, dim};
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < p_end - p_begin;
         r += (int64_t)gridDim.x * blockDim.x)
        out[r] = point_weight<T>(sw, gi, (uint64_t)(p_begin + r));
}

// 128-bit register <-> element helpers (keep the prefetched vector in registers, not in a local array)
template <typename T> __device__ __forceinline__ T vec_elem(const uint4& q, int j);
template <> __device__ __forceinline__ float vec_elem<float>(const uint4& q, int j) {
    return __uint_as_float(j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : q.w);
}
template <> __device__ __forceinline__ double vec_elem<double>(const uint4& q, int j) {
    return j == 0 ? __hiloint2double((int)q.y, (int)q.x) : __hiloint2double((int)q.w, (int)q.z);
}
template <typename T> __device__ __forceinline__ uint4 pack_vec(const T* v, int count);
template <> __device__ __forceinline__ uint4 pack_vec<float>(const float* v, int count) {
    return make_uint4(__float_as_uint(v[0]), count > 1 ? __float_as_uint(v[1]) : 0u, count > 2 ? __float_as_uint(v[2]) : 0u,
                      count > 3 ? __float_as_uint(v[3]) : 0u);
}
template <> __device__ __forceinline__ uint4 pack_vec<double>(const double* v, int count) {
    const double b = count > 1 ? v[1] : 0.0;
    return make_uint4((unsigned)__double2loint(v[0]), (unsigned)__double2hiint(v[0]), (unsigned)__double2loint(b),
                      (unsigned)__double2hiint(b));
}

// Weight of the leading dim-1 digits of row h = p / n (fastest leading digit first).  Out of line: the hot loop
// of contract1_kernel only needs it when a carry ripples past the fastest digit.
template <typename T>
__device__ __noinline__ T row_prefix(const T* sw, uint32_t n, int dim, FastDiv fd, uint64_t h) {
    T pr = (T)1;
    if (h <= 0xffffffffull) {
        uint32_t q = (uint32_t)h;
        for (int d = dim - 2; d >= 0; --d) {
            const uint32_t t = fd.div(q);
            pr *= sw[d * n + (q - t * n)];
            q = t;
        }
    } else {
        for (int d = dim - 2; d >= 0; --d) {
            const uint64_t t = h / n;
            pr *= sw[d * n + (uint32_t)(h - t * n)];
            h = t;
        }
    }
    return pr;
}

// cols == 1 contraction: fp64 accumulation of f[p]*W[p], deterministic two-stage reduction.
// The grid is viewed as (dim - m) leading dimensions of n nodes and ONE trailing super-dimension of R = n^m <= C1_MAXR
// nodes whose weights (products of the last m dimensions' weights) are tabulated once per CTA in shared memory.  The weight
// of point p is then prefix(p / R) * wR[p % R]: a CTA owns a contiguous range of points, tabulates the prefixes of the
// (few) super-rows it covers -- one digit walk each -- and streams its range with 128-bit loads, one multiply-shift
// division, two shared-memory reads and an fp64 FMA per point; no barrier inside the stream.
// (Was: one digit walk per 32-byte vector, 207 thread instructions per vector, 3.7 TB/s; profiles/r2 for this version.)
constexpr int C1_MAXR = 2048;      // entries of the super-dimension table
constexpr int C1_MAXROWS = 1024;   // super-rows tabulated per pass over a CTA's range

struct C1Plan {
    FastDiv fdn;     // division by n
    FastDiv fdR;     // division by R
    uint32_t R;      // n^m
    int m;           // trailing dimensions folded into the super-dimension
    int64_t chunk;   // points per CTA (multiple of the vector width)
};

template <typename T, int V>
__global__ void __launch_bounds__(256, 4)
contract1_kernel(const T* __restrict__ f, const T* __restrict__ w, uint32_t n, int dim, int64_t p_begin,
                 int64_t p_end, C1Plan plan, double* partials, unsigned int* ticket, double* out) {
    __shared__ double sh[32];
    // V copies of the table, copy c shifted by c entries: the V weights of a 128-bit vector of f starting at ANY offset
    // are one aligned 128-bit shared-memory read (consecutive lanes -> consecutive 16-byte words, no bank conflicts; the
    // scalar reads of a stride-V pattern were 4-way conflicted and capped the kernel at 3.9 TB/s)
    constexpr int RP = C1_MAXR + 4;  // padded copy length (a multiple of every V)
    __shared__ __align__(16) T s_wR[V * RP];
    __shared__ T s_prefix[C1_MAXROWS];
    const uint32_t R = plan.R;
    const int lead = dim - plan.m;  // leading dimensions (digits of the super-row index)
    // super-dimension table: wR[j] = prod over the last m dimensions, slowest of them first
    const bool table = R <= (uint32_t)C1_MAXR;
    const T* wR = table ? s_wR : w + (size_t)(dim - 1) * n;  // n > C1_MAXR: m == 1, the weights themselves (global memory)
    for (uint32_t j = threadIdx.x; j < R && table; j += blockDim.x) {
        uint32_t q = j;
        T pr = (T)1;
        for (int d = dim - 1; d >= lead; --d) {
            const uint32_t t = plan.fdn.div(q);
            pr *= w[d * n + (q - t * n)];
            q = t;
        }
#pragma unroll
        for (int c = 0; c < V; ++c)
            if (j >= (uint32_t)c) s_wR[c * RP + (j - c)] = pr;  // copy c holds wR[i + c] at position i
    }
    const int64_t npts = p_end - p_begin;
    double acc[1] = {0.0};
    // chunks are dealt round-robin: at any time the CTAs of the grid stream one contiguous window of f
    for (int64_t c0 = (int64_t)blockIdx.x * plan.chunk; c0 < npts; c0 += (int64_t)gridDim.x * plan.chunk) {
    const int64_t c1 = c0 + plan.chunk < npts ? c0 + plan.chunk : npts;
    int64_t r0 = c0;  // start of the current pass, relative to p_begin (f is indexed from there)
    while (r0 < c1) {
        const uint64_t p0 = (uint64_t)(p_begin + r0);
        const uint64_t h0 = p0 / R;                          // first super-row of the pass
        const uint32_t l0 = (uint32_t)(p0 - h0 * R);         // offset of the pass's first point inside it
        // as many points as C1_MAXROWS super-rows hold (and 32-bit local offsets allow)
        int64_t span = (int64_t)C1_MAXROWS * R - l0;
        if (span > (int64_t)0x7fffffff - R) span = (int64_t)0x7fffffff - R;
        span -= span % V;                                    // keeps the next pass vector-aligned
        const int64_t r1 = r0 + span < c1 ? r0 + span : c1;
        const uint32_t here = (uint32_t)(r1 - r0);
        const int nrows = (int)plan.fdR.div(l0 + here - 1) + 1;
        __syncthreads();                                     // table written / previous pass's prefixes no longer read
        for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
            // prefix of super-row h: product over the leading digits, fastest of them first (same order as row_prefix)
            uint64_t h = h0 + r;
            T pr = (T)1;
            for (int d = lead - 1; d >= 0; --d) {
                uint32_t digit;
                if (h <= 0xffffffffull) {
                    const uint32_t t = plan.fdn.div((uint32_t)h);
                    digit = (uint32_t)h - t * n;
                    h = t;
                } else {
                    const uint64_t t = h / n;
                    digit = (uint32_t)(h - t * n);
                    h = t;
                }
                pr *= w[d * n + digit];
            }
            s_prefix[r] = pr;
        }
        __syncthreads();
        constexpr int UNROLL = 4;
        for (uint32_t e0 = threadIdx.x * V; e0 < here; e0 += 256 * V * UNROLL) {
            T fv[UNROLL][V];
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) {
                const uint32_t e = e0 + k * 256 * V;
                if (V > 1 && e + V <= here) {
                    const uint4 q = __ldcs(reinterpret_cast<const uint4*>(f + r0 + e));
#pragma unroll
                    for (int j = 0; j < V; ++j) fv[k][j] = vec_elem<T>(q, j);
                } else {
#pragma unroll
                    for (int j = 0; j < V; ++j) fv[k][j] = e + j < here ? f[r0 + e + j] : (T)0;
                }
            }
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) {
                const uint32_t e = e0 + k * 256 * V;
                if (e < here) {
                    uint32_t row = plan.fdR.div(l0 + e);
                    uint32_t last = l0 + e - row * R;
                    if (V > 1 && table && last + V <= R && e + V <= here) {  // the whole vector inside one super-row
                        const uint32_t c = last % V;
                        const uint4 q = *reinterpret_cast<const uint4*>(s_wR + c * RP + (last - c));
                        const T pr = s_prefix[row];
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[0] += (double)fv[k][j] * (double)(pr * vec_elem<T>(q, j));
                    } else {
#pragma unroll
                        for (int j = 0; j < V; ++j) {
                            if (e + j < here) acc[0] += (double)fv[k][j] * (double)(s_prefix[row] * wR[last]);
                            if (++last >= R) { last = 0; ++row; }
                        }
                    }
                }
            }
        }
        r0 = r1;
    }
    }
    grid_sum_finish<1>(acc, sh, partials, ticket, out);
}

// cols > 1 (vector-valued integrands, grid_integrator.py:70-82): a CTA walks tiles of CK_ROWS grid points.  Per tile
// the point weights are formed ONCE per row (one thread each) into shared memory, then the tile's rows x cols values
// are read coalesced by threads that keep a FIXED column (stride a multiple of cols), so a thread accumulates in
// registers and touches the per-CTA column accumulators once, at the end.  Up to CK_COLS_PER_THREAD * 256 columns.
constexpr int CK_ROWS = 1024;
constexpr int CK_COLS_PER_THREAD = 8;

template <typename T>
__global__ void __launch_bounds__(256)
contractk_kernel(const T* __restrict__ f, const T* __restrict__ w, uint32_t n, int dim, int64_t p_begin,
                 int64_t p_end, int cols, double* partials, unsigned int* ticket, double* out, bool use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ bool is_last;
    __shared__ T s_pw[CK_ROWS];
    double* sacc = reinterpret_cast<double*>(smem_raw);  // [cols]
    for (int c = threadIdx.x; c < cols; c += 256) sacc[c] = 0.0;
    const T* sw = stage_weights<T>(w, reinterpret_cast<T*>(sacc + cols), dim, n, use_smem);  // [dim*n]; syncs when staged
    const GridIndex gi{n, dim};
    const int64_t rows = p_end - p_begin;
    const int64_t ntiles = (rows + CK_ROWS - 1) / CK_ROWS;
    double acc[CK_COLS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < CK_COLS_PER_THREAD; ++k) acc[k] = 0.0;
    // narrow: threads [0, span) own column t % cols and rows t / cols, t / cols + rstep, ...
    const int span = cols <= 256 ? (256 / cols) * cols : 0;
    const int rstep = cols <= 256 ? 256 / cols : 1;
    const int my_col = cols <= 256 ? (int)threadIdx.x % cols : 0;
    const int my_row0 = cols <= 256 ? (int)threadIdx.x / cols : 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t r0 = t * CK_ROWS;
        const int rows_here = (int)(rows - r0 < CK_ROWS ? rows - r0 : CK_ROWS);
        __syncthreads();  // previous tile's weights are no longer read
        for (int r = threadIdx.x; r < rows_here; r += 256) s_pw[r] = point_weight<T>(sw, gi, (uint64_t)(p_begin + r0 + r));
        __syncthreads();
        const T* ft = f + r0 * cols;
        if (cols <= 256) {
            if ((int)threadIdx.x < span) {
                int r = my_row0;
                for (; r + 3 * rstep < rows_here; r += 4 * rstep) {  // four independent loads in flight
                    const T v0 = __ldcs(ft + (int64_t)r * cols + my_col);
                    const T v1 = __ldcs(ft + (int64_t)(r + rstep) * cols + my_col);
                    const T v2 = __ldcs(ft + (int64_t)(r + 2 * rstep) * cols + my_col);
                    const T v3 = __ldcs(ft + (int64_t)(r + 3 * rstep) * cols + my_col);
                    acc[0] += (double)v0 * (double)s_pw[r];
                    acc[0] += (double)v1 * (double)s_pw[r + rstep];
                    acc[0] += (double)v2 * (double)s_pw[r + 2 * rstep];
                    acc[0] += (double)v3 * (double)s_pw[r + 3 * rstep];
                }
                for (; r < rows_here; r += rstep) acc[0] += (double)__ldcs(ft + (int64_t)r * cols + my_col) * (double)s_pw[r];
            }
        } else {  // wide: thread owns columns threadIdx.x + 256 k
            for (int r = 0; r < rows_here; ++r) {
                const double pw = (double)s_pw[r];
                const T* fr = ft + (int64_t)r * cols;
#pragma unroll
                for (int k = 0; k < CK_COLS_PER_THREAD; ++k) {
                    const int c = (int)threadIdx.x + 256 * k;
                    if (c < cols) acc[k] += (double)__ldcs(fr + c) * pw;
                }
            }
        }
    }
    __syncthreads();
    if (cols <= 256) {  // fixed-order sum of the rstep threads that share a column (deterministic)
        __shared__ double s_thr[256];
        s_thr[threadIdx.x] = acc[0];
        __syncthreads();
        if ((int)threadIdx.x < cols) {
            double a = 0.0;
            for (int j = 0; j < rstep; ++j) a += s_thr[threadIdx.x + j * cols];
            sacc[threadIdx.x] = a;
        }
    } else {
#pragma unroll
        for (int k = 0; k < CK_COLS_PER_THREAD; ++k) {
            const int c = (int)threadIdx.x + 256 * k;
            if (c < cols) sacc[c] = acc[k];  // one owner per column
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += 256) partials[(size_t)blockIdx.x * cols + c] = sacc[c];
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int c = threadIdx.x; c < cols; c += 256) {
            double a = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) a += __ldcg(&partials[(size_t)b * cols + c]);
            out[c] = a;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

}  // namespace tq

using namespace tq;

static int check_grid_args(const char* who, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end) {
    TQ_REQUIRE(n >= 1 && dim >= 1 && dim <= TQ_MAX_DIM, "%s: need n >= 1 and 1 <= dim <= %d", who, TQ_MAX_DIM);
    TQ_REQUIRE(p_begin >= 0 && p_end >= p_begin, "%s: bad point range", who);
    return TQ_OK;
}

extern "C" {

int tq_nc_grid_points(const void* nodes, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                      void* points, int32_t dtype, void* stream) {
    int rc = check_grid_args("tq_nc_grid_points", n, dim, p_begin, p_end);
    if (rc) return rc;
    if (p_end == p_begin) return TQ_OK;
    TQ_DISPATCH_DTYPE(dtype, {
        constexpr int V = 16 / sizeof(T);
        const int nvb = (dim + V - 1) / V;
        const int rows_per_pass = 256 / nvb;
        const int64_t passes = (p_end - p_begin + rows_per_pass - 1) / rows_per_pass;
        const int64_t cap = (int64_t)num_sms() * 8;
        const int grid = (int)(passes < cap ? passes : cap);
        const bool vec_ok = (dim % V) == 0 && (reinterpret_cast<uintptr_t>(points) & 15) == 0;
        grid_points_kernel<T, V><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((const T*)nodes, (uint32_t)n, dim, nvb, p_begin, p_end,
                                                                     (T*)points, vec_ok);
    });
    return check_launch("grid_points_kernel");
}

int tq_nc_grid_points_backward(const void* grad_points, int32_t n, int32_t dim, int64_t p_begin,
                               int64_t p_end, double* grad_nodes_f64, int32_t dtype, void* stream) {
    int rc = check_grid_args("tq_nc_grid_points_backward", n, dim, p_begin, p_end);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(grad_nodes_f64, 0, (size_t)dim * n * sizeof(double), st);
    if (p_end == p_begin) return TQ_OK;
    const int grid = grid_for((p_end - p_begin) * dim, 256, 2);
    const bool use_smem = (size_t)dim * n * sizeof(double) <= NC_TABLE_SMEM;
    const size_t smem = use_smem ? (size_t)dim * n * sizeof(double) : 0;
    TQ_DISPATCH_DTYPE(dtype, {
        cudaFuncSetAttribute(grid_points_backward_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NC_TABLE_SMEM);
        grid_points_backward_kernel<T><<<TQ_GRID(grid), 256, smem, st>>>((const T*)grad_points, (uint32_t)n, dim, p_begin, p_end, grad_nodes_f64, use_smem);
    });
    return check_launch("grid_points_backward_kernel");
}

int tq_nc_point_weights(const void* w, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                        void* out, int32_t dtype, void* stream) {
    int rc = check_grid_args("tq_nc_point_weights", n, dim, p_begin, p_end);
    if (rc) return rc;
    if (p_end == p_begin) return TQ_OK;
    const int grid = grid_for(p_end - p_begin, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        const bool use_smem = (size_t)dim * n * sizeof(T) <= NC_TABLE_SMEM;
        const size_t smem = use_smem ? (size_t)dim * n * sizeof(T) : 0;
        cudaFuncSetAttribute(point_weights_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NC_TABLE_SMEM);
        point_weights_kernel<T><<<TQ_GRID(grid), 256, smem, as_stream(stream)>>>((const T*)w, (uint32_t)n, dim, p_begin, p_end, (T*)out, use_smem);
    });
    return check_launch("point_weights_kernel");
}

int tq_nc_contract(const void* f, const void* w, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                   int64_t cols, int32_t dtype, double* out_f64, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_grid_args("tq_nc_contract", n, dim, p_begin, p_end);
    if (rc) return rc;
    TQ_REQUIRE(cols >= 1 && cols <= 2048, "tq_nc_contract: 1 <= cols <= 2048 (got %lld)", (long long)cols);
    Workspace wk(ws, ws_bytes);
    unsigned int* ticket = wk.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    cudaStream_t st = as_stream(stream);
    const int64_t rows = p_end - p_begin;
    if (cols == 1) {
        const bool aligned = (reinterpret_cast<uintptr_t>(f) & 15) == 0;
        TQ_DISPATCH_DTYPE(dtype, {
            constexpr int V = 16 / sizeof(T);
            C1Plan plan;
            plan.fdn.set((uint32_t)n);
            plan.R = 1;
            plan.m = 0;
            while (plan.m < dim && (uint64_t)plan.R * (uint64_t)n <= (uint64_t)C1_MAXR) { plan.R *= (uint32_t)n; ++plan.m; }
            if (plan.m == 0) {  // n > C1_MAXR (few, very long dimensions): the last dimension's weights are read from global memory
                plan.R = (uint32_t)n;
                plan.m = 1;
            }
            {
                plan.fdR.set(plan.R);
                // chunks of C1_CHUNK points (a multiple of the vector width), dealt round-robin to a persistent grid
                static const int64_t chunk_env = getenv("TQ_C1_CHUNK") ? atoll(getenv("TQ_C1_CHUNK")) : 0;
                int64_t ctas = (int64_t)num_sms() * 4;
                int64_t chunk = (rows + ctas - 1) / ctas;  // one contiguous range per CTA (measured faster than dealing
                if (chunk < 8192) chunk = 8192;            // 16K..256K-point chunks round-robin: 3.9 vs 3.6 TB/s)
                if (chunk_env > 0) chunk = chunk_env;
                chunk = (chunk + 255) / 256 * 256;
                ctas = (rows + chunk - 1) / chunk;
                if (ctas > (int64_t)num_sms() * 4) ctas = (int64_t)num_sms() * 4;
                plan.chunk = chunk;
                double* partials = wk.take<double>((size_t)ctas);
                if (!ticket || !partials) { set_error("tq_nc_contract: workspace too small"); return TQ_ERR_WORKSPACE; }
                if (aligned)
                    contract1_kernel<T, V><<<TQ_GRID((unsigned)ctas), 256, 0, st>>>((const T*)f, (const T*)w, (uint32_t)n, dim, p_begin, p_end,
                                                                                   plan, partials, ticket, out_f64);
                else
                    contract1_kernel<T, 1><<<TQ_GRID((unsigned)ctas), 256, 0, st>>>((const T*)f, (const T*)w, (uint32_t)n, dim, p_begin, p_end,
                                                                                   plan, partials, ticket, out_f64);
                return check_launch("contract1_kernel");
            }
        });
    }
    const int64_t ntiles = (rows + CK_ROWS - 1) / CK_ROWS;
    int64_t grid = ntiles < (int64_t)num_sms() * 4 ? ntiles : (int64_t)num_sms() * 4;
    if (grid < 1) grid = 1;
    // the per-CTA partials (cols doubles each) must fit the caller's workspace
    const size_t ws_left = ws_bytes > (size_t)(64 << 10) ? ws_bytes - (size_t)(64 << 10) : 0;
    const int64_t grid_cap = (int64_t)(ws_left / ((size_t)cols * sizeof(double)));
    if (grid > grid_cap) grid = grid_cap > 0 ? grid_cap : 1;
    double* partials = wk.take<double>((size_t)grid * cols);
    if (!ticket || !partials) { set_error("tq_nc_contract: workspace too small"); return TQ_ERR_WORKSPACE; }
    TQ_DISPATCH_DTYPE(dtype, {
        const bool use_smem = (size_t)dim * n * sizeof(T) <= NC_TABLE_SMEM;
        const size_t smem = (size_t)cols * sizeof(double) + (use_smem ? (size_t)dim * n * sizeof(T) : 0);
        if (smem > 40 * 1024)
            cudaFuncSetAttribute(contractk_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NC_TABLE_SMEM + 2048 * sizeof(double)));
        contractk_kernel<T><<<TQ_GRID((int)grid), 256, smem, st>>>((const T*)f, (const T*)w, (uint32_t)n, dim, p_begin, p_end, (int)cols,
                                                                  partials, ticket, out_f64, use_smem);
    });
    return check_launch("contractk_kernel");
}

}  // extern "C"
