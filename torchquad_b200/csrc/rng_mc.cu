// Philox uniform streams, Monte Carlo sample generation and the fp64 column reduction.
// Replaces rng.py:119-125, monte_carlo.py:84-106 and the sum of monte_carlo.py:72-77.
#include "common.cuh"

namespace tq {

// One thread per (row, Philox block) item, with the block index FIXED per thread: thread t of a CTA owns
// Philox block `t % nblk` of rows `t / nblk + k*rows_per_pass`, so everything that depends on the column
// (domain start/size, number of valid lanes, store width) is loop-invariant and lives in registers, the row
// advances by a constant stride, and the threads of a warp still write one contiguous span of the row-major
// output.  Why this shape: on B200 the kernel is bound by instruction issue, not by HBM -- IMAD.WIDE (2 per
// Philox round) is quarter rate, which caps the chip at ~4.8e11 Philox blocks/s = 7.6 TB/s of uniforms
// (tq_peak_microbench kind 2), so every ALU-pipe instruction spent on indexing or predicates costs bandwidth.
// Conversion: u = m * 2^-24 (m < 2^24, exact) and x = u*size + start; the exact power-of-two scale is folded
// into `size` on the host side of the loop (RN(m*2^-24*size) is unchanged), keeping bit parity with
// mul-then-add of the reference (monte_carlo.py:106).
template <typename T> struct RawBits;
template <> struct RawBits<float> {
    __device__ __forceinline__ static void get(const uint4& r, float* m) {
        m[0] = __uint2float_rn(r.x >> 8); m[1] = __uint2float_rn(r.y >> 8);
        m[2] = __uint2float_rn(r.z >> 8); m[3] = __uint2float_rn(r.w >> 8);
    }
    static constexpr float SCALE = 5.9604644775390625e-08f;  // 2^-24
};
template <> struct RawBits<double> {
    __device__ __forceinline__ static void get(const uint4& r, double* m) {
        m[0] = __ull2double_rn((((unsigned long long)r.y << 32) | r.x) >> 11);
        m[1] = __ull2double_rn((((unsigned long long)r.w << 32) | r.z) >> 11);
    }
    static constexpr double SCALE = 1.1102230246251565e-16;  // 2^-53
};

#ifndef TQ_UNIFORM_ROWS
#define TQ_UNIFORM_ROWS 1
#endif
template <typename T, bool AFFINE>
__global__ void __launch_bounds__(256)
uniform_kernel(T* __restrict__ out, const T* __restrict__ domain, int64_t row_begin, int64_t nrows,
               int dim, int nblk, uint64_t seed, uint32_t call, const uint32_t* __restrict__ call_offset) {
    constexpr int LANES = U01<T>::LANES;
    if (call_offset) call += *call_offset;  // replayed launches (CUDA graphs) advance the stream on the device
    const int rows_per_pass = 256 / nblk;
    const int rloc = threadIdx.x / nblk;
    const int blk = threadIdx.x - rloc * nblk;
    if (rloc >= rows_per_pass) return;  // idle tail threads when 256 % nblk != 0 (no barriers below)
    const int d0 = blk * LANES;
    const int nvalid = dim - d0 < LANES ? dim - d0 : LANES;
    T sz[LANES], st[LANES];
#pragma unroll
    for (int j = 0; j < LANES; ++j) {
        sz[j] = RawBits<T>::SCALE;
        st[j] = (T)0;
        if (AFFINE && j < nvalid) {
            const T a = domain[2 * (d0 + j)], b = domain[2 * (d0 + j) + 1];
            sz[j] = mul_rn(sub_rn(b, a), RawBits<T>::SCALE);  // exact: power-of-two scaling
            st[j] = a;
        }
    }
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int64_t stride = (int64_t)gridDim.x * rows_per_pass;
    int64_t row = (int64_t)blockIdx.x * rows_per_pass + rloc;
    T* p = out + row * dim + d0;
    const int64_t pstride = stride * dim;
    // store width: 16 B when the column block is complete and every row start is 16-byte aligned, else 8 B
    // pairs when rows start 8-byte aligned, else scalars (decided once per thread)
    const bool base16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int mode = (nvalid == LANES && base16 && (dim % LANES) == 0) ? 2
                   : (sizeof(T) == 4 && base16 && (dim % 2) == 0 && (nvalid % 2) == 0) ? 1 : 0;
    // TQ_UNIFORM_ROWS rows per step side by side (independent Philox chains).  Unlike the fused Monte Carlo kernel this one does
    // NOT gain from the extra instruction-level parallelism: it is bound by the quarter-rate IMAD.WIDE pipe (math-pipe throttle),
    // not by issue latency -- measured dim 10 fp32: 1 / 2 / 4 rows = 3784 / 3330 / 3347 GB/s (profiles/r2/exp_uniform_rows.txt).
    constexpr int R = TQ_UNIFORM_ROWS;
    for (; row < nrows; row += R * stride, p += R * pstride) {
        uint4 r[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const uint64_t grow = (uint64_t)(row_begin + row + k * stride);
            r[k] = Philox::run((uint32_t)grow, (uint32_t)(grow >> 32), (uint32_t)blk, call, k0, k1);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (k > 0 && row + k * stride >= nrows) break;
            T* pk = p + k * pstride;
            T v[LANES];
            RawBits<T>::get(r[k], v);
#pragma unroll
            for (int j = 0; j < LANES; ++j) v[j] = AFFINE ? add_rn(mul_rn(v[j], sz[j]), st[j]) : mul_rn(v[j], sz[j]);
            if (mode == 2) {
                if constexpr (LANES == 4) __stcs(reinterpret_cast<float4*>(pk), make_float4(v[0], v[1], v[2], v[3]));
                else __stcs(reinterpret_cast<double2*>(pk), make_double2(v[0], v[1]));
            } else if (mode == 1) {
                if constexpr (LANES == 4) {
                    __stcs(reinterpret_cast<float2*>(pk), make_float2(v[0], v[1]));
                    if (nvalid == 4) __stcs(reinterpret_cast<float2*>(pk) + 1, make_float2(v[2], v[3]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < LANES; ++j)
                    if (j < nvalid) pk[j] = v[j];
            }
        }
    }
}

template <typename T>
static int launch_uniform(T* out, const T* domain, int64_t row_begin, int64_t row_end, int dim,
                          uint64_t seed, uint32_t call, cudaStream_t st, const uint32_t* call_offset = nullptr) {
    const int64_t nrows = row_end - row_begin;
    if (nrows <= 0) return TQ_OK;
    const int nblk = (dim + U01<T>::LANES - 1) / U01<T>::LANES;
    const int rows_per_pass = 256 / nblk;
    const int64_t passes = (nrows + rows_per_pass - 1) / rows_per_pass;
    const int64_t cap = (int64_t)num_sms() * 8;
    const int grid = (int)(passes < cap ? passes : cap);
    if (domain)
        uniform_kernel<T, true><<<TQ_GRID(grid), 256, 0, st>>>(out, domain, row_begin, nrows, dim, nblk, seed, call, call_offset);
    else
        uniform_kernel<T, false><<<TQ_GRID(grid), 256, 0, st>>>(out, nullptr, row_begin, nrows, dim, nblk, seed, call, call_offset);
    return check_launch("uniform_kernel");
}

// grad wrt domain of x = u*(b-a) + a:  d/da = 1-u, d/db = u.  One partial per CTA and dimension, then a
// deterministic last-CTA sum.  Thread t walks rows t, t+stride, ... of a fixed block of dimensions.
template <typename T>
__global__ void __launch_bounds__(256)
mc_sample_backward_kernel(const T* __restrict__ g, int64_t row_begin, int64_t nrows, int dim, uint64_t seed,
                          uint32_t call, double* __restrict__ partials, unsigned int* ticket,
                          double* __restrict__ out) {
    constexpr int LANES = U01<T>::LANES;
    __shared__ double sh[32 * 2];
    __shared__ bool is_last;
    const int nblk = (dim + LANES - 1) / LANES;
    // partials layout: [gridDim.x][dim][2]
    for (int blk = 0; blk < nblk; ++blk) {
        double ga[LANES], gb[LANES];
#pragma unroll
        for (int j = 0; j < LANES; ++j) ga[j] = gb[j] = 0.0;
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
            const uint64_t grow = (uint64_t)(row_begin + r);
            T u[LANES];
            philox_block<T>(seed, call, (uint32_t)grow, (uint32_t)(grow >> 32), blk, u);
#pragma unroll
            for (int j = 0; j < LANES; ++j) {
                int d = blk * LANES + j;
                if (d < dim) {
                    double gv = (double)g[r * dim + d];
                    gb[j] += gv * (double)u[j];
                    ga[j] += gv * (1.0 - (double)u[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < LANES; ++j) {
            int d = blk * LANES + j;
            double v[2] = {ga[j], gb[j]};
            block_sum<2>(v, sh);
            if (threadIdx.x == 0 && d < dim) {
                partials[((size_t)blockIdx.x * dim + d) * 2 + 0] = v[0];
                partials[((size_t)blockIdx.x * dim + d) * 2 + 1] = v[1];
            }
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int e = threadIdx.x; e < dim * 2; e += blockDim.x) {
            double acc = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) acc += __ldcg(&partials[(size_t)b * dim * 2 + e]);
            out[e] = acc;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// Column sums of f[rows, cols] in fp64.  cols == 1: vectorised grid-stride loads, deterministic tree.
template <typename T, bool SQ>
__global__ void __launch_bounds__(256)
sum1_kernel(const T* __restrict__ f, int64_t n, double* partials, unsigned int* ticket, double* out_s, double* out_q) {
    constexpr int V = 16 / sizeof(T);
    __shared__ double sh[32 * 2];
    double s = 0.0, q = 0.0;
    const int64_t nvec = ((reinterpret_cast<uintptr_t>(f) & 15) == 0) ? n / V : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = tid; i < nvec; i += stride) {
        uint4 raw = __ldcs(reinterpret_cast<const uint4*>(f) + i);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            double x = (double)e[j];
            s += x;
            if (SQ) q += x * x;
        }
    }
    for (int64_t i = nvec * V + tid; i < n; i += stride) {
        double x = (double)f[i];
        s += x;
        if (SQ) q += x * x;
    }
    double v[2] = {s, q};
    grid_sum_finish_split(v, sh, partials, ticket, out_s, SQ ? out_q : nullptr);  // straight into the caller's tensors
}

// cols > 1: thread t of the grid owns flat elements t, t+S, ... with S a multiple of cols, so its
// column is fixed; per-CTA shared accumulators per column, then last-CTA finish.
template <typename T, bool SQ>
__global__ void __launch_bounds__(256)
sumk_kernel(const T* __restrict__ f, int64_t rows, int64_t cols, int64_t S, double* partials,
            unsigned int* ticket, double* out_s, double* out_q) {
    extern __shared__ double sacc[];  // [cols*2]
    __shared__ bool is_last;
    for (int64_t c = threadIdx.x; c < cols * 2; c += blockDim.x) sacc[c] = 0.0;
    __syncthreads();
    const int64_t n = rows * cols;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < S) {
        double s = 0.0, q = 0.0;
        constexpr int U = 8;  // independent loads in flight per thread (HBM latency x bandwidth needs ~5 MB in flight)
        int64_t i = tid;
        for (; i + (U - 1) * S < n; i += U * S) {
            T v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldcs(f + i + u * S);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double x = (double)v[u];
                s += x;
                if (SQ) q += x * x;
            }
        }
        for (; i < n; i += S) {
            const double x = (double)f[i];
            s += x;
            if (SQ) q += x * x;
        }
        const int64_t c = tid % cols;
        atomicAdd(&sacc[c], s);
        if (SQ) atomicAdd(&sacc[cols + c], q);
    }
    __syncthreads();
    for (int64_t c = threadIdx.x; c < cols * 2; c += blockDim.x) partials[(size_t)blockIdx.x * cols * 2 + c] = sacc[c];
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int64_t c = threadIdx.x; c < cols * 2; c += blockDim.x) {
            double acc = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) acc += __ldcg(&partials[(size_t)b * cols * 2 + c]);
            if (c < cols) out_s[c] = acc;
            else if (SQ) out_q[c - cols] = acc;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

}  // namespace tq

using namespace tq;

extern "C" {

int tq_philox_uniform(void* out, int64_t row_begin, int64_t row_end, int32_t dim, int32_t dtype,
                      uint64_t seed, uint32_t call_idx, void* stream) {
    TQ_REQUIRE(dim >= 1 && dim <= 512, "tq_philox_uniform: dim %d out of range (max 512)", dim);
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_philox_uniform: bad row range");
    TQ_DISPATCH_DTYPE(dtype, {
        const int nblk = (dim + U01<T>::LANES - 1) / U01<T>::LANES;
        TQ_REQUIRE(nblk <= 256, "tq_philox_uniform: dim %d too large", dim);
        return launch_uniform<T>((T*)out, nullptr, row_begin, row_end, dim, seed, call_idx, as_stream(stream));
    });
    return TQ_OK;
}

int tq_mc_sample(void* out, const void* domain, int64_t row_begin, int64_t row_end, int32_t dim,
                 int32_t dtype, uint64_t seed, uint32_t call_idx, void* stream) {
    TQ_REQUIRE(dim >= 1 && dim <= 512, "tq_mc_sample: dim %d out of range (max 512)", dim);
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_mc_sample: bad row range");
    TQ_REQUIRE(domain != nullptr, "tq_mc_sample: domain is NULL");
    TQ_DISPATCH_DTYPE(dtype, {
        return launch_uniform<T>((T*)out, (const T*)domain, row_begin, row_end, dim, seed, call_idx, as_stream(stream));
    });
    return TQ_OK;
}

int tq_mc_sample_replayable(void* out, const void* domain, int64_t row_begin, int64_t row_end, int32_t dim,
                            int32_t dtype, uint64_t seed, uint32_t call_idx, const uint32_t* call_offset_dev,
                            void* stream) {
    TQ_REQUIRE(dim >= 1 && dim <= 512, "tq_mc_sample_replayable: dim %d out of range (max 512)", dim);
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_mc_sample_replayable: bad row range");
    TQ_REQUIRE(domain != nullptr && call_offset_dev != nullptr, "tq_mc_sample_replayable: NULL argument");
    TQ_DISPATCH_DTYPE(dtype, {
        return launch_uniform<T>((T*)out, (const T*)domain, row_begin, row_end, dim, seed, call_idx, as_stream(stream),
                                 call_offset_dev);
    });
    return TQ_OK;
}

int tq_mc_sample_backward(const void* grad_out, int64_t row_begin, int64_t row_end, int32_t dim,
                          int32_t dtype, uint64_t seed, uint32_t call_idx, double* grad_domain_f64,
                          void* ws, size_t ws_bytes, void* stream) {
    TQ_REQUIRE(dim >= 1 && dim <= 512, "tq_mc_sample_backward: dim %d out of range (max 512)", dim);
    const int64_t nrows = row_end - row_begin;
    TQ_REQUIRE(nrows >= 0, "tq_mc_sample_backward: bad row range");
    Workspace w(ws, ws_bytes);
    unsigned int* ticket = w.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    const int grid = grid_for(nrows, 256, 4);
    double* partials = w.take<double>((size_t)grid * dim * 2);
    if (!ticket || !partials) { set_error("tq_mc_sample_backward: workspace too small"); return TQ_ERR_WORKSPACE; }
    TQ_DISPATCH_DTYPE(dtype, {
        mc_sample_backward_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((const T*)grad_out, row_begin, nrows, dim, seed,
                                                                         call_idx, partials, ticket, grad_domain_f64);
    });
    return check_launch("mc_sample_backward_kernel");
}

int tq_sum_columns(const void* f, int64_t rows, int64_t cols, int32_t dtype, double* sum_f64,
                   double* sumsq_f64, void* ws, size_t ws_bytes, void* stream) {
    TQ_REQUIRE(rows >= 0 && cols >= 1, "tq_sum_columns: bad shape [%lld, %lld]", (long long)rows, (long long)cols);
    Workspace w(ws, ws_bytes);
    unsigned int* ticket = w.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    cudaStream_t st = as_stream(stream);
    if (cols == 1) {
        const int grid = grid_for((rows + 3) / 4, 256, 4);
        double* partials = w.take<double>((size_t)grid * 2);
        if (!ticket || !partials) { set_error("tq_sum_columns: workspace too small"); return TQ_ERR_WORKSPACE; }
        TQ_DISPATCH_DTYPE(dtype, {
            if (sumsq_f64) sum1_kernel<T, true><<<TQ_GRID(grid), 256, 0, st>>>((const T*)f, rows, partials, ticket, sum_f64, sumsq_f64);
            else sum1_kernel<T, false><<<TQ_GRID(grid), 256, 0, st>>>((const T*)f, rows, partials, ticket, sum_f64, nullptr);
        });
        return check_launch("sum1_kernel");
    }
    TQ_REQUIRE(cols <= 2048, "tq_sum_columns: at most 2048 integrand components (got %lld)", (long long)cols);
    int grid = grid_for((rows * cols + 7) / 8, 256, 8);
    // the per-CTA partials (cols * 2 doubles each) must fit the caller's workspace
    const size_t ws_left = ws_bytes > (size_t)(64 << 10) ? ws_bytes - (size_t)(64 << 10) : 0;
    const int64_t grid_cap = (int64_t)(ws_left / ((size_t)cols * 2 * sizeof(double)));
    if (grid > grid_cap) grid = (int)(grid_cap > 0 ? grid_cap : 1);
    // stride S: largest multiple of cols not exceeding the thread count (at least cols)
    int64_t threads = (int64_t)grid * 256;
    if (threads < cols) { grid = (int)((cols + 255) / 256); threads = (int64_t)grid * 256; }
    const int64_t S = (threads / cols) * cols;
    double* partials = w.take<double>((size_t)grid * cols * 2);
    if (!ticket || !partials) { set_error("tq_sum_columns: workspace too small"); return TQ_ERR_WORKSPACE; }
    const size_t smem = (size_t)cols * 2 * sizeof(double);
    TQ_DISPATCH_DTYPE(dtype, {
        if (sumsq_f64) sumk_kernel<T, true><<<TQ_GRID(grid), 256, smem, st>>>((const T*)f, rows, cols, S, partials, ticket, sum_f64, sumsq_f64);
        else sumk_kernel<T, false><<<TQ_GRID(grid), 256, smem, st>>>((const T*)f, rows, cols, S, partials, ticket, sum_f64, nullptr);
    });
    return check_launch("sumk_kernel");
}

}  // extern "C"
