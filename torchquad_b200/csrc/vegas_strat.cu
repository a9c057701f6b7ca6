// VEGAS+ adaptive stratification: per-cube sample counts, stratified sampling, per-cube sums, damped
// variance update and the per-iteration estimator.
// Replaces torchquad/integration/vegas_stratification.py (get_NH :92-103, _get_indices/get_Y :105-165,
// accumulate_weight :46-70, update_DH :72-90) and the estimator of vegas.py:293-303.
#include "common.cuh"
#include "strat_tile.cuh"
#include "internal.cuh"
#include "vegas_dev.cuh"

namespace tq {

constexpr int ST_TILE = 1024;  // cubes (or rows) per CTA tile: 256 threads x 4

// ---- get_NH + exclusive scan (3 phases: tile sums, scan of tile sums, tile scans)
template <typename T>
__global__ void __launch_bounds__(256)
nh_tile_sum_kernel(const T* __restrict__ dh, int64_t n_cubes, T nev, long long* __restrict__ tile_sums) {
    __shared__ long long sh[33];
    const int64_t c0 = (int64_t)blockIdx.x * ST_TILE + threadIdx.x * 4;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (c0 + i < n_cubes) s += nh_of<T>(dh[c0 + i], nev);
    long long total;
    block_excl_scan<long long>(s, sh, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
i64_tile_scan_kernel(long long* __restrict__ tile_sums, int64_t ntiles, long long* __restrict__ total_out) {
    __shared__ long long sh[33];
    long long carry = 0;
    for (int64_t base = 0; base < ntiles; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < ntiles ? tile_sums[i] : 0;
        long long total;
        const long long ex = block_excl_scan<long long>(v, sh, total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

template <typename T>
__global__ void __launch_bounds__(256)
nh_scan_kernel(const T* __restrict__ dh, int64_t n_cubes, T nev, const long long* __restrict__ tile_offsets,
               long long* __restrict__ nh, long long* __restrict__ offsets) {
    __shared__ long long sh[33];
    const int64_t c0 = (int64_t)blockIdx.x * ST_TILE + threadIdx.x * 4;
    long long v[4], run = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = (c0 + i < n_cubes) ? nh_of<T>(dh[c0 + i], nev) : 0;
        run += v[i];
    }
    long long total;
    long long ex = block_excl_scan<long long>(run, sh, total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (c0 + i < n_cubes) {
            nh[c0 + i] = v[i];
            offsets[c0 + i] = ex;
        }
        ex += v[i];
    }
}

// ---- exclusive scan of a caller-provided nh
__global__ void __launch_bounds__(256)
i64_tile_sum_kernel(const long long* __restrict__ v, int64_t n, long long* __restrict__ tile_sums) {
    __shared__ long long sh[33];
    const int64_t c0 = (int64_t)blockIdx.x * ST_TILE + threadIdx.x * 4;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (c0 + i < n) s += v[c0 + i];
    long long total;
    block_excl_scan<long long>(s, sh, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
i64_scan_kernel(const long long* __restrict__ in, int64_t n, const long long* __restrict__ tile_offsets,
                long long* __restrict__ offsets) {
    __shared__ long long sh[33];
    const int64_t c0 = (int64_t)blockIdx.x * ST_TILE + threadIdx.x * 4;
    long long v[4], run = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = (c0 + i < n) ? in[c0 + i] : 0;
        run += v[i];
    }
    long long total;
    long long ex = block_excl_scan<long long>(run, sh, total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (c0 + i < n) offsets[c0 + i] = ex;
        ex += v[i];
    }
}

// ---- row -> cube lookup shared by the sampling and backward kernels.
// A CTA tile of ST_TILE consecutive rows overlaps at most ST_TILE/2 + 1 cubes (nh >= 2), so the slice of
// `offsets` it needs fits in shared memory: two global binary searches per tile, then per-row searches
// in shared memory.
struct CubeSlice {
    long long c_lo;
    int count;  // cubes in the slice; s_off holds count+1 entries
};

__device__ __forceinline__ long long upper_cube(const long long* __restrict__ offsets, int64_t n_cubes, long long row) {
    // largest c in [0, n_cubes) with offsets[c] <= row
    long long lo = 0, hi = n_cubes - 1;
    while (lo < hi) {
        const long long mid = (lo + hi + 1) >> 1;
        if (__ldg(&offsets[mid]) <= row) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ CubeSlice load_cube_slice(const long long* __restrict__ offsets, int64_t n_cubes,
                                                     long long row_lo, long long row_hi /*inclusive*/,
                                                     long long* s_off, long long* s_bounds) {
    if (threadIdx.x == 0) s_bounds[0] = upper_cube(offsets, n_cubes, row_lo);
    if (threadIdx.x == 32) s_bounds[1] = upper_cube(offsets, n_cubes, row_hi);
    __syncthreads();
    CubeSlice s;
    s.c_lo = s_bounds[0];
    s.count = (int)(s_bounds[1] - s_bounds[0] + 1);
    for (int i = threadIdx.x; i <= s.count; i += blockDim.x) s_off[i] = __ldg(&offsets[s.c_lo + i]);
    __syncthreads();
    return s;
}

__device__ __forceinline__ int cube_in_slice(const long long* s_off, int count, long long row) {
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_off[mid] <= row) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- get_Y
// Each CTA walks a contiguous chunk of rows in tiles of ST_ROWS (row -> cube through StratTile).  Inside a
// tile, thread t owns Philox block `t % nblk` (LANES consecutive dimensions) of rows t / nblk + k*rows_per_pass:
// the column block is loop-invariant (digit divisor, store width) and a warp writes one contiguous span of y.
template <typename T>
__global__ void __launch_bounds__(256)
strat_sample_kernel(const long long* __restrict__ offsets, int64_t n_cubes, int n_strat, int dim, int nblk,
                    const T* __restrict__ u_in, uint64_t seed, uint32_t call, int64_t row_begin, int64_t row_end,
                    int64_t rows_per_cta, FastDiv fdn, T* __restrict__ y) {
    constexpr int LANES = U01<T>::LANES;
    __shared__ StratTile st;
    const int rows_per_pass = 256 / nblk;
    const int rloc = threadIdx.x / nblk;
    const int blk = threadIdx.x - rloc * nblk;
    const bool worker = rloc < rows_per_pass;
    const int d0 = blk * LANES;
    const int nvalid = dim - d0 < LANES ? dim - d0 : LANES;
    // n_strat^d0, saturated: the first digit of this thread's block is (cube / div0) % n_strat
    unsigned long long pw = 1;
    for (int i = 0; i < d0 && pw <= 0xffffffffull; ++i) pw *= (unsigned long long)n_strat;
    FastDiv fd0;
    fd0.set(pw > 0xffffffffull ? 0xffffffffu : (uint32_t)pw);
    const uint32_t ns = (uint32_t)n_strat;
    const T nsf = (T)n_strat;
    const T inv_ns = div_rn((T)1, nsf);
    const bool base16 = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(u_in)) & 15) == 0;
    const int mode = (nvalid == LANES && base16 && (dim % LANES) == 0) ? 2 : 0;
    const int64_t r_lo = row_begin + (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_hi = r_lo + rows_per_cta < row_end ? r_lo + rows_per_cta : row_end;
    if (threadIdx.x == 0 && r_lo < r_hi) st.first = cube_of_row(offsets, n_cubes, r_lo);
    __syncthreads();
    long long c_lo = r_lo < r_hi ? st.first : 0;
    int buf = 0;
    for (int64_t rb = r_lo; rb < r_hi; rb += ST_ROWS, buf ^= 1) {
        const int64_t re = rb + ST_ROWS < r_hi ? rb + ST_ROWS : r_hi;
        strat_tile_fill(st, buf, offsets, n_cubes, c_lo, rb, re);
        const int n_here = (int)(re - rb);
        if (worker) {
            for (int rl = rloc; rl < n_here; rl += rows_per_pass) {
                const int ci = st.cube[buf][rl];
                const uint32_t cube = (uint32_t)(c_lo + ci);
                const int64_t row = rb + rl;
                const uint32_t k = (uint32_t)(row - st.off[buf][ci]);
                T* out = y + (row - row_begin) * dim + d0;
                alignas(16) T u[LANES];
                if (u_in) {
                    const T* uin = u_in + (row - row_begin) * dim + d0;
                    if (mode == 2) *reinterpret_cast<uint4*>(u) = __ldcs(reinterpret_cast<const uint4*>(uin));
                    else {
#pragma unroll
                        for (int j = 0; j < LANES; ++j) u[j] = j < nvalid ? uin[j] : (T)0;
                    }
                } else {
                    philox_block<T>(seed, call, cube, k, (uint32_t)blk, u);
                }
                uint32_t c = fd0.div(cube);
#pragma unroll
                for (int j = 0; j < LANES; ++j) {
                    const uint32_t q = fdn.div(c);
                    const uint32_t p = c - q * ns;
                    c = q;
                    // Philox uniforms are multiples of 2^-24 / 2^-53: the multiply-FMA-FMA quotient is the correctly rounded
                    // one (div_by_const); injected uniforms may be anything (subnormals), they keep the IEEE division
                    const T a = add_rn((T)p, u[j]);
                    T v = u_in ? div_rn(a, nsf) : div_by_const(a, nsf, inv_ns);
                    if (v >= (T)1) v = (T)0.999999;
                    u[j] = v;
                }
                if (mode == 2) {
                    __stcs(reinterpret_cast<uint4*>(out), *reinterpret_cast<uint4*>(u));
                } else {
#pragma unroll
                    for (int j = 0; j < LANES; ++j)
                        if (j < nvalid) out[j] = u[j];
                }
            }
        }
        c_lo += st.cube[buf][n_here - 1];
    }
}

// ---- accumulate_weight: one thread per cube, rows summed in order (bit-identical to the CPU scatter_add_).
// Cubes holding more than ST_HEAVY rows (peaked dh) would serialise one thread for a long time; those are
// summed by the whole warp instead (stride-32 coalesced loads, fp64 accumulation, one rounding), which is
// within 1 ulp of the exact sum rather than bit-identical to the reference's sequential order.
constexpr long long ST_HEAVY = 256;

template <typename T>
__global__ void __launch_bounds__(256)
strat_accumulate_kernel(const T* __restrict__ jf, int64_t row_base, const long long* __restrict__ offsets,
                        int64_t cube_begin, int64_t cube_end, T* __restrict__ JF, T* __restrict__ JF2) {
    const int lane = threadIdx.x & 31;
    for (int64_t cw = cube_begin + (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); cw < cube_end;
         cw += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = cw + lane;
        const bool valid = c < cube_end;
        const long long r0 = valid ? offsets[c] - row_base : 0, r1 = valid ? offsets[c + 1] - row_base : 0;
        const bool heavy = r1 - r0 > ST_HEAVY;
        if (valid && !heavy) {
            T s = (T)0, q = (T)0;
            for (long long r = r0; r < r1; ++r) {
                const T v = jf[r];
                s = add_rn(s, v);
                q = add_rn(q, mul_rn(v, v));
            }
            JF[c] = s;
            JF2[c] = q;
        }
        unsigned m = __ballot_sync(0xffffffffu, heavy);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const long long R0 = __shfl_sync(0xffffffffu, r0, src), R1 = __shfl_sync(0xffffffffu, r1, src);
            double s = 0.0, q = 0.0;
            for (long long r = R0 + lane; r < R1; r += 32) {
                const double v = (double)jf[r];
                s += v;
                q += v * v;
            }
            s = warp_sum(s);
            q = warp_sum(q);
            if (lane == 0) {
                JF[cw + src] = (T)s;
                JF2[cw + src] = (T)q;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
strat_accumulate_backward_kernel(const T* __restrict__ gJF, const long long* __restrict__ offsets, int64_t n_cubes,
                                 int64_t row_begin, int64_t row_end, T* __restrict__ gjf) {
    __shared__ long long s_off[ST_TILE / 2 + 4];
    __shared__ long long s_bounds[2];
    for (int64_t rb = row_begin + (int64_t)blockIdx.x * ST_TILE; rb < row_end; rb += (int64_t)gridDim.x * ST_TILE) {
        const int64_t re = rb + ST_TILE < row_end ? rb + ST_TILE : row_end;
        const CubeSlice sl = load_cube_slice(offsets, n_cubes, rb, re - 1, s_off, s_bounds);
        for (int64_t row = rb + threadIdx.x; row < re; row += blockDim.x) {
            const int ci = cube_in_slice(s_off, sl.count, row);
            gjf[row - row_begin] = gJF[sl.c_lo + ci];
        }
        __syncthreads();
    }
}

// ---- estimator + update_DH
template <typename T>
__global__ void __launch_bounds__(256)
strat_update_kernel(const T* __restrict__ JF, const T* __restrict__ JF2, const long long* __restrict__ nh,
                    int64_t n_cubes, T V, T V2, T beta, T* __restrict__ dh, double* partials, unsigned int* ticket,
                    double* scalars) {
    __shared__ double sh[32 * 4];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cubes;
         c += (int64_t)gridDim.x * blockDim.x) {
        const long long nc = nh[c];
        const T n = (T)nc;
        const T inv = div_rn((T)1, n);
        const T jf = JF[c], jf2 = JF2[c];
        // vegas.py:293-303
        const T ih = mul_rn(jf, mul_rn(inv, V));
        const T sig2 = fabs(sub_rn(mul_rn(mul_rn(jf2, inv), V2), mul_rn(ih, ih)));
        acc[0] += (double)ih;
        acc[1] += (double)mul_rn(sig2, inv);
        // vegas_stratification.py:78-85
        const T m = div_rn(mul_rn(V, jf), n);
        T dv = sub_rn(div_rn(mul_rn(V2, jf2), n), mul_rn(m, m));
        if (dv < (T)0) dv = (T)0;
        const T p = pow(dv, beta);
        dh[c] = p;
        acc[2] += (double)p;
        acc[3] += (double)nc;
    }
    grid_sum_finish<4>(acc, sh, partials, ticket, scalars);
}

template <typename T>
__global__ void __launch_bounds__(256)
strat_normalise_kernel(T* __restrict__ dh, int64_t n_cubes, const double* __restrict__ scalars) {
    const T s = (T)scalars[2];
    if (s == (T)0) return;  // vegas_stratification.py:89-90
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cubes;
         c += (int64_t)gridDim.x * blockDim.x)
        dh[c] = div_rn(dh[c], s);
}

// get_NH + offsets; `clear` (optional, 4-byte aligned) is zeroed on the same stream before the pass that
// follows -- the native loop (vegas_driver.cu) folds its JF/JF2 reset into this launch.
int strat_nh_launch(const void* dh, int64_t n_cubes, double nevals_exp, int32_t dtype, int64_t* nh, int64_t* offsets,
                    void* clear, size_t clear_bytes, void* ws, size_t ws_bytes, void* stream) {
    TQ_REQUIRE(n_cubes >= 1, "tq_vegas_strat_nh: n_cubes must be positive");
    Workspace w(ws, ws_bytes);
    w.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    const int64_t ntiles = (n_cubes + ST_TILE - 1) / ST_TILE;
    long long* tile_sums = w.take<long long>((size_t)ntiles);
    if (!tile_sums) { set_error("tq_vegas_strat_nh: workspace too small for %lld cubes", (long long)n_cubes); return TQ_ERR_WORKSPACE; }
    cudaStream_t st = as_stream(stream);
    if (small_strat_ok(n_cubes)) return small_nh_launch(dh, n_cubes, nevals_exp, dtype, nh, offsets, clear, clear_bytes, stream);
    if (clear && clear_bytes) cudaMemsetAsync(clear, 0, clear_bytes, st);
    TQ_DISPATCH_DTYPE(dtype, {
        nh_tile_sum_kernel<T><<<TQ_GRID((unsigned)ntiles), 256, 0, st>>>((const T*)dh, n_cubes, (T)nevals_exp, tile_sums);
        i64_tile_scan_kernel<<<TQ_GRID(1), 256, 0, st>>>(tile_sums, ntiles, (long long*)offsets + n_cubes);
        nh_scan_kernel<T><<<TQ_GRID((unsigned)ntiles), 256, 0, st>>>((const T*)dh, n_cubes, (T)nevals_exp, tile_sums,
                                                           (long long*)nh, (long long*)offsets);
    });
    return check_launch("tq_vegas_strat_nh");
}

int strat_update_partial_launch(const void* JF, const void* JF2, const int64_t* nh, int64_t n_cubes, double v_cubes, double beta,
                                int32_t dtype, void* dh, double* scalars, void* ws, size_t ws_bytes, void* stream) {
    Workspace w(ws, ws_bytes);
    unsigned int* ticket = w.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    const int grid = grid_for(n_cubes, 256, 4);
    double* partials = w.take<double>((size_t)grid * 4);
    if (!ticket || !partials) { set_error("strat_update: workspace too small"); return TQ_ERR_WORKSPACE; }
    TQ_DISPATCH_DTYPE(dtype, {
        strat_update_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((const T*)JF, (const T*)JF2, (const long long*)nh, n_cubes,
                                                                   (T)v_cubes, (T)(v_cubes * v_cubes), (T)beta, (T*)dh, partials,
                                                                   ticket, scalars);
    });
    return check_launch("strat_update_kernel");
}

int strat_normalise_launch(void* dh, int64_t n_cubes, const double* scalars, int32_t dtype, void* stream) {
    const int grid = grid_for(n_cubes, 256, 4);
    TQ_DISPATCH_DTYPE(dtype, {
        strat_normalise_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((T*)dh, n_cubes, scalars);
    });
    return check_launch("strat_normalise_kernel");
}

}  // namespace tq

using namespace tq;

extern "C" {

int tq_vegas_strat_nh(const void* dh, int64_t n_cubes, double nevals_exp, int32_t dtype, int64_t* nh,
                      int64_t* offsets, void* ws, size_t ws_bytes, void* stream) {
    return strat_nh_launch(dh, n_cubes, nevals_exp, dtype, nh, offsets, nullptr, 0, ws, ws_bytes, stream);
}

int tq_vegas_strat_offsets(const int64_t* nh, int64_t n_cubes, int64_t* offsets, void* ws, size_t ws_bytes,
                           void* stream) {
    TQ_REQUIRE(n_cubes >= 1, "tq_vegas_strat_offsets: n_cubes must be positive");
    Workspace w(ws, ws_bytes);
    w.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    const int64_t ntiles = (n_cubes + ST_TILE - 1) / ST_TILE;
    long long* tile_sums = w.take<long long>((size_t)ntiles);
    if (!tile_sums) { set_error("tq_vegas_strat_offsets: workspace too small for %lld cubes", (long long)n_cubes); return TQ_ERR_WORKSPACE; }
    cudaStream_t st = as_stream(stream);
    i64_tile_sum_kernel<<<TQ_GRID((unsigned)ntiles), 256, 0, st>>>((const long long*)nh, n_cubes, tile_sums);
    i64_tile_scan_kernel<<<TQ_GRID(1), 256, 0, st>>>(tile_sums, ntiles, (long long*)offsets + n_cubes);
    i64_scan_kernel<<<TQ_GRID((unsigned)ntiles), 256, 0, st>>>((const long long*)nh, n_cubes, tile_sums, (long long*)offsets);
    return check_launch("tq_vegas_strat_offsets");
}

int tq_vegas_strat_sample(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim,
                          int32_t dtype, const void* u_in, uint64_t seed, uint32_t call_idx,
                          int64_t row_begin, int64_t row_end, void* y, void* stream) {
    TQ_REQUIRE(n_cubes >= 1 && n_cubes < (1LL << 31), "tq_vegas_strat_sample: n_cubes out of range");
    TQ_REQUIRE(n_strat >= 1 && dim >= 1, "tq_vegas_strat_sample: bad n_strat/dim");
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_vegas_strat_sample: bad row range");
    if (row_end == row_begin) return TQ_OK;
    const int64_t nrows = row_end - row_begin;
    const int64_t tiles = (nrows + ST_ROWS - 1) / ST_ROWS;
    int64_t ctas = tiles < (int64_t)num_sms() * 8 ? tiles : (int64_t)num_sms() * 8;
    int64_t rows_per_cta = (nrows + ctas - 1) / ctas;
    rows_per_cta = ((rows_per_cta + ST_ROWS - 1) / ST_ROWS) * ST_ROWS;
    ctas = (nrows + rows_per_cta - 1) / rows_per_cta;
    TQ_DISPATCH_DTYPE(dtype, {
        const int nblk = (dim + U01<T>::LANES - 1) / U01<T>::LANES;
        TQ_REQUIRE(nblk <= 128, "tq_vegas_strat_sample: dim %d too large", dim);
        FastDiv fdn;
        fdn.set((uint32_t)n_strat);
        strat_sample_kernel<T><<<TQ_GRID((unsigned)ctas), 256, 0, as_stream(stream)>>>((const long long*)offsets, n_cubes, n_strat, dim, nblk,
                                                                             (const T*)u_in, seed, call_idx, row_begin, row_end,
                                                                             rows_per_cta, fdn, (T*)y);
    });
    return check_launch("strat_sample_kernel");
}

int tq_vegas_strat_accumulate(const void* jf, int64_t row_base, const int64_t* offsets,
                              int64_t cube_begin, int64_t cube_end, void* JF, void* JF2, int32_t dtype,
                              void* stream) {
    TQ_REQUIRE(cube_end >= cube_begin && cube_begin >= 0, "tq_vegas_strat_accumulate: bad cube range");
    if (cube_end == cube_begin) return TQ_OK;
    const int grid = grid_for(cube_end - cube_begin, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        strat_accumulate_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((const T*)jf, row_base, (const long long*)offsets,
                                                                       cube_begin, cube_end, (T*)JF, (T*)JF2);
    });
    return check_launch("strat_accumulate_kernel");
}

int tq_vegas_strat_accumulate_backward(const void* grad_JF, const int64_t* offsets, int64_t n_cubes,
                                       int64_t row_begin, int64_t row_end, void* grad_jf, int32_t dtype,
                                       void* stream) {
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_vegas_strat_accumulate_backward: bad row range");
    if (row_end == row_begin) return TQ_OK;
    const int grid = grid_for(row_end - row_begin, ST_TILE, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        strat_accumulate_backward_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>(
            (const T*)grad_JF, (const long long*)offsets, n_cubes, row_begin, row_end, (T*)grad_jf);
    });
    return check_launch("strat_accumulate_backward_kernel");
}

int tq_vegas_strat_update(const void* JF, const void* JF2, const int64_t* nh, int64_t n_cubes,
                          double v_cubes, double beta, int32_t dtype, void* dh, double* scalars_f64,
                          void* ws, size_t ws_bytes, void* stream) {
    TQ_REQUIRE(n_cubes >= 1, "tq_vegas_strat_update: n_cubes must be positive");
    Workspace w(ws, ws_bytes);
    unsigned int* ticket = w.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    const int grid = grid_for(n_cubes, 256, 4);
    double* partials = w.take<double>((size_t)grid * 4);
    if (!ticket || !partials) { set_error("tq_vegas_strat_update: workspace too small"); return TQ_ERR_WORKSPACE; }
    cudaStream_t st = as_stream(stream);
    if (small_strat_ok(n_cubes)) {
        SmallStrat a = {JF, JF2, const_cast<int64_t*>(nh), n_cubes, v_cubes, beta, dh, scalars_f64, 0.0, nullptr, nullptr, 0};
        return small_update_launch(&a, nullptr, dtype, stream);
    }
    TQ_DISPATCH_DTYPE(dtype, {
        strat_update_kernel<T><<<TQ_GRID(grid), 256, 0, st>>>((const T*)JF, (const T*)JF2, (const long long*)nh, n_cubes,
                                                    (T)v_cubes, (T)(v_cubes * v_cubes), (T)beta, (T*)dh, partials, ticket,
                                                    scalars_f64);
        strat_normalise_kernel<T><<<TQ_GRID(grid), 256, 0, st>>>((T*)dh, n_cubes, scalars_f64);
    });
    return check_launch("tq_vegas_strat_update");
}

}  // extern "C"
