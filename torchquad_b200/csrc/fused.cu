// Fused generate -> map -> evaluate -> accumulate kernels for built-in integrands: no sample ever
// reaches HBM.  The integrands are the Genz families (not in the reference) and the
// reference's test integrands (tests/integration_test_functions.py:146-325), all of which factor as
// finish(combine_d step(x_d)), so a thread streams over the dimensions without holding the point.
#include <stdlib.h>

#include "common.cuh"
#include "vegas_dev.cuh"
#include "internal.cuh"

namespace tq {

constexpr double TWO_PI = 6.283185307179586476925286766559;

// Per-dimension parameters staged in shared memory in the working type.
template <typename T>
struct FnShared {
    T a[TQ_MAX_DIM];      // difficulty
    T u[TQ_MAX_DIM];      // shift
    T c[TQ_MAX_DIM];      // family-specific precomputation (a^-2, a^2)
    T start[TQ_MAX_DIM];  // domain start
    T size[TQ_MAX_DIM];   // domain size
    T coeff[8];
    T scale;
    T phase;
    T expo;
    int dim, ncoeff;
};

template <typename T>
__device__ __forceinline__ void stage_integrand(const tq_integrand& P, FnShared<T>& S) {
    for (int d = threadIdx.x; d < P.dim; d += blockDim.x) {
        S.a[d] = (T)P.a[d];
        S.u[d] = (T)P.u[d];
        S.start[d] = (T)P.start[d];
        S.size[d] = (T)P.size[d];
        T c = (T)0;
        if (P.family == TQ_F_GENZ_PRODUCT_PEAK) c = (T)(1.0 / (P.a[d] * P.a[d]));
        if (P.family == TQ_F_GENZ_GAUSSIAN) c = (T)(P.a[d] * P.a[d]);
        S.c[d] = c;
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < 8; ++k) S.coeff[k] = (T)P.coeff[k];
        S.scale = (T)P.scale;
        S.phase = (T)(TWO_PI * P.u[0]);
        S.expo = (T)(-(double)(P.dim + 1));
        S.dim = P.dim;
        S.ncoeff = P.ncoeff;
    }
    __syncthreads();
}

template <int FAM, typename T>
struct Integrand {
    T acc;
    bool inside;
    __device__ __forceinline__ void init() {
        acc = (FAM == TQ_F_GENZ_PRODUCT_PEAK || FAM == TQ_F_PROD_COS || FAM == TQ_F_PROD_COS_FAST) ? (T)1 : (T)0;
        inside = true;
    }
    __device__ __forceinline__ void step(T x, int d, const FnShared<T>& S) {
        if constexpr (FAM == TQ_F_GENZ_OSCILLATORY || FAM == TQ_F_GENZ_CORNER_PEAK) {
            acc += S.a[d] * x;
        } else if constexpr (FAM == TQ_F_GENZ_PRODUCT_PEAK) {
            const T t = x - S.u[d];
            acc *= (T)1 / (S.c[d] + t * t);
        } else if constexpr (FAM == TQ_F_GENZ_GAUSSIAN) {
            const T t = x - S.u[d];
            acc += S.c[d] * t * t;
        } else if constexpr (FAM == TQ_F_GENZ_C0) {
            acc += S.a[d] * fabs(x - S.u[d]);
        } else if constexpr (FAM == TQ_F_GENZ_DISCONTINUOUS) {
            acc += S.a[d] * x;
            if (d < 2 && x > S.u[d]) inside = false;
        } else if constexpr (FAM == TQ_F_SUM_SIN) {
            acc += sin(x);
        } else if constexpr (FAM == TQ_F_SUM_EXP) {
            acc += exp(x);
        } else if constexpr (FAM == TQ_F_PROD_COS) {
            acc *= cos(x);
        } else if constexpr (FAM == TQ_F_SUM_SIN_FAST) {
            if constexpr (sizeof(T) == 4) acc += __sinf(x); else acc += sin(x);
        } else if constexpr (FAM == TQ_F_SUM_EXP_FAST) {
            if constexpr (sizeof(T) == 4) acc += __expf(x); else acc += exp(x);
        } else if constexpr (FAM == TQ_F_PROD_COS_FAST) {
            if constexpr (sizeof(T) == 4) acc *= __cosf(x); else acc *= cos(x);
        } else if constexpr (FAM == TQ_F_POLYNOMIAL) {
            T h = S.coeff[S.ncoeff - 1];
            for (int k = S.ncoeff - 2; k >= 0; --k) h = h * x + S.coeff[k];
            acc += h;
        }
    }
    __device__ __forceinline__ T finish(const FnShared<T>& S) const {
        if constexpr (FAM == TQ_F_GENZ_OSCILLATORY) return cos(S.phase + acc);
        else if constexpr (FAM == TQ_F_GENZ_CORNER_PEAK) return pow((T)1 + acc, S.expo);
        else if constexpr (FAM == TQ_F_GENZ_GAUSSIAN || FAM == TQ_F_GENZ_C0) return exp(-acc);
        else if constexpr (FAM == TQ_F_GENZ_DISCONTINUOUS) return inside ? exp(acc) : (T)0;
        else return acc;
    }
};

// ------------------------------------------------------------------ fused Monte Carlo
#ifndef TQ_MC_MIN_CTAS
#define TQ_MC_MIN_CTAS 1
#endif
#ifndef TQ_MC_ROWS
#define TQ_MC_ROWS 2  // rows a thread evaluates side by side (independent Philox chains and integrand evaluations): the kernel is
                      // issue-bound and gains from ILP, not occupancy -- measured on configs[1] (profiles/r2/exp_mc_variants.txt):
                      // 1 row 4.96e10 evals/s, 2 rows 5.50e10, 3 / 4 / 6 rows 5.29 / 5.39 / 5.15e10; capping registers for 6 / 8
                      // CTAs per SM instead: 4.60 / 4.34e10
#endif
template <int FAM, typename T>
__global__ void __launch_bounds__(256, TQ_MC_MIN_CTAS)
fused_mc_kernel(const tq_integrand P, int64_t row_begin, int64_t nrows, uint64_t seed, uint32_t call,
                double* partials, unsigned int* ticket, double* out) {
    constexpr int LANES = U01<T>::LANES;
    constexpr int R = TQ_MC_ROWS;
    __shared__ FnShared<T> S;
    __shared__ double sh[32 * 2];
    stage_integrand<T>(P, S);
    const int dim = S.dim;
    double acc[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += R * stride) {
        Integrand<FAM, T> fn[R];
        uint64_t grow[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            fn[k].init();
            grow[k] = (uint64_t)(row_begin + r + k * stride);
        }
        for (int d0 = 0; d0 < dim; d0 += LANES) {
            T u[R][LANES];
#pragma unroll
            for (int k = 0; k < R; ++k)
                philox_block<T>(seed, call, (uint32_t)grow[k], (uint32_t)(grow[k] >> 32), (uint32_t)(d0 / LANES), u[k]);
#pragma unroll
            for (int j = 0; j < LANES; ++j) {
                if (d0 + j < dim) {
                    const T sz = S.size[d0 + j], st = S.start[d0 + j];
#pragma unroll
                    for (int k = 0; k < R; ++k) fn[k].step(add_rn(mul_rn(u[k][j], sz), st), d0 + j, S);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (k == 0 || r + k * stride < nrows) {  // the extra rows of the last step are evaluated but not counted
                const double f = (double)(fn[k].finish(S) * S.scale);
                acc[0] += f;
                acc[1] += f * f;
            }
        }
    }
    grid_sum_finish<2>(acc, sh, partials, ticket, out);
}

// ------------------------------------------------------------------ fused Newton-Cotes
// Point p of the flattened grid has the multi-index (i_0 .. i_{dim-1}), dimension 0 slowest (integration_grid.py:98-99).
// A thread keeps its point as (hi, lo) = (p / B, p % B) with B = n^k <= 2^31 and advances both parts by the constant
// grid stride (an add and a conditional carry), so the loop never divides 64-bit numbers -- grids beyond 2^32 points
// (6-D Boole at n >= 41) cost the same per point as small ones; the digits come from `lo` (the k fastest dimensions) and
// `hi` by multiply-shift divisions by n (FastDiv).
struct NcWalk {
    FastDiv fd;          // division by n
    uint32_t B;          // n^k
    int k;               // digits taken from lo
    uint64_t step_hi;    // grid stride / B
    uint32_t step_lo;    // grid stride % B
};

#ifndef TQ_NC_ROWS
#define TQ_NC_ROWS 2  // grid points a thread evaluates side by side (see TQ_MC_ROWS)
#endif
template <int FAM, typename T>
__global__ void __launch_bounds__(256)
fused_nc_kernel(const tq_integrand P, const T* __restrict__ nodes, const T* __restrict__ w, uint32_t n,
                int64_t p_begin, int64_t p_end, NcWalk wk, double* partials, unsigned int* ticket, double* out, bool use_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FnShared<T> S;
    __shared__ double sh[32];
    stage_integrand<T>(P, S);
    const int dim = S.dim;
    const T* sn = nodes;
    const T* sw = w;
    if (use_smem) {
        T* a = reinterpret_cast<T*>(smem_raw);
        T* b = a + dim * n;
        for (int i = threadIdx.x; i < dim * (int)n; i += blockDim.x) { a[i] = nodes[i]; b[i] = w[i]; }
        __syncthreads();
        sn = a;
        sw = b;
    }
    double acc[1] = {0.0};
    // R points per thread side by side (p, p + stride, ...): independent digit walks and integrand evaluations give the
    // issue-bound kernel instruction-level parallelism; every copy keeps its own (hi, lo) and advances by R grid strides.
    constexpr int R = TQ_NC_ROWS;
    const int64_t r0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = p_end - p_begin;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint64_t hi[R];
    uint32_t lo[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const uint64_t pfirst = (uint64_t)(p_begin + r0 + k * stride);
        hi[k] = pfirst / wk.B;                       // once per thread
        lo[k] = (uint32_t)(pfirst - hi[k] * wk.B);
    }
    const int d_split = dim - wk.k;                    // dimensions [d_split, dim) come from lo
    for (int64_t r = r0; r < total; r += R * stride) {
        Integrand<FAM, T> fn[R];
        T wt[R];
        uint32_t q[R];
        bool wide = false;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            fn[k].init();
            wt[k] = (T)1;
            q[k] = lo[k];
            wide |= hi[k] > 0xffffffffull;
        }
        for (int d = dim - 1; d >= d_split; --d) {
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const uint32_t t = wk.fd.div(q[k]);
                const uint32_t i = q[k] - t * n;
                q[k] = t;
                wt[k] *= sw[d * n + i];
                fn[k].step(sn[d * n + i], d, S);
            }
        }
        if (!wide) {
#pragma unroll
            for (int k = 0; k < R; ++k) q[k] = (uint32_t)hi[k];
            for (int d = d_split - 1; d >= 0; --d) {
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const uint32_t t = wk.fd.div(q[k]);
                    const uint32_t i = q[k] - t * n;
                    q[k] = t;
                    wt[k] *= sw[d * n + i];
                    fn[k].step(sn[d * n + i], d, S);
                }
            }
        } else {  // more than 2^63 / n points: never in practice, kept exact
#pragma unroll
            for (int k = 0; k < R; ++k) {
                uint64_t h = hi[k];
                for (int d = d_split - 1; d >= 0; --d) {
                    const uint64_t t = h / n;
                    const uint32_t i = (uint32_t)(h - t * n);
                    h = t;
                    wt[k] *= sw[d * n + i];
                    fn[k].step(sn[d * n + i], d, S);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (k == 0 || r + k * stride < total)  // copies beyond the grid's end are evaluated (on valid table entries) but not added
                acc[0] += (double)(fn[k].finish(S) * S.scale) * (double)wt[k];
            lo[k] += wk.step_lo;
            hi[k] += wk.step_hi;
            if (lo[k] >= wk.B) { lo[k] -= wk.B; ++hi[k]; }
        }
    }
    grid_sum_finish<1>(acc, sh, partials, ticket, out);
}

// ------------------------------------------------------------------ fused VEGAS pass
// One thread per sample row.  Each CTA owns a contiguous chunk of the (cube-sorted) rows and walks it in
// tiles of FV_BLOCK rows.  Per tile, one thread per overlapped cube (<= FV_BLOCK/2 + 2 because nh >= 2) reads
// its two offsets and writes its slice index into s_cube[row] for the rows it owns, so the row -> cube lookup
// is one shared-memory byte; only the first tile of a chunk pays a global binary search, later tiles start
// from the previous tile's last cube.  Shared buffers are double-buffered: one barrier per tile.
// Bin ids of a row are parked in shared memory until jf is known, then the f^2 histogram is updated
// (L2 reductions; shared-memory privatised only for tiny maps).  Per-cube sums: rows of a cube are
// consecutive lanes, so a key-segmented warp shuffle reduction leaves one partial per (cube, warp), added to
// JF/JF2 with one L2 reduction each -- no shared-memory atomics.
// Map edges arrive packed as {x_edge[k], dx_edge[k]} pairs: one 8/16-byte gather per dimension.
constexpr int FV_BLOCK = 256;

// Multi-GPU: cubes are dealt to the ranks in blocks of 2^lb cubes, so that the hot regions VEGAS concentrates its samples
// on are spread over all ranks.  Round b (= `world` consecutive blocks) gives every rank one block; WHICH one rotates with
// the round: rank r takes position (r + skew(b)) % world.  A plain r-th-block rule would hand each rank a fixed digit of
// the cube index whenever world and the block size are powers of N_strat (8 GPUs, N_strat = 8: rank = digit 4 = a slab of
// dimension 4), and the samples VEGAS allocates per slab differ (measured: 47 % strong-scaling efficiency at 8 GPUs).
// A rank's state arrays (dh, nh, offsets, JF, JF2) hold only its own cubes, numbered locally 0..n_local.
struct CubeShard {
    uint32_t lb, rank;
    FastDiv world;
    __host__ __device__ static uint32_t skew(uint32_t b) { return b + (b >> 3) + (b >> 6) + (b >> 9) + (b >> 12) + (b >> 15) + (b >> 18); }
    __device__ __forceinline__ uint32_t global_cube(uint32_t l) const {
        if (world.d <= 1) return l;
        const uint32_t b = l >> lb;
        const uint32_t t = rank + skew(b);
        const uint32_t pos = t - world.div(t) * world.d;  // (rank + skew(b)) % world
        return ((b * world.d + pos) << lb) + (l & ((1u << lb) - 1u));
    }
    // inverse: does this rank own global cube c, and under which local id?  (local block index = round)
    __device__ __forceinline__ bool local_cube(uint32_t c, uint32_t& l) const {
        if (world.d <= 1) { l = c; return true; }
        const uint32_t blk = c >> lb;
        const uint32_t round = world.div(blk);
        const uint32_t pos = blk - round * world.d;
        const uint32_t t = rank + skew(round);
        l = (round << lb) + (c & ((1u << lb) - 1u));
        return t - world.div(t) * world.d == pos;
    }
};
#ifndef FV_MIN_CTAS
#define FV_MIN_CTAS 6
#endif
constexpr int FV_SLICE = FV_BLOCK / 2 + 4;
#ifndef TQ_FV_DIM_UNROLL
#define TQ_FV_DIM_UNROLL 1  // measured (profiles/r2/exp_variants.txt): 1 / 2 / 4 x {3..8 CTAs per SM} all within 5 %
#endif
constexpr int FV_DIM_UNROLL = TQ_FV_DIM_UNROLL;  // Philox blocks of a sample processed per unrolled step of the dimension loop

template <typename T> struct Pair2;
template <> struct Pair2<float> { using type = float2; };
template <> struct Pair2<double> { using type = double2; };

// Sum of (a, b) over the lanes that follow this one inside its segment (a run of equal keys; rows of a cube are consecutive
// lanes).  The segment's last lane comes from ONE ballot of the head flags, so the five rounds shuffle only the two values.
template <typename T>
__device__ __forceinline__ void segmented_warp_sum2(unsigned key, T& a, T& b) {
    const int lane = threadIdx.x & 31;
    const unsigned prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != key);
    const unsigned later = lane == 31 ? 0u : heads & (0xfffffffeu << lane);  // heads of the segments after this lane
    const int last = later ? __ffs(later) - 2 : 31;                         // last lane of this lane's segment
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T a2 = __shfl_down_sync(0xffffffffu, a, off);
        const T b2 = __shfl_down_sync(0xffffffffu, b, off);
        if (lane + off <= last) { a += a2; b += b2; }
    }
}

// Where the f^2 histogram of a pass goes (vegas_map.py:99-111).
enum HistMode {
    HIST_NONE = 0,     // no grid improvement
    HIST_ARRAYS = 1,   // weights[T] / counts[u64]: two L2 reductions per (sample, dim); small passes
    HIST_SMEM = 2,     // whole map privatised in shared memory (tiny maps), flushed to weights / counts
    HIST_PAIRS = 3,    // {sum jf^2, count} as an fp64 pair per bin: ONE reduction sector per (sample, dim), see below
    HIST_RECORDS = 4,  // the {weight, count} words of the bin's record (large maps)
    HIST_DEFER = 5,    // jf^2 of every row is written out; hist_sweep_kernel bins the rows band by band afterwards
};

// Sector-paired reductions.  L2 executes a reduction per 32-byte SECTOR, not per lane (measured on B200,
// scripts/microbench/lsu_rates.cu: RED.F64 to 32 scattered sectors 1.82 cyc/lane/SM; RED.F64 + RED.U64 to adjacent words
// of the same bins 3.65; ONE RED.F64 whose lane pairs hit the two words of 16 bins 1.82 per bin).  So the count lives
// next to the weight as an fp64 (exact below 2^53) and lanes 2i / 2i+1 of one instruction add {jf^2, 1.0} of sample i.
// Bin ids and jf^2 of the warp's 32 samples are parked in shared memory; two rounds of 16 samples cover the warp.
template <int FAM, typename T, bool STRAT>
__global__ void __launch_bounds__(FV_BLOCK, FV_MIN_CTAS)
fused_vegas_kernel(const tq_integrand P, const long long* __restrict__ offsets, int64_t n_cubes, CubeShard shard, FastDiv ns_div, T inv_ns,
                   T nsf, T nif,
                   int64_t row_begin, int64_t row_end, bool rows_from_offsets, int64_t rows_per_cta,
                   const void* __restrict__ edges_raw, bool records, long long ni, T* __restrict__ weights,
                   unsigned long long* __restrict__ counts, double* __restrict__ hist_pairs, T* __restrict__ jf2_rows,
                   T* __restrict__ JF, T* __restrict__ JF2, uint64_t seed, uint32_t call, int hist_mode, double* partials,
                   unsigned int* ticket, double* out) {
    constexpr int LANES = U01<T>::LANES;
    using P2 = typename Pair2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FnShared<T> S;
    __shared__ double sh[32 * 2];
    __shared__ long long s_off[2][FV_SLICE];
    __shared__ unsigned char s_cube[2][FV_BLOCK];
    stage_integrand<T>(P, S);
    const int dim = S.dim;
    double* s_jf2 = reinterpret_cast<double*>(smem_raw);                   // [FV_BLOCK] (paired reductions)
    int* s_ids = reinterpret_cast<int*>(s_jf2 + FV_BLOCK);                 // [dim][FV_BLOCK]
    T* s_w = reinterpret_cast<T*>(s_ids + dim * FV_BLOCK);                 // [dim*ni] when hist_smem
    const bool hist_smem = hist_mode == HIST_SMEM;
    unsigned int* s_c = reinterpret_cast<unsigned int*>(s_w + (hist_smem ? dim * ni : 0));
    MapRecord<T>* recs = reinterpret_cast<MapRecord<T>*>(const_cast<void*>(edges_raw));
    const bool do_hist = hist_mode != HIST_NONE && hist_mode != HIST_DEFER;
    // fp64 records keep their count as an fp64 next to the weight: the same sector-paired reduction applies
    const bool paired = hist_mode == HIST_PAIRS || (hist_mode == HIST_RECORDS && sizeof(T) == 8);
    double* pair_base = hist_mode == HIST_PAIRS ? hist_pairs : reinterpret_cast<double*>(recs) + 2;
    const int pair_stride = hist_mode == HIST_PAIRS ? 2 : 4;
    if (do_hist && hist_smem) {
        for (int i = threadIdx.x; i < dim * (int)ni; i += blockDim.x) { s_w[i] = (T)0; s_c[i] = 0u; }
    }
    __syncthreads();
    if (STRAT && rows_from_offsets) row_end = __ldg(&offsets[n_cubes]);  // sample count of the pass, never read back
    // nif = (T)ni and nsf = (T)N_strat arrive as kernel parameters: operands straight from the constant bank (computed here
    // they were re-converted from integers for every dimension to save registers)
    const int ni32 = (int)ni;                       // n_intervals < 2^31 (checked by the launcher)
    const int estride = records ? 2 : 1, eshift = records ? 1 : 0;  // a record's first half is its {x, dx} pair
    const P2* __restrict__ etab = reinterpret_cast<const P2*>(edges_raw);
    double acc[2] = {0.0, 0.0};
    int buf = 0;
    // The host sizes the grid so that one chunk per CTA covers its row estimate; the stride loop keeps the pass
    // complete when the device-side count exceeds it.
    for (int64_t r_lo = row_begin + (int64_t)blockIdx.x * rows_per_cta; r_lo < row_end;
         r_lo += (int64_t)gridDim.x * rows_per_cta) {
        const int64_t r_hi = r_lo + rows_per_cta < row_end ? r_lo + rows_per_cta : row_end;
        long long c_lo = 0;
        if (STRAT) {
            // largest c with offsets[c] <= r_lo, by a CTA-wide FV_BLOCK-ary search: every round all threads probe
            // one point each, so 10^4 cubes take two load latencies instead of fourteen dependent ones.
            long long hi = n_cubes;  // offsets[c_lo] <= r_lo < offsets[hi]
            while (hi - c_lo > 1) {
                const long long step = (hi - c_lo + FV_BLOCK - 1) / FV_BLOCK;
                const long long p = c_lo + (long long)(threadIdx.x + 1) * step;
                const int below = __syncthreads_count(p < hi && __ldg(&offsets[p]) <= r_lo);
                c_lo += below * step;
                if (c_lo + step < hi) hi = c_lo + step;
            }
        }
        for (int64_t rb = r_lo; rb < r_hi; rb += FV_BLOCK, buf ^= 1) {
            const int64_t re = rb + FV_BLOCK < r_hi ? rb + FV_BLOCK : r_hi;
            const int64_t row = rb + threadIdx.x;
            const bool active = row < re;
            if (STRAT) {
                if (threadIdx.x < FV_SLICE) {
                    const long long c = c_lo + threadIdx.x;
                    const long long lo = __ldg(&offsets[c < n_cubes ? c : n_cubes]);
                    const long long hi = __ldg(&offsets[c + 1 < n_cubes ? c + 1 : n_cubes]);
                    s_off[buf][threadIdx.x] = lo;
                    const long long a = lo > rb ? lo : rb, b = hi < re ? hi : re;
                    for (long long r = a; r < b; ++r) s_cube[buf][r - rb] = (unsigned char)threadIdx.x;
                }
                __syncthreads();
            }
            T jf = (T)0, jf2 = (T)0;
            unsigned key = 0xffffu;
            if (active) {
                uint32_t i0, i1, c = 0;
                if (STRAT) {
                    key = s_cube[buf][threadIdx.x];
                    i0 = shard.global_cube((uint32_t)(c_lo + key));  // Philox key and digits come from the GLOBAL cube id
                    i1 = (uint32_t)(row - s_off[buf][key]);
                    c = i0;
                } else {
                    i0 = (uint32_t)(uint64_t)row;
                    i1 = (uint32_t)((uint64_t)row >> 32);
                }
                Integrand<FAM, T> fn;
                fn.init();
                T jac = (T)1;
#pragma unroll FV_DIM_UNROLL
                for (int d0 = 0; d0 < dim; d0 += LANES) {  // unrolled: the edge gathers of several Philox blocks are in flight together
                    T u[LANES];
                    philox_block<T>(seed, call, i0, i1, (uint32_t)(d0 / LANES), u);
    #pragma unroll
                    for (int j = 0; j < LANES; ++j) {
                        const int d = d0 + j;
                        if (d < dim) {
                            T y;
                            if (STRAT) {
                                const uint32_t q = ns_div.div(c);
                                const uint32_t p = c - q * ns_div.d;
                                c = q;
                                y = div_by_const(add_rn((T)p, u[j]), nsf, inv_ns);
                                if (y >= (T)1) y = (T)0.999999;
                            } else {
                                y = mul_rn(u[j], (T)0.999999);
                            }
                            const T t = mul_rn(y, nif);
                            const T fl = floor(t);
                            int k = (int)fl;  // 0 <= t < 2^31
                            k = max(0, min(k, ni32 - 1));
                            const T o = sub_rn(t, fl);
                            const P2 e = __ldg(etab + (int64_t)d * ni * estride + (k << eshift));  // uniform table base + a 32-bit offset
                            const T x = add_rn(e.x, mul_rn(e.y, o));
                            jac = mul_rn(jac, mul_rn(nif, e.y));
                            s_ids[d * FV_BLOCK + threadIdx.x] = k;
                            fn.step(add_rn(mul_rn(x, S.size[d]), S.start[d]), d, S);
                        }
                    }
                }
                const T f = mul_rn(fn.finish(S), S.scale);
                jf = mul_rn(f, jac);
                jf2 = mul_rn(jf, jf);
                if (hist_mode == HIST_DEFER) jf2_rows[row - row_begin] = jf2;
                if (do_hist && !paired) {
                    if (hist_smem) {
                        for (int d = 0; d < dim; ++d) {
                            const int b = d * (int)ni + s_ids[d * FV_BLOCK + threadIdx.x];
                            atomicAdd(&s_w[b], jf2);
                            atomicAdd(&s_c[b], 1u);
                        }
                    } else if (records) {
                        for (int d = 0; d < dim; ++d) {
                            MapRecord<T>* r = &recs[(int64_t)d * ni + s_ids[d * FV_BLOCK + threadIdx.x]];
                            atomicAdd(&r->w, jf2);
                            atomicAdd(&r->c, (decltype(r->c))1);
                        }
                    } else {
                        for (int d = 0; d < dim; ++d) {
                            const int64_t b = (int64_t)d * ni + s_ids[d * FV_BLOCK + threadIdx.x];
                            atomicAdd(&weights[b], jf2);
                            atomicAdd(&counts[b], 1ull);
                        }
                    }
                }
            }
            if (paired) {  // warp-convergent: every lane serves half a sample of its warp
                s_jf2[threadIdx.x] = (double)jf2;
                if (!active) s_ids[threadIdx.x] = -1;  // dimension 0 marks the row: inactive rows have no bins at all
                __syncwarp();
                {
                    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31, word = lane & 1;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int src = wbase + (lane >> 1) + 16 * h;
                        if (s_ids[src] >= 0) {
                            const double v = word ? 1.0 : s_jf2[src];
                            for (int d = 0; d < dim; ++d)
                                atomicAdd(pair_base + ((int64_t)d * ni + s_ids[d * FV_BLOCK + src]) * pair_stride + word, v);
                        }
                    }
                }
                __syncwarp();
            }
            if (STRAT) {
                const unsigned prev = __shfl_up_sync(0xffffffffu, key, 1);
                T a = jf, b = jf2;
                segmented_warp_sum2<T>(key, a, b);
                if (active && ((threadIdx.x & 31) == 0 || prev != key)) {
                    atomicAdd(&JF[c_lo + key], a);
                    atomicAdd(&JF2[c_lo + key], b);
                }
                c_lo += s_cube[buf][(int)(re - rb) - 1];  // cube of the tile's last row: where the next tile starts
            } else {
                acc[0] += (double)jf;
                acc[1] += (double)jf2;
            }
        }
    }
    if (do_hist && hist_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < dim * (int)ni; i += blockDim.x) {
            const unsigned int cc = s_c[i];
            if (cc) {
                atomicAdd(&weights[i], s_w[i]);
                atomicAdd(&counts[i], (unsigned long long)cc);
            }
        }
    }
    if (!STRAT) grid_sum_finish<2>(acc, sh, partials, ticket, out);
}

// ------------------------------------------------------------------ fused VEGAS pass, band-privatised (tile) version
// Shared-memory privatised histograms for maps that stay in L2 (north_star).  Cube c = sum_d p_d N_strat^d
// has its samples of dimension d inside the band of ~Ni / N_strat bins selected by digit p_d.  A TILE is an aligned
// block of N_strat^g consecutive cubes: the g low digits run over everything, the digits of the dim - g HIGH dimensions are
// fixed, so for those dimensions one CTA stages the band's {x, dx} edges in shared memory (the gather becomes a 16-byte
// shared-memory read: 0.3 instead of ~1 L1TEX cycles per lane) and keeps a private {sum jf^2, count} band that takes
// shared-memory atomics (0.65 cycles per sample and dimension instead of 1.8 for an L2 reduction sector) and is flushed
// once per tile with sector-paired reductions.  The low dimensions go through L2 exactly like in fused_vegas_kernel.
// Tiles are handed out in order by a global counter; one 1024-thread CTA per SM (the bands need up to ~200 KB).
constexpr int FT_BLOCK = 1024;
constexpr size_t FT_SMEM_CAP = 210 * 1024;  // dynamic shared memory of the tile kernel (227 KB per CTA minus ~14 KB of static arrays)
constexpr int FT_SLICE = FT_BLOCK / 2 + 4;

template <int FAM, typename T>
__global__ void __launch_bounds__(FT_BLOCK, 1)
fused_vegas_tile_kernel(const tq_integrand P, const long long* __restrict__ offsets, uint32_t n_cubes, CubeShard shard, FastDiv ns_div,
                        T inv_ns, T nsf, T nif, const typename Pair2<T>::type* __restrict__ edges, long long ni, double* __restrict__ hist,
                        T* __restrict__ JF, T* __restrict__ JF2, uint64_t seed, uint32_t call, int g, uint32_t tile_cubes,
                        FastDiv tile_div, int band_w, int ne, bool hist_smem, unsigned int* next_tile) {
    constexpr int LANES = U01<T>::LANES;
    using P2 = typename Pair2<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FnShared<T> S;
    __shared__ long long s_off[2][FT_SLICE];
    __shared__ unsigned short s_cube[2][FT_BLOCK];
    __shared__ int s_blo[TQ_MAX_DIM];
    __shared__ unsigned int s_tile;
    stage_integrand<T>(P, S);
    const int dim = S.dim;
    const int ns = dim - g;  // band (high) dimensions
    // dynamic shared memory: edges bands | weight bands | jf^2 of the rows in flight | count bands | low-dim bin ids | band ids
    // Every band dimension has a private {weight, count} band (a shared-memory atomic replaces an L2 reduction sector: the
    // larger saving); the first `ne` of them also have their {x, dx} edges staged (a shared-memory read replaces a gather).
    // hist_smem == false (fp64: a shared-memory fp64 atomicAdd is a compare-and-swap loop): only the edges are staged, every
    // dimension's histogram goes through the sector-paired L2 reductions and s_ids holds the bin ids of ALL dimensions.
    const size_t nband = hist_smem ? (size_t)ns * band_w : 0;
    const int nid = hist_smem ? g : dim;  // dimensions binned through L2
    P2* s_edge = reinterpret_cast<P2*>(smem_raw);                                  // [ne][band_w]
    T* s_hw = reinterpret_cast<T*>(s_edge + (size_t)ne * band_w);                 // [ns][band_w]
    double* s_jf2 = reinterpret_cast<double*>(s_hw + ((nband + 1) & ~(size_t)1)); // [FT_BLOCK]
    unsigned int* s_hc = reinterpret_cast<unsigned int*>(s_jf2 + FT_BLOCK);       // [ns][band_w]
    int* s_ids = reinterpret_cast<int*>(s_hc + nband);                            // [nid][FT_BLOCK]
    unsigned short* s_bid = reinterpret_cast<unsigned short*>(s_ids + (size_t)nid * FT_BLOCK);  // [ns][FT_BLOCK] (hist_smem)
    const int ni32 = (int)ni;  // n_intervals <= 2^20 on this path; nif = (T)ni, nsf = (T)N_strat are kernel parameters
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31, word = lane & 1;
    int buf = 0;
    for (;;) {
        __syncthreads();  // previous tile flushed
        if (threadIdx.x == 0) s_tile = atomicAdd(next_tile, 1u);
        __syncthreads();
        const uint64_t c_first64 = (uint64_t)s_tile * tile_cubes;
        if (c_first64 >= n_cubes) break;
        const uint32_t c_first = (uint32_t)c_first64;
        const uint32_t c_end = c_first + tile_cubes < n_cubes ? c_first + tile_cubes : n_cubes;
        const int64_t r_lo = __ldg(&offsets[c_first]), r_hi = __ldg(&offsets[c_end]);
        if (threadIdx.x < ns) {  // the band of every high dimension: digit (g + sd) of the tile's GLOBAL cube ids
            uint32_t q = tile_div.div(shard.global_cube(c_first));
            uint32_t p = 0;
            for (int sd = 0; sd <= (int)threadIdx.x; ++sd) {
                const uint32_t qq = ns_div.div(q);
                p = q - qq * ns_div.d;
                q = qq;
            }
            long long lo = ((long long)p * ni) / (long long)ns_div.d - 1;  // one bin of slack on either side (rounding of y * Ni)
            s_blo[threadIdx.x] = (int)(lo < 0 ? 0 : lo);
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < ns * band_w; idx += FT_BLOCK) {
            const int sd = idx / band_w, wdx = idx - sd * band_w;
            const long long k = (long long)s_blo[sd] + wdx;
            if (sd < ne) {
                P2 e;
                e.x = (T)0;
                e.y = (T)0;
                if (k < ni) e = __ldg(&edges[(int64_t)(g + sd) * ni + k]);
                s_edge[idx] = e;
            }
            if (hist_smem) {
                s_hw[idx] = (T)0;
                s_hc[idx] = 0u;
            }
        }
        __syncthreads();
        long long c_lo = c_first;
        for (int64_t rb = r_lo; rb < r_hi; rb += FT_BLOCK, buf ^= 1) {
            const int64_t re = rb + FT_BLOCK < r_hi ? rb + FT_BLOCK : r_hi;
            const int64_t row = rb + threadIdx.x;
            const bool active = row < re;
            if (threadIdx.x < FT_SLICE) {
                const long long c = c_lo + threadIdx.x;
                const long long lo = __ldg(&offsets[c < n_cubes ? c : n_cubes]);
                const long long hi = __ldg(&offsets[c + 1 < n_cubes ? c + 1 : n_cubes]);
                s_off[buf][threadIdx.x] = lo;
                const long long a = lo > rb ? lo : rb, b = hi < re ? hi : re;
                for (long long r = a; r < b; ++r) s_cube[buf][r - rb] = (unsigned short)threadIdx.x;
            }
            __syncthreads();
            T jf = (T)0, jf2 = (T)0;
            unsigned key = 0xffffu;
            if (active) {
                key = s_cube[buf][threadIdx.x];
                const uint32_t i0 = shard.global_cube((uint32_t)(c_lo + key));
                const uint32_t i1 = (uint32_t)(row - s_off[buf][key]);
                uint32_t c = i0;
                Integrand<FAM, T> fn;
                fn.init();
                T jac = (T)1;
                for (int d0 = 0; d0 < dim; d0 += LANES) {
                    T u[LANES];
                    philox_block<T>(seed, call, i0, i1, (uint32_t)(d0 / LANES), u);
#pragma unroll
                    for (int j = 0; j < LANES; ++j) {
                        const int d = d0 + j;
                        if (d < dim) {
                            const uint32_t q = ns_div.div(c);
                            const uint32_t p = c - q * ns_div.d;
                            c = q;
                            T y = div_by_const(add_rn((T)p, u[j]), nsf, inv_ns);
                            if (y >= (T)1) y = (T)0.999999;
                            const T t = mul_rn(y, nif);
                            const T fl = floor(t);
                            int k = (int)fl;
                            k = max(0, min(k, ni32 - 1));
                            const T o = sub_rn(t, fl);
                            P2 e;
                            if (d < g) {  // uniform over the CTA
                                e = __ldg(&edges[(int64_t)d * ni + k]);
                                s_ids[d * FT_BLOCK + threadIdx.x] = k;
                            } else {
                                const int sd = d - g;
                                const unsigned loc = (unsigned)(k - s_blo[sd]);
                                const bool in_band = loc < (unsigned)band_w;
                                // outside the staged band: cannot happen with the slack for Ni <= 2^20; kept exact anyway
                                e = (in_band && sd < ne) ? s_edge[sd * band_w + (int)loc] : __ldg(&edges[(int64_t)d * ni + k]);
                                if (hist_smem) s_bid[sd * FT_BLOCK + threadIdx.x] = in_band ? (unsigned short)loc : (unsigned short)0xffffu;
                                else s_ids[d * FT_BLOCK + threadIdx.x] = k;
                            }
                            const T x = add_rn(e.x, mul_rn(e.y, o));
                            jac = mul_rn(jac, mul_rn(nif, e.y));
                            fn.step(add_rn(mul_rn(x, S.size[d]), S.start[d]), d, S);
                        }
                    }
                }
                const T f = mul_rn(fn.finish(S), S.scale);
                jf = mul_rn(f, jac);
                jf2 = mul_rn(jf, jf);
                for (int sd = 0; sd < ns && hist_smem; ++sd) {  // band dimensions: private shared-memory histogram
                    const unsigned loc = s_bid[sd * FT_BLOCK + threadIdx.x];
                    if (loc != 0xffffu) {
                        atomicAdd(&s_hw[sd * band_w + loc], jf2);
                        atomicAdd(&s_hc[sd * band_w + loc], 1u);
                    } else {  // recompute the bin of dimension g + sd and update the global table directly
                        const int d = g + sd;
                        T u[LANES];
                        philox_block<T>(seed, call, i0, i1, (uint32_t)(d / LANES), u);
                        T ud = u[0];
#pragma unroll
                        for (int l = 1; l < LANES; ++l)
                            if (d % LANES == l) ud = u[l];
                        uint32_t q = i0, p = 0;
                        for (int dd = 0; dd <= d; ++dd) {
                            const uint32_t qq = ns_div.div(q);
                            p = q - qq * ns_div.d;
                            q = qq;
                        }
                        T y = div_by_const(add_rn((T)p, ud), nsf, inv_ns);
                        if (y >= (T)1) y = (T)0.999999;
                        long long k = (long long)floor(mul_rn(y, nif));
                        k = k < 0 ? 0 : (k >= ni ? ni - 1 : k);
                        atomicAdd(hist + ((int64_t)d * ni + k) * 2, (double)jf2);
                        atomicAdd(hist + ((int64_t)d * ni + k) * 2 + 1, 1.0);
                    }
                }
            }
            // low dimensions: sector-paired L2 reductions (lanes 2i / 2i+1 serve sample i of a 16-sample half)
            s_jf2[threadIdx.x] = active ? (double)jf2 : -1.0;
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int src = wbase + (lane >> 1) + 16 * h;
                const double sv = s_jf2[src];
                if (sv >= 0.0 || sv != sv) {  // active source row (NaN weights are kept, like the direct path keeps them)
                    const double v = word ? 1.0 : sv;
                    for (int d = 0; d < nid; ++d) atomicAdd(hist + ((int64_t)d * ni + s_ids[d * FT_BLOCK + src]) * 2 + word, v);
                }
            }
            __syncwarp();
            {
                const unsigned prev = __shfl_up_sync(0xffffffffu, key, 1);
                T a = jf, b = jf2;
                segmented_warp_sum2<T>(key, a, b);
                if (active && (lane == 0 || prev != key)) {
                    atomicAdd(&JF[c_lo + key], a);
                    atomicAdd(&JF2[c_lo + key], b);
                }
                c_lo += s_cube[buf][(int)(re - rb) - 1];
            }
        }
        __syncthreads();  // every row of the tile is binned
        for (int idx = threadIdx.x; idx < ns * band_w * 2 && hist_smem; idx += FT_BLOCK) {  // flush: lanes 2i / 2i+1 -> {sum, count} of bin i
            const int b = idx >> 1, wd = idx & 1;
            const unsigned int cnt = s_hc[b];
            if (cnt) {
                const int sd = b / band_w, wdx = b - sd * band_w;
                atomicAdd(hist + ((int64_t)(g + sd) * ni + s_blo[sd] + wdx) * 2 + wd, wd ? (double)cnt : (double)s_hw[b]);
            }
        }
    }
}

// {x_edges[d,k], dx_edges[d,k]} -> packed pairs [dim, Ni]
template <typename T>
__global__ void __launch_bounds__(256)
pack_edges_kernel(const T* __restrict__ xe, const T* __restrict__ dxe, typename Pair2<T>::type* __restrict__ out,
                  int dim, long long ni) {
    const int64_t total = (int64_t)dim * ni;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = i / ni, k = i - d * ni;
        typename Pair2<T>::type e;
        e.x = xe[d * (ni + 1) + k];
        e.y = dxe[i];
        out[i] = e;
    }
}

// records[d,k] = {x_edges[d,k], dx_edges[d,k], 0, 0}
template <typename T>
__global__ void __launch_bounds__(256)
pack_records_kernel(const T* __restrict__ xe, const T* __restrict__ dxe, MapRecord<T>* __restrict__ out, int dim, long long ni) {
    const int64_t total = (int64_t)dim * ni;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = i / ni, k = i - d * ni;
        MapRecord<T> r;
        r.x = xe[d * (ni + 1) + k];
        r.dx = dxe[i];
        r.w = (T)0;
        r.c = 0;
        out[i] = r;
    }
}

// weights += records.w, counts += records.c, and the record fields go back to zero
template <typename T>
__global__ void __launch_bounds__(256)
unpack_records_kernel(MapRecord<T>* __restrict__ recs, T* __restrict__ weights, long long* __restrict__ counts, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const MapRecord<T> r = recs[i];
        if (r.c != 0) {
            weights[i] = add_rn(weights[i], r.w);
            counts[i] += (long long)r.c;
            MapRecord<T> z = r;
            z.w = (T)0;
            z.c = 0;
            recs[i] = z;
        }
    }
}

// weights += (T)hist.w, counts += (int64)hist.c, and the pair goes back to zero (fp64 pair table of HIST_PAIRS)
template <typename T>
__global__ void __launch_bounds__(256)
unpack_hist_kernel(double2* __restrict__ hist, T* __restrict__ weights, long long* __restrict__ counts, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 h = hist[i];
        if (h.y != 0.0) {
            weights[i] = (T)((double)weights[i] + h.x);
            counts[i] += (long long)h.y;
            hist[i] = make_double2(0.0, 0.0);
        }
    }
}

// ---- maps beyond L2: the histogram of a stratified pass, band by band -------------------------------------------------
// A sample of cube c falls, in dimension d, into the band of Ni / N_strat bins selected by digit d of c
// (y_d = (digit + u) / N_strat, vegas_stratification.py:140-165).  With the reference's map size (Ni = 1e7 per dimension)
// the histogram table is far beyond L2 and every reduction of the fused pass is a random read-modify-write in HBM.  Instead
// the pass only stores jf^2 per row (HIST_DEFER) and this kernel re-bins the rows afterwards, one GROUP of g <= LANES
// consecutive dimensions (= one Philox block) per launch, walking the cubes in the order of the group's digits: all cubes
// that are in flight at a time share the group's g bands (g * Ni / N_strat * 16 bytes, 40 MB for configs[3]), so the
// reductions hit L2 and every table line goes to HBM once per band.  The uniforms are regenerated from the cube-keyed
// Philox stream (same block, same lanes, same arithmetic as the pass: identical bins), jf^2 is read back (8 bytes / row).
// One warp handles 32 cubes; per step every lane has at most one sample, binned with the sector-paired reductions.
constexpr uint32_t SWEEP_CHUNK = 1024;  // cubes per scheduling unit

template <typename T>
__global__ void __launch_bounds__(256)
hist_sweep_kernel(const long long* __restrict__ offsets, uint32_t n_cubes, CubeShard shard, const T* __restrict__ jf2_rows,
                  double* __restrict__ hist, long long ni, int d0, int g, FastDiv ns_div, FastDiv low_div, uint32_t pow_g,
                  FastDiv cpc_div, T inv_ns, uint64_t seed, uint32_t call, int lane0, unsigned int* next_chunk) {
    constexpr int LANES = U01<T>::LANES;
    __shared__ unsigned int s_chunk;
    const int lane = threadIdx.x & 31, word = lane & 1;
    const T nif = (T)ni, nsf = (T)ns_div.d;
    const uint32_t cubes_per_combo = cpc_div.d, pow_low = low_div.d;
    // Chunks of SWEEP_CHUNK cubes are handed out IN ORDER by a global counter: the cubes in flight then always form one
    // compact window of the band-major order, whatever the spread of the per-cube work.  (A static grid-stride assignment
    // lets the CTAs drift apart once nh is adapted -- measured: 16 HBM round trips per table line instead of one.)
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(next_chunk, 1u);
        __syncthreads();
        const uint64_t chunk0 = (uint64_t)s_chunk * SWEEP_CHUNK;
        if (chunk0 >= n_cubes) break;
    for (uint32_t base = (uint32_t)chunk0 + (threadIdx.x - lane); base < n_cubes && base < chunk0 + SWEEP_CHUNK; base += blockDim.x) {
        const uint32_t j = base + lane;  // position in the band-major order
        uint32_t c = 0, combo = 0;
        long long row0 = 0;
        int nh = 0;
        if (j < n_cubes) {
            combo = cpc_div.div(j);                       // the group's digits (dimension d0 fastest)
            const uint32_t r = j - combo * cubes_per_combo;
            const uint32_t high = low_div.div(r), low = r - high * pow_low;
            c = (high * pow_g + combo) * pow_low + low;   // the (GLOBAL) cube with those digits at positions d0 .. d0+g-1
            // multi-GPU: the sweep walks the global band-major order and bins only the cubes this rank owns; `offsets`,
            // `jf2_rows` are the rank's local arrays (a skipped cube costs the index arithmetic above and nothing else)
            uint32_t l = c;
            const bool mine = shard.local_cube(c, l);
            if (mine) {
                row0 = __ldcs(&offsets[l]);
                nh = (int)(__ldcs(&offsets[l + 1]) - row0);
            }
        }
        // The warp's 32 cubes hold `total` samples; they are flattened so that every step gives each lane one sample
        // whatever the spread of nh (after adaptation nh varies by an order of magnitude between neighbouring cubes).
        int incl = nh;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - nh;
        for (int s0 = 0; s0 < total; s0 += 32) {
            const int sidx = s0 + lane;
            const bool act = sidx < total;
            // owner = first lane whose inclusive count exceeds sidx (5-step search over the lanes' counts)
            int lo = 0, hi = 31;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int mid = (lo + hi) >> 1;
                const int vmid = __shfl_sync(0xffffffffu, incl, mid);
                if (vmid <= sidx) lo = mid + 1; else hi = mid;
            }
            const int owner = lo;
            const uint32_t oc = __shfl_sync(0xffffffffu, c, owner);
            const uint32_t ocombo = __shfl_sync(0xffffffffu, combo, owner);
            const long long orow0 = __shfl_sync(0xffffffffu, row0, owner);
            const int oi = sidx - __shfl_sync(0xffffffffu, excl, owner);
            double v = 0.0;
            int k[LANES];
#pragma unroll
            for (int t = 0; t < LANES; ++t) k[t] = 0;
            if (act) {
                T u[LANES];
                philox_block<T>(seed, call, oc, (uint32_t)oi, (uint32_t)(d0 / LANES), u);
                v = (double)__ldcs(&jf2_rows[orow0 + oi]);
                uint32_t q = ocombo;
#pragma unroll
                for (int t = 0; t < LANES; ++t) {
                    if (t < g) {
                        const uint32_t qq = ns_div.div(q);
                        const uint32_t p = q - qq * ns_div.d;
                        q = qq;
                        T ut = u[0];  // lane lane0 + t of the block (dimension d0 + t), selected without dynamic indexing
#pragma unroll
                        for (int l = 1; l < LANES; ++l)
                            if (lane0 + t == l) ut = u[l];
                        T y = div_by_const(add_rn((T)p, ut), nsf, inv_ns);
                        if (y >= (T)1) y = (T)0.999999;
                        const T fl = floor(mul_rn(y, nif));
                        long long kk = (long long)fl;
                        kk = kk < 0 ? 0 : (kk >= ni ? ni - 1 : kk);
                        k[t] = (int)kk;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // lanes 2i / 2i+1 serve sample i of a 16-sample half: {sum jf^2, count} of one bin
                const int src = (lane >> 1) + 16 * h;
                const bool a = __shfl_sync(0xffffffffu, (int)act, src) != 0;
                const double sv = __shfl_sync(0xffffffffu, v, src);  // every lane takes part in the shuffle
                const double val = word ? 1.0 : sv;
#pragma unroll
                for (int t = 0; t < LANES; ++t) {
                    const int kk = __shfl_sync(0xffffffffu, k[t], src);
                    if (t < g && a) atomicAdd(hist + ((int64_t)(d0 + t) * ni + kk) * 2 + word, val);
                }
            }
        }
    }
    }
}

// pairs[i] += {record.w, record.c}; the record's histogram fields go back to zero
template <typename T>
__global__ void __launch_bounds__(256)
records_to_pairs_kernel(MapRecord<T>* __restrict__ recs, double2* __restrict__ pairs, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const MapRecord<T> r = recs[i];
        if (r.c != 0) {
            double2 p = pairs[i];
            p.x += (double)r.w;
            p.y += (double)r.c;
            pairs[i] = p;
            MapRecord<T> z = r;
            z.w = (T)0;
            z.c = 0;
            recs[i] = z;
        }
    }
}

int records_to_pairs_launch(void* records, double* pairs, int32_t dim, int64_t ni, int32_t dtype, void* stream) {
    const int64_t total = (int64_t)dim * ni;
    const int grid = grid_for(total, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        records_to_pairs_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((MapRecord<T>*)records, (double2*)pairs, total);
    });
    return check_launch("records_to_pairs_kernel");
}

#define TQ_DISPATCH_FAMILY(fam, ...)                                                                  \
    switch (fam) {                                                                                    \
        case TQ_F_GENZ_OSCILLATORY: { constexpr int FAM = TQ_F_GENZ_OSCILLATORY; __VA_ARGS__; break; }      \
        case TQ_F_GENZ_PRODUCT_PEAK: { constexpr int FAM = TQ_F_GENZ_PRODUCT_PEAK; __VA_ARGS__; break; }    \
        case TQ_F_GENZ_CORNER_PEAK: { constexpr int FAM = TQ_F_GENZ_CORNER_PEAK; __VA_ARGS__; break; }      \
        case TQ_F_GENZ_GAUSSIAN: { constexpr int FAM = TQ_F_GENZ_GAUSSIAN; __VA_ARGS__; break; }            \
        case TQ_F_GENZ_C0: { constexpr int FAM = TQ_F_GENZ_C0; __VA_ARGS__; break; }                        \
        case TQ_F_GENZ_DISCONTINUOUS: { constexpr int FAM = TQ_F_GENZ_DISCONTINUOUS; __VA_ARGS__; break; }  \
        case TQ_F_SUM_SIN: { constexpr int FAM = TQ_F_SUM_SIN; __VA_ARGS__; break; }                        \
        case TQ_F_SUM_EXP: { constexpr int FAM = TQ_F_SUM_EXP; __VA_ARGS__; break; }                        \
        case TQ_F_PROD_COS: { constexpr int FAM = TQ_F_PROD_COS; __VA_ARGS__; break; }                      \
        case TQ_F_POLYNOMIAL: { constexpr int FAM = TQ_F_POLYNOMIAL; __VA_ARGS__; break; }                  \
        case TQ_F_SUM_SIN_FAST: { constexpr int FAM = TQ_F_SUM_SIN_FAST; __VA_ARGS__; break; }              \
        case TQ_F_SUM_EXP_FAST: { constexpr int FAM = TQ_F_SUM_EXP_FAST; __VA_ARGS__; break; }              \
        case TQ_F_PROD_COS_FAST: { constexpr int FAM = TQ_F_PROD_COS_FAST; __VA_ARGS__; break; }            \
        default: tq::set_error("unknown integrand family %d", (int)(fam)); return TQ_ERR_INVALID_ARGUMENT;  \
    }

static int check_integrand(const char* who, const tq_integrand* fn) {
    TQ_REQUIRE(fn != nullptr, "%s: integrand is NULL", who);
    TQ_REQUIRE(fn->dim >= 1 && fn->dim <= TQ_MAX_DIM, "%s: integrand dim %d out of range (max %d)", who, fn->dim, TQ_MAX_DIM);
    TQ_REQUIRE(fn->family >= 0 && fn->family < TQ_F_COUNT, "%s: unknown integrand family %d", who, fn->family);
    TQ_REQUIRE(fn->family != TQ_F_POLYNOMIAL || (fn->ncoeff >= 1 && fn->ncoeff <= 8), "%s: polynomial needs 1..8 coefficients", who);
    return TQ_OK;
}

}  // namespace tq

using namespace tq;

extern "C" {

int tq_fused_mc(const tq_integrand* fn_host, int32_t dtype, int64_t row_begin, int64_t row_end,
                uint64_t seed, uint32_t call_idx, double* out_f64, void* ws, size_t ws_bytes,
                void* stream) {
    int rc = check_integrand("tq_fused_mc", fn_host);
    if (rc) return rc;
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_fused_mc: bad row range");
    Workspace w(ws, ws_bytes);
    unsigned int* ticket = w.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    const int64_t nrows = row_end - row_begin;
    const int grid = grid_for((nrows + TQ_MC_ROWS - 1) / TQ_MC_ROWS, 256, 8);  // a thread evaluates TQ_MC_ROWS rows per step
    double* partials = w.take<double>((size_t)grid * 2);
    if (!ticket || !partials) { set_error("tq_fused_mc: workspace too small"); return TQ_ERR_WORKSPACE; }
    cudaStream_t st = as_stream(stream);
    TQ_DISPATCH_DTYPE(dtype, {
        TQ_DISPATCH_FAMILY(fn_host->family, {
            fused_mc_kernel<FAM, T><<<TQ_GRID(grid), 256, 0, st>>>(*fn_host, row_begin, nrows, seed, call_idx, partials, ticket, out_f64);
        });
    });
    return check_launch("fused_mc_kernel");
}

int tq_fused_nc(const tq_integrand* fn_host, const void* nodes, const void* w, int32_t n, int32_t dtype,
                int64_t p_begin, int64_t p_end, double* out_f64, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_integrand("tq_fused_nc", fn_host);
    if (rc) return rc;
    TQ_REQUIRE(n >= 1 && p_begin >= 0 && p_end >= p_begin, "tq_fused_nc: bad grid arguments");
    Workspace wk(ws, ws_bytes);
    unsigned int* ticket = wk.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    const int grid = grid_for((p_end - p_begin + TQ_NC_ROWS - 1) / TQ_NC_ROWS, 256, 8);  // TQ_NC_ROWS points per thread and step
    double* partials = wk.take<double>((size_t)grid);
    if (!ticket || !partials) { set_error("tq_fused_nc: workspace too small"); return TQ_ERR_WORKSPACE; }
    cudaStream_t st = as_stream(stream);
    const int dim = fn_host->dim;
    NcWalk walk;
    walk.fd.set((uint32_t)n);
    walk.B = 1;
    walk.k = 0;
    while (walk.k < dim && (uint64_t)walk.B * (uint64_t)n <= (1ull << 31)) { walk.B *= (uint32_t)n; ++walk.k; }
    if (walk.k == 0) { walk.B = (uint32_t)n; walk.k = 1; }  // n > 2^31 is refused by the int32 argument; n = 1 lands here
    const uint64_t stride = (uint64_t)grid * 256 * TQ_NC_ROWS;  // a thread's copies advance by TQ_NC_ROWS grid strides per step
    walk.step_hi = stride / walk.B;
    walk.step_lo = (uint32_t)(stride % walk.B);
    TQ_DISPATCH_DTYPE(dtype, {
        const size_t table = (size_t)2 * dim * n * sizeof(T);
        const bool use_smem = table <= 96 * 1024;
        TQ_DISPATCH_FAMILY(fn_host->family, {
            cudaFuncSetAttribute(fused_nc_kernel<FAM, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            fused_nc_kernel<FAM, T><<<TQ_GRID(grid), 256, use_smem ? table : 0, st>>>(*fn_host, (const T*)nodes, (const T*)w, (uint32_t)n,
                                                                           p_begin, p_end, walk, partials, ticket, out_f64, use_smem);
        });
    });
    return check_launch("fused_nc_kernel");
}

int tq_vegas_map_pack_edges(const void* x_edges, const void* dx_edges, void* edges_packed, int32_t dim,
                            int64_t n_intervals, int32_t dtype, void* stream) {
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1, "tq_vegas_map_pack_edges: bad shape");
    const int grid = grid_for((int64_t)dim * n_intervals, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        pack_edges_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((const T*)x_edges, (const T*)dx_edges,
                                                                 (typename Pair2<T>::type*)edges_packed, dim, n_intervals);
    });
    return check_launch("pack_edges_kernel");
}

size_t tq_vegas_map_records_bytes(int32_t dim, int64_t n_intervals, int32_t dtype) {
    return (size_t)dim * (size_t)n_intervals * (dtype == TQ_F64 ? sizeof(MapRecord<double>) : sizeof(MapRecord<float>));
}

int tq_vegas_map_pack_records(const void* x_edges, const void* dx_edges, void* records, int32_t dim,
                              int64_t n_intervals, int32_t dtype, void* stream) {
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1, "tq_vegas_map_pack_records: bad shape");
    const int grid = grid_for((int64_t)dim * n_intervals, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        pack_records_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((const T*)x_edges, (const T*)dx_edges,
                                                                   (MapRecord<T>*)records, dim, n_intervals);
    });
    return check_launch("pack_records_kernel");
}

int tq_vegas_map_unpack_records(void* records, void* weights, int64_t* counts, int32_t dim, int64_t n_intervals,
                                int32_t dtype, void* stream) {
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1, "tq_vegas_map_unpack_records: bad shape");
    const int64_t total = (int64_t)dim * n_intervals;
    const int grid = grid_for(total, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        unpack_records_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((MapRecord<T>*)records, (T*)weights, (long long*)counts, total);
    });
    return check_launch("unpack_records_kernel");
}

int tq_vegas_map_unpack_hist(void* hist_pairs, void* weights, int64_t* counts, int32_t dim, int64_t n_intervals,
                             int32_t dtype, void* stream) {
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1 && hist_pairs && weights && counts, "tq_vegas_map_unpack_hist: bad arguments");
    const int64_t total = (int64_t)dim * n_intervals;
    const int grid = grid_for(total, 256, 8);
    TQ_DISPATCH_DTYPE(dtype, {
        unpack_hist_kernel<T><<<TQ_GRID(grid), 256, 0, as_stream(stream)>>>((double2*)hist_pairs, (T*)weights, (long long*)counts, total);
    });
    return check_launch("unpack_hist_kernel");
}

int tq_fused_vegas(const tq_integrand* fn_host, int32_t dtype, const int64_t* offsets, int64_t n_cubes,
                   int32_t n_strat, int64_t row_begin, int64_t row_end, const void* edges_packed,
                   int32_t edges_layout, int64_t n_intervals, void* weights, int64_t* counts, void* hist_pairs,
                   void* JF, void* JF2, uint64_t seed, uint32_t call_idx, double* out_f64, void* ws,
                   size_t ws_bytes, void* stream) {
    return tq_fused_vegas_sharded(fn_host, dtype, offsets, n_cubes, n_strat, row_begin, row_end, edges_packed, edges_layout,
                                  n_intervals, weights, counts, hist_pairs, JF, JF2, seed, call_idx, 0, 0, 1, out_f64, ws, ws_bytes,
                                  stream);
}

int tq_fused_vegas_sharded(const tq_integrand* fn_host, int32_t dtype, const int64_t* offsets, int64_t n_cubes,
                           int32_t n_strat, int64_t row_begin, int64_t row_end, const void* edges_packed,
                           int32_t edges_layout, int64_t n_intervals, void* weights, int64_t* counts, void* hist_pairs,
                           void* JF, void* JF2, uint64_t seed, uint32_t call_idx, int32_t cube_block_log2, int32_t rank,
                           int32_t world, double* out_f64, void* ws, size_t ws_bytes, void* stream) {
    return fused_vegas_launch(fn_host, dtype, offsets, n_cubes, n_strat, row_begin, row_end, edges_packed, edges_layout, n_intervals,
                              weights, counts, hist_pairs, nullptr, JF, JF2, seed, call_idx, cube_block_log2, rank, world, out_f64,
                              ws, ws_bytes, stream);
}

int tq_fused_vegas_deferred(const tq_integrand* fn_host, int32_t dtype, const int64_t* offsets, int64_t n_cubes,
                            int32_t n_strat, int64_t row_begin, int64_t row_end, const void* edges_pairs, int64_t n_intervals,
                            void* jf2_rows, void* JF, void* JF2, uint64_t seed, uint32_t call_idx, void* ws, size_t ws_bytes,
                            void* stream) {
    TQ_REQUIRE(offsets && jf2_rows, "tq_fused_vegas_deferred: a stratified pass (offsets) and jf2_rows are required");
    return fused_vegas_launch(fn_host, dtype, offsets, n_cubes, n_strat, row_begin, row_end, edges_pairs, TQ_EDGES_PAIRS, n_intervals,
                              nullptr, nullptr, nullptr, jf2_rows, JF, JF2, seed, call_idx, 0, 0, 1, nullptr, ws, ws_bytes, stream);
}

int tq_vegas_hist_sweep(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim, int32_t dtype,
                        const void* jf2_rows, int64_t n_intervals, void* hist_pairs, int32_t dims_per_group, uint64_t seed,
                        uint32_t call_idx, void* ws, size_t ws_bytes, void* stream) {
    return hist_sweep_launch(offsets, n_cubes, n_strat, dim, dtype, jf2_rows, n_intervals, hist_pairs, dims_per_group, seed, call_idx,
                             0, 0, 1, ws, ws_bytes, stream);
}

}  // extern "C"

namespace tq {

// `n_cubes` is the GLOBAL cube count; with world > 1 `offsets` / `jf2_rows` are the rank's local arrays (CubeShard deal).
int hist_sweep_launch(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim, int32_t dtype, const void* jf2_rows,
                      int64_t n_intervals, void* hist_pairs, int32_t dims_per_group, uint64_t seed, uint32_t call_idx,
                      int32_t cube_block_log2, int32_t rank, int32_t world, void* ws, size_t ws_bytes, void* stream) {
    TQ_REQUIRE(offsets && jf2_rows && hist_pairs, "tq_vegas_hist_sweep: NULL argument");
    TQ_REQUIRE(world >= 1 && rank >= 0 && rank < world && cube_block_log2 >= 0 && cube_block_log2 < 31,
               "tq_vegas_hist_sweep: bad cube shard (block 2^%d, rank %d of %d)", cube_block_log2, rank, world);
    CubeShard shard;
    shard.lb = (uint32_t)cube_block_log2;
    shard.rank = (uint32_t)rank;
    shard.world.set((uint32_t)world);
    Workspace wsp(ws, ws_bytes);
    wsp.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    unsigned int* counters = wsp.take<unsigned int>(TQ_MAX_DIM);  // one chunk counter per launch of this call
    if (!counters) { set_error("tq_vegas_hist_sweep: workspace too small"); return TQ_ERR_WORKSPACE; }
    TQ_REQUIRE(n_cubes >= 1 && n_cubes < (1LL << 31) && n_strat >= 1 && dim >= 1 && dim <= TQ_MAX_DIM && n_intervals >= 1 &&
                   n_intervals < (1LL << 31), "tq_vegas_hist_sweep: bad sizes");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(counters, 0, TQ_MAX_DIM * sizeof(unsigned int), st);
    int launch = 0;
    FastDiv ns_div;
    ns_div.set((uint32_t)n_strat);
    TQ_DISPATCH_DTYPE(dtype, {
        constexpr int LANES = U01<T>::LANES;
        TQ_REQUIRE(dims_per_group >= 1 && dims_per_group <= LANES, "tq_vegas_hist_sweep: dims_per_group must be 1..%d for this dtype", LANES);
        const T inv_ns = (T)1 / (T)n_strat;
        static const int ctas_env = getenv("TQ_SWEEP_CTAS_PER_SM") ? atoi(getenv("TQ_SWEEP_CTAS_PER_SM")) : 0;
        // few CTAs: the cubes in flight then stay inside ONE band (measured, 8-D Ni=1e7: 1 / 2 / 4 / 8 / 16 CTAs per SM ->
        // 1.40 / 0.81 / 0.80 / 1.54 / 1.78 ms per dimension, profiles/r2/exp_sweep_grid.txt)
        const int grid = grid_for(n_cubes, 256, ctas_env > 0 ? ctas_env : 4);
        // groups never straddle a Philox block: dimensions [b * LANES, (b + 1) * LANES) in pieces of dims_per_group
        for (int b0 = 0; b0 < dim; b0 += LANES) {
            for (int d0 = b0; d0 < dim && d0 < b0 + LANES; d0 += dims_per_group) {
                // dimensions d0 .. d0+g-1 are lanes d0-b0 .. of Philox block b0 / LANES (a piece inside a block re-runs it)
                int g = dims_per_group;
                if (d0 + g > dim) g = dim - d0;
                if (d0 + g > b0 + LANES) g = b0 + LANES - d0;
                uint64_t pow_low = 1;
                for (int i = 0; i < d0; ++i) pow_low *= (uint64_t)n_strat;
                uint64_t pow_g = 1;
                for (int i = 0; i < g; ++i) pow_g *= (uint64_t)n_strat;
                FastDiv low_div, cpc_div;
                low_div.set((uint32_t)pow_low);
                cpc_div.set((uint32_t)((uint64_t)n_cubes / pow_g));
                hist_sweep_kernel<T><<<TQ_GRID(grid), 256, 0, st>>>((const long long*)offsets, (uint32_t)n_cubes, shard, (const T*)jf2_rows,
                                                                   (double*)hist_pairs, n_intervals, d0, g, ns_div, low_div,
                                                                   (uint32_t)pow_g, cpc_div, inv_ns, seed, call_idx, d0 - b0,
                                                                   counters + launch++);
            }
        }
    });
    return check_launch("hist_sweep_kernel");
}

int fused_vegas_launch(const tq_integrand* fn_host, int32_t dtype, const int64_t* offsets, int64_t n_cubes,
                       int32_t n_strat, int64_t row_begin, int64_t row_end, const void* edges_packed,
                       int32_t edges_layout, int64_t n_intervals, void* weights, int64_t* counts, void* hist_pairs,
                       void* jf2_rows, void* JF, void* JF2, uint64_t seed, uint32_t call_idx, int32_t cube_block_log2,
                       int32_t rank, int32_t world, double* out_f64, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_integrand("tq_fused_vegas", fn_host);
    if (rc) return rc;
    TQ_REQUIRE(world >= 1 && rank >= 0 && rank < world && cube_block_log2 >= 0 && cube_block_log2 < 31,
               "tq_fused_vegas: bad cube shard (block 2^%d, rank %d of %d)", cube_block_log2, rank, world);
    CubeShard shard;
    shard.lb = (uint32_t)cube_block_log2;
    shard.rank = (uint32_t)rank;
    shard.world.set((uint32_t)world);
    const bool strat = offsets != nullptr;
    const bool rows_from_offsets = row_end < 0;  // count stays on the device; -row_end is the sizing estimate
    if (rows_from_offsets) {
        TQ_REQUIRE(strat && row_begin == 0, "tq_fused_vegas: a device-side row count needs a stratified pass from row 0");
        row_end = -row_end;
    }
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_fused_vegas: bad row range");
    TQ_REQUIRE(n_intervals >= 1 && n_intervals < (1LL << 31), "tq_fused_vegas: n_intervals out of range");
    TQ_REQUIRE(!strat || (n_cubes >= 1 && n_cubes < (1LL << 31) && n_strat >= 1 && JF && JF2),
               "tq_fused_vegas: stratified pass needs n_cubes, n_strat, JF, JF2");
    TQ_REQUIRE(strat || out_f64 != nullptr, "tq_fused_vegas: warm-up pass needs out_f64");
    TQ_REQUIRE(edges_layout == TQ_EDGES_PAIRS || edges_layout == TQ_EDGES_RECORDS, "tq_fused_vegas: unknown edges layout %d", edges_layout);
    const bool records = edges_layout == TQ_EDGES_RECORDS;
    TQ_REQUIRE(!records || (weights == nullptr && counts == nullptr && hist_pairs == nullptr),
               "tq_fused_vegas: with TQ_EDGES_RECORDS the histogram lives in the records; pass weights = counts = hist_pairs = NULL");
    TQ_REQUIRE(!(weights && hist_pairs), "tq_fused_vegas: pass either weights/counts or hist_pairs, not both");
    TQ_REQUIRE((weights == nullptr) == (counts == nullptr) || records, "tq_fused_vegas: weights and counts go together");
    const int64_t nrows = row_end - row_begin;
    if (nrows == 0 && strat) return TQ_OK;
    Workspace w(ws, ws_bytes);
    unsigned int* ticket = w.take<unsigned int>(WS_HEADER / sizeof(unsigned int));
    const int dim = fn_host->dim;
    const size_t elt = dtype == TQ_F64 ? 8 : 4;
    const size_t ids_bytes = (size_t)FV_BLOCK * sizeof(double) + (size_t)dim * FV_BLOCK * sizeof(int);
    const size_t hist_bytes = (size_t)dim * n_intervals * (elt + 4);
    // CTAs: persistent, each with a contiguous chunk of rows; fewer CTAs when the problem is small.
    const int sms = num_sms();
    int64_t tiles = (nrows + FV_BLOCK - 1) / FV_BLOCK;
    if (tiles < 1) tiles = 1;
    // Shared-memory privatised histogram only for SMALL maps (<= 4096 bins, where same-address contention on
    // L2 atomics would serialise) and only when every CTA sees enough rows to amortise zero + flush.
    const int64_t bins = (int64_t)dim * n_intervals;
    int hist_mode = jf2_rows ? HIST_DEFER : records ? HIST_RECORDS : hist_pairs ? HIST_PAIRS : weights ? HIST_ARRAYS : HIST_NONE;
    if (hist_mode == HIST_ARRAYS && bins <= 4096 && nrows >= 64 * bins) hist_mode = HIST_SMEM;
    // Record layout = tables beyond L2: a sample's histogram reductions find its record still in L2 only if few
    // samples are in flight between the gather and the reduction.  Measured (8-D fp64, Ni=1e7): 6 CTAs/SM refetch
    // every record from HBM for the reductions (5 DRAM sectors read per gather, 1.7e9 evals/s); 2 CTAs/SM: 2.07e9.
    const int per_sm = records ? 2 : FV_MIN_CTAS;
    int64_t ctas = tiles < (int64_t)sms * per_sm ? tiles : (int64_t)sms * per_sm;
    if (hist_mode == HIST_SMEM) {
        const int64_t cap = nrows / (16 * bins) > 0 ? nrows / (16 * bins) : 1;
        if (ctas > cap) ctas = cap;
    }
    int64_t rows_per_cta = (nrows + ctas - 1) / ctas;
    static const int64_t rpc_env = getenv("TQ_FV_ROWS_PER_CTA") ? atoll(getenv("TQ_FV_ROWS_PER_CTA")) : 0;
    // maps beyond L2: deal the rows in small chunks, so that the rows in flight share the bands of the slow dimensions
    // (1024 rows per chunk: 13.4 ms vs 14.0 ms with one contiguous range per CTA, profiles/r2/exp_defer_interleave.txt)
    if (hist_mode == HIST_DEFER) {
        const int64_t rpc = rpc_env > 0 ? rpc_env : 1024;
        if (rpc < rows_per_cta) rows_per_cta = rpc;
    } else if (rpc_env > 0 && strat && rpc_env < rows_per_cta) {
        // experiment only: interleaved chunks for L2-resident maps make every CTA in flight hit the SAME bands, and the
        // reductions then contend for the same L2 sectors: 8-D fp64 pairs 1.48e10 -> 1.24e10 samples/s at 1024-row chunks
        // (profiles/r2/exp_interleave_l2_resident.txt).  One contiguous range per CTA stays the default.
        rows_per_cta = rpc_env;
    }
    rows_per_cta = ((rows_per_cta + FV_BLOCK - 1) / FV_BLOCK) * FV_BLOCK;
    if (rows_per_cta < FV_BLOCK) rows_per_cta = FV_BLOCK;
    ctas = nrows > 0 ? (nrows + rows_per_cta - 1) / rows_per_cta : 1;
    if (ctas > (int64_t)sms * per_sm) ctas = (int64_t)sms * per_sm;  // small chunks are dealt round-robin (the kernel's stride loop)
    double* partials = w.take<double>((size_t)ctas * 2);
    if (!ticket || !partials) { set_error("tq_fused_vegas: workspace too small"); return TQ_ERR_WORKSPACE; }
    const size_t smem = ids_bytes + (hist_mode == HIST_SMEM ? hist_bytes : 0);
    cudaStream_t st = as_stream(stream);
    FastDiv ns_div;
    ns_div.set((uint32_t)(strat ? n_strat : 1));
    // Band-privatised tile version (fused_vegas_tile_kernel): whole stratified passes into the pair table, when some high
    // dimensions' bands fit shared memory and there are enough tiles to balance one CTA per SM.
    // Measured on B200 (profiles/r2/exp_tile.txt): fp32 +8 .. +10 % over the all-L2 version (16-D N_strat=3: 8.07e9 -> 8.75e9
    // samples/s; 8-D N_strat=8: 1.65e10 -> 1.81e10), fp64 -5 % (8-D: 1.476e10 -> 1.405e10: a shared-memory fp64 atomicAdd is a
    // 64-bit compare-and-swap loop that retries when two lanes of a warp meet in a 512-bin band).  Default: fp32 only.
    static const int tile_env = getenv("TQ_FV_TILE") ? atoi(getenv("TQ_FV_TILE")) : -1;  // 0: never, 1: always, default: fp32
    // fp64: edges-only variant (hist_smem = false: staged edges, every histogram update through L2), behind TQ_FV_TILE64=1.
    // Measured slower than the plain kernel (8-D Ni=4096: 1.23e10 vs 1.50e10 samples/s): tiles are handed out in order, so all
    // CTAs reduce into the same slow-dimension bands at once (same-sector contention) at half the occupancy.  Off by default.
    static const int tile64_env = getenv("TQ_FV_TILE64") ? atoi(getenv("TQ_FV_TILE64")) : 0;
    const bool edges_only = dtype == TQ_F64 && tile_env != 1;
    const bool tile_ok = tile_env == 1 || (tile_env != 0 && (dtype == TQ_F32 || tile64_env == 1));
    if (tile_ok && hist_mode == HIST_PAIRS && strat && rows_from_offsets && n_strat >= 2 && dim >= 2 &&
        n_intervals <= (1 << 20)) {
        const int64_t band_w = (n_intervals + n_strat - 1) / n_strat + 3;
        const size_t hist_band = (size_t)band_w * (elt + 4), edge_band = (size_t)band_w * 2 * elt;
        static const int flush_factor = getenv("TQ_FV_TILE_FLUSH") ? atoi(getenv("TQ_FV_TILE_FLUSH")) : 4;  // rows per flushed bin
        static const int sb_env = getenv("TQ_FV_TILE_SB") ? atoi(getenv("TQ_FV_TILE_SB")) : 0;  // experiments: cap on band dimensions
        int best_g = -1, best_ne = 0;
        size_t best_smem = 0;
        uint64_t best_tile = 0;
        for (int sb = sb_env > 0 && sb_env < dim ? sb_env : dim - 1; sb >= 1 && band_w < 65535; --sb) {
            const int g = dim - sb;
            uint64_t tile_cubes = 1;
            for (int i = 0; i < g; ++i) tile_cubes *= (uint64_t)n_strat;
            if (tile_cubes > (uint64_t)n_cubes || (uint64_t)n_cubes % tile_cubes) continue;
            const uint64_t n_tiles = (uint64_t)n_cubes / tile_cubes;
            if (n_tiles < (uint64_t)sms * 4) break;                               // larger tiles only get fewer
            if (world > 1 && ((1ull << cube_block_log2) % tile_cubes)) continue;  // a tile must not straddle a rank's cube block
            // as many band dimensions as possible get a private histogram band; the shared memory left over stages edges
            const size_t fixed = edges_only ? 16 + (size_t)FT_BLOCK * 8 + (size_t)dim * FT_BLOCK * 4
                                            : (size_t)sb * hist_band + 16 + (size_t)FT_BLOCK * 8 + (size_t)g * FT_BLOCK * 4 + (size_t)sb * FT_BLOCK * 2;
            if (fixed > FT_SMEM_CAP) continue;
            if (edges_only) {  // nothing to flush: a tile only has to amortise its staging and barriers (>= 4 row steps)
                if ((uint64_t)nrows / n_tiles < 4ull * FT_BLOCK) continue;
            } else if ((uint64_t)nrows / n_tiles < (uint64_t)flush_factor * sb * band_w) continue;  // the flush must stay small next to the tile's rows
            int ne = (int)((FT_SMEM_CAP - fixed) / edge_band);
            if (ne > sb) ne = sb;
            if (edges_only && ne < 1) continue;
            best_g = g;
            best_ne = ne;
            best_smem = fixed + (size_t)ne * edge_band;
            best_tile = tile_cubes;
            break;
        }
        if (best_g > 0) {
            unsigned int* next_tile = w.take<unsigned int>(64);
            if (!next_tile) { set_error("tq_fused_vegas: workspace too small"); return TQ_ERR_WORKSPACE; }
            cudaMemsetAsync(next_tile, 0, sizeof(unsigned int), st);
            FastDiv tile_div;
            tile_div.set((uint32_t)best_tile);
            const uint64_t n_tiles = (uint64_t)n_cubes / best_tile;
            const unsigned grid = (unsigned)(n_tiles < (uint64_t)sms ? n_tiles : (uint64_t)sms);
            TQ_DISPATCH_DTYPE(dtype, {
                using P2 = typename Pair2<T>::type;
                const T inv_ns = (T)1 / (T)n_strat;
                TQ_DISPATCH_FAMILY(fn_host->family, {
                    cudaFuncSetAttribute(fused_vegas_tile_kernel<FAM, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FT_SMEM_CAP);
                    fused_vegas_tile_kernel<FAM, T><<<TQ_GRID(grid), FT_BLOCK, best_smem, st>>>(
                        *fn_host, (const long long*)offsets, (uint32_t)n_cubes, shard, ns_div, inv_ns, (T)n_strat, (T)n_intervals,
                        (const P2*)edges_packed, n_intervals,
                        (double*)hist_pairs, (T*)JF, (T*)JF2, seed, call_idx, best_g, (uint32_t)best_tile, tile_div, (int)band_w, best_ne, !edges_only, next_tile);
                });
            });
            return check_launch("fused_vegas_tile_kernel");
        }
    }
    TQ_DISPATCH_DTYPE(dtype, {
        const T inv_ns = (T)1 / (T)(strat ? n_strat : 1);  // RN(1 / N_strat) in the working precision (div_by_const)
        TQ_DISPATCH_FAMILY(fn_host->family, {
            if (strat) {
                if (smem > 48 * 1024) cudaFuncSetAttribute(fused_vegas_kernel<FAM, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
                fused_vegas_kernel<FAM, T, true><<<TQ_GRID((unsigned)ctas), FV_BLOCK, smem, st>>>(
                    *fn_host, (const long long*)offsets, n_cubes, shard, ns_div, inv_ns, (T)n_strat, (T)n_intervals, row_begin, row_end,
                    rows_from_offsets, rows_per_cta, edges_packed, records, n_intervals, (T*)weights, (unsigned long long*)counts, (double*)hist_pairs,
                    (T*)jf2_rows, (T*)JF, (T*)JF2, seed, call_idx, hist_mode, partials, ticket, out_f64);
            } else {
                if (smem > 48 * 1024) cudaFuncSetAttribute(fused_vegas_kernel<FAM, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
                fused_vegas_kernel<FAM, T, false><<<TQ_GRID((unsigned)ctas), FV_BLOCK, smem, st>>>(
                    *fn_host, nullptr, 0, shard, ns_div, inv_ns, (T)1, (T)n_intervals, row_begin, row_end, false, rows_per_cta, edges_packed,
                    records, n_intervals, (T*)weights, (unsigned long long*)counts, (double*)hist_pairs, nullptr, nullptr, nullptr, seed,
                    call_idx, hist_mode, partials, ticket, out_f64);
            }
        });
    });
    return check_launch("fused_vegas_kernel");
}

}  // namespace tq
