// VEGAS+ bookkeeping for SMALL problems (the launch-latency-bound regime: the reference's own N = 1e6
// configuration has 10^4 cubes and 4 x 4000 map bins).  get_NH + scan, the stratification update and the
// whole map update each run inside ONE thread-block cluster (8 CTAs that exchange partial sums through
// distributed shared memory and meet at cluster barriers), and one launch can carry the stratification
// update and every dimension's map update side by side.  The arithmetic is that of the large-problem kernels
// in vegas_strat.cu / vegas_map.cu (same rounding, fixed summation order).
// Reference: vegas_stratification.py:72-103, vegas.py:293-303, vegas_map.py:113-261.
#include <cooperative_groups.h>

#include "common.cuh"
#include "internal.cuh"
#include "vegas_dev.cuh"

namespace cg = cooperative_groups;

namespace tq {

constexpr int SM_CL = 8;          // CTAs per cluster (portable maximum)
constexpr int SM_THREADS = 512;   // threads per CTA
constexpr int SM_ITEMS = 8;       // items per thread at the size limits below
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr int64_t SM_MAX_CUBES = (int64_t)SM_CL * SM_THREADS * SM_ITEMS;  // 32768
constexpr long long SM_MAX_NI = (long long)SM_CL * SM_THREADS * SM_ITEMS;
constexpr int SM_MAX_DIM = 64;
constexpr int SM_COARSE = 32;     // stride of the shared-memory search table over the prefix sums

bool small_strat_ok(int64_t n_cubes) { return n_cubes >= 1 && n_cubes <= SM_MAX_CUBES; }
bool small_map_ok(int32_t dim, int64_t ni) { return dim >= 1 && dim <= SM_MAX_DIM && ni >= 2 && ni <= SM_MAX_NI; }

template <typename T>
struct StratArgs {
    const T* JF;
    const T* JF2;
    long long* nh;
    int64_t n_cubes;
    T V, V2, beta;
    T* dh;
    double* scalars;
    T next_nev;
    long long* offsets;
    uint32_t* clear;
    int64_t clear_words;
};

template <typename T>
struct MapArgs {
    T* xe;
    T* dxe;
    T* weights;
    long long* counts;
    typename EdgePair<T>::type* packed;
    T* avg;
    T* smoothed;
    double* S;
    T* x_new;
    int dim;
    long long ni;
    T alpha;
    int32_t* status;
    bool do_edges;
};

__device__ __forceinline__ void cluster_clear(uint32_t* __restrict__ clear, int64_t words, unsigned rank) {
    for (int64_t i = (int64_t)rank * SM_THREADS + threadIdx.x; i < words; i += SM_CL * SM_THREADS) clear[i] = 0u;
}

// Exclusive scan over the cluster of `per` consecutive counts per thread (thread t of CTA `rank` owns cubes
// [c0, c0 + per)); writes nh, offsets and the total.  One CTA scan + one exchange of CTA totals.
__device__ __forceinline__ void cluster_nh_scan(cg::cluster_group& cluster, unsigned rank, const long long (&v)[SM_ITEMS],
                                                int per, int64_t c0, int64_t n_cubes, long long* __restrict__ nh,
                                                long long* __restrict__ offsets) {
    __shared__ long long sh[33];
    __shared__ long long s_total;
    long long run = 0;
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) run += v[i];
    long long total;
    long long ex = block_excl_scan<long long>(run, sh, total);
    if (threadIdx.x == 0) s_total = total;
    cluster.sync();
    long long all = 0;
    for (unsigned r = 0; r < SM_CL; ++r) {
        const long long t = *cluster.map_shared_rank(&s_total, r);
        if (r < rank) ex += t;
        all += t;
    }
    cluster.sync();  // no CTA may leave (or reuse s_total) while its shared memory is still being read
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
        if (i < per && c0 + i < n_cubes) {
            nh[c0 + i] = v[i];
            offsets[c0 + i] = ex;
        }
        ex += v[i];
    }
    if (rank == 0 && threadIdx.x == 0) offsets[n_cubes] = all;
}

// ---------------------------------------------------------------- get_NH + offsets (+ accumulator reset)
template <typename T>
__global__ void __cluster_dims__(SM_CL, 1, 1) __launch_bounds__(SM_THREADS)
nh_small_kernel(const T* __restrict__ dh, int64_t n_cubes, T nev, long long* __restrict__ nh,
                long long* __restrict__ offsets, uint32_t* __restrict__ clear, int64_t clear_words) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    cluster_clear(clear, clear_words, rank);
    const int per = (int)((n_cubes + SM_CL * SM_THREADS - 1) / (SM_CL * SM_THREADS));
    const int64_t c0 = ((int64_t)rank * SM_THREADS + threadIdx.x) * per;
    long long v[SM_ITEMS];
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) v[i] = (i < per && c0 + i < n_cubes) ? nh_of<T>(dh[c0 + i], nev) : 0;
    cluster_nh_scan(cluster, rank, v, per, c0, n_cubes, nh, offsets);
}

// ---------------------------------------------------------------- estimator + update_DH (+ next get_NH)
// The d^beta values stay in registers between the reduction and the normalisation; the sums cross the cluster
// in rank order, so every CTA normalises by the same value.  With next_nev > 0 the cluster goes straight on to
// the NEXT pass's get_NH + scan from the normalised dh it still holds, and resets the JF/JF2 accumulators it
// has just consumed, which removes a launch from every iteration of the native loop.
template <typename T>
__device__ __forceinline__ void strat_update_body(cg::cluster_group& cluster, const StratArgs<T>& a) {
    __shared__ double sh[32 * 4];
    __shared__ double s_part[4];
    const unsigned rank = cluster.block_rank();
    const int64_t n_cubes = a.n_cubes;
    const int per = (int)((n_cubes + SM_CL * SM_THREADS - 1) / (SM_CL * SM_THREADS));
    const int64_t c0 = ((int64_t)rank * SM_THREADS + threadIdx.x) * per;
    T p[SM_ITEMS];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
        const int64_t c = c0 + i;
        p[i] = (T)0;
        if (i < per && c < n_cubes) {
            const long long nc = a.nh[c];
            const T n = (T)nc;
            const T inv = div_rn((T)1, n);
            const T jf = a.JF[c], jf2 = a.JF2[c];
            // vegas.py:293-303
            const T ih = mul_rn(jf, mul_rn(inv, a.V));
            const T sig2 = fabs(sub_rn(mul_rn(mul_rn(jf2, inv), a.V2), mul_rn(ih, ih)));
            acc[0] += (double)ih;
            acc[1] += (double)mul_rn(sig2, inv);
            // vegas_stratification.py:78-85
            const T m = div_rn(mul_rn(a.V, jf), n);
            T dv = sub_rn(div_rn(mul_rn(a.V2, jf2), n), mul_rn(m, m));
            if (dv < (T)0) dv = (T)0;
            p[i] = pow(dv, a.beta);
            acc[2] += (double)p[i];
            acc[3] += (double)nc;
        }
    }
    block_sum<4>(acc, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s_part[k] = acc[k];
    }
    cluster.sync();  // every CTA has consumed its JF/JF2/nh slice
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    for (unsigned r = 0; r < SM_CL; ++r) {
        const double* rp = cluster.map_shared_rank(s_part, r);
#pragma unroll
        for (int k = 0; k < 4; ++k) tot[k] += rp[k];
    }
    if (rank == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) a.scalars[k] = tot[k];
    }
    const T s = (T)tot[2];
    long long v[SM_ITEMS];
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
        const int64_t c = c0 + i;
        v[i] = 0;
        if (i < per && c < n_cubes) {
            const T dhc = s == (T)0 ? p[i] : div_rn(p[i], s);  // vegas_stratification.py:89-90
            a.dh[c] = dhc;
            if (a.next_nev > (T)0) v[i] = nh_of<T>(dhc, a.next_nev);
        }
    }
    if (a.next_nev > (T)0) {
        cluster_clear(a.clear, a.clear_words, rank);
        cluster_nh_scan(cluster, rank, v, per, c0, n_cubes, a.nh, a.offsets);  // syncs the cluster before it returns
    } else {
        cluster.sync();  // no CTA may leave while its s_part is still being read
    }
}

// ---------------------------------------------------------------- map update: one cluster per dimension
// average -> smooth -> fp64 prefix -> new edges -> repair/diff/reset back to back.  Each CTA owns a contiguous
// slice of the dimension's bins and each thread `per` consecutive bins, kept in registers from the smoothing
// to the prefix sums.  The "any dimension sums to zero" decision needs all row sums; each cluster recomputes
// them (dim * Ni reads per cluster, tiny at these sizes) instead of synchronising across clusters.
template <typename T>
__device__ __forceinline__ void map_update_body(cg::cluster_group& cluster, const MapArgs<T>& a, int d, bool first_cluster) {
    __shared__ double sh[33];
    __shared__ double s_part[SM_MAX_DIM];  // this CTA's share of every dimension's row sum
    __shared__ double s_warp[SM_MAX_DIM * SM_WARPS];
    __shared__ double s_tot[SM_MAX_DIM];
    __shared__ double s_slice;             // this CTA's share of the smoothed row sum
    __shared__ double s_coarse[SM_MAX_NI / SM_COARSE];
    const unsigned rank = cluster.block_rank();
    const int tid = threadIdx.x;
    const int dim = a.dim;
    const long long ni = a.ni;
    const int per = (int)((ni + SM_CL * SM_THREADS - 1) / (SM_CL * SM_THREADS));
    const long long slice = (long long)per * SM_THREADS;
    const long long j_lo = (long long)rank * slice;
    const long long j_hi = j_lo + slice < ni ? j_lo + slice : ni;
    T* avg = a.avg + (int64_t)d * ni;
    // ---- averages with the zero-count fill; row sums of every dimension (vegas_map.py:118-144,150).
    // Warp partials go to shared memory and ONE barrier follows the loop, so the loads of all dimensions are
    // in flight together.
    const int lane = tid & 31, warp = tid >> 5;
    for (int dd = 0; dd < dim; ++dd) {
        const T* w = a.weights + (int64_t)dd * ni;
        const long long* c = a.counts + (int64_t)dd * ni;
        double part = 0.0;
        for (long long j = j_lo + tid; j < j_hi; j += SM_THREADS) {
            const T v = filled_average<T>(w, c, j, ni);
            if (dd == d) avg[j] = v;
            part += (double)v;
        }
        part = warp_sum(part);
        if (lane == 0) s_warp[dd * SM_WARPS + warp] = part;
    }
    __syncthreads();
    if (tid < dim) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < SM_WARPS; ++k) t += s_warp[tid * SM_WARPS + k];
        s_part[tid] = t;
    }
    cluster.sync();  // also publishes this cluster's avg[] slices to its other CTAs
    if (tid < dim) {
        double t = 0.0;
        for (unsigned r = 0; r < SM_CL; ++r) t += cluster.map_shared_rank(s_part, r)[tid];
        s_tot[tid] = t;
    }
    __syncthreads();
    bool any_zero = false;
    for (int dd = 0; dd < dim; ++dd) any_zero |= ((T)s_tot[dd] == (T)0);
    if (any_zero) {  // the reference skips the whole update (vegas_map.py:192-197), keeping only the reset
        if (first_cluster && rank == 0 && tid == 0) a.status[0] = 1;
        if (a.do_edges) {
            for (long long j = j_lo + tid; j < j_hi; j += SM_THREADS) {
                a.weights[(int64_t)d * ni + j] = (T)0;
                a.counts[(int64_t)d * ni + j] = 0;
            }
        }
        cluster.sync();  // no CTA may leave while its shared memory is still being read
        return;
    }
    // ---- smoothing + compression (vegas_map.py:146-170); thread owns bins [jt, jt + per)
    T* sm = a.smoothed + (int64_t)d * ni;
    const T denom = mul_rn((T)8, (T)s_tot[d]);
    const long long jt = j_lo + (long long)tid * per;
    T v[SM_ITEMS];
    double run = 0.0;
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
        const long long j = jt + i;
        v[i] = (T)0;
        if (i < per && j < ni) {
            T x;
            if (j == 0) x = add_rn(mul_rn((T)7, avg[0]), avg[1]);
            else if (j == ni - 1) x = add_rn(avg[ni - 2], mul_rn((T)7, avg[ni - 1]));
            else x = add_rn(add_rn(avg[j - 1], mul_rn((T)6, avg[j])), avg[j + 1]);
            x = div_rn(x, denom);
            if (x != (T)0) {
                const T base = div_rn(sub_rn(x, (T)1), log(x));
                x = (a.alpha == (T)0.5) ? sqrt(base) : pow(base, a.alpha);  // ATen evaluates x**0.5 as sqrt
            }
            v[i] = x;
            sm[j] = x;
            run += (double)x;
        }
    }
    if (!a.do_edges) {
        cluster.sync();
        return;
    }
    // ---- fp64 inclusive prefix sums (vegas_map.py:207-213): CTA scan + rank-ordered slice totals
    double total;
    double ex = block_excl_scan<double>(run, sh, total);
    if (tid == 0) s_slice = total;
    cluster.sync();
    double row_total = 0.0;
    for (unsigned r = 0; r < SM_CL; ++r) {
        const double t = *cluster.map_shared_rank(&s_slice, r);
        if (r < rank) ex += t;
        row_total += t;
    }
    double* Sd = a.S + (int64_t)d * ni;
#pragma unroll
    for (int i = 0; i < SM_ITEMS; ++i) {
        const long long j = jt + i;
        if (i < per && j < ni) {
            ex += (double)v[i];
            Sd[j] = ex;
        }
    }
    cluster.sync();
    // ---- new inner edges (vegas_map.py:214-239).  For m = 0..Ni-2:
    //   idx = #{j <= Ni-2 : trunc(S_j/delta) <= m}   (the reference builds it as histogram + cumsum)
    //   acc = (m+1)*delta - S_{idx-1}                (reference: cumsum of delta - val_per_multiple)
    //   x_new[m+1] = xe[idx] + acc/sm[idx]*dxe[idx]
    // Two-level search: a coarse table of every SM_COARSE-th prefix sum in shared memory, then the block.
    const T* x_old = a.xe + (int64_t)d * (ni + 1);
    const T* dx_old = a.dxe + (int64_t)d * ni;
    T* xn = a.x_new + (int64_t)d * (ni + 1);
    const T delta_t = div_rn((T)row_total, (T)ni);
    const double delta = (double)delta_t;
    const long long last = ni - 2;  // largest searchable index
    const int nblk = (int)((last + SM_COARSE) / SM_COARSE);  // blocks of SM_COARSE covering [0, last]
    for (int b = tid; b < nblk; b += SM_THREADS) {
        const long long e = (long long)b * SM_COARSE + SM_COARSE - 1;
        s_coarse[b] = Sd[e < last ? e : last];
    }
    if (rank == 0 && tid == 0) { xn[0] = x_old[0]; xn[ni] = x_old[ni]; }
    __syncthreads();
    for (long long m = j_lo + tid; m < j_hi && m <= last; m += SM_THREADS) {
        // smallest j in [0, last] with trunc(S_j / delta) > m; ni - 1 when there is none.  The predicate is
        // S_j >= thr (division_threshold, vegas_dev.cuh); thr < 0 keeps the division for degenerate delta.
        const double thr = division_threshold((double)(m + 1), delta);
        auto reached = [&](double sj) { return thr >= 0.0 ? sj >= thr : (long long)__ddiv_rn(sj, delta) > m; };
        int lo = 0, hi = nblk;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (reached(s_coarse[mid])) hi = mid; else lo = mid + 1;
        }
        long long idx = ni - 1;
        if (lo < nblk) {
            long long flo = (long long)lo * SM_COARSE, fhi = flo + SM_COARSE - 1;
            if (fhi > last) fhi = last;
            while (flo < fhi) {
                const long long mid = (flo + fhi) >> 1;
                if (reached(Sd[mid])) fhi = mid; else flo = mid + 1;
            }
            idx = flo;
        }
        const double below = idx > 0 ? Sd[idx - 1] : 0.0;
        const T acc = (T)((double)(m + 1) * delta - below);
        xn[m + 1] = add_rn(x_old[idx], mul_rn(div_rn(acc, sm[idx]), dx_old[idx]));
    }
    cluster.sync();
    // ---- repair, diff, pack, reset (vegas_map.py:240-261)
    const long long e_hi = (j_lo < ni && j_hi == ni) ? ni + 1 : j_hi;  // edge ni goes with the last non-empty slice
    for (long long e = j_lo + tid; e < e_hi; e += SM_THREADS) {
        bool bad, still;
        const T val = repaired_edge<T>(xn, e, ni, bad, still);
        if (bad) atomicAdd(&a.status[1], 1);
        if (still) a.status[2] = 1;
        a.xe[(int64_t)d * (ni + 1) + e] = val;
        if (e < ni) {
            bool b2, s2;
            const T dv = sub_rn(repaired_edge<T>(xn, e + 1, ni, b2, s2), val);
            a.dxe[(int64_t)d * ni + e] = dv;
            if (a.packed) {
                typename EdgePair<T>::type pr;
                pr.x = val;
                pr.y = dv;
                a.packed[(int64_t)d * ni + e] = pr;
            }
            a.weights[(int64_t)d * ni + e] = (T)0;
            a.counts[(int64_t)d * ni + e] = 0;
        }
    }
}

// One launch: cluster 0 runs the stratification update when HAS_STRAT, the next `dim` clusters one map
// dimension each when HAS_MAP.
template <typename T, bool HAS_STRAT, bool HAS_MAP>
__global__ void __cluster_dims__(SM_CL, 1, 1) __launch_bounds__(SM_THREADS)
vegas_update_small_kernel(const StratArgs<T> sa, const MapArgs<T> ma) {
    cg::cluster_group cluster = cg::this_cluster();
    const int cl = blockIdx.x / SM_CL;
    if (HAS_STRAT && cl == 0) {
        strat_update_body<T>(cluster, sa);
        return;
    }
    if (HAS_MAP) {
        const int d = cl - (HAS_STRAT ? 1 : 0);
        map_update_body<T>(cluster, ma, d, d == 0);
    }
}

int small_nh_launch(const void* dh, int64_t n_cubes, double nevals_exp, int32_t dtype, int64_t* nh, int64_t* offsets,
                    void* clear, size_t clear_bytes, void* stream) {
    TQ_REQUIRE(small_strat_ok(n_cubes), "small_nh_launch: %lld cubes out of range", (long long)n_cubes);
    TQ_DISPATCH_DTYPE(dtype, {
        nh_small_kernel<T><<<TQ_GRID(SM_CL), SM_THREADS, 0, as_stream(stream)>>>((const T*)dh, n_cubes, (T)nevals_exp, (long long*)nh,
                                                                       (long long*)offsets, (uint32_t*)clear,
                                                                       (int64_t)(clear ? clear_bytes / 4 : 0));
    });
    return check_launch("nh_small_kernel");
}

int small_update_launch(const SmallStrat* st, const SmallMap* mp, int32_t dtype, void* stream) {
    TQ_REQUIRE(st || mp, "small_update_launch: nothing to do");
    TQ_REQUIRE(!st || small_strat_ok(st->n_cubes), "small_update_launch: cube count out of range");
    TQ_REQUIRE(!mp || small_map_ok(mp->dim, mp->ni), "small_update_launch: map shape out of range");
    cudaStream_t s = as_stream(stream);
    TQ_DISPATCH_DTYPE(dtype, {
        StratArgs<T> sa = {};
        MapArgs<T> ma = {};
        if (st) {
            sa.JF = (const T*)st->JF;
            sa.JF2 = (const T*)st->JF2;
            sa.nh = (long long*)st->nh;
            sa.n_cubes = st->n_cubes;
            sa.V = (T)st->v_cubes;
            sa.V2 = (T)(st->v_cubes * st->v_cubes);
            sa.beta = (T)st->beta;
            sa.dh = (T*)st->dh;
            sa.scalars = st->scalars;
            sa.next_nev = st->next_nevals > 0 ? (T)st->next_nevals : (T)0;
            sa.offsets = (long long*)st->offsets;
            sa.clear = (uint32_t*)st->clear;
            sa.clear_words = st->clear ? (int64_t)(st->clear_bytes / 4) : 0;
        }
        if (mp) {
            ma.xe = (T*)mp->x_edges;
            ma.dxe = (T*)mp->dx_edges;
            ma.weights = (T*)mp->weights;
            ma.counts = (long long*)mp->counts;
            ma.packed = (typename EdgePair<T>::type*)mp->edges_packed;
            ma.avg = (T*)mp->scratch.avg;
            ma.smoothed = (T*)mp->scratch.smoothed;
            ma.S = mp->scratch.S;
            ma.x_new = (T*)mp->scratch.x_new;
            ma.dim = mp->dim;
            ma.ni = mp->ni;
            ma.alpha = (T)mp->alpha;
            ma.status = mp->status;
            ma.do_edges = mp->do_edges;
        }
        if (st && mp) vegas_update_small_kernel<T, true, true><<<TQ_GRID((1 + mp->dim) * SM_CL), SM_THREADS, 0, s>>>(sa, ma);
        else if (st) vegas_update_small_kernel<T, true, false><<<TQ_GRID(SM_CL), SM_THREADS, 0, s>>>(sa, ma);
        else vegas_update_small_kernel<T, false, true><<<TQ_GRID(mp->dim * SM_CL), SM_THREADS, 0, s>>>(sa, ma);
    });
    return check_launch("vegas_update_small_kernel");
}

}  // namespace tq
