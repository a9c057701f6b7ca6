// VEGAS passes with a CALLBACK integrand (any Python callable -- the path every user of the reference is on), as two
// kernels around the user's evaluation instead of four:
//   sample_map_kernel        get_Y (vegas_stratification.py:140-165) + get_X / get_Jac (vegas_map.py:44-74) + the unit-cube ->
//                            domain transform (vegas.py:109-110): the stratified y never reaches HBM, only x and jac do
//   accumulate_regen_kernel  jf = (f * V) * jac (vegas.py:104-112,284-287) + VEGASMap.accumulate_weight (vegas_map.py:99-111):
//                            the bin ids are REGENERATED from the cube-keyed Philox stream (same block, same arithmetic as
//                            sample_map_kernel: identical bins) instead of being read back from a materialised y, and the
//                            histogram uses the sector-paired reductions of the fused pass (fused.cu)
// Per sample the library's own HBM traffic drops from 5*dim*s + 5s bytes (y written, read twice; x; jac; f; jf) to
// dim*s + 5s.  Arithmetic is operation for operation that of strat_sample_kernel -> map_forward_kernel ->
// map_accumulate_global_kernel, so x, jac and jf are bit-identical to that pipeline (tests/test_gpu_kernels.py).
#include "common.cuh"
#include "internal.cuh"
#include "vegas_dev.cuh"

namespace tq {

constexpr int UF_BLOCK = 256;
constexpr int UF_SLICE = UF_BLOCK / 2 + 4;  // cubes a tile of UF_BLOCK rows can overlap (nh >= 2) + slack
constexpr int UF_XPAD = UF_BLOCK + 1;       // row stride of the staging tiles (odd: conflict-free transposed reads)
#ifndef TQ_UF_MIN_CTAS
#define TQ_UF_MIN_CTAS 4  // 58-64 registers without spills (6 CTAs: 40 registers, spills in every instantiation)
#endif
constexpr int UF_MIN_CTAS = TQ_UF_MIN_CTAS;

// Row -> (cube, index in cube) for CTA tiles of UF_BLOCK consecutive rows of the cube-sorted order, as in the fused pass:
// one thread per overlapped cube publishes its offset and marks its rows; one barrier per tile (double-buffered).
struct RowCubes {
    long long off[2][UF_SLICE];
    unsigned char cube[2][UF_BLOCK];
};

// largest c with offsets[c] <= r_lo, by a CTA-wide UF_BLOCK-ary search (two load latencies for 10^4 cubes)
__device__ __forceinline__ long long first_cube_of_chunk(const long long* __restrict__ offsets, int64_t n_cubes, int64_t r_lo) {
    long long c_lo = 0, hi = n_cubes;
    while (hi - c_lo > 1) {
        const long long step = (hi - c_lo + UF_BLOCK - 1) / UF_BLOCK;
        const long long p = c_lo + (long long)(threadIdx.x + 1) * step;
        const int below = __syncthreads_count(p < hi && __ldg(&offsets[p]) <= r_lo);
        c_lo += below * step;
        if (c_lo + step < hi) hi = c_lo + step;
    }
    return c_lo;
}

__device__ __forceinline__ void fill_row_cubes(RowCubes& rc, int buf, const long long* __restrict__ offsets, int64_t n_cubes,
                                               long long c_lo, int64_t rb, int64_t re) {
    if (threadIdx.x < UF_SLICE) {
        const long long c = c_lo + threadIdx.x;
        const long long lo = __ldg(&offsets[c < n_cubes ? c : n_cubes]);
        const long long hi = __ldg(&offsets[c + 1 < n_cubes ? c + 1 : n_cubes]);
        rc.off[buf][threadIdx.x] = lo;
        const long long a = lo > rb ? lo : rb, b = hi < re ? hi : re;
        for (long long r = a; r < b; ++r) rc.cube[buf][r - rb] = (unsigned char)threadIdx.x;
    }
    __syncthreads();
}

// y of dimension d and its bin: the digit walk and arithmetic of strat_sample_kernel / bin_of (warm-up: y = u * 0.999999)
template <typename T, bool STRAT>
__device__ __forceinline__ int sample_bin(T u, uint32_t& c, const FastDiv& ns_div, T nsf, T inv_ns, T nif, int ni, T& o) {
    T y;
    if (STRAT) {
        const uint32_t q = ns_div.div(c);
        const uint32_t p = c - q * ns_div.d;
        c = q;
        y = div_by_const(add_rn((T)p, u), nsf, inv_ns);
        if (y >= (T)1) y = (T)0.999999;
    } else {
        y = mul_rn(u, (T)0.999999);
    }
    const T t = mul_rn(y, nif);
    const T fl = floor(t);
    o = sub_rn(t, fl);
    int k = (int)fl;  // 0 <= t < 2^31 (n_intervals < 2^31 is checked by the launcher)
    k = k < 0 ? 0 : (k >= ni ? ni - 1 : k);
    return k;
}

// ------------------------------------------------------------------ sample + map
// One thread per row; the row's x values are staged in shared memory, DG dimensions at a time, and leave as coalesced
// stores (a group of DG dimensions is 64 bytes of a row: whole sectors).
template <typename T, bool STRAT>
__global__ void __launch_bounds__(UF_BLOCK, UF_MIN_CTAS)
sample_map_kernel(const long long* __restrict__ offsets, int64_t n_cubes, FastDiv ns_div, T inv_ns, int64_t row_begin,
                  int64_t row_end, int64_t rows_per_cta, const typename EdgePair<T>::type* __restrict__ edges, int ep_stride, int ni,
                  int dim, const T* __restrict__ domain, T* __restrict__ x_out, T* __restrict__ jac_out, uint64_t seed,
                  uint32_t call) {
    constexpr int LANES = U01<T>::LANES;
    constexpr int DG = 64 / sizeof(T);  // dimensions per staging group (a multiple of LANES)
    using P2 = typename EdgePair<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ RowCubes rc;
    __shared__ T s_x[DG * UF_XPAD];
    T* s_start = reinterpret_cast<T*>(smem_raw);  // [dim]
    T* s_size = s_start + dim;                    // [dim]
    for (int d = threadIdx.x; d < dim; d += UF_BLOCK) {
        const T a = domain[2 * d], b = domain[2 * d + 1];
        s_start[d] = a;
        s_size[d] = sub_rn(b, a);
    }
    __syncthreads();
    const T nif = (T)ni, nsf = (T)ns_div.d;
    int buf = 0;
    for (int64_t r_lo = row_begin + (int64_t)blockIdx.x * rows_per_cta; r_lo < row_end; r_lo += (int64_t)gridDim.x * rows_per_cta) {
        const int64_t r_hi = r_lo + rows_per_cta < row_end ? r_lo + rows_per_cta : row_end;
        long long c_lo = STRAT ? first_cube_of_chunk(offsets, n_cubes, r_lo) : 0;
        for (int64_t rb = r_lo; rb < r_hi; rb += UF_BLOCK, buf ^= 1) {
            const int64_t re = rb + UF_BLOCK < r_hi ? rb + UF_BLOCK : r_hi;
            const int rows_here = (int)(re - rb);
            const int64_t row = rb + threadIdx.x;
            const bool active = row < re;
            if (STRAT) fill_row_cubes(rc, buf, offsets, n_cubes, c_lo, rb, re);
            uint32_t i0 = 0, i1 = 0, c = 0;
            if (active) {
                if (STRAT) {
                    const int key = rc.cube[buf][threadIdx.x];
                    i0 = (uint32_t)(c_lo + key);
                    i1 = (uint32_t)(row - rc.off[buf][key]);
                    c = i0;
                } else {
                    i0 = (uint32_t)(uint64_t)row;
                    i1 = (uint32_t)((uint64_t)row >> 32);
                }
            }
            T jac = (T)1;
            for (int g0 = 0; g0 < dim; g0 += DG) {
                const int gw = dim - g0 < DG ? dim - g0 : DG;
                if (active) {
                    for (int d0 = g0; d0 < g0 + gw; d0 += LANES) {
                        T u[LANES];
                        philox_block<T>(seed, call, i0, i1, (uint32_t)(d0 / LANES), u);
#pragma unroll
                        for (int j = 0; j < LANES; ++j) {
                            const int d = d0 + j;
                            if (d < dim) {
                                T o;
                                const int k = sample_bin<T, STRAT>(u[j], c, ns_div, nsf, inv_ns, nif, ni, o);
                                const P2 e = __ldg(&edges[((int64_t)d * ni + k) * ep_stride]);
                                const T x = add_rn(e.x, mul_rn(e.y, o));
                                jac = mul_rn(jac, mul_rn(nif, e.y));
                                s_x[(d - g0) * UF_XPAD + threadIdx.x] = add_rn(mul_rn(x, s_size[d]), s_start[d]);
                            }
                        }
                    }
                }
                __syncthreads();
                {   // rows_here x gw values -> x_out[(row - row_begin) * dim + g0 ..], consecutive threads = consecutive addresses of a row
                    const uint32_t magic = (65536u + (uint32_t)gw - 1u) / (uint32_t)gw;  // e / gw exactly for e * gw < 65536
                    T* out = x_out + (rb - row_begin) * dim + g0;
                    const int n_el = rows_here * gw;
                    for (int e = threadIdx.x; e < n_el; e += UF_BLOCK) {
                        const int r = (int)(((uint32_t)e * magic) >> 16);
                        const int dl = e - r * gw;
                        out[(int64_t)r * dim + dl] = s_x[dl * UF_XPAD + r];
                    }
                }
                __syncthreads();
            }
            if (active) jac_out[row - row_begin] = jac;
            if (STRAT) c_lo += rc.cube[buf][rows_here - 1];  // cube of the tile's last row: where the next tile starts
        }
    }
}

// ------------------------------------------------------------------ accumulate with regenerated bins
// jf = (f * volume) * jac (written to jf_out; jf^2 to jf2_out for tq_vegas_hist_sweep), and when a histogram target is given weights[d, k] += jf^2,
// counts[d, k] += 1 for the row's bins.  Targets: the fp64 pair table {sum jf^2, count} (hist != NULL), the
// {weight, count} words of the large-map records (recs != NULL), or the weights / counts arrays themselves.  Bin ids of a warp's 32 rows are parked in shared memory,
// DG dimensions at a time; lanes 2i / 2i+1 then add {jf^2, 1.0} of row i with ONE reduction per bin sector (fused.cu).
template <typename T, bool STRAT>
__global__ void __launch_bounds__(UF_BLOCK, UF_MIN_CTAS)
accumulate_regen_kernel(const long long* __restrict__ offsets, int64_t n_cubes, FastDiv ns_div, T inv_ns, int64_t row_begin,
                        int64_t row_end, int64_t rows_per_cta, int ni, int dim, const T* __restrict__ f,
                        const T* __restrict__ jacp, T volume, T* __restrict__ jf_out, T* __restrict__ jf2_out, double* __restrict__ hist,
                        MapRecord<T>* __restrict__ recs, T* __restrict__ weights, unsigned long long* __restrict__ counts, uint64_t seed,
                        uint32_t call) {
    constexpr int LANES = U01<T>::LANES;
    constexpr int DG = 16;
    __shared__ RowCubes rc;
    __shared__ int s_ids[DG * UF_BLOCK];
    __shared__ double s_jf2[UF_BLOCK];
    const bool do_hist = hist != nullptr || recs != nullptr || weights != nullptr;
    // fp64 records keep their count as an fp64 next to the weight: the same sector-paired reduction applies
    const bool paired = hist != nullptr || (recs != nullptr && sizeof(T) == 8);
    double* pair_base = hist ? hist : reinterpret_cast<double*>(recs) + 2;
    const int pair_stride = hist ? 2 : 4;
    const T nif = (T)ni, nsf = (T)ns_div.d;
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31, word = lane & 1;
    int buf = 0;
    for (int64_t r_lo = row_begin + (int64_t)blockIdx.x * rows_per_cta; r_lo < row_end; r_lo += (int64_t)gridDim.x * rows_per_cta) {
        const int64_t r_hi = r_lo + rows_per_cta < row_end ? r_lo + rows_per_cta : row_end;
        long long c_lo = (STRAT && do_hist) ? first_cube_of_chunk(offsets, n_cubes, r_lo) : 0;
        for (int64_t rb = r_lo; rb < r_hi; rb += UF_BLOCK, buf ^= 1) {
            const int64_t re = rb + UF_BLOCK < r_hi ? rb + UF_BLOCK : r_hi;
            const int64_t row = rb + threadIdx.x;
            const bool active = row < re;
            T jf2 = (T)0;
            if (active) {
                const T jf = mul_rn(mul_rn(__ldcs(&f[row - row_begin]), volume), __ldcs(&jacp[row - row_begin]));
                jf2 = mul_rn(jf, jf);
                if (jf_out) jf_out[row - row_begin] = jf;
                if (jf2_out) jf2_out[row - row_begin] = jf2;
            }
            if (!do_hist) continue;  // no grid improvement: jf only
            if (STRAT) fill_row_cubes(rc, buf, offsets, n_cubes, c_lo, rb, re);
            uint32_t i0 = 0, i1 = 0, c = 0;
            if (active) {
                if (STRAT) {
                    const int key = rc.cube[buf][threadIdx.x];
                    i0 = (uint32_t)(c_lo + key);
                    i1 = (uint32_t)(row - rc.off[buf][key]);
                    c = i0;
                } else {
                    i0 = (uint32_t)(uint64_t)row;
                    i1 = (uint32_t)((uint64_t)row >> 32);
                }
            }
            s_jf2[threadIdx.x] = (double)jf2;
            for (int g0 = 0; g0 < dim; g0 += DG) {
                const int gw = dim - g0 < DG ? dim - g0 : DG;
                if (active) {
                    for (int d0 = g0; d0 < g0 + gw; d0 += LANES) {
                        T u[LANES];
                        philox_block<T>(seed, call, i0, i1, (uint32_t)(d0 / LANES), u);
#pragma unroll
                        for (int j = 0; j < LANES; ++j) {
                            const int d = d0 + j;
                            if (d < dim) {
                                T o;
                                s_ids[(d - g0) * UF_BLOCK + threadIdx.x] = sample_bin<T, STRAT>(u[j], c, ns_div, nsf, inv_ns, nif, ni, o);
                            }
                        }
                    }
                } else {
                    s_ids[threadIdx.x] = -1;  // the group's first dimension marks the row: inactive rows have no bins at all
                }
                __syncwarp();
                if (paired) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int src = wbase + (lane >> 1) + 16 * h;
                        if (s_ids[src] >= 0) {
                            const double v = word ? 1.0 : s_jf2[src];
                            for (int dl = 0; dl < gw; ++dl)
                                atomicAdd(pair_base + ((int64_t)(g0 + dl) * ni + s_ids[dl * UF_BLOCK + src]) * pair_stride + word, v);
                        }
                    }
                } else if (active && recs) {  // fp32 records: {float weight, u32 count}
                    for (int dl = 0; dl < gw; ++dl) {
                        MapRecord<T>* r = &recs[(int64_t)(g0 + dl) * ni + s_ids[dl * UF_BLOCK + threadIdx.x]];
                        atomicAdd(&r->w, jf2);
                        atomicAdd(&r->c, (decltype(r->c))1);
                    }
                } else if (active) {  // weights / counts arrays (small passes: what the one-launch cluster update reads)
                    for (int dl = 0; dl < gw; ++dl) {
                        const int64_t b = (int64_t)(g0 + dl) * ni + s_ids[dl * UF_BLOCK + threadIdx.x];
                        atomicAdd(&weights[b], jf2);
                        atomicAdd(&counts[b], 1ull);
                    }
                }
                __syncwarp();
            }
            if (STRAT) c_lo += rc.cube[buf][(int)(re - rb) - 1];
        }
    }
}

static void chunking(int64_t nrows, int64_t& ctas, int64_t& rows_per_cta) {
    int64_t tiles = (nrows + UF_BLOCK - 1) / UF_BLOCK;
    if (tiles < 1) tiles = 1;
    const int64_t cap = (int64_t)num_sms() * UF_MIN_CTAS;
    ctas = tiles < cap ? tiles : cap;
    rows_per_cta = (nrows + ctas - 1) / ctas;
    rows_per_cta = ((rows_per_cta + UF_BLOCK - 1) / UF_BLOCK) * UF_BLOCK;
    ctas = (nrows + rows_per_cta - 1) / rows_per_cta;
    if (ctas < 1) ctas = 1;
}

}  // namespace tq

using namespace tq;

extern "C" {

int tq_vegas_sample_map(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim, int32_t dtype, int64_t row_begin,
                        int64_t row_end, const void* edges_packed, int32_t edges_layout, int64_t n_intervals, const void* domain,
                        uint64_t seed, uint32_t call_idx, void* x, void* jac, void* stream) {
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_vegas_sample_map: bad row range");
    TQ_REQUIRE(dim >= 1 && dim <= 1024 && n_intervals >= 1 && n_intervals < (1LL << 31), "tq_vegas_sample_map: bad map shape (dim <= 1024)");
    TQ_REQUIRE(edges_packed && domain && x && jac, "tq_vegas_sample_map: NULL argument");
    TQ_REQUIRE(edges_layout == TQ_EDGES_PAIRS || edges_layout == TQ_EDGES_RECORDS, "tq_vegas_sample_map: unknown edges layout %d", edges_layout);
    const bool strat = offsets != nullptr;
    TQ_REQUIRE(!strat || (n_cubes >= 1 && n_cubes < (1LL << 31) && n_strat >= 1), "tq_vegas_sample_map: bad stratification sizes");
    const int64_t nrows = row_end - row_begin;
    if (nrows == 0) return TQ_OK;
    int64_t ctas, rows_per_cta;
    chunking(nrows, ctas, rows_per_cta);
    FastDiv ns_div;
    ns_div.set((uint32_t)(strat ? n_strat : 1));
    cudaStream_t st = as_stream(stream);
    const int ep_stride = edges_layout == TQ_EDGES_RECORDS ? 2 : 1;
    TQ_DISPATCH_DTYPE(dtype, {
        using P2 = typename EdgePair<T>::type;
        const T inv_ns = (T)1 / (T)(strat ? n_strat : 1);
        const size_t smem = 2 * (size_t)dim * sizeof(T);
        if (strat)
            sample_map_kernel<T, true><<<TQ_GRID((unsigned)ctas), UF_BLOCK, smem, st>>>(
                (const long long*)offsets, n_cubes, ns_div, inv_ns, row_begin, row_end, rows_per_cta, (const P2*)edges_packed, ep_stride,
                (int)n_intervals, dim, (const T*)domain, (T*)x, (T*)jac, seed, call_idx);
        else
            sample_map_kernel<T, false><<<TQ_GRID((unsigned)ctas), UF_BLOCK, smem, st>>>(
                nullptr, 0, ns_div, inv_ns, row_begin, row_end, rows_per_cta, (const P2*)edges_packed, ep_stride, (int)n_intervals, dim,
                (const T*)domain, (T*)x, (T*)jac, seed, call_idx);
    });
    return check_launch("sample_map_kernel");
}

int tq_vegas_accumulate_regen(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim, int32_t dtype,
                              int64_t row_begin, int64_t row_end, int64_t n_intervals, const void* f, const void* jac,
                              double volume, void* jf_out, void* jf2_out, void* hist_pairs, void* records, void* weights,
                              int64_t* counts, uint64_t seed, uint32_t call_idx, void* stream) {
    TQ_REQUIRE(row_end >= row_begin && row_begin >= 0, "tq_vegas_accumulate_regen: bad row range");
    TQ_REQUIRE(dim >= 1 && n_intervals >= 1 && n_intervals < (1LL << 31), "tq_vegas_accumulate_regen: bad map shape");
    TQ_REQUIRE(f && jac, "tq_vegas_accumulate_regen: NULL argument");
    TQ_REQUIRE((hist_pairs != nullptr) + (records != nullptr) + (weights != nullptr) <= 1,
               "tq_vegas_accumulate_regen: pass at most one histogram target (hist_pairs, records or weights + counts)");
    TQ_REQUIRE((weights == nullptr) == (counts == nullptr), "tq_vegas_accumulate_regen: weights and counts go together");
    TQ_REQUIRE(jf_out || jf2_out || hist_pairs || records || weights, "tq_vegas_accumulate_regen: nothing to do");
    const bool strat = offsets != nullptr;
    TQ_REQUIRE(!strat || (n_cubes >= 1 && n_cubes < (1LL << 31) && n_strat >= 1), "tq_vegas_accumulate_regen: bad stratification sizes");
    const int64_t nrows = row_end - row_begin;
    if (nrows == 0) return TQ_OK;
    int64_t ctas, rows_per_cta;
    chunking(nrows, ctas, rows_per_cta);
    FastDiv ns_div;
    ns_div.set((uint32_t)(strat ? n_strat : 1));
    cudaStream_t st = as_stream(stream);
    TQ_DISPATCH_DTYPE(dtype, {
        const T inv_ns = (T)1 / (T)(strat ? n_strat : 1);
        if (strat)
            accumulate_regen_kernel<T, true><<<TQ_GRID((unsigned)ctas), UF_BLOCK, 0, st>>>(
                (const long long*)offsets, n_cubes, ns_div, inv_ns, row_begin, row_end, rows_per_cta, (int)n_intervals, dim, (const T*)f,
                (const T*)jac, (T)volume, (T*)jf_out, (T*)jf2_out, (double*)hist_pairs, (MapRecord<T>*)records, (T*)weights,
                (unsigned long long*)counts, seed, call_idx);
        else
            accumulate_regen_kernel<T, false><<<TQ_GRID((unsigned)ctas), UF_BLOCK, 0, st>>>(
                nullptr, 0, ns_div, inv_ns, row_begin, row_end, rows_per_cta, (int)n_intervals, dim, (const T*)f, (const T*)jac,
                (T)volume, (T*)jf_out, (T*)jf2_out, (double*)hist_pairs, (MapRecord<T>*)records, (T*)weights,
                (unsigned long long*)counts, seed, call_idx);
    });
    return check_launch("accumulate_regen_kernel");
}

}  // extern "C"
