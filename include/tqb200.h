/* tqb200 -- C-ABI of the B200-native sampling-and-reduction hot path of torchquad.
 *
 * The reference (esa/torchquad v0.5.0) has NO native/FFI interface: every step below is Python over
 * ATen reached through `autoray`.  Each entry point therefore cites the reference Python function
 * (file:line under /root/reference) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain C: pointers are DEVICE pointers unless the name ends in _host; sizes are int64_t.
 *   - `dtype` is TQ_F32 or TQ_F64 (the working dtype of the reference tensors); integer state is int64
 *     exactly as in the reference (`counts`, `nh`).
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and never
 *     synchronises the device.  No call allocates device memory: temporaries come from the caller's
 *     workspace `ws` (`ws_bytes` >= tq_workspace_bytes(), zero-initialised once by the caller).
 *   - return value: TQ_OK (0) or a negative TQ_ERR_* / positive cudaError_t; tq_last_error() gives text.
 *   - tensors are dense row-major; `[rows, dim]` means element (r, d) at r*dim + d.
 */
#ifndef TQB200_H
#define TQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TQ_API __attribute__((visibility("default")))
#else
#define TQ_API
#endif

#define TQ_OK 0
#define TQ_ERR_INVALID_ARGUMENT (-1)
#define TQ_ERR_WORKSPACE (-2)
#define TQ_ERR_UNSUPPORTED (-3)
#define TQ_ERR_CALLBACK (-4) /* an integrand callback returned non-zero */

#define TQ_F32 0
#define TQ_F64 1

#define TQ_MAX_DIM 32 /* integration dimensions supported by the fused kernels */

/* Newton-Cotes rules (trapezoid.py, simpson.py, boole.py) */
#define TQ_RULE_TRAPEZOID 0
#define TQ_RULE_SIMPSON 1
#define TQ_RULE_BOOLE 2

/* Built-in integrand families for the fused kernels.  Genz families: SURVEY 8(d); the last four are
 * the reference's test integrands (tests/integration_test_functions.py:146-325). */
#define TQ_F_GENZ_OSCILLATORY 0
#define TQ_F_GENZ_PRODUCT_PEAK 1
#define TQ_F_GENZ_CORNER_PEAK 2
#define TQ_F_GENZ_GAUSSIAN 3
#define TQ_F_GENZ_C0 4
#define TQ_F_GENZ_DISCONTINUOUS 5
#define TQ_F_SUM_SIN 6
#define TQ_F_SUM_EXP 7
#define TQ_F_PROD_COS 8
#define TQ_F_POLYNOMIAL 9
/* fp32 fast-math variants (MUFU.SIN/COS/EX2 via __sinf/__cosf/__expf, |abs error| <= ~5e-7 on the
 * reference test domains); identical to the precise family in fp64. */
#define TQ_F_SUM_SIN_FAST 10
#define TQ_F_SUM_EXP_FAST 11
#define TQ_F_PROD_COS_FAST 12
#define TQ_F_COUNT 13

/* Parameters of a built-in integrand, passed by value from the host.  `a`/`u` are the Genz difficulty
 * and shift vectors; `coeff[0..ncoeff)` the polynomial coefficients (same for each dimension).
 * The integrand is evaluated at x*size + start and multiplied by `scale` (VEGAS passes the domain
 * volume, vegas.py:104-112). */
typedef struct tq_integrand {
    int32_t family;
    int32_t dim;
    int32_t ncoeff;
    int32_t _pad;
    double a[TQ_MAX_DIM];
    double u[TQ_MAX_DIM];
    double coeff[8];
    double start[TQ_MAX_DIM];
    double size[TQ_MAX_DIM];
    double scale;
} tq_integrand;

/* layouts of the packed VEGAS map tables (tq_vegas_map_pack_edges / tq_vegas_map_pack_records) */
#define TQ_EDGES_PAIRS 0
#define TQ_EDGES_RECORDS 1

/* ---- library ---------------------------------------------------------------------------------- */
TQ_API const char* tq_last_error(void);
TQ_API int tq_version(void);
/* Kernels launched by this library in this process so far (bench.py reports the difference as gpu_launches). */
TQ_API uint64_t tq_kernel_launches(void);
TQ_API size_t tq_workspace_bytes(void);
/* sm count / compute capability of the current device */
TQ_API int tq_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- RNG: replaces RNG.uniform (torchquad/integration/rng.py:119-125) ----------------------------
 * out[r - row_begin, d] = U[0,1)(seed, call_idx, r, d), r in [row_begin, row_end): counter-based
 * Philox4x32-10, a pure function of its arguments, so ranks drawing disjoint row ranges of one call
 * reproduce the single-process stream exactly. */
TQ_API int tq_philox_uniform(void* out, int64_t row_begin, int64_t row_end, int32_t dim, int32_t dtype,
                      uint64_t seed, uint32_t call_idx, void* stream);

/* ---- Monte Carlo: replaces MonteCarlo.calculate_sample_points (monte_carlo.py:84-106) ------------
 * out[r, d] = u*(domain[d,1]-domain[d,0]) + domain[d,0] with u as above (mul then add, not fused). */
TQ_API int tq_mc_sample(void* out, const void* domain, int64_t row_begin, int64_t row_end, int32_t dim,
                 int32_t dtype, uint64_t seed, uint32_t call_idx, void* stream);
/* The same op for launches that are REPLAYED (CUDA-graph capture of a whole integrate() call, the counterpart
 * of the reference's get_jit_compiled_integrate, monte_carlo.py:108-225): the stream call index used is
 * call_idx + *call_offset_dev, read on the device, so a replay that follows an increment of that word draws
 * fresh samples without a new launch configuration. */
TQ_API int tq_mc_sample_replayable(void* out, const void* domain, int64_t row_begin, int64_t row_end, int32_t dim,
                            int32_t dtype, uint64_t seed, uint32_t call_idx, const uint32_t* call_offset_dev,
                            void* stream);
/* d(loss)/d(domain) for the op above: grad_domain[d,0] = sum_r g*(1-u), grad_domain[d,1] = sum_r g*u,
 * uniforms regenerated from the counter.  grad_domain_f64 is double[dim*2] (device). */
TQ_API int tq_mc_sample_backward(const void* grad_out, int64_t row_begin, int64_t row_end, int32_t dim,
                          int32_t dtype, uint64_t seed, uint32_t call_idx, double* grad_domain_f64,
                          void* ws, size_t ws_bytes, void* stream);
/* Column sums of f[rows, cols] accumulated in fp64: sum_f64[c] = sum_r f[r,c], and if sumsq_f64 is not
 * NULL sumsq_f64[c] = sum_r f[r,c]^2.  Replaces the `anp.sum(function_values, axis=0)` of
 * MonteCarlo.calculate_result (monte_carlo.py:72-77); the variance moment is an extension. */
TQ_API int tq_sum_columns(const void* f, int64_t rows, int64_t cols, int32_t dtype, double* sum_f64,
                   double* sumsq_f64, void* ws, size_t ws_bytes, void* stream);

/* ---- VEGAS map (torchquad/integration/vegas_map.py) ---------------------------------------------
 * forward: get_X (:44-58) + get_Jac (:60-74) + _get_interval_ID (:76-85) in one pass.
 *   k = floor(y*Ni); o = y*Ni - k; x = xe[d,k] + dxe[d,k]*o; jac = prod_d Ni*dxe[d,k] (left to right).
 *   x, jac, ids, offset may each be NULL.  ids is int32[rows, dim]; offset[rows, dim] is `o`
 *   (_get_interval_offset :87-97). */
TQ_API int tq_vegas_map_forward(const void* y, const void* x_edges, const void* dx_edges, void* x, void* jac,
                         int32_t* ids, void* offset, int64_t rows, int32_t dim, int64_t n_intervals,
                         int32_t dtype, void* stream);
/* Same pass with the edges packed as {x_edge, dx_edge} pairs (tq_vegas_map_pack_edges; one gather per element)
 * and, when `domain` ([dim,2], nullable) is given, the unit-cube -> domain transform x*size + start of
 * vegas.py:109-110 applied in the same kernel.  edges_layout = TQ_EDGES_RECORDS reads the pairs out of the
 * large-map record table (see tq_vegas_map_pack_records below). */
TQ_API int tq_vegas_map_forward_packed(const void* y, const void* edges_packed, int32_t edges_layout, const void* domain,
                                void* x, void* jac, int32_t* ids, int64_t rows, int32_t dim, int64_t n_intervals,
                                int32_t dtype, void* stream);
/* accumulate_weight (:99-111): weights[d,k] += jf2[r], counts[d,k] += 1 (int64, bit-exact). */
TQ_API int tq_vegas_map_accumulate(const void* y, const void* jf2, void* weights, int64_t* counts, int64_t rows,
                            int32_t dim, int64_t n_intervals, int32_t dtype, void* stream);
/* The tail of an unfused VEGAS pass in one kernel (vegas.py:104-112,284-290): jf = (f*volume)*jac,
 * weights[d,k] += jf^2, counts[d,k] += 1, and jf written to jf_out (nullable) for tq_vegas_strat_accumulate.
 * Large maps: pass `records` (and weights = counts = NULL) to accumulate into the record table instead;
 * weights = counts = records = NULL computes jf only (no grid improvement). */
TQ_API int tq_vegas_accumulate_fused(const void* y, const void* f, const void* jac, double volume, void* jf_out, void* weights,
                              int64_t* counts, void* records, int64_t rows, int32_t dim, int64_t n_intervals,
                              int32_t dtype, void* stream);
/* An unfused VEGAS pass as TWO kernels around the user's integrand (round 2; replaces tq_vegas_strat_sample ->
 * tq_vegas_map_forward_packed -> ... -> tq_vegas_accumulate_fused whenever the uniforms come from the Philox stream):
 *  tq_vegas_sample_map: get_Y (vegas_stratification.py:140-165) + get_X / get_Jac (vegas_map.py:44-74) + the
 *   x*size + start transform of vegas.py:109-110 for rows [row_begin,row_end): x[row - row_begin, :] and
 *   jac[row - row_begin]; the stratified y is never written.  offsets == NULL: a warm-up pass, y = u * 0.999999 from the
 *   row-keyed stream (vegas.py:230-233).  x, jac are bit-identical to the three-kernel pipeline.
 *  tq_vegas_accumulate_regen: jf = (f*volume)*jac -> jf_out (nullable), and VEGASMap.accumulate_weight
 *   (vegas_map.py:99-111) with the bin ids REGENERATED from the same Philox blocks (same arithmetic: identical bins)
 *   into `hist_pairs` (fp64 {sum jf^2, count} per bin, tq_vegas_map_unpack_hist), into the large-map `records`, or
 *   into `weights` / `counts` directly (two reductions per bin: small passes); at most one target, none: jf only
 *   (no grid improvement).  jf2_out (nullable) receives jf^2 per row: maps beyond L2 bin those rows afterwards, band by
 *   band, with tq_vegas_hist_sweep.  f and jac point at row `row_begin` (the rows of the matching tq_vegas_sample_map). */
TQ_API int tq_vegas_sample_map(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim, int32_t dtype,
                        int64_t row_begin, int64_t row_end, const void* edges_packed, int32_t edges_layout,
                        int64_t n_intervals, const void* domain, uint64_t seed, uint32_t call_idx, void* x, void* jac,
                        void* stream);
TQ_API int tq_vegas_accumulate_regen(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim, int32_t dtype,
                              int64_t row_begin, int64_t row_end, int64_t n_intervals, const void* f, const void* jac,
                              double volume, void* jf_out, void* jf2_out, void* hist_pairs, void* records,
                              void* weights, int64_t* counts, uint64_t seed, uint32_t call_idx, void* stream);
/* Scratch bytes tq_vegas_map_smooth / tq_vegas_map_update need for a [dim, Ni] map (pass as ws). */
TQ_API size_t tq_vegas_map_workspace_bytes(int32_t dim, int64_t n_intervals, int32_t dtype);
/* _smooth_map (:113-172) written to `smoothed[dim, Ni]`; status[0] = 1 when a dimension sums to zero
 * (the reference returns None).  weights/counts are not modified. */
TQ_API int tq_vegas_map_smooth(const void* weights, const int64_t* counts, void* smoothed, int32_t dim,
                        int64_t n_intervals, double alpha, int32_t dtype, int32_t* status, void* ws,
                        size_t ws_bytes, void* stream);
/* update_map (:185-261): smooth, equal-mass rebin (fp64 prefix sums), inf repair, dx = diff(x), reset
 * of weights/counts.  status (device int32[4]): [0]=1 update skipped (zero dimension), [1]=number of
 * non-finite edges that were repaired, [2]=1 unrepairable (reference raises RuntimeError), [3]=unused.
 * edges_packed (nullable): also receives the new edges as {x_edge, dx_edge} pairs (tq_vegas_map_pack_edges).
 * Maps of up to 32768 intervals per dimension run the whole update in a single launch. */
TQ_API int tq_vegas_map_update(void* x_edges, void* dx_edges, void* weights, int64_t* counts, void* edges_packed,
                        int32_t dim, int64_t n_intervals, double alpha, int32_t dtype, int32_t* status, void* ws,
                        size_t ws_bytes, void* stream);

/* ---- VEGAS stratification (torchquad/integration/vegas_stratification.py) -----------------------
 * get_NH (:92-103): nh[c] = max(2, floor(dh[c]*nevals_exp)) as int64, plus offsets[c] = sum_{c'<c} nh
 * (offsets has n_cubes+1 entries; offsets[n_cubes] = M, the number of samples of the iteration). */
TQ_API int tq_vegas_strat_nh(const void* dh, int64_t n_cubes, double nevals_exp, int32_t dtype, int64_t* nh,
                      int64_t* offsets, void* ws, size_t ws_bytes, void* stream);
/* offsets[c] = sum_{c'<c} nh[c'] (n_cubes+1 entries) for a caller-provided nh; the `anp.repeat(arange, nevals)`
 * of get_Y / accumulate_weight (:58-59,:154-155) is never materialised, rows find their cube by search. */
TQ_API int tq_vegas_strat_offsets(const int64_t* nh, int64_t n_cubes, int64_t* offsets, void* ws, size_t ws_bytes,
                           void* stream);
/* get_Y (:140-165) for sample rows [row_begin,row_end) of the cube-sorted order:
 *   y = (digit_d(cube) + u)/N_strat, digit_d(c) = (c / N_strat^d) % N_strat, y >= 1 -> 0.999999.
 * u_in != NULL: uniforms are read from u_in[row - row_begin, d] (injected RNG, tests/vegas_test.py:143-156);
 * u_in == NULL: cube-keyed Philox stream (seed, call_idx, cube, index within cube, d). */
TQ_API int tq_vegas_strat_sample(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim,
                          int32_t dtype, const void* u_in, uint64_t seed, uint32_t call_idx,
                          int64_t row_begin, int64_t row_end, void* y, void* stream);
/* accumulate_weight (:46-70) for cubes [cube_begin,cube_end): JF[c] = sum jf, JF2[c] = sum jf^2 over the
 * cube's rows in row order (bit-identical to the CPU scatter_add_).  jf points at row `row_base`. */
TQ_API int tq_vegas_strat_accumulate(const void* jf, int64_t row_base, const int64_t* offsets,
                              int64_t cube_begin, int64_t cube_end, void* JF, void* JF2, int32_t dtype,
                              void* stream);
/* Backward of JF wrt jf: grad_jf[r] = grad_JF[cube(r)] for rows [row_begin,row_end). */
TQ_API int tq_vegas_strat_accumulate_backward(const void* grad_JF, const int64_t* offsets, int64_t n_cubes,
                                       int64_t row_begin, int64_t row_end, void* grad_jf, int32_t dtype,
                                       void* stream);
/* Iteration estimator (vegas.py:293-303) + update_DH (vegas_stratification.py:72-90) in one pass:
 *   scalars_f64[0] = I_it = sum JF*V/n, [1] = sigma2_it = sum |JF2*V^2/n - ih^2|/n, [2] = sum d^beta,
 *   [3] = sum nh (the pass's sample count, exact below 2^53) -- scalars_f64 holds FOUR doubles;
 *   dh[c] = d^beta / sum (left unnormalised when the sum is 0). */
TQ_API int tq_vegas_strat_update(const void* JF, const void* JF2, const int64_t* nh, int64_t n_cubes,
                          double v_cubes, double beta, int32_t dtype, void* dh, double* scalars_f64,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- Newton-Cotes grids (integration_grid.py:64-99, grid_integrator.py:57-91, rule files) -------
 * points[p - p_begin, d] = nodes[d, i_d(p)], i_d(p) = (p / n^(dim-1-d)) % n  (dim 0 slowest). */
TQ_API int tq_nc_grid_points(const void* nodes, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                      void* points, int32_t dtype, void* stream);
/* Backward: grad_nodes_f64[d, j] = sum over rows with i_d(p) = j of grad_points[p, d]. */
TQ_API int tq_nc_grid_points_backward(const void* grad_points, int32_t n, int32_t dim, int64_t p_begin,
                               int64_t p_end, double* grad_nodes_f64, int32_t dtype, void* stream);
/* out_f64[k] = sum_{p in [p_begin,p_end)} f[p - p_begin, k] * prod_d w[d, i_d(p)]: the tensor-product
 * weight contraction that `_apply_composite_rule` evaluates axis by axis. */
TQ_API int tq_nc_contract(const void* f, const void* w, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                   int64_t cols, int32_t dtype, double* out_f64, void* ws, size_t ws_bytes, void* stream);
/* W[p - p_begin] = prod_d w[d, i_d(p)] (used by the backward of tq_nc_contract). */
TQ_API int tq_nc_point_weights(const void* w, int32_t n, int32_t dim, int64_t p_begin, int64_t p_end,
                        void* out, int32_t dtype, void* stream);

/* ---- fused paths for built-in integrands (no sample traffic to HBM) ------------------------------
 * Monte Carlo over rows [row_begin,row_end) of the row-keyed stream: out_f64 = {sum f, sum f^2}. */
TQ_API int tq_fused_mc(const tq_integrand* fn_host, int32_t dtype, int64_t row_begin, int64_t row_end,
                uint64_t seed, uint32_t call_idx, double* out_f64, void* ws, size_t ws_bytes,
                void* stream);
/* Newton-Cotes: out_f64[0] = sum_p f(nodes[., i(p)]) * prod_d w[d, i_d(p)]. */
TQ_API int tq_fused_nc(const tq_integrand* fn_host, const void* nodes, const void* w, int32_t n, int32_t dtype,
                int64_t p_begin, int64_t p_end, double* out_f64, void* ws, size_t ws_bytes, void* stream);
/* {x_edges[d,k], dx_edges[d,k]} interleaved as pairs [dim, Ni] (float2 / double2): the layout the fused
 * kernel gathers from, one 8/16-byte load per dimension instead of two. */
TQ_API int tq_vegas_map_pack_edges(const void* x_edges, const void* dx_edges, void* edges_packed, int32_t dim,
                            int64_t n_intervals, int32_t dtype, void* stream);
/* One VEGAS pass without writing samples: generate -> map -> evaluate -> accumulate.
 *   stratified (offsets != NULL): rows [row_begin,row_end) of the cube-sorted order (vegas.py:268-291);
 *     JF/JF2 (pre-zeroed by the caller) receive the per-cube sums.  row_begin == 0 and row_end < 0 selects the
 *     whole pass [0, offsets[n_cubes]) with the count read ON THE DEVICE (no host read-back of get_NH's total);
 *     -row_end is then the caller's estimate of that count and only sizes the grid.
 *   warm-up (offsets == NULL): rows are plain samples y = u*0.999999 (vegas.py:236); out_f64 receives
 *     {sum jf, sum jf^2}.
 *   edges_packed: see tq_vegas_map_pack_edges (edges_layout = TQ_EDGES_PAIRS) or the record layout below.
 *   Map histogram (vegas_map.py:99-111), one of:
 *     weights/counts != NULL   accumulated directly (two L2 reductions per sample and dimension; small passes);
 *     hist_pairs != NULL       fp64 [dim, Ni, 2] = {sum jf^2, count} per bin (both dtypes; the count is an exact fp64
 *                              integer): ONE reduction sector per sample and dimension.  Zero it once; fold it into
 *                              weights/counts with tq_vegas_map_unpack_hist before tq_vegas_map_update;
 *     all NULL                 no grid improvement (or TQ_EDGES_RECORDS: the histogram lives in the records). */
TQ_API int tq_fused_vegas(const tq_integrand* fn_host, int32_t dtype, const int64_t* offsets, int64_t n_cubes,
                   int32_t n_strat, int64_t row_begin, int64_t row_end, const void* edges_packed,
                   int32_t edges_layout, int64_t n_intervals, void* weights, int64_t* counts, void* hist_pairs,
                   void* JF, void* JF2, uint64_t seed, uint32_t call_idx, double* out_f64, void* ws,
                   size_t ws_bytes, void* stream);
/* The same pass on one rank of a multi-GPU run.  Cubes are dealt to the ranks in blocks of B = 2^cube_block_log2 cubes:
 * round b (world consecutive blocks) gives every rank one block, rank r the one at position (r + skew(b)) % world with
 * skew(b) = b + (b>>3) + (b>>6) + ... + (b>>18) (a rotation, so that no rank owns a fixed digit of the cube index).
 * offsets/JF/JF2 index this rank's cubes only (n_cubes = how many it owns); local cube l = global cube
 * ((l / B * world + (rank + skew(l / B)) % world) * B + l % B, which keys its Philox stream and gives its position in the
 * unit cube -- so every world size draws exactly the samples of the single-GPU run.
 * Warm-up passes (offsets == NULL) are sharded by the caller through [row_begin, row_end). */
TQ_API int tq_fused_vegas_sharded(const tq_integrand* fn_host, int32_t dtype, const int64_t* offsets, int64_t n_cubes,
                           int32_t n_strat, int64_t row_begin, int64_t row_end, const void* edges_packed,
                           int32_t edges_layout, int64_t n_intervals, void* weights, int64_t* counts, void* hist_pairs,
                           void* JF, void* JF2, uint64_t seed, uint32_t call_idx, int32_t cube_block_log2, int32_t rank,
                           int32_t world, double* out_f64, void* ws, size_t ws_bytes, void* stream);
/* Maps BEYOND L2 (the reference's Ni = N_increment / 10 of vegas.py:117: 8 x 1e7 bins at N = 2.5e9): a stratified pass in two
 * steps.  tq_fused_vegas_deferred runs the pass without the histogram (edges gathered from the pair table) and stores
 * jf^2 of every row in jf2_rows[row - row_begin] (working dtype).  tq_vegas_hist_sweep then accumulates
 * weights[d,k] += jf^2, counts[d,k] += 1 (vegas_map.py:99-111) into hist_pairs band by band: a sample of cube c can only
 * fall into the Ni / N_strat bins selected by digit d of c (vegas_stratification.py:140-165), so walking the cubes in the
 * order of those digits keeps the touched part of the table (dims_per_group * Ni / N_strat * 16 bytes) resident in L2 and
 * every table line is written to HBM once per band, instead of one random HBM read-modify-write per sample and dimension.
 * dims_per_group <= 2 (fp64) / 4 (fp32) dimensions share one launch (they share a Philox block).  Same seed / call_idx /
 * offsets as the pass: the uniforms are regenerated, the bins are identical, counts are exact. */
TQ_API int tq_fused_vegas_deferred(const tq_integrand* fn_host, int32_t dtype, const int64_t* offsets, int64_t n_cubes,
                            int32_t n_strat, int64_t row_begin, int64_t row_end, const void* edges_pairs, int64_t n_intervals,
                            void* jf2_rows, void* JF, void* JF2, uint64_t seed, uint32_t call_idx, void* ws, size_t ws_bytes,
                            void* stream);
TQ_API int tq_vegas_hist_sweep(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim, int32_t dtype,
                        const void* jf2_rows, int64_t n_intervals, void* hist_pairs, int32_t dims_per_group, uint64_t seed,
                        uint32_t call_idx, void* ws, size_t ws_bytes, void* stream);
/* weights += hist.sum (rounded once to the working dtype), counts += hist.count, hist = 0 (vegas_map.py:99-111). */
TQ_API int tq_vegas_map_unpack_hist(void* hist_pairs, void* weights, int64_t* counts, int32_t dim, int64_t n_intervals,
                             int32_t dtype, void* stream);

/* Record layout for LARGE maps (tables beyond L2, e.g. the reference's Ni = N_increment/10 of vegas.py:117 at
 * N = 2.5e9: 8 x 1e7 bins).  One bin = one record {x_edge, dx_edge, weight, count}: 32 bytes {f64,f64,f64,f64}
 * for TQ_F64 (the count is an exact fp64 integer), 16 bytes {f32,f32,f32,u32} for TQ_F32, [dim, Ni] records.  With edges_layout = TQ_EDGES_RECORDS
 * tq_fused_vegas gathers the edges from the records and accumulates the histogram (vegas_map.py:99-111) INTO
 * them (weights = counts = NULL): the three scattered accesses per sample and dimension fall into one DRAM
 * sector.  tq_vegas_map_unpack_records then adds the record fields to weights/counts (the arrays
 * tq_vegas_map_update reads) and zeroes them; tq_vegas_map_pack_records rewrites the records from new edges. */
TQ_API size_t tq_vegas_map_records_bytes(int32_t dim, int64_t n_intervals, int32_t dtype);
TQ_API int tq_vegas_map_pack_records(const void* x_edges, const void* dx_edges, void* records, int32_t dim,
                              int64_t n_intervals, int32_t dtype, void* stream);
TQ_API int tq_vegas_map_unpack_records(void* records, void* weights, int64_t* counts, int32_t dim,
                                int64_t n_intervals, int32_t dtype, void* stream);

/* ---- whole fused VEGAS run in one call (host loop in C++) ----------------------------------------
 * Runs VEGAS.integrate's warm-up, iterations and chi^2 / budget schedule (vegas.py:137-209,211-315) for a
 * built-in integrand, calling the step kernels above; sample-for-sample identical to driving them one by
 * one.  All device buffers are owned by the caller (no allocation); records/status need
 * TQ_VEGAS_MAX_PASSES*4 entries.  JF and JF2 must be contiguous ([2, n_cubes], JF2 = JF + n_cubes). */
#define TQ_VEGAS_MAX_PASSES 128
typedef struct tq_vegas_state {
    void* x_edges;      /* [dim, Ni+1] */
    void* dx_edges;     /* [dim, Ni] */
    void* edges_packed; /* initial map, already packed: pairs [dim, Ni, 2] or records (edges_layout) */
    void* weights;      /* [dim, Ni], zero */
    int64_t* counts;    /* [dim, Ni], zero */
    void* hist_pairs;   /* fp64 [dim, Ni, 2], zero, or NULL: see tq_fused_vegas (used for passes of >= 2^20 rows) */
    void* jf2_rows;     /* working dtype [jf2_rows_cap] or NULL.  Non-NULL (with hist_pairs, TQ_EDGES_PAIRS): stratified passes run
                           deferred + tq_vegas_hist_sweep (maps beyond L2); cap >= 4 * (N / (max_iterations + 5)) + 2 * n_cubes */
    int64_t jf2_rows_cap;
    int32_t sweep_dims_per_group; /* dims_per_group of tq_vegas_hist_sweep */
    int32_t _pad0;
    void* dh;           /* [n_cubes], initial 1/n_cubes */
    int64_t* nh;        /* [n_cubes] */
    int64_t* offsets;   /* [n_cubes + 1] */
    void* JF;           /* [2, n_cubes] */
    void* JF2;
    double* records;    /* [TQ_VEGAS_MAX_PASSES * 4]: per iteration I, sigma^2, sum d^beta, sum nh */
    int32_t* status;    /* [TQ_VEGAS_MAX_PASSES * 4]: per map update, see tq_vegas_map_update */
    void* map_ws;       /* tq_vegas_map_workspace_bytes(dim, Ni, dtype) */
    size_t map_ws_bytes;
    void* ws;           /* tq_workspace_bytes(), zero-initialised */
    size_t ws_bytes;
    int32_t edges_layout; /* TQ_EDGES_PAIRS or TQ_EDGES_RECORDS (large maps) */
} tq_vegas_state;
typedef struct tq_vegas_result {
    int32_t it;          /* iterations performed (VEGAS.it) */
    int32_t n_block;     /* entries of results / sigma2: the last block of up to 5 iterations */
    int64_t fevals;      /* VEGAS._nr_of_fevals */
    int64_t starting_N;  /* per-iteration budget when the run stopped */
    int32_t calls_used;  /* next free call index of the Philox stream */
    int32_t n_passes;    /* map updates performed = valid status words */
    double results[8];
    double sigma2[8];
    int32_t status[TQ_VEGAS_MAX_PASSES * 4];
} tq_vegas_result;
TQ_API int tq_vegas_run_fused(const tq_integrand* fn_host, int32_t dtype, int64_t N, int32_t max_iterations,
                       double eps_rel, double eps_abs, int32_t use_grid_improve, int32_t use_warmup,
                       int64_t n_intervals, int32_t n_strat, int64_t n_cubes, double v_cubes, double alpha,
                       double beta, uint64_t seed, uint32_t first_call, const tq_vegas_state* state,
                       tq_vegas_result* result_host, void* stream);

/* ---- the same run on one rank of a multi-GPU job (one process per GPU) -----------------------------------
 * Hypercubes are dealt to the ranks block-cyclically (tq_fused_vegas_sharded): `state`'s dh / nh / offsets / JF / JF2 hold
 * this rank's `n_cubes_local` cubes only (dh initialised to 1 / n_cubes, the GLOBAL count passed as n_cubes); the map is
 * replicated.  Per pass the library calls `allreduce(user, offset, count)` ONCE: sum `comm[offset, offset + count)` (fp64,
 * device memory) over all ranks, in place, stream-ordered on `stream` (NCCL all-reduce in the Python host,
 * torchquad_b200/ops.py).  comm = [{sum jf^2, count} pairs of the map histogram, dim * Ni * 2 | 8 scalars]; it must be zero
 * on entry.  Every rank receives the same summed statistics, takes the same schedule decisions and returns the same result. */
typedef int (*tq_allreduce_callback)(void* user, int64_t offset, int64_t count);
typedef struct tq_vegas_shard {
    int32_t rank, world;
    int32_t cube_block_log2;   /* cubes are dealt in blocks of 2^cube_block_log2 */
    int32_t _pad;
    int64_t n_cubes_local;     /* cubes this rank owns */
    double* comm;              /* device fp64 [dim * Ni * 2 + 8], zero */
    tq_allreduce_callback allreduce;
    void* user;
} tq_vegas_shard;
TQ_API int tq_vegas_run_fused_sharded(const tq_integrand* fn_host, int32_t dtype, int64_t N, int32_t max_iterations,
                               double eps_rel, double eps_abs, int32_t use_grid_improve, int32_t use_warmup,
                               int64_t n_intervals, int32_t n_strat, int64_t n_cubes, double v_cubes, double alpha,
                               double beta, uint64_t seed, uint32_t first_call, const tq_vegas_state* state,
                               const tq_vegas_shard* shard, tq_vegas_result* result_host, void* stream);

/* ---- whole VEGAS run with a CALLBACK integrand (the drop-in path for arbitrary Python callables) -------
 * Same loop and schedule as tq_vegas_run_fused, but every pass materialises x: tq_vegas_sample_map (stratified y ->
 * x, jac with the unit-cube -> domain transform, y itself stays in registers) ->
 * `eval(user, rows, &f)` -> jf, histogram with regenerated bins (tq_vegas_accumulate_regen) -> per-cube sums
 * (tq_vegas_strat_accumulate) -> updates.  The callback evaluates the integrand on the first `rows` rows of
 * buffers->x on the SAME stream and stores the device pointer of its `rows` values (working dtype) in *f;
 * non-zero return aborts the run with TQ_ERR_CALLBACK.  One 8-byte read-back per pass (rows sizes the
 * callback's view).  Buffers hold `cap_rows` rows; a pass that needs more fails with TQ_ERR_WORKSPACE. */
typedef int (*tq_eval_callback)(void* user, int64_t rows, const void** f);
typedef struct tq_vegas_unfused_buffers {
    void* y;            /* unused since round 2 (the samples y are never materialised); may be NULL */
    void* x;            /* [cap_rows, dim] */
    void* jac;          /* [cap_rows] */
    void* jf;           /* [cap_rows] */
    const void* domain; /* [dim, 2] integration domain (device) */
    const void* warm_domain; /* [dim, 2] rows {0, 0.999999}: the warm-up's y = u * 0.999999 (vegas.py:236) */
    int64_t cap_rows;
    double volume;      /* prod(domain sizes) in the working precision */
} tq_vegas_unfused_buffers;
TQ_API int tq_vegas_run_unfused(tq_eval_callback eval, void* user, int32_t dim, int32_t dtype, int64_t N,
                         int32_t max_iterations, double eps_rel, double eps_abs, int32_t use_grid_improve,
                         int32_t use_warmup, int64_t n_intervals, int32_t n_strat, int64_t n_cubes, double v_cubes,
                         double alpha, double beta, uint64_t seed, uint32_t first_call, const tq_vegas_state* state,
                         const tq_vegas_unfused_buffers* buffers, tq_vegas_result* result_host, void* stream);

/* The schedule decision tq_vegas_run_fused takes after every fifth iteration, as a host-only function (no GPU
 * work; tested against the reference's tensor arithmetic): VEGAS._check_abort_conditions + the weighted mean /
 * error / chi^2 of vegas.py:161-209,318-362 on a block of n_block <= 5 (result, sigma^2) pairs, evaluated in the
 * working precision `dtype`.  *stop_out = 1 to stop; otherwise *starting_N_inout holds the next per-iteration
 * budget.  *mean_out = the block's weighted mean (VEGAS._get_result). */
TQ_API int tq_vegas_schedule(const double* results_host, const double* sigma2_host, int32_t n_block, int32_t dtype,
                      double eps_rel, double eps_abs, int64_t N, int64_t fevals, int32_t it, int32_t max_iterations,
                      int64_t increment, int64_t* starting_N_inout, double* mean_out, int32_t* stop_out);

/* Get (and, when bytes > 0, set) the device's L2 fetch granularity hint (cudaLimitMaxL2FetchGranularity:
 * 32, 64 or 128).  The VEGAS kernels gather map edges and update histogram bins at random positions of
 * tables that exceed L2 when the reference's map size formula is used (Ni = N/250 per dimension); with
 * the default 128-byte granularity every 8/16-byte access moves 128 bytes of HBM. */
TQ_API int tq_l2_fetch_granularity(int32_t bytes, int32_t* previous_host);

/* Peak-rate microbenchmarks used by bench.py for the fused-path roofline denominators:
 * kind 0 = dependent-free FP32 FMA chains, 1 = FP64 FMA chains, 2 = Philox4x32-10 blocks.
 * Performs iters*threads*ops_per_iter operations; returns ops per launch through ops_out_host. */
TQ_API int tq_peak_microbench(int32_t kind, int64_t iters, double* sink, double* ops_out_host, void* stream);
/* The reduction pattern of the fused VEGAS pass alone: every thread adds {1.5, 1.0} to the two fp64 words of a random bin
 * of table[bins][2] per iteration (lane pairs of one RED.F64 share a bin = one 32-byte sector).  ops_out_host = reduction
 * sectors issued by the launch; bench.py times it for the L2-reduction roofline of the L2-resident VEGAS workloads. */
TQ_API int tq_red_microbench(double* table, int64_t bins, int64_t iters, double* ops_out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TQB200_H */
