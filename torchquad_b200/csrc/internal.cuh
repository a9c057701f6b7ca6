// Launch helpers shared between translation units: the public entry points of include/tqb200.h are thin
// wrappers over these; the native VEGAS loop (vegas_driver.cu) uses the extra options.
#pragma once
#include "common.cuh"

namespace tq {

// ---- vegas_strat.cu / vegas_map.cu
int strat_nh_launch(const void* dh, int64_t n_cubes, double nevals_exp, int32_t dtype, int64_t* nh, int64_t* offsets,
                    void* clear, size_t clear_bytes, void* ws, size_t ws_bytes, void* stream);

// estimator partial sums + unnormalised d^beta (first half of tq_vegas_strat_update): scalars[0..4) = this rank's
// {sum ih, sum sig2/n, sum d^beta, sum nh}; strat_normalise_launch divides dh by scalars[2] once those are global.
int strat_update_partial_launch(const void* JF, const void* JF2, const int64_t* nh, int64_t n_cubes, double v_cubes, double beta,
                                int32_t dtype, void* dh, double* scalars, void* ws, size_t ws_bytes, void* stream);
int strat_normalise_launch(void* dh, int64_t n_cubes, const double* scalars, int32_t dtype, void* stream);
// fused.cu: the one launcher behind tq_fused_vegas / _sharded / _deferred (jf2_rows != NULL: HIST_DEFER)
int fused_vegas_launch(const tq_integrand* fn_host, int32_t dtype, const int64_t* offsets, int64_t n_cubes,
                       int32_t n_strat, int64_t row_begin, int64_t row_end, const void* edges_packed,
                       int32_t edges_layout, int64_t n_intervals, void* weights, int64_t* counts, void* hist_pairs,
                       void* jf2_rows, void* JF, void* JF2, uint64_t seed, uint32_t call_idx, int32_t cube_block_log2,
                       int32_t rank, int32_t world, double* out_f64, void* ws, size_t ws_bytes, void* stream);
// fused.cu: tq_vegas_hist_sweep for one rank's share of the cubes (n_cubes = GLOBAL count, offsets / jf2_rows local)
int hist_sweep_launch(const int64_t* offsets, int64_t n_cubes, int32_t n_strat, int32_t dim, int32_t dtype, const void* jf2_rows,
                      int64_t n_intervals, void* hist_pairs, int32_t dims_per_group, uint64_t seed, uint32_t call_idx,
                      int32_t cube_block_log2, int32_t rank, int32_t world, void* ws, size_t ws_bytes, void* stream);
// fused.cu: pairs += {record.w, record.c}, record fields back to zero (records -> the all-reduce buffer)
int records_to_pairs_launch(void* records, double* pairs, int32_t dim, int64_t ni, int32_t dtype, void* stream);

int map_update_launch(void* x_edges, void* dx_edges, void* weights, int64_t* counts, void* edges_packed, int32_t dim,
                      int64_t n_intervals, double alpha, int32_t dtype, int32_t* status, bool clear_status, void* ws,
                      size_t ws_bytes, void* stream);

// Scratch of one map update, carved from the caller's buffer (tq_vegas_map_workspace_bytes).
struct MapScratch {
    void* avg;          // T[dim*ni]
    void* smoothed;     // T[dim*ni]
    void* x_new;        // T[dim*(ni+1)]
    double* S;          // [dim*ni] fp64 prefix sums
    double* tile_sums;  // [dim*ntiles]
    double* totals;     // [dim] row sums of avg
    double* totals2;    // [dim] row sums of smoothed
    int ntiles;
};
bool map_scratch_carve(void* ws, size_t ws_bytes, int dim, long long ni, int32_t dtype, bool need_edges, MapScratch& s);

// ---- vegas_small.cu: the launch-latency-bound regime, one thread-block cluster per job
struct SmallStrat {  // arguments of the stratification update (+ optional get_NH of the NEXT pass)
    const void* JF;
    const void* JF2;
    int64_t* nh;        // read (this pass); rewritten with the next pass's counts when next_nevals > 0
    int64_t n_cubes;
    double v_cubes, beta;
    void* dh;
    double* scalars;    // fp64[4]
    double next_nevals; // > 0: also run get_NH + scan for the next pass ...
    int64_t* offsets;   // ... into offsets[n_cubes + 1] ...
    void* clear;        // ... and zero `clear_bytes` (the JF/JF2 accumulators) for it
    size_t clear_bytes;
};
struct SmallMap {  // arguments of the map update / smoothing
    void* x_edges;
    void* dx_edges;
    void* weights;
    int64_t* counts;
    void* edges_packed;
    MapScratch scratch;
    int32_t dim;
    int64_t ni;
    double alpha;
    int32_t* status;
    bool do_edges;      // false: stop after the smoothed weights (tq_vegas_map_smooth)
};
bool small_strat_ok(int64_t n_cubes);
bool small_map_ok(int32_t dim, int64_t ni);
int small_nh_launch(const void* dh, int64_t n_cubes, double nevals_exp, int32_t dtype, int64_t* nh, int64_t* offsets,
                    void* clear, size_t clear_bytes, void* stream);
// Either argument may be NULL; with both, ONE launch runs the stratification update and the map update side by side.
int small_update_launch(const SmallStrat* strat, const SmallMap* map, int32_t dtype, void* stream);

}  // namespace tq
