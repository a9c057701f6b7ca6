// Host-side VEGAS+ loops: one C-ABI call runs warm-up, all iterations and the chi^2 / budget schedule of
// torchquad/integration/vegas.py:137-209,211-315 without returning to Python between passes.  Small problems
// (the reference's own N = 1e6 configuration is ~0.4 ms of GPU work) are bound by per-launch host overhead.
// The kernels are exactly those of the step-by-step API, so a run draws the same samples as the Python-driven
// loop (tests/test_gpu_integrators.py).
//   tq_vegas_run_fused    built-in integrands: nothing is read back inside a block of five iterations -- the
//                         fused pass takes its sample count from get_NH's offsets on the device, the
//                         stratification update records it next to the iteration estimate, and the schedule
//                         reads the block's records in one copy.
//   tq_vegas_run_unfused  any integrand through a per-pass callback: samples are materialised, the callback
//                         evaluates them, one 8-byte read-back per pass sizes its view.
//   tq_vegas_schedule     the checkpoint decision alone (host only; tested against the oracle on CPU).
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "internal.cuh"

namespace tq {

static inline float tsqrt(float x) { return sqrtf(x); }
static inline double tsqrt(double x) { return sqrt(x); }

// vegas.py:318-362 in the working precision T (the reference does this arithmetic on 0-dim tensors of dtype T).
template <typename T>
struct Block {
    std::vector<T> res, sig;
    T mean() const {
        bool zero = false;
        for (T s : sig) zero |= (s == (T)0);
        if (zero) {
            T acc = (T)0;
            for (T r : res) acc += r;
            return acc / (T)res.size();
        }
        T num = (T)0, den = (T)0;
        for (size_t k = 0; k < res.size(); ++k) {
            num += res[k] / sig[k];
            den += (T)1 / sig[k];
        }
        return num / den;
    }
    T error() const {
        T inv = (T)0;
        for (T s : sig)
            if (s != (T)0) inv += (T)1 / s;
        return inv == (T)0 ? sig[0] : (T)1 / tsqrt(inv);
    }
    T chisq(T m) const {
        T acc = res[0] * (T)0;
        for (size_t k = 0; k < res.size(); ++k)
            if (res[k] != m) acc += (res[k] - m) * (res[k] - m) / sig[k];
        return acc;
    }
};

// vegas.py:161-209 at an iteration with it % 5 == 0: true = stop; otherwise `starting` holds the next budget.
template <typename T>
static bool schedule_checkpoint(const Block<T>& blk, double eps_rel, double eps_abs, int64_t N, int64_t fevals, int it,
                                int max_it, int64_t increment, int64_t& starting) {
    const T mean = blk.mean();
    const T res_abs = (T)fabs((double)mean);
    const T err = blk.error();
    const T chi2 = blk.chisq(mean);
    if ((err <= (T)eps_rel * res_abs || err <= (T)eps_abs) && chi2 / (T)5 < (T)1) return true;
    if (chi2 / (T)5 < (T)1) {
        if (res_abs == (T)0) {
            starting += increment;
        } else {
            const T acc = err / res_abs;
            const T scaled = (T)starting * tsqrt((T)(acc / (T)(eps_rel + 1e-8)));
            const int64_t alt = isfinite((double)scaled) && (double)scaled < 9.0e18 ? (int64_t)scaled : INT64_MAX;
            starting = starting + increment < alt ? starting + increment : alt;
        }
    } else if (chi2 / (T)5 > (T)1) {
        starting += increment;
    }
    if (fevals + starting * 5 > N) return true;
    return it + 5 > max_it;
}

template <typename T>
static int run_fused_sharded(const tq_integrand* fn, int32_t dtype, int64_t N, int32_t max_it, double eps_rel, double eps_abs,
                             bool grid_improve, bool warmup, int64_t ni, int32_t n_strat, int64_t n_cubes, double v_cubes,
                             double alpha, double beta, uint64_t seed, uint32_t call, const tq_vegas_state* s,
                             const tq_vegas_shard* sh, tq_vegas_result* out, void* stream);

template <typename T>
static int run_fused(const tq_integrand* fn, int32_t dtype, int64_t N, int32_t max_it, double eps_rel, double eps_abs,
                     bool grid_improve, bool warmup, int64_t ni, int32_t n_strat, int64_t n_cubes, double v_cubes,
                     double alpha, double beta, uint64_t seed, uint32_t call, const tq_vegas_state* s,
                     tq_vegas_result* out, void* stream) {
    cudaStream_t st = as_stream(stream);
    const int dim = fn->dim;
    const size_t elt = sizeof(T);
    const int64_t increment = N / (max_it + 5);  // vegas.py:90-91
    int64_t starting = increment;
    int64_t fevals = 0;
    int passes = 0;
    const int max_passes = TQ_VEGAS_MAX_PASSES;
    // Large maps keep {x, dx, weight, count} records (tq_fused_vegas, TQ_EDGES_RECORDS): the histogram is moved to
    // the weights/counts arrays before an update and the records are rewritten from the new edges after it.
    const bool recs = s->edges_layout == TQ_EDGES_RECORDS;
    const int layout = recs ? TQ_EDGES_RECORDS : TQ_EDGES_PAIRS;
    void* hist_w = recs ? nullptr : s->weights;
    int64_t* hist_c = recs ? nullptr : s->counts;
    // Small problems: the stratification update, every dimension's map update and the NEXT pass's get_NH share one
    // launch (vegas_small.cu); get_NH runs on its own only after the schedule may have changed the sample budget.
    const bool small = !recs && small_strat_ok(n_cubes) && (!grid_improve || small_map_ok(dim, ni));
    // Passes of >= 2^20 rows accumulate the histogram as fp64 {sum, count} pairs (one reduction sector per sample and
    // dimension instead of two, tq_fused_vegas); it is folded into weights / counts right before the map update.
    const bool pairs_ok = !recs && !small && s->hist_pairs != nullptr;
    bool pairs_dirty = false;
    // Maps beyond L2 (pair layout + jf2_rows): a stratified pass stores jf^2 per row and tq_vegas_hist_sweep bins the rows band
    // by band afterwards, so the reductions hit L2 instead of being random HBM read-modify-writes.
    const bool sweep_ok = pairs_ok && grid_improve && s->jf2_rows != nullptr && s->sweep_dims_per_group >= 1 && n_strat >= 2;
    auto hist_args = [&](int64_t rows, void*& w, int64_t*& c, void*& h) {
        const bool use_pairs = pairs_ok && rows >= (1 << 20);
        w = use_pairs ? nullptr : hist_w;
        c = use_pairs ? nullptr : hist_c;
        h = use_pairs ? s->hist_pairs : nullptr;
        pairs_dirty |= use_pairs;
    };
    auto update_map = [&]() -> int {
        if (passes >= max_passes) { set_error("tq_vegas_run_fused: more than %d passes", max_passes); return TQ_ERR_UNSUPPORTED; }
        int rc = TQ_OK;
        if (recs && (rc = tq_vegas_map_unpack_records(s->edges_packed, s->weights, s->counts, dim, ni, dtype, stream))) return rc;
        if (pairs_dirty && (rc = tq_vegas_map_unpack_hist(s->hist_pairs, s->weights, s->counts, dim, ni, dtype, stream))) return rc;
        pairs_dirty = false;
        rc = map_update_launch(s->x_edges, s->dx_edges, s->weights, s->counts, recs ? nullptr : s->edges_packed, dim, ni, alpha,
                               dtype, s->status + 4 * passes, false, s->map_ws, s->map_ws_bytes, stream);
        ++passes;
        if (!rc && recs) rc = tq_vegas_map_pack_records(s->x_edges, s->dx_edges, s->edges_packed, dim, ni, dtype, stream);
        return rc;
    };
    cudaMemsetAsync(s->status, 0, 4 * (size_t)max_passes * sizeof(int32_t), st);
    if (warmup) {  // vegas.py:211-266: 5 unstratified passes of starting//5 samples, results discarded
        const int64_t ns = starting / 5;
        for (int w = 0; w < 5; ++w) {
            void *hw, *hp;
            int64_t* hc;
            hist_args(ns, hw, hc, hp);
            int rc = tq_fused_vegas(fn, dtype, nullptr, 0, 1, 0, ns, s->edges_packed, layout, ni, hw, hc, hp, nullptr, nullptr,
                                    seed, call++, s->records, s->ws, s->ws_bytes, stream);
            if (rc) return rc;
            fevals += ns;
            if ((rc = update_map())) return rc;
        }
    }
    MapScratch scratch = {};
    if (small && grid_improve && !map_scratch_carve(s->map_ws, s->map_ws_bytes, dim, ni, dtype, true, scratch)) {
        set_error("tq_vegas_run_fused: map workspace too small");
        return TQ_ERR_WORKSPACE;
    }
    const size_t jf_bytes = 2 * (size_t)n_cubes * elt;  // JF and JF2 are one [2, n_cubes] allocation (tq_vegas_state)
    Block<T> blk;
    int it = 0;
    int first_rec = 0;      // record index of the first iteration of the current block
    bool have_nh = false;   // nh/offsets of the coming pass already computed (and JF/JF2 zeroed) by the last update
    while (true) {
        ++it;
        int rc = TQ_OK;
        if (!have_nh) {  // get_NH's launch also zeroes JF/JF2 for this pass
            rc = strat_nh_launch(s->dh, n_cubes, (double)starting, dtype, s->nh, s->offsets, s->JF, jf_bytes, s->ws, s->ws_bytes,
                                 stream);
            if (rc) return rc;
        }
        // sum nh <= starting * sum(dh) + 2 * n_cubes; the estimate only sizes the grid
        const int64_t m_est = starting + 2 * n_cubes + 1024;
        // without grid improvement nothing is accumulated: the pass only needs the {x, dx} gather
        if (sweep_ok && starting >= (1 << 20) && starting + 2 * n_cubes <= s->jf2_rows_cap) {
            rc = tq_fused_vegas_deferred(fn, dtype, s->offsets, n_cubes, n_strat, 0, -m_est, s->edges_packed, ni, s->jf2_rows, s->JF,
                                         s->JF2, seed, call, s->ws, s->ws_bytes, stream);
            if (!rc)
                rc = tq_vegas_hist_sweep(s->offsets, n_cubes, n_strat, dim, dtype, s->jf2_rows, ni, s->hist_pairs,
                                         s->sweep_dims_per_group, seed, call, s->ws, s->ws_bytes, stream);
            ++call;
            pairs_dirty = true;
        } else {
            void *hw = nullptr, *hp = nullptr;
            int64_t* hc = nullptr;
            if (grid_improve) hist_args(starting, hw, hc, hp);
            rc = tq_fused_vegas(fn, dtype, s->offsets, n_cubes, n_strat, 0, -m_est, s->edges_packed, layout, ni, hw, hc, hp, s->JF,
                                s->JF2, seed, call++, nullptr, s->ws, s->ws_bytes, stream);
        }
        if (rc) return rc;
        if (it > TQ_VEGAS_MAX_PASSES) { set_error("tq_vegas_run_fused: too many iterations"); return TQ_ERR_UNSUPPORTED; }
        double* record = s->records + 4 * (it - 1);
        if (small) {
            have_nh = it % 5 > 0;  // inside a block the next pass keeps this one's budget
            SmallStrat sa = {s->JF, s->JF2, s->nh, n_cubes, v_cubes, beta, s->dh, record, have_nh ? (double)starting : 0.0,
                             s->offsets, s->JF, jf_bytes};
            if (grid_improve) {
                if (passes >= max_passes) { set_error("tq_vegas_run_fused: more than %d passes", max_passes); return TQ_ERR_UNSUPPORTED; }
                SmallMap ma = {s->x_edges, s->dx_edges, s->weights, s->counts, s->edges_packed, scratch, dim, ni, alpha,
                               s->status + 4 * passes, true};
                ++passes;
                rc = small_update_launch(&sa, &ma, dtype, stream);
            } else {
                rc = small_update_launch(&sa, nullptr, dtype, stream);
            }
            if (rc) return rc;
        } else {
            rc = tq_vegas_strat_update(s->JF, s->JF2, s->nh, n_cubes, v_cubes, beta, dtype, s->dh, record, s->ws, s->ws_bytes, stream);
            if (rc) return rc;
            if (grid_improve && (rc = update_map())) return rc;
        }
        if (it % 5 > 0) continue;
        // vegas.py:161-209 on the block of the last (up to) five iterations
        const int nrec = it - first_rec;
        std::vector<double> rec(4 * nrec);
        cudaMemcpyAsync(rec.data(), s->records + 4 * first_rec, rec.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (passes > 0) cudaMemcpyAsync(out->status, s->status, 4 * (size_t)passes * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("tq_vegas_run_fused: %s", cudaGetErrorString(e)); return (int)e; }
        blk.res.clear();
        blk.sig.clear();
        for (int k = 0; k < nrec; ++k) {
            blk.res.push_back((T)rec[4 * k]);
            blk.sig.push_back((T)rec[4 * k + 1]);
            fevals += (int64_t)rec[4 * k + 3];  // sum nh of the pass (vegas.py:291)
        }
        const bool stop = schedule_checkpoint<T>(blk, eps_rel, eps_abs, N, fevals, it, max_it, increment, starting);
        if (stop) break;
        first_rec = it;
    }
    out->it = it;
    out->n_block = (int32_t)blk.res.size();
    out->fevals = fevals;
    out->starting_N = starting;
    out->calls_used = (int32_t)(call);
    out->n_passes = passes;
    for (size_t k = 0; k < blk.res.size() && k < 8; ++k) {
        out->results[k] = (double)blk.res[k];
        out->sigma2[k] = (double)blk.sig[k];
    }
    return TQ_OK;
}

// Optional phase timing of the sharded loop (TQ_VEGAS_TIMING=1): CUDA events around every phase of every pass, summed and
// printed per rank when the run ends.  The all-reduce phase includes the wait for the slowest rank, i.e. the load imbalance.
struct PhaseTimer {
    bool on;
    cudaStream_t st;
    std::vector<cudaEvent_t> ev;
    std::vector<int> phase;
    explicit PhaseTimer(cudaStream_t s) : on(getenv("TQ_VEGAS_TIMING") != nullptr), st(s) {}
    void mark(int ph) {  // start of phase `ph` (= end of the previous one)
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        phase.push_back(ph);
    }
    void report(int rank, const char* const* names, int nph) {
        if (!on) return;
        cudaStreamSynchronize(st);
        std::vector<double> tot(nph, 0.0);
        for (size_t i = 0; i + 1 < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            if (phase[i] >= 0 && phase[i] < nph) tot[phase[i]] += ms;
        }
        fprintf(stderr, "[tq timing rank %d]", rank);
        for (int k = 0; k < nph; ++k) fprintf(stderr, " %s %.3f ms;", names[k], tot[k]);
        fprintf(stderr, "\n");
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
};

// Multi-GPU fused run: every rank owns a block-cyclic share of the hypercubes -- its slice of dh / nh /
// offsets / JF / JF2 never leaves the GPU -- and the map is replicated.  Per pass ONE collective: the fp64 buffer
// [{sum jf^2, count} pairs of the map histogram | I, sigma^2, sum d^beta, sum nh] is summed over the ranks through the
// caller's all-reduce (NCCL over NVLink in torch.distributed); nothing else crosses the links and nothing is read back
// inside a block of five iterations.  The map update runs redundantly on every rank from the summed histogram (it is
// O(dim * Ni), independent of the sample count), or -- `shard_map_update`, for maps beyond L2 -- each rank rebins
// dim / world dimensions after a reduce-scatter and the new edges are all-gathered.
template <typename T>
static int run_fused_sharded(const tq_integrand* fn, int32_t dtype, int64_t N, int32_t max_it, double eps_rel, double eps_abs,
                             bool grid_improve, bool warmup, int64_t ni, int32_t n_strat, int64_t n_cubes, double v_cubes,
                             double alpha, double beta, uint64_t seed, uint32_t call, const tq_vegas_state* s,
                             const tq_vegas_shard* sh, tq_vegas_result* out, void* stream) {
    cudaStream_t st = as_stream(stream);
    const int dim = fn->dim;
    const size_t elt = sizeof(T);
    const int64_t increment = N / (max_it + 5);
    int64_t starting = increment;
    int64_t fevals = 0;
    int passes = 0;
    const int max_passes = TQ_VEGAS_MAX_PASSES;
    const int rank = sh->rank, world = sh->world, lb = sh->cube_block_log2;
    const int64_t n_local = sh->n_cubes_local;
    const bool recs = s->edges_layout == TQ_EDGES_RECORDS;
    const int layout = recs ? TQ_EDGES_RECORDS : TQ_EDGES_PAIRS;
    double* comm = sh->comm;                               // [hist_len + 8]
    const int64_t hist_len = (int64_t)dim * ni * 2;
    double* tail = comm + hist_len;                        // I, sigma^2, sum d^beta, sum nh of the pass
    auto allreduce = [&](int64_t offset, int64_t count) -> int {
        if (sh->allreduce(sh->user, offset, count) != 0) {
            set_error("tq_vegas_run_fused_sharded: the all-reduce callback failed");
            return TQ_ERR_CALLBACK;
        }
        return TQ_OK;
    };
    // summed histogram -> weights / counts -> new edges (replicated on every rank)
    auto update_map = [&]() -> int {
        if (passes >= max_passes) { set_error("tq_vegas_run_fused_sharded: more than %d passes", max_passes); return TQ_ERR_UNSUPPORTED; }
        int rc = tq_vegas_map_unpack_hist(comm, s->weights, s->counts, dim, ni, dtype, stream);
        if (rc) return rc;
        rc = map_update_launch(s->x_edges, s->dx_edges, s->weights, s->counts, recs ? nullptr : s->edges_packed, dim, ni, alpha,
                               dtype, s->status + 4 * passes, false, s->map_ws, s->map_ws_bytes, stream);
        ++passes;
        if (!rc && recs) rc = tq_vegas_map_pack_records(s->x_edges, s->dx_edges, s->edges_packed, dim, ni, dtype, stream);
        return rc;
    };
    const bool sweep_ok = !recs && grid_improve && s->jf2_rows != nullptr && s->sweep_dims_per_group >= 1 && n_strat >= 2;
    cudaMemsetAsync(s->status, 0, 4 * (size_t)max_passes * sizeof(int32_t), st);
    cudaMemsetAsync(tail, 0, 8 * sizeof(double), st);
    PhaseTimer tm(st);
    static const char* const PH[] = {"get_NH", "pass", "strat_update", "all_reduce", "normalise+map_update", "warmup_pass"};
    if (warmup) {  // vegas.py:211-266, rows split evenly over the ranks
        const int64_t ns = starting / 5;
        const int64_t r0 = ns * rank / world, r1 = ns * (rank + 1) / world;
        for (int w = 0; w < 5; ++w) {
            tm.mark(5);
            int rc = tq_fused_vegas(fn, dtype, nullptr, 0, 1, r0, r1, s->edges_packed, layout, ni, nullptr, nullptr,
                                    recs ? nullptr : comm, nullptr, nullptr, seed, call++, s->records, s->ws, s->ws_bytes, stream);
            if (rc) return rc;
            if (recs && (rc = records_to_pairs_launch(s->edges_packed, comm, dim, ni, dtype, stream))) return rc;
            fevals += ns;
            tm.mark(3);
            if ((rc = allreduce(0, hist_len))) return rc;
            tm.mark(4);
            if ((rc = update_map())) return rc;
        }
    }
    const size_t jf_bytes = 2 * (size_t)n_local * elt;
    Block<T> blk;
    int it = 0;
    int first_rec = 0;
    while (true) {
        ++it;
        tm.mark(0);
        int rc = strat_nh_launch(s->dh, n_local, (double)starting, dtype, s->nh, s->offsets, s->JF, jf_bytes, s->ws, s->ws_bytes, stream);
        if (rc) return rc;
        tm.mark(1);
        const int64_t m_est = starting / world + 2 * n_local + 1024;
        if (sweep_ok && starting >= (1 << 20) && starting + 2 * n_local <= s->jf2_rows_cap) {
            // maps beyond L2: the pass stores jf^2 per local row, the band sweeps bin this rank's cubes into `comm`
            // (a rank's rows never exceed starting * sum(dh) + 2 * n_local <= the capacity checked above)
            rc = fused_vegas_launch(fn, dtype, s->offsets, n_local, n_strat, 0, -m_est, s->edges_packed, TQ_EDGES_PAIRS, ni, nullptr,
                                    nullptr, nullptr, s->jf2_rows, s->JF, s->JF2, seed, call, lb, rank, world, nullptr, s->ws,
                                    s->ws_bytes, stream);
            if (!rc)
                rc = hist_sweep_launch(s->offsets, n_cubes, n_strat, dim, dtype, s->jf2_rows, ni, comm, s->sweep_dims_per_group, seed,
                                       call, lb, rank, world, s->ws, s->ws_bytes, stream);
            ++call;
        } else {
            rc = tq_fused_vegas_sharded(fn, dtype, s->offsets, n_local, n_strat, 0, -m_est, s->edges_packed, layout, ni, nullptr, nullptr,
                                        grid_improve && !recs ? comm : nullptr, s->JF, s->JF2, seed, call++, lb, rank, world, nullptr,
                                        s->ws, s->ws_bytes, stream);
        }
        if (rc) return rc;
        if (grid_improve && recs && (rc = records_to_pairs_launch(s->edges_packed, comm, dim, ni, dtype, stream))) return rc;
        if (it > TQ_VEGAS_MAX_PASSES) { set_error("tq_vegas_run_fused_sharded: too many iterations"); return TQ_ERR_UNSUPPORTED; }
        // local estimator sums + unnormalised d^beta, then the pass's one collective, then the normalisation
        tm.mark(2);
        if ((rc = strat_update_partial_launch(s->JF, s->JF2, s->nh, n_local, v_cubes, beta, dtype, s->dh, tail, s->ws, s->ws_bytes, stream))) return rc;
        tm.mark(3);
        if ((rc = grid_improve ? allreduce(0, hist_len + 4) : allreduce(hist_len, 4))) return rc;
        tm.mark(4);
        double* record = s->records + 4 * (it - 1);
        cudaMemcpyAsync(record, tail, 4 * sizeof(double), cudaMemcpyDeviceToDevice, st);
        if ((rc = strat_normalise_launch(s->dh, n_local, record, dtype, stream))) return rc;
        if (grid_improve && (rc = update_map())) return rc;
        tm.mark(-1);
        if (it % 5 > 0) continue;
        const int nrec = it - first_rec;
        std::vector<double> rec(4 * nrec);
        cudaMemcpyAsync(rec.data(), s->records + 4 * first_rec, rec.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (passes > 0) cudaMemcpyAsync(out->status, s->status, 4 * (size_t)passes * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("tq_vegas_run_fused_sharded: %s", cudaGetErrorString(e)); return (int)e; }
        blk.res.clear();
        blk.sig.clear();
        for (int k = 0; k < nrec; ++k) {
            blk.res.push_back((T)rec[4 * k]);
            blk.sig.push_back((T)rec[4 * k + 1]);
            fevals += (int64_t)rec[4 * k + 3];
        }
        // every rank holds the same summed records, so every rank takes the same decision
        if (schedule_checkpoint<T>(blk, eps_rel, eps_abs, N, fevals, it, max_it, increment, starting)) break;
        first_rec = it;
    }
    tm.report(rank, PH, 6);
    out->it = it;
    out->n_block = (int32_t)blk.res.size();
    out->fevals = fevals;
    out->starting_N = starting;
    out->calls_used = (int32_t)(call);
    out->n_passes = passes;
    for (size_t k = 0; k < blk.res.size() && k < 8; ++k) {
        out->results[k] = (double)blk.res[k];
        out->sigma2[k] = (double)blk.sig[k];
    }
    (void)n_cubes;
    return TQ_OK;
}

// The same loop with materialised samples and a callback integrand (tq_vegas_run_unfused).
template <typename T>
static int run_unfused(tq_eval_callback eval, void* user, int dim, int32_t dtype, int64_t N, int32_t max_it, double eps_rel,
                       double eps_abs, bool grid_improve, bool warmup, int64_t ni, int32_t n_strat, int64_t n_cubes,
                       double v_cubes, double alpha, double beta, uint64_t seed, uint32_t call, const tq_vegas_state* s,
                       const tq_vegas_unfused_buffers* b, tq_vegas_result* out, void* stream) {
    cudaStream_t st = as_stream(stream);
    const size_t elt = sizeof(T);
    const int64_t increment = N / (max_it + 5);
    int64_t starting = increment;
    int64_t fevals = 0;
    int passes = 0;
    const int max_passes = TQ_VEGAS_MAX_PASSES;
    const bool recs = s->edges_layout == TQ_EDGES_RECORDS;
    const int layout = recs ? TQ_EDGES_RECORDS : TQ_EDGES_PAIRS;
    const bool small = !recs && small_strat_ok(n_cubes) && (!grid_improve || small_map_ok(dim, ni));
    MapScratch scratch = {};
    if (small && grid_improve && !map_scratch_carve(s->map_ws, s->map_ws_bytes, dim, ni, dtype, true, scratch)) {
        set_error("tq_vegas_run_unfused: map workspace too small");
        return TQ_ERR_WORKSPACE;
    }
    const bool to_pairs = !recs && !small && s->hist_pairs != nullptr;
    auto update_map = [&]() -> int {
        if (passes >= max_passes) { set_error("tq_vegas_run_unfused: more than %d passes", max_passes); return TQ_ERR_UNSUPPORTED; }
        int rc = TQ_OK;
        if (recs && (rc = tq_vegas_map_unpack_records(s->edges_packed, s->weights, s->counts, dim, ni, dtype, stream))) return rc;
        if (to_pairs && (rc = tq_vegas_map_unpack_hist(s->hist_pairs, s->weights, s->counts, dim, ni, dtype, stream))) return rc;
        rc = map_update_launch(s->x_edges, s->dx_edges, s->weights, s->counts, recs ? nullptr : s->edges_packed, dim, ni, alpha,
                               dtype, s->status + 4 * passes, false, s->map_ws, s->map_ws_bytes, stream);
        ++passes;
        if (!rc && recs) rc = tq_vegas_map_pack_records(s->x_edges, s->dx_edges, s->edges_packed, dim, ni, dtype, stream);
        return rc;
    };
    // One pass around the integrand in two kernels (vegas_unfused.cu): x, jac straight from the Philox stream -> f = eval(x) ->
    // jf + histogram with regenerated bins.  The samples y never exist in HBM.  `offs` NULL: a warm-up pass.
    // Histogram target: the records (large maps), the weights / counts arrays (small problems: what the one-launch
    // cluster update reads), else the fp64 pair table (one reduction sector per sample and dimension).
    auto evaluate = [&](const int64_t* offs, int64_t rows, uint32_t call_idx, bool hist, void* jf_out) -> int {
        if (rows > b->cap_rows) {
            set_error("tq_vegas_run_unfused: a pass of %lld rows exceeds the buffers (%lld)", (long long)rows, (long long)b->cap_rows);
            return TQ_ERR_WORKSPACE;
        }
        int rc = tq_vegas_sample_map(offs, n_cubes, n_strat, dim, dtype, 0, rows, s->edges_packed, layout, ni, b->domain, seed, call_idx,
                                     b->x, b->jac, stream);
        if (rc) return rc;
        const void* f = nullptr;
        if (eval(user, rows, &f) != 0 || f == nullptr) {
            set_error("tq_vegas_run_unfused: the integrand callback failed");
            return TQ_ERR_CALLBACK;
        }
        fevals += rows;
        if (!hist && !jf_out) return TQ_OK;
        const bool arrays = hist && !recs && !to_pairs;
        return tq_vegas_accumulate_regen(offs, n_cubes, n_strat, dim, dtype, 0, rows, ni, f, b->jac, b->volume, jf_out, nullptr,
                                         hist && to_pairs ? s->hist_pairs : nullptr, hist && recs ? s->edges_packed : nullptr,
                                         arrays ? s->weights : nullptr, arrays ? s->counts : nullptr, seed, call_idx, stream);
    };
    cudaMemsetAsync(s->status, 0, 4 * (size_t)max_passes * sizeof(int32_t), st);
    if (warmup) {  // vegas.py:211-266
        const int64_t ns = starting / 5;
        for (int w = 0; w < 5; ++w) {
            int rc = evaluate(nullptr, ns, call++, true, nullptr);
            if (rc) return rc;
            if ((rc = update_map())) return rc;
        }
    }
    Block<T> blk;
    int it = 0;
    int first_rec = 0;
    bool have_nh = false;
    while (true) {
        ++it;
        int rc = TQ_OK;
        if (!have_nh) {
            rc = strat_nh_launch(s->dh, n_cubes, (double)starting, dtype, s->nh, s->offsets, nullptr, 0, s->ws, s->ws_bytes, stream);
            if (rc) return rc;
        }
        long long M = 0;
        cudaMemcpyAsync(&M, s->offsets + n_cubes, sizeof(long long), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);  // the callback's view of x needs the row count
        if (e != cudaSuccess) { set_error("tq_vegas_run_unfused: %s", cudaGetErrorString(e)); return (int)e; }
        if (M > b->cap_rows) {
            set_error("tq_vegas_run_unfused: a pass of %lld rows exceeds the buffers (%lld)", M, (long long)b->cap_rows);
            return TQ_ERR_WORKSPACE;
        }
        if ((rc = evaluate(s->offsets, M, call++, grid_improve, b->jf))) return rc;
        rc = tq_vegas_strat_accumulate(b->jf, 0, s->offsets, 0, n_cubes, s->JF, s->JF2, dtype, stream);
        if (rc) return rc;
        if (it > TQ_VEGAS_MAX_PASSES) { set_error("tq_vegas_run_unfused: too many iterations"); return TQ_ERR_UNSUPPORTED; }
        double* record = s->records + 4 * (it - 1);
        if (small) {
            have_nh = it % 5 > 0;
            SmallStrat sa = {s->JF, s->JF2, s->nh, n_cubes, v_cubes, beta, s->dh, record, have_nh ? (double)starting : 0.0,
                             s->offsets, nullptr, 0};
            if (grid_improve) {
                if (passes >= max_passes) { set_error("tq_vegas_run_unfused: more than %d passes", max_passes); return TQ_ERR_UNSUPPORTED; }
                SmallMap ma = {s->x_edges, s->dx_edges, s->weights, s->counts, s->edges_packed, scratch, dim, ni, alpha,
                               s->status + 4 * passes, true};
                ++passes;
                rc = small_update_launch(&sa, &ma, dtype, stream);
            } else {
                rc = small_update_launch(&sa, nullptr, dtype, stream);
            }
            if (rc) return rc;
        } else {
            rc = tq_vegas_strat_update(s->JF, s->JF2, s->nh, n_cubes, v_cubes, beta, dtype, s->dh, record, s->ws, s->ws_bytes, stream);
            if (rc) return rc;
            if (grid_improve && (rc = update_map())) return rc;
        }
        if (it % 5 > 0) continue;
        const int nrec = it - first_rec;
        std::vector<double> rec(4 * nrec);
        cudaMemcpyAsync(rec.data(), s->records + 4 * first_rec, rec.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (passes > 0) cudaMemcpyAsync(out->status, s->status, 4 * (size_t)passes * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("tq_vegas_run_unfused: %s", cudaGetErrorString(e)); return (int)e; }
        blk.res.clear();
        blk.sig.clear();
        for (int k = 0; k < nrec; ++k) {
            blk.res.push_back((T)rec[4 * k]);
            blk.sig.push_back((T)rec[4 * k + 1]);
        }
        if (schedule_checkpoint<T>(blk, eps_rel, eps_abs, N, fevals, it, max_it, increment, starting)) break;
        first_rec = it;
    }
    out->it = it;
    out->n_block = (int32_t)blk.res.size();
    out->fevals = fevals;
    out->starting_N = starting;
    out->calls_used = (int32_t)(call);
    out->n_passes = passes;
    for (size_t k = 0; k < blk.res.size() && k < 8; ++k) {
        out->results[k] = (double)blk.res[k];
        out->sigma2[k] = (double)blk.sig[k];
    }
    (void)elt;
    return TQ_OK;
}

}  // namespace tq

template <typename T>
static int schedule_entry(const double* results, const double* sigma2, int n_block, double eps_rel, double eps_abs, int64_t N,
                          int64_t fevals, int it, int max_it, int64_t increment, int64_t* starting, double* mean_out,
                          int32_t* stop_out) {
    tq::Block<T> blk;
    for (int k = 0; k < n_block; ++k) {
        blk.res.push_back((T)results[k]);
        blk.sig.push_back((T)sigma2[k]);
    }
    *mean_out = (double)blk.mean();
    *stop_out = tq::schedule_checkpoint<T>(blk, eps_rel, eps_abs, N, fevals, it, max_it, increment, *starting) ? 1 : 0;
    return TQ_OK;
}

extern "C" int tq_vegas_schedule(const double* results_host, const double* sigma2_host, int32_t n_block, int32_t dtype,
                                 double eps_rel, double eps_abs, int64_t N, int64_t fevals, int32_t it,
                                 int32_t max_iterations, int64_t increment, int64_t* starting_N_inout, double* mean_out,
                                 int32_t* stop_out) {
    TQ_REQUIRE(results_host && sigma2_host && starting_N_inout && mean_out && stop_out && n_block >= 1,
               "tq_vegas_schedule: NULL argument or empty block");
    if (dtype == TQ_F32)
        return schedule_entry<float>(results_host, sigma2_host, n_block, eps_rel, eps_abs, N, fevals, it, max_iterations, increment,
                                     starting_N_inout, mean_out, stop_out);
    if (dtype == TQ_F64)
        return schedule_entry<double>(results_host, sigma2_host, n_block, eps_rel, eps_abs, N, fevals, it, max_iterations, increment,
                                      starting_N_inout, mean_out, stop_out);
    tq::set_error("tq_vegas_schedule: unsupported dtype %d", dtype);
    return TQ_ERR_INVALID_ARGUMENT;
}

extern "C" int tq_vegas_run_unfused(tq_eval_callback eval, void* user, int32_t dim, int32_t dtype, int64_t N,
                                    int32_t max_iterations, double eps_rel, double eps_abs, int32_t use_grid_improve,
                                    int32_t use_warmup, int64_t n_intervals, int32_t n_strat, int64_t n_cubes, double v_cubes,
                                    double alpha, double beta, uint64_t seed, uint32_t first_call, const tq_vegas_state* state,
                                    const tq_vegas_unfused_buffers* buffers, tq_vegas_result* result_host, void* stream) {
    TQ_REQUIRE(eval && state && buffers && result_host, "tq_vegas_run_unfused: NULL argument");
    TQ_REQUIRE(dim >= 1 && dim <= 255, "tq_vegas_run_unfused: dim %d out of range", dim);
    TQ_REQUIRE(N >= 1 && max_iterations >= 1 && max_iterations + 5 <= TQ_VEGAS_MAX_PASSES,
               "tq_vegas_run_unfused: max_iterations must be in [1, %d]", TQ_VEGAS_MAX_PASSES - 5);
    TQ_REQUIRE(n_cubes >= 1 && n_strat >= 1 && n_intervals >= 2, "tq_vegas_run_unfused: bad map / stratification sizes");
    if (dtype == TQ_F32)
        return tq::run_unfused<float>(eval, user, dim, dtype, N, max_iterations, eps_rel, eps_abs, use_grid_improve != 0,
                                      use_warmup != 0, n_intervals, n_strat, n_cubes, v_cubes, alpha, beta, seed, first_call, state,
                                      buffers, result_host, stream);
    if (dtype == TQ_F64)
        return tq::run_unfused<double>(eval, user, dim, dtype, N, max_iterations, eps_rel, eps_abs, use_grid_improve != 0,
                                       use_warmup != 0, n_intervals, n_strat, n_cubes, v_cubes, alpha, beta, seed, first_call, state,
                                       buffers, result_host, stream);
    tq::set_error("tq_vegas_run_unfused: unsupported dtype %d", dtype);
    return TQ_ERR_INVALID_ARGUMENT;
}

extern "C" int tq_vegas_run_fused_sharded(const tq_integrand* fn_host, int32_t dtype, int64_t N, int32_t max_iterations,
                                          double eps_rel, double eps_abs, int32_t use_grid_improve, int32_t use_warmup,
                                          int64_t n_intervals, int32_t n_strat, int64_t n_cubes, double v_cubes, double alpha,
                                          double beta, uint64_t seed, uint32_t first_call, const tq_vegas_state* state,
                                          const tq_vegas_shard* shard, tq_vegas_result* result_host, void* stream) {
    TQ_REQUIRE(fn_host && state && result_host && shard, "tq_vegas_run_fused_sharded: NULL argument");
    TQ_REQUIRE(N >= 1 && max_iterations >= 1 && max_iterations + 5 <= TQ_VEGAS_MAX_PASSES,
               "tq_vegas_run_fused_sharded: max_iterations must be in [1, %d]", TQ_VEGAS_MAX_PASSES - 5);
    TQ_REQUIRE(n_cubes >= 1 && n_strat >= 1 && n_intervals >= 2, "tq_vegas_run_fused_sharded: bad map / stratification sizes");
    TQ_REQUIRE(shard->world >= 1 && shard->rank >= 0 && shard->rank < shard->world && shard->n_cubes_local >= 1 &&
                   shard->cube_block_log2 >= 0 && shard->cube_block_log2 < 31 && shard->comm && shard->allreduce,
               "tq_vegas_run_fused_sharded: bad shard description");
    if (dtype == TQ_F32)
        return tq::run_fused_sharded<float>(fn_host, dtype, N, max_iterations, eps_rel, eps_abs, use_grid_improve != 0, use_warmup != 0,
                                            n_intervals, n_strat, n_cubes, v_cubes, alpha, beta, seed, first_call, state, shard,
                                            result_host, stream);
    if (dtype == TQ_F64)
        return tq::run_fused_sharded<double>(fn_host, dtype, N, max_iterations, eps_rel, eps_abs, use_grid_improve != 0, use_warmup != 0,
                                             n_intervals, n_strat, n_cubes, v_cubes, alpha, beta, seed, first_call, state, shard,
                                             result_host, stream);
    tq::set_error("tq_vegas_run_fused_sharded: unsupported dtype %d", dtype);
    return TQ_ERR_INVALID_ARGUMENT;
}

extern "C" int tq_vegas_run_fused(const tq_integrand* fn_host, int32_t dtype, int64_t N, int32_t max_iterations,
                                  double eps_rel, double eps_abs, int32_t use_grid_improve, int32_t use_warmup,
                                  int64_t n_intervals, int32_t n_strat, int64_t n_cubes, double v_cubes, double alpha,
                                  double beta, uint64_t seed, uint32_t first_call, const tq_vegas_state* state,
                                  tq_vegas_result* result_host, void* stream) {
    TQ_REQUIRE(fn_host && state && result_host, "tq_vegas_run_fused: NULL argument");
    TQ_REQUIRE(N >= 1 && max_iterations >= 1 && max_iterations + 5 <= TQ_VEGAS_MAX_PASSES,
               "tq_vegas_run_fused: max_iterations must be in [1, %d]", TQ_VEGAS_MAX_PASSES - 5);
    TQ_REQUIRE(n_cubes >= 1 && n_strat >= 1 && n_intervals >= 2, "tq_vegas_run_fused: bad map / stratification sizes");
    if (dtype == TQ_F32)
        return tq::run_fused<float>(fn_host, dtype, N, max_iterations, eps_rel, eps_abs, use_grid_improve != 0, use_warmup != 0,
                                    n_intervals, n_strat, n_cubes, v_cubes, alpha, beta, seed, first_call, state, result_host, stream);
    if (dtype == TQ_F64)
        return tq::run_fused<double>(fn_host, dtype, N, max_iterations, eps_rel, eps_abs, use_grid_improve != 0, use_warmup != 0,
                                     n_intervals, n_strat, n_cubes, v_cubes, alpha, beta, seed, first_call, state, result_host, stream);
    tq::set_error("tq_vegas_run_fused: unsupported dtype %d", dtype);
    return TQ_ERR_INVALID_ARGUMENT;
}
