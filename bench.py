#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: integrand evaluations per second (VEGAS, MC, Boole).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME|all]

Workloads (BASELINE.json `configs`):
  mc10             configs[1]  MonteCarlo, 10-D sum-of-sines, N=1e9 evals per GPU, fp32       (HEADLINE: the metric config)
  vegas4           configs[0]  VEGAS 4-D Genz Gaussian N=1e6 fp64 (the reference's CPU-runnable case)
  boole6           configs[2]  Boole 6-D product-of-cosines, 33^6 points, fp64, grid sharded over the ranks
  vegas8_cap4096   configs[3]  VEGAS 8-D Genz oscillatory fp64, N=2.5e9 (1e8-sample iterations), map capped at 4096
                               intervals per dimension (L2-resident); fixed N at every GPU count (strong scaling)
  vegas8           configs[3]  the same at the reference's own map size Ni = N/250 = 1e7 per dimension (2.6 GB of records)
  vegas16_cap4096  configs[4]  VEGAS 16-D Genz product-peak fp32, N=1e10, fused path, map capped at 4096 (the reference
                               itself raises "Could not replace all infinite edges" for its fp32 map with Ni=4e7)
The default run (`--workload all`) prints ONE JSON line: the headline record (mc10) at top level, plus `workloads`:
one sub-record per other workload, each with its own value / e2e / roofline (and cpu_baseline at N=1).
A "step" is one complete integration of the workload.  `value` times the fused functor path (inputs: the [dim, 2]
domain, resident in HBM); `e2e` times the same call through the public drop-in API with the domain in host memory and
the result read back to the host every step; `unfused` reports the torch-callable path (points materialised in HBM).
Under torchrun each rank owns one GPU; time is the max over ranks of CUDA-event time.
`--impl reference` times the UNMODIFIED reference (baseline/_ref, staged by baseline/stage_ref.sh) on the host cores.
"""
import argparse
import json
import math
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "mc10": dict(kind="mc", dim=10, N=10**9, dtype="float32", integrand="sum_sin", scaling="weak", config_index=1),
    "vegas4": dict(kind="vegas", dim=4, N=10**6, dtype="float64", integrand="genz_gaussian", scaling="strong", config_index=0),
    "boole6": dict(kind="boole", dim=6, N=33**6, dtype="float64", integrand="prod_cos", scaling="strong", config_index=2),
    "vegas8_cap4096": dict(kind="vegas", dim=8, N=2_500_000_000, dtype="float64", integrand="genz_oscillatory", map_cap=4096,
                           scaling="strong", config_index=3),
    "vegas8": dict(kind="vegas", dim=8, N=2_500_000_000, dtype="float64", integrand="genz_oscillatory", map_cap=None,
                   scaling="strong", config_index=3),
    "vegas16_cap4096": dict(kind="vegas", dim=16, N=10**10, dtype="float32", integrand="genz_product_peak", map_cap=4096,
                            scaling="strong", config_index=4,
                            note="map capped: the unmodified reference raises 'Could not replace all infinite edges' for an fp32 "
                                 "map with Ni = N/250 = 4e7 (only 2.6e7 of its initial edges are distinct)"),
}
HEADLINE = "mc10"
SUB_WORKLOADS = ["vegas8_cap4096", "vegas8", "vegas16_cap4096", "vegas4", "boole6"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=sorted(WORKLOADS) + ["all"])
    ap.add_argument("--no-unfused", action="store_true", help="skip the unfused torch-callable arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--map-cap", type=int, default=None,
                    help="VEGAS workloads: override the cap on map intervals per dimension (VEGAS.max_map_intervals)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------ integrands
def make_integrand(name, dim, fast_math=False):
    from torchquad_b200 import integrands as F

    if name == "sum_sin":
        return F.SumOfSines(dim, fast_math=fast_math)
    if name == "prod_cos":
        return F.ProductOfCosines(dim)
    if name == "genz_gaussian":
        return F.GenzGaussian(dim, a=5.0, u=0.5)
    if name == "genz_oscillatory":
        return F.GenzOscillatory(dim, a=0.5, u=0.3)
    if name == "genz_product_peak":
        return F.GenzProductPeak(dim, a=2.0, u=0.5)
    raise ValueError(name)


def torch_callable(name):
    """The same integrand written the way a torchquad user writes it (plain torch ops)."""
    if name == "sum_sin":
        return lambda x: torch.sum(torch.sin(x), dim=1)
    if name == "prod_cos":
        return lambda x: torch.prod(torch.cos(x), dim=1)
    if name == "genz_gaussian":
        return lambda x: torch.exp(-torch.sum(25.0 * (x - 0.5) ** 2, dim=1))
    if name == "genz_oscillatory":
        return lambda x: torch.cos(2.0 * 3.141592653589793 * 0.3 + torch.sum(0.5 * x, dim=1))
    if name == "genz_product_peak":
        return lambda x: torch.prod(1.0 / (0.25 + (x - 0.5) ** 2), dim=1)
    raise ValueError(name)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            # nvidia-smi's start-up (NVML initialisation) stalls kernel launches for a while: wait for its first
            # sample so that only the periodic queries overlap the timed region
            t0 = time.time()
            while not self.rows and time.time() - t0 < 10.0:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        return len(self.rows)

    def summary(self, first=0, last=None):
        rows = self.rows[first:last]
        sm = [float(r[1]) for r in rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if len(r) > 8 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) > 8 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        return self.summary()


# ------------------------------------------------------------------------------------ timing helpers
def l2_flush(buf):
    buf.add_(1)  # 512 MiB read+write: evicts the 126 MB L2 between timed iterations


def agree_max(x):
    """max over ranks of a host float (identity in a single-process run)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def timed_steps(step, steps, warmup, flush_buf, barrier, min_warm_s=0.3):
    """Per-step CUDA-event times (seconds) on the current stream; L2 flushed between iterations."""
    # W warm-up steps, and at least 0.3 s of them: the GPU drops its clocks while the host sets things up (e.g. waits
    # for nvidia-smi), and a millisecond-scale step would otherwise be timed on the ramp
    # The number of extra steps is derived from a rank-agreed time (steps contain collectives under torchrun).
    t0 = time.time()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    spent = agree_max(time.time() - t0)
    if spent < min_warm_s:
        for _ in range(min(200, int(math.ceil((min_warm_s - spent) / max(spent / max(warmup, 1), 1e-4))))):
            step()
    torch.cuda.synchronize()
    times = []
    from torchquad_b200 import _lib as _tq_lib

    timed_steps.launches = 0
    for _ in range(steps):
        l2_flush(flush_buf)
        barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _tq_lib.kernel_launches()
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        timed_steps.launches += _tq_lib.kernel_launches() - n0
        times.append(a.elapsed_time(b) * 1e-3)
    return times


def max_over_ranks(x, device, world):
    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def time_call(fn, reps=5):
    """Mean CUDA-event time (s) of fn() alone on the current stream after one warm-up call."""
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return statistics.mean(ts)


# ------------------------------------------------------------------------------------ reference / CPU arm
CPU_SAMPLE = {  # bounded samples of each workload for the host-core baseline (BASELINE.md "CPU-baseline plan")
    "mc10": dict(N=10**7), "boole6": dict(n_per_dim=17), "vegas4": dict(N=10**6), "vegas8": dict(N=2_500_000),
    "vegas8_cap4096": dict(N=2_500_000), "vegas16_cap4096": dict(N=2_000_000),
}


def cpu_reference_runner(name, wl):
    """(run, sample description, kind): one bounded repetition of the workload on the host cores through the UNMODIFIED
    reference (baseline/_ref, kind "reference"); the oracle port (same ATen CPU kernels, kind "port") only when the
    staged copy is missing.  run() returns the evaluations it did."""
    from baseline import ref as ref_loader

    torch.set_num_threads(os.cpu_count() or 1)
    dt = getattr(torch, wl["dtype"])
    fn = torch_callable(wl["integrand"])
    dim = wl["dim"]
    sample = CPU_SAMPLE[name]
    torch.set_default_dtype(dt)  # the reference builds list domains in torch's default dtype
    dom_list = [[0.0, 1.0]] * dim
    if ref_loader.available():
        tqr = ref_loader.import_reference()
        where = f"unmodified esa/torchquad {getattr(tqr, '__version__', '0.5.0')} from baseline/_ref, torch CPU backend"
        if wl["kind"] == "mc":
            n = sample["N"]
            integ = tqr.MonteCarlo()

            def run():
                integ.integrate(fn, dim, N=n, integration_domain=dom_list, seed=0, backend="torch")
                return n

            return run, f"MonteCarlo {dim}-D N={n:,} per repetition; {where}", "reference"
        if wl["kind"] == "boole":
            npd = sample["n_per_dim"]
            integ = tqr.Boole()

            def run():
                integ.integrate(fn, dim, N=npd**dim, integration_domain=dom_list, backend="torch")
                return npd**dim

            return run, f"Boole {dim}-D n={npd} per dim per repetition; {where}", "reference"
        n = min(wl["N"], sample["N"])
        integ = tqr.VEGAS()

        def run():
            integ.integrate(fn, dim, N=n, integration_domain=dom_list, seed=0, backend="torch")
            return integ._nr_of_fevals

        return run, f"VEGAS {dim}-D N={n:,} per repetition (the reference's own map size Ni = N/250); {where}", "reference"
    from oracle import ref_oracle as O

    where = "oracle/ref_oracle.py (restates the reference over the same ATen CPU kernels; baseline/_ref is not staged here)"
    dom = torch.tensor(dom_list, dtype=dt)
    if wl["kind"] == "mc":
        n = sample["N"]

        def run():
            O.mc_integrate(fn, dim, n, dom, seed=0)
            return n

        return run, f"MonteCarlo {dim}-D N={n:,} per repetition; {where}", "port"
    if wl["kind"] == "boole":
        npd = sample["n_per_dim"]

        def run():
            pts, hs, n_ = O.nc_grid("boole", npd**dim, dom)
            O.nc_result("boole", fn(pts), dim, n_, hs)
            return npd**dim

        return run, f"Boole {dim}-D n={npd} per dim per repetition; {where}", "port"
    n = min(wl["N"], sample["N"])

    def run():
        g = torch.Generator().manual_seed(0)
        r = O.VegasRun(fn, dim, n, dom, lambda size, dtype: torch.rand(size, dtype=dtype, generator=g))
        r.run()
        return r.fevals

    return run, f"VEGAS {dim}-D N={n:,} per repetition; {where}", "port"


def cpu_reference_rate(name, wl, budget_s=8.0, steps=None, warmup=1):
    """Time the reference on the host: `steps` repetitions when given, else a time budget."""
    run, sample, kind = cpu_reference_runner(name, wl)
    for _ in range(max(1, warmup)):
        run()  # warm-up (thread pool, allocator)
    evals, reps, t0 = 0, 0, time.perf_counter()
    while (reps < steps) if steps else (time.perf_counter() - t0 < budget_s or reps < 2):
        evals += run()
        reps += 1
    dt_s = time.perf_counter() - t0
    return {"value": evals / dt_s, "unit": "evals/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": kind,
            "sample": f"{sample} x {reps} repetitions ({dt_s:.1f} s)", "ms_per_step": dt_s / reps * 1e3}


def workload_config(name, wl, path):
    cfg = {"workload": name, "baseline_config": f"configs[{wl['config_index']}]", "kind": wl["kind"], "dim": wl["dim"],
           ("N_per_gpu" if wl["scaling"] == "weak" else "N_total"): wl["N"], "integrand": wl["integrand"], "path": path}
    if wl["kind"] == "vegas":
        cfg["map_cap"] = wl.get("map_cap")
    if wl.get("note"):
        cfg["note"] = wl["note"]
    return cfg


def run_reference(args, names):
    rank, _, world = dist_env()
    if rank != 0:
        return
    records = {}
    for name in names:
        wl = WORKLOADS[name]
        head = name == names[0]
        base = cpu_reference_rate(name, wl, steps=args.steps if head else 2, warmup=args.warmup if head else 1)
        ms = base.pop("ms_per_step")
        records[name] = {
            "metric": "integrand evals/s", "value": base["value"], "unit": "evals/s", "ms_per_step": ms,
            "higher_is_better": True, "scaling": wl["scaling"], "dtype": "f32" if wl["dtype"] == "float32" else "f64",
            "config": workload_config(name, wl, "reference implementation on the host cores, bounded sample per step"),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
    head = records[names[0]]
    line = {"impl": "reference", **head, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "vs_baseline": None,
            "data": "synthetic", "gpu_launches": 0}
    if len(names) > 1:
        line["workloads"] = {n: records[n] for n in names[1:]}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def build_steps(name, wl, device, world, map_cap, seed0=0):
    """Returns (fused_step, e2e_step, unfused_step, evals(), info)."""
    import torchquad_b200 as tq

    dt = getattr(torch, wl["dtype"])
    dim = wl["dim"]
    fn = make_integrand(wl["integrand"], dim)
    call = torch_callable(wl["integrand"])
    dom_host = [[0.0, 1.0]] * dim
    dom_dev = torch.tensor(dom_host, dtype=dt, device=device)
    N = wl["N"] * (world if wl["scaling"] == "weak" else 1)
    state = {"seed": seed0}
    torch.set_default_dtype(dt)  # list domains take torch's default dtype, like in the reference
    if wl["kind"] == "mc":
        integ = tq.MonteCarlo()

        def fused():
            state["seed"] += 1
            return integ.integrate(fn, dim, N=N, integration_domain=dom_dev, seed=state["seed"])

        def e2e():
            state["seed"] += 1
            return float(integ.integrate(fn, dim, N=N, integration_domain=dom_host, seed=state["seed"], backend="torch"))

        def unfused():
            state["seed"] += 1
            return float(integ.integrate(call, dim, N=N, integration_domain=dom_host, seed=state["seed"], backend="torch"))

        evals = lambda: N  # noqa: E731
    elif wl["kind"] == "boole":
        integ = tq.Boole()

        def fused():
            return integ.integrate(fn, dim, N=N, integration_domain=dom_dev)

        def e2e():
            return float(integ.integrate(fn, dim, N=N, integration_domain=dom_host, backend="torch"))

        def unfused():
            return float(integ.integrate(call, dim, N=N, integration_domain=dom_host, backend="torch"))

        evals = lambda: integ._nr_of_fevals  # noqa: E731
    else:
        integ = tq.VEGAS()
        integ.max_map_intervals = map_cap

        def fused():
            state["seed"] += 1
            return integ.integrate(fn, dim, N=N, integration_domain=dom_dev, seed=state["seed"])

        def e2e():
            state["seed"] += 1
            return float(integ.integrate(fn, dim, N=N, integration_domain=dom_host, seed=state["seed"], backend="torch"))

        def unfused():
            state["seed"] += 1
            return float(integ.integrate(call, dim, N=N, integration_domain=dom_host, seed=state["seed"], backend="torch"))

        evals = lambda: integ._nr_of_fevals  # noqa: E731
    info = {"dtype": dt, "dim": dim, "exact": fn.exact(), "integrator": integ, "fn": fn, "N": N}
    if wl["kind"] == "mc":
        fast_fn = make_integrand(wl["integrand"], dim, fast_math=True)

        def fused_fast():
            state["seed"] += 1
            return integ.integrate(fast_fn, dim, N=N, integration_domain=dom_dev, seed=state["seed"])

        info["fused_fast"] = fused_fast
    return fused, e2e, unfused, evals, info


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_metrics():
    """Per-kernel figures taken from the `ncu --set full` captures of this round (written by scripts/ncu_summary.py --json
    into profiles/r2/ncu_metrics.json): dram bytes and warp instructions per launch + the launch's work units."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2", "ncu_metrics.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def ncu_entry(key):
    return ncu_metrics().get(key) or {}


def issue_roofline(kernel, ncu_key, evals_per_s, sm_count, sm_mhz):
    """Fused kernels move no sample data: they are bound by instruction issue.  achieved = warp instructions per eval
    (ncu smsp__inst_executed.sum / evals of the captured launch) x evals/s; peak = 4 schedulers x SMs x the SM clock
    sampled during the timed region."""
    e = ncu_entry(ncu_key)
    if not e.get("warp_inst_per_unit") or not sm_mhz:
        return {"kernel": kernel, "bound": "instruction issue (no tensor, no sample traffic)", "achieved": None, "peak": None,
                "unit": "G warp-inst/s", "frac": None, "traffic": e.get("dram_bytes"), "note": f"no ncu capture for {ncu_key}"}
    winst = e["warp_inst_per_unit"] * evals_per_s
    peak = 4.0 * sm_count * sm_mhz * 1e6
    return {"kernel": kernel, "bound": "instruction issue (no tensor, no sample traffic)", "achieved": winst / 1e9,
            "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": winst / peak, "traffic": e.get("dram_bytes"),
            "warp_inst_per_eval": e["warp_inst_per_unit"], "ncu_issue_active_pct": e.get("issue_active_pct"),
            "source": f"profiles/r2/ncu_metrics.json[{ncu_key}]", "peak_source": "4 issue slots x SMs x sampled SM clock"}


def mc_kernel_rooflines(wl, device):
    """Live CUDA-event timings of the unfused path's own kernels alone (HBM-bound) + FP issue microbenchmarks."""
    import ctypes

    from torchquad_b200 import _lib, ops

    out = {}
    peaks, peak_src = measured_peaks()
    dt = getattr(torch, wl["dtype"])
    dim = wl["dim"]
    sink = torch.zeros(1, dtype=torch.float64, device=device)
    micro = {}
    for kind, name in [(0, "fp32_fma_per_s"), (1, "fp64_fma_per_s"), (2, "philox_blocks_per_s")]:
        ops_out = ctypes.c_double()
        t = time_call(lambda: _lib.call("tq_peak_microbench", kind, 20000 if kind != 2 else 2000, sink.data_ptr(),
                                        ctypes.byref(ops_out), _lib.stream_ptr(device)))
        micro[name] = ops_out.value / t
    out["microbench"] = micro
    rows = 2 * 10**8
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt, device=device)
    buf = torch.empty((rows, dim), dtype=dt, device=device)

    def gen():
        _lib.call("tq_mc_sample", buf.data_ptr(), dom.data_ptr(), 0, rows, dim, _lib.dtype_code(dt), 1, 0,
                  _lib.stream_ptr(device))

    t = time_call(gen)
    bytes_alg = rows * dim * buf.element_size()  # algorithmic: every sample coordinate written once
    out["roofline_unfused"] = {
        "kernel": "uniform_kernel<T,true> (tq_mc_sample: Philox + affine map, points written to HBM)",
        "bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": bytes_alg / t / 1e9 / peaks["hbm_gbs"], "traffic": ncu_entry("uniform_kernel_f32_d10").get("dram_bytes"),
        "peak_source": peak_src, "launch_ms": t * 1e3, "algorithmic_bytes_per_launch": bytes_alg,
    }
    del buf
    f = torch.rand(rows, dtype=dt, device=device)
    t = time_call(lambda: ops.sum_columns(f))
    out["roofline_reduce"] = {
        "kernel": "sum1_kernel<T> (tq_sum_columns: fp64-accumulated reduction of f)", "bound": "hbm",
        "achieved": rows * f.element_size() / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": rows * f.element_size() / t / 1e9 / peaks["hbm_gbs"], "traffic": ncu_entry("sum1_kernel_f32").get("dram_bytes"),
        "peak_source": peak_src, "launch_ms": t * 1e3,
    }
    return out


def vegas_kernel_roofline(name, wl, info, device):
    """The dominant kernel of a fused VEGAS run -- one stratified pass of fused_vegas_kernel over the run's final map and
    sample allocation -- timed alone with CUDA events, against the roofline that bounds it:
      map in L2 (capped):   the L2 reduction rate.  Algorithmic work per sample = dim reduction sectors ({sum jf^2, count} of
                            one bin each); peak = the same paired RED.F64 pattern alone on a table of the same size, measured
                            live (tq_red_microbench).
      map beyond L2:        HBM.  Algorithmic bytes per sample = dim x 64 (one 32-byte record sector read, one written back)."""
    import ctypes

    from torchquad_b200 import _lib, ops

    v = info["integrator"]
    vmap, strat = v.map, v.strat
    dim, dt = info["dim"], info["dtype"]
    offsets = strat._offsets
    rows = int(offsets[-1].item())
    s = v._fn_struct
    JF = torch.zeros((2, strat.N_cubes), dtype=dt, device=device)
    peaks, peak_src = measured_peaks()
    g = vmap.sweep_group(strat.N_strat)
    if vmap.wants_records() and g >= 1:
        # maps beyond L2 on one GPU: the pass stores jf^2 per row (tq_fused_vegas_deferred), the band sweeps bin the rows
        jf2 = torch.empty(rows, dtype=dt, device=device)
        h = vmap.hist_pairs()
        h.zero_()
        edges = vmap.packed_edges()
        t_pass = time_call(lambda: ops.fused_vegas_deferred(s, edges, 0, rows, 1, 7, offsets, strat.N_strat, JF[0], JF[1], jf2))
        t_sweep = time_call(lambda: ops.hist_sweep(offsets, strat.N_strat, dim, jf2, vmap.N_intervals, h, g, 1, 7))
        h.zero_()
        elt = jf2.element_size()
        per_sample = dim * 32 + elt + dim * elt  # one 32-byte edge sector per dimension, jf^2 written once and read per sweep
        t = t_pass + t_sweep
        e = ncu_entry(name + ":fused_vegas_kernel")
        e2 = ncu_entry(name + ":hist_sweep_kernel")
        traffic = None
        if e.get("dram_bytes_per_unit") and e2.get("dram_bytes_per_unit"):
            traffic = e["dram_bytes_per_unit"] + dim / g * e2["dram_bytes_per_unit"]
        return {"kernel": "fused_vegas_kernel<STRAT> (deferred histogram) + hist_sweep_kernel x dim (one pass)", "bound": "hbm",
                "achieved": rows * per_sample / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": rows * per_sample / t / 1e9 / peaks["hbm_gbs"], "traffic": traffic * rows if traffic else None,
                "peak_source": peak_src, "launch_ms": t * 1e3, "pass_ms": t_pass * 1e3, "sweeps_ms": t_sweep * 1e3,
                "rows_per_launch": rows, "samples_per_s": rows / t, "algorithmic_bytes_per_sample": per_sample,
                "ncu_dram_bytes_per_sample": traffic,
                "note": "the edge gathers are random 32-byte sector reads of a 1.3 GB table: HBM delivers 64-byte bursts, so they "
                        "cost twice their algorithmic bytes; the histogram itself stays in L2 (one band at a time)"}
    if vmap.wants_records():
        rec = vmap.records()
        run = lambda: ops.fused_vegas(s, None, None, None, 0, rows, 1, 7, offsets=offsets, n_strat=strat.N_strat, JF=JF[0],  # noqa: E731
                                      JF2=JF[1], records=rec, dtype=dt, n_intervals=vmap.N_intervals)
        t = time_call(run)
        bytes_alg = rows * dim * 64
        e = ncu_entry(name + ":fused_vegas_kernel")
        traffic = e.get("dram_bytes_per_unit")
        return {"kernel": "fused_vegas_kernel<STRAT> (record layout, one pass)", "bound": "hbm", "achieved": bytes_alg / t / 1e9,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": bytes_alg / t / 1e9 / peaks["hbm_gbs"],
                "traffic": traffic * rows if traffic else None, "peak_source": peak_src, "launch_ms": t * 1e3,
                "rows_per_launch": rows, "samples_per_s": rows / t, "algorithmic_bytes_per_sample": dim * 64,
                "ncu_dram_bytes_per_sample": traffic,
                "note": "random 32-byte sector accesses: HBM delivers 64-byte bursts, so 0.5 is the ceiling of this fraction"}
    h = vmap.hist_pairs()
    h.zero_()
    # row_end < 0: the pass reads its row count on the device, exactly like the passes of the run (which selects the
    # band-privatised tile kernel for fp32)
    run = lambda: ops.fused_vegas(s, vmap.packed_edges(), None, None, 0, -rows, 1, 7, offsets=offsets, n_strat=strat.N_strat,  # noqa: E731
                                  JF=JF[0], JF2=JF[1], hist_pairs=h)
    t = time_call(run)
    tile = dt == torch.float32
    h.zero_()
    bins = dim * vmap.N_intervals
    table = torch.zeros((bins, 2), dtype=torch.float64, device=device)
    ops_out = ctypes.c_double()
    t_red = time_call(lambda: _lib.call("tq_red_microbench", table.data_ptr(), bins, 2000, ctypes.byref(ops_out),
                                        _lib.stream_ptr(device)))
    peak = ops_out.value / t_red
    ach = rows * dim / t
    e = ncu_entry(name + (":fused_vegas_tile_kernel" if tile else ":fused_vegas_kernel"))
    return {"kernel": ("fused_vegas_tile_kernel (slow dimensions' bands in shared memory, one pass)" if tile
                       else "fused_vegas_kernel<STRAT> (pair layout, one pass)"), "bound": "l2 reduction sectors (map resident in L2)",
            "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G reduction sectors/s", "frac": ach / peak,
            "traffic": e.get("dram_bytes"), "launch_ms": t * 1e3, "rows_per_launch": rows, "samples_per_s": rows / t,
            "algorithmic_sectors_per_sample": dim, "ncu_issue_active_pct": e.get("issue_active_pct"),
            "ncu_lts_throughput_pct": e.get("lts_throughput_pct"), "ncu_l1tex_throughput_pct": e.get("l1tex_throughput_pct"),
            "peak_source": f"tq_red_microbench: paired RED.F64 alone on a {bins}-bin fp64 pair table, measured in this run"}


def measure(name, wl, args, ctx, headline):
    """One workload on this rank: fused value, e2e, (N=1) kernel roofline + unfused arm.  Returns the record on rank 0."""
    device, world, rank = ctx["device"], ctx["world"], ctx["rank"]
    flush, barrier, sampler = ctx["flush"], ctx["barrier"], ctx["sampler"]
    map_cap = args.map_cap if (args.map_cap is not None and wl["kind"] == "vegas") else wl.get("map_cap")
    # A run of 15 passes of ~6e4 samples (configs[0], 0.8 ms) is launch latency, not work: it does not shard.  Under torchrun
    # every rank integrates the whole problem with its OWN seed -- N independent replicas, no collective -- and the line
    # reports their aggregate rate as weak scaling.
    replicas = world > 1 and wl["kind"] == "vegas" and wl["N"] < 10**8
    if replicas:
        import torchquad_b200 as tq

        tq.distributed.disable()
    fused, e2e, unfused, evals, info = build_steps(name, wl, device, world, map_cap, seed0=1000 * rank if replicas else 0)
    steps = args.steps if headline else max(2, min(args.steps, 4))
    warmup = max(3, args.warmup) if headline else 3
    mark0 = sampler.mark() if sampler else 0
    times = timed_steps(fused, steps, warmup, flush, barrier)
    launches = timed_steps.launches
    mark1 = sampler.mark() if sampler else 0
    t_fused = max_over_ranks(sum(times), device, world)
    n_evals = evals()
    result_check = float(fused())
    e2e_steps = max(2, steps // 2)
    t_e2e = max_over_ranks(sum(timed_steps(e2e, e2e_steps, 1, flush, barrier, min_warm_s=0.0)), device, world)
    n_evals_e2e = evals()
    fast = unf = None
    if "fused_fast" in info and headline:
        f_steps = max(2, steps // 2)
        t_fast = max_over_ranks(sum(timed_steps(info["fused_fast"], f_steps, 1, flush, barrier)), device, world)
        fast = {"value": evals() * f_steps / t_fast, "unit": "evals/s", "ms_per_step": t_fast / f_steps * 1e3,
                "last_integral": float(info["fused_fast"]()),
                "note": "same fused kernel with sin() evaluated by the SFU (__sinf, abs error ~5e-7): opt-in "
                        "SumOfSines(dim, fast_math=True); not used for `value`"}
    if not args.no_unfused and world == 1 and (headline or name in ("boole6", "vegas4", "vegas8_cap4096")):
        u_steps = 2
        t_unf = sum(timed_steps(unfused, u_steps, 1, flush, barrier, min_warm_s=0.0))
        unf = {"value": evals() * u_steps / t_unf, "unit": "evals/s", "ms_per_step": t_unf / u_steps * 1e3,
               "path": "torch-callable integrand, points materialised in HBM (chunked), e2e through the public API"}
    if replicas:
        import torchquad_b200 as tq

        tq.distributed.enable()
        both = torch.tensor([float(n_evals), float(n_evals_e2e)], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(both)  # every replica does the whole job: the evaluations add up
        n_evals, n_evals_e2e = int(both[0]), int(both[1])
    if rank != 0:
        return None
    elt = 4 if wl["dtype"] == "float32" else 8
    value = n_evals * steps / t_fused
    rec = {
        "metric": "integrand evals/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": t_fused / steps * 1e3, "higher_is_better": True, "scaling": "weak" if replicas else wl["scaling"],
        "dtype": "f32" if wl["dtype"] == "float32" else "f64", "data": "synthetic",
        "config": {**workload_config(name, wl, "fused functor (generate+map+evaluate+accumulate in one kernel)"),
                   "l2": "512 MiB buffer rewritten between timed iterations", "rng": "Philox4x32-10, fresh seed per step"},
        "e2e": {"value": n_evals_e2e * e2e_steps / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": wl["dim"] * 2 * elt,
                "d2h_bytes_per_step": elt, "steps": e2e_steps,
                "call": f"torchquad_b200.{'MonteCarlo' if wl['kind']=='mc' else 'Boole' if wl['kind']=='boole' else 'VEGAS'}()"
                        ".integrate(fn, dim, N, integration_domain=<host list>, backend='torch') -> float"},
        "gpu_launches": launches, "evals_per_step": n_evals,
        "result": {"last_integral": result_check, "exact": info["exact"]},
    }
    if wl["kind"] == "vegas":
        v = info["integrator"]
        rec["config"].update(map_intervals=v.map.N_intervals, n_cubes=v.strat.N_cubes, iterations=v.it, map_cap=map_cap,
                             map_layout=("pairs + deferred histogram, band sweeps" if v.map.wants_records() and v._shard is None
                                         and v.map.sweep_group(v.strat.N_strat) >= 1 else "records" if v.map.wants_records() else "pairs"),
                             sharding=("block-cyclic cubes, one fp64 all-reduce per pass" if v._shard is not None
                                       else "independent replicas, one seed per rank (a 0.8 ms latency-bound run does not shard)"
                                       if replicas else "replicas" if world > 1 else "single GPU"))
        rec["result"]["error_estimate"] = float(v._get_error())
    clocks = sampler.summary(mark0, mark1) if sampler else None
    rec["clocks"] = clocks
    if world == 1:
        from torchquad_b200 import _lib

        sm_count = _lib.device_info()[0]
        sm_mhz = (clocks or {}).get("sm_mhz")
        if wl["kind"] == "mc":
            rec["roofline"] = issue_roofline("fused_mc_kernel<SUM_SIN,float>", "mc10:fused_mc_kernel", value, sm_count, sm_mhz)
            rec.update(mc_kernel_rooflines(wl, device))
        elif wl["kind"] == "boole":
            rec["roofline"] = issue_roofline("fused_nc_kernel<PROD_COS,double>", "boole6:fused_nc_kernel", value, sm_count, sm_mhz)
        else:
            if wl["N"] >= 10**8:
                rec["roofline"] = vegas_kernel_roofline(name, wl, info, device)
            else:
                rec["roofline"] = {"kernel": "fused_vegas_kernel + vegas_update_small_kernel (15 passes of ~6e4 samples)",
                                   "bound": "launch latency (a chain of ~30 dependent launches of 5-15 us)", "achieved": None,
                                   "peak": None, "unit": None, "frac": None, "traffic": None,
                                   "note": "no throughput roofline applies at this size; see profiles/r1/launches_vegas4.txt"}
    if fast:
        rec["fast_math"] = fast
    if unf:
        rec["unfused"] = unf
    return rec


def run_ours(args, names):
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    # stdout carries exactly ONE JSON line: everything else that reaches fd 1 (NCCL_DEBUG=INFO banners and channel
    # reports, whenever NCCL chooses to print them) goes to stderr, where the rank / transport evidence stays visible
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torchquad_b200 as tq
    from torchquad_b200 import _lib

    _lib.load()  # fail loudly if the CUDA library is missing
    barrier = lambda: None  # noqa: E731
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        dist.barrier()
        torch.cuda.synchronize()
        tq.distributed.enable()
        barrier = dist.barrier
    flush = torch.zeros(128 << 20, dtype=torch.float32, device=device)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx = {"device": device, "world": world, "rank": rank, "flush": flush, "barrier": barrier, "sampler": sampler}
    records = {}
    for i, name in enumerate(names):
        records[name] = measure(name, WORKLOADS[name], args, ctx, headline=(i == 0))
        torch.cuda.empty_cache()
    all_clocks = sampler.stop() if sampler else None
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    if world == 1 and not args.no_cpu_baseline:
        for i, name in enumerate(names):
            base = cpu_reference_rate(name, WORKLOADS[name], budget_s=10.0 if i == 0 else 3.0)
            base.pop("ms_per_step", None)
            records[name]["cpu_baseline"] = base
    head = records[names[0]]
    line = {**head, "vs_baseline": None}
    line["clocks"] = head.get("clocks") or all_clocks
    line["clocks_whole_run"] = all_clocks
    if len(names) > 1:
        line["workloads"] = {n: records[n] for n in names[1:]}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    names = [HEADLINE] + SUB_WORKLOADS if args.workload == "all" else [args.workload]
    if args.impl == "reference":
        run_reference(args, names)
    else:
        run_ours(args, names)


if __name__ == "__main__":
    main()
