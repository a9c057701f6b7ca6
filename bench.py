#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: integrand evaluations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (BASELINE.json `configs`):
  mc10      configs[1]  MonteCarlo, 10-D sum-of-sines, N=1e9 evals per GPU, fp32   (default; the metric config)
  vegas4    configs[0]  VEGAS 4-D Genz Gaussian N=1e6 fp64 (the reference's CPU-runnable case)
  boole6    configs[2]  Boole 6-D product-of-cosines, 33^6 points, fp64
  vegas8    configs[3]  VEGAS 8-D Genz oscillatory, fp64, 1e8-sample iterations
  vegas16   configs[4]  VEGAS 16-D Genz product-peak, fp32 fused, N=1e10 total
A "step" is one complete integration of the workload.  `value` times the fused functor path (inputs: none
beyond the 80-byte domain, resident in HBM); `e2e` times the same call through the public drop-in API with
the domain in host memory and the result read back to the host every step; `unfused` reports the
torch-callable path (points materialised in HBM) with its own HBM roofline.
Under torchrun each rank owns one GPU and a disjoint row range of the same Philox stream (weak scaling:
N per GPU fixed); time is the max over ranks of CUDA-event time.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "mc10": dict(kind="mc", dim=10, N=10**9, dtype="float32", integrand="sum_sin"),
    "vegas4": dict(kind="vegas", dim=4, N=10**6, dtype="float64", integrand="genz_gaussian"),
    "boole6": dict(kind="boole", dim=6, N=33**6, dtype="float64", integrand="prod_cos"),
    "vegas8": dict(kind="vegas", dim=8, N=2_500_000_000, dtype="float64", integrand="genz_oscillatory"),
    "vegas16": dict(kind="vegas", dim=16, N=10**10, dtype="float32", integrand="genz_product_peak"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mc10", choices=sorted(WORKLOADS))
    ap.add_argument("--no-unfused", action="store_true", help="skip the unfused torch-callable arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--map-cap", type=int, default=None,
                    help="VEGAS workloads: cap the map at this many intervals per dimension (VEGAS.max_map_intervals); "
                         "default keeps the reference's Ni = N/250")
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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) > 8 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


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


def timed_steps(step, steps, warmup, flush_buf, barrier):
    """Per-step CUDA-event times (seconds) on the current stream; L2 flushed between iterations."""
    # W warm-up steps, and at least 0.3 s of them: the GPU drops its clocks while the host sets things up (e.g. waits
    # for nvidia-smi), and a millisecond-scale step would otherwise be timed on the ramp
    # The number of extra steps is derived from a rank-agreed time (steps contain collectives under torchrun).
    t0 = time.time()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    spent = agree_max(time.time() - t0)
    if spent < 0.3:
        for _ in range(min(200, int(math.ceil((0.3 - spent) / max(spent / max(warmup, 1), 1e-4))))):
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


# ------------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_runner(wl):
    """(run, sample description): one bounded repetition of the reference algorithm on the host cores
    (oracle port: torch CPU ops = the reference's own arithmetic).  run() returns the evaluations it did."""
    from oracle import ref_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    dt = getattr(torch, wl["dtype"])
    fn = torch_callable(wl["integrand"])
    dim = wl["dim"]
    dom = torch.tensor([[0.0, 1.0]] * dim, dtype=dt)
    if wl["kind"] == "mc":
        n = 10**7

        def run():
            O.mc_integrate(fn, dim, n, dom, seed=0)
            return n

        return run, f"MonteCarlo {dim}-D N={n:.0e} per repetition"
    if wl["kind"] == "boole":
        npd = 17

        def run():
            pts, hs, n_ = O.nc_grid("boole", npd**dim, dom)
            O.nc_result("boole", fn(pts), dim, n_, hs)
            return npd**dim

        return run, f"Boole {dim}-D n={npd} per dim per repetition"
    n = min(wl["N"], 2 * 10**6 if dim > 4 else 10**6)

    def run():
        g = torch.Generator().manual_seed(0)
        r = O.VegasRun(fn, dim, n, dom, lambda size, dtype: torch.rand(size, dtype=dtype, generator=g))
        r.run()
        return r.fevals

    return run, f"VEGAS {dim}-D N={n:.0e} per repetition"


def cpu_reference_rate(wl, budget_s=12.0, steps=None, warmup=1):
    """Time the reference algorithm on the host: `steps` repetitions when given, else a time budget."""
    run, sample = cpu_reference_runner(wl)
    for _ in range(max(1, warmup)):
        run()  # warm-up (thread pool, allocator)
    evals, reps, t0 = 0, 0, time.perf_counter()
    while (reps < steps) if steps else (time.perf_counter() - t0 < budget_s or reps < 2):
        evals += run()
        reps += 1
    dt_s = time.perf_counter() - t0
    return {"value": evals / dt_s, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample} x {reps} repetitions ({dt_s:.1f} s); oracle/ref_oracle.py restates the reference over the same ATen CPU kernels",
            "ms_per_step": dt_s / reps * 1e3}


def run_reference(args, wl):
    rank, _, world = dist_env()
    if rank != 0:
        return
    base = cpu_reference_rate(wl, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "integrand evals/s", "value": base["value"], "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base.pop("ms_per_step"),
        "higher_is_better": True, "scaling": "strong" if wl["kind"] == "boole" else "weak", "vs_baseline": None,
        "dtype": "f32" if wl["dtype"] == "float32" else "f64", "data": "synthetic",
        "config": {"workload": args.workload, "kind": wl["kind"], "dim": wl["dim"],
                   ("N_total" if wl["kind"] == "boole" else "N_per_gpu"): wl["N"],
                   "integrand": wl["integrand"], "path": "reference algorithm on the host cores, bounded sample per step"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def build_steps(wl, device, world, map_cap=None):
    """Returns (fused_step, e2e_step, unfused_step, evals_per_step_per_job, info)."""
    import torchquad_b200 as tq

    dt = getattr(torch, wl["dtype"])
    dim = wl["dim"]
    fn = make_integrand(wl["integrand"], dim)
    call = torch_callable(wl["integrand"])
    dom_host = [[0.0, 1.0]] * dim
    dom_dev = torch.tensor(dom_host, dtype=dt, device=device)
    N = wl["N"] * world  # weak scaling: per-GPU work fixed
    state = {"evals": N, "seed": 0}
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
            return integ.integrate(fn, dim, N=wl["N"], integration_domain=dom_dev)

        def e2e():
            return float(integ.integrate(fn, dim, N=wl["N"], integration_domain=dom_host, backend="torch"))

        def unfused():
            return float(integ.integrate(call, dim, N=wl["N"], integration_domain=dom_host, backend="torch"))

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
    info = {"dtype": dt, "dim": dim, "exact": fn.exact(), "integrator": integ}
    if wl["kind"] == "mc":
        fast_fn = make_integrand(wl["integrand"], dim, fast_math=True)

        def fused_fast():
            state["seed"] += 1
            return integ.integrate(fast_fn, dim, N=N, integration_domain=dom_dev, seed=state["seed"])

        info["fused_fast"] = fused_fast
    return fused, e2e, unfused, evals, info


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures summarised in
# profiles/r1 (same launch shapes as kernel_rooflines below / the default workload).
NCU_TRAFFIC = {
    "fused_mc_kernel": 28_160,            # profiles/r1/prof_fused_mc.txt (1e9 evals: no sample traffic)
    "sum1_kernel": 865_042_944,           # profiles/r1/prof_sum1.txt (2e8 fp32 values = 800 MB algorithmic)
    "uniform_kernel": 8_028_438_320,      # profiles/r1/prof_uniform_f32_d10.txt (2e8 x 10 fp32 = 8.0 GB algorithmic)
}


# smsp__inst_executed.sum / evaluations from the same captures (10-D sum-of-sines, fp32: 1.772e10 / 1e9)
NCU_WARP_INST_PER_EVAL = {"fused_mc_kernel": 17.72}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def kernel_rooflines(wl, device, clocks_mhz):
    """Live CUDA-event timings of the dominant kernels alone + measured issue-rate denominators."""
    import ctypes

    from torchquad_b200 import _lib, ops

    out = {}
    peaks, peak_src = measured_peaks()
    dt = getattr(torch, wl["dtype"])
    dim = wl["dim"]

    def time_call(fn, reps=5):
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

    # measured instruction-issue denominators (FP32 FMA chains, FP64 FMA chains, Philox blocks)
    sink = torch.zeros(1, dtype=torch.float64, device=device)
    micro = {}
    for kind, name in [(0, "fp32_fma_per_s"), (1, "fp64_fma_per_s"), (2, "philox_blocks_per_s")]:
        ops_out = ctypes.c_double()
        t = time_call(lambda: _lib.call("tq_peak_microbench", kind, 20000 if kind != 2 else 2000, sink.data_ptr(),
                                        ctypes.byref(ops_out), _lib.stream_ptr(device)))
        micro[name] = ops_out.value / t
    out["microbench"] = micro
    if wl["kind"] == "mc":
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
            "frac": bytes_alg / t / 1e9 / peaks["hbm_gbs"], "traffic": NCU_TRAFFIC.get("uniform_kernel"),
            "peak_source": peak_src, "launch_ms": t * 1e3, "algorithmic_bytes_per_launch": bytes_alg,
        }
        del buf
        f = torch.rand(rows, dtype=dt, device=device)
        t = time_call(lambda: ops.sum_columns(f))
        out["roofline_reduce"] = {
            "kernel": "sum1_kernel<T> (tq_sum_columns: fp64-accumulated reduction of f)", "bound": "hbm",
            "achieved": rows * f.element_size() / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": rows * f.element_size() / t / 1e9 / peaks["hbm_gbs"], "traffic": NCU_TRAFFIC.get("sum1_kernel"),
            "peak_source": peak_src,
            "launch_ms": t * 1e3,
        }
    return out


def run_ours(args, wl):
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torchquad_b200 as tq
    from torchquad_b200 import _lib

    _lib.load()  # fail loudly if the CUDA library is missing
    barrier = lambda: None  # noqa: E731
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION/INFO; stdout must carry ONE JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("TQ_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"
        # ... and whatever still reaches fd 1 during communicator set-up goes to stderr instead
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        tq.distributed.enable()
        barrier = dist.barrier
    flush = torch.zeros(128 << 20, dtype=torch.float32, device=device)
    fused, e2e, unfused, evals, info = build_steps(wl, device, world, args.map_cap)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    times = timed_steps(fused, args.steps, args.warmup, flush, barrier)
    launches = timed_steps.launches  # kernels libtqb200 launched inside the timed region (tq_kernel_launches)
    clocks = sampler.stop() if rank == 0 else None
    t_fused = max_over_ranks(sum(times), device, world)
    n_evals = evals()
    result_check = float(fused())

    t_e2e = max_over_ranks(sum(timed_steps(e2e, max(2, args.steps // 2), 1, flush, barrier)), device, world)
    e2e_steps = max(2, args.steps // 2)
    n_evals_e2e = evals()

    fast = None
    if "fused_fast" in info:
        f_steps = max(2, args.steps // 2)
        t_fast = max_over_ranks(sum(timed_steps(info["fused_fast"], f_steps, 1, flush, barrier)), device, world)
        fast = {"value": evals() * f_steps / t_fast, "unit": "evals/s", "ms_per_step": t_fast / f_steps * 1e3,
                "last_integral": float(info["fused_fast"]()),
                "note": "same fused kernel with sin() evaluated by the SFU (__sinf, abs error ~5e-7): opt-in "
                        "SumOfSines(dim, fast_math=True); not used for `value`"}
    unf = None
    if not args.no_unfused and wl["kind"] in ("mc", "boole", "vegas") and wl["N"] <= 3 * 10**9:
        u_steps = 2
        t_unf = max_over_ranks(sum(timed_steps(unfused, u_steps, 1, flush, barrier)), device, world)
        unf = {"value": evals() * u_steps / t_unf, "unit": "evals/s", "ms_per_step": t_unf / u_steps * 1e3,
               "path": "torch-callable integrand, points materialised in HBM (chunked), e2e through the public API"}
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    roof = kernel_rooflines(wl, device, clocks) if world == 1 else {}
    micro = roof.get("microbench", {})
    elt = 4 if wl["dtype"] == "float32" else 8
    line = {
        "metric": "integrand evals/s", "value": n_evals * args.steps / t_fused, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_fused / args.steps * 1e3, "higher_is_better": True,
        # MC / VEGAS: N per GPU fixed (weak); the Newton-Cotes grid is one fixed grid sharded over the ranks (strong)
        "scaling": "strong" if wl["kind"] == "boole" else "weak", "vs_baseline": None,
        "dtype": "f32" if wl["dtype"] == "float32" else "f64", "data": "synthetic",
        "config": {"workload": args.workload, "kind": wl["kind"], "dim": wl["dim"],
                   ("N_total" if wl["kind"] == "boole" else "N_per_gpu"): wl["N"],
                   "integrand": wl["integrand"], "path": "fused functor (generate+evaluate+accumulate in one kernel)",
                   "l2": "512 MiB buffer rewritten between timed iterations", "rng": "Philox4x32-10, fresh seed per step"},
        "clocks": clocks,
        "e2e": {"value": n_evals_e2e * e2e_steps / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": wl["dim"] * 2 * elt,
                "d2h_bytes_per_step": elt, "steps": e2e_steps,
                "call": f"torchquad_b200.{'MonteCarlo' if wl['kind']=='mc' else 'Boole' if wl['kind']=='boole' else 'VEGAS'}().integrate(fn, dim, N, integration_domain=<host list>, backend='torch') -> float"},
        "gpu_launches": launches,
        "result": {"last_integral": result_check, "exact": info["exact"]},
    }
    if wl["kind"] == "mc" and micro:
        # Fused kernel: no sample traffic, bound by instruction issue (INT32 Philox + FP32 sin).  Work model per
        # eval: ceil(dim/4) Philox blocks; denominator: measured Philox-block rate of this GPU.
        blocks_per_eval = -(-wl["dim"] // (4 if elt == 4 else 2))
        ach = n_evals * args.steps / t_fused * blocks_per_eval
        line["roofline"] = {"kernel": "fused_mc_kernel<SUM_SIN,float>", "bound": "int32/fp32 issue (no tensor, no HBM traffic)",
                            "achieved": ach / 1e9, "peak": micro["philox_blocks_per_s"] / 1e9, "unit": "G Philox blocks/s",
                            "frac": ach / micro["philox_blocks_per_s"], "traffic": NCU_TRAFFIC.get("fused_mc_kernel"),
                            "note": "peak = Philox-only microbenchmark on this GPU; the kernel also evaluates dim sin() per eval"}
        # Second view of the same kernel: instruction issue.  Warp instructions per eval come from the ncu capture
        # of this kernel (profiles/r1/prof_fused_mc.txt: smsp__inst_executed.sum / evals); the denominator is
        # 4 schedulers x SMs x the SM clock sampled during the timed region.
        if elt == 4 and wl["dim"] == 10 and clocks and clocks.get("sm_mhz"):
            import ctypes as _ct

            sm, maj, mnr = _ct.c_int(), _ct.c_int(), _ct.c_int()
            _lib.call("tq_device_info", _ct.byref(sm), _ct.byref(maj), _ct.byref(mnr))
            winst = NCU_WARP_INST_PER_EVAL["fused_mc_kernel"] * n_evals * args.steps / t_fused
            peak_issue = 4.0 * sm.value * clocks["sm_mhz"] * 1e6
            line["roofline"]["issue"] = {"warp_inst_per_eval": NCU_WARP_INST_PER_EVAL["fused_mc_kernel"],
                                         "achieved_gwarp_inst_s": winst / 1e9, "peak_gwarp_inst_s": peak_issue / 1e9,
                                         "frac": winst / peak_issue}
    if wl["kind"] == "vegas":
        line["config"]["map_intervals"] = info["integrator"].map.N_intervals
        line["config"]["n_cubes"] = info["integrator"].strat.N_cubes
        line["config"]["iterations"] = info["integrator"].it
    if roof:
        line.update({k: v for k, v in roof.items()})
    if fast:
        line["fast_math"] = fast
    if unf:
        line["unfused"] = unf
    if not args.no_cpu_baseline and world == 1:
        base = cpu_reference_rate(wl)
        base.pop("ms_per_step", None)
        line["cpu_baseline"] = base
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
