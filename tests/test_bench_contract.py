"""bench.py's reference arm runs without a GPU and prints ONE JSON line with the contract's keys (the driver
computes the speed-up from it); the workload table covers every BASELINE.json config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "vegas4",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "integrand evals/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "vegas4"


def test_workloads_cover_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench

    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(bench.WORKLOADS) >= len(base.get("configs", [])) >= 1
    kinds = {w["kind"] for w in bench.WORKLOADS.values()}
    assert {"mc", "vegas", "boole"} <= kinds
    mc = bench.WORKLOADS["mc10"]
    assert mc["dim"] == 10 and mc["N"] == 10**9 and mc["dtype"] == "float32"  # configs[1], the default workload


def test_rooflines_have_their_ncu_captures():
    """bench.py takes warp instructions per eval and DRAM bytes from profiles/r2/ncu_metrics.json (scripts/capture_ncu.py);
    a missing key would silently turn a workload's `roofline.frac` / `traffic` into null."""
    sys.path.insert(0, ROOT)
    import bench

    m = bench.ncu_metrics()
    for key in ("mc10:fused_mc_kernel", "boole6:fused_nc_kernel"):
        assert m[key]["warp_inst_per_unit"] > 0 and m[key]["units_per_launch"] > 0, key
    for key in ("vegas8_cap4096:fused_vegas_kernel", "vegas16_cap4096:fused_vegas_tile_kernel", "vegas8:fused_vegas_kernel",
                "vegas8:hist_sweep_kernel", "uniform_kernel_f32_d10", "sum1_kernel_f32"):
        assert m[key]["dram_bytes"] > 0 and m[key]["time_ns"] > 0, key
    # every sub-workload of the default run is a known workload, the headline is configs[1]
    assert bench.HEADLINE == "mc10" and set(bench.SUB_WORKLOADS) <= set(bench.WORKLOADS)
