"""Run on the GPU box (one GPU):  python scripts/capture_ncu.py [out_dir] [key ...]
For every entry of CAPTURES: `ncu --set full` of the LAST launch of the named kernel in scripts/profile_target.py <target>,
a text summary (scripts/ncu_summary.py) and one record in <out_dir>/ncu_metrics.json that bench.py reads for its
rooflines (dram bytes and warp instructions per launch and per work unit)."""
import csv
import json
import os
import re
import subprocess
import sys

OUT = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2"
ONLY = set(sys.argv[2:])
os.makedirs(OUT, exist_ok=True)
REPS = os.environ.get("TQ_NCU_REP_DIR", "/tmp/tq_ncu_reps")
os.makedirs(REPS, exist_ok=True)
# key -> (profile_target argument, kernel regex, launches of that kernel BEFORE the one to capture)
CAPTURES = {
    "mc10:fused_mc_kernel": ("mc10", "fused_mc_kernel", 1),
    "boole6:fused_nc_kernel": ("boole6", "fused_nc_kernel", 1),
    "vegas8_cap4096:fused_vegas_kernel": ("vegas8_cap4096", "fused_vegas_kernel", 15),
    "vegas8:fused_vegas_kernel": ("vegas8", "fused_vegas_kernel", 15),
    "vegas8:hist_sweep_kernel": ("vegas8", "hist_sweep_kernel", 87),   # 10 passes x 8 dims + the 8 of the final pass: the last one
    "vegas16_cap4096:fused_vegas_tile_kernel": ("vegas16_cap4096", "fused_vegas_tile_kernel", 10),  # fp32: 10 passes + the final one
    "vegas8_unfused:sample_map_kernel": ("unfused8_cap", "sample_map_kernel", 14),
    "vegas8_unfused:accumulate_regen_kernel": ("unfused8_cap", "accumulate_regen_kernel", 14),
    "uniform_kernel_f32_d10": ("uniform_f32_d10", "uniform_kernel", 1),
    "sum1_kernel_f32": ("sum1_f32", "sum1_kernel", 1),
    "contract1_kernel_f64": ("contract1_f64", "contract1_kernel", 1),
}
PICK = {
    "time_ns": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
    "warp_inst": "smsp__inst_executed.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "registers": "launch__registers_per_thread",
    "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "lts_red_sectors": "lts__t_sectors_op_red.sum", "lts_sectors": "lts__t_sectors.sum",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
              "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}


def num(text, unit):
    v = float(text.replace(",", ""))
    return v * UNIT_SCALE.get(unit, 1.0)


metrics_path = os.path.join(OUT, "ncu_metrics.json")
metrics = json.load(open(metrics_path)) if os.path.exists(metrics_path) else {}
for key, (target, kernel, skip) in CAPTURES.items():
    if ONLY and key not in ONLY and target not in ONLY:
        continue
    rep = os.path.join(REPS, "prof_" + key.replace(":", "_"))  # .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB)
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{kernel}", "-s", str(skip),
           "-c", "1", "-f", "-o", rep, sys.executable, "scripts/profile_target.py", target]
    print("+", " ".join(cmd), flush=True)
    run = subprocess.run(cmd, capture_output=True, text=True)
    units = re.search(r"^UNITS (\d+)", run.stdout, re.M)
    if run.returncode != 0 or not units or not os.path.exists(rep + ".ncu-rep"):
        print("  FAILED", run.stdout[-1500:], run.stderr[-1500:], flush=True)
        continue
    units = int(units.group(1))
    raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, unit_row, vals = rows[0], rows[1], rows[-1]
    idx = {h: i for i, h in enumerate(hdr)}
    rec = {"kernel": vals[idx["Kernel Name"]][:160], "units_per_launch": units, "target": f"scripts/profile_target.py {target}"}
    for name, metric in PICK.items():
        if metric in idx and vals[idx[metric]] not in ("", "n/a"):
            rec[name] = num(vals[idx[metric]], unit_row[idx[metric]])
    rec["dram_bytes"] = rec.get("dram_read", 0.0) + rec.get("dram_write", 0.0)
    rec["dram_bytes_per_unit"] = rec["dram_bytes"] / units
    if "warp_inst" in rec:
        rec["warp_inst_per_unit"] = rec["warp_inst"] / units
    if "time_ns" in rec:
        rec["units_per_s_under_ncu"] = units / (rec["time_ns"] * 1e-9)
    metrics[key] = rec
    with open(os.path.join(OUT, "prof_" + key.replace(":", "_") + ".txt"), "w") as f:
        f.write(subprocess.run([sys.executable, "scripts/ncu_summary.py", rep + ".ncu-rep"], capture_output=True, text=True).stdout)
        f.write(f"\nUNITS per launch: {units}\n")
    if "vegas" in key:  # per-source-line stall picture of the VEGAS pass (needs -lineinfo; small CSV)
        src = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "source", "--csv"], capture_output=True, text=True).stdout
        with open(os.path.join(OUT, "source_" + key.replace(":", "_") + ".csv"), "w") as f:
            f.write(src)
    print("  ", json.dumps(rec), flush=True)
    json.dump(metrics, open(metrics_path, "w"), indent=1, sort_keys=True)
print("wrote", metrics_path)
