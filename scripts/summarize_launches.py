"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += val * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"# {path}: {sum(cnt.values())} launches, {total:.3f} ms of kernel time")
print(f"{'kernel':90s} {'launches':>8s} {'ms':>12s} {'share':>7s} {'avg_us':>10s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{k[:90]:90s} {cnt[k]:8d} {v:12.3f} {100*v/total:6.1f}% {1e3*v/cnt[k]:10.1f}")
