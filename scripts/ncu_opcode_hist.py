"""Opcode histogram of an `ncu --page source --csv` dump (scripts/capture_ncu.py writes source_<key>.csv):
thread instructions per work unit and share of the stall samples per SASS opcode.
    python scripts/ncu_opcode_hist.py <source.csv> <units per launch> [top]"""
import collections
import csv
import sys

path, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
inst, samp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    words = r[ix["Source"]].split()
    op = (words[1] if words[0].startswith("@") else words[0]).rstrip(";")
    parts = op.split(".")
    keep2 = ("IMAD", "LDG", "RED", "ATOM", "LDS", "STS", "F2I", "I2F", "MUFU", "SHFL", "LDC")
    key = ".".join(parts[:2]) if parts[0] in keep2 else parts[0]
    inst[key] += int(r[ix["Instructions Executed"]])
    samp[key] += int(r[ix["# Samples"]])
total, stotal = sum(inst.values()), max(1, sum(samp.values()))
print(f"# {path}: {total / units:.2f} warp instructions = {32 * total / units:.0f} thread instructions per unit")
print(f"{'opcode':16s} {'inst/unit':>10s} {'share':>7s} {'stall samples':>14s}")
for k, v in inst.most_common(top):
    print(f"{k:16s} {32 * v / units:10.1f} {100 * v / total:6.1f}% {100 * samp[k] / stotal:13.1f}%")
