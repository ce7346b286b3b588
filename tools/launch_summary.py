"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name (last `--last N` launches).
python tools/launch_summary.py gpurun_out/x.csv [--last N]"""
import csv
import re
import sys

path = sys.argv[1]
last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 0
rows = []
for d in csv.DictReader([l for l in open(path) if not l.startswith("==")]):
    if d.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((d["Kernel Name"].split("(")[0], float(d["Metric Value"].replace(",", "")) / 1000.0))
if last:
    rows = rows[-last:]
tot = sum(t for _, t in rows)
agg = {}
for n, t in rows:
    n = re.sub(r"ppv::|_kernel|void |\(anonymous namespace\)::", "", n)
    c = agg.setdefault(n, [0.0, 0])
    c[0] += t
    c[1] += 1
print(f"# {len(rows)} launches, total {tot:.1f} us")
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{k:40s} {c:5d} launches {v:10.1f} us {100 * v / tot:5.1f}%  avg {v / c:8.1f} us")
