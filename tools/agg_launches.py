"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
   python tools/agg_launches.py gpurun_out/launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
for row in rows:
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name).replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.2f} ms total over {steps:g} steps -> {tot / steps:.2f} ms of kernel time per step")
print(f"{'ms/step':>9} {'share':>6} {'n/step':>7}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{v[1] / steps:9.3f} {100 * v[1] / tot:5.1f}% {v[0] / steps:7.1f}  {k[:100]}")
