"""Top warp-stall sites per kernel from `ncu -i X.ncu-rep --page source --csv` (captured with --import-source on).
  python tools/ncu_source_stalls.py gpurun_out/k1_src.csv [top_n]"""
import csv
import sys

path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kernels, cur, hdr = [], None, None
for row in csv.reader(open(path, newline="")):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        kernels.append(cur)
        hdr = None
    elif row[0] == "Address":
        hdr = row
    elif cur is not None and hdr is not None and len(row) >= len(hdr) - 2:
        cur["rows"].append(dict(zip(hdr, row)))
for k in kernels:
    rows = k["rows"]
    def f(r, c):
        try:
            return float(r.get(c) or 0)
        except ValueError:
            return 0.0
    total = sum(f(r, "Warp Stall Sampling (All Samples)") for r in rows)
    if total <= 0:
        continue
    reasons = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    by = sorted(((sum(f(r, c) for r in rows), c) for c in reasons), reverse=True)
    print(f"### {k['name'][:80]}: {int(total)} samples; " + ", ".join(f"{c[6:]} {100 * v / total:.0f}%" for v, c in by[:6]))
    for r in sorted(rows, key=lambda r: -f(r, "Warp Stall Sampling (All Samples)"))[:top_n]:
        s = f(r, "Warp Stall Sampling (All Samples)")
        why = max(reasons, key=lambda c: f(r, c))[6:]
        print(f"   {100 * s / total:5.1f}%  exec {int(f(r, 'Instructions Executed')):8d}  {r['Source'][:70]:70s} {why}")
