"""print value / ms_per_step / e2e ms_per_step of a bench.py JSON line read from stdin (profiling aid)"""
import json
import sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(tag, round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["ms_per_step"], 1), d.get("warmup"))
