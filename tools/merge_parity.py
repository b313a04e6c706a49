"""Merge the parity JSON a GPU run wrote (gpurun_out/r02_parity.json) into profiles/r02_parity.json (union by key: entries
of the newest run win; entries only an earlier run produced — e.g. the 2-GPU tests — are kept)."""
import json
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "gpurun_out", "r02_parity.json")
dst = os.path.join(root, "profiles", "r02_parity.json")
cur = json.load(open(dst)) if os.path.exists(dst) else {}
new = json.load(open(src))
cur.update(new)
json.dump(cur, open(dst, "w"), indent=1, sort_keys=True)
print(f"{len(new)} entries merged, {len(cur)} total")
