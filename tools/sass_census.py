"""SASS mnemonic census of the built library: cuobjdump -sass, counted per kernel (runs without a GPU).
  python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "vae-channel-dynamics_b200", "libvcd_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = [("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("LDTM", r"\bLDTM"), ("UTMASTG", r"\bUTMASTG"), ("HMMA", r"\bHMMA"),
        ("F*2", r"\b(FFMA2|FADD2|FMUL2)\b"), ("ST128", r"\bSTG?\.E\.128"), ("LDGSTS", r"\bLDGSTS"), ("PREEXIT", r"\bPREEXIT")]
counts = collections.OrderedDict()
cur = None
it = iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        counts[cur] = [0] * len(cols)
        continue
    if cur is None:
        continue
    for i, (_, pat) in enumerate(cols):
        if re.search(pat, line):
            counts[cur][i] += 1
print("SASS mnemonic census of vae-channel-dynamics_b200/libvcd_b200.so (tools/sass_census.py: cuobjdump -sass, count per kernel).")
print("UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTMASTG = TMA store (none: epilogues store through "
      "shared-memory\nslabs + 128-bit stores; ST128 counts STG.E.128 and generic ST.E.128), HMMA = mma.sync (conv_small.cu only), F*2 = FFMA2/FADD2/FMUL2 packed fp32x2 arithmetic, "
      "LDGSTS = cp.async,\nPREEXIT = griddepcontrol.launch_dependents (programmatic dependent launch trigger).")
print(" ".join(f"{c:>7}" for c, _ in cols) + "  kernel")
for k in sorted(counts):
    if any(counts[k]):
        print(" ".join(f"{v:7d}" for v in counts[k]) + "  " + k)
