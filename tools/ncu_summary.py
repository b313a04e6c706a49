"""Summarise an `ncu --set full` report (or its `--page raw --csv` export) for profiles/: per-launch key metrics as CSV + the JSON bench.py reads for
roofline.traffic.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_conv [dominant-kernel-regex]"""
import csv
import io
import json
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
dom = re.compile(sys.argv[3] if len(sys.argv) > 3 else r"umma_pair_kernel<256")
if rep.endswith(".csv"):      # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x_raw.csv
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


with open(out + "_launches.csv", "w", newline="") as f:
    wr = csv.writer(f)
    wr.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in idx])
    for r in rows[2:]:
        wr.writerow([re.sub(r"\(.*", "", r[i]).replace("void ", "").replace("<unnamed>::", "") if w == "Kernel Name" else r[i]
                     for w, i in idx])
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
ti, pi = hdr.index("gpu__time_duration.sum"), hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
sel = [r for r in rows[2:] if dom.search(r[ki])]
if sel:
    tr = [to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]) for r in sel]
    summary = {"report": rep, "dominant_kernel": dom.pattern, "launches": len(sel),
               "dominant_kernel_dram_bytes_per_launch": sum(tr) / len(tr),
               "dominant_kernel_time_us": sum(float(r[ti]) for r in sel) / len(sel),
               "dominant_kernel_tensor_pipe_active_pct": sum(float(r[pi]) for r in sel) / len(sel)}
    with open(out + "_summary.json", "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1))
