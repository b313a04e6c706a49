"""Per-pass CUDA-event timing of the conv entry points (fprop / dgrad / wgrad) on the model's dominant shapes.
  python tools/prof_conv2.py [--batch 8] [--shapes 128,128,512 256,256,256 ...]   (Cin,Cout,H ; 3x3 s1)
Inputs are rotated through > 126 MB so nothing is L2-resident."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vcd_b200
from vcd_b200.ops import _p, _st, call, dtype_code

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--shapes", nargs="*", default=["128,128,512", "256,256,256", "512,512,128", "512,512,64", "256,128,512",
                                                 "512,256,256"])
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--gn-sums", action="store_true", help="fprop also emits the GroupNorm sums of its output from the epilogue "
                "(as every conv that feeds a GroupNorm does in the model)")
a = ap.parse_args()
lib = vcd_b200._lib.lib()
B, k = a.batch, a.k
pad = 1 if k == 3 else 0
for sh in a.shapes:
    ci, co, h = (int(v) for v in sh.split(","))
    nbuf = min(8, max(2, int(300e6 // (B * h * h * max(ci, co) * 2)) + 1))
    xs = [torch.randn(B, h, h, ci, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    gs = [torch.randn(B, h, h, co, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    w = (torch.randn(co, ci, k, k, device="cuda") * 0.02).to(torch.bfloat16)
    bias = torch.zeros(co, device="cuda", dtype=torch.bfloat16)
    packs = vcd_b200.ops.PackedWeights()
    wf, wd, b32 = packs.get(w, bias)
    y = torch.empty(B, h, h, co, dtype=torch.bfloat16, device="cuda")
    dx = torch.empty(B, h, h, ci, dtype=torch.bfloat16, device="cuda")
    dw = torch.empty_like(w)
    db = torch.empty_like(bias)
    ws = torch.empty(lib.vcd_conv2d_wgrad_ws_bytes(B, h, h, ci, co, k, k, 1) // 4, dtype=torch.float32, device="cuda")
    colsum = torch.zeros(co, dtype=torch.float32, device="cuda")

    osums = torch.empty(B * 32 * 2, dtype=torch.float64, device="cuda") if a.gn_sums else None

    def fprop(i):
        call("vcd_conv2d_fprop", _p(xs[i % nbuf]), _p(wf), _p(b32), None, _p(y), None, B, h, h, ci, co, k, k, 1, pad, pad, h, h,
             0, 0, _p(osums), 32 if a.gn_sums else 0, _st())

    def dgrad(i):
        call("vcd_conv2d_dgrad", _p(gs[i % nbuf]), _p(wf), _p(wd), _p(dx), None, B, h, h, ci, co, k, k, 1, pad, pad, h, h, 0, 0,
             _st())

    def wgrad(i):
        call("vcd_conv2d_wgrad", _p(xs[i % nbuf]), _p(gs[i % nbuf]), _p(dw), _p(db), _p(colsum), dtype_code(w), _p(ws), B, h, h,
             ci, co, k, k, 1, pad, pad, h, h, 0, 0, _st())

    # dgrad with the fused GroupNorm backward prologue (3x3 only)
    gsums = torch.empty(B * 32 * 2, dtype=torch.float64, device="cuda")
    call("vcd_gn_stats", _p(xs[0]), _p(gsums), None, 0.0, B, h * h, ci, 32, _st())
    gamma = torch.ones(ci, device="cuda", dtype=torch.bfloat16)
    beta = torch.zeros(ci, device="cuda", dtype=torch.bfloat16)
    dsdb = torch.empty(B * ci * 2, dtype=torch.float32, device="cuda")
    ab = torch.empty(B * ci * 2, dtype=torch.float32, device="cuda")

    def dgrad_gn(i):
        call("vcd_conv2d_dgrad_gn", _p(gs[i % nbuf]), _p(wd), _p(dx), B, h, h, ci, co, k, k, pad, pad, _p(xs[i % nbuf]),
             _p(gsums), _p(gamma), _p(beta), dtype_code(gamma), 32, 1e-6, 1, _p(dsdb), _p(ab), _st())

    fl = 2.0 * B * h * h * co * ci * k * k
    out = []
    passes = [("fprop", fprop), ("dgrad", dgrad), ("wgrad", wgrad)]
    if k == 3 and lib.vcd_conv2d_dgrad_gn_supported(B, h, h, ci, co, k, k, 1):
        passes.append(("dgrad+gn", dgrad_gn))
    for name, fn in passes:
        for i in range(int(os.environ.get('VCD_PROF_WARM', '2'))):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(a.iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        out.append(f"{name} {ms:.3f} ms {fl / ms / 1e9:6.0f} TF")
    print(f"{ci:4d}->{co:4d} k{k} @{h:4d} B={B}: " + " | ".join(out), flush=True)
