"""Profiling aid: GroupNorm+SiLU forward/backward on the model's tensor shapes (CUDA-event timing per kernel).
  python tools/prof_gn.py [--batch 8] [--shapes 128,512 256,256 512,128 512,64]     (C,H)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vcd_b200
from vcd_b200.ops import _p, _st, call, dtype_code

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--shapes", nargs="*", default=["128,512", "256,512", "256,256", "512,256", "512,128", "512,64"])
a = ap.parse_args()
B = a.batch
for sh in a.shapes:
    C, h = (int(v) for v in sh.split(","))
    n = B * h * h * C
    nbuf = min(6, max(2, int(300e6 // (n * 2)) + 1))
    xs = [torch.randn(B, h, h, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    gs = [torch.randn(B, h, h, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    rs = [torch.randn(B, h, h, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    out = torch.empty_like(xs[0])
    gamma = torch.ones(C, device="cuda", dtype=torch.bfloat16)
    beta = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
    sums = torch.empty(B * 32 * 2, dtype=torch.float64, device="cuda")
    dsdb = torch.empty(B * C * 2, dtype=torch.float32, device="cuda")
    colsum = torch.empty(C, dtype=torch.float32, device="cuda")
    pdt = dtype_code(gamma)
    slot, slot2 = vcd_b200.ops.TrackSlot(C, "cuda", 0.0), vcd_b200.ops.TrackSlot(C, "cuda", 0.0)
    hw = h * h
    call("vcd_gn_stats", _p(xs[0]), _p(sums), None, 0.0, B, hw, C, 32, _st())

    fns = {
        "stats": (lambda i: call("vcd_gn_stats", _p(xs[i % nbuf]), _p(sums), None, 0.0, B, hw, C, 32, _st()), 2),
        "apply": (lambda i: call("vcd_gn_apply_fwd", _p(xs[i % nbuf]), _p(sums), _p(gamma), _p(beta), pdt, _p(out), None, None,
                                 0.0, 1e-6, 1, B, hw, C, 32, _st()), 4),
        "apply+stats_out": (lambda i: call("vcd_gn_apply_fwd", _p(xs[i % nbuf]), _p(sums), _p(gamma), _p(beta), pdt, _p(out), None,
                                           _p(slot.raw), 0.0, 1e-6, 1, B, hw, C, 32, _st()), 4),
        "apply+stats_in+out": (lambda i: call("vcd_gn_apply_fwd", _p(xs[i % nbuf]), _p(sums), _p(gamma), _p(beta), pdt, _p(out),
                                              _p(slot2.raw), _p(slot.raw), 0.0, 1e-6, 1, B, hw, C, 32, _st()), 4),
        "bwd_reduce": (lambda i: call("vcd_gn_bwd_reduce", _p(xs[i % nbuf]), _p(gs[i % nbuf]), _p(sums), _p(gamma), _p(beta),
                                      pdt, _p(dsdb), 1e-6, 1, B, hw, C, 32, _st()), 4),
        "bwd_apply": (lambda i: call("vcd_gn_bwd_apply", _p(xs[i % nbuf]), _p(gs[i % nbuf]), _p(sums), _p(gamma), _p(beta), pdt,
                                     _p(dsdb), _p(out), None, _p(colsum), None, None, 1e-6, 1, B, hw, C, 32, _st()), 6),
        "bwd_apply+res": (lambda i: call("vcd_gn_bwd_apply", _p(xs[i % nbuf]), _p(gs[i % nbuf]), _p(sums), _p(gamma), _p(beta),
                                         pdt, _p(dsdb), _p(out), _p(rs[i % nbuf]), _p(colsum), None, None, 1e-6, 1, B, hw, C, 32, _st()), 8),
    }
    line = []
    for name, (fn, bpe) in fns.items():
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
        line.append(f"{name} {ms * 1e3:6.1f} us {bpe * n / ms / 1e6:5.0f} GB/s")
    print(f"C={C:4d} @{h:4d} B={B}: " + " | ".join(line), flush=True)
