"""Profiling aid: run fprop/dgrad/wgrad of one conv shape a few times (for ncu captures and quick CUDA-event timing).
  python tools/prof_conv.py --cin 128 --cout 128 --h 512 --batch 8 [--k 3] [--iters 5] [--residual]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vcd_b200

ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=128)
ap.add_argument("--cout", type=int, default=128)
ap.add_argument("--h", type=int, default=512)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--residual", action="store_true")
ap.add_argument("--fwd-only", action="store_true")
a = ap.parse_args()
ops = vcd_b200.ops
B, h, ci, co, k = a.batch, a.h, a.cin, a.cout, a.k
nbuf = 3
xs = [torch.randn(B, h, h, ci, device="cuda").to(torch.bfloat16).requires_grad_() for _ in range(nbuf)]
res = torch.randn(B, h, h, co, device="cuda").to(torch.bfloat16) if a.residual else None
w = (torch.randn(co, ci, k, k, device="cuda") * 0.02).to(torch.bfloat16).requires_grad_()
bias = torch.zeros(co, device="cuda", dtype=torch.bfloat16).requires_grad_()
packs = ops.PackedWeights()
g = torch.randn(B, h, h, co, device="cuda").to(torch.bfloat16)
pad = 1 if k == 3 else 0


def run(i):
    y = ops.conv2d(xs[i % nbuf], w, bias, packs, stride=1, pad_t=pad, pad_l=pad, residual=res)
    if not a.fwd_only:
        y.backward(g)


for i in range(2):
    run(i)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(a.iters):
    run(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
fl = (1 if a.fwd_only else 3) * 2.0 * B * h * h * co * ci * k * k
print(f"{ci}->{co} k{k} @{h} B={B} residual={a.residual} fwd_only={a.fwd_only}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
