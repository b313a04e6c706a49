"""Profiling aid: GroupNorm+SiLU forward/backward on one tensor shape (ncu captures / CUDA-event timing).
  python tools/prof_gn.py --c 128 --h 512 --batch 8"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vcd_b200

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=128)
ap.add_argument("--h", type=int, default=512)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
ops = vcd_b200.ops
B, h, C = a.batch, a.h, a.c
xs = [torch.randn(B, h, h, C, device="cuda").to(torch.bfloat16).requires_grad_() for _ in range(3)]
g = torch.randn(B, h, h, C, device="cuda").to(torch.bfloat16)
gamma = torch.ones(C, device="cuda", dtype=torch.bfloat16, requires_grad=True)
beta = torch.zeros(C, device="cuda", dtype=torch.bfloat16, requires_grad=True)


def run(i):
    y, xid = ops.group_norm(xs[i % 3], gamma, beta, 32, 1e-6, True, None, None, True)
    torch.autograd.backward([y, xid], [g, g])


for i in range(2):
    run(i)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record()
for i in range(a.iters):
    run(i)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / a.iters
n = B * h * h * C
print(f"GN+SiLU fwd+bwd C={C} @{h} B={B}: {ms:.3f} ms; algorithmic (4+8 B/elem incl. skip grad) {12 * n / ms / 1e6:.0f} GB/s")
