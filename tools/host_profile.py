"""Host-side (Python) profile of the training step: where the ~80 ms of enqueue time per step go.
  python tools/host_profile.py [--steps 3]"""
import argparse
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vcd_b200

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
vcd_b200.add_src_to_path()
from models.sdxl_vae_wrapper import SDXLVAEWrapper

w = SDXLVAEWrapper("random-init:42", torch_dtype=torch.bfloat16).cuda()
opt = torch.optim.AdamW(w.parameters(), lr=5e-5, fused=True)
x = torch.rand(a.batch, 3, a.res, a.res, device="cuda") * 2 - 1


def step():
    out = w(x, sample_posterior=True)
    total, rec, kl = vcd_b200.vae_loss(out, x, 1e-6)
    total.backward()
    torch.nn.utils.clip_grad_norm_(w.parameters(), 1.0)
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(a.steps):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
