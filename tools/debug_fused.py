import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vcd_b200
from oracle.torch_vae import build_oracle, oracle_forward
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import rel_err
vcd_b200.add_src_to_path()
from models.sdxl_vae_wrapper import SDXLVAEWrapper
oracle = build_oracle(42).cuda()
res = {}
for fused in (False, True):
    w = SDXLVAEWrapper("random-init:42").cuda()
    w.vae.load_state_dict(oracle.state_dict())
    opt = torch.optim.AdamW(w.parameters(), lr=2e-3, weight_decay=0.0, fused=fused)
    torch.manual_seed(3)
    x = torch.rand(2, 3, 64, 64, device="cuda") * 2 - 1
    out = w(x, sample_posterior=False)
    total, _, _ = vcd_b200.vae_loss(out, x, 1e-6)
    total.backward()
    gn = {n: p.grad.detach().clone() for n, p in w.vae.named_parameters()}
    opt.step()
    opt.zero_grad(set_to_none=True)
    after = w(x, sample_posterior=False)["reconstruction"].detach()
    after2 = w(x, sample_posterior=False)["reconstruction"].detach()
    for m in w.vae.modules():
        if hasattr(m, "_packs"):
            m._packs.key = None
        if hasattr(m, "_up_packs"):
            m._up_packs.key = None
    after3 = w(x, sample_posterior=False)["reconstruction"].detach()
    ref = copy.deepcopy(oracle)
    ref.load_state_dict(w.vae.state_dict())
    with torch.no_grad():
        expect = oracle_forward(ref, x, False)["reconstruction"]
    res[fused] = ({n: p.detach().clone() for n, p in w.vae.named_parameters()}, gn)
    nan = any(not torch.isfinite(p).all() for p in w.vae.parameters())
    print(f"fused={fused}: ours vs oracle {rel_err(after, expect):.4f}; repeat {rel_err(after2, after):.2e}; after pack reset "
          f"{rel_err(after3, after):.2e}; nonfinite params {nan}; |rec| {float(after.abs().max()):.3f} oracle {float(expect.abs().max()):.3f}")
dp = max(float((res[True][0][n] - res[False][0][n]).abs().max()) for n in res[True][0])
dg = max(float((res[True][1][n] - res[False][1][n]).abs().max()) for n in res[True][1])
print("max |param(fused) - param(unfused)|", dp, " max |grad diff|", dg)
