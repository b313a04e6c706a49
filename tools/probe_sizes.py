import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import vcd_b200 as vcd
from oracle.torch_vae import build_oracle, oracle_forward, oracle_losses
from util import rel_err
oracle = build_oracle(42).cuda()
model = vcd.B200AutoencoderKL().cuda(); model.load_state_dict(oracle.state_dict())
for (H, W, B) in [(8, 8, 2), (16, 24, 1), (32, 32, 5), (8, 1024, 1), (24, 8, 3), (100, 100, 1), (36, 44, 2), (520, 520, 1)]:
    torch.manual_seed(0)
    x = torch.rand(B, 3, H, W, device='cuda') * 2 - 1
    try:
        oo = oracle_forward(oracle, x, False)
        oshape = tuple(oo['reconstruction'].shape)
    except Exception as e:
        oo = None; oshape = 'oracle fails: ' + str(e)[:60]
    try:
        model.zero_grad(set_to_none=True)
        d = model.encode(x).latent_dist
        rec = model.decode(d.mode()).sample
        msg = f"ours {tuple(rec.shape)}"
        if oo is not None and tuple(rec.shape) == oshape:
            msg += f" rec err {rel_err(rec, oo['reconstruction']):.3e} mean err {rel_err(d.mean, oo['latent_dist'].mean):.3e}"
            if rec.shape == x.shape:
                vcd.vae_loss({"reconstruction": rec, "latent_dist": d}, x, 1e-6)[0].backward()
                oracle.zero_grad(set_to_none=True)
                oracle_losses(oo, x, 1e-6)[0].backward()
                og = dict(oracle.named_parameters()); big = max(float(p.grad.norm()) for p in og.values())
                named = dict(model.named_parameters())
                e = torch.tensor([rel_err(named[n].grad, p.grad) for n, p in og.items() if float(p.grad.norm()) > 1e-4 * big])
                msg += f" grad median {float(e.median()):.3e} max {float(e.max()):.3e}"
    except Exception as e:
        msg = "ours fails: " + type(e).__name__ + ": " + str(e)[:200]
    print((H, W, B), "oracle", oshape, "|", msg, flush=True)
