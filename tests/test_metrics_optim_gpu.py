"""GPU parity of the SURVEY 8(f) kernels: fused clip+AdamW (f2) against torch.optim.AdamW + clip_grad_norm_, on-device
PSNR / SSIM (f4) against a plain-torch restatement of torchmetrics' published algorithm, uint8 preprocessing (f1)
against the reference's own PIL transform chain (data_utils.py:13-30)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import record_parity, rel_err

pytestmark = pytest.mark.gpu


def _torch_ssim(pred, target, data_range=1.0, k=11, sigma=1.5):
    """[upstream] torchmetrics functional SSIM: reflect-pad, depth-wise Gaussian conv of (x, y, xx, yy, xy), crop the pad,
    per-image mean."""
    C = pred.shape[1]
    d = torch.arange(k, dtype=torch.float64, device=pred.device) - (k - 1) / 2
    g = torch.exp(-0.5 * (d / sigma) ** 2)
    g = (g / g.sum())
    w = (g[:, None] * g[None, :]).to(pred.dtype).expand(C, 1, k, k).contiguous()
    pad = (k - 1) // 2
    p, t = F.pad(pred, (pad,) * 4, mode="reflect"), F.pad(target, (pad,) * 4, mode="reflect")
    x = torch.cat([p, t, p * p, t * t, p * t])
    o = F.conv2d(x, w, groups=C)
    mu_p, mu_t, pp, tt, pt = o.split(pred.shape[0])
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s_p, s_t, s_pt = pp - mu_p ** 2, tt - mu_t ** 2, pt - mu_p * mu_t
    m = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (s_p + s_t + c2))
    m = m[..., pad:-pad, pad:-pad]
    return m.reshape(m.shape[0], -1).mean(-1)


@pytest.mark.parametrize("shape", [(3, 3, 64, 64), (2, 3, 75, 41), (1, 1, 11, 11), (4, 3, 512, 512)])
def test_psnr_ssim_match_torchmetrics_algorithm(vcd, shape):
    torch.manual_seed(0)
    t = torch.rand(*shape, device="cuda", dtype=torch.float64)
    low = F.interpolate(F.avg_pool2d(t, 4, ceil_mode=True), size=shape[2:], mode="bilinear")   # structured images
    t = (0.7 * low + 0.3 * t).clamp(0, 1)
    p = (t + 0.05 * torch.randn_like(t)).clamp(0, 1)
    psnr, ssim = vcd.metrics.PeakSignalNoiseRatio(data_range=1.0), vcd.metrics.StructuralSimilarityIndexMeasure(
        data_range=1.0, gaussian_kernel=True, sigma=1.5, kernel_size=11)
    want_ssim, sse, n, imgs = 0.0, 0.0, 0, 0
    for lo in range(0, shape[0], 2):               # several updates, like evaluate.py's batches
        pb, tb = p[lo:lo + 2], t[lo:lo + 2]
        psnr.update(pb.float(), tb.float())
        ssim.update(pb.float(), tb.float())
        want_ssim += float(_torch_ssim(pb, tb).sum())
        sse += float(((pb.float().double() - tb.float().double()) ** 2).sum())
        n += pb.numel()
        imgs += pb.shape[0]
    got_psnr, got_ssim = float(psnr.compute()), float(ssim.compute())
    want_psnr = 10 * math.log10(1.0 / (sse / n))
    record_parity(f"psnr_ssim {shape}", {"psnr": got_psnr, "psnr_ref": want_psnr, "ssim": got_ssim, "ssim_ref": want_ssim / imgs})
    assert abs(got_psnr - want_psnr) < 1e-3
    assert abs(got_ssim - want_ssim / imgs) < 2e-5


def test_preprocess_u8_matches_the_reference_transform_chain(vcd):
    """data_utils.py:13-30 on CIFAR-sized inputs (32 -> 64: up-sampling, where PIL's bilinear filter has no anti-aliasing
    support widening).  PIL rounds the resized image back to uint8 before ToTensor: agreement within one grey level."""
    from PIL import Image
    from torchvision import transforms as T
    imgs = vcd.data.synthetic_images(4, 32, "smooth", seed=1)
    tf = T.Compose([T.Resize(64, interpolation=T.InterpolationMode.BILINEAR), T.CenterCrop(64), T.ToTensor(), T.Normalize([0.5], [0.5])])
    want = torch.stack([tf(Image.fromarray(a)) for a in imgs])
    got = vcd.data.preprocess_uint8_batch(torch.from_numpy(imgs).cuda(), 64).cpu()
    assert got.shape == want.shape == (4, 3, 64, 64)
    err = float((got - want).abs().max())
    record_parity("preprocess_u8 32->64", {"max_abs_err": err, "one_grey_level": 2 / 255})
    assert err <= 1.01 * 2 / 255
    # non-square source: shorter side -> R, centre crop
    rect = np.ascontiguousarray(vcd.data.synthetic_images(2, 48, "smooth", seed=2)[:, :32])      # [2, 32, 48, 3]
    want = torch.stack([tf(Image.fromarray(a)) for a in rect])
    got = vcd.data.preprocess_uint8_batch(torch.from_numpy(rect).cuda(), 64).cpu()
    assert float((got - want).abs().max()) <= 1.01 * 2 / 255


def _make_params(dtype, misalign):
    torch.manual_seed(3)
    shapes = [(128, 3, 3, 3), (128,), (3,), (256, 128, 1, 1), (512, 512, 3, 3), (512,), (8, 512, 3, 3), (1031,), (5,)]
    ps = [torch.nn.Parameter((torch.randn(s, device="cuda") * 0.05).to(dtype)) for s in shapes]
    flat = torch.empty(sum(p.numel() for p in ps) + 64, dtype=dtype, device="cuda") if misalign else None
    return ps, flat


def _set_grads(ps, flat, gen, scale):
    off = 3 if flat is not None else 0          # gradient views at odd element offsets (DDP bucket views)
    for p in ps:
        g = (torch.randn(p.shape, device="cuda", generator=gen) * scale).to(p.dtype)
        if flat is not None:
            v = flat[off:off + p.numel()].view_as(p)
            v.copy_(g)
            p.grad = v
            off += p.numel()
        else:
            p.grad = g


@pytest.mark.parametrize("dtype,misalign", [(torch.float32, False), (torch.bfloat16, False), (torch.bfloat16, True), (torch.float32, True)])
def test_fused_clip_adamw_matches_torch_over_10_steps(vcd, dtype, misalign):
    """fp32 parameters: against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW themselves.  bf16 parameters: torch keeps
    bf16 moments, the fused kernel fp32 moments — the reference trajectory is torch AdamW on an fp32 master copy whose
    parameters are rounded to bf16 after every step (same arithmetic, fp32 moments)."""
    ps, flat = _make_params(dtype, misalign)
    ref = [torch.nn.Parameter(p.detach().float().clone()) for p in ps]
    hp = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    topt = torch.optim.AdamW(ref, **hp)
    fopt = vcd.FusedClipAdamW.from_torch(torch.optim.AdamW(ps, **hp))
    sched = torch.optim.lr_scheduler.LambdaLR(fopt.adopted, lambda s: min(1.0, (s + 1) / 4))     # built on the ADOPTED optimizer
    tsched = torch.optim.lr_scheduler.LambdaLR(topt, lambda s: min(1.0, (s + 1) / 4))
    gen = torch.Generator(device="cuda").manual_seed(11)
    worst_norm = 0.0
    for step in range(10):
        _set_grads(ps, flat, gen, scale=0.5 if step % 2 else 0.01)       # alternately clipped / not clipped
        for r, p in zip(ref, ps):
            r.grad = p.grad.detach().float().clone()
        if step == 4:                                                     # a parameter without gradient is skipped
            ps[3].grad = None
            ref[3].grad = None
        n_ref = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        n_got = fopt.clip_grad_norm_(ps, 1.0)
        worst_norm = max(worst_norm, abs(float(n_got) - float(n_ref)) / float(n_ref))
        topt.step()
        fopt.step()
        tsched.step()
        sched.step()
        if dtype == torch.bfloat16:
            with torch.no_grad():
                for r in ref:
                    r.copy_(r.to(torch.bfloat16).float())
        fopt.zero_grad(set_to_none=True)
        topt.zero_grad(set_to_none=True)
    err = max(rel_err(p.detach(), r.detach()) for p, r in zip(ps, ref))
    m_err = max(rel_err(fopt.state[p]["exp_avg"], topt.state[r]["exp_avg"]) for p, r in zip(ps, ref))
    v_err = max(rel_err(fopt.state[p]["exp_avg_sq"], topt.state[r]["exp_avg_sq"]) for p, r in zip(ps, ref))
    record_parity(f"fused_clip_adamw {dtype} misaligned_grads={misalign}", {"param": err, "exp_avg": m_err, "exp_avg_sq": v_err,
                                                                            "grad_norm": worst_norm})
    assert worst_norm < 1e-5
    assert m_err < 1e-5 and v_err < 1e-5
    assert err < (1e-5 if dtype == torch.float32 else 8e-3)      # bf16: one ulp where an fp32 update straddles a rounding boundary
    assert fopt.param_groups[0]["lr"] == pytest.approx(3e-3) and sched.get_last_lr()[0] == pytest.approx(3e-3)


def test_fused_clip_adamw_on_the_vae(vcd):
    """All 248 tensors of the SDXL VAE (bf16 parameters as bench.py builds them): one step equals torch's update."""
    import copy
    torch.manual_seed(0)
    vae = vcd.B200AutoencoderKL.from_pretrained("random-init:42", torch_dtype=torch.bfloat16).cuda()
    ref = copy.deepcopy(vae).float()
    gen = torch.Generator(device="cuda").manual_seed(1)
    for p, r in zip(vae.parameters(), ref.parameters()):
        p.grad = (torch.randn(p.shape, device="cuda", generator=gen) * 1e-3).to(torch.bfloat16)
        r.grad = p.grad.float()
    hp = dict(lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    topt = torch.optim.AdamW(ref.parameters(), **hp)
    fopt = vcd.FusedClipAdamW(vae.parameters(), **hp)
    n_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    n_got = fopt.clip_grad_norm_(None, 1.0)
    topt.step()
    l0 = vcd._lib.launches
    fopt.step()
    assert vcd._lib.launches - l0 == 1
    assert abs(float(n_got) - float(n_ref)) < 1e-5 * float(n_ref)
    worst = max(float((p.detach().float() - r.detach().to(torch.bfloat16).float()).abs().max() / r.detach().abs().max())
                for p, r in zip(vae.parameters(), ref.parameters()))
    assert worst < 8e-3, worst
    moved = sum(int((p.detach().float() != q.detach().float()).sum()) for p, q in
                zip(vae.parameters(), vcd.B200AutoencoderKL.from_pretrained("random-init:42", torch_dtype=torch.bfloat16).cuda().parameters()))
    assert moved > 0


def test_device_prefetcher_yields_every_batch_once_in_order(vcd):
    """input pipeline: tensors and collate_fn-style dicts, copy stream ordering, empty iterable"""
    from vcd_b200.data import DevicePrefetcher
    host = [torch.full((4, 3, 16, 16), float(i)).pin_memory() for i in range(5)]
    got = [b for b in DevicePrefetcher(host, "cuda")]
    assert len(got) == 5 and all(b.is_cuda for b in got)
    assert [float(b.mean()) for b in got] == [0.0, 1.0, 2.0, 3.0, 4.0]
    dicts = [{"pixel_values": h, "labels": torch.tensor([i])} for i, h in enumerate(host)]
    out = list(DevicePrefetcher(dicts, "cuda"))
    assert [int(d["labels"]) for d in out] == [0, 1, 2, 3, 4] and all(d["pixel_values"].is_cuda for d in out)
    assert list(DevicePrefetcher([], "cuda")) == []
    with pytest.raises(ValueError):
        DevicePrefetcher(host, "cpu")
