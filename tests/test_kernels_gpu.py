"""Per-kernel parity (GPU): each libvcd_b200 kernel against a plain PyTorch fp32 reference of the same op
on identical (bf16-rounded) inputs.  Tolerances: bf16 outputs 1e-2 max-rel."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import bf16_round, nchw, nhwc, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)


def _conv_case(vcd, N, H, W, Cin, Cout, k, stride, impl, residual=False):
    ops = vcd.ops
    dev = "cuda"
    pad = 1 if (k == 3 and stride == 1) else 0
    x = bf16_round(torch.randn(N, Cin, H, W, device=dev))
    w = (torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * k * k)).requires_grad_()
    b = (torch.randn(Cout, device=dev) * 0.1).requires_grad_()
    wr, br = bf16_round(w.detach()).requires_grad_(), b.detach().clone().requires_grad_()
    xr = x.clone().requires_grad_()
    if stride == 2:
        ref = F.conv2d(F.pad(xr, (0, 1, 0, 1)), wr, br, stride=2)
    else:
        ref = F.conv2d(xr, wr, br, padding=pad)
    Ho, Wo = ref.shape[2], ref.shape[3]
    res = bf16_round(torch.randn_like(ref)) if residual else None
    if residual:
        ref = ref + res
    xp = nhwc(x).requires_grad_()
    packs = ops.PackedWeights()
    y = ops.conv2d(xp, w, b, packs, stride=stride, pad_t=pad, pad_l=pad, out_hw=(Ho, Wo),
                   residual=None if res is None else nhwc(res), impl=impl)
    assert rel_err(nchw(y), ref) < TOL, "fprop"
    g = bf16_round(torch.randn_like(ref))
    ref.backward(g)
    y.backward(nhwc(g))
    assert rel_err(nchw(xp.grad), xr.grad) < TOL, "dgrad"
    assert rel_err(w.grad, wr.grad) < TOL, "wgrad"
    assert rel_err(b.grad, br.grad) < TOL, "bias grad"


@pytest.mark.parametrize("cin,cout,k", [(3, 128, 3), (128, 3, 3), (512, 8, 3), (4, 512, 3), (8, 8, 1), (4, 4, 1), (16, 24, 3)])
def test_conv_simt_small_channels(vcd, cin, cout, k):
    _conv_case(vcd, 2, 12, 10, cin, cout, k, 1, vcd._lib.IMPL_SIMT)


@pytest.mark.parametrize("cin,cout", [(3, 128), (128, 3), (512, 8), (4, 512), (128, 64)])
def test_conv_small_channels_on_tcgen05(vcd, cin, cout):
    """the small-channel layers of the VAE on the tcgen05 kernel: narrow-N implicit GEMM / im2col patch + GEMM"""
    _conv_case(vcd, 2, 12, 10, cin, cout, 3, 1, vcd._lib.IMPL_UMMA if (cin, cout) != (128, 64) else vcd._lib.IMPL_AUTO)
    _conv_case(vcd, 1, 16, 16, cin, cout, 3, 1, vcd._lib.IMPL_AUTO)


def test_conv_simt_stride2_and_residual(vcd):
    _conv_case(vcd, 2, 12, 8, 16, 16, 3, 2, vcd._lib.IMPL_SIMT)
    _conv_case(vcd, 1, 6, 6, 32, 32, 3, 1, vcd._lib.IMPL_SIMT, residual=True)


def _gn_ref(x, g, b, G, eps, act):
    y = F.group_norm(x, G, g, b, eps)
    return F.silu(y) if act else y


@pytest.mark.parametrize("C,H,W,act", [(64, 16, 16, True), (128, 32, 24, True), (512, 8, 8, False), (256, 5, 7, True)])
def test_groupnorm_silu_fwd_bwd(vcd, C, H, W, act):
    ops = vcd.ops
    N, G, eps = 3, 32, 1e-6
    x = bf16_round(torch.randn(N, C, H, W, device="cuda") * 2 + 0.3)
    gamma = (torch.rand(C, device="cuda") + 0.5).requires_grad_()
    beta = (torch.randn(C, device="cuda") * 0.1).requires_grad_()
    xr, gr, br = x.clone().requires_grad_(), gamma.detach().clone().requires_grad_(), beta.detach().clone().requires_grad_()
    ref = _gn_ref(xr, gr, br, G, eps, act)
    xp = nhwc(x).requires_grad_()
    y = ops.group_norm(xp, gamma, beta, G, eps, act)
    assert rel_err(nchw(y), ref) < TOL
    go = bf16_round(torch.randn_like(ref))
    ref.backward(go)
    y.backward(nhwc(go))
    assert rel_err(nchw(xp.grad), xr.grad) < TOL
    assert rel_err(gamma.grad, gr.grad) < TOL
    assert rel_err(beta.grad, br.grad) < TOL


def test_groupnorm_split_skip_gradient_and_colsum(vcd):
    """split=True: the skip-connection gradient is added inside the dx kernel and the column sums of dx
    (bias gradient of the producing conv) come out of the same pass."""
    ops = vcd.ops
    N, C, H, W, G = 2, 128, 12, 12, 32
    x = bf16_round(torch.randn(N, C, H, W, device="cuda"))
    gamma = (torch.rand(C, device="cuda") + 0.5).requires_grad_()
    beta = (torch.randn(C, device="cuda") * 0.1).requires_grad_()
    xr = x.clone().requires_grad_()
    ref = _gn_ref(xr, gamma.detach(), beta.detach(), G, 1e-6, True)
    go, gs = bf16_round(torch.randn_like(ref)), bf16_round(torch.randn_like(ref))
    (ref * go).sum().add((xr * gs).sum()).backward()
    xp = nhwc(x).requires_grad_()
    seen = []
    xp.register_hook(lambda g: seen.append((g, ops.pop_colsum(g))))   # what a producing conv's backward receives
    y, xid = ops.group_norm(xp, gamma, beta, G, 1e-6, True, None, None, True)
    assert torch.equal(xid, xp)
    torch.autograd.backward([y, xid], [nhwc(go), nhwc(gs)])
    assert rel_err(nchw(xp.grad), xr.grad) < TOL
    g, cs = seen[0]
    assert cs is not None and rel_err(cs, xr.grad.sum(dim=[0, 2, 3])) < TOL
    assert ops.pop_colsum(g) is None


def test_groupnorm_bf16_params(vcd):
    ops = vcd.ops
    N, C, H, W, G = 2, 128, 8, 8, 32
    x = bf16_round(torch.randn(N, C, H, W, device="cuda"))
    gamma = (torch.rand(C, device="cuda") + 0.5).to(torch.bfloat16).requires_grad_()
    beta = (torch.randn(C, device="cuda") * 0.1).to(torch.bfloat16).requires_grad_()
    ref = _gn_ref(x, gamma.float(), beta.float(), G, 1e-6, True)
    y = ops.group_norm(nhwc(x), gamma, beta, G, 1e-6, True)
    assert rel_err(nchw(y), ref) < TOL
    y.float().sum().backward()
    assert gamma.grad.dtype == torch.bfloat16 and torch.isfinite(gamma.grad.float()).all()


def test_fused_channel_stats_match_monitor_formula(vcd):
    """monitor.py:64-75 on the GroupNorm output (pre-SiLU) and input, through the fused slots."""
    ops = vcd.ops
    N, C, H, W, G = 2, 128, 16, 16, 32
    x = bf16_round(torch.randn(N, C, H, W, device="cuda") + 0.2)
    gamma = torch.rand(C, device="cuda") + 0.1
    beta = torch.randn(C, device="cuda") * 0.1
    s_in, s_out = ops.TrackSlot(C, "cuda", 0.05), ops.TrackSlot(C, "cuda", 0.05)
    for _ in range(2):
        ops.group_norm(nhwc(x), gamma, beta, G, 1e-6, True, s_in, s_out)
    y = F.group_norm(x, G, gamma, beta, 1e-6)
    for slot, t in ((s_in, x), (s_out, y)):
        run = slot.run.view(5, C)
        F_ = float(slot.scal[2])
        assert F_ == 2.0
        ref_abs = t.abs().mean(dim=[0, 2, 3])
        assert rel_err(run[0] / F_, ref_abs) < 1e-4
        assert abs(float(slot.scal[0]) / F_ - float(t.mean())) < 1e-4 * max(1.0, abs(float(t.mean())))
        assert abs(float(slot.scal[1]) / F_ - float(t.std())) < 1e-4 * float(t.std())
        assert rel_err(run[3], t.abs().amax(dim=[0, 2, 3])) < 1e-6
        assert rel_err(run[4] / F_, (t.abs() < 0.05).float().mean(dim=[0, 2, 3])) < 1e-4 + 1e-9


@pytest.mark.parametrize("act,tau,C,H,W", [(True, 0.05, 128, 16, 16), (False, 0.0, 256, 9, 7), (True, 0.0, 512, 8, 8),
                                             (True, 0.05, 128, 64, 64)])
def test_input_and_output_statistics_in_one_apply_pass(vcd, act, tau, C, H, W):
    """The group sums come from the producing GEMM's epilogue (here: pushed by hand), so the statistics pass does not run:
    vcd_gn_apply_fwd then fills BOTH slots in its single pass over x — the input slot directly, the output slot's sum /
    sum of squares derived from the input sums (y is affine in x per image and channel), its |y| sum, max and near-zero
    count accumulated.  All five rows of both slots against torch."""
    ops = vcd.ops
    N, G = 3, 32
    x = bf16_round(torch.randn(N, C, H, W, device="cuda") * 1.5 + 0.2)
    gamma = torch.rand(C, device="cuda") + 0.1
    beta = torch.randn(C, device="cuda") * 0.1
    xp = nhwc(x)
    xg = x.reshape(N, G, -1).double()
    sums = torch.stack([xg.sum(-1), (xg * xg).sum(-1)], dim=-1).reshape(-1).contiguous()       # [N][G][2] fp64
    s_in, s_out, s_extra = ops.TrackSlot(C, "cuda", tau), ops.TrackSlot(C, "cuda", tau), ops.TrackSlot(C, "cuda", tau)
    l0 = vcd._lib.launches
    ops.push_gn_sums(xp, sums, G)
    out = ops.group_norm(xp, gamma, beta, G, 1e-6, act, s_in, s_out, slot_in_extra=s_extra)
    assert vcd._lib.launches - l0 == 1 + 3          # one apply launch + three finalize launches: no vcd_gn_stats
    y = F.group_norm(x, G, gamma, beta, 1e-6)
    assert rel_err(nchw(out), F.silu(y) if act else y) < TOL
    n = N * H * W
    for slot, t in ((s_in, x), (s_extra, x), (s_out, y)):
        run = slot.run.view(5, C)
        assert float(slot.scal[2]) == 1.0
        assert rel_err(run[0], t.abs().mean(dim=[0, 2, 3])) < 1e-4                       # mean |.|   (monitor.py:64-67)
        assert rel_err(run[1], t.mean(dim=[0, 2, 3])) < 1e-4                             # mean
        assert rel_err(run[2], t.var(dim=[0, 2, 3], unbiased=False)) < 2e-4              # variance
        assert rel_err(run[3], t.abs().amax(dim=[0, 2, 3])) < 1e-6                       # max |.|
        if tau > 0:
            assert rel_err(run[4], (t.abs() < tau).float().mean(dim=[0, 2, 3])) < 1e-4 + 2.0 / n
        assert abs(float(slot.scal[0]) - float(t.mean())) < 1e-4 * max(1.0, abs(float(t.mean())))
        assert abs(float(slot.scal[1]) - float(t.std())) < 1e-4 * float(t.std())


def test_chan_stats_standalone_layouts(vcd):
    ops = vcd.ops
    t = torch.randn(3, 24, 7, 9, device="cuda")
    ref = t.abs().mean(dim=[0, 2, 3])
    for x in (t, t.contiguous(memory_format=torch.channels_last), t.to(torch.bfloat16)):
        slot = ops.TrackSlot(24, "cuda")
        ops.chan_stats(x, slot)
        tol = 1e-4 if x.dtype == torch.float32 else 1e-2
        assert rel_err(slot.run.view(5, 24)[0], ref) < tol


def test_upsample_planes_layout_roundtrip(vcd):
    ops, call = vcd.ops, vcd._lib.call
    x = bf16_round(torch.randn(2, 64, 6, 8, device="cuda"))
    xp = nhwc(x).requires_grad_()
    y = ops.upsample2x(xp)
    ref = F.interpolate(x, scale_factor=2.0, mode="nearest")
    assert torch.equal(nchw(y), ref)
    g = bf16_round(torch.randn_like(ref))
    y.backward(nhwc(g))
    refg = F.avg_pool2d(g, 2) * 4
    assert rel_err(nchw(xp.grad), refg) < TOL
    st = torch.cuda.current_stream().cuda_stream
    a = nhwc(x)
    planes = torch.empty(2, 4, 3, 4, 64, dtype=torch.bfloat16, device="cuda")
    back = torch.empty_like(a)
    call("vcd_space_to_planes", a.data_ptr(), planes.data_ptr(), 2, 6, 8, 64, st)
    assert torch.equal(planes[:, 1 * 2 + 0], a[:, 1::2, 0::2])
    call("vcd_planes_to_space", planes.data_ptr(), back.data_ptr(), 2, 6, 8, 64, st)
    assert torch.equal(back, a)


def test_layout_conversions(vcd):
    ops = vcd.ops
    x = torch.rand(2, 3, 8, 6, device="cuda") * 2 - 1
    y = ops.to_nhwc(x)
    assert y.shape == (2, 8, 6, 3) and torch.equal(y, nhwc(x))
    z = ops.to_nchw(y, torch.float32)
    assert z.is_contiguous() and torch.equal(z, bf16_round(x))


def test_softmax_and_transpose(vcd):
    call = vcd._lib.call
    st = torch.cuda.current_stream().cuda_stream
    s = bf16_round(torch.randn(5, 64, 64, device="cuda") * 3).to(torch.bfloat16)
    p = torch.empty_like(s)
    call("vcd_softmax_fwd", s.data_ptr(), p.data_ptr(), 5 * 64, 64, st)
    ref = torch.softmax(s.float(), -1)
    assert rel_err(p, ref) < TOL
    dp = torch.randn_like(s)
    ds = torch.empty_like(s)
    call("vcd_softmax_bwd", p.data_ptr(), dp.data_ptr(), ds.data_ptr(), 0.5, 5 * 64, 64, st)
    pr = p.float()
    refd = 0.5 * pr * (dp.float() - (dp.float() * pr).sum(-1, keepdim=True))
    assert rel_err(ds, refd) < TOL
    x = torch.randn(3, 40, 72, device="cuda").to(torch.bfloat16)
    y = torch.empty(3, 72, 40, dtype=torch.bfloat16, device="cuda")
    call("vcd_transpose_bf16", x.data_ptr(), y.data_ptr(), 3, 40, 72, st)
    assert torch.equal(y, x.transpose(1, 2))


def test_gauss_sample_kl_and_mse(vcd):
    ops = vcd.ops
    N, h, w = 3, 8, 8
    mom = bf16_round(torch.randn(N, 8, h, w, device="cuda") * 2)
    mom[0, 4, 0, 0] = 25.0   # beyond the logvar clamp
    mom[0, 5, 0, 0] = -40.0
    noise = torch.randn(N, 4, h, w, device="cuda")
    mr = mom.clone().requires_grad_()
    mean, logvar = torch.chunk(mr, 2, dim=1)
    lv = torch.clamp(logvar, -30.0, 20.0)
    z_ref = mean + torch.exp(0.5 * lv) * noise
    kl_ref = 0.5 * torch.sum(mean ** 2 + torch.exp(lv) - 1.0 - lv, dim=[1, 2, 3])
    mp = nhwc(mom).requires_grad_()
    z, kl, m_out, lv_out = ops.gauss_sample_kl(mp, noise)
    assert rel_err(nchw(z), z_ref) < TOL
    assert rel_err(kl, kl_ref) < 1e-4
    assert torch.equal(m_out, mean.detach()) and rel_err(lv_out, lv) < 1e-6
    gz = bf16_round(torch.randn_like(z_ref))
    (z_ref * gz).sum().add(kl_ref.mean() * 0.3).backward()
    (z.float() * nhwc(gz).float()).sum().add(kl.mean() * 0.3).backward()
    assert rel_err(nchw(mp.grad), mr.grad) < TOL
    # mode(): no noise
    z2, kl2, _, _ = ops.gauss_sample_kl(nhwc(mom), None)
    assert rel_err(nchw(z2), mean.detach()) < 1e-6 and rel_err(kl2, kl_ref) < 1e-4
    # mse
    x = torch.rand(2, 3, 16, 12, device="cuda") * 2 - 1
    rec = bf16_round(torch.randn(2, 3, 16, 12, device="cuda"))
    rr = rec.clone().requires_grad_()
    ref = F.mse_loss(rr, x)
    rp = nhwc(rec).requires_grad_()
    loss = ops.mse_loss(rp, x)
    assert abs(float(loss) - float(ref)) < 1e-5 * float(ref)
    (ref * 2).backward()
    (loss * 2).backward()
    assert rel_err(nchw(rp.grad), rr.grad) < TOL
