"""tcgen05 implicit-GEMM kernel (umma_gemm.cu) parity: against torch fp32 conv/matmul on bf16-rounded
inputs AND against the independent CUDA-core kernels.  Shapes cover every layer class of the SDXL VAE
(SURVEY 8a5.1): 128/256/512 channels, 3x3 s1, 3x3 s2 with (0,1,0,1) pad, 1x1 shortcuts, Linear,
ragged tiles (W not a multiple of the tile, batch not a multiple of the tile)."""
import math

import pytest
import torch
import torch.nn.functional as F

from test_kernels_gpu import _conv_case
from util import bf16_round, nchw, nhwc, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(1)


@pytest.mark.parametrize("N,H,W,cin,cout,k,stride", [
    (2, 16, 16, 128, 128, 3, 1),
    (1, 8, 8, 128, 128, 3, 1),      # 64 pixels: tile spans a missing second image
    (3, 8, 8, 512, 512, 3, 1),      # odd batch, BLOCK_N 256, 2 n-tiles
    (2, 32, 32, 128, 256, 3, 1),
    (1, 16, 16, 256, 128, 1, 1),    # 1x1 shortcut
    (2, 12, 20, 128, 128, 3, 1),    # ragged: W, H not powers of two
    (1, 6, 136, 128, 128, 3, 1),    # W > 128 with a partial tile
    (2, 16, 16, 128, 128, 3, 2),    # Downsample2D
    (1, 32, 32, 256, 256, 3, 2),
])
def test_conv_umma(vcd, N, H, W, cin, cout, k, stride):
    _conv_case(vcd, N, H, W, cin, cout, k, stride, vcd._lib.IMPL_UMMA)


def test_conv_umma_residual(vcd):
    _conv_case(vcd, 2, 16, 16, 256, 256, 3, 1, vcd._lib.IMPL_UMMA, residual=True)


def test_conv_umma_matches_simt_bitwise_close(vcd):
    """same bf16 inputs through both device paths: only fp32 summation order differs."""
    ops = vcd.ops
    x = torch.randn(2, 16, 16, 128, device="cuda").to(torch.bfloat16)
    w = torch.randn(128, 128, 3, 3, device="cuda") / 34.0
    b = torch.randn(128, device="cuda")
    ya = ops.conv2d(x, w, b, ops.PackedWeights(), impl=vcd._lib.IMPL_UMMA)
    yb = ops.conv2d(x, w, b, ops.PackedWeights(), impl=vcd._lib.IMPL_SIMT)
    assert rel_err(ya, yb) < 4e-3


@pytest.mark.parametrize("batch,M,N,K,bb", [(1, 256, 512, 512, False), (3, 64, 64, 512, True), (2, 200, 128, 64, True),
                                            (2, 4096, 512, 512, False)])
def test_gemm_nt(vcd, batch, M, N, K, bb):
    ops = vcd.ops
    A = torch.randn(batch, M, K, device="cuda").to(torch.bfloat16)
    B = (torch.randn(batch if bb else 1, N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(batch, M, N, device="cuda").to(torch.bfloat16)
    D = torch.empty(batch, M, N, dtype=torch.bfloat16, device="cuda")
    ops._gemm_nt(A, B, D, batch, M, N, K, bb, alpha=0.5, bias=bias, residual=res)
    ref = 0.5 * torch.matmul(A.float(), B.float().transpose(1, 2)) + bias + res.float()
    assert rel_err(D, ref) < TOL


@pytest.mark.parametrize("batch,M,N,K,red", [(2, 128, 512, 256, False), (3, 64, 512, 64, False), (2, 512, 512, 1000, True),
                                             (1, 256, 128, 64, True)])
def test_gemm_tn(vcd, batch, M, N, K, red):
    ops = vcd.ops
    A = torch.randn(batch, K, M, device="cuda").to(torch.bfloat16)
    B = (torch.randn(batch, K, N, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    nb = 1 if red else batch
    D = torch.empty(nb, M, N, dtype=torch.float32, device="cuda")
    ops._gemm_tn(A, B, D, batch, M, N, K, red)
    ref = torch.matmul(A.float().transpose(1, 2), B.float())
    if red:
        ref = ref.sum(0, keepdim=True)
    assert rel_err(D, ref) < TOL


@pytest.mark.parametrize("T", [64, 256, 1024, 99, 9, 1287])
def test_attention_core(vcd, T):
    """T = 99 (11 x 9), 9 (3 x 3), 1287 (33 x 39): token counts that are no multiple of the 64-wide K step are zero-padded,
    the padded keys masked with -inf scores"""
    ops = vcd.ops
    N, C = 2, 512
    q, k, v = [bf16_round(torch.randn(N, T, C, device="cuda")) for _ in range(3)]
    qr, kr, vr = [t.clone().requires_grad_() for t in (q, k, v)]
    ref = torch.softmax(qr @ kr.transpose(1, 2) / math.sqrt(C), -1) @ vr
    qp, kp, vp = [t.to(torch.bfloat16).requires_grad_() for t in (q, k, v)]
    o = ops.attention_core(qp, kp, vp)
    assert rel_err(o, ref) < 2e-2
    g = bf16_round(torch.randn_like(ref))
    ref.backward(g)
    o.backward(g.to(torch.bfloat16))
    for a, b in ((qp, qr), (kp, kr), (vp, vr)):
        assert rel_err(a.grad, b.grad) < 3e-2


@pytest.mark.parametrize("N,H,W,cin,cout", [(2, 8, 8, 128, 128), (1, 16, 16, 256, 256), (3, 8, 8, 512, 512),
                                            (2, 12, 20, 128, 256), (1, 6, 72, 256, 128)])
def test_upconv_fused_matches_torch(vcd, N, H, W, cin, cout):
    """Upsample2D (nearest x2 + conv3x3 pad 1) as four phase convolutions on the low-resolution tensor
    (vcd_upconv2d_*) against F.interpolate + F.conv2d in fp32 on the same bf16-rounded inputs: output,
    input gradient, weight and bias gradients."""
    ops = vcd.ops
    x = bf16_round(torch.randn(N, cin, H, W, device="cuda"))
    w = bf16_round(torch.randn(cout, cin, 3, 3, device="cuda") / math.sqrt(9 * cin))
    b = torch.randn(cout, device="cuda")
    g = bf16_round(torch.randn(N, cout, 2 * H, 2 * W, device="cuda"))
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    ref = F.conv2d(F.interpolate(xr, scale_factor=2.0, mode="nearest"), wr, br, padding=1)
    ref.backward(g)
    xp = nhwc(x).requires_grad_()
    wp, bp = w.clone().requires_grad_(), b.clone().requires_grad_()
    y = ops.upconv2d(xp, wp, bp, ops.UpconvPackedWeights())
    assert y.shape == (N, 2 * H, 2 * W, cout)
    assert rel_err(nchw(y), ref) < TOL
    y.backward(nhwc(g))
    assert rel_err(nchw(xp.grad), xr.grad) < TOL
    assert rel_err(wp.grad, wr.grad) < TOL
    assert rel_err(bp.grad, br.grad) < TOL


@pytest.mark.parametrize("N,H,W,cin,cout,k,stride", [
    (2, 16, 16, 128, 128, 3, 1),    # 2 tiles per image, 4 tiles = 2 pairs
    (3, 16, 8, 512, 512, 3, 1),     # 3 tiles: the last pair has an empty second CTA; 2 n-tiles of 256
    (1, 40, 20, 128, 256, 3, 1),    # ragged: H, W not multiples of the 16 x 8 tile
    (2, 32, 32, 256, 128, 3, 1),
    (1, 64, 64, 128, 128, 3, 1),    # 32 tiles, several items per cluster: pipeline phases wrap
    (2, 32, 16, 256, 256, 3, 2),    # Downsample2D on parity planes: four halo groups
    (2, 16, 16, 256, 512, 1, 1),    # 1x1 shortcut as a ROWS-mode GEMM
    (1, 24, 24, 512, 256, 1, 1),
])
def test_conv_pair_kernel(vcd, N, H, W, cin, cout, k, stride):
    """shapes that must be served by the CTA-pair halo kernel (umma_pair.cu): parity as for test_conv_umma, plus
    the launch counter proves which kernel ran."""
    lib = vcd._lib.lib()
    n0 = lib.vcd_pair_kernel_launches()
    _conv_case(vcd, N, H, W, cin, cout, k, stride, vcd._lib.IMPL_UMMA)
    assert lib.vcd_pair_kernel_launches() > n0


def test_conv_pair_residual_and_narrow(vcd):
    lib = vcd._lib.lib()
    n0 = lib.vcd_pair_kernel_launches()
    _conv_case(vcd, 2, 32, 32, 256, 256, 3, 1, vcd._lib.IMPL_UMMA, residual=True)
    assert lib.vcd_pair_kernel_launches() > n0
    # decoder.conv_out-like: 128 -> 3 channels (narrow N, masked columns) and its dgrad
    ops = vcd.ops
    x = bf16_round(torch.randn(2, 128, 32, 32, device="cuda"))
    w = bf16_round(torch.randn(3, 128, 3, 3, device="cuda") / 34.0)
    b = torch.randn(3, device="cuda")
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    ref = F.conv2d(xr, wr, b, padding=1)
    g = bf16_round(torch.randn_like(ref))
    ref.backward(g)
    xp = nhwc(x).requires_grad_()
    wp = w.clone().requires_grad_()
    n1 = lib.vcd_pair_kernel_launches()
    y = ops.conv2d(xp, wp, b, ops.PackedWeights())
    assert lib.vcd_pair_kernel_launches() > n1
    assert rel_err(nchw(y), ref) < TOL
    y.backward(nhwc(g))
    assert rel_err(nchw(xp.grad), xr.grad) < TOL
    assert rel_err(wp.grad, wr.grad) < TOL


def test_upconv_pair_kernel(vcd):
    lib = vcd._lib.lib()
    n0 = lib.vcd_pair_kernel_launches()
    test_upconv_fused_matches_torch(vcd, 2, 32, 16, 256, 256)
    test_upconv_fused_matches_torch(vcd, 1, 16, 24, 512, 512)
    assert lib.vcd_pair_kernel_launches() >= n0 + 10   # 4 phase fprops + 1 dgrad per call


@pytest.mark.parametrize("N,H,W,cin,cout,k,fused", [
    (2, 32, 16, 128, 128, 3, True),     # D = 4 channels per group
    (3, 16, 8, 256, 256, 3, True),      # D = 8, odd tile count
    (1, 40, 24, 128, 512, 3, True),     # D = 16, ragged tiles, two n-tiles
    (2, 16, 16, 256, 128, 1, True),     # 1x1 shortcut (ROWS mode, 256 rows per image)
    (2, 12, 10, 128, 128, 3, False),    # not eligible for the pair kernel: sums from vcd_gn_stats inside the call
])
def test_conv_epilogue_groupnorm_sums(vcd, N, H, W, cin, cout, k, fused):
    """gn_sums of vcd_conv2d_fprop: sum / sum of squares of the bf16 output per (image, group), whichever path
    serves the layer, equal the sums of the stored tensor (the GroupNorm that follows never re-reads it)."""
    ops, lib = vcd.ops, vcd._lib.lib()
    x = torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k)).to(torch.bfloat16)
    b = torch.randn(cout, device="cuda").to(torch.bfloat16)
    res = torch.randn(N, H, W, cout, device="cuda").to(torch.bfloat16)
    n0 = lib.vcd_pair_kernel_launches()
    y = ops.conv2d(x, w, b, ops.PackedWeights(), pad_t=k // 2, pad_l=k // 2, residual=res, gn_groups=32)
    assert (lib.vcd_pair_kernel_launches() > n0) == fused
    sums = ops.pop_gn_sums(y, 32)
    assert sums is not None
    yf = y.double().reshape(N, H * W, 32, cout // 32)
    ref = torch.stack([yf.sum(dim=(1, 3)), (yf * yf).sum(dim=(1, 3))], dim=-1).reshape(-1)
    assert torch.allclose(sums, ref, rtol=1e-5, atol=1e-3), float((sums - ref).abs().max())


def test_upconv_epilogue_groupnorm_sums(vcd):
    ops = vcd.ops
    x = torch.randn(2, 16, 16, 256, device="cuda").to(torch.bfloat16)
    w = (torch.randn(256, 256, 3, 3, device="cuda") / 48.0).to(torch.bfloat16)
    b = torch.randn(256, device="cuda").to(torch.bfloat16)
    y = ops.upconv2d(x, w, b, ops.UpconvPackedWeights(), 32)
    sums = ops.pop_gn_sums(y, 32)
    yf = y.double().reshape(2, 32 * 32, 32, 8)
    ref = torch.stack([yf.sum(dim=(1, 3)), (yf * yf).sum(dim=(1, 3))], dim=-1).reshape(-1)
    assert torch.allclose(sums, ref, rtol=1e-5, atol=1e-3), float((sums - ref).abs().max())


@pytest.mark.parametrize("N,H,W,C,cout,split", [(2, 32, 16, 256, 128, False), (3, 16, 8, 256, 512, True),
                                                (1, 40, 24, 512, 256, True), (2, 16, 16, 128, 256, False)])
def test_groupnorm_backward_fused_into_conv_dgrad(vcd, N, H, W, C, cout, split):
    """GroupNorm+SiLU -> conv3x3: with sole_consumer_is_conv the conv's dgrad epilogue applies SiLU' and reduces the
    GroupNorm backward sums (vcd_conv2d_dgrad_gn); gradients must match the torch fp32 chain and the unfused path."""
    ops = vcd.ops
    x = bf16_round(torch.randn(N, C, H, W, device="cuda") * 1.5 + 0.3)
    gamma = bf16_round(torch.rand(C, device="cuda") + 0.5)
    beta = bf16_round(torch.randn(C, device="cuda") * 0.2)
    w = bf16_round(torch.randn(cout, C, 3, 3, device="cuda") / math.sqrt(9 * C))
    b = torch.randn(cout, device="cuda")
    g = bf16_round(torch.randn(N, cout, H, W, device="cuda"))
    gskip = bf16_round(torch.randn(N, C, H, W, device="cuda"))
    xr, gr, br, wr = (t.clone().requires_grad_() for t in (x, gamma, beta, w))
    h = F.silu(F.group_norm(xr, 32, gr, br, 1e-6))
    ref = F.conv2d(h, wr, b, padding=1)
    (ref * g).sum().backward() if not split else ((ref * g).sum() + (xr * gskip).sum()).backward()

    def run(fuse):
        ops.clear_colsums()
        xp = nhwc(x).requires_grad_()
        gp, bp, wp = gamma.clone().requires_grad_(), beta.clone().requires_grad_(), w.clone().requires_grad_()
        out = ops.group_norm(xp, gp, bp, 32, 1e-6, True, None, None, split, fuse)
        hh, xid = out if split else (out, None)
        y = ops.conv2d(hh, wp, b, ops.PackedWeights())
        loss = (y.float() * nhwc(g).float()).sum()
        if split:
            loss = loss + (xid.float() * nhwc(gskip).float()).sum()
        loss.backward()
        return y, xp.grad, gp.grad, bp.grad, wp.grad

    lib = vcd._lib.lib()
    y1, dx1, dg1, db1, dw1 = run(True)
    assert not ops._GN_BWD and not ops._GN_FWD          # both hand-offs were consumed (C = 128: fusion declined)
    y0, dx0, dg0, db0, dw0 = run(False)
    assert rel_err(nchw(y1), ref) < TOL
    for a, r, name in ((nchw(dx1), xr.grad, "dx"), (dg1, gr.grad, "dgamma"), (db1, br.grad, "dbeta"), (dw1, wr.grad, "dw")):
        assert rel_err(a, r) < 2e-2, name
    for a, r, name in ((dx1, dx0, "dx"), (dg1, dg0, "dgamma"), (db1, db0, "dbeta")):
        assert rel_err(a, r) < 1e-2, name + " fused vs unfused"


@pytest.mark.parametrize("N,H,W,cin,cout,k,stride", [
    (8, 64, 64, 512, 512, 3, 1),     # 256 dgrad items on 74 clusters: a ragged last wave for the weight gradient to fill
    (4, 128, 128, 128, 128, 3, 1),   # tap-pair wgrad on the single-CTA kernel behind a B-resident dgrad
    (2, 64, 64, 256, 128, 1, 1),     # 1x1 shortcut
    (2, 64, 64, 256, 256, 3, 2),     # Downsample2D
    (2, 12, 10, 128, 3, 3, 1),       # conv_out class: the flag only skips the zeroing (no tcgen05 wgrad)
])
def test_wgrad_overlap_modes_agree(vcd, monkeypatch, N, H, W, cin, cout, k, stride):
    """VCD_WGRAD_OVERLAP = off | stream | pdl (programmatic dependent launch behind the dgrad kernel, the default) give
    the same gradients: dx bit-identical (no atomics on that path), dw / db equal up to the order of the fp32 split-K
    atomics; repeated so that a race between the overlapped kernels would show."""
    ops = vcd.ops
    pad = 1 if (k == 3 and stride == 1) else 0
    x = nhwc(bf16_round(torch.randn(N, cin, H, W, device="cuda")))
    w0 = torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k)
    b0 = torch.randn(cout, device="cuda") * 0.1
    Ho, Wo = (H // 2, W // 2) if stride == 2 else (H, W)
    g = nhwc(bf16_round(torch.randn(N, cout, Ho, Wo, device="cuda")))

    def run(mode):
        monkeypatch.setenv("VCD_WGRAD_OVERLAP", mode)
        xp = x.clone().requires_grad_()
        w, b = w0.clone().requires_grad_(), b0.clone().requires_grad_()
        y = ops.conv2d(xp, w, b, ops.PackedWeights(), stride=stride, pad_t=pad, pad_l=pad, out_hw=(Ho, Wo))
        y.backward(g)
        torch.cuda.synchronize()
        return xp.grad, w.grad, b.grad

    dx0, dw0, db0 = run("off")
    assert float(dw0.abs().max()) > 0
    for mode in ("pdl", "stream", "pdl", "pdl"):
        for _ in range(3):
            dx, dw, db = run(mode)
            assert torch.equal(dx, dx0), mode
            assert rel_err(dw, dw0) < 1e-4, (mode, rel_err(dw, dw0))
            assert rel_err(db, db0) < 1e-4, (mode, rel_err(db, db0))
