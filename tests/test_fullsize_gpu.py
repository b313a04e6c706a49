"""Parity at BASELINE.json's full layer sizes (512^2 fonts_nudge), where a CPU oracle would take minutes: size-
independent properties instead.
  * two independent device implementations (CTA-pair halo kernels vs the single-CTA tcgen05 kernel) agree on fprop,
    dgrad, wgrad of the dominant layer shapes;
  * the GroupNorm sums produced by a conv epilogue equal the sums of the stored tensor (checksum of checksums);
  * GroupNorm forward output has zero mean / unit variance per (image, group) before the affine (idempotence of the
    normalisation), statistics slot == direct reduction."""
import math

import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,H,cin,cout,k,stride", [
    (2, 512, 128, 128, 3, 1),     # decoder.up_blocks.3 / encoder.down_blocks.0 resnets
    (2, 256, 256, 256, 3, 1),
    (2, 128, 512, 512, 3, 1),
    (2, 512, 256, 128, 3, 1),     # decoder.up_blocks.3.resnets.0.conv1
    (2, 512, 256, 128, 1, 1),     # ... and its 1x1 shortcut
    (2, 512, 128, 128, 3, 2),     # encoder.down_blocks.0.downsamplers.0
    (2, 512, 128, 3, 3, 1),       # decoder.conv_out: narrow N with the weight operand resident in shared memory
])
def test_pair_and_single_cta_kernels_agree_at_full_size(vcd, N, H, cin, cout, k, stride):
    ops, lib = vcd.ops, vcd._lib.lib()
    torch.manual_seed(3)
    x = torch.randn(N, H, H, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * k * k)).to(torch.bfloat16)
    b = torch.randn(cout, device="cuda").to(torch.bfloat16)
    Ho = H // stride
    g = torch.randn(N, Ho, Ho, cout, device="cuda").to(torch.bfloat16)
    pad = 1 if (k == 3 and stride == 1) else 0

    def run(pair_on):
        prev = lib.vcd_set_pair_kernels(pair_on)
        try:
            n0 = lib.vcd_pair_kernel_launches()
            xp, wp, bp = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
            y = ops.conv2d(xp, wp, bp, ops.PackedWeights(), stride=stride, pad_t=pad, pad_l=pad, out_hw=(Ho, Ho))
            y.backward(g)
            used = lib.vcd_pair_kernel_launches() - n0
            return y.detach(), xp.grad, wp.grad, bp.grad, used
        finally:
            lib.vcd_set_pair_kernels(prev)

    y1, dx1, dw1, db1, used1 = run(1)
    y0, dx0, dw0, db0, used0 = run(0)
    # decoder.conv_out with VCD_SMALL_CONV=1: only the narrow-N fprop is a pair-kernel launch (its 3 -> 128 data gradient
    # then runs in conv_small.cu)
    assert used1 >= (1 if cout < 8 else 2) and used0 == 0
    # same bf16 inputs, fp32 accumulation in a different order: at most ~1 bf16 ulp (2^-8) of the largest element
    assert rel_err(y1, y0) < 8e-3
    assert rel_err(dx1, dx0) < 8e-3
    assert rel_err(dw1, dw0) < 1e-2      # bf16 parameter-dtype gradients, fp32 atomics in a different order
    assert rel_err(db1, db0) < 1e-2


def test_upconv_matches_materialised_upsample_at_full_size(vcd):
    """decoder.up_blocks.2.upsamplers.0 (256 -> 256, 256^2 -> 512^2): four phase convolutions == nearest x2 + conv."""
    ops = vcd.ops
    torch.manual_seed(4)
    x = torch.randn(2, 256, 256, 256, device="cuda").to(torch.bfloat16)
    w = (torch.randn(256, 256, 3, 3, device="cuda") / 48.0).to(torch.bfloat16)
    b = torch.randn(256, device="cuda").to(torch.bfloat16)
    g = torch.randn(2, 512, 512, 256, device="cuda").to(torch.bfloat16)
    xa, wa, ba = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    ya = ops.upconv2d(xa, wa, ba, ops.UpconvPackedWeights())
    ya.backward(g)
    xb, wb, bb = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    yb = ops.conv2d(ops.upsample2x(xb), wb, bb, ops.PackedWeights())
    yb.backward(g)
    assert rel_err(ya, yb) < 1e-2          # the phase weights are sums of 1-4 taps rounded to bf16 once more
    assert rel_err(xa.grad, xb.grad) < 1e-2
    assert rel_err(wa.grad, wb.grad) < 1.5e-2
    assert rel_err(ba.grad, bb.grad) < 1e-2


def test_epilogue_groupnorm_sums_and_normalisation_at_full_size(vcd):
    ops = vcd.ops
    torch.manual_seed(5)
    N, H, C = 2, 512, 128
    x = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, device="cuda") / 34.0).to(torch.bfloat16)
    b = torch.randn(C, device="cuda").to(torch.bfloat16)
    ops.clear_colsums()
    y = ops.conv2d(x, w, b, ops.PackedWeights(), gn_groups=32)
    fused = ops._GNSUMS[y.data_ptr()][1].clone()
    yf = y.double().reshape(N, H * H, 32, C // 32)
    ref = torch.stack([yf.sum(dim=(1, 3)), (yf * yf).sum(dim=(1, 3))], dim=-1).reshape(-1)
    assert torch.allclose(fused, ref, rtol=1e-6, atol=1e-2), float((fused - ref).abs().max())
    gamma = torch.ones(C, device="cuda", dtype=torch.bfloat16)
    beta = torch.zeros(C, device="cuda", dtype=torch.bfloat16)
    slot = ops.TrackSlot(C, "cuda", 0.0)
    out = ops.group_norm(y, gamma, beta, 32, 1e-6, False, None, slot)      # consumes the fused sums
    assert y.data_ptr() not in ops._GNSUMS
    of = out.float().reshape(N, H * H, 32, C // 32)
    assert float(of.mean(dim=(1, 3)).abs().max()) < 2e-3
    assert float((of.var(dim=(1, 3), unbiased=False) - 1).abs().max()) < 5e-3
    # tracker slot (monitor.py:64-67 mean |y| per channel) == direct reduction of the stored output
    mean_abs = slot.run[0 * C:1 * C]          # run[0] = sum over forwards of per-forward mean|x|
    direct = out.float().abs().mean(dim=(0, 1, 2))
    assert rel_err(mean_abs, direct) < 2e-3
