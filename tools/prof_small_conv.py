"""Times the small-channel layers of the VAE at full size (conv_in 3->128 and conv_out 128->3 at R^2, batch B): fprop and
fprop+backward through ops.conv2d (im2col patch + tcgen05 GEMM / narrow-N implicit GEMM), CUDA events."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcd_b200  # noqa: E402

ops = vcd_b200.ops


def main(R=512, B=8, iters=10):
    for cin, cout in ((3, 128), (128, 3)):
        x = torch.randn(B, R, R, cin, device="cuda", dtype=torch.bfloat16).requires_grad_()
        w = (torch.randn(cout, cin, 3, 3, device="cuda") / math.sqrt(9 * cin)).to(torch.bfloat16).requires_grad_()
        b = torch.zeros(cout, device="cuda", dtype=torch.bfloat16).requires_grad_()
        g = torch.randn(B, R, R, cout, device="cuda", dtype=torch.bfloat16)
        packs = ops.PackedWeights()

        def fwd():
            return ops.conv2d(x, w, b, packs, stride=1, pad_t=1, pad_l=1, out_hw=(R, R))

        def both():
            fwd().backward(g)

        for name, fn in (("fprop", fwd), ("fprop+bwd", both)):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"{cin}->{cout} @{R}^2 B={B} {name}: {e0.elapsed_time(e1) / iters * 1e3:.0f} us")


if __name__ == "__main__":
    main()
