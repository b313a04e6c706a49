"""Times vcd_multi_pack_weights (all GEMM operand packs of the encoder / decoder in one launch each) with CUDA events.
Algorithmic bytes: parameter bytes read + 2 bf16 packs written (Upsample2D convs: 16/9 of that)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcd_b200  # noqa: E402


def main():
    for dtype in (torch.bfloat16, torch.float32):
        vae = vcd_b200.B200AutoencoderKL.from_pretrained("random-init:5", torch_dtype=dtype).cuda()
        tot_ms, tot_b = 0.0, 0
        for which in ("encoder", "decoder"):
            plan = vae._pack_weights(which)
            nbytes = 0
            for w, b, packs, mode in plan.layers:
                nbytes += w.numel() * w.element_size() + packs.wf.numel() * 2 + (packs.wd.numel() * 2 if packs.wd is not None else 0)
            for _ in range(3):
                plan.run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                plan.run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            tot_ms += ms
            tot_b += nbytes
            print(f"{dtype} {which}: {plan.n_tiles} tiles, {ms * 1e3:.1f} us, {nbytes / 1e6:.1f} MB, {nbytes / ms / 1e6:.0f} GB/s")
        print(f"{dtype} both: {tot_ms * 1e3:.1f} us per forward, {tot_b / tot_ms / 1e6:.0f} GB/s (weights stay in the 126 MB L2 between "
              f"iterations only partly: {tot_b / 1e6:.0f} MB)")


if __name__ == "__main__":
    main()
