"""Loss of the training step (src/train.py:289-291) on the fused kernels.

train.py itself calls ``F.mse_loss`` on the fp32 reconstruction (that file is an unchanged caller);
this helper is what bench.py and the parity tests use: one fused kernel reads the bf16 NHWC
reconstruction and the fp32 loader tensor, produces the loss and the gradient in a single pass."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def vae_loss(model_output: dict, pixel_values: torch.Tensor, kl_weight: float):
    """returns (total, rec, kl) exactly as train.py:289-291 defines them."""
    rec = model_output["reconstruction"]
    nhwc = getattr(rec, "_vcd_nhwc", None)
    if nhwc is not None:
        rec_loss = ops.mse_loss(nhwc, pixel_values)
    else:
        rec_loss = F.mse_loss(rec.float(), pixel_values.float(), reduction="mean")
    kl = model_output["latent_dist"].kl().mean()
    return rec_loss + kl_weight * kl, rec_loss, kl
