"""TEST SHIM: torchmetrics is absent from this image; evaluate.py:15,163-176 reaches PSNR / SSIM through
``torchmetrics.image`` — served here by the product's device kernels (vae-channel-dynamics_b200/metrics.py)."""
from . import image  # noqa: F401

__version__ = "0.0-vcd-test-shim"
