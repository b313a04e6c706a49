"""On-device PSNR / SSIM with the torchmetrics call surface evaluate.py uses.

Reference: src/evaluate.py:163-176 builds ``torchmetrics.image.PeakSignalNoiseRatio(data_range=1.0)`` and
``StructuralSimilarityIndexMeasure(data_range=1.0, gaussian_kernel=True, sigma=1.5, kernel_size=11)``, calls
``.update(reconstructions_0_1, originals_0_1)`` per batch (:238-249) and ``.compute().item()`` at the end (:281-288).
Here both metrics accumulate in two fp64 device scalars fed by ONE fused kernel per update (`vcd_ssim_psnr_update`:
Gaussian-window moments, SSIM map, per-image mean, and the squared error, in a single read of the two batches).
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .ops import _p, _st, call


class _PairAccumulator:
    """Shared device state: running sum of per-image SSIM, running squared error, image / element counts."""

    def __init__(self, data_range: float, kernel_size: int, sigma: float):
        self.data_range, self.kernel_size, self.sigma = float(data_range), int(kernel_size), float(sigma)
        self.acc = None
        self.images = 0
        self.elements = 0

    def update(self, preds: torch.Tensor, target: torch.Tensor):
        if not preds.is_cuda:
            raise _lib.VcdError("PSNR / SSIM run on the device only (no CPU path)")
        if preds.shape != target.shape or preds.dim() != 4:
            raise _lib.VcdError(f"expected two [N, C, H, W] batches of equal shape, got {tuple(preds.shape)} / {tuple(target.shape)}")
        p = preds.detach().to(torch.float32).contiguous()
        t = target.detach().to(device=p.device, dtype=torch.float32).contiguous()
        if self.acc is None or self.acc.device != p.device:
            self.acc = torch.zeros(2, dtype=torch.float64, device=p.device)
        N, C, H, W = p.shape
        with torch.cuda.device(p.device):
            call("vcd_ssim_psnr_update", _p(p), _p(t), N, C, H, W, self.data_range, self.kernel_size, self.sigma,
                 self.acc.data_ptr(), self.acc.data_ptr() + 8, _st())
        self.images += N
        self.elements += p.numel()


class _Metric:
    def __init__(self, data_range: float = 1.0, kernel_size: int = 11, sigma: float = 1.5):
        self._s = _PairAccumulator(data_range, kernel_size, sigma)

    def to(self, *_a, **_k):
        return self

    def update(self, preds, target):
        self._s.update(preds, target)

    def __call__(self, preds, target):
        one = type(self)(self._s.data_range, self._s.kernel_size, self._s.sigma) if isinstance(self, StructuralSimilarityIndexMeasure) \
            else type(self)(self._s.data_range)
        one.update(preds, target)
        self.update(preds, target)
        return one.compute()

    def reset(self):
        self._s.acc = None
        self._s.images = self._s.elements = 0


class PeakSignalNoiseRatio(_Metric):
    """[upstream] torchmetrics.image.PeakSignalNoiseRatio(data_range): 10 log10(R^2 / mean squared error over everything seen)."""

    def __init__(self, data_range: float = 1.0, **_):
        super().__init__(data_range)

    def compute(self) -> torch.Tensor:
        s = self._s
        if s.acc is None or s.elements == 0:
            return torch.tensor(float("nan"))
        mse = s.acc[1] / s.elements
        return (10.0 * torch.log10(torch.tensor(s.data_range ** 2, dtype=torch.float64, device=mse.device) / mse)).to(torch.float32)


class StructuralSimilarityIndexMeasure(_Metric):
    """[upstream] torchmetrics.image.StructuralSimilarityIndexMeasure(gaussian_kernel=True, sigma, kernel_size, data_range),
    reduction 'elementwise_mean': mean over all images seen of the per-image mean of the SSIM map."""

    def __init__(self, data_range: float = 1.0, gaussian_kernel: bool = True, sigma: float = 1.5, kernel_size: int = 11, **_):
        if not gaussian_kernel:
            raise _lib.VcdError("StructuralSimilarityIndexMeasure: only the Gaussian window of evaluate.py:168-172 is implemented")
        super().__init__(data_range, kernel_size, sigma)

    def compute(self) -> torch.Tensor:
        s = self._s
        if s.acc is None or s.images == 0:
            return torch.tensor(float("nan"))
        return (s.acc[0] / s.images).to(torch.float32)
