"""Offline input pipeline helpers (reference src/data_utils.py:13-30,66-72,163-225).

The reference loads its images with ``datasets.load_dataset(dataset_name, name=config, split=split)`` from the HF hub
(data_utils.py:66-72).  ``load_dataset`` resolves a *relative directory of the same name* before it looks at the hub,
so a local parquet tree  ``<cwd>/<dataset_name>/data/<split>-00000-of-00001.parquet``  satisfies the unchanged call
with no network.  ``write_synthetic_image_dataset`` writes such a tree (PIL-encoded images under the configured
``image_column`` — 'img' for uoft-cs/cifar10, 'image' elsewhere — plus a ``label`` column), with the synthetic image
families bench.py uses: uniform noise, or 'glyph-like' +/-1 blocks.

``preprocess_uint8_batch`` is the device-side replacement of the reference's per-image PIL transform chain
(Resize(bilinear) -> CenterCrop -> ToTensor -> Normalize(0.5, 0.5), data_utils.py:13-30) for batches that are already
decoded to uint8: one kernel (vcd_preprocess_u8) reads [N, H, W, 3] uint8 and writes the fp32 NCHW tensor in [-1, 1]
the training loop consumes.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np


def synthetic_images(n: int, size: int, kind: str = "noise", seed: int = 0) -> np.ndarray:
    """[n, size, size, 3] uint8.  kind: 'noise' (uniform), 'glyph' (8x8-pixel blocks, 10 % ink), 'smooth' (low-pass noise)."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, size=(n, size, size, 3), dtype=np.uint8)
    if kind == "glyph":
        cell = max(1, size // 8) if size < 64 else 8
        g = max(1, size // cell)
        ink = rng.random((n, g, g, 3)) < 0.1
        img = np.repeat(np.repeat(ink, cell, axis=1), cell, axis=2)[:, :size, :size]
        return np.where(img, 255, 0).astype(np.uint8)
    if kind == "smooth":
        g = max(2, size // 8)
        low = rng.random((n, g, g, 3)).astype(np.float32)
        rep = -(-size // g)
        img = np.repeat(np.repeat(low, rep, axis=1), rep, axis=2)[:, :size, :size]
        return np.clip(img * 255.0, 0, 255).astype(np.uint8)
    raise ValueError(f"unknown synthetic image kind '{kind}'")


def write_synthetic_image_dataset(root: str, dataset_name: str, splits: Dict[str, int], size: int = 32,
                                  image_column: str = "image", kind: str = "noise", seed: int = 0,
                                  num_classes: int = 10) -> str:
    """Write ``<root>/<dataset_name>/data/<split>-00000-of-00001.parquet`` for every split; returns the dataset dir.
    Run the unchanged train.py / evaluate.py with ``cwd=root`` and ``data.dataset_name: <dataset_name>``."""
    from datasets import Dataset, Features, ClassLabel, Image
    from PIL import Image as PILImage
    ddir = os.path.join(root, *dataset_name.split("/"))
    os.makedirs(os.path.join(ddir, "data"), exist_ok=True)
    feats = Features({image_column: Image(), "label": ClassLabel(num_classes=num_classes)})
    for k, (split, n) in enumerate(sorted(splits.items())):
        arr = synthetic_images(n, size, kind, seed + 1000 * k)
        rng = np.random.default_rng(seed + 1000 * k + 1)
        ds = Dataset.from_dict({image_column: [PILImage.fromarray(a) for a in arr],
                                "label": rng.integers(0, num_classes, size=n).tolist()}, features=feats)
        ds.to_parquet(os.path.join(ddir, "data", f"{split}-00000-of-00001.parquet"))
    return ddir


def preprocess_uint8_batch(images_u8, resolution: int, out_dtype=None):
    """[N, H, W, 3] uint8 CUDA tensor -> [N, 3, R, R] fp32 in [-1, 1]: bilinear resize of the shorter side to R
    (align_corners=False, antialias off — torchvision's tensor path), centre crop, x/127.5 - 1.  One kernel."""
    import torch
    from . import _lib
    from .ops import _p, _st, call
    if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
        raise _lib.VcdError("preprocess_uint8_batch expects a CUDA uint8 tensor [N, H, W, 3]")
    x = images_u8.contiguous()
    N, H, W, _ = x.shape
    out = torch.empty((N, 3, resolution, resolution), dtype=torch.float32, device=x.device)
    call("vcd_preprocess_u8", _p(x), _p(out), N, H, W, int(resolution), _st())
    return out if out_dtype in (None, torch.float32) else out.to(out_dtype)


class DevicePrefetcher:
    """Iterates an iterable of pinned host batches (tensors, or dicts of tensors as the reference's collate_fn returns them,
    data_utils.py:204-211) and yields them on `device`, copying batch i+1 on a copy stream while the caller trains on batch i
    — the device-side half of `DataLoader(pin_memory=True)` + `batch.to(device)` (train.py:283-286 via accelerate) without
    the copy on the critical path.  Every batch is still copied exactly once, from pinned memory, when it is about to be
    used; only its position on the timeline changes."""

    def __init__(self, batches, device):
        import torch
        self._torch = torch
        self.batches = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("DevicePrefetcher: device must be a CUDA device (no CPU path)")
        self.stream = torch.cuda.Stream(self.device)

    def _issue(self, batch):
        torch = self._torch
        with torch.cuda.stream(self.stream):
            if isinstance(batch, dict):
                out = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
            else:
                out = batch.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return out, ev

    def _ready(self, issued):
        torch = self._torch
        out, ev = issued
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in (out.values() if isinstance(out, dict) else (out,)):
            if torch.is_tensor(t):
                t.record_stream(cur)      # allocated on the copy stream, consumed on the caller's stream
        return out

    def __iter__(self):
        it = iter(self.batches)
        try:
            cur = self._issue(next(it))
        except StopIteration:
            return
        for nxt_host in it:
            nxt = self._issue(nxt_host)
            yield self._ready(cur)
            cur = nxt
        yield self._ready(cur)

    def __len__(self):
        return len(self.batches)
