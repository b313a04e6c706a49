"""CUDA-graph capture of the training step's device work (forward + loss + backward).

The step issues ~1 100 short kernel launches; at the low-resolution layers the GPU finishes them faster than the
host can enqueue the next one.  Capturing forward + loss (src/train.py:287-291) + backward (train.py:299) once
and replaying the graph removes those launch gaps.  The optimizer, gradient clipping, monitor.step / classify /
intervene stay eager, exactly where train.py:300-330 runs them.

Gradients accumulate into views of ONE flat buffer, so data-parallel training needs a single NCCL all-reduce of
that buffer after the replay (167 MB in bf16; no per-bucket launches).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from .losses import vae_loss


class GraphedVAEStep:
    """step(pixel_values) -> (total, rec, kl) detached scalars; parameters' .grad are filled (and averaged
    across ranks) after the call.  `wrapper` is an SDXLVAEWrapper whose parameters live on a CUDA device."""

    def __init__(self, wrapper: torch.nn.Module, kl_weight: float, example: torch.Tensor, warmup: int = 3):
        self.wrapper = wrapper
        self.kl_weight = float(kl_weight)
        self.params = [p for p in wrapper.parameters() if p.requires_grad]
        dt = self.params[0].dtype
        dev = self.params[0].device
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=dt, device=dev)
        self._views = []
        off = 0
        for p in self.params:
            self._views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._attach_grads()
        self.static_x = example.detach().clone()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._out = None
        self._capture(warmup)

    def _attach_grads(self):
        """`p.grad` of every parameter IS its view of the flat buffer the captured backward accumulates into.
        The reference loop calls optimizer.zero_grad(set_to_none=True) every step (train.py:304), which drops these
        aliases; they are restored before every replay (the buffer itself is zeroed inside the graph)."""
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is None or g.data_ptr() != v.data_ptr():
                p.grad = v

    def _slots(self):
        seen = []
        for m in self.wrapper.modules():
            for a in ("_track_in", "_track_out"):
                s = getattr(m, a, None)
                if s is not None and all(s is not t for t in seen):
                    seen.append(s)
        return seen

    def _fwd_bwd(self):
        self.flat.zero_()
        out = self.wrapper(self.static_x, sample_posterior=True)
        total, rec, kl = vae_loss(out, self.static_x, self.kl_weight)
        (total / self.world).backward()        # DDP semantics: mean of the per-rank means
        return total.detach(), rec.detach(), kl.detach()

    def _capture(self, warmup: int):
        # the warm-up and capture forwards run on an example batch: the statistics slots of a subscribed
        # ActivityMonitor must not count them (snapshot here, restore after the capture)
        saved = [(s, s.raw.clone(), s.run.clone(), s.scal.clone()) for s in self._slots()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):     # allocator warm-up, weight packs, autograd threads' CUDA context
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.launches
        with torch.cuda.graph(self.graph):
            self._out = self._fwd_bwd()
        self.launches_per_replay = _lib.launches - l0   # C-ABI compute calls recorded in the graph
        for s, raw, run, scal in saved:
            s.raw.copy_(raw)
            s.run.copy_(run)
            s.scal.copy_(scal)

    def refresh(self):
        """Re-capture (e.g. after hooks were added/removed, which changes the kernel sequence)."""
        self._capture(1)

    def step(self, pixel_values: torch.Tensor):
        vae = getattr(self.wrapper, "vae", None)
        if vae is not None and hasattr(vae, "_sync_gamma_if_pending"):
            vae._sync_gamma_if_pending()        # eager: Python-side flag, not part of the graph
        self._attach_grads()
        self.static_x.copy_(pixel_values, non_blocking=True)
        self.graph.replay()
        if self.world > 1:
            dist.all_reduce(self.flat)          # gradients were pre-scaled by 1/world
        return self._out
