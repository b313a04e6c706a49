"""FusedClipAdamW — gradient-norm clip + AdamW in two kernel launches.

The reference builds ``torch.optim.AdamW(params, lr, betas, weight_decay, eps)`` itself (train.py:184-187) and calls
``accelerator.clip_grad_norm_`` + ``optimizer.step()`` (train.py:301-302); with torch's foreach implementation that
is ~25 multi-tensor launches and ~10 passes over 84 M parameters every step.  This class keeps torch's AdamW
arithmetic (decoupled weight decay, bias correction, the 1e-6 of clip_grad_norm_) but runs it as ONE pass over the
gradients (sum of squares) and ONE pass over (param, grad, exp_avg, exp_avg_sq) — `vcd_multi_sqnorm`,
`vcd_clip_adamw_step`.  Parameters / gradients stay fp32 or bf16 exactly as train.py:150-154 loads them; the moments
are fp32 (torch keeps them in the parameter dtype, i.e. bf16 moments for bf16 parameters).

Opt-in that leaves train.py untouched: ``FusedClipAdamW.from_torch(optimizer)`` adopts an already constructed
torch.optim.AdamW — it SHARES that optimizer's ``param_groups`` dicts, so a LambdaLR built on the original object
(train.py:202) keeps driving the learning rate.  An accelerate-compatible ``prepare()`` does the swap when
``VCD_FUSED_OPT=1`` (tests/shims/accelerate); bench.py uses the class directly.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import _lib
from .ops import _st, call, dtype_code


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = None
        self._pending_clip: Optional[float] = None
        self._sqnorm: Optional[torch.Tensor] = None
        self._steps = 0

    @classmethod
    def from_torch(cls, opt: torch.optim.Optimizer) -> "FusedClipAdamW":
        """Adopt a constructed torch.optim.AdamW (no steps taken yet): same parameters and hyper-parameters, and the
        SAME param_group dicts, so schedulers holding `opt` keep working."""
        if any(len(s) for s in opt.state.values()):
            raise _lib.VcdError("FusedClipAdamW.from_torch: the optimizer already holds state (call it before the first step)")
        for g in opt.param_groups:
            if g.get("amsgrad") or g.get("maximize"):
                raise _lib.VcdError("FusedClipAdamW: amsgrad / maximize are not supported")
        self = cls.__new__(cls)
        torch.optim.Optimizer.__init__(self, [dict(g) for g in opt.param_groups],
                                       dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2))
        self.param_groups = opt.param_groups           # shared with the adopted optimizer (LambdaLR writes 'lr' here)
        self._tables = None
        self._pending_clip = None
        self._sqnorm = None
        self._steps = 0
        self.adopted = opt
        return self

    # ---- device tables -------------------------------------------------------------------------------
    def _build(self):
        groups = []
        chunk = _lib.lib().vcd_optim_chunk_elems()
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise _lib.VcdError("FusedClipAdamW: parameters must live on a CUDA device (no CPU path)")
            for p in ps:
                if not p.is_contiguous():
                    raise _lib.VcdError("FusedClipAdamW: parameters must be contiguous")
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros(p.shape, dtype=torch.float32, device=dev)
                    st["exp_avg_sq"] = torch.zeros(p.shape, dtype=torch.float32, device=dev)
            T = len(ps)
            ct, co = [], []
            for t, p in enumerate(ps):
                for off in range(0, p.numel(), chunk):
                    ct.append(t)
                    co.append(off)
            i64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
            tab = {
                "params": ps, "device": dev, "T": T, "n_chunks": len(ct),
                "p": i64([p.data_ptr() for p in ps]),
                "m": i64([self.state[p]["exp_avg"].data_ptr() for p in ps]),
                "v": i64([self.state[p]["exp_avg_sq"].data_ptr() for p in ps]),
                "numel": i64([p.numel() for p in ps]),
                "dtype": torch.tensor([dtype_code(p) for p in ps], dtype=torch.int32, device=dev),
                "chunk_tensor": torch.tensor(ct, dtype=torch.int32, device=dev),
                "chunk_off": i64(co),
                "g": torch.zeros(T, dtype=torch.int64, device=dev),
                "g_host": [torch.zeros(T, dtype=torch.int64).pin_memory() for _ in range(4)],
                "g_ptrs": None, "turn": 0,
                "sqnorm": torch.zeros(1, dtype=torch.float64, device=dev),
                "steps_dev": torch.zeros(T, dtype=torch.int32, device=dev), "uniform": True,
            }
            groups.append((g, tab))
        self._tables = groups

    def _refresh_grad_table(self, tab):
        """Gradients are re-allocated every step (zero_grad(set_to_none=True), train.py:304): their addresses travel in
        one 2 KB pinned-memory copy, skipped when nothing moved."""
        ptrs = []
        for p in tab["params"]:
            g = p.grad
            if g is None:
                ptrs.append(0)
                continue
            if g.dtype != p.dtype or not g.is_contiguous() or g.is_sparse:
                raise _lib.VcdError("FusedClipAdamW: gradients must be dense, contiguous and of the parameter dtype")
            ptrs.append(g.data_ptr())
        if ptrs != tab["g_ptrs"]:
            host = tab["g_host"][tab["turn"] % 4]
            tab["turn"] += 1
            host.copy_(torch.tensor(ptrs, dtype=torch.int64))
            tab["g"].copy_(host, non_blocking=True)
            tab["g_ptrs"] = ptrs

    # ---- torch.nn.utils.clip_grad_norm_ (train.py:301) -----------------------------------------------
    @torch.no_grad()
    def clip_grad_norm_(self, parameters: Optional[Iterable[torch.Tensor]] = None, max_norm: float = 1.0) -> torch.Tensor:
        """Computes the global 2-norm of all gradients now (one launch; the value stays on the device and is returned
        as a 0-d tensor like torch's) and remembers `max_norm`: the scaling itself happens inside the next step()."""
        if self._tables is None:
            self._build()
        if len(self._tables) != 1:
            raise _lib.VcdError("FusedClipAdamW.clip_grad_norm_: a single param group is supported (train.py:184 builds one)")
        _, tab = self._tables[0]
        self._refresh_grad_table(tab)
        with torch.cuda.device(tab["device"]):
            call("vcd_multi_sqnorm", tab["g"].data_ptr(), tab["numel"].data_ptr(), tab["dtype"].data_ptr(),
                 tab["chunk_tensor"].data_ptr(), tab["chunk_off"].data_ptr(), tab["n_chunks"], tab["sqnorm"].data_ptr(), _st())
        self._pending_clip = float(max_norm)
        return tab["sqnorm"].sqrt().to(torch.float32).reshape(())

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._tables is None:
            self._build()
        self._steps += 1
        for g, tab in self._tables:
            self._refresh_grad_table(tab)
            b1, b2 = g["betas"]
            clip = self._pending_clip if self._pending_clip is not None else 0.0
            # torch counts steps per parameter (a parameter without a gradient is skipped and lags behind); the table
            # only travels to the device once the counters have diverged — never in the reference's training loop
            steps = []
            for p in tab["params"]:
                st = self.state[p]
                if p.grad is not None:
                    st["step"] = int(st["step"]) + 1
                steps.append(int(st["step"]))
            tab["uniform"] = tab["uniform"] and all(s == steps[0] for s in steps)
            steps_ptr = None
            if not tab["uniform"]:
                tab["steps_dev"].copy_(torch.tensor(steps, dtype=torch.int32))
                steps_ptr = tab["steps_dev"].data_ptr()
            with torch.cuda.device(tab["device"]):
                call("vcd_clip_adamw_step", tab["p"].data_ptr(), tab["g"].data_ptr(), tab["m"].data_ptr(), tab["v"].data_ptr(),
                     tab["numel"].data_ptr(), tab["dtype"].data_ptr(), tab["chunk_tensor"].data_ptr(),
                     tab["chunk_off"].data_ptr(), tab["n_chunks"], tab["sqnorm"].data_ptr() if clip > 0 else None,
                     float(clip), float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                     steps_ptr, max(1, steps[0]), _st())
        self._pending_clip = None
        return loss
