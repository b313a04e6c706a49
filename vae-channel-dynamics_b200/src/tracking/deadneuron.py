"""DeadNeuronTracker — the reference's src/tracking/deadneuron.py:11-115 API on one multi-tensor kernel.

The reference walks ``named_parameters()`` and issues 1-3 blocking ``.item()`` reductions per tensor
(248 tensors => several hundred host syncs per call).  Here all selected tensors are scanned by two
launches of ``vcd_dead_weight_count`` (sum|w| pass, predicate-count pass) followed by ONE device->host
copy; percentages are then formed on the host in float64 exactly as the reference does
(``(count / numel) * 100.0``, deadneuron.py:82,94,115)."""
from __future__ import annotations

import importlib
import logging
from collections import defaultdict
from typing import List, Tuple, Type

import numpy as np
import torch
import torch.nn as nn

_pkg = importlib.import_module("vae-channel-dynamics_b200")
_lib = _pkg._lib
ops = _pkg.ops

logger = logging.getLogger(__name__)

_DEAD_TYPES = {"threshold": 0, "percent_of_mean": 1, "both": 2}


class DeadNeuronTracker:
    def __init__(self, target_layer_classes: Tuple[Type[nn.Module], ...], target_layer_names_for_raw_weights: List[str],
                 threshold: float, mean_percentage: float, dead_type: str = "threshold"):
        self.threshold = threshold
        self.mean_percentage = mean_percentage
        self.target_layer_classes = target_layer_classes
        self.target_layer_names_for_raw_weights = target_layer_names_for_raw_weights
        self.dead_type = dead_type
        if dead_type not in _DEAD_TYPES:
            logger.warning(f"Unknown dead_type: {dead_type}. Defaulting to no-op for percentage calculation.")
        self.get_percentage = self._single_percentage
        self.weights_history = defaultdict(list)
        self.percent_history = defaultdict(list)
        self._table_key = None
        self._table = None

    # ------------------------------------------------------------------ selection (reference :37-76)
    def _select(self, vae: nn.Module):
        picked = []
        for name, param in vae.named_parameters():
            if not param.requires_grad:
                continue
            if name in self.target_layer_names_for_raw_weights:
                p = param.detach()
                self.weights_history[name] = [(p.float() if p.dtype == torch.bfloat16 else p).cpu().numpy()]
            if "weight" in name or "bias" in name:
                try:
                    module = vae.get_submodule(".".join(name.split(".")[:-1]))
                except AttributeError:
                    continue
                if isinstance(module, self.target_layer_classes):
                    picked.append((name, param))
        return picked

    def _percentages(self, params: List[torch.Tensor]) -> List[float]:
        if not params:
            return []
        if self.dead_type not in _DEAD_TYPES:
            return [0.0] * len(params)                                   # reference noop (:75-76)
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.VcdError("DeadNeuronTracker: parameters must be on a CUDA device (no CPU path)")
        key = tuple((p.data_ptr(), p.numel(), p.dtype) for p in params)
        if key != self._table_key:
            ptrs = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64, device=dev)
            numels = torch.tensor([p.numel() for p in params], dtype=torch.int64, device=dev)
            dts = torch.tensor([ops.dtype_code(p) for p in params], dtype=torch.int32, device=dev)
            self._table = (ptrs, numels, dts)
            self._table_key = key
        ptrs, numels, dts = self._table
        T = len(params)
        sums = torch.empty(T, dtype=torch.float64, device=dev)
        counts = torch.empty(T, dtype=torch.int64, device=dev)
        _lib.call("vcd_dead_weight_count", ptrs.data_ptr(), numels.data_ptr(), dts.data_ptr(), T,
                  float(self.threshold), float(self.mean_percentage), _DEAD_TYPES[self.dead_type],
                  sums.data_ptr(), counts.data_ptr(), torch.cuda.current_stream().cuda_stream)
        host = torch.stack([sums, counts.double()]).cpu().numpy()        # the single D2H of the call
        out = []
        for i, p in enumerate(params):
            n = p.numel()
            if n == 0:
                out.append(0.0)
                continue
            cnt = int(host[1, i])
            if self.dead_type == "percent_of_mean":
                mean_abs = float(torch.tensor(host[0, i] / n, dtype=torch.float64).to(p.dtype))
                if abs(mean_abs) < 1e-9:                                 # reference :88-90
                    out.append(100.0 if cnt == n else 0.0)
                    continue
            out.append((cnt / n) * 100.0)
        return out

    def _single_percentage(self, param: torch.Tensor) -> float:
        return self._percentages([param.detach()])[0]

    def track_dead_neurons(self, model_wrapper: nn.Module, global_step: int):
        if hasattr(model_wrapper, "vae") and model_wrapper.vae is not None:
            vae = model_wrapper.vae
        elif isinstance(model_wrapper, nn.Module):
            vae = model_wrapper
        else:
            logger.error("DeadNeuronTracker: model_wrapper is not an nn.Module or has no .vae attribute.")
            return
        try:
            picked = self._select(vae)
            pcts = self._percentages([p.detach() for _, p in picked])
        except Exception as e:  # reference :72-74 logs and carries on
            logger.error(f"DeadNeuronTracker (step {global_step}): error computing percentages: {e}")
            return
        for (name, _), pct in zip(picked, pcts):
            self.percent_history[name].append((global_step, pct))

    # kept for API compatibility with the reference's public helpers
    def noop(self, param):
        return 0.0

    def smaller_than_threshold(self, param):
        return DeadNeuronTracker((), [], self.threshold, self.mean_percentage, "threshold")._single_percentage(param)

    def percent_of_mean(self, param):
        return DeadNeuronTracker((), [], self.threshold, self.mean_percentage, "percent_of_mean")._single_percentage(param)

    def both(self, param):
        return DeadNeuronTracker((), [], self.threshold, self.mean_percentage, "both")._single_percentage(param)
