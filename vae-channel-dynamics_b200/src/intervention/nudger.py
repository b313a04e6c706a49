"""InterventionHandler — the reference's src/intervention/nudger.py:10-172 API; the per-channel
``.item()`` read / scalar write loop (nudger.py:128-143) is one ``vcd_nudge_gamma`` launch per layer
that mutates the parameter's own storage in place (the optimizer and DDP keep seeing it)."""
from __future__ import annotations

import importlib
import logging
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

_pkg = importlib.import_module("vae-channel-dynamics_b200")
_lib = _pkg._lib
ops = _pkg.ops

logger = logging.getLogger(__name__)


class InterventionHandler:
    def __init__(self, model: nn.Module, config: Dict[str, Any]):
        self.model = model
        self.config = config
        self.strategy = config.get("strategy", "none")
        self.nudge_factor = float(config.get("nudge_factor", 1.1))
        self.nudge_value_add = float(config.get("nudge_value_add", 0.01))
        self.max_scale_value = float(config.get("max_scale_value", 2.0))
        self.num_nudges_applied = 0
        logger.info(f"InterventionHandler initialized (strategy: {self.strategy}, model type: {type(model)})")
        if not isinstance(model, nn.Module):
            logger.warning(f"InterventionHandler received a model of type {type(model)}, expected nn.Module.")

    def _get_parameter(self, param_name: str) -> Optional[nn.Parameter]:
        """reference :49-72 — dotted path from self.model; None (and a log line) when it does not resolve."""
        try:
            cur = self.model
            for part in param_name.split("."):
                if not hasattr(cur, part):
                    logger.error(f"Model does not have attribute '{part}' in path '{param_name}'.")
                    return None
                cur = getattr(cur, part)
            if isinstance(cur, nn.Parameter):
                return cur
            logger.error(f"Attribute '{param_name}' is not a Parameter, but {type(cur)}.")
            return None
        except Exception as e:
            logger.error(f"Error getting parameter '{param_name}': {e}", exc_info=True)
            return None

    def _apply(self, param: nn.Parameter, indices, mode: int) -> int:
        data = param.data
        if data.device.type != "cuda":
            raise _lib.VcdError("InterventionHandler: parameters must be on a CUDA device (no CPU path)")
        if not data.is_contiguous():
            raise _lib.VcdError("InterventionHandler: GroupNorm scale must be contiguous")
        # the reference walks the list sequentially (nudger.py:128-143), so a repeated index is nudged (and rounded to
        # the parameter dtype, and counted) once per occurrence: pass k holds the indices that occur more than k times —
        # no two threads of one launch touch the same element
        ints = [int(i) for i in indices]
        mult: Dict[int, int] = {}
        for i in ints:
            mult[i] = mult.get(i, 0) + 1
        n = 0
        applied = torch.zeros(1, dtype=torch.int32, device=data.device)
        with torch.cuda.device(data.device):
            for k in range(max(mult.values())):
                idx = torch.tensor([i for i, m in mult.items() if m > k], dtype=torch.int64, device=data.device)
                _lib.call("vcd_nudge_gamma", data.data_ptr(), ops.dtype_code(data), data.numel(), idx.data_ptr(),
                          idx.numel(), self.nudge_factor, self.max_scale_value, mode, applied.data_ptr(),
                          torch.cuda.current_stream().cuda_stream)
                n += int(applied.item())
        for i in ints:
            if not (0 <= i < data.numel()):
                logger.warning(f"Inactive index {i} out of bounds (size: {data.numel()})")
        return n

    def intervene(self, classification_results: Dict[str, Any], global_step: int):
        """reference :74-172 — same guards, same counters."""
        if not self.config.get("enabled", False) or self.strategy == "none":
            return
        interval = self.config.get("intervention_interval", 200)
        if global_step == 0 or global_step % interval != 0:
            if not (interval == 1 and global_step > 0):
                return
        logger.info(f"InterventionHandler attempting intervention at step {global_step} with strategy '{self.strategy}'.")
        if not classification_results:
            logger.info(f"Step {global_step}: No regions classified by RegionClassifier, skipping intervention.")
            return
        self.num_nudges_applied = 0
        if self.strategy == "gentle_nudge_groupnorm_scale":
            mode, warn = 0, True
        elif self.strategy == "reset_groupnorm_scale":
            mode, warn = 1, False
        else:
            logger.warning(f"Unknown intervention strategy: {self.strategy}")
            return
        for layer_key, data in classification_results.items():
            name = data.get("param_name_scale")
            indices = data.get("inactive_channel_indices")
            if not name or indices is None:
                if warn:
                    logger.warning(f"Missing 'param_name_scale' or 'inactive_channel_indices' for '{layer_key}'. Skipping.")
                continue
            param = self._get_parameter(name)
            if param is None:
                if warn:
                    logger.warning(f"Could not retrieve scale parameter '{name}' for '{layer_key}'. Skipping.")
                continue
            if len(indices) == 0:
                continue
            with torch.no_grad():
                self.num_nudges_applied += self._apply(param, indices, mode)
        if self.num_nudges_applied > 0:
            logger.info(f"Applied '{self.strategy}' to {self.num_nudges_applied} channel scales at step {global_step}.")
