"""RegionClassifier — the reference's src/classification/classifier.py:10-151 API; the threshold
compare (classifier.py:135, strict '<' in float32) runs as ``vcd_classify_mask`` on the device."""
from __future__ import annotations

import importlib
import logging
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

_pkg = importlib.import_module("vae-channel-dynamics_b200")
_lib = _pkg._lib

logger = logging.getLogger(__name__)


class RegionClassifier:
    def __init__(self, model: torch.nn.Module, config: Dict[str, Any]):
        self.config = config
        self.method = config.get("method", "threshold_groupnorm_activity")
        self.threshold = float(config.get("threshold", 1e-3))
        self.target_metric_key = config.get("target_metric_key", "mean_abs_activation_per_channel")
        self.layers_to_classify: List[str] = config.get("layers_to_classify", [])
        self._layer_to_param_map: Dict[str, Tuple[str, int]] = {}
        self._device = None
        if model is not None:
            self._build_groupnorm_map(model)
            try:
                self._device = next(model.parameters()).device
            except StopIteration:
                pass
        else:
            logger.warning("RegionClassifier initialised without a model – parameter mapping will be heuristic only.")
        logger.info(f"RegionClassifier initialised (method={self.method}, thr={self.threshold}, "
                    f"metric={self.target_metric_key}, map_size={len(self._layer_to_param_map)})")
        if not self._layer_to_param_map:
            logger.warning("RegionClassifier: no GroupNorm layers found / mapped.")

    # ------------------------------------------------------------------ mapping (reference :43-95)
    def _build_groupnorm_map(self, model: torch.nn.Module):
        """monitor key '<gn>.output' (and its 'vae.'-prefixed alias) -> ('<gn>.weight', num_channels)."""
        for mod_name, mod in model.named_modules():
            if not isinstance(mod, torch.nn.GroupNorm):
                continue
            if not isinstance(getattr(mod, "weight", None), torch.nn.Parameter):
                logger.debug(f"Skipping {mod_name}: scale param is not a Parameter.")
                continue
            info = (f"{mod_name}.weight", mod.num_channels)
            self._layer_to_param_map[f"{mod_name}.output"] = info
            if not mod_name.startswith("vae."):
                self._layer_to_param_map[f"vae.{mod_name}.output"] = info

    def _lookup_param_info(self, layer_id: str) -> Optional[Tuple[str, int]]:
        info = self._layer_to_param_map.get(layer_id)
        if info is None and "." in layer_id:      # retry once without the leading scope (e.g. "vae.")
            info = self._layer_to_param_map.get(layer_id.split(".", 1)[1])
        return info

    # ------------------------------------------------------------------ device compare
    def _inactive_indices(self, vals: np.ndarray) -> np.ndarray:
        dev = self._device if self._device is not None and self._device.type == "cuda" else None
        if dev is None:
            if not torch.cuda.is_available():
                raise _lib.VcdError("RegionClassifier: no CUDA device (the classifier has no CPU path)")
            dev = torch.device("cuda", torch.cuda.current_device())
        C = int(vals.shape[0])
        v = torch.from_numpy(np.ascontiguousarray(vals, dtype=np.float32)).to(dev)
        mask = torch.empty(C, dtype=torch.uint8, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("vcd_classify_mask", v.data_ptr(), float(np.float32(self.threshold)), mask.data_ptr(),
                      count.data_ptr(), C, torch.cuda.current_stream().cuda_stream)
        return np.nonzero(mask.cpu().numpy())[0]

    # ------------------------------------------------------------------ public API (reference :100-151)
    def classify(self, tracked_data_for_step: Dict[str, Any], global_step: int) -> Dict[str, Any]:
        if not self.config.get("enabled", False):
            return {}
        results: Dict[str, Any] = {}
        if self.method != "threshold_groupnorm_activity":
            logger.warning(f"Unknown classification method: {self.method}")
            return results
        if not tracked_data_for_step:
            return results
        for layer_id, metrics in tracked_data_for_step.items():
            if self.layers_to_classify and layer_id not in self.layers_to_classify:
                continue
            vals = metrics.get(self.target_metric_key)
            if not (isinstance(vals, np.ndarray) and vals.ndim == 1):
                continue
            info = self._lookup_param_info(layer_id)
            if info is None:
                logger.debug(f"{layer_id}: no GN mapping found – skipped.")
                continue
            param_name_scale, num_ch = info
            if vals.shape[0] != num_ch:
                logger.warning(f"{layer_id}: channel mismatch ({vals.shape[0]} vs {num_ch}) – skipped.")
                continue
            idx = self._inactive_indices(vals)
            if idx.size == 0:
                continue
            results[layer_id] = {
                "param_name_scale": param_name_scale,
                "inactive_channel_indices": idx.tolist(),
                "metric_used": self.target_metric_key,
                "threshold_value": self.threshold,
                "values_of_inactive_channels": vals[idx].tolist(),
            }
            logger.info(f"Step {global_step}: {layer_id} → {len(idx)} inactive channels (param {param_name_scale})")
        logger.info(f"Classification complete — {len(results)} layer(s) flagged.")
        return results
