"""ActivityMonitor — the reference's src/tracking/monitor.py:11-271 API on device-side accumulators.

Reference behaviour kept: config keys, ``<layer>.<capture_point>`` identifiers, the metric names, the
mean-over-forwards aggregation (monitor.py:176-202), off-interval ``step()`` returning {} without draining
(:150-152), W&B key names (:183-201) and the CSV record schema (:229-270).

What changed underneath: the reference's hook computes ``tensor.abs().mean(dim=[0,2,3])`` and copies it
to the host on EVERY forward (:64-67).  Here GroupNorm / conv targets get a ``TrackSlot``: the statistics
come out of the GroupNorm kernels' own pass over the tensor (no extra HBM read), are normalised per
forward and accumulated on the device, and cross the PCIe bus once, inside ``step()``.  Targets that are
not B200 layers (or ask for ``full_activation_map``) use a real forward hook + the stand-alone
``vcd_chan_stats`` kernel.  With torch.distributed initialised, ``step()`` all-reduces the packed
accumulators once so every rank reports global-batch statistics.
"""
from __future__ import annotations

import importlib
import logging
from collections import defaultdict
from typing import Any, Dict, List, Optional

import numpy as np
import torch

_pkg = importlib.import_module("vae-channel-dynamics_b200")
ops = _pkg.ops
_vae = importlib.import_module("vae-channel-dynamics_b200.vae")

logger = logging.getLogger(__name__)

PER_CHANNEL = "mean_abs_activation_per_channel"
FULL_MAP = "full_activation_map"
SCALARS = ("mean_activation", "std_activation")
KNOWN = (PER_CHANNEL, FULL_MAP) + SCALARS


class _Handle:
    """Quacks like torch.utils.hooks.RemovableHandle for fused (hook-less) subscriptions."""

    def __init__(self, fn):
        self._fn = fn

    def remove(self):
        if self._fn is not None:
            self._fn()
            self._fn = None


class _Target:
    def __init__(self, identifier: str, metrics: List[str]):
        self.identifier = identifier
        self.metrics = metrics
        self.slot: Optional[ops.TrackSlot] = None
        self.channels: Optional[int] = None
        self.first_map: Optional[torch.Tensor] = None   # monitor.py:166-167 keeps only the first map
        self.map_count = 0


class ActivityMonitor:
    def __init__(self, model: torch.nn.Module, tracking_config: Dict[str, Any]):
        self.model = model
        self.config = tracking_config
        self.target_layers_config: List[Dict[str, Any]] = self.config.get("target_layers", [])
        self.hook_collected_buffer = defaultdict(lambda: defaultdict(list))  # kept for API parity; unused by slots
        self.processed_data_by_step = defaultdict(dict)
        self.extended_data_by_step = defaultdict(dict)   # variance / near-zero fraction / max-abs (north star)
        self.hooks: List[Any] = []
        self._targets: Dict[str, _Target] = {}
        self._fired: List[str] = []
        self.near_zero_threshold = float(self.config.get("near_zero_threshold", 0.0))
        if self.config.get("enabled", False):
            self._register_hooks()
            logger.info(f"ActivityMonitor initialized for {len(self.target_layers_config)} target(s).")
        else:
            logger.info("ActivityMonitor is disabled in config.")

    # ------------------------------------------------------------------ layer resolution (reference :41-54)
    def _get_layer(self, layer_name: str) -> torch.nn.Module:
        cur = self.model
        for part in layer_name.split("."):
            if hasattr(cur, part):
                cur = getattr(cur, part)
            elif hasattr(cur, "module") and hasattr(cur.module, part):   # DDP / accelerate wrapper
                cur = getattr(cur.module, part)
            else:
                raise AttributeError(f"Model (or its .module) does not have a layer named '{layer_name}' (path: {part})")
        return cur

    # ------------------------------------------------------------------ registration (reference :108-139)
    def _slot_for(self, tgt: _Target, channels: int, device) -> ops.TrackSlot:
        if tgt.slot is None or tgt.slot.C != channels or tgt.slot.raw.device != device:
            tgt.slot = ops.TrackSlot(channels, device, self.near_zero_threshold)
            tgt.slot.on_finalize = lambda ident=tgt.identifier: self._note_fired(ident)
            tgt.channels = channels
        return tgt.slot

    def _note_fired(self, ident: str):
        """The reference's buffer is keyed in first-fire order (monitor.py:101); step() reports in that order."""
        if ident not in self._fired:
            self._fired.append(ident)

    def _register_hooks(self):
        self.remove_hooks()
        self.hook_collected_buffer.clear()
        self._targets.clear()
        for conf in self.target_layers_config:
            name = conf.get("name")
            point = conf.get("capture_point", "output")
            if not name:
                logger.warning("Skipping a target_layer entry with no name.")
                continue
            ident = f"{name}.{point}"
            if point not in ("input", "output"):
                logger.warning(f"Unknown capture_point '{point}' for layer {name}. Skipping.")
                continue
            metrics = list(conf.get("metrics", [PER_CHANNEL]))
            for m in metrics:
                if m not in KNOWN:
                    logger.warning(f"Unknown metric '{m}' requested.")
            try:
                layer = self._get_layer(name)
            except AttributeError as e:
                logger.error(f"Could not register hook for {ident} (AttributeError): {e}")
                continue
            except Exception as e:  # reference :138-139
                logger.error(f"Unexpected error registering hook for {ident}: {e}", exc_info=True)
                continue
            tgt = self._targets.get(ident)
            if tgt is None:
                tgt = self._targets[ident] = _Target(ident, [])
            for m in metrics:
                if m in KNOWN and m not in tgt.metrics:
                    tgt.metrics.append(m)
            stat_metrics = [m for m in metrics if m == PER_CHANNEL or m in SCALARS]
            fused = False
            if stat_metrics:
                fused = self._subscribe_fused(layer, point, tgt)
            need_hook = (FULL_MAP in metrics) or (stat_metrics and not fused)
            if need_hook:
                self._register_torch_hook(layer, point, tgt, stats=bool(stat_metrics) and not fused,
                                          full_map=FULL_MAP in metrics)
            logger.info(f"Registered {'fused statistics' if fused else 'forward hook'} for layer: {name} ({point})")

    def _subscribe_fused(self, layer, point: str, tgt: _Target) -> bool:
        """GroupNorm input/output and conv output statistics come out of the layer's own kernels."""
        try:
            dev = next(layer.parameters()).device
        except StopIteration:
            return False
        if dev.type != "cuda":
            return False
        if isinstance(layer, _vae.B200GroupNorm):
            slot = self._slot_for(tgt, layer.num_channels, dev)
            attr = "_track_in" if point == "input" else "_track_out"
        elif isinstance(layer, _vae.B200Conv2d) and point == "output":
            slot = self._slot_for(tgt, layer.out_channels, dev)
            attr = "_track_out"
        else:
            return False
        setattr(layer, attr, slot)

        def undo(layer=layer, attr=attr, slot=slot):
            if getattr(layer, attr, None) is slot:
                setattr(layer, attr, None)
        self.hooks.append(_Handle(undo))
        return True

    def _register_torch_hook(self, layer, point: str, tgt: _Target, stats: bool, full_map: bool):
        monitor = self

        def consume(t):
            if not isinstance(t, torch.Tensor):
                return
            try:
                if stats and t.dim() >= 2:
                    slot = monitor._slot_for(tgt, t.shape[1], t.device)
                    ops.chan_stats(t.detach(), slot)
                if full_map:
                    monitor._note_fired(tgt.identifier)
                    if tgt.first_map is None:
                        d = t.detach()
                        tgt.first_map = (d.float() if d.dtype == torch.bfloat16 else d.clone()).cpu()
                    tgt.map_count += 1
            except Exception as e:  # reference :77-79: log, never raise into the training loop
                logger.error(f"Error calculating metrics for {tgt.identifier}: {e}", exc_info=True)

        if point == "input":
            def pre_hook(module, args):
                consume(args[0] if isinstance(args, tuple) and len(args) > 0 else args)
            self.hooks.append(layer.register_forward_pre_hook(pre_hook))
        else:
            def post_hook(module, args, output):
                consume(output)
            self.hooks.append(layer.register_forward_hook(post_hook))

    def remove_hooks(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []

    # ------------------------------------------------------------------ step (reference :146-216)
    def _gather(self):
        """One packed D2H (and, multi-rank, one packed all-reduce) of every slot."""
        tgts = [t for t in self._targets.values() if t.slot is not None]
        if not tgts:
            return {}
        dev = tgts[0].slot.run.device
        if any(t.slot.run.device != dev for t in tgts):   # model split over devices: pack on the first slot's device
            runs = [t.slot.run.to(dev) for t in tgts]
            scals = [t.slot.scal.to(dev) for t in tgts]
        else:
            runs, scals = [t.slot.run for t in tgts], [t.slot.scal for t in tgts]
        flat = torch.cat([r.double() for r in runs] + scals)
        if torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            # sums of per-forward means over all ranks in ONE packed SUM all-reduce; the running max|x| rows (row 3 of
            # every slot) are not additive: they travel in a second, tiny MAX all-reduce
            mx = torch.cat([r.view(5, -1)[3] for r in runs]).double()
            torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
            torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
            off = 0
            for r in runs:
                c = r.numel() // 5
                flat[off + 3 * c: off + 4 * c] = mx[off // 5: off // 5 + c]
                off += 5 * c
        host = flat.cpu().numpy()
        out, off = {}, 0
        for t in tgts:
            n = 5 * t.slot.C
            out[t.identifier] = [host[off:off + n].reshape(5, t.slot.C)]
            off += n
        for t in tgts:
            out[t.identifier].append(host[off:off + 3])
            off += 3
        return out

    def step(self, global_step: int) -> Dict[str, Any]:
        if not self.config.get("enabled", False):
            return {}
        if global_step % self.config.get("track_interval", 100) != 0:
            return {}
        wandb_metrics: Dict[str, Any] = {}
        processed: Dict[str, Dict[str, Any]] = {}
        extended: Dict[str, Dict[str, Any]] = {}
        gathered = self._gather()
        order = [i for i in self._fired if i in self._targets]
        # forwards replayed from a CUDA graph do not run the Python callback: append the remaining subscribed
        # targets in registration order (their device accumulators say whether they fired)
        order += [i for i in self._targets if i not in order]
        for ident in order:
            tgt = self._targets[ident]
            entry: Dict[str, Any] = {}
            run, scal = gathered.get(ident, (None, None))
            forwards = float(scal[2]) if scal is not None else 0.0
            for metric in tgt.metrics:
                if metric == FULL_MAP:
                    if tgt.first_map is None:
                        continue
                    arr = tgt.first_map.numpy().astype(np.float32)
                    wandb_metrics[f"tracking/{ident}/{metric}_mean"] = np.mean(arr)
                    wandb_metrics[f"tracking/{ident}/{metric}_std"] = np.std(arr)
                    entry[metric] = tgt.first_map
                elif forwards > 0 and metric == PER_CHANNEL:
                    vec = (run[0] / forwards).astype(np.float32)
                    wandb_metrics[f"tracking/{ident}/{metric}_overall_mean"] = np.mean(vec)
                    wandb_metrics[f"tracking/{ident}/{metric}_overall_std"] = np.std(vec)
                    entry[metric] = vec
                elif forwards > 0 and metric in SCALARS:
                    val = np.float64(scal[0 if metric == "mean_activation" else 1] / forwards)
                    wandb_metrics[f"tracking/{ident}/{metric}"] = val
                    entry[metric] = val
            if entry:
                processed[ident] = entry
            if forwards > 0 and run is not None:
                extended[ident] = {
                    "mean_per_channel": (run[1] / forwards).astype(np.float32),
                    "variance_per_channel": (run[2] / forwards).astype(np.float32),
                    "max_abs_per_channel": run[3].astype(np.float32),
                    "near_zero_fraction_per_channel": (run[4] / forwards).astype(np.float32),
                    "forwards": int(forwards),
                }
            if tgt.slot is not None:
                tgt.slot.reset()
            tgt.first_map, tgt.map_count = None, 0
        if processed:
            self.processed_data_by_step[global_step] = processed
            self.extended_data_by_step[global_step] = extended
            logger.info(f"ActivityMonitor collected and processed data for step {global_step}.")
        self.hook_collected_buffer.clear()
        self._fired = []
        self._mark_gamma_sync()
        return wandb_metrics

    def _mark_gamma_sync(self):
        """All ranks pass here on the same step (train.py:308-309); rank 0 may nudge right after
        (train.py:315-319), so every rank re-synchronises GroupNorm scales at its next forward."""
        for m in self.model.modules():
            if isinstance(m, _pkg.B200AutoencoderKL):
                m._gamma_sync_pending = True

    def get_data_for_step(self, global_step: int) -> Dict[str, Any]:
        return self.processed_data_by_step.get(global_step, {})

    def get_extended_stats_for_step(self, global_step: int) -> Dict[str, Any]:
        return self.extended_data_by_step.get(global_step, {})

    # ------------------------------------------------------------------ CSV records (reference :221-271)
    def export_all_processed_data_to_records(self) -> List[Dict[str, Any]]:
        records: List[Dict[str, Any]] = []

        def add(base, kind, value):
            records.append({**base, "metric_type": kind, "metric_value": value})

        for step, step_data in self.processed_data_by_step.items():
            for ident, metrics in step_data.items():
                for metric, value in metrics.items():
                    base = {"global_step": step, "layer_identifier": ident, "original_metric_name": metric}
                    if isinstance(value, torch.Tensor):
                        arr = value.numpy()
                    elif isinstance(value, np.ndarray):
                        arr = value
                    else:
                        add(base, "scalar", float(value))
                        continue
                    if arr.ndim == 0:
                        add(base, "scalar", float(arr.item()))
                    elif metric == FULL_MAP:
                        f = arr.astype(np.float32)
                        add(base, "full_map_shape", str(arr.shape))
                        add(base, "full_map_mean", float(np.mean(f)))
                        add(base, "full_map_std", float(np.std(f)))
                        add(base, "full_map_min", float(np.min(f)))
                        add(base, "full_map_max", float(np.max(f)))
                    elif PER_CHANNEL in metric:
                        add(base, "per_channel_overall_mean", float(np.mean(arr)))
                        add(base, "per_channel_overall_std", float(np.std(arr)))
                        add(base, "per_channel_overall_min", float(np.min(arr)))
                        add(base, "per_channel_overall_max", float(np.max(arr)))
                    else:
                        f = arr.astype(np.float32)
                        add(base, "array_mean", float(np.mean(f)))
                        add(base, "array_std", float(np.std(f)))
        return records

    def __del__(self):
        try:
            self.remove_hooks()
        except Exception:
            pass
