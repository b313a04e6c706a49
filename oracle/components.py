"""ORACLE (test infrastructure, never the product path).

numpy / torch-CPU restatement of the reference's tracker, classifier, nudger and
dead-weight arithmetic.  Each function cites the reference lines it follows.
Pinned against the reference itself: tests/golden/make_golden.py imports the
reference modules from /root/reference in the build container, runs them on seeded
inputs and commits inputs + outputs under tests/golden/*.npz; tests/test_oracle.py
replays those fixtures through the functions below.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch


# ---------------------------------------------------------------- tracker (monitor.py)
def mean_abs_per_channel(t: torch.Tensor) -> np.ndarray:
    """src/tracking/monitor.py:64-67 — |t|.mean over every dim but 1."""
    if t.ndim >= 2:
        return t.abs().mean(dim=[0] + list(range(2, t.ndim))).detach().cpu().numpy()
    return t.abs().mean().detach().cpu().numpy()


def mean_activation(t: torch.Tensor) -> np.ndarray:
    """src/tracking/monitor.py:72-73."""
    return t.mean().detach().cpu().numpy()


def std_activation(t: torch.Tensor) -> np.ndarray:
    """src/tracking/monitor.py:74-75 — torch.std is the unbiased estimator."""
    return t.std().detach().cpu().numpy()


def aggregate_per_channel(values: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    """src/tracking/monitor.py:176-186 — mean of the per-forward vectors (not pooled)."""
    agg = np.mean(np.stack(list(values)), axis=0)
    return {"value": agg, "overall_mean": np.mean(agg), "overall_std": np.std(agg)}


def aggregate_scalar(values: Sequence) -> float:
    """src/tracking/monitor.py:199-202."""
    return float(np.mean([v.item() if hasattr(v, "item") else float(v) for v in values]))


def per_channel_records(agg: np.ndarray) -> Dict[str, float]:
    """src/tracking/monitor.py:257-265 — CSV record values."""
    return {
        "per_channel_overall_mean": float(np.mean(agg)),
        "per_channel_overall_std": float(np.std(agg)),
        "per_channel_overall_min": float(np.min(agg)),
        "per_channel_overall_max": float(np.max(agg)),
    }


# ---------------------------------------------------------------- classifier (classifier.py)
def classify_indices(vals: np.ndarray, threshold: float) -> np.ndarray:
    """src/classification/classifier.py:135 — strict '<', float32 array vs python float."""
    return np.where(vals < threshold)[0]


# ---------------------------------------------------------------- nudger (nudger.py)
def nudge_gamma(gamma: torch.Tensor, idx: Sequence[int], factor: float, cap: float) -> int:
    """src/intervention/nudger.py:127-143 — in place; python-float product, min with cap,
    stored back in the parameter dtype; out-of-range indices skipped.  Returns #applied."""
    n = 0
    for i in idx:
        if 0 <= i < gamma.numel():
            v = gamma[i].item()
            gamma[i] = min(v * factor, cap)
            n += 1
    return n


def reset_gamma(gamma: torch.Tensor, idx: Sequence[int]) -> int:
    """src/intervention/nudger.py:162-168."""
    n = 0
    for i in idx:
        if 0 <= i < gamma.numel():
            gamma[i] = 1.0
            n += 1
    return n


def intervention_due(step: int, interval: int) -> bool:
    """src/intervention/nudger.py:94-97."""
    if step == 0 or step % interval != 0:
        if not (interval == 1 and step > 0):
            return False
    return True


# ---------------------------------------------------------------- dead-weight scan (deadneuron.py)
def dead_percent(param: torch.Tensor, threshold: float, mean_percentage: float, dead_type: str) -> float:
    """src/tracking/deadneuron.py:78-115."""
    n = param.numel()
    if n == 0:
        return 0.0
    a = param.abs()
    if dead_type == "threshold":
        return ((a < threshold).sum().item() / n) * 100.0
    mean_abs = a.mean().item()
    if dead_type == "percent_of_mean":
        if abs(mean_abs) < 1e-9:
            return 100.0 if (a < 1e-9).all().item() else 0.0
        return ((a < mean_percentage * mean_abs).sum().item() / n) * 100.0
    if dead_type == "both":
        fixed = a < threshold
        adaptive = (a < 1e-9) if abs(mean_abs) < 1e-9 else (a < mean_percentage * mean_abs)
        return ((fixed & adaptive).sum().item() / n) * 100.0
    return 0.0
