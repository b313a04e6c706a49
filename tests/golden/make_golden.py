"""Generate golden fixtures by running the REFERENCE's own Python modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Nothing at test time reads /root/reference; the tests
replay these files through oracle/components.py (CPU) and the CUDA kernels (GPU).
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("VCD_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))
from tracking.monitor import ActivityMonitor  # noqa: E402
from tracking.deadneuron import DeadNeuronTracker  # noqa: E402
from classification.classifier import RegionClassifier  # noqa: E402
from intervention.nudger import InterventionHandler  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class Tiny(nn.Module):
    """conv -> GN(32,64) -> silu -> conv; channel counts legal for the CUDA GN kernel."""

    def __init__(self):
        super().__init__()
        self.conv_in = nn.Conv2d(3, 64, 3, padding=1)
        self.norm1 = nn.GroupNorm(32, 64, eps=1e-6)
        self.conv1 = nn.Conv2d(64, 64, 3, padding=1)

    def forward(self, x):
        return self.conv1(torch.nn.functional.silu(self.norm1(self.conv_in(x))))


def golden_monitor():
    torch.manual_seed(1234)
    m = Tiny()
    with torch.no_grad():
        m.norm1.weight.copy_(torch.rand(64) * 1.5 + 0.05)
        m.norm1.weight[::8] = 1e-3          # planted dead channels
        m.norm1.bias.copy_(torch.randn(64) * 0.05)
        m.norm1.bias[::8] = 0.0
    cfg = {
        "enabled": True, "track_interval": 2,
        "target_layers": [
            {"name": "conv_in", "capture_point": "output", "metrics": ["mean_abs_activation_per_channel"]},
            {"name": "norm1", "capture_point": "output",
             "metrics": ["mean_abs_activation_per_channel", "mean_activation", "std_activation"]},
            {"name": "norm1", "capture_point": "input", "metrics": ["mean_abs_activation_per_channel"]},
        ],
    }
    cap = {}
    h1 = m.norm1.register_forward_hook(lambda mod, i, o: cap.setdefault("gn_in", []).append(i[0].detach().clone())
                                       or cap.setdefault("gn_out", []).append(o.detach().clone()))
    mon = ActivityMonitor(m, cfg)
    xs = [torch.rand(2, 3, 16, 16) * 2 - 1, torch.rand(2, 3, 16, 16) * 2 - 1, torch.rand(1, 3, 16, 16) * 2 - 1]
    with torch.no_grad():
        for x in xs:
            m(x)
    assert mon.step(1) == {}            # off-interval: buffers keep accumulating (monitor.py:150-152)
    wb = mon.step(2)
    data = mon.get_data_for_step(2)
    recs = mon.export_all_processed_data_to_records()
    h1.remove()
    out = {f"x{i}": x.numpy() for i, x in enumerate(xs)}
    for i in range(3):
        out[f"gn_in{i}"] = cap["gn_in"][i].numpy()
        out[f"gn_out{i}"] = cap["gn_out"][i].numpy()
    out["gamma"] = m.norm1.weight.detach().numpy()
    out["beta"] = m.norm1.bias.detach().numpy()
    out["conv_in_w"] = m.conv_in.weight.detach().numpy()
    out["conv_in_b"] = m.conv_in.bias.detach().numpy()
    for lid, metrics in data.items():
        for k, v in metrics.items():
            out[f"data/{lid}/{k}"] = np.asarray(v)
    for k, v in wb.items():
        out[f"wandb/{k}"] = np.asarray(v)
    out["records_keys"] = np.array([f"{r['layer_identifier']}|{r['original_metric_name']}|{r['metric_type']}" for r in recs])
    out["records_vals"] = np.array([float(r["metric_value"]) for r in recs])
    np.savez_compressed(os.path.join(OUT, "monitor.npz"), **out)

    # classifier on the monitor output (thr 0.2 like the shipped configs)
    ccfg = {"enabled": True, "threshold": 0.2, "target_metric_key": "mean_abs_activation_per_channel",
            "layers_to_classify": ["norm1.output"]}
    clf = RegionClassifier(model=m, config=ccfg)
    res = clf.classify(data, 2)
    cout = {"threshold": np.float64(0.2)}
    for lid, r in res.items():
        cout[f"{lid}/idx"] = np.array(r["inactive_channel_indices"], dtype=np.int64)
        cout[f"{lid}/vals"] = np.array(r["values_of_inactive_channels"], dtype=np.float32)
        cout[f"{lid}/param"] = np.array(r["param_name_scale"])
    # synthetic edge vector: values at / next to the float32 threshold
    thr32 = np.float32(0.2)
    edge = np.array([thr32, np.nextafter(thr32, np.float32(0)), np.nextafter(thr32, np.float32(1)), 0.0, 0.19999, 0.2001,
                     1e-3, 5.0] + [0.5] * 56, dtype=np.float32)
    res2 = clf.classify({"norm1.output": {"mean_abs_activation_per_channel": edge}}, 4)
    cout["edge/vals_in"] = edge
    cout["edge/idx"] = np.array(res2["norm1.output"]["inactive_channel_indices"], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "classifier.npz"), **cout)

    # nudger on those results, fp32 and bf16 params, both strategies
    nout = {}
    for dt_name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        for strat in ("gentle_nudge_groupnorm_scale", "reset_groupnorm_scale"):
            torch.manual_seed(7)
            mm = Tiny().to(dt)
            g0 = (torch.rand(64) * 2.2 - 0.4)
            g0[3] = 1.45; g0[4] = 1.2499; g0[5] = 1.25; g0[6] = -0.7; g0[7] = 0.0
            with torch.no_grad():
                mm.norm1.weight.copy_(g0.to(dt))
            idx = [0, 3, 4, 5, 6, 7, 9, 17, 63, 64, 200, -1]
            ih = InterventionHandler(mm, {"enabled": True, "strategy": strat, "nudge_factor": 1.2,
                                          "max_scale_value": 1.5, "intervention_interval": 2})
            before = mm.norm1.weight.detach().float().numpy().copy()
            ih.intervene({"norm1.output": {"param_name_scale": "norm1.weight", "inactive_channel_indices": idx}}, 2)
            key = f"{dt_name}/{strat}"
            nout[f"{key}/before"] = before
            nout[f"{key}/after"] = mm.norm1.weight.detach().float().numpy()
            nout[f"{key}/count"] = np.int64(ih.num_nudges_applied)
            nout[f"{key}/idx"] = np.array(idx, dtype=np.int64)
            ih.intervene({"norm1.output": {"param_name_scale": "norm1.weight", "inactive_channel_indices": idx}}, 3)
            nout[f"{key}/after_offinterval"] = mm.norm1.weight.detach().float().numpy()
    np.savez_compressed(os.path.join(OUT, "nudger.npz"), **nout)


def golden_deadneuron():
    out = {}
    torch.manual_seed(99)
    tensors = {
        "w_conv": torch.randn(16, 8, 3, 3) * 0.02,
        "w_lin": torch.randn(32, 32) * 0.001,
        "w_gn": torch.rand(64) * 1e-3,
        "w_zero": torch.zeros(40),
        "w_tiny": torch.full((24,), 1e-12),
    }
    tensors["w_conv"].view(-1)[::7] = 1e-7
    tensors["w_lin"].view(-1)[::5] = 0.0
    for dt_name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        for dead_type in ("threshold", "percent_of_mean", "both"):
            trk = DeadNeuronTracker((nn.Conv2d,), [], threshold=1e-3, mean_percentage=0.1, dead_type=dead_type)
            for name, t in tensors.items():
                p = t.to(dt)
                out[f"{dt_name}/{dead_type}/{name}"] = np.float64(trk.get_percentage(p))
    for name, t in tensors.items():
        out[f"tensor/{name}"] = t.numpy()
    np.savez_compressed(os.path.join(OUT, "deadneuron.npz"), **out)


if __name__ == "__main__":
    golden_monitor()
    golden_deadneuron()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
