"""CPU: the oracle (oracle/) pinned against the golden fixtures generated from the REFERENCE's own modules
(tests/golden/make_golden.py) and against the reference's `__main__` known answers."""
import os

import numpy as np
import torch
import torch.nn as nn

from oracle import components as oc
from oracle.torch_vae import build_oracle, oracle_forward, oracle_losses


def test_oracle_module_tree_matches_published_sdxl_vae_size():
    vae = build_oracle(42)
    params = list(vae.named_parameters())
    assert len(params) == 248
    assert sum(p.numel() for _, p in params) == 83_653_863
    assert params[0][0] == "encoder.conv_in.weight" and params[-1][0] == "post_quant_conv.bias"
    gns = [m for m in vae.modules() if isinstance(m, nn.GroupNorm)]
    assert len(gns) == 52 and sum(m.num_channels for m in gns) == 19_840
    assert all(m.num_groups == 32 and m.eps == 1e-6 for m in gns)
    assert len([m for m in vae.modules() if isinstance(m, nn.Conv2d)]) == 64
    for name in ("encoder.down_blocks.0.resnets.0.norm1", "decoder.up_blocks.1.resnets.0.norm1", "decoder.conv_norm_out",
                 "encoder.down_blocks.1.resnets.0.conv_shortcut", "decoder.up_blocks.2.resnets.2.norm2",
                 "encoder.mid_block.attentions.0.to_out.0"):
        vae.get_submodule(name)   # names the shipped configs / evaluate.py reference


def test_oracle_forward_shapes_and_loss_definition():
    torch.manual_seed(0)
    vae = build_oracle(42)
    x = torch.rand(1, 3, 32, 32) * 2 - 1
    noise = torch.randn(1, 4, 4, 4)
    out = oracle_forward(vae, x, True, noise=noise)
    assert out["reconstruction"].shape == x.shape and out["latents_sampled"].shape == (1, 4, 4, 4)
    d = out["latent_dist"]
    assert torch.allclose(out["latents_sampled"], d.mean + d.std * noise)
    total, rec, kl = oracle_losses(out, x, 1e-6)
    assert torch.allclose(rec, ((out["reconstruction"] - x) ** 2).mean())
    assert torch.allclose(kl, (0.5 * (d.mean ** 2 + d.var - 1 - d.logvar).sum(dim=[1, 2, 3])).mean())
    assert torch.allclose(total, rec + 1e-6 * kl)
    assert torch.equal(oracle_forward(vae, x, False)["latents_sampled"], d.mean)


def test_tracker_formulas_match_reference_monitor_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "monitor.npz"))
    cases = {"norm1.input": [torch.from_numpy(g[f"gn_in{i}"]) for i in range(3)],
             "norm1.output": [torch.from_numpy(g[f"gn_out{i}"]) for i in range(3)],
             "conv_in.output": [torch.from_numpy(g[f"gn_in{i}"]) for i in range(3)]}   # conv_in output IS norm1's input
    for lid, tensors in cases.items():
        agg = oc.aggregate_per_channel([oc.mean_abs_per_channel(t) for t in tensors])
        want = g[f"data/{lid}/mean_abs_activation_per_channel"]
        assert agg["value"].dtype == np.float32 and np.array_equal(agg["value"], want), lid   # bit-exact
        assert np.float32(agg["overall_mean"]) == g[f"wandb/tracking/{lid}/mean_abs_activation_per_channel_overall_mean"]
        assert np.float32(agg["overall_std"]) == g[f"wandb/tracking/{lid}/mean_abs_activation_per_channel_overall_std"]
    outs = cases["norm1.output"]
    assert oc.aggregate_scalar([oc.mean_activation(t) for t in outs]) == float(g["data/norm1.output/mean_activation"])
    assert oc.aggregate_scalar([oc.std_activation(t) for t in outs]) == float(g["data/norm1.output/std_activation"])
    keys, vals = list(g["records_keys"]), g["records_vals"]
    rec = oc.per_channel_records(g["data/norm1.output/mean_abs_activation_per_channel"])
    for kind, v in rec.items():
        assert v == vals[keys.index(f"norm1.output|mean_abs_activation_per_channel|{kind}")]


def test_classifier_formula_matches_reference_golden(golden_dir):
    gm = np.load(os.path.join(golden_dir, "monitor.npz"))
    gc = np.load(os.path.join(golden_dir, "classifier.npz"))
    vals = gm["data/norm1.output/mean_abs_activation_per_channel"]
    assert np.array_equal(oc.classify_indices(vals, 0.2), gc["norm1.output/idx"])
    assert np.array_equal(oc.classify_indices(gc["edge/vals_in"], 0.2), gc["edge/idx"])
    assert set(range(0, 64, 8)) <= set(gc["norm1.output/idx"].tolist())  # the planted small-gamma channels


def test_nudger_formula_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nudger.npz"))
    for dt_name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        for strat in ("gentle_nudge_groupnorm_scale", "reset_groupnorm_scale"):
            key = f"{dt_name}/{strat}"
            gamma = torch.from_numpy(g[f"{key}/before"]).to(dt)
            idx = g[f"{key}/idx"].tolist()
            n = oc.nudge_gamma(gamma, idx, 1.2, 1.5) if strat.startswith("gentle") else oc.reset_gamma(gamma, idx)
            assert n == int(g[f"{key}/count"])
            assert np.array_equal(gamma.float().numpy(), g[f"{key}/after"]), key
    assert [oc.intervention_due(s, 10) for s in (0, 5, 10, 20)] == [False, False, True, True]
    assert [oc.intervention_due(s, 1) for s in (0, 1, 2)] == [False, True, True]


def test_deadneuron_formula_matches_reference_golden_and_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "deadneuron.npz"))
    for dt_name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        for dead_type in ("threshold", "percent_of_mean", "both"):
            for name in ("w_conv", "w_lin", "w_gn", "w_zero", "w_tiny"):
                p = torch.from_numpy(g[f"tensor/{name}"]).to(dt)
                assert oc.dead_percent(p, 1e-3, 0.1, dead_type) == float(g[f"{dt_name}/{dead_type}/{name}"])
    # deadneuron.py:183-202
    w = torch.full((8, 3, 3, 3), 0.001)
    w[0, 0, 0, 0] = 1.0
    w[1, 0, 0, 0] = 1e-7
    assert oc.dead_percent(w, 1e-5, 0.1, "both") == (1 / 216) * 100.0
    assert oc.dead_percent(torch.full((8,), 1e-6), 1e-5, 0.1, "both") == 0.0
    assert oc.dead_percent(torch.full((8,), 1e-7), 1e-5, 0.1, "both") == 0.0
