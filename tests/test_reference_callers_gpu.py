"""north_star: "train.py, evaluate.py and the YAML configs run unchanged".

Executes the reference's OWN src/train.py and src/evaluate.py (unmodified files staged under baseline/_ref by
__graft_entry__.build()) twice on the GPU — once on the B200 drop-in (vae-channel-dynamics_b200/launch.py), once with every
module the reference's own and diffusers.AutoencoderKL served by the plain-torch oracle — from the shipped YAML
configs (offline overrides only: local synthetic dataset, a local init checkpoint with planted dead channels, fewer
samples / epochs) and compares what the two runs WRITE: tracked_activation_stats.csv, intervention_history.csv,
dead_neuron_percentage_history.csv, the saved final_model/vae, and evaluate.py's eval_metrics.txt.
Covers SURVEY rows a7 (the three gathers + .item()), a9 (AdamW + LambdaLR through accelerate.prepare), a17 (cadence),
(b) the boundary, (f3) save_pretrained/from_pretrained round trip, (f4) run_validation + PSNR/SSIM.
"""
import os

import numpy as np
import pandas as pd
import pytest
import torch

import ref_harness as rh
from util import record_parity

pytestmark = pytest.mark.gpu

PLANTED = ["encoder.down_blocks.0.resnets.0.norm1", "decoder.up_blocks.1.resnets.0.norm1"]


def _prepare_workdir(vcd, wd, n_train, n_test):
    vcd.data.write_synthetic_image_dataset(wd, "uoft-cs/cifar10", {"train": n_train, "test": n_test}, size=32,
                                           image_column="img", kind="smooth", seed=3)
    vae = vcd.B200AutoencoderKL.from_pretrained("random-init:42")
    with torch.no_grad():
        for n in PLANTED:                      # dead channels the classifier must find
            vae.get_submodule(n).weight[::8] = 1e-3
    init = os.path.join(wd, "init_vae")
    vae.save_pretrained(init)
    return init


def _read_metrics(path):
    out = {}
    with open(path) as f:
        for line in f:
            if ":" in line:
                k, v = line.split(":", 1)
                try:
                    out[k.strip()] = float(v)
                except ValueError:
                    out[k.strip()] = v.strip()
    return out


def _compare_runs(ours, ref, tag, stat_tol, weight_tol):
    res = {}
    # ---- tracked_activation_stats.csv: same records in the same order, values at the network tolerance
    a = pd.read_csv(os.path.join(ours, "tracked_activation_stats.csv"))
    b = pd.read_csv(os.path.join(ref, "tracked_activation_stats.csv"))
    key = ["global_step", "layer_identifier", "original_metric_name", "metric_type"]
    ka, kb = [tuple(r) for r in a[key].values.tolist()], [tuple(r) for r in b[key].values.tolist()]
    missing_in_ref = sorted({k[1] for k in ka} - {k[1] for k in kb})
    res["layers_only_in_b200_run"] = missing_in_ref      # bf16: the reference's conv-output hook dies in .numpy() (monitor.py:78)
    assert [k for k in ka if k[1] not in missing_in_ref] == kb, "record order / set differs"
    m = a.merge(b, on=key, suffixes=("_b200", "_ref"))
    worst = 0.0
    for _, r in m.iterrows():
        if r["metric_type"] == "full_map_shape":
            assert r["metric_value_b200"] == r["metric_value_ref"], r
            continue
        va, vb = float(r["metric_value_b200"]), float(r["metric_value_ref"])
        scale = max(abs(vb), 0.05)
        tol = stat_tol * (4 if r["metric_type"] in ("full_map_min", "full_map_max") else 1)
        err = abs(va - vb) / scale
        worst = max(worst, err / (4 if r["metric_type"] in ("full_map_min", "full_map_max") else 1))
        assert err < tol, (dict(r), err)
    res["tracked_stats_rows"] = len(m)
    res["tracked_stats_worst_rel_err"] = worst
    # ---- intervention_history.csv: step, #inactive, #nudged — exact
    ia = open(os.path.join(ours, "intervention_history.csv")).read()
    ib = open(os.path.join(ref, "intervention_history.csv")).read()
    assert ia == ib and ia.strip(), (ia, ib)
    res["intervention_history"] = ia.strip().splitlines()
    # ---- dead_neuron_percentage_history.csv
    da = pd.read_csv(os.path.join(ours, "dead_neuron_percentage_history.csv"))
    db = pd.read_csv(os.path.join(ref, "dead_neuron_percentage_history.csv"))
    assert da[["step", "layer"]].values.tolist() == db[["step", "layer"]].values.tolist()
    dd = np.abs(da["percentage"].values - db["percentage"].values)
    res["dead_weight_rows"] = len(da)
    res["dead_weight_max_abs_diff_percentage_points"] = float(dd.max())
    # ---- final_model/vae (save_pretrained) — same tensors, close values
    from safetensors.torch import load_file
    sa = load_file(os.path.join(ours, "final_model", "vae", "diffusion_pytorch_model.safetensors"))
    sb = load_file(os.path.join(ref, "final_model", "vae", "diffusion_pytorch_model.safetensors"))
    assert list(sa) == list(sb) and len(sa) == 248
    res["final_weights_max_abs_diff"] = max(float((sa[k].float() - sb[k].float()).abs().max()) for k in sa)
    for n in PLANTED:
        ga, gb = sa[n + ".weight"].float(), sb[n + ".weight"].float()
        res[f"planted gamma max abs diff {n}"] = float((ga[::8] - gb[::8]).abs().max())
        res[f"planted gamma mean {n}"] = {"b200": float(ga[::8].mean()), "oracle": float(gb[::8].mean())}
    assert res["final_weights_max_abs_diff"] < weight_tol, res["final_weights_max_abs_diff"]
    assert os.path.isfile(os.path.join(ours, "final_model", "model.safetensors"))      # accelerator.save_state layout
    record_parity(tag, res)
    return res


def _run_both(vcd, tmp_path, base_yaml, overrides, tag, stat_tol, eval_tol, weight_tol, env_b200=None, evaluate=True):
    if rh.reference_dir() is None:
        pytest.skip("baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists")
    wd = str(tmp_path)
    init = _prepare_workdir(vcd, wd, n_train=overrides["data"]["max_samples"], n_test=overrides["data"]["validation_max_samples"])
    runs = {}
    for arm in ("b200", "oracle"):
        ov = {k: (dict(v) if isinstance(v, dict) else v) for k, v in overrides.items()}
        ov["output_dir"] = os.path.join(wd, "results_" + arm)
        ov["model"] = {"pretrained_vae_name": init}
        cfg = rh.make_config(wd, base_yaml, ov, f"{arm}.yaml")
        log = rh.run_script(arm, "train.py", ["--config_path", cfg], wd, env_extra=env_b200 if arm == "b200" else None)
        assert "Training finished." in log
        import yaml
        run_name = yaml.safe_load(open(cfg))["run_name"]
        runs[arm] = (cfg, os.path.join(ov["output_dir"], run_name), log)
    res = _compare_runs(runs["b200"][1], runs["oracle"][1], tag, stat_tol, weight_tol)
    if not evaluate:
        return res
    # ---- evaluate.py on each run's own final_model (from_pretrained of what save_pretrained wrote)
    ev = {}
    for arm in ("b200", "oracle"):
        cfg, out, _ = runs[arm]
        rh.run_script(arm, "evaluate.py", ["--config_path", cfg, "--checkpoint_path", os.path.join(out, "final_model"),
                                           "--eval_split", "test", "--num_samples_to_save", "2"], wd)
        ev[arm] = _read_metrics(os.path.join(out, "final_model", "eval_results_test", "eval_metrics.txt"))
        assert os.path.isfile(os.path.join(out, "final_model", "eval_results_test", "sample_0_recon.png"))
    e = {k: {"b200": ev["b200"][k], "oracle": ev["oracle"][k]} for k in ("Average MSE", "Average KL", "Average PSNR", "Average SSIM")}
    record_parity(tag + " evaluate.py", e)
    assert ev["b200"]["Number of Samples Processed"] == ev["oracle"]["Number of Samples Processed"]
    for k in ("Average MSE", "Average KL"):
        assert abs(e[k]["b200"] - e[k]["oracle"]) < eval_tol * abs(e[k]["oracle"]), (k, e[k])
    assert abs(e["Average PSNR"]["b200"] - e["Average PSNR"]["oracle"]) < 0.2, e
    assert abs(e["Average SSIM"]["b200"] - e["Average SSIM"]["oracle"]) < 0.01, e
    return res


def test_unchanged_train_py_and_evaluate_py_cifar10_test_config(vcd, tmp_path):
    """configs[0] experiment_cifar10_test.yaml (fp32 parameters, tracking incl. full_activation_map hooks, dead-weight
    tracker, classification, nudge, validation every epoch, periodic save_state): 48 optimizer steps."""
    res = _run_both(vcd, tmp_path, "experiment_cifar10_test.yaml", {
        "data": {"max_samples": 48, "validation_max_samples": 16},
        "training": {"num_train_epochs": 8},
    }, "train.py unchanged: experiment_cifar10_test.yaml 48 steps", stat_tol=3e-2, eval_tol=3e-2,
        weight_tol=1.5e-3)   # sum of the warm-up learning rates over 48 steps is 5.9e-4: a sign flip moves a weight by 2 lr
    assert len(res["intervention_history"]) == 2          # steps 20 and 40: lcm(track_interval 10, intervention_interval 20)
    assert res["intervention_history"][0].split(",") == ["20", "80", "80"]


def test_unchanged_train_py_bf16_nudge_config_with_fused_optimizer(vcd, tmp_path):
    """configs[1] experiment_cifar10_nudge.yaml with mixed_precision bf16 (BASELINE.json), batch 16 x 40 steps; the
    drop-in run additionally opts into the fused clip+AdamW (VCD_FUSED_OPT=1 in accelerate.prepare)."""
    res = _run_both(vcd, tmp_path, "experiment_cifar10_nudge.yaml", {
        "data": {"max_samples": 128, "validation_max_samples": 16, "batch_size": 16, "validation_batch_size": 16, "num_workers": 0},
        "training": {"num_train_epochs": 5, "mixed_precision": "bf16"},
    }, "train.py unchanged: experiment_cifar10_nudge.yaml bf16 40 steps (fused clip+AdamW)", stat_tol=4e-2, eval_tol=5e-2,
        weight_tol=5e-3, env_b200={"VCD_FUSED_OPT": "1"},
        # the reference's evaluate.py cannot run a bf16 config in EITHER arm: evaluate.py:97 does getattr(torch, "bf16")
        evaluate=False)
    assert [r.split(",")[0] for r in res["intervention_history"]] == ["20", "40"]


def test_unchanged_train_py_on_two_ranks(vcd, tmp_path):
    """torchrun --nproc-per-node 2 launch.py train.py (needs 2 GPUs): the accelerate surface at world size 2 — DDP wrap,
    per-rank batch sharding, the three loss all-gathers of train.py:292-294, rank-0-only classifier / nudger / CSV
    writers — with the drop-in's global-batch statistics and gamma re-broadcast.  The run must finish, intervene on the
    planted channels at steps 20 and 40, and save a model whose nudged scales moved."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    if rh.reference_dir() is None:
        pytest.skip("baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists")
    wd = str(tmp_path)
    init = _prepare_workdir(vcd, wd, n_train=96, n_test=16)
    cfg = rh.make_config(wd, "experiment_cifar10_test.yaml", {
        "output_dir": os.path.join(wd, "results_b200_2gpu"), "model": {"pretrained_vae_name": init},
        "data": {"max_samples": 96, "validation_max_samples": 16},           # 96 / (8 per rank x 2 ranks) = 6 steps per epoch
        "training": {"num_train_epochs": 8}}, "two_ranks.yaml")
    log = rh.run_script("b200", "train.py", ["--config_path", cfg], wd, nproc=2)
    assert "Training finished." in log
    out = os.path.join(wd, "results_b200_2gpu", "sdxl_vae_cifar10_test_run")
    hist = open(os.path.join(out, "intervention_history.csv")).read().strip().splitlines()
    assert hist == ["20,80,80", "40,80,80"], hist
    df = pd.read_csv(os.path.join(out, "tracked_activation_stats.csv"))
    assert sorted(df["global_step"].unique().tolist()) == [10, 20, 30, 40]
    from safetensors.torch import load_file
    sd = load_file(os.path.join(out, "final_model", "vae", "diffusion_pytorch_model.safetensors"))
    for n in PLANTED:
        g = sd[n + ".weight"].float()[::8]
        assert float(g.min()) > 1.0e-3 * 1.05 * 1.05 * 0.5 and float(g.max()) < 5e-3, (n, float(g.min()), float(g.max()))
    record_parity("train.py unchanged on 2 ranks (torchrun + launch.py): experiment_cifar10_test.yaml 48 steps",
                  {"intervention_history": hist, "tracked_rows": int(len(df))})
