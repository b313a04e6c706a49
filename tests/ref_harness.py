"""Harness that executes the reference's OWN, UNMODIFIED entry scripts (baseline/_ref/src/train.py, evaluate.py —
staged from /root/reference by __graft_entry__.build(), or $VCD_REFERENCE) in a subprocess, in one of two arms:

  arm "b200"   : through vae-channel-dynamics_b200/launch.py — models / tracking / classification / intervention resolve to the
                 B200 drop-in, everything else (utils, data_utils, analysis, the script itself) is the reference's;
  arm "oracle" : plain `python train.py` — every module is the reference's own; its `from diffusers import
                 AutoencoderKL` is served by tests/shims_oracle (the plain-torch oracle).  This is the expected output.

Both arms get tests/shims on PYTHONPATH (accelerate / matplotlib / torchmetrics are absent from the image), W&B in
disabled mode, and an offline HF datasets tree written by vcd_b200.data.write_synthetic_image_dataset.
"""
import os
import subprocess
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "tests", "shims")
SHIMS_ORACLE = os.path.join(ROOT, "tests", "shims_oracle")
LAUNCH = os.path.join(ROOT, "vae-channel-dynamics_b200", "launch.py")


def reference_dir():
    for d in (os.environ.get("VCD_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if d and os.path.isfile(os.path.join(d, "src", "train.py")):
            return d
    return None


def make_config(workdir: str, base_yaml: str, overrides: dict, name: str) -> str:
    """Shipped experiment YAML + its base_config (the reference's `defaults:` inheritance, utils/config_utils.py:37-52),
    with the offline overrides merged section by section."""
    ref = reference_dir()
    with open(os.path.join(ref, "configs", base_yaml)) as f:
        cfg = yaml.safe_load(f)
    for k, v in overrides.items():
        if isinstance(v, dict) and isinstance(cfg.get(k), dict):
            cfg[k].update(v)
        else:
            cfg[k] = v
    cdir = os.path.join(workdir, "configs")
    os.makedirs(cdir, exist_ok=True)
    import shutil
    shutil.copy(os.path.join(ref, "configs", "base_config.yaml"), os.path.join(cdir, "base_config.yaml"))
    path = os.path.join(cdir, name)
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)
    return path


def run_script(arm: str, script: str, args, workdir: str, env_extra=None, timeout=1500, nproc: int = 1):
    ref = reference_dir()
    spath = os.path.join(ref, "src", script)
    env = dict(os.environ)
    env.update({"WANDB_MODE": "disabled", "WANDB_SILENT": "true", "HF_DATASETS_OFFLINE": "1", "HF_HUB_OFFLINE": "1",
                "HF_HOME": os.path.join(workdir, "hf_home"), "TQDM_DISABLE": "1", "TOKENIZERS_PARALLELISM": "false",
                "CUBLAS_WORKSPACE_CONFIG": ":4096:8"})
    paths = [SHIMS] + ([SHIMS_ORACLE, ROOT] if arm == "oracle" else [ROOT])
    env["PYTHONPATH"] = os.pathsep.join(paths + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env.update(env_extra or {})
    py = [sys.executable]
    if nproc > 1:
        py += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
               "--master-port", str(29400 + os.getpid() % 500)]
    cmd = py + ([LAUNCH, spath] if arm == "b200" else [spath]) + list(args)
    r = subprocess.run(cmd, cwd=workdir, env=env, capture_output=True, text=True, timeout=timeout)
    log = os.path.join(workdir, f"{arm}_{script}.log")
    with open(log, "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr)
    if r.returncode != 0:
        raise AssertionError(f"{arm} {script} exited {r.returncode}\n{(r.stdout + r.stderr)[-6000:]}")
    return r.stdout + r.stderr
