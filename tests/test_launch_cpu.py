"""CPU: the drop-in launcher orders sys.path so that the reference's unchanged scripts import the B200 classes for
models / tracking / classification / intervention and the reference's own files for everything else; the accelerate
test shim exposes the surface train.py / evaluate.py use."""
import os
import subprocess
import sys

import pytest

import ref_harness as rh

ROOT = rh.ROOT


def test_launcher_resolves_dropin_modules_first(tmp_path):
    ref = rh.reference_dir()
    if ref is None:
        pytest.skip("baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists")
    script = tmp_path / "src"
    script.mkdir()
    # a stand-in for train.py living next to a copy of the reference's utils package
    import shutil
    shutil.copytree(os.path.join(ref, "src", "utils"), script / "utils")
    (script / "models").mkdir()
    (script / "models" / "__init__.py").write_text("raise ImportError('the reference models package must be shadowed')\n")
    (script / "probe.py").write_text(
        "import sys\n"
        "import models.sdxl_vae_wrapper as w, tracking.monitor as m, classification.classifier as c, intervention.nudger as n\n"
        "import tracking.deadneuron as d\n"
        "import utils.config_utils as u\n"
        "print('ARGS', sys.argv[1:])\n"
        "for x in (w, m, c, n, d, u): print('MOD', x.__name__, x.__file__)\n")
    r = subprocess.run([sys.executable, rh.LAUNCH, str(script / "probe.py"), "--config_path", "x.yaml"], capture_output=True,
                       text=True, timeout=300, env={**os.environ, "PYTHONPATH": rh.SHIMS})
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l.split() for l in r.stdout.splitlines() if l.startswith("MOD")]
    where = {l[1]: l[2] for l in lines}
    pkg_src = os.path.join(ROOT, "vae-channel-dynamics_b200", "src")
    for mod in ("models.sdxl_vae_wrapper", "tracking.monitor", "classification.classifier", "intervention.nudger", "tracking.deadneuron"):
        assert where[mod].startswith(pkg_src), (mod, where[mod])
    assert where["utils.config_utils"].startswith(str(script)), where
    assert "ARGS ['--config_path', 'x.yaml']" in r.stdout


def test_accelerate_shim_surface():
    sys.path.insert(0, rh.SHIMS)
    try:
        for k in [k for k in sys.modules if k == "accelerate" or k.startswith("accelerate.")]:
            del sys.modules[k]
        import torch
        from accelerate import Accelerator
        from accelerate.logging import get_logger
        from accelerate.utils import ProjectConfiguration, set_seed
        acc = Accelerator(gradient_accumulation_steps=2, mixed_precision="no", log_with=None,
                          project_config=ProjectConfiguration(project_dir="/tmp/x", logging_dir="/tmp/x/logs"), cpu=True)
        assert acc.is_main_process and acc.is_local_main_process and acc.num_processes == 1 and acc.device.type == "cpu"
        get_logger("t", log_level="INFO").info("hello", main_process_only=False)
        set_seed(1)
        model = torch.nn.Linear(4, 2)
        opt = torch.optim.AdamW(model.parameters(), lr=0.1)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (s + 1))
        data = torch.utils.data.DataLoader([{"pixel_values": torch.randn(4)} for _ in range(6)], batch_size=2)
        model, opt, data, sched = acc.prepare(model, opt, data, sched)
        assert acc.unwrap_model(model) is model and len(data) == 3
        w0 = model.weight.detach().clone()
        stepped = []
        for batch in data:
            with acc.accumulate(model):
                loss = model(batch["pixel_values"]).pow(2).mean()
                assert acc.gather(loss.detach()).mean().item() == pytest.approx(loss.item())
                acc.backward(loss)
                if acc.sync_gradients:
                    acc.clip_grad_norm_(model.parameters(), 1.0)
                opt.step()
                sched.step()
                opt.zero_grad(set_to_none=True)
                stepped.append(acc.sync_gradients)
        # accumulation 2 over 3 batches: sync on batch 2 and on the last batch of the loader
        assert stepped == [False, True, True]
        assert not torch.equal(model.weight, w0) and sched.get_last_lr()[0] == pytest.approx(0.1 / 3)
        d = acc.save_state(os.path.join("/tmp", f"vcd_acc_state_{os.getpid()}"))
        assert all(os.path.exists(os.path.join(d, f)) for f in ("model.safetensors", "optimizer.bin", "scheduler.bin",
                                                                 "random_states_0.pkl"))
    finally:
        sys.path.remove(rh.SHIMS)
        for k in [k for k in sys.modules if k == "accelerate" or k.startswith("accelerate.")]:
            del sys.modules[k]
