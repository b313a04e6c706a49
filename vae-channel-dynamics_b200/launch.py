"""Run one of the reference's own entry scripts (src/train.py, src/evaluate.py) UNCHANGED on the B200 drop-in.

    python /path/to/repo/vae-channel-dynamics_b200/launch.py /path/to/reference/src/train.py --config_path cfg.yaml
    torchrun --nproc-per-node 8 /path/to/repo/vae-channel-dynamics_b200/launch.py /path/to/reference/src/train.py ...

`python script.py` puts the script's own directory FIRST on sys.path, so the reference's `from models.sdxl_vae_wrapper
import SDXLVAEWrapper` (train.py:27-31, evaluate.py:26) would find the reference's diffusers-based modules.  This
launcher orders the path as  [<repo>/vae-channel-dynamics_b200/src, <reference>/src, ...]  — `models`, `tracking`,
`classification`, `intervention` resolve to the B200 classes; `utils`, `data_utils`, `analysis` to the reference's
own files — and then executes the script as `__main__` with its arguments untouched.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _extend_package(name: str, ref_src: str) -> None:
    """`tracking`, `models`, ... exist in both trees.  The B200 package wins for the modules it defines; everything
    else in the reference's package of the same name (none today) stays importable through __path__."""
    mod = sys.modules.get(name)
    extra = os.path.join(ref_src, name)
    if mod is not None and os.path.isdir(extra) and extra not in list(getattr(mod, "__path__", [])):
        mod.__path__.append(extra)


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: launch.py <reference script.py> [script arguments]")
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"launch.py: no such script: {script}")
    ref_src = os.path.dirname(script)
    ours = os.path.join(HERE, "src")
    for p in (HERE, ref_src, ours, ROOT):           # drop stale copies, then insert in priority order
        while p in sys.path:
            sys.path.remove(p)
    sys.path[0:0] = [ours, ref_src, ROOT]
    import importlib
    for name in ("models", "tracking", "classification", "intervention"):
        importlib.import_module(name)
        _extend_package(name, ref_src)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
