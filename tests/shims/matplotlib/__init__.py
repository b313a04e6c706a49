"""TEST SHIM: matplotlib is absent from this image; the reference's plotting code (utils/plotting_utils.py,
analysis/logit_lens.py — out of scope) only needs its calls not to fail."""
from unittest.mock import MagicMock

__version__ = "0.0-vcd-test-shim"


class _Figure(MagicMock):
    def savefig(self, path, *a, **k):
        _touch(path)


def _touch(path):
    try:
        with open(path, "wb") as f:
            f.write(b"matplotlib test shim: no image rendered\n")
    except Exception:
        pass


def use(*a, **k):
    pass


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return MagicMock(name=f"matplotlib.{name}")
