"""TEST SHIM: matplotlib.pyplot (see matplotlib/__init__.py)."""
from unittest.mock import MagicMock

from . import _Figure, _touch


def subplots(nrows=1, ncols=1, *a, **k):
    fig = _Figure(name="Figure")
    if nrows == 1 and ncols == 1:
        return fig, MagicMock(name="Axes")
    import numpy as np
    axes = np.empty((nrows, ncols), dtype=object)
    for i in range(nrows):
        for j in range(ncols):
            axes[i, j] = MagicMock(name=f"Axes[{i},{j}]")
    if k.get("squeeze", True):
        axes = axes.squeeze()
    return fig, axes


def figure(*a, **k):
    return _Figure(name="Figure")


def savefig(path, *a, **k):
    _touch(path)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return MagicMock(name=f"pyplot.{name}")
