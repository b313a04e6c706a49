"""TEST SHIM: matplotlib.ticker."""
from unittest.mock import MagicMock


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return MagicMock(name=f"ticker.{name}")
