"""Importable alias for the package directory ``vae-channel-dynamics_b200`` (its name is not a Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("vae-channel-dynamics_b200")
sys.modules[__name__] = _pkg
