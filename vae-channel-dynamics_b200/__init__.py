"""vae-channel-dynamics_b200 — B200-native hot path of olegroshka/vae-channel-dynamics.

Import with ``importlib.import_module("vae-channel-dynamics_b200")`` or, shorter, ``import vcd_b200``
(alias module at the repository root).  ``src/`` mirrors the reference's own ``src/`` package layout
(models / tracking / classification / intervention) so the reference's ``train.py`` / ``evaluate.py``
import these classes unchanged when ``vae-channel-dynamics_b200/src`` is first on ``sys.path``.
"""
import os
import sys

from . import _lib, ops  # noqa: F401
from ._lib import LIB_PATH, VcdError  # noqa: F401
from .vae import B200AutoencoderKL, DiagonalGaussianDistribution  # noqa: F401
from .losses import vae_loss  # noqa: F401
from .graph_step import GraphedVAEStep  # noqa: F401
from .optim import FusedClipAdamW  # noqa: F401
from . import data, metrics  # noqa: F401

SRC_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "src")


def add_src_to_path() -> str:
    """Make ``models.sdxl_vae_wrapper`` etc. importable exactly as the reference's train.py imports them."""
    if SRC_DIR not in sys.path:
        sys.path.insert(0, SRC_DIR)
    return SRC_DIR
