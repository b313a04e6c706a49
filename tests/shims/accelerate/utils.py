"""TEST SHIM: accelerate.utils.ProjectConfiguration / set_seed (train.py:14,112,131)."""
import random
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass
class ProjectConfiguration:
    project_dir: Optional[str] = None
    logging_dir: Optional[str] = None
    automatic_checkpoint_naming: bool = False
    total_limit: Optional[int] = None
    iteration: int = 0
    save_on_each_node: bool = False

    def __post_init__(self):
        if self.logging_dir is None:
            self.logging_dir = self.project_dir


def set_seed(seed: int, device_specific: bool = False, deterministic: bool = False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
