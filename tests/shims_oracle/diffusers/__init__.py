"""TEST SHIM (reference-arm only): `diffusers.AutoencoderKL` backed by the plain-torch oracle
(oracle/torch_vae.py), so the reference's OWN src/models/sdxl_vae_wrapper.py, monitor, classifier, nudger, dead-neuron
tracker, train.py and evaluate.py run unmodified in an image without diffusers.  The result of such a run is the
expected output the B200 drop-in is compared with (tests/test_reference_callers_gpu.py).  Never on the product path.
"""
import json
import os
from types import SimpleNamespace

import torch

from oracle.torch_vae import SDXL_VAE_CONFIG, ODiagonalGaussian, OracleAutoencoderKL

__version__ = "0.0-vcd-oracle-shim"


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _Dist(ODiagonalGaussian):
    """diffusers samples with randn_tensor(shape, generator, device, dtype) — the same draw as torch.randn on the
    parameters' device in the parameters' dtype."""

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise


class AutoencoderKL(OracleAutoencoderKL):
    config_name = "config.json"
    weights_name = "diffusion_pytorch_model.safetensors"

    def __init__(self, cfg=None):
        super().__init__(None)
        self.config = _Config(dict(SDXL_VAE_CONFIG, _class_name="AutoencoderKL", **(cfg or {})))

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def encode(self, x, return_dict=True):
        return SimpleNamespace(latent_dist=_Dist(self.quant_conv(self.encoder(x))))

    @classmethod
    def from_pretrained(cls, path, torch_dtype=None, **_):
        path = str(path)
        if path.startswith("random-init"):
            torch.manual_seed(int(path.split(":")[1]) if ":" in path else 42)
            m = cls()
        else:
            from safetensors.torch import load_file
            m = cls()
            m.load_state_dict(load_file(os.path.join(path, cls.weights_name)), strict=True)
        return m.to(torch_dtype) if torch_dtype is not None else m

    def save_pretrained(self, save_directory, **_):
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump({k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()}, f, indent=2, sort_keys=True)
        save_file({k: v.detach().contiguous().cpu() for k, v in self.state_dict().items()},
                  os.path.join(save_directory, self.weights_name), metadata={"format": "pt"})
