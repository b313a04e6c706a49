"""SDXLVAEWrapper — same class, attributes and methods as the reference's
src/models/sdxl_vae_wrapper.py:10-179, with ``self.vae`` a B200AutoencoderKL (libvcd_b200 kernels)
instead of ``diffusers.AutoencoderKL``."""
from __future__ import annotations

import importlib
import logging
from typing import Callable, Dict, List, Optional, Union

import torch

_pkg = importlib.import_module("vae-channel-dynamics_b200")
B200AutoencoderKL = _pkg.B200AutoencoderKL

logger = logging.getLogger(__name__)


class SDXLVAEWrapper(torch.nn.Module):
    def __init__(self, pretrained_model_name_or_path: str = "stabilityai/sdxl-vae",
                 torch_dtype: Optional[Union[str, torch.dtype]] = None):
        super().__init__()
        self.pretrained_model_name_or_path = pretrained_model_name_or_path
        self.torch_dtype = torch_dtype
        self.vae = self._load_vae()
        self.scaling_factor = self.vae.config.scaling_factor      # reference :22
        self._hook_handles: List[torch.utils.hooks.RemovableHandle] = []
        self._captured_activations: Dict[str, torch.Tensor] = {}

    def _load_vae(self):
        """reference :27-40 — the only place that re-raises."""
        logger.info(f"Loading VAE model from: {self.pretrained_model_name_or_path}")
        try:
            dt = self.torch_dtype
            if isinstance(dt, str):
                dt = getattr(torch, dt)
            vae = B200AutoencoderKL.from_pretrained(self.pretrained_model_name_or_path, torch_dtype=dt)
            logger.info(f"VAE model loaded successfully. scaling factor: {vae.config.scaling_factor}")
            return vae
        except Exception as e:
            logger.error(f"Failed to load VAE model from {self.pretrained_model_name_or_path}: {e}")
            raise

    def forward(self, pixel_values: torch.Tensor, sample_posterior: bool = True):
        """reference :42-77 — encode, sample (train) or mode (eval), decode unscaled latents."""
        latent_dist = self.vae.encode(pixel_values).latent_dist
        latents = latent_dist.sample() if sample_posterior else latent_dist.mode()
        reconstruction = self.vae.decode(latents).sample
        return {"reconstruction": reconstruction, "latent_dist": latent_dist, "latents_sampled": latents}

    # ---- activation capture for evaluate.py / logit lens (reference :79-146) ---------------
    def _capture_activation_hook_fn(self, name: str) -> Callable:
        def hook(module, input_data, output_data):
            self._captured_activations[name] = output_data.detach().float().cpu() \
                if output_data.dtype == torch.bfloat16 else output_data.detach().cpu()
        return hook

    def add_hooks(self, layer_names: List[str]):
        self.remove_hooks()
        found = False
        for name, module in self.vae.named_modules():
            if name in layer_names:
                self._hook_handles.append(module.register_forward_hook(self._capture_activation_hook_fn(name)))
                logger.info(f"Registered activation hook for VAE layer: '{name}'")
                found = True
        if not found and layer_names:
            logger.warning(f"No hooks registered. Ensure layer names {layer_names} are correct and exist in the VAE.")

    def remove_hooks(self):
        if not self._hook_handles:
            return
        for h in self._hook_handles:
            h.remove()
        self._hook_handles.clear()
        self._captured_activations.clear()
        logger.info("Cleared all VAE model hooks and captured activations.")

    def get_captured_activations(self) -> Dict[str, torch.Tensor]:
        return self._captured_activations

    def clear_captured_activations(self):
        self._captured_activations.clear()

    # ---- inference helpers (reference :148-179) --------------------------------------------
    @torch.no_grad()
    def encode(self, pixel_values: torch.Tensor) -> torch.Tensor:
        self.vae.eval()
        dist = self.vae.encode(pixel_values.to(self.vae.device, dtype=self.vae.dtype)).latent_dist
        return dist.sample() * self.scaling_factor

    @torch.no_grad()
    def decode(self, latents: torch.Tensor) -> torch.Tensor:
        self.vae.eval()
        latents = latents / self.scaling_factor
        image = self.vae.decode(latents.to(self.vae.device, dtype=self.vae.dtype)).sample
        return image.clamp(-1, 1)
