"""ORACLE (test infrastructure, never the product path).

Plain-PyTorch restatement of the arithmetic the reference reaches through
``diffusers.AutoencoderKL`` with the ``stabilityai/sdxl-vae`` config.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this file.

PARITY UNPINNED at the diffusers boundary: ``diffusers`` is an un-vendored,
unpinned pip dependency of the reference (requirements.txt:8) and is absent from
this image, and the reference holds no golden tensors for the VAE arithmetic
(SURVEY.md section 8c).  What pins this file: the module tree must reproduce the
published SDXL-VAE size (83 653 863 parameters in 248 tensors, checked in
tests/test_oracle.py) and the call sites below; and — an independent implementation
that IS installed — HF transformers' port of the latent-diffusion autoencoder the
SDXL-VAE was trained with and diffusers' AutoencoderKL was ported from
(JanusVQVAEEncoder / JanusVQVAEDecoder, configured with the SDXL-VAE hyper-parameters):
with this file's weights copied in by name, encoder and decoder outputs agree to
1e-5 (tests/test_oracle_crosscheck.py).

Reference call sites restated here:
  * src/models/sdxl_vae_wrapper.py:60   vae.encode(x).latent_dist
  * src/models/sdxl_vae_wrapper.py:64   latent_dist.sample()
  * src/models/sdxl_vae_wrapper.py:66   latent_dist.mode()
  * src/models/sdxl_vae_wrapper.py:71   vae.decode(z).sample
  * src/train.py:289-291                mse + kl_weight * kl().mean()
  * src/train.py:77-78                  validation: mse(sum), kl().sum()
Upstream semantics (diffusers AutoencoderKL / ResnetBlock2D / Attention /
Downsample2D(padding=0) / Upsample2D / DiagonalGaussianDistribution) are written
from the published algorithm: GroupNorm(32, C, eps=1e-6), SiLU, 3x3 convs,
asymmetric (0,1,0,1) pad before the stride-2 downsample conv, nearest x2 before
the upsample conv, single-head attention of width 512 in both mid blocks,
logvar clamp [-30, 20].
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

SDXL_VAE_CONFIG = dict(
    in_channels=3,
    out_channels=3,
    latent_channels=4,
    block_out_channels=(128, 256, 512, 512),
    layers_per_block=2,
    norm_num_groups=32,
    norm_eps=1e-6,
    scaling_factor=0.13025,
)


class OResnet(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class OAttention(nn.Module):
    """Single-head self attention over the (H*W) tokens, residual connection."""

    def __init__(self, c: int, groups: int, eps: float):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, hh, ww = x.shape
        res = x
        t = self.group_norm(x.view(b, c, hh * ww)).transpose(1, 2)  # [B, HW, C]
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        s = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(c), dim=-1)
        o = self.to_out[0](s @ v)
        return o.transpose(1, 2).reshape(b, c, hh, ww) + res


class OMid(nn.Module):
    def __init__(self, c: int, groups: int, eps: float):
        super().__init__()
        self.attentions = nn.ModuleList([OAttention(c, groups, eps)])
        self.resnets = nn.ModuleList([OResnet(c, c, groups, eps), OResnet(c, c, groups, eps)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class ODownsample(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class OUpsample(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class ODownBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, eps, down: bool):
        super().__init__()
        self.resnets = nn.ModuleList([OResnet(cin if i == 0 else cout, cout, groups, eps) for i in range(n)])
        self.downsamplers = nn.ModuleList([ODownsample(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class OUpBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, eps, up: bool):
        super().__init__()
        self.resnets = nn.ModuleList([OResnet(cin if i == 0 else cout, cout, groups, eps) for i in range(n)])
        self.upsamplers = nn.ModuleList([OUpsample(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class OEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, g, eps, n = cfg["block_out_channels"], cfg["norm_num_groups"], cfg["norm_eps"], cfg["layers_per_block"]
        self.conv_in = nn.Conv2d(cfg["in_channels"], ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            self.down_blocks.append(ODownBlock(cin, cout, n, g, eps, down=i < len(ch) - 1))
        self.mid_block = OMid(ch[-1], g, eps)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg["latent_channels"], 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class ODecoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, g, eps, n = cfg["block_out_channels"], cfg["norm_num_groups"], cfg["norm_eps"], cfg["layers_per_block"]
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(cfg["latent_channels"], rev[0], 3, padding=1)
        self.up_blocks = nn.ModuleList()
        self.mid_block = OMid(rev[0], g, eps)
        cout = rev[0]
        for i, c in enumerate(rev):
            cin, cout = cout, c
            self.up_blocks.append(OUpBlock(cin, cout, n + 1, g, eps, up=i < len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg["out_channels"], 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class ODiagonalGaussian:
    """[upstream] DiagonalGaussianDistribution: chunk, clamp(-30, 20), exp."""

    def __init__(self, moments: torch.Tensor):
        self.parameters = moments
        self.mean, self.logvar = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean

    def kl(self):
        return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])


class OracleAutoencoderKL(nn.Module):
    def __init__(self, cfg: Optional[dict] = None):
        super().__init__()
        cfg = dict(SDXL_VAE_CONFIG if cfg is None else cfg)
        self.config = SimpleNamespace(**cfg)
        self.encoder = OEncoder(cfg)
        self.decoder = ODecoder(cfg)
        lc = cfg["latent_channels"]
        self.quant_conv = nn.Conv2d(2 * lc, 2 * lc, 1)
        self.post_quant_conv = nn.Conv2d(lc, lc, 1)

    def encode(self, x):
        return SimpleNamespace(latent_dist=ODiagonalGaussian(self.quant_conv(self.encoder(x))))

    def decode(self, z):
        return SimpleNamespace(sample=self.decoder(self.post_quant_conv(z)))


def oracle_forward(vae: OracleAutoencoderKL, x: torch.Tensor, sample_posterior: bool = True,
                   noise: Optional[torch.Tensor] = None):
    """src/models/sdxl_vae_wrapper.py:42-77 restated."""
    dist = vae.encode(x).latent_dist
    z = dist.sample(noise=noise) if sample_posterior else dist.mode()
    rec = vae.decode(z).sample
    return {"reconstruction": rec, "latent_dist": dist, "latents_sampled": z}


def oracle_losses(out: dict, x: torch.Tensor, kl_weight: float):
    """src/train.py:289-291 restated."""
    rec = F.mse_loss(out["reconstruction"].float(), x.float(), reduction="mean")
    kl = out["latent_dist"].kl().mean()
    return rec + kl_weight * kl, rec, kl


def build_oracle(seed: int = 42, cfg: Optional[dict] = None, dtype=torch.float32) -> OracleAutoencoderKL:
    """Random-init SDXL-VAE architecture (torch default inits), seeded like train.py:131."""
    torch.manual_seed(seed)
    return OracleAutoencoderKL(cfg).to(dtype)
