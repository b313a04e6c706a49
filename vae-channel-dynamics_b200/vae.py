"""B200-native drop-in for ``diffusers.AutoencoderKL`` (SDXL-VAE config).

Same module paths, parameter names and shapes as the model the reference loads at
src/models/sdxl_vae_wrapper.py:31, same call surface
(``encode(x).latent_dist.{sample,mode,kl}``, ``decode(z).sample``, ``config.scaling_factor``,
``save_pretrained`` / ``from_pretrained``), but every layer runs a libvcd_b200 kernel.

Layers are real ``nn.Conv2d`` / ``nn.GroupNorm`` / ``nn.Linear`` subclasses so the reference's
``isinstance`` checks (classifier.py:56, train.py:38, deadneuron.py:62) and its dotted-path parameter
lookups (nudger.py:49-72) keep working.  Modules exchange *logically NCHW* bf16 tensors stored
channels-last, so forward hooks registered by foreign code (sdxl_vae_wrapper.py:91-113, evaluate.py:209)
still see [N, C, H, W] tensors; when a layer carries no foreign hook the block uses the fused
GroupNorm+SiLU / conv+residual kernels instead.
"""
from __future__ import annotations

import contextlib
import json
import logging
import os
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import VcdError

SDXL_VAE_CONFIG = {
    "_class_name": "AutoencoderKL",
    "act_fn": "silu",
    "block_out_channels": [128, 256, 512, 512],
    "down_block_types": ["DownEncoderBlock2D"] * 4,
    "up_block_types": ["UpDecoderBlock2D"] * 4,
    "in_channels": 3,
    "out_channels": 3,
    "latent_channels": 4,
    "layers_per_block": 2,
    "norm_num_groups": 32,
    "sample_size": 1024,
    "scaling_factor": 0.13025,
    "force_upcast": True,
}
_NORM_EPS = 1e-6


def _device_of(t: torch.Tensor):
    """Context manager that makes `t`'s CUDA device current (a no-op context for CPU tensors, whose first kernel call raises)."""
    return torch.cuda.device(t.device) if t.is_cuda else contextlib.nullcontext()


def _phys(x: torch.Tensor) -> torch.Tensor:
    """logical [N, C, *S] -> physical [N, *S, C] view."""
    return x.permute(0, *range(2, x.dim()), 1)


def _logi(y: torch.Tensor) -> torch.Tensor:
    """physical [N, *S, C] -> logical [N, C, *S] view."""
    return y.permute(0, y.dim() - 1, *range(1, y.dim() - 1))


def _hooked(m: nn.Module) -> bool:
    return bool(m._forward_hooks) or bool(m._forward_pre_hooks)


class B200Conv2d(nn.Conv2d):
    """3x3 / 1x1 convolution.  Implicit-GEMM tcgen05 kernel when Cin, Cout are multiples of 128,
    CUDA-core direct kernel for the six small-channel layers."""

    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__(cin, cout, k, stride=stride, padding=padding)
        self._packs = ops.PackedWeights()
        self._track_out: Optional[ops.TrackSlot] = None
        self._gn_groups = 0   # > 0: the output feeds a GroupNorm of that many groups (sums fused into the epilogue)
        # the GroupNorm whose INPUT is exactly this conv's output (conv_in -> first norm1, conv1 -> norm2): a statistics
        # slot on this conv's output is then filled by that GroupNorm's own pass over the tensor (no extra read)
        self._stats_consumer = None

    def forward(self, x, residual=None):
        xp = ops.to_nhwc(x)
        N, H, W, _ = xp.shape
        s = self.stride[0]
        if s == 1:
            out_hw = (H, W)
        else:  # Downsample2D: pad (0,1,0,1) then stride 2, no padding
            out_hw = (H // 2, W // 2)
        y = ops.conv2d(xp, self.weight, self.bias, self._packs, stride=s, pad_t=self.padding[0], pad_l=self.padding[1],
                       out_hw=out_hw, residual=None if residual is None else _phys(residual),
                       gn_groups=0 if _hooked(self) else self._gn_groups)
        out = _logi(y)
        if self._track_out is not None:
            cons = self._stats_consumer[0] if self._stats_consumer is not None else None
            if cons is not None and residual is None and not _hooked(self):
                # the reference's buffer is keyed in first-fire order (monitor.py:101): this target fires HERE, even though
                # its statistics are produced a moment later by the consuming GroupNorm's pass over the same tensor
                if self._track_out.on_finalize is not None:
                    self._track_out.on_finalize()
                cons._pending_in = (self._track_out, y.data_ptr())
            else:
                ops.chan_stats(out.detach(), self._track_out)
        return out


class B200GroupNorm(nn.GroupNorm):
    """GroupNorm(32, C, eps=1e-6) with optional fused SiLU and fused per-channel statistics."""

    def __init__(self, groups, channels, eps=_NORM_EPS):
        super().__init__(groups, channels, eps=eps, affine=True)
        self._track_in: Optional[ops.TrackSlot] = None
        self._track_out: Optional[ops.TrackSlot] = None

    def forward(self, x, act: bool = False, split: bool = False, feeds_conv_only: bool = False):
        xp = _phys(x)
        extra, pend = None, self.__dict__.pop("_pending_in", None)
        if pend is not None:
            if pend[1] == xp.data_ptr():          # the producing conv's output really is this input tensor
                extra = pend[0]
            else:   # cannot happen through this module tree; the reference's convention is log + skip, never raise into the loop
                logging.getLogger(__name__).error("statistics hand-off: the tracked conv output is not the input of its GroupNorm; "
                                                  "this forward's statistics of that target are dropped")
        y = ops.group_norm(xp, self.weight, self.bias, self.num_groups, self.eps, act,
                           self._track_in, self._track_out, split, feeds_conv_only, slot_in_extra=extra)
        if split:   # (normalised, input routed through for the block's skip connection)
            return _logi(y[0]), _logi(y[1])
        return _logi(y)


class B200Linear(nn.Linear):
    """Linear over the channel dim of a [N, T, C] token tensor (tcgen05 GEMM, K-major weights as stored)."""

    def __init__(self, cin, cout):
        super().__init__(cin, cout)
        self._packs = ops.PackedWeights()
        self._gn_groups = 0

    def forward(self, x, residual=None):
        # x: physical [N, T, C]
        N, T, C = x.shape
        y = ops.conv2d(x.reshape(N, T, 1, C), self.weight, self.bias, self._packs, stride=1, pad_t=0, pad_l=0,
                       out_hw=(T, 1), residual=None if residual is None else residual.reshape(N, T, 1, -1),
                       gn_groups=self._gn_groups)
        return y.reshape(N, T, -1)


def _norm_act(norm: B200GroupNorm, x, split: bool = False, conv: Optional[nn.Module] = None):
    """GroupNorm followed by SiLU; unfused only when a foreign hook must observe the pre-activation.
    split=True additionally returns the tensor the block's skip connection must use (see ops._GroupNormFn).
    conv: the conv that is the ONLY consumer of the result (its dgrad then carries this GroupNorm's backward
    reduction, ops._GN_FWD); not used when either module carries a foreign hook."""
    if _hooked(norm):
        h = _logi(ops.silu(_phys(norm(x))))
        return (h, x) if split else h
    return norm(x, act=True, split=split, feeds_conv_only=conv is not None and not _hooked(conv))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = B200GroupNorm(groups, cin)
        self.conv1 = B200Conv2d(cin, cout, 3, padding=1)
        self.norm2 = B200GroupNorm(groups, cout)
        self.conv2 = B200Conv2d(cout, cout, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = B200Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h, x = _norm_act(self.norm1, x, split=True, conv=self.conv1)
        h = self.conv1(h)
        h = _norm_act(self.norm2, h, conv=self.conv2)
        sc = x if self.conv_shortcut is None else self.conv_shortcut(x)
        if _hooked(self.conv2) or self.conv2._track_out is not None:
            # a hook / statistics slot on conv2 must observe the PRE-residual tensor (the reference adds the skip
            # connection outside the module): explicit add instead of the residual epilogue
            return _logi(ops.add(_phys(self.conv2(h)), _phys(sc)))
        return self.conv2(h, residual=sc)


class Attention(nn.Module):
    """mid_block.attentions.0: GroupNorm -> q,k,v -> softmax(qk^T/sqrt(C)) v -> out proj -> + residual."""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = B200GroupNorm(groups, c)
        self.to_q = B200Linear(c, c)
        self.to_k = B200Linear(c, c)
        self.to_v = B200Linear(c, c)
        self.to_out = nn.ModuleList([B200Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        N, C, H, W = x.shape
        tokens = _phys(x).reshape(N, H * W, C)                      # physical [N, T, C]
        if _hooked(self.group_norm):
            h = _phys(self.group_norm(_logi(tokens)))               # GroupNorm sees logical [N, C, T]
        else:
            h, tokens = (_phys(t) for t in self.group_norm(_logi(tokens), split=True))
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        o = ops.attention_core(q, k, v)
        if _hooked(self.to_out[0]):   # the hook sees the projection alone; the residual is added outside the module
            o = ops.add(self.to_out[0](o), tokens)
        else:
            o = self.to_out[0](o, residual=tokens)
        return _logi(o.reshape(N, H, W, C))


class UNetMidBlock2D(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(c, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, groups), ResnetBlock2D(c, c, groups)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = B200Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        if x.shape[2] % 2 or x.shape[3] % 2:
            raise VcdError("Downsample2D needs even spatial sizes")
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = B200Conv2d(c, c, 3, padding=1)

        self._up_packs = ops.UpconvPackedWeights()

    def forward(self, x):
        c = self.conv
        if _hooked(c) or c._track_out is not None or not ops.upconv_supported(c.in_channels, c.out_channels):
            # a hook on the conv must observe the upsampled tensor: materialise it
            return c(_logi(ops.upsample2x(_phys(x))))
        return _logi(ops.upconv2d(_phys(x), c.weight, c.bias, self._up_packs, c._gn_groups))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, n, groups, down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(n)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, n, groups, up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(n)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, g, n = cfg["block_out_channels"], cfg["norm_num_groups"], cfg["layers_per_block"]
        self.conv_in = B200Conv2d(cfg["in_channels"], ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            self.down_blocks.append(DownEncoderBlock2D(cin, cout, n, g, down=i < len(ch) - 1))
        self.mid_block = UNetMidBlock2D(ch[-1], g)
        self.conv_norm_out = B200GroupNorm(g, ch[-1])
        self.conv_act = nn.SiLU()
        self.conv_out = B200Conv2d(ch[-1], 2 * cfg["latent_channels"], 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(_norm_act(self.conv_norm_out, x))


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, g, n = cfg["block_out_channels"], cfg["norm_num_groups"], cfg["layers_per_block"]
        rev = list(reversed(ch))
        self.conv_in = B200Conv2d(cfg["latent_channels"], rev[0], 3, padding=1)
        self.up_blocks = nn.ModuleList()
        self.mid_block = UNetMidBlock2D(rev[0], g)
        cout = rev[0]
        for i, c in enumerate(rev):
            cin, cout = cout, c
            self.up_blocks.append(UpDecoderBlock2D(cin, cout, n + 1, g, up=i < len(ch) - 1))
        self.conv_norm_out = B200GroupNorm(g, ch[0])
        self.conv_act = nn.SiLU()
        self.conv_out = B200Conv2d(ch[0], cfg["out_channels"], 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(_norm_act(self.conv_norm_out, x))


class DiagonalGaussianDistribution:
    """[upstream] diffusers DiagonalGaussianDistribution on top of the fused sample+KL kernel.

    ``moments``: logical [N, 2L, h, w] bf16.  sample()/mode() return logical [N, L, h, w] bf16 (channels-last);
    kl() returns fp32 [N].  mean/logvar/std/var are fp32 [N, L, h, w] (logvar clamped to [-30, 20])."""

    def __init__(self, moments: torch.Tensor):
        self.parameters = moments
        self._res = {}

    def _run(self, noise):
        key = "mode" if noise is None else "sample"
        z, kl, mean, logvar = ops.gauss_sample_kl(_phys(self.parameters), noise)
        self._res[key] = (z, kl, mean, logvar)
        self._res["any"] = (z, kl, mean, logvar)
        return z, kl, mean, logvar

    def _any(self):
        return self._res["any"] if "any" in self._res else self._run(None)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        N, C2, h, w = self.parameters.shape
        noise = torch.randn((N, C2 // 2, h, w), generator=generator, device=self.parameters.device,
                            dtype=torch.float32)
        return _logi(self._run(noise)[0])

    def mode(self) -> torch.Tensor:
        res = self._res.get("mode") or self._run(None)
        return _logi(res[0])

    def kl(self, other=None) -> torch.Tensor:
        if other is not None:
            raise VcdError("kl(other) is not on the reference's path (train.py:290 calls kl())")
        return self._any()[1]

    @property
    def mean(self):
        return self._any()[2]

    @property
    def logvar(self):
        return self._any()[3]

    @property
    def std(self):
        return torch.exp(0.5 * self.logvar)

    @property
    def var(self):
        return torch.exp(self.logvar)


class _Config(dict):
    """dict with attribute access (``vae.config.scaling_factor``, sdxl_vae_wrapper.py:22)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class B200AutoencoderKL(nn.Module):
    config_name = "config.json"
    weights_name = "diffusion_pytorch_model.safetensors"

    def __init__(self, config: Optional[dict] = None):
        super().__init__()
        cfg = dict(SDXL_VAE_CONFIG)
        if config:
            cfg.update({k: v for k, v in config.items() if not k.startswith("_") or k == "_class_name"})
        self.config = _Config(cfg)
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        lc = cfg["latent_channels"]
        self.quant_conv = B200Conv2d(2 * lc, 2 * lc, 1)
        self.post_quant_conv = B200Conv2d(lc, lc, 1)
        self.output_dtype = torch.float32  # accelerate converts forward outputs to fp32 [upstream]
        self._mark_groupnorm_producers(cfg["norm_num_groups"])

    def _mark_groupnorm_producers(self, g: int):
        """Every conv / linear whose output tensor is the input of a GroupNorm emits that GroupNorm's sums from its
        GEMM epilogue (ops.conv2d gn_groups).  Block outputs feed the next block's norm1, the attention's group_norm
        or conv_norm_out — except the last resnet of a block that ends in a Down/Upsample2D conv."""
        def block(resnets, sampler):
            for i, r in enumerate(resnets):
                r.conv1._gn_groups = g
                r.conv1._stats_consumer = (r.norm2,)      # tuple: not registered as a child module of the conv
                if sampler is None or i < len(resnets) - 1:
                    r.conv2._gn_groups = g
            if sampler is not None:
                sampler.conv._gn_groups = g
        self.encoder.conv_in._stats_consumer = (self.encoder.down_blocks[0].resnets[0].norm1,)
        self.decoder.conv_in._stats_consumer = (self.decoder.mid_block.resnets[0].norm1,)
        for net in (self.encoder, self.decoder):
            net.conv_in._gn_groups = g
            blocks = net.down_blocks if net is self.encoder else net.up_blocks
            for b in blocks:
                samplers = b.downsamplers if net is self.encoder else b.upsamplers
                block(b.resnets, None if samplers is None else samplers[0])
            block(net.mid_block.resnets, None)
            net.mid_block.attentions[0].to_out[0]._gn_groups = g

    # ---- diffusers-compatible surface ---------------------------------------------------
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    # ---- GEMM operand packs: one launch per forward for ALL layers -----------------------------------------
    def _pack_layers(self, root: nn.Module):
        up_convs = {id(m.conv): m for m in root.modules() if isinstance(m, Upsample2D)}
        layers = []
        for m in root.modules():
            if isinstance(m, (B200Conv2d, B200Linear)):
                up = up_convs.get(id(m))
                if up is not None and ops.upconv_supported(m.in_channels, m.out_channels):
                    layers.append((m.weight, m.bias, up._up_packs, 1))       # phase-form packs of the fused Upsample2D path
                else:
                    layers.append((m.weight, m.bias, m._packs, 0))
        return layers

    def _pack_weights(self, which: str):
        """Rebuild the bf16 operand packs of every conv / linear layer of the encoder (+ quant_conv) or the decoder
        (+ post_quant_conv) in ONE kernel launch (they are rebuilt on every forward: fused optimizers update parameters
        without bumping `_version`, ops.PackedWeights).  Returns the plan; its packs are valid until plan.expire().
        Layers that take an unfused path (a hooked Upsample2D conv) find their own pack not valid and pack themselves."""
        cache = self.__dict__.setdefault("_pack_layer_cache", {})
        ck = (which, ops.get_conv_impl())
        if ck not in cache:          # the module tree is fixed; which Upsample2D convs use the phase form depends on the impl
            roots = [self.encoder, self.quant_conv] if which == "encoder" else [self.decoder, self.post_quant_conv]
            cache[ck] = [l for r in roots for l in self._pack_layers(r)]
        layers = cache[ck]
        if not layers or not layers[0][0].is_cuda:
            return None
        plans = self.__dict__.setdefault("_pack_plans", {})
        plan = plans.get(ck)
        if plan is None or plan.key != ops.PackPlan.make_key(layers):
            plan = plans[ck] = ops.PackPlan(layers)
        plan.run()
        return plan

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        if x.dim() != 4:
            raise VcdError(f"encode expects [N, 3, H, W], got {tuple(x.shape)}")
        self._sync_gamma_if_pending()
        ops.clear_colsums()
        # every kernel goes to torch's CURRENT stream of the CURRENT device: make the input's device current for the call
        # (a model on cuda:1 in a process whose current device is cuda:0; backward nodes run under autograd's own guard)
        with _device_of(x):
            plan = self._pack_weights("encoder") if x.is_cuda else None
            try:
                moments = self.quant_conv(self.encoder(x))
            finally:
                if plan is not None:
                    plan.expire()
            dist = DiagonalGaussianDistribution(moments)
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        if not torch.is_grad_enabled():
            ops.clear_colsums()     # decode-only loops (wrapper.decode, logit lens): drop stale producer->consumer hand-offs
        with _device_of(z):
            plan = self._pack_weights("decoder") if z.is_cuda else None
            try:
                y = self.decoder(self.post_quant_conv(z))          # logical [N, 3, H, W] bf16
            finally:
                if plan is not None:
                    plan.expire()
            sample = ops.to_nchw(_phys(y), self.output_dtype)  # contiguous NCHW, fp32
            sample._vcd_nhwc = _phys(y)                        # fused-loss fast path (vcd_b200.losses)
        return SimpleNamespace(sample=sample) if return_dict else (sample,)

    def forward(self, sample, sample_posterior: bool = False, generator=None):
        dist = self.encode(sample).latent_dist
        z = dist.sample(generator) if sample_posterior else dist.mode()
        return self.decode(z)

    def _sync_gamma_if_pending(self):
        """Multi-rank only: train.py nudges on rank 0 alone (train.py:244-246,315-319) and DDP never
        re-broadcasts parameters, so replicas would drift.  One packed broadcast of the 52 GroupNorm
        scales (19 840 values) from rank 0 at the first forward after a tracking step keeps them equal."""
        if not getattr(self, "_gamma_sync_pending", False):
            return
        self._gamma_sync_pending = False
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        gammas = [m.weight.data for m in self.modules() if isinstance(m, B200GroupNorm)]
        flat = torch.cat([g.reshape(-1).float() for g in gammas])
        dist.broadcast(flat, src=0)
        off = 0
        for g in gammas:
            n = g.numel()
            g.copy_(flat[off:off + n].to(g.dtype))
            off += n

    # ---- checkpoints (train.py:412, evaluate.py:99) --------------------------------------
    def save_pretrained(self, save_directory: str, **_):
        from safetensors.torch import save_file
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump(dict(self.config), f, indent=2, sort_keys=True)
        sd = {k: v.detach().contiguous().cpu() for k, v in self.state_dict().items()}
        save_file(sd, os.path.join(save_directory, self.weights_name), metadata={"format": "pt"})

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, torch_dtype=None, **_):
        path = str(pretrained_model_name_or_path)
        if path.startswith("random-init") or (not os.path.isdir(path) and os.environ.get("VCD_ALLOW_RANDOM_INIT") == "1"):
            # "random-init[:seed]" — offline benchmarking / tests (BASELINE.json: random-init SDXL-VAE weights)
            seed = int(path.split(":")[1]) if path.startswith("random-init") and ":" in path else 42
            torch.manual_seed(seed)
            model = cls()
            return model.to(torch_dtype) if torch_dtype is not None else model
        if not os.path.isdir(path):
            try:
                from huggingface_hub import snapshot_download
                path = snapshot_download(path, allow_patterns=["*.json", "*.safetensors"])
            except Exception as e:  # no network in the build image
                raise VcdError(f"cannot resolve VAE '{pretrained_model_name_or_path}': not a local directory and the "
                               f"hub is unreachable ({e}); use a local diffusers-layout directory or "
                               f"'random-init[:seed]'") from e
        with open(os.path.join(path, cls.config_name)) as f:
            cfg = json.load(f)
        model = cls(cfg)
        wfile = os.path.join(path, cls.weights_name)
        from safetensors.torch import load_file
        sd = load_file(wfile)
        sd = {_remap_legacy_key(k): v for k, v in sd.items()}
        for k in list(sd):  # legacy attention projections were stored as 1x1 convs
            if k.endswith("weight") and sd[k].dim() == 4 and ".attentions." in k:
                sd[k] = sd[k][:, :, 0, 0]
        model.load_state_dict(sd, strict=True)
        return model.to(torch_dtype) if torch_dtype is not None else model


def _remap_legacy_key(k: str) -> str:
    """[upstream] old sdxl-vae checkpoints name the attention projections query/key/value/proj_attn."""
    for old, new in ((".query.", ".to_q."), (".key.", ".to_k."), (".value.", ".to_v."), (".proj_attn.", ".to_out.0.")):
        k = k.replace(old, new)
    return k
