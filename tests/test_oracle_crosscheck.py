"""Cross-check of the ORACLE's VAE arithmetic against an independent, installed implementation of the same architecture.

diffusers (whose AutoencoderKL the reference calls) is absent from this image, so the oracle cannot be pinned at that
boundary (oracle/torch_vae.py header).  HF transformers, which IS installed, ships its own port of the latent-diffusion /
taming autoencoder that the SDXL-VAE checkpoint was trained with and that diffusers' AutoencoderKL was ported from:
`JanusVQVAEEncoder` / `JanusVQVAEDecoder` (GroupNorm(32, eps 1e-6) -> x*sigmoid(x) -> conv3x3 ResNet blocks with 1x1
`nin_shortcut`, single-head attention block with scale C^-0.5, pad (0,1,0,1) + stride-2 conv down-sampling, nearest x2 + conv
up-sampling, 3 blocks per decoder level).  With the SDXL-VAE hyper-parameters (128 base channels, multipliers 1/2/4/4, two
blocks per level, double latent) and its extra last-level attention blocks removed, it is the same network: same 83.65 M
parameters less the two 1x1 quant convs.  The oracle's weights are copied in by name and the two must agree to fp32
rounding.  This does not replace a diffusers pin; it shows the restatement is the published architecture."""
import pytest
import torch

from oracle.torch_vae import build_oracle


def _janus():
    try:
        from transformers.models.janus.configuration_janus import JanusVQVAEConfig
        from transformers.models.janus.modeling_janus import JanusVQVAEDecoder, JanusVQVAEEncoder
    except Exception as e:      # transformers without the Janus model
        pytest.skip(f"transformers' Janus VQ-VAE is not available: {e}")
    cfg = JanusVQVAEConfig(embed_dim=4, latent_channels=4, double_latent=True, in_channels=3, out_channels=3, base_channels=128,
                           channel_multiplier=[1, 2, 4, 4], num_res_blocks=2, dropout=0.0)
    enc, dec = JanusVQVAEEncoder(cfg), JanusVQVAEDecoder(cfg)
    enc.down[-1].attn = torch.nn.ModuleList()      # SDXL-VAE: attention in the mid block only
    dec.up[0].attn = torch.nn.ModuleList()
    return enc, dec


def _copy_resnet(dst, src):
    for n in ("norm1", "conv1", "norm2", "conv2"):
        getattr(dst, n).load_state_dict(getattr(src, n).state_dict())
    if src.conv_shortcut is not None:
        dst.nin_shortcut.load_state_dict(src.conv_shortcut.state_dict())


def _copy_attn(dst, src):
    dst.norm.load_state_dict(src.group_norm.state_dict())
    for d, s in (("q", src.to_q), ("k", src.to_k), ("v", src.to_v), ("proj_out", src.to_out[0])):
        getattr(dst, d).weight.data.copy_(s.weight.data[:, :, None, None])      # Linear [C, C] == 1x1 conv
        getattr(dst, d).bias.data.copy_(s.bias.data)


def _copy_mid(dst, src):
    _copy_resnet(dst.block_1, src.resnets[0])
    _copy_attn(dst.attn_1, src.attentions[0])
    _copy_resnet(dst.block_2, src.resnets[1])


def test_oracle_encoder_decoder_equal_the_transformers_port_of_the_ldm_autoencoder():
    torch.manual_seed(0)
    oracle = build_oracle(42)
    with torch.no_grad():       # non-trivial affine parameters everywhere
        for m in oracle.modules():
            if isinstance(m, torch.nn.GroupNorm):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    enc, dec = _janus()
    assert sum(p.numel() for p in enc.parameters()) == sum(p.numel() for p in oracle.encoder.parameters()) == 34_163_592
    assert sum(p.numel() for p in dec.parameters()) == sum(p.numel() for p in oracle.decoder.parameters()) == 49_490_179
    e, d = oracle.encoder, oracle.decoder
    enc.conv_in.load_state_dict(e.conv_in.state_dict())
    for i, blk in enumerate(e.down_blocks):
        for j, r in enumerate(blk.resnets):
            _copy_resnet(enc.down[i].block[j], r)
        if blk.downsamplers is not None:
            enc.down[i].downsample.conv.load_state_dict(blk.downsamplers[0].conv.state_dict())
    _copy_mid(enc.mid, e.mid_block)
    enc.norm_out.load_state_dict(e.conv_norm_out.state_dict())
    enc.conv_out.load_state_dict(e.conv_out.state_dict())
    dec.conv_in.load_state_dict(d.conv_in.state_dict())
    _copy_mid(dec.mid, d.mid_block)
    for i, blk in enumerate(d.up_blocks):
        for j, r in enumerate(blk.resnets):
            _copy_resnet(dec.up[i].block[j], r)
        if blk.upsamplers is not None:
            dec.up[i].upsample.conv.load_state_dict(blk.upsamplers[0].conv.state_dict())
    dec.norm_out.load_state_dict(d.conv_norm_out.state_dict())
    dec.conv_out.load_state_dict(d.conv_out.state_dict())
    x = torch.rand(2, 3, 32, 32) * 2 - 1
    z = torch.randn(2, 4, 4, 4)
    with torch.no_grad():
        want_e, got_e = enc(x.clone()), e(x)
        want_d, got_d = dec(z.clone()), d(z)
    assert got_e.shape == want_e.shape == (2, 8, 4, 4) and got_d.shape == want_d.shape == (2, 3, 32, 32)
    err_e = float((got_e - want_e).abs().max() / want_e.abs().max())
    err_d = float((got_d - want_d).abs().max() / want_d.abs().max())
    assert err_e < 1e-5 and err_d < 1e-5, (err_e, err_d)
