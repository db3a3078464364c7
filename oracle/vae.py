"""ORACLE (test infrastructure): AutoencoderKL.decode and .encode restated functionally.

Reference: models/autoencoders/autoencoder_kl.py:199-233 (_decode / decode), autoencoders/vae.py:180-310 (Decoder:
conv_in -> UNetMidBlock2D -> UpDecoderBlock2D x N -> GroupNorm(eps 1e-6) -> SiLU -> conv_out; every resnet/attention
eps is 1e-6 and temb is None), unets/unet_2d_blocks.py:589-750 (UNetMidBlock2D: one head of dim C, group_norm inside
the attention, residual connection), :2575-2660 (UpDecoderBlock2D).
"""
import torch.nn.functional as F

from . import blocks as Bk
from . import nn as O


def _count(sd, fmt):
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


def vae_decode(sd, cfg, z):
    groups = cfg.get("norm_num_groups", 32)
    eps = 1e-6
    if "post_quant_conv.weight" in sd:
        z = O.conv2d(sd, "post_quant_conv", z, padding=0)
    x = O.conv2d(sd, "decoder.conv_in", z)
    x = Bk.unet_mid_block_2d(sd, "decoder.mid_block", x, None, None, groups, eps)
    for i in range(_count(sd, "decoder.up_blocks.{}.resnets.0.norm1.weight")):
        x = Bk.up_decoder_block_2d(sd, f"decoder.up_blocks.{i}", x, groups, eps)
    x = O.group_norm(sd, "decoder.conv_norm_out", x, groups, eps)
    x = F.silu(x)
    return O.conv2d(sd, "decoder.conv_out", x)


def vae_encode(sd, cfg, x):
    """AutoencoderKL._encode (models/autoencoders/autoencoder_kl.py:158-168) -> the moments [B, 2*latent_channels, H/f, W/f]
    DiagonalGaussianDistribution is built from.  Encoder.forward autoencoders/vae.py:152-177: conv_in -> DownEncoderBlock2D x N
    (unets/unet_2d_blocks.py DownEncoderBlock2D: resnets with temb None, Downsample2D(padding=0) = F.pad(0,1,0,1) + stride-2
    conv) -> UNetMidBlock2D -> GroupNorm(eps 1e-6) -> SiLU -> conv_out; then quant_conv."""
    groups = cfg.get("norm_num_groups", 32)
    eps = 1e-6
    h = O.conv2d(sd, "encoder.conv_in", x)
    for i in range(_count(sd, "encoder.down_blocks.{}.resnets.0.norm1.weight")):
        h, _ = Bk.down_block_2d(sd, f"encoder.down_blocks.{i}", h, None, groups, eps, downsample_padding=0)
    h = Bk.unet_mid_block_2d(sd, "encoder.mid_block", h, None, None, groups, eps)
    h = O.group_norm(sd, "encoder.conv_norm_out", h, groups, eps)
    h = F.silu(h)
    h = O.conv2d(sd, "encoder.conv_out", h)
    if "quant_conv.weight" in sd:
        h = O.conv2d(sd, "quant_conv", h, padding=0)
    return h
