"""ORACLE (test infrastructure): AutoencoderKL.decode restated functionally.

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
