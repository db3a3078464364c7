"""ORACLE (test infrastructure): AutoencoderKL.decode restated functionally.

Reference: models/autoencoders/autoencoder_kl.py:199-233 (_decode / decode), autoencoders/vae.py:180-310 (Decoder:
conv_in -> UNetMidBlock2D -> UpDecoderBlock2D x N -> GroupNorm(eps 1e-6) -> SiLU -> conv_out; every resnet/attention
eps is 1e-6 and temb is None), unets/unet_2d_blocks.py:589-750 (UNetMidBlock2D: one head of dim C, group_norm inside
the attention, residual connection), :2575-2660 (UpDecoderBlock2D).
"""
import torch.nn.functional as F

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
    x = O.resnet_block(sd, "decoder.mid_block.resnets.0", x, None, groups, eps)
    if "decoder.mid_block.attentions.0.to_q.weight" in sd:
        x = O.attention(sd, "decoder.mid_block.attentions.0", x, None, heads=1, norm_groups=groups, group_norm_eps=eps,
                        residual_connection=True, rescale_output_factor=1.0)
    x = O.resnet_block(sd, "decoder.mid_block.resnets.1", x, None, groups, eps)
    n_up = _count(sd, "decoder.up_blocks.{}.resnets.0.norm1.weight")
    for i in range(n_up):
        p = f"decoder.up_blocks.{i}"
        for j in range(_count(sd, p + ".resnets.{}.norm1.weight")):
            x = O.resnet_block(sd, f"{p}.resnets.{j}", x, None, groups, eps)
        if (p + ".upsamplers.0.conv.weight") in sd:
            x = O.upsample2d(sd, p + ".upsamplers.0", x)
    x = O.group_norm(sd, "decoder.conv_norm_out", x, groups, eps)
    x = F.silu(x)
    return O.conv2d(sd, "decoder.conv_out", x)
