"""ORACLE (test infrastructure): the 2-D UNet / VAE block containers restated functionally, one function per
reference class (models/unets/unet_2d_blocks.py).  Pinned by the reference's own golden slices
(tests/models/unets/test_unet_2d_blocks.py) through tests/golden/blocks.pt.
"""
import torch

from . import nn as O


def _count(sd, fmt):
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


def _resnets(sd, p):
    return _count(sd, p + ".resnets.{}.norm1.weight")


def _layers(sd, p):
    return _count(sd, p + ".transformer_blocks.{}.norm1.weight")


def _legacy_attn(sd, p, x, head_dim, groups, eps):
    c = x.shape[1]
    heads = c // head_dim if head_dim is not None else 1
    return O.attention(sd, p, x, None, heads, norm_groups=groups, group_norm_eps=eps, residual_connection=True,
                       rescale_output_factor=1.0)


def down_block_2d(sd, p, x, temb, groups=32, eps=1e-6, downsample_padding=1):
    """DownBlock2D.forward :1346-1370"""
    outs = ()
    for j in range(_resnets(sd, p)):
        x = O.resnet_block(sd, f"{p}.resnets.{j}", x, temb, groups, eps)
        outs += (x,)
    if (p + ".downsamplers.0.conv.weight") in sd:
        x = O.downsample2d(sd, p + ".downsamplers.0", x, padding=downsample_padding)
        outs += (x,)
    return x, outs


def attn_down_block_2d(sd, p, x, temb, head_dim, groups=32, eps=1e-6, downsample_padding=1):
    """AttnDownBlock2D.forward :1105-1140 (downsample_type='conv')"""
    outs = ()
    for j in range(_resnets(sd, p)):
        x = O.resnet_block(sd, f"{p}.resnets.{j}", x, temb, groups, eps)
        x = _legacy_attn(sd, f"{p}.attentions.{j}", x, head_dim, groups, eps)
        outs += (x,)
    if (p + ".downsamplers.0.conv.weight") in sd:
        x = O.downsample2d(sd, p + ".downsamplers.0", x, padding=downsample_padding)
        outs += (x,)
    return x, outs


def cross_attn_down_block_2d(sd, p, x, temb, ehs, heads, groups=32, eps=1e-6, use_linear_projection=False):
    """CrossAttnDownBlock2D.forward :1239-1290"""
    outs = ()
    for j in range(_resnets(sd, p)):
        x = O.resnet_block(sd, f"{p}.resnets.{j}", x, temb, groups, eps)
        a = f"{p}.attentions.{j}"
        x = O.transformer_2d(sd, a, x, ehs, heads, _layers(sd, a), groups, use_linear_projection)
        outs += (x,)
    if (p + ".downsamplers.0.conv.weight") in sd:
        x = O.downsample2d(sd, p + ".downsamplers.0", x)
        outs += (x,)
    return x, outs


def unet_mid_block_2d(sd, p, x, temb, head_dim, groups=32, eps=1e-6):
    """UNetMidBlock2D.forward :736-750"""
    x = O.resnet_block(sd, p + ".resnets.0", x, temb, groups, eps)
    n_attn = _count(sd, p + ".attentions.{}.to_q.weight")
    for j in range(_resnets(sd, p) - 1):
        if j < n_attn:
            x = _legacy_attn(sd, f"{p}.attentions.{j}", x, head_dim, groups, eps)
        x = O.resnet_block(sd, f"{p}.resnets.{j + 1}", x, temb, groups, eps)
    return x


def unet_mid_block_2d_cross_attn(sd, p, x, temb, ehs, heads, groups=32, eps=1e-6, use_linear_projection=False):
    """UNetMidBlock2DCrossAttn.forward :854-900"""
    x = O.resnet_block(sd, p + ".resnets.0", x, temb, groups, eps)
    for j in range(_count(sd, p + ".attentions.{}.norm.weight")):
        a = f"{p}.attentions.{j}"
        x = O.transformer_2d(sd, a, x, ehs, heads, _layers(sd, a), groups, use_linear_projection)
        x = O.resnet_block(sd, f"{p}.resnets.{j + 1}", x, temb, groups, eps)
    return x


def up_block_2d(sd, p, x, res, temb, groups=32, eps=1e-6):
    """UpBlock2D.forward :2524-2570"""
    for j in range(_resnets(sd, p)):
        r, res = res[-1], res[:-1]
        x = torch.cat([x, r], dim=1)
        x = O.resnet_block(sd, f"{p}.resnets.{j}", x, temb, groups, eps)
    if (p + ".upsamplers.0.conv.weight") in sd:
        x = O.upsample2d(sd, p + ".upsamplers.0", x)
    return x


def attn_up_block_2d(sd, p, x, res, temb, head_dim, groups=32, eps=1e-6):
    """AttnUpBlock2D.forward :2270-2300 (upsample_type='conv')"""
    for j in range(_resnets(sd, p)):
        r, res = res[-1], res[:-1]
        x = torch.cat([x, r], dim=1)
        x = O.resnet_block(sd, f"{p}.resnets.{j}", x, temb, groups, eps)
        x = _legacy_attn(sd, f"{p}.attentions.{j}", x, head_dim, groups, eps)
    if (p + ".upsamplers.0.conv.weight") in sd:
        x = O.upsample2d(sd, p + ".upsamplers.0", x)
    return x


def cross_attn_up_block_2d(sd, p, x, res, temb, ehs, heads, groups=32, eps=1e-6, use_linear_projection=False):
    """CrossAttnUpBlock2D.forward :2405-2470"""
    for j in range(_resnets(sd, p)):
        r, res = res[-1], res[:-1]
        x = torch.cat([x, r], dim=1)
        x = O.resnet_block(sd, f"{p}.resnets.{j}", x, temb, groups, eps)
        a = f"{p}.attentions.{j}"
        x = O.transformer_2d(sd, a, x, ehs, heads, _layers(sd, a), groups, use_linear_projection)
    if (p + ".upsamplers.0.conv.weight") in sd:
        x = O.upsample2d(sd, p + ".upsamplers.0", x)
    return x


def up_decoder_block_2d(sd, p, x, groups=32, eps=1e-6):
    """UpDecoderBlock2D.forward :2637-2646"""
    for j in range(_resnets(sd, p)):
        x = O.resnet_block(sd, f"{p}.resnets.{j}", x, None, groups, eps)
    if (p + ".upsamplers.0.conv.weight") in sd:
        x = O.upsample2d(sd, p + ".upsamplers.0", x)
    return x
