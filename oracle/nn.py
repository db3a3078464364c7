"""ORACLE (test infrastructure, never shipped / never on the product path).

Functional CPU restatement of the reference building blocks of the denoising hot path.  Every function
takes a reference-style ``state_dict`` slice (``sd`` + ``prefix``) and calls the same ATen ops, in the same
order, as the reference module it cites (paths relative to /root/reference/src/diffusers/).  Pinned against
the reference itself by oracle/make_golden.py -> tests/golden/*.pt (see tests/test_oracle_golden.py).
"""
import math

import torch
import torch.nn.functional as F


def _w(sd, name):
    return sd[name]


def _b(sd, name):
    return sd.get(name)


def linear(sd, prefix, x):
    """nn.Linear"""
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def conv2d(sd, prefix, x, stride=1, padding=1):
    """nn.Conv2d"""
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding)


def group_norm(sd, prefix, x, groups, eps):
    """nn.GroupNorm"""
    return F.group_norm(x, groups, sd.get(prefix + ".weight"), sd.get(prefix + ".bias"), eps)


def layer_norm(sd, prefix, x, eps=1e-5, affine=True):
    """nn.LayerNorm"""
    w = sd.get(prefix + ".weight") if affine else None
    b = sd.get(prefix + ".bias") if affine else None
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000):
    """models/embeddings.py:27-79"""
    assert timesteps.dim() == 1
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def timestep_embedding_mlp(sd, prefix, sample, act="silu"):
    """models/embeddings.py:1262-1308 TimestepEmbedding.forward (no cond_proj, no post_act)"""
    sample = linear(sd, prefix + ".linear_1", sample)
    sample = F.silu(sample) if act == "silu" else F.gelu(sample)
    return linear(sd, prefix + ".linear_2", sample)


def resnet_block(sd, prefix, x, temb, groups=32, eps=1e-5, output_scale_factor=1.0, pre_norm=True):
    """models/resnet.py:319-377 ResnetBlock2D.forward (time_embedding_norm='default', no up/down, silu)"""
    h = group_norm(sd, prefix + ".norm1", x, groups, eps)
    h = F.silu(h)
    h = conv2d(sd, prefix + ".conv1", h)
    if temb is not None and (prefix + ".time_emb_proj.weight") in sd:
        t = linear(sd, prefix + ".time_emb_proj", F.silu(temb))[:, :, None, None]
        h = h + t
    h = group_norm(sd, prefix + ".norm2", h, groups, eps)
    h = F.silu(h)
    h = conv2d(sd, prefix + ".conv2", h)
    if (prefix + ".conv_shortcut.weight") in sd:
        x = conv2d(sd, prefix + ".conv_shortcut", x, padding=0)
    return (x + h) / output_scale_factor


def downsample2d(sd, prefix, x, padding=1):
    """models/downsampling.py:130-150 Downsample2D.forward (use_conv=True)"""
    if padding == 0:
        x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    return conv2d(sd, prefix + ".conv", x, stride=2, padding=padding)


def upsample2d(sd, prefix, x):
    """models/upsampling.py:140-190 Upsample2D.forward (nearest 2x + conv)"""
    x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    return conv2d(sd, prefix + ".conv", x)


def attention(sd, prefix, hidden_states, encoder_hidden_states=None, heads=8, norm_groups=None, group_norm_eps=1e-5,
              residual_connection=False, rescale_output_factor=1.0):
    """models/attention_processor.py:2705-2789 AttnProcessor2_0.__call__ (no mask, no norm_q/k)"""
    residual = hidden_states
    input_ndim = hidden_states.ndim
    if input_ndim == 4:
        b, c, hgt, wid = hidden_states.shape
        hidden_states = hidden_states.view(b, c, hgt * wid).transpose(1, 2)
    batch = hidden_states.shape[0]
    if norm_groups is not None:
        hidden_states = F.group_norm(hidden_states.transpose(1, 2), norm_groups, sd[prefix + ".group_norm.weight"],
                                     sd[prefix + ".group_norm.bias"], group_norm_eps).transpose(1, 2)
    q = linear(sd, prefix + ".to_q", hidden_states)
    ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
    k = linear(sd, prefix + ".to_k", ctx)
    v = linear(sd, prefix + ".to_v", ctx)
    inner = k.shape[-1]
    hd = inner // heads
    q = q.view(batch, -1, heads, hd).transpose(1, 2)
    k = k.view(batch, -1, heads, hd).transpose(1, 2)
    v = v.view(batch, -1, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(batch, -1, heads * hd).to(q.dtype)
    o = linear(sd, prefix + ".to_out.0", o)
    if input_ndim == 4:
        o = o.transpose(-1, -2).reshape(b, c, hgt, wid)
    if residual_connection:
        o = o + residual
    return o / rescale_output_factor


def geglu_feed_forward(sd, prefix, x):
    """models/attention.py:1736-1744 FeedForward.forward with GEGLU (models/activations.py:113-123)"""
    h = linear(sd, prefix + ".net.0.proj", x)
    h, gate = h.chunk(2, dim=-1)
    h = h * F.gelu(gate)
    return linear(sd, prefix + ".net.2", h)


def basic_transformer_block(sd, prefix, x, encoder_hidden_states, heads):
    """models/attention.py:960-1075 BasicTransformerBlock.forward, norm_type='layer_norm'"""
    n = layer_norm(sd, prefix + ".norm1", x)
    x = attention(sd, prefix + ".attn1", n, None, heads) + x
    if (prefix + ".attn2.to_q.weight") in sd:
        n = layer_norm(sd, prefix + ".norm2", x)
        x = attention(sd, prefix + ".attn2", n, encoder_hidden_states, heads) + x
    n = layer_norm(sd, prefix + ".norm3", x)
    x = geglu_feed_forward(sd, prefix + ".ff", n) + x
    return x


def transformer_2d(sd, prefix, x, encoder_hidden_states, heads, num_layers, groups=32, use_linear_projection=True):
    """models/transformers/transformer_2d.py:324-512 (continuous inputs)"""
    b, c, hgt, wid = x.shape
    residual = x
    h = group_norm(sd, prefix + ".norm", x, groups, 1e-6)
    if use_linear_projection:
        h = h.permute(0, 2, 3, 1).reshape(b, hgt * wid, c)
        h = linear(sd, prefix + ".proj_in", h)
    else:
        h = conv2d(sd, prefix + ".proj_in", h, padding=0)
        inner = h.shape[1]
        h = h.permute(0, 2, 3, 1).reshape(b, hgt * wid, inner)
    for i in range(num_layers):
        h = basic_transformer_block(sd, f"{prefix}.transformer_blocks.{i}", h, encoder_hidden_states, heads)
    if use_linear_projection:
        h = linear(sd, prefix + ".proj_out", h)
        h = h.reshape(b, hgt, wid, -1).permute(0, 3, 1, 2).contiguous()
    else:
        h = h.reshape(b, hgt, wid, -1).permute(0, 3, 1, 2).contiguous()
        h = conv2d(sd, prefix + ".proj_out", h, padding=0)
    return h + residual
