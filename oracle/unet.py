"""ORACLE (test infrastructure): functional restatement of UNet2DConditionModel.forward (SDXL-style
configs) and UNet2DModel.forward (DDPM, config 0).  Structure is read off the reference state_dict keys, so
the only config values used are the ones the reference forward itself consults.

Reference: models/unets/unet_2d_condition.py:979-1240, unet_2d_blocks.py:1239-1290 (CrossAttnDownBlock2D),
:1346-1370 (DownBlock2D), :854-900 (UNetMidBlock2DCrossAttn), :2405-2470 (CrossAttnUpBlock2D), :2524-2570
(UpBlock2D); models/unets/unet_2d.py:249-350 (UNet2DModel), unet_2d_blocks.py:736-750 (UNetMidBlock2D),
:1018-1060 (AttnDownBlock2D), :2185-2230 (AttnUpBlock2D).
"""
import torch
import torch.nn.functional as F

from . import blocks as Bk
from . import nn as O


def _count(sd, fmt):
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


def _as_tuple(v, n):
    return tuple(v) if isinstance(v, (list, tuple)) else (v,) * n


def unet2d_condition_forward(sd, cfg, sample, timestep, encoder_hidden_states, added_cond_kwargs=None):
    """Returns the noise prediction, same dtype as `sample`."""
    groups = cfg.get("norm_num_groups", 32)
    eps = cfg.get("norm_eps", 1e-5)
    n_down = _count(sd, "down_blocks.{}.resnets.0.norm1.weight")
    heads_cfg = cfg.get("num_attention_heads") or cfg["attention_head_dim"]
    heads = _as_tuple(heads_cfg, n_down)
    use_lin = cfg.get("use_linear_projection", False)

    # ---- time embedding  (get_time_embed :852-874, time_embedding :1082)
    timesteps = timestep
    if not torch.is_tensor(timesteps):
        timesteps = torch.tensor([timesteps], dtype=torch.float64 if isinstance(timestep, float) else torch.int64)
    elif timesteps.dim() == 0:
        timesteps = timesteps[None]
    timesteps = timesteps.expand(sample.shape[0])
    t_emb = O.get_timestep_embedding(timesteps, cfg["block_out_channels"][0], flip_sin_to_cos=cfg.get("flip_sin_to_cos", True),
                                     downscale_freq_shift=cfg.get("freq_shift", 0)).to(sample.dtype)
    emb = O.timestep_embedding_mlp(sd, "time_embedding", t_emb)
    # ---- get_aug_embed :890-930 (text_time)
    if cfg.get("addition_embed_type") == "text_time":
        text_embeds = added_cond_kwargs["text_embeds"]
        time_ids = added_cond_kwargs["time_ids"]
        time_embeds = O.get_timestep_embedding(time_ids.flatten(), cfg["addition_time_embed_dim"], flip_sin_to_cos=cfg.get("flip_sin_to_cos", True),
                                               downscale_freq_shift=cfg.get("freq_shift", 0))
        time_embeds = time_embeds.reshape((text_embeds.shape[0], -1))
        add_embeds = torch.concat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
        emb = emb + O.timestep_embedding_mlp(sd, "add_embedding", add_embeds)

    x = O.conv2d(sd, "conv_in", sample)
    skips = (x,)
    # ---- down blocks
    for i in range(n_down):
        p = f"down_blocks.{i}"
        if (p + ".attentions.0.norm.weight") in sd:
            x, outs = Bk.cross_attn_down_block_2d(sd, p, x, emb, encoder_hidden_states, heads[i], groups, eps, use_lin)
        else:
            x, outs = Bk.down_block_2d(sd, p, x, emb, groups, eps)
        skips += outs
    # ---- mid block
    if "mid_block.resnets.0.norm1.weight" in sd:
        x = Bk.unet_mid_block_2d_cross_attn(sd, "mid_block", x, emb, encoder_hidden_states, heads[-1], groups, eps, use_lin)
    # ---- up blocks
    n_up = _count(sd, "up_blocks.{}.resnets.0.norm1.weight")
    rev_heads = tuple(reversed(heads))
    for i in range(n_up):
        p = f"up_blocks.{i}"
        n_res = _count(sd, p + ".resnets.{}.norm1.weight")
        res, skips = skips[-n_res:], skips[:-n_res]
        if (p + ".attentions.0.norm.weight") in sd:
            x = Bk.cross_attn_up_block_2d(sd, p, x, res, emb, encoder_hidden_states, rev_heads[i], groups, eps, use_lin)
        else:
            x = Bk.up_block_2d(sd, p, x, res, emb, groups, eps)
    x = O.group_norm(sd, "conv_norm_out", x, groups, eps)
    x = F.silu(x)
    return O.conv2d(sd, "conv_out", x)


def unet2d_forward(sd, cfg, sample, timestep):
    """UNet2DModel.forward (models/unets/unet_2d.py:249-350): positional time embedding, DownBlock2D /
    AttnDownBlock2D, UNetMidBlock2D (one attention), AttnUpBlock2D / UpBlock2D, no skip-conv variants."""
    groups = cfg.get("norm_num_groups", 32)
    eps = cfg.get("norm_eps", 1e-5)
    head_dim = cfg.get("attention_head_dim", 8)
    timesteps = timestep
    if not torch.is_tensor(timesteps):
        timesteps = torch.tensor([timesteps], dtype=torch.long)
    elif timesteps.dim() == 0:
        timesteps = timesteps[None]
    timesteps = timesteps * torch.ones(sample.shape[0], dtype=timesteps.dtype)
    t_emb = O.get_timestep_embedding(timesteps, cfg["block_out_channels"][0], flip_sin_to_cos=cfg.get("flip_sin_to_cos", True),
                                     downscale_freq_shift=cfg.get("freq_shift", 0)).to(sample.dtype)
    emb = O.timestep_embedding_mlp(sd, "time_embedding", t_emb)

    x = O.conv2d(sd, "conv_in", sample)
    skips = (x,)
    n_down = _count(sd, "down_blocks.{}.resnets.0.norm1.weight")
    pad = cfg.get("downsample_padding", 1)
    for i in range(n_down):
        p = f"down_blocks.{i}"
        if (p + ".attentions.0.to_q.weight") in sd:
            x, outs = Bk.attn_down_block_2d(sd, p, x, emb, head_dim, groups, eps, pad)
        else:
            x, outs = Bk.down_block_2d(sd, p, x, emb, groups, eps, pad)
        skips += outs
    x = Bk.unet_mid_block_2d(sd, "mid_block", x, emb, head_dim, groups, eps)
    n_up = _count(sd, "up_blocks.{}.resnets.0.norm1.weight")
    for i in range(n_up):
        p = f"up_blocks.{i}"
        n_res = _count(sd, p + ".resnets.{}.norm1.weight")
        res, skips = skips[-n_res:], skips[:-n_res]
        if (p + ".attentions.0.to_q.weight") in sd:
            x = Bk.attn_up_block_2d(sd, p, x, res, emb, head_dim, groups, eps)
        else:
            x = Bk.up_block_2d(sd, p, x, res, emb, groups, eps)
    x = O.group_norm(sd, "conv_norm_out", x, groups, eps)
    x = F.silu(x)
    return O.conv2d(sd, "conv_out", x)
