"""Parameter names and shapes of the reference modules on the hot path, derived from their configs.

The product has to run where the reference is absent (the GPU box), so the structure logic of the reference
constructors is restated here: UNet2DConditionModel.__init__ (models/unets/unet_2d_condition.py:170-520),
get_down_block / get_up_block (unet_2d_blocks.py:41-480), Transformer2DModel.__init__ (transformer_2d.py:73-320),
BasicTransformerBlock.__init__ (attention.py:790-950), AutoencoderKL / Decoder (autoencoders/vae.py:180-277),
UNet2DModel.__init__ (unets/unet_2d.py:95-245), FluxTransformer2DModel.__init__ (transformer_flux.py:596-669).
tests/test_specs_vs_reference.py checks every name/shape against the reference state_dict (meta device).
"""
import math

import torch


def _t(v, n):
    return tuple(v) if isinstance(v, (list, tuple)) else (v,) * n


class _Spec(dict):
    def lin(self, name, i, o, bias=True):
        self[name + ".weight"] = (o, i)
        if bias:
            self[name + ".bias"] = (o,)

    def conv(self, name, i, o, k):
        self[name + ".weight"] = (o, i, k, k)
        self[name + ".bias"] = (o,)

    def norm(self, name, c):
        self[name + ".weight"] = (c,)
        self[name + ".bias"] = (c,)


def _resnet(s, p, cin, cout, temb):
    s.norm(p + ".norm1", cin)
    s.conv(p + ".conv1", cin, cout, 3)
    if temb:
        s.lin(p + ".time_emb_proj", temb, cout)
    s.norm(p + ".norm2", cout)
    s.conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        s.conv(p + ".conv_shortcut", cin, cout, 1)


def _transformer2d(s, p, channels, n_layers, cross_dim, use_linear_projection):
    s.norm(p + ".norm", channels)
    if use_linear_projection:
        s.lin(p + ".proj_in", channels, channels)
        s.lin(p + ".proj_out", channels, channels)
    else:
        s.conv(p + ".proj_in", channels, channels, 1)
        s.conv(p + ".proj_out", channels, channels, 1)
    for k in range(n_layers):
        b = f"{p}.transformer_blocks.{k}"
        s.norm(b + ".norm1", channels)
        for nm in ("to_q", "to_k", "to_v"):
            s.lin(f"{b}.attn1.{nm}", channels, channels, bias=False)
        s.lin(b + ".attn1.to_out.0", channels, channels)
        s.norm(b + ".norm2", channels)
        s.lin(b + ".attn2.to_q", channels, channels, bias=False)
        s.lin(b + ".attn2.to_k", cross_dim, channels, bias=False)
        s.lin(b + ".attn2.to_v", cross_dim, channels, bias=False)
        s.lin(b + ".attn2.to_out.0", channels, channels)
        s.norm(b + ".norm3", channels)
        s.lin(b + ".ff.net.0.proj", channels, channels * 8)
        s.lin(b + ".ff.net.2", channels * 4, channels)


def _attn_block(s, p, channels):
    """legacy Attention used by UNet2DModel / VAE mid block (attention_processor.py:107-300, bias=True)"""
    s.norm(p + ".group_norm", channels)
    for nm in ("to_q", "to_k", "to_v"):
        s.lin(f"{p}.{nm}", channels, channels)
    s.lin(p + ".to_out.0", channels, channels)


SDXL_UNET_CONFIG = dict(
    sample_size=128, in_channels=4, out_channels=4,
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn",
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
    block_out_channels=(320, 640, 1280), layers_per_block=2, cross_attention_dim=2048,
    transformer_layers_per_block=(1, 2, 10), attention_head_dim=(5, 10, 20), use_linear_projection=True,
    addition_embed_type="text_time", addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816,
    norm_num_groups=32, norm_eps=1e-5, act_fn="silu", flip_sin_to_cos=True, freq_shift=0, downsample_padding=1,
    time_cond_proj_dim=None, center_input_sample=False,
)

SDXL_VAE_CONFIG = dict(
    in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
    up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
    latent_channels=4, norm_num_groups=32, sample_size=1024, scaling_factor=0.13025, act_fn="silu",
    force_upcast=True, use_quant_conv=True, use_post_quant_conv=True, mid_block_add_attention=True,
    shift_factor=None, latents_mean=None, latents_std=None,
)

FLUX_DEV_CONFIG = dict(
    patch_size=1, in_channels=64, out_channels=None, num_layers=19, num_single_layers=38, attention_head_dim=128,
    num_attention_heads=24, joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=True,
    axes_dims_rope=(16, 56, 56),
)

DDPM_TINY_CONFIG = dict(
    sample_size=32, in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(32, 64),
    down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D"),
    attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5, act_fn="silu", flip_sin_to_cos=True, freq_shift=0,
    downsample_padding=1, time_embedding_type="positional", add_attention=True,
)


def unet2d_condition_params(cfg):
    s = _Spec()
    boc = tuple(cfg["block_out_channels"])
    n = len(boc)
    temb = boc[0] * 4
    lpb = _t(cfg.get("layers_per_block", 2), n)
    tlpb = _t(cfg.get("transformer_layers_per_block", 1), n)
    cross = _t(cfg["cross_attention_dim"], n)
    ulp = cfg.get("use_linear_projection", False)
    s.conv("conv_in", cfg["in_channels"], boc[0], 3)
    s.lin("time_embedding.linear_1", boc[0], temb)
    s.lin("time_embedding.linear_2", temb, temb)
    if cfg.get("addition_embed_type") == "text_time":
        s.lin("add_embedding.linear_1", cfg["projection_class_embeddings_input_dim"], temb)
        s.lin("add_embedding.linear_2", temb, temb)
    elif cfg.get("addition_embed_type") is not None:
        raise NotImplementedError(f"addition_embed_type {cfg['addition_embed_type']!r}")
    out_ch = boc[0]
    for i, bt in enumerate(cfg["down_block_types"]):
        in_ch, out_ch = out_ch, boc[i]
        p = f"down_blocks.{i}"
        if bt not in ("DownBlock2D", "CrossAttnDownBlock2D"):
            raise NotImplementedError(bt)
        for j in range(lpb[i]):
            _resnet(s, f"{p}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch, temb)
            if bt == "CrossAttnDownBlock2D":
                _transformer2d(s, f"{p}.attentions.{j}", out_ch, _t(tlpb[i], lpb[i])[j], cross[i], ulp)
        if i != n - 1:
            s.conv(f"{p}.downsamplers.0.conv", out_ch, out_ch, 3)
    if cfg.get("mid_block_type", "UNetMidBlock2DCrossAttn") != "UNetMidBlock2DCrossAttn":
        raise NotImplementedError(cfg["mid_block_type"])
    _resnet(s, "mid_block.resnets.0", boc[-1], boc[-1], temb)
    _transformer2d(s, "mid_block.attentions.0", boc[-1], _t(tlpb[-1], 1)[0], cross[-1], ulp)
    _resnet(s, "mid_block.resnets.1", boc[-1], boc[-1], temb)
    rboc, rlpb, rtl, rcross = boc[::-1], lpb[::-1], tlpb[::-1], cross[::-1]
    out_ch = rboc[0]
    for i, bt in enumerate(cfg["up_block_types"]):
        prev, out_ch = out_ch, rboc[i]
        in_ch = rboc[min(i + 1, n - 1)]
        p = f"up_blocks.{i}"
        if bt not in ("UpBlock2D", "CrossAttnUpBlock2D"):
            raise NotImplementedError(bt)
        nl = rlpb[i] + 1
        for j in range(nl):
            skip = in_ch if j == nl - 1 else out_ch
            rin = prev if j == 0 else out_ch
            _resnet(s, f"{p}.resnets.{j}", rin + skip, out_ch, temb)
            if bt == "CrossAttnUpBlock2D":
                _transformer2d(s, f"{p}.attentions.{j}", out_ch, _t(rtl[i], nl)[j], rcross[i], ulp)
        if i != n - 1:
            s.conv(f"{p}.upsamplers.0.conv", out_ch, out_ch, 3)
    s.norm("conv_norm_out", boc[0])
    s.conv("conv_out", boc[0], cfg["out_channels"], 3)
    return s


def unet2d_params(cfg):
    """UNet2DModel (DDPM)."""
    s = _Spec()
    boc = tuple(cfg["block_out_channels"])
    n = len(boc)
    temb = boc[0] * 4
    lpb = cfg.get("layers_per_block", 2)
    s.conv("conv_in", cfg["in_channels"], boc[0], 3)
    s.lin("time_embedding.linear_1", boc[0], temb)
    s.lin("time_embedding.linear_2", temb, temb)
    out_ch = boc[0]
    for i, bt in enumerate(cfg["down_block_types"]):
        in_ch, out_ch = out_ch, boc[i]
        p = f"down_blocks.{i}"
        if bt not in ("DownBlock2D", "AttnDownBlock2D"):
            raise NotImplementedError(bt)
        for j in range(lpb):
            _resnet(s, f"{p}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch, temb)
            if bt == "AttnDownBlock2D":
                _attn_block(s, f"{p}.attentions.{j}", out_ch)
        if i != n - 1:
            s.conv(f"{p}.downsamplers.0.conv", out_ch, out_ch, 3)
    _resnet(s, "mid_block.resnets.0", boc[-1], boc[-1], temb)
    if cfg.get("add_attention", True):
        _attn_block(s, "mid_block.attentions.0", boc[-1])
    _resnet(s, "mid_block.resnets.1", boc[-1], boc[-1], temb)
    rboc = boc[::-1]
    out_ch = rboc[0]
    for i, bt in enumerate(cfg["up_block_types"]):
        prev, out_ch = out_ch, rboc[i]
        in_ch = rboc[min(i + 1, n - 1)]
        p = f"up_blocks.{i}"
        if bt not in ("UpBlock2D", "AttnUpBlock2D"):
            raise NotImplementedError(bt)
        nl = lpb + 1
        for j in range(nl):
            skip = in_ch if j == nl - 1 else out_ch
            rin = prev if j == 0 else out_ch
            _resnet(s, f"{p}.resnets.{j}", rin + skip, out_ch, temb)
            if bt == "AttnUpBlock2D":
                _attn_block(s, f"{p}.attentions.{j}", out_ch)
        if i != n - 1:
            s.conv(f"{p}.upsamplers.0.conv", out_ch, out_ch, 3)
    s.norm("conv_norm_out", boc[0])
    s.conv("conv_out", boc[0], cfg["out_channels"], 3)
    return s


def vae_decoder_params(cfg):
    """AutoencoderKL: post_quant_conv + decoder.* (the only part the text->image path runs)."""
    s = _Spec()
    boc = tuple(cfg["block_out_channels"])
    lc = cfg["latent_channels"]
    lpb = cfg.get("layers_per_block", 1)
    if cfg.get("use_post_quant_conv", True):
        s.conv("post_quant_conv", lc, lc, 1)
    d = "decoder"
    s.conv(d + ".conv_in", lc, boc[-1], 3)
    _resnet(s, d + ".mid_block.resnets.0", boc[-1], boc[-1], 0)
    if cfg.get("mid_block_add_attention", True):
        _attn_block(s, d + ".mid_block.attentions.0", boc[-1])
    _resnet(s, d + ".mid_block.resnets.1", boc[-1], boc[-1], 0)
    rboc = boc[::-1]
    out_ch = rboc[0]
    for i in range(len(boc)):
        prev, out_ch = out_ch, rboc[i]
        for j in range(lpb + 1):
            _resnet(s, f"{d}.up_blocks.{i}.resnets.{j}", prev if j == 0 else out_ch, out_ch, 0)
        if i != len(boc) - 1:
            s.conv(f"{d}.up_blocks.{i}.upsamplers.0.conv", out_ch, out_ch, 3)
    s.norm(d + ".conv_norm_out", boc[0])
    s.conv(d + ".conv_out", boc[0], cfg["out_channels"], 3)
    return s


def vae_encoder_params(cfg):
    """AutoencoderKL: encoder.* + quant_conv (autoencoders/vae.py:59-180, autoencoder_kl.py:126) - image -> latent moments."""
    s = _Spec()
    boc = tuple(cfg["block_out_channels"])
    lc = cfg["latent_channels"]
    lpb = cfg.get("layers_per_block", 1)
    for t in tuple(cfg.get("down_block_types", ("DownEncoderBlock2D",) * len(boc))):
        if t != "DownEncoderBlock2D":
            raise NotImplementedError(t)
    e = "encoder"
    s.conv(e + ".conv_in", cfg["in_channels"], boc[0], 3)
    out_ch = boc[0]
    for i in range(len(boc)):
        in_ch, out_ch = out_ch, boc[i]
        for j in range(lpb):
            _resnet(s, f"{e}.down_blocks.{i}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch, 0)
        if i != len(boc) - 1:
            s.conv(f"{e}.down_blocks.{i}.downsamplers.0.conv", out_ch, out_ch, 3)
    _resnet(s, e + ".mid_block.resnets.0", boc[-1], boc[-1], 0)
    if cfg.get("mid_block_add_attention", True):
        _attn_block(s, e + ".mid_block.attentions.0", boc[-1])
    _resnet(s, e + ".mid_block.resnets.1", boc[-1], boc[-1], 0)
    s.norm(e + ".conv_norm_out", boc[-1])
    s.conv(e + ".conv_out", boc[-1], 2 * lc, 3)
    if cfg.get("use_quant_conv", True):
        s.conv("quant_conv", 2 * lc, 2 * lc, 1)
    return s


def vae_params(cfg):
    """The whole AutoencoderKL state_dict: encoder + quant_conv + post_quant_conv + decoder."""
    s = vae_encoder_params(cfg)
    s.update(vae_decoder_params(cfg))
    return s


def flux_params(cfg):
    s = _Spec()
    H, hd = cfg["num_attention_heads"], cfg["attention_head_dim"]
    D = H * hd
    outc = cfg.get("out_channels") or cfg["in_channels"]
    s.lin("time_text_embed.timestep_embedder.linear_1", 256, D)
    s.lin("time_text_embed.timestep_embedder.linear_2", D, D)
    if cfg.get("guidance_embeds", False):
        s.lin("time_text_embed.guidance_embedder.linear_1", 256, D)
        s.lin("time_text_embed.guidance_embedder.linear_2", D, D)
    s.lin("time_text_embed.text_embedder.linear_1", cfg["pooled_projection_dim"], D)
    s.lin("time_text_embed.text_embedder.linear_2", D, D)
    s.lin("context_embedder", cfg["joint_attention_dim"], D)
    s.lin("x_embedder", cfg["in_channels"], D)
    for i in range(cfg["num_layers"]):
        p = f"transformer_blocks.{i}"
        s.lin(p + ".norm1.linear", D, 6 * D)
        s.lin(p + ".norm1_context.linear", D, 6 * D)
        for nm in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj"):
            s.lin(f"{p}.attn.{nm}", D, D)
        for nm in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            s[f"{p}.attn.{nm}.weight"] = (hd,)
        s.lin(p + ".attn.to_out.0", D, D)
        s.lin(p + ".attn.to_add_out", D, D)
        s.lin(p + ".ff.net.0.proj", D, 4 * D)
        s.lin(p + ".ff.net.2", 4 * D, D)
        s.lin(p + ".ff_context.net.0.proj", D, 4 * D)
        s.lin(p + ".ff_context.net.2", 4 * D, D)
    for i in range(cfg["num_single_layers"]):
        p = f"single_transformer_blocks.{i}"
        s.lin(p + ".norm.linear", D, 3 * D)
        s.lin(p + ".proj_mlp", D, 4 * D)
        s.lin(p + ".proj_out", 5 * D, D)
        for nm in ("to_q", "to_k", "to_v"):
            s.lin(f"{p}.attn.{nm}", D, D)
        for nm in ("norm_q", "norm_k"):
            s[f"{p}.attn.{nm}.weight"] = (hd,)
    s.lin("norm_out.linear", D, 2 * D)
    s.lin("proj_out", D, cfg.get("patch_size", 1) ** 2 * outc)
    return s


def random_state_dict(spec, seed=0, dtype=torch.bfloat16, device="cpu"):
    """Random weights with torch's default-init distributions (kaiming-uniform bound 1/sqrt(fan_in) for
    Linear/Conv weights and biases, ones/zeros for norms).  Values differ from a seeded reference module;
    parity tests use fixtures exported from the reference instead.

    device="cpu" (default) draws from a CPU generator: the stream the committed fixtures (tests/golden) were recorded
    with.  A CUDA device draws from that device's Philox generator instead (different values, same distributions): the
    11.9 B parameters of the Flux.1-dev shape take seconds instead of minutes."""
    device = torch.device(device)
    g = torch.Generator(device=device if device.type == "cuda" else "cpu").manual_seed(seed)
    sd = {}
    for name, shape in spec.items():
        is_norm = (".norm" in name or name.startswith("norm") or "group_norm" in name or "conv_norm_out" in name) \
            and len(shape) == 1 and not name.endswith("linear.bias")
        if is_norm:
            t = torch.ones(shape, device=g.device) if name.endswith("weight") else torch.zeros(shape, device=g.device)
        else:
            if len(shape) > 1:
                fan_in = math.prod(shape[1:])
            else:
                w = spec.get(name[:-4] + "weight")
                fan_in = math.prod(w[1:]) if w is not None and len(w) > 1 else shape[0]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, device=g.device) * 2 - 1) * bound
        sd[name] = t.to(dtype=dtype, device=device)
    return sd
