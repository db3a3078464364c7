"""ORACLE (test infrastructure): FluxTransformer2DModel.forward restated functionally.

Reference: models/transformers/transformer_flux.py:671-821 (model forward), :443-497 (FluxTransformerBlock),
:383-413 (FluxSingleTransformerBlock), :84-137 (FluxAttnProcessor), :500-526 (FluxPosEmbed);
models/embeddings.py:1120-1183 (get_1d_rotary_pos_embed), :1187-1231 (apply_rotary_emb), :1604-1626
(CombinedTimestepGuidanceTextProjEmbeddings), :2192-2222 (PixArtAlphaTextProjection);
models/normalization.py:157-170 (AdaLayerNormZero), :194-202 (AdaLayerNormZeroSingle), :346-351 (AdaLayerNormContinuous).
"""
import torch
import torch.nn.functional as F

from . import nn as O


def rope_tables(ids, axes_dim, theta=10000):
    """FluxPosEmbed.forward: fp64 frequencies -> fp32 cos/sin [S, sum(axes_dim)] (repeat-interleaved)."""
    cos_out, sin_out = [], []
    pos = ids.float()
    for i in range(ids.shape[-1]):
        dim = axes_dim[i]
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64, device=pos.device) / dim))
        freqs = torch.outer(pos[:, i], freqs)
        cos_out.append(freqs.cos().repeat_interleave(2, dim=1, output_size=freqs.shape[1] * 2).float())
        sin_out.append(freqs.sin().repeat_interleave(2, dim=1, output_size=freqs.shape[1] * 2).float())
    return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)


def apply_rotary_emb(x, cos, sin):
    """embeddings.py:1187-1231 with use_real=True, use_real_unbind_dim=-1, sequence_dim=1; x [B, S, H, D]"""
    cos = cos[None, :, None, :]
    sin = sin[None, :, None, :]
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rotated = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rotated.float() * sin).to(x.dtype)


def _rms(sd, name, x, eps=1e-6):
    return F.rms_norm(x, (x.shape[-1],), sd[name + ".weight"], eps)


def _time_text_embed(sd, timestep, guidance, pooled, guidance_embeds):
    p = "time_text_embed"
    t_proj = O.get_timestep_embedding(timestep, 256, flip_sin_to_cos=True, downscale_freq_shift=0)
    emb = O.timestep_embedding_mlp(sd, p + ".timestep_embedder", t_proj.to(pooled.dtype))
    if guidance_embeds:
        g_proj = O.get_timestep_embedding(guidance, 256, flip_sin_to_cos=True, downscale_freq_shift=0)
        emb = emb + O.timestep_embedding_mlp(sd, p + ".guidance_embedder", g_proj.to(pooled.dtype))
    txt = O.linear(sd, p + ".text_embedder.linear_2", F.silu(O.linear(sd, p + ".text_embedder.linear_1", pooled)))
    return emb + txt


def _attention(sd, p, heads, hd, x, ctx, rope):
    B = x.shape[0]
    q = O.linear(sd, p + ".to_q", x).unflatten(-1, (-1, hd))
    k = O.linear(sd, p + ".to_k", x).unflatten(-1, (-1, hd))
    v = O.linear(sd, p + ".to_v", x).unflatten(-1, (-1, hd))
    q = _rms(sd, p + ".norm_q", q)
    k = _rms(sd, p + ".norm_k", k)
    if ctx is not None:
        eq = _rms(sd, p + ".norm_added_q", O.linear(sd, p + ".add_q_proj", ctx).unflatten(-1, (-1, hd)))
        ek = _rms(sd, p + ".norm_added_k", O.linear(sd, p + ".add_k_proj", ctx).unflatten(-1, (-1, hd)))
        ev = O.linear(sd, p + ".add_v_proj", ctx).unflatten(-1, (-1, hd))
        q = torch.cat([eq, q], dim=1)
        k = torch.cat([ek, k], dim=1)
        v = torch.cat([ev, v], dim=1)
    q = apply_rotary_emb(q, *rope)
    k = apply_rotary_emb(k, *rope)
    # dispatch_attention_fn -> _native_attention (attention_dispatch.py:3678-3717): permute to (B,H,S,D), SDPA, back
    o = F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3))
    o = o.permute(0, 2, 1, 3).flatten(2, 3).to(q.dtype)
    if ctx is not None:
        c_o, x_o = o.split_with_sizes([ctx.shape[1], o.shape[1] - ctx.shape[1]], dim=1)
        return O.linear(sd, p + ".to_out.0", x_o.contiguous()), O.linear(sd, p + ".to_add_out", c_o.contiguous())
    return o


def _ln(x):
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)


def _ff(sd, p, x):
    h = F.gelu(O.linear(sd, p + ".net.0.proj", x), approximate="tanh")
    return O.linear(sd, p + ".net.2", h)


def flux_forward(sd, cfg, hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance=None):
    heads, hd = cfg["num_attention_heads"], cfg["attention_head_dim"]
    x = O.linear(sd, "x_embedder", hidden_states)
    timestep = timestep.to(x.dtype) * 1000
    if guidance is not None:
        guidance = guidance.to(x.dtype) * 1000
    temb = _time_text_embed(sd, timestep, guidance, pooled_projections, cfg.get("guidance_embeds", False) and guidance is not None)
    ctx = O.linear(sd, "context_embedder", encoder_hidden_states)
    rope = rope_tables(torch.cat((txt_ids, img_ids), dim=0), cfg["axes_dims_rope"])
    n_double = 0
    while f"transformer_blocks.{n_double}.norm1.linear.weight" in sd:
        n_double += 1
    for i in range(n_double):
        p = f"transformer_blocks.{i}"
        e = O.linear(sd, p + ".norm1.linear", F.silu(temb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = e.chunk(6, dim=1)
        nx = _ln(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        ec = O.linear(sd, p + ".norm1_context.linear", F.silu(temb))
        c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = ec.chunk(6, dim=1)
        nc = _ln(ctx) * (1 + c_scale_msa[:, None]) + c_shift_msa[:, None]
        a_x, a_c = _attention(sd, p + ".attn", heads, hd, nx, nc, rope)
        x = x + gate_msa.unsqueeze(1) * a_x
        nx = _ln(x) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        x = x + gate_mlp.unsqueeze(1) * _ff(sd, p + ".ff", nx)
        ctx = ctx + c_gate_msa.unsqueeze(1) * a_c
        nc = _ln(ctx) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        ctx = ctx + c_gate_mlp.unsqueeze(1) * _ff(sd, p + ".ff_context", nc)
        if ctx.dtype == torch.float16:
            ctx = ctx.clip(-65504, 65504)
    n_single = 0
    while f"single_transformer_blocks.{n_single}.norm.linear.weight" in sd:
        n_single += 1
    for i in range(n_single):
        p = f"single_transformer_blocks.{i}"
        T = ctx.shape[1]
        h = torch.cat([ctx, x], dim=1)
        residual = h
        e = O.linear(sd, p + ".norm.linear", F.silu(temb))
        shift, scale, gate = e.chunk(3, dim=1)
        nh = _ln(h) * (1 + scale[:, None]) + shift[:, None]
        mlp = F.gelu(O.linear(sd, p + ".proj_mlp", nh), approximate="tanh")
        attn = _attention(sd, p + ".attn", heads, hd, nh, None, rope)
        h = torch.cat([attn, mlp], dim=2)
        h = gate.unsqueeze(1) * O.linear(sd, p + ".proj_out", h)
        h = residual + h
        if h.dtype == torch.float16:
            h = h.clip(-65504, 65504)
        ctx, x = h[:, :T], h[:, T:]
    e = O.linear(sd, "norm_out.linear", F.silu(temb).to(x.dtype))
    scale, shift = torch.chunk(e, 2, dim=1)
    x = _ln(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    return O.linear(sd, "proj_out", x)
