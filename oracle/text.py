"""ORACLE (test infrastructure): the text encoders the reference's pipelines call, restated functionally in plain torch.

Third-party algorithms (the reference imports them from `transformers`, pinned here to the 5.5.0 in this image):
  * clip_text_forward  - transformers/models/clip/modeling_clip.py: CLIPTextEmbeddings (token + position embedding),
    CLIPEncoderLayer (pre-LN; CLIPAttention with scale head_dim^-0.5 and a causal mask; CLIPMLP with quick_gelu / gelu),
    final_layer_norm, pooled = features at the EOS token (argmax of the ids for the legacy eos_token_id == 2),
    text_projection (CLIPTextModelWithProjection);
  * t5_encoder_forward - transformers/models/t5/modeling_t5.py: T5LayerNorm (RMS, no bias), T5Attention (no scaling,
    relative_attention_bias[bucket(j - i)] of block 0 added in every block), T5DenseGatedActDense (gelu_new(wi_0 x) * wi_1 x),
    final_layer_norm.
Called by the reference at pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl.py:283 (encode_prompt) and
pipelines/flux/pipeline_flux.py:217-387.  Pinned by tests/test_oracle_golden.py against outputs of the real transformers classes
recorded by oracle/make_golden.py text (tests/golden/text.pt).  Only tests/, smoke() and bench.py's CPU legs may import this."""
import math

import torch
import torch.nn.functional as F


def _lin(sd, p, x, bias=True):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"] if bias else None)


def _act(name, x):
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return F.gelu(x)
    if name in ("gelu_new", "gelu_pytorch_tanh"):
        return F.gelu(x, approximate="tanh")
    raise NotImplementedError(name)


def clip_text_forward(sd, cfg, ids, with_projection=False):
    """-> dict(last_hidden_state, pooler_output, hidden_states (tuple, embeddings first), text_embeds or None)"""
    B, S = ids.shape
    D, nh, eps = cfg["hidden_size"], cfg["num_attention_heads"], cfg["layer_norm_eps"]
    hd = D // nh
    e = "text_model.embeddings"
    x = sd[e + ".token_embedding.weight"][ids] + sd[e + ".position_embedding.weight"][:S]
    hidden = [x]
    mask = torch.full((S, S), float("-inf"), dtype=x.dtype, device=x.device).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}"
        n = F.layer_norm(x, (D,), sd[p + ".layer_norm1.weight"], sd[p + ".layer_norm1.bias"], eps)
        q, k, v = (_lin(sd, f"{p}.self_attn.{nm}", n).view(B, S, nh, hd).transpose(1, 2) for nm in ("q_proj", "k_proj", "v_proj"))
        w = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5 + mask, dim=-1)
        a = (w @ v).transpose(1, 2).reshape(B, S, D)
        x = x + _lin(sd, p + ".self_attn.out_proj", a)
        n = F.layer_norm(x, (D,), sd[p + ".layer_norm2.weight"], sd[p + ".layer_norm2.bias"], eps)
        x = x + _lin(sd, p + ".mlp.fc2", _act(cfg["hidden_act"], _lin(sd, p + ".mlp.fc1", n)))
        hidden.append(x)
    last = F.layer_norm(x, (D,), sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"], eps)
    if cfg.get("eos_token_id", 2) == 2:
        idx = ids.to(torch.int).argmax(-1)
    else:
        idx = (ids.to(torch.int) == cfg["eos_token_id"]).int().argmax(-1)
    pooled = last[torch.arange(B, device=ids.device), idx]
    te = F.linear(pooled, sd["text_projection.weight"]) if with_projection else None
    return dict(last_hidden_state=last, pooler_output=pooled, hidden_states=tuple(hidden), text_embeds=te)


def t5_bucket(rel, num_buckets, max_distance):
    nb = num_buckets // 2
    out = (rel > 0).long() * nb
    n = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, nb - 1))
    return out + torch.where(n < max_exact, n, large)


def _rms(x, w, eps):
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def t5_encoder_forward(sd, cfg, ids):
    """-> dict(last_hidden_state, hidden_states)"""
    B, S = ids.shape
    nh, hd, eps = cfg["num_heads"], cfg["d_kv"], cfg["layer_norm_epsilon"]
    x = sd["shared.weight"][ids]
    pos = torch.arange(S, device=ids.device)
    bucket = t5_bucket(pos[None, :] - pos[:, None], cfg["relative_attention_num_buckets"], cfg["relative_attention_max_distance"])
    bias = sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"][bucket].permute(2, 0, 1)[None]
    hidden = [x]
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer"
        n = _rms(x, sd[p + ".0.layer_norm.weight"], eps)
        q, k, v = (_lin(sd, f"{p}.0.SelfAttention.{nm}", n, bias=False).view(B, S, nh, hd).transpose(1, 2) for nm in ("q", "k", "v"))
        w = torch.softmax(q @ k.transpose(-1, -2) + bias, dim=-1)
        a = (w @ v).transpose(1, 2).reshape(B, S, nh * hd)
        x = x + _lin(sd, p + ".0.SelfAttention.o", a, bias=False)
        n = _rms(x, sd[p + ".1.layer_norm.weight"], eps)
        h = F.gelu(_lin(sd, p + ".1.DenseReluDense.wi_0", n, bias=False), approximate="tanh") * _lin(sd, p + ".1.DenseReluDense.wi_1", n, bias=False)
        x = x + _lin(sd, p + ".1.DenseReluDense.wo", h, bias=False)
        hidden.append(x)
    last = _rms(x, sd["encoder.final_layer_norm.weight"], eps)
    hidden[-1] = last
    return dict(last_hidden_state=last, hidden_states=tuple(hidden))
