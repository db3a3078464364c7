"""Plug-ins for the reference's two attention extension points, for users who keep the reference's own
modules and only swap the attention operator:

* `B200AttnProcessor` - an `AttnProcessor2_0`-compatible processor for the legacy `Attention` module
  (models/attention_processor.py:2696-2789).  Install with `unet.set_attn_processor(B200AttnProcessor())`
  (models/attention.py:64).  The weights stay owned by the `Attention` module; nn.Linear weights are already the
  K-major [N, K] layout the tcgen05 GEMM wants.
* `b200_attention_backend` - a `dispatch_attention_fn` backend ((B, S, H, D) in/out, models/attention_dispatch.py:390,
  signature of `_native_attention` :3678).  `AttentionBackendName` is a closed Enum (:214), so it is installed into the
  NATIVE slot: `install_native_backend()`.
Both run the fused tcgen05 attention kernel; masks, dropout, causal and GQA are not on the hot path and raise.
"""
import torch

from . import ops, packing


def _w(linear):
    w = linear.weight
    if w.shape[1] % 64 == 0 and w.is_contiguous():
        return w  # identical to packing.pack_linear_weight(w)
    cached = getattr(linear, "_b200_packed", None)
    if cached is None or cached[0] != w._version:
        linear._b200_packed = (w._version, packing.pack_linear_weight(w.detach()))
    return linear._b200_packed[1]


def _linear(linear, x2d, residual=None):
    return ops.linear(x2d, _w(linear), linear.weight.shape[0], bias=linear.bias, residual=residual)


class B200AttnProcessor:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, *args, **kwargs):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is outside the accelerated hot path")
        if getattr(attn, "spatial_norm", None) is not None or getattr(attn, "norm_q", None) is not None or \
                getattr(attn, "norm_k", None) is not None or getattr(attn, "norm_cross", False):
            raise NotImplementedError("spatial_norm / qk-norm / norm_cross variants are outside the hot path")
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b, c, h, w = hidden_states.shape
            hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
        if getattr(attn, "group_norm", None) is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        hidden_states = hidden_states.contiguous()
        B, S, C = hidden_states.shape
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states.contiguous()
        Sk = ctx.shape[1]
        x2 = hidden_states.view(B * S, C)
        q = _linear(attn.to_q, x2)
        k = _linear(attn.to_k, ctx.view(B * Sk, ctx.shape[-1]))
        v = _linear(attn.to_v, ctx.view(B * Sk, ctx.shape[-1]))
        inner = k.shape[-1]
        head_dim = inner // attn.heads
        o = ops.attention(q.view(B, S, inner), k.view(B, Sk, inner), v.view(B, Sk, inner), heads=attn.heads, head_dim=head_dim)
        o = _linear(attn.to_out[0], o.view(B * S, inner)).view(B, S, -1)
        if input_ndim == 4:
            o = o.transpose(-1, -2).reshape(b, c, h, w)
        if getattr(attn, "residual_connection", False):
            o = o + residual
        rescale = getattr(attn, "rescale_output_factor", 1.0)
        return o if rescale == 1.0 else o / rescale


def b200_attention_backend(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False,
                           return_lse=False, _parallel_config=None):
    """(B, S, H, D) -> (B, S, H, D), the contract of the reference's attention backends."""
    if attn_mask is not None or dropout_p != 0.0 or is_causal or enable_gqa or return_lse or _parallel_config is not None:
        raise NotImplementedError("mask / dropout / causal / GQA / LSE / context parallel are outside the accelerated hot path")
    B, S, H, D = query.shape

    def flat(t):
        if t.stride(3) != 1 or t.stride(2) != D:
            t = t.contiguous()
        return t.as_strided((t.shape[0], t.shape[1], H * D), (t.stride(0), t.stride(1), 1))

    o = ops.attention(flat(query), flat(key), flat(value), heads=H, head_dim=D, scale=scale)
    return o.view(B, S, H, D)


def install_native_backend():
    """Route the reference's NATIVE attention backend (what FluxAttnProcessor dispatches to by default) to the kernel.
    Needs the reference package; returns the previous function so callers can restore it."""
    from diffusers.models import attention_dispatch as ad
    reg = ad._AttentionBackendRegistry
    prev = reg._backends.get(ad.AttentionBackendName.NATIVE)
    reg._backends[ad.AttentionBackendName.NATIVE] = b200_attention_backend
    return prev
