"""Drop-in text encoders on libb200diff.so: the step before the denoising loop.

The reference does not own these models - its pipelines call `transformers` classes:
  * SDXL  pipeline_stable_diffusion_xl.py:283 `encode_prompt`: `text_encoder(ids, output_hidden_states=True)` for
    CLIPTextModel (CLIP-L) and CLIPTextModelWithProjection (OpenCLIP bigG); uses `.hidden_states[-2]` and, of the second
    encoder, `[0]` (= text_embeds);
  * Flux  pipeline_flux.py:217-387: `text_encoder(ids, output_hidden_states=False).pooler_output` (CLIP-L) and
    `text_encoder_2(ids, output_hidden_states=False)[0]` (T5EncoderModel, T5-XXL v1.1, 512 tokens).
Third-party dependency: transformers 5.5.0 (the version in this image; `oracle/text.py` restates the algorithms of
models/clip/modeling_clip.py and models/t5/modeling_t5.py and is pinned against that package's own outputs).

Same call surface and output objects as those classes for what the pipelines touch (`.config`, `.dtype`, `.device`,
`__call__(input_ids, attention_mask=None, output_hidden_states=...)`, outputs indexable and with named fields).  Parameter
names are transformers' state_dict names.

All numerics run in the CUDA kernels of the denoisers: the tcgen05 GEMM (fused QKV, bias / QuickGELU / GELU / tanh-gated GEGLU /
residual epilogues), the LayerNorm kernel (with its RMSNorm mode for T5) and b200_text_attention (causal mask / relative-position
bias).  The token-embedding gather and the T5 bucket table are index arithmetic done once per call with torch (model boundary).
"""
import math

import torch

from . import ops, packing
from .config import FrozenConfig
from .ops import ACT_GELU_ERF, ACT_GELU_TANH, ACT_QUICK_GELU

CLIP_L_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                     max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=768, eos_token_id=2,
                     bos_token_id=0, pad_token_id=1)
CLIP_BIGG_CONFIG = dict(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                        max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5, projection_dim=1280, eos_token_id=2,
                        bos_token_id=0, pad_token_id=1)
T5_XXL_CONFIG = dict(vocab_size=32128, d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64, relative_attention_num_buckets=32,
                     relative_attention_max_distance=128, layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu")


def clip_text_params(cfg, with_projection=False):
    """transformers CLIPTextModel(/WithProjection).state_dict() names and shapes."""
    D, I = cfg["hidden_size"], cfg["intermediate_size"]
    s = {"text_model.embeddings.token_embedding.weight": (cfg["vocab_size"], D),
         "text_model.embeddings.position_embedding.weight": (cfg["max_position_embeddings"], D)}
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}"
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[f"{p}.self_attn.{nm}.weight"], s[f"{p}.self_attn.{nm}.bias"] = (D, D), (D,)
        for nm in ("layer_norm1", "layer_norm2"):
            s[f"{p}.{nm}.weight"], s[f"{p}.{nm}.bias"] = (D,), (D,)
        s[f"{p}.mlp.fc1.weight"], s[f"{p}.mlp.fc1.bias"] = (I, D), (I,)
        s[f"{p}.mlp.fc2.weight"], s[f"{p}.mlp.fc2.bias"] = (D, I), (D,)
    s["text_model.final_layer_norm.weight"], s["text_model.final_layer_norm.bias"] = (D,), (D,)
    if with_projection:
        s["text_projection.weight"] = (cfg["projection_dim"], D)
    return s


def t5_encoder_params(cfg):
    """transformers T5EncoderModel.state_dict() names and shapes (`shared.weight` is tied to `encoder.embed_tokens.weight`)."""
    D, inner, F = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]
    s = {"shared.weight": (cfg["vocab_size"], D)}
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer"
        for nm in ("q", "k", "v"):
            s[f"{p}.0.SelfAttention.{nm}.weight"] = (inner, D)
        s[f"{p}.0.SelfAttention.o.weight"] = (D, inner)
        if i == 0:
            s[f"{p}.0.SelfAttention.relative_attention_bias.weight"] = (cfg["relative_attention_num_buckets"], cfg["num_heads"])
        s[f"{p}.0.layer_norm.weight"] = (D,)
        s[f"{p}.1.DenseReluDense.wi_0.weight"], s[f"{p}.1.DenseReluDense.wi_1.weight"] = (F, D), (F, D)
        s[f"{p}.1.DenseReluDense.wo.weight"] = (D, F)
        s[f"{p}.1.layer_norm.weight"] = (D,)
    s["encoder.final_layer_norm.weight"] = (D,)
    return s


def random_state_dict(spec, seed=0, dtype=torch.bfloat16, device="cpu"):
    """Random text-encoder weights (tests / benchmarks: there are no checkpoints offline): N(0, 1/fan_in) matrices, N(0, 0.02)
    biases, norm weights around 1, unit-scale embeddings.  Deterministic in (spec, seed) on a CPU generator."""
    device = torch.device(device)
    on_gpu = device.type == "cuda"  # a CUDA device draws from its own Philox stream (other values, same distributions): seconds for T5-XXL
    g = torch.Generator(device=device if on_gpu else "cpu").manual_seed(seed)
    sd = {}
    for name, shape in spec.items():
        r = torch.randn(shape, generator=g, device=g.device)
        if "norm" in name and name.endswith("weight") and len(shape) == 1:
            t = 1.0 + 0.1 * r
        elif "embedding" in name or name in ("shared.weight",) or "relative_attention_bias" in name:
            t = 0.5 * r
        elif len(shape) == 1:
            t = 0.02 * r
        else:
            t = r * (1.0 / math.sqrt(shape[1]))
            if name.endswith("SelfAttention.q.weight"):
                t = t * 0.125  # T5 does not scale q.k: its own init draws q with std (d_model * d_kv)^-0.5 (modeling_t5.py _init_weights);
                               # unit-variance queries would saturate every softmax and make the network chaotic in any 16-bit dtype
        sd[name] = t.to(dtype=dtype, device=device)
    return sd


class _Output:
    """ModelOutput stand-in: named fields + tuple indexing over the fields that are not None, in declaration order."""

    def __init__(self, **fields):
        self._names = [k for k, v in fields.items() if v is not None]
        for k, v in fields.items():
            setattr(self, k, v)

    def __getitem__(self, i):
        if isinstance(i, str):
            return getattr(self, i)
        return getattr(self, self._names[i])

    def __len__(self):
        return len(self._names)

    def keys(self):
        return list(self._names)


class _Base(torch.nn.Module):
    def _reg(self, t, device, dtype=None):
        name = f"w{self._n}"
        self._n += 1
        self.register_buffer(name, t.to(device=device, dtype=dtype or self._dtype).contiguous(), persistent=False)
        return name

    def W(self, name):
        return self._buffers[name]

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._buffers["w0"].device

    _hf_architectures = ()

    @classmethod
    def from_pretrained(cls, path, subfolder=None, variant=None, torch_dtype=torch.bfloat16, device="cuda"):
        """A transformers `save_pretrained` directory (config.json with `architectures`, model[.variant].safetensors or its sharded
        index, or pytorch_model.bin) - e.g. `text_encoder/`, `text_encoder_2/` of an SDXL or Flux checkpoint - -> this shell.
        Local directories only."""
        import os

        from . import checkpoint
        root = os.path.join(path, subfolder) if subfolder else path
        raw = checkpoint.load_config(root)
        arch = raw.get("architectures") or []
        if arch and not set(arch) & set(cls._hf_architectures):
            raise ValueError(f"{root} holds a {arch}, not one of {cls._hf_architectures}")
        cfg = {k: raw[k] for k in cls._default_config if k in raw}
        sd = checkpoint.load_state_dict(root, variant=variant, safetensors_name="model.safetensors", pickle_name="pytorch_model.bin")
        return cls(cfg, sd, dtype=torch_dtype, device=device)

    @staticmethod
    def _check(spec, sd):
        for k, shp in spec.items():
            if k not in sd:
                raise ValueError(f"state_dict is missing {k}")
            if tuple(sd[k].shape) != tuple(shp):
                raise ValueError(f"{k}: expected shape {tuple(shp)}, got {tuple(sd[k].shape)}")


class CLIPTextModel(_Base):
    """transformers CLIPTextModel (models/clip/modeling_clip.py: CLIPTextTransformer) - and, with `with_projection=True`,
    CLIPTextModelWithProjection."""
    with_projection = False
    _hf_architectures = ("CLIPTextModel",)
    _default_config = CLIP_L_CONFIG

    def __init__(self, config, state_dict, dtype=torch.bfloat16, device="cuda"):
        super().__init__()
        cfg = dict(CLIP_L_CONFIG)
        cfg.update(config)
        self.config = FrozenConfig(cfg)
        if cfg["hidden_size"] != 64 * cfg["num_attention_heads"]:
            raise NotImplementedError("the text attention kernel is built for head_dim 64 (CLIP-L, OpenCLIP bigG)")
        acts = {"quick_gelu": ACT_QUICK_GELU, "gelu": ACT_GELU_ERF, "gelu_new": ACT_GELU_TANH, "gelu_pytorch_tanh": ACT_GELU_TANH}
        if cfg["hidden_act"] not in acts:
            raise NotImplementedError(f"hidden_act={cfg['hidden_act']!r}")
        self._act = acts[cfg["hidden_act"]]
        self._dtype, self._n = dtype, 0
        self._check(clip_text_params(cfg, self.with_projection), state_dict)
        dev = torch.device(device)
        g = lambda k: state_dict[k].to(torch.float32)  # noqa: E731
        R = lambda t: self._reg(t, dev)  # noqa: E731
        e = "text_model.embeddings"
        self.tok, self.pos = R(g(e + ".token_embedding.weight")), R(g(e + ".position_embedding.weight"))
        self.layers = []
        for i in range(cfg["num_hidden_layers"]):
            p = f"text_model.encoder.layers.{i}"
            a = p + ".self_attn"
            qkv_w = torch.cat([g(a + ".q_proj.weight"), g(a + ".k_proj.weight"), g(a + ".v_proj.weight")], 0)
            qkv_b = torch.cat([g(a + ".q_proj.bias"), g(a + ".k_proj.bias"), g(a + ".v_proj.bias")], 0)
            self.layers.append(dict(
                l1w=R(g(p + ".layer_norm1.weight")), l1b=R(g(p + ".layer_norm1.bias")),
                qkv=R(packing.pack_linear_weight(qkv_w)), qkvb=R(qkv_b),
                ow=R(packing.pack_linear_weight(g(a + ".out_proj.weight"))), ob=R(g(a + ".out_proj.bias")),
                l2w=R(g(p + ".layer_norm2.weight")), l2b=R(g(p + ".layer_norm2.bias")),
                f1w=R(packing.pack_linear_weight(g(p + ".mlp.fc1.weight"))), f1b=R(g(p + ".mlp.fc1.bias")),
                f2w=R(packing.pack_linear_weight(g(p + ".mlp.fc2.weight"))), f2b=R(g(p + ".mlp.fc2.bias"))))
        self.fw, self.fb = R(g("text_model.final_layer_norm.weight")), R(g("text_model.final_layer_norm.bias"))
        self.proj = R(g("text_projection.weight")) if self.with_projection else None

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, output_attentions=None, output_hidden_states=None,
                return_dict=True, **kw):
        if attention_mask is not None or position_ids is not None or output_attentions:
            raise NotImplementedError("attention_mask / position_ids / output_attentions: the diffusers pipelines pass input_ids only")
        cfg = self.config
        dev = self.device
        ids = input_ids.to(dev)
        B, S = ids.shape
        if S > cfg["max_position_embeddings"]:
            raise ValueError(f"sequence length {S} exceeds max_position_embeddings {cfg['max_position_embeddings']}")
        D, I, nh, eps = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_attention_heads"], cfg["layer_norm_eps"]
        # CLIPTextEmbeddings: token_embedding(ids) + position_embedding(arange): a gather and one 16-bit add
        x = (self.W(self.tok).index_select(0, ids.reshape(-1)).view(B, S, D) + self.W(self.pos)[:S]).view(B * S, D)
        hidden = [x.view(B, S, D)]
        for L in self.layers:
            n = ops.layer_norm(x, eps=eps, gamma=self.W(L["l1w"]), beta=self.W(L["l1b"]))
            qkv = ops.linear(n, self.W(L["qkv"]), 3 * D, bias=self.W(L["qkvb"])).view(B, S, 3 * D)
            a = ops.text_attention(qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:], heads=nh, scale=0.125, causal=True)
            x = ops.linear(a.view(B * S, D), self.W(L["ow"]), D, bias=self.W(L["ob"]), residual=x)
            n = ops.layer_norm(x, eps=eps, gamma=self.W(L["l2w"]), beta=self.W(L["l2b"]))
            h = ops.linear(n, self.W(L["f1w"]), I, bias=self.W(L["f1b"]), act=self._act)
            x = ops.linear(h, self.W(L["f2w"]), D, bias=self.W(L["f2b"]), residual=x)
            hidden.append(x.view(B, S, D))
        last = ops.layer_norm(x, eps=eps, gamma=self.W(self.fw), beta=self.W(self.fb)).view(B, S, D)
        # pooled = the features at the EOS token (CLIPTextTransformer.forward: argmax of the ids for the legacy eos_token_id == 2)
        if cfg["eos_token_id"] == 2:
            idx = ids.to(torch.int).argmax(dim=-1)
        else:
            idx = (ids.to(torch.int) == cfg["eos_token_id"]).int().argmax(dim=-1)
        pooled = last[torch.arange(B, device=dev), idx]
        hs = tuple(hidden) if output_hidden_states else None
        if self.with_projection:
            te = ops.small_linear(pooled.contiguous(), self.W(self.proj))
            out = _Output(text_embeds=te, last_hidden_state=last, hidden_states=hs)
        else:
            out = _Output(last_hidden_state=last, pooler_output=pooled, hidden_states=hs)
        return out if return_dict else tuple(out[i] for i in range(len(out)))


class CLIPTextModelWithProjection(CLIPTextModel):
    with_projection = True
    _hf_architectures = ("CLIPTextModelWithProjection",)


def t5_relative_position_bucket(relative_position, num_buckets=32, max_distance=128):
    """T5Attention._relative_position_bucket, bidirectional (encoder) form: half of the buckets per sign; within a sign, exact
    buckets up to num_buckets/4 and logarithmically spaced ones up to max_distance."""
    nb = num_buckets // 2
    ret = (relative_position > 0).to(torch.long) * nb
    n = relative_position.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


class T5EncoderModel(_Base):
    """transformers T5EncoderModel (models/t5/modeling_t5.py: T5Stack of T5Block(T5LayerSelfAttention, T5LayerFF)), gated-GELU
    feed-forward (T5 v1.1 / XXL).  Flux calls it without an attention mask."""
    _hf_architectures = ("T5EncoderModel", "T5ForConditionalGeneration", "T5Model")
    _default_config = T5_XXL_CONFIG

    def __init__(self, config, state_dict, dtype=torch.bfloat16, device="cuda"):
        super().__init__()
        cfg = dict(T5_XXL_CONFIG)
        cfg.update(config)
        self.config = FrozenConfig(cfg)
        if cfg["d_kv"] != 64:
            raise NotImplementedError("the text attention kernel is built for d_kv 64 (every T5 v1.1 size)")
        if cfg["feed_forward_proj"] != "gated-gelu":
            raise NotImplementedError(f"feed_forward_proj={cfg['feed_forward_proj']!r}: only the gated-GELU feed-forward of T5 v1.1 is built")
        self._dtype, self._n = dtype, 0
        sd = dict(state_dict)
        if "shared.weight" not in sd and "encoder.embed_tokens.weight" in sd:
            sd["shared.weight"] = sd["encoder.embed_tokens.weight"]
        self._check(t5_encoder_params(cfg), sd)
        dev = torch.device(device)
        g = lambda k: sd[k].to(torch.float32)  # noqa: E731
        R = lambda t: self._reg(t, dev)  # noqa: E731
        self.tok = R(g("shared.weight"))
        self.inner = cfg["num_heads"] * 64
        self.rel = R(g("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"))
        F = cfg["d_ff"]
        self.ff_tile = ops.pick_tile_n(1 << 20, 2 * F, True)
        self.blocks = []
        for i in range(cfg["num_layers"]):
            p = f"encoder.block.{i}.layer"
            a = p + ".0.SelfAttention"
            qkv = torch.cat([g(a + ".q.weight"), g(a + ".k.weight"), g(a + ".v.weight")], 0)
            # hidden = gelu_new(wi_0 x) * (wi_1 x): the GEGLU epilogue multiplies VALUE rows by act(GATE rows) -> value = wi_1, gate = wi_0
            ffw, _ = packing.pack_geglu(torch.cat([g(p + ".1.DenseReluDense.wi_1.weight"), g(p + ".1.DenseReluDense.wi_0.weight")], 0), None, self.ff_tile)
            self.blocks.append(dict(n1=R(g(p + ".0.layer_norm.weight")), qkv=R(packing.pack_linear_weight(qkv)),
                                    o=R(packing.pack_linear_weight(g(a + ".o.weight"))), n2=R(g(p + ".1.layer_norm.weight")),
                                    ff=R(ffw), wo=R(packing.pack_linear_weight(g(p + ".1.DenseReluDense.wo.weight")))))
        self.fn = R(g("encoder.final_layer_norm.weight"))
        self._bias_cache = {}

    def _position_bias(self, S):
        """T5Attention.compute_bias: [heads, S, S] fp32, shared by every block (only block 0 owns the embedding)."""
        if S not in self._bias_cache:
            cfg = self.config
            dev = self.device
            ctx = torch.arange(S, dtype=torch.long, device=dev)[:, None]
            mem = torch.arange(S, dtype=torch.long, device=dev)[None, :]
            bucket = t5_relative_position_bucket(mem - ctx, cfg["relative_attention_num_buckets"], cfg["relative_attention_max_distance"])
            self._bias_cache[S] = self.W(self.rel)[bucket].permute(2, 0, 1).float().contiguous()
        return self._bias_cache[S]

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, output_attentions=None, output_hidden_states=None, return_dict=True, **kw):
        if attention_mask is not None or output_attentions:
            raise NotImplementedError("attention_mask / output_attentions: FluxPipeline passes input_ids only")
        cfg = self.config
        dev = self.device
        ids = input_ids.to(dev)
        B, S = ids.shape
        if S > 512:
            raise NotImplementedError("at most 512 tokens (the text attention kernel keeps K and V of a head in shared memory)")
        D, inner, nh, eps, F = cfg["d_model"], self.inner, cfg["num_heads"], cfg["layer_norm_epsilon"], cfg["d_ff"]
        x = self.W(self.tok).index_select(0, ids.reshape(-1))
        bias = self._position_bias(S)
        hidden = [x.view(B, S, D)]
        for blk in self.blocks:
            n = ops.layer_norm(x, eps=eps, gamma=self.W(blk["n1"]), rms=True)
            qkv = ops.linear(n, self.W(blk["qkv"]), 3 * inner).view(B, S, 3 * inner)
            a = ops.text_attention(qkv[:, :, :inner], qkv[:, :, inner:2 * inner], qkv[:, :, 2 * inner:], heads=nh, scale=1.0, bias=bias)
            x = ops.linear(a.view(B * S, inner), self.W(blk["o"]), D, residual=x)
            n = ops.layer_norm(x, eps=eps, gamma=self.W(blk["n2"]), rms=True)
            h = ops.linear(n, self.W(blk["ff"]), 2 * F, geglu=True, act=ACT_GELU_TANH, tile_n=self.ff_tile)
            x = ops.linear(h, self.W(blk["wo"]), D, residual=x)
            hidden.append(x.view(B, S, D))
        last = ops.layer_norm(x, eps=eps, gamma=self.W(self.fn), rms=True).view(B, S, D)
        hidden[-1] = last  # T5Stack appends the normalised final state
        out = _Output(last_hidden_state=last, hidden_states=tuple(hidden) if output_hidden_states else None)
        return out if return_dict else tuple(out[i] for i in range(len(out)))
