"""Drop-in UNet2DConditionModel (SDXL-style) running on libb200diff.so.

Same constructor config, `forward(sample, timestep, encoder_hidden_states, ..., added_cond_kwargs,
return_dict)` signature, `.config`, `.dtype`, `.device`, `.add_embedding.linear_1.in_features` as the reference
(models/unets/unet_2d_condition.py:76,979; what StableDiffusionXLPipeline touches is listed in SURVEY.md §8b).
Internals are not a port: activations are NHWC bf16/fp16 buffers, every op is a C-ABI call into the sm_100a
kernels (ops.py), skip-connection concats and layout permutes never materialise, the 17 resnet time
projections are one skinny GEMM, and the text K/V projections of all 70 cross-attention layers are one GEMM
that is cached per prompt.  There is no PyTorch fallback.
"""
import types

import torch

from . import ops, packing, specs
from .checkpoint import FromPretrainedMixin
from .config import FrozenConfig
from .ops import ACT_NONE, ACT_SILU


def _t(v, n):
    return tuple(v) if isinstance(v, (list, tuple)) else (v,) * n


class UNet2DConditionOutput:
    def __init__(self, sample):
        self.sample = sample


class UNet2DConditionModel(torch.nn.Module, FromPretrainedMixin):
    _ref_class_names = ("UNet2DConditionModel",)

    @classmethod
    def _param_spec(cls, cfg):
        full = dict(specs.SDXL_UNET_CONFIG)
        full.update(cfg)
        return specs.unet2d_condition_params(full)

    _supports_cuda_graph = True

    def __init__(self, config, state_dict, dtype=torch.bfloat16, device="cuda", fold_norms=True):
        """fold_norms: the three LayerNorms of every BasicTransformerBlock are folded into the Linear that follows them
        (b200_conv_gemm_args.ln_*: gamma and the mean subtraction go into the packed weight, the row statistics come from the
        epilogue of the GEMM that produced the residual stream) - no LayerNorm kernel runs.  The fold re-rounds those weights
        once; fold_norms=False keeps bit-exact copies of the checkpoint's to_q/k/v, attn2.to_q and ff.net.0.proj weights
        (exact `save_pretrained` round trips) and runs b200_layer_norm instead."""
        super().__init__()
        cfg = dict(specs.SDXL_UNET_CONFIG)
        cfg.update(config)
        self.config = FrozenConfig(cfg)
        self._dtype = dtype
        self._fold_norms = bool(fold_norms)
        self._n = 0
        spec = specs.unet2d_condition_params(cfg)
        missing = [k for k in spec if k not in state_dict]
        if missing:
            raise ValueError(f"state_dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
        for k, shp in spec.items():
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: expected shape {tuple(shp)}, got {tuple(state_dict[k].shape)}")
        self._validate_config(cfg)
        self._build(state_dict, torch.device(device))
        # attribute the SDXL pipeline reads (pipeline_stable_diffusion_xl.py:737)
        self.add_embedding = types.SimpleNamespace(
            linear_1=types.SimpleNamespace(in_features=cfg.get("projection_class_embeddings_input_dim")))
        self._kv_key = None
        self._kv = None
        self._graphs = {}
        self.use_cuda_graph = False

    # Constructor options of the reference (models/unets/unet_2d_condition.py:178-239) whose non-default values change the
    # numerics and are NOT implemented here: a checkpoint that sets one must fail loudly, not be approximated.
    _ONLY = dict(act_fn=("silu", "swish"), mid_block_type=("UNetMidBlock2DCrossAttn",), addition_embed_type=(None, "text_time"),
                 time_embedding_type=("positional",), resnet_time_scale_shift=("default",), attention_type=("default",),
                 class_embed_type=(None,), num_class_embeds=(None,), time_cond_proj_dim=(None,), timestep_post_act=(None,),
                 time_embedding_act_fn=(None,), time_embedding_dim=(None,), encoder_hid_dim=(None,), encoder_hid_dim_type=(None,),
                 cross_attention_norm=(None,), reverse_transformer_layers_per_block=(None,), mid_block_only_cross_attention=(None, False),
                 only_cross_attention=(False,), dual_cross_attention=(False,), upcast_attention=(False, None),
                 resnet_skip_time_act=(False,), class_embeddings_concat=(False,), center_input_sample=(False,),
                 resnet_out_scale_factor=(1.0, 1), mid_block_scale_factor=(1.0, 1), conv_in_kernel=(3,), conv_out_kernel=(3,),
                 downsample_padding=(1,))
    _BLOCKS = ("DownBlock2D", "CrossAttnDownBlock2D", "UpBlock2D", "CrossAttnUpBlock2D")

    @classmethod
    def _validate_config(cls, cfg):
        for k, allowed in cls._ONLY.items():
            if k in cfg:
                v = cfg[k]
                v = tuple(v) if isinstance(v, list) else v
                if isinstance(v, tuple) and len(set(v)) == 1:
                    v = v[0]
                if v not in allowed:
                    raise NotImplementedError(f"UNet2DConditionModel option {k}={cfg[k]!r} is outside the accelerated hot path "
                                              f"(supported: {allowed})")
        for t in tuple(cfg["down_block_types"]) + tuple(cfg["up_block_types"]):
            if t not in cls._BLOCKS:
                raise NotImplementedError(f"block type {t} is outside the accelerated hot path ({cls._BLOCKS})")
        if cfg.get("norm_num_groups") is None:
            raise NotImplementedError("norm_num_groups=None (no GroupNorm) is outside the hot path")
        if isinstance(cfg.get("cross_attention_dim"), (list, tuple)) and len(set(cfg["cross_attention_dim"])) > 1:
            raise NotImplementedError("per-block cross_attention_dim is outside the hot path")

    # ------------------------------------------------------------------ weights
    def _reg(self, t, device, dtype=None):
        name = f"w{self._n}"
        self._n += 1
        self.register_buffer(name, t.to(device=device, dtype=dtype or self._dtype).contiguous(), persistent=False)
        return name

    def _build(self, sd, device):
        cfg = self.config
        dt = self._dtype
        R = lambda t: self._reg(t, device)  # noqa: E731
        R32 = lambda t: self._reg(t, device, torch.float32)  # noqa: E731  (per-column fp32 vectors of a folded LayerNorm)
        g = lambda k: sd[k].to(torch.float32)  # noqa: E731  (packing in fp32, stored in model dtype)
        boc = tuple(cfg["block_out_channels"])
        n = len(boc)
        heads = _t(cfg.get("num_attention_heads") or cfg["attention_head_dim"], n)
        self.time_dim = boc[0]
        self.temb_dim = boc[0] * 4
        self.in_pad = packing.rup(cfg["in_channels"], 8)

        def lin_small(p):
            return dict(w=R(g(p + ".weight")), b=R(g(p + ".bias")))

        self.time_embedding = [lin_small("time_embedding.linear_1"), lin_small("time_embedding.linear_2")]
        self.add_emb = None
        if cfg.get("addition_embed_type") == "text_time":
            self.add_emb = [lin_small("add_embedding.linear_1"), lin_small("add_embedding.linear_2")]

        w_in = g("conv_in.weight")
        w_in = torch.nn.functional.pad(w_in, (0, 0, 0, 0, 0, self.in_pad - w_in.shape[1]))
        self.conv_in = dict(w=R(packing.pack_conv_weight(w_in)), b=R(g("conv_in.bias")), n=boc[0])

        temb_w, temb_b = [], []
        self._temb_total = 0
        kv_w = []
        self._kv_total = 0

        def resnet(p, split=None):
            w1 = g(p + ".conv1.weight")
            cout, cin = w1.shape[0], w1.shape[1]
            r = dict(cin=cin, cout=cout, split=split,
                     n1w=R(g(p + ".norm1.weight")), n1b=R(g(p + ".norm1.bias")),
                     c1w=R(packing.pack_conv_weight(w1, split)), c1b=R(g(p + ".conv1.bias")),
                     n2w=R(g(p + ".norm2.weight")), n2b=R(g(p + ".norm2.bias")),
                     c2w=R(packing.pack_conv_weight(g(p + ".conv2.weight"))), c2b=R(g(p + ".conv2.bias")))
            r["temb_off"] = self._temb_total
            temb_w.append(g(p + ".time_emb_proj.weight"))
            temb_b.append(g(p + ".time_emb_proj.bias"))
            self._temb_total += cout
            if (p + ".conv_shortcut.weight") in sd:
                r["scw"] = R(packing.pack_conv_weight(g(p + ".conv_shortcut.weight"), split))
                r["scb"] = R(g(p + ".conv_shortcut.bias"))
            elif split is not None:
                raise NotImplementedError("two-source resnet without conv_shortcut")
            return r

        def transformer(p, ch, n_layers, nheads):
            t = dict(ch=ch, heads=nheads, nw=R(g(p + ".norm.weight")), nb=R(g(p + ".norm.bias")),
                     piw=R(packing.pack_linear_weight(g(p + ".proj_in.weight").reshape(ch, ch))), pib=R(g(p + ".proj_in.bias")),
                     pow=R(packing.pack_linear_weight(g(p + ".proj_out.weight").reshape(ch, ch))), pob=R(g(p + ".proj_out.bias")),
                     blocks=[])
            if ch % nheads or (ch // nheads) not in (64, 128):
                raise NotImplementedError(f"attention head_dim {ch // nheads if nheads else '?'}: the tcgen05 kernel supports 64 and 128")
            for k in range(n_layers):
                b = f"{p}.transformer_blocks.{k}"
                qkv = torch.cat([g(b + ".attn1.to_q.weight"), g(b + ".attn1.to_k.weight"), g(b + ".attn1.to_v.weight")], 0)
                q2 = g(b + ".attn2.to_q.weight")
                ff1, ff1bias = g(b + ".ff.net.0.proj.weight"), g(b + ".ff.net.0.proj.bias")
                tile = ops.pick_tile_n(1 << 20, ff1.shape[0], True)
                fold = {}
                if self._fold_norms:
                    # LN(x) W^T + b = rstd (x W'^T) + (b + W beta), W' = W gamma - its row means: packing.fold_layer_norm
                    qkv, lb1, sh1 = packing.fold_layer_norm(qkv, g(b + ".norm1.weight"), g(b + ".norm1.bias"), None, dt)
                    q2, lb2, sh2 = packing.fold_layer_norm(q2, g(b + ".norm2.weight"), g(b + ".norm2.bias"), None, dt)
                    ff1, ff1bias, sh3 = packing.fold_layer_norm(ff1, g(b + ".norm3.weight"), g(b + ".norm3.bias"), ff1bias, dt)
                    fold = dict(qkv_b=R(lb1), q2_b=R(lb2), qkv_sh=R32(sh1), q2_sh=R32(sh2), ff1_sh=R32(sh3))
                ff1w, ff1b = packing.pack_geglu(ff1, ff1bias, tile)
                qkvw, q2w = packing.pack_linear_weight(qkv), packing.pack_linear_weight(q2)
                blk = dict(
                    l1w=R(g(b + ".norm1.weight")), l1b=R(g(b + ".norm1.bias")),
                    qkv=R(qkvw),
                    ow=R(packing.pack_linear_weight(g(b + ".attn1.to_out.0.weight"))), ob=R(g(b + ".attn1.to_out.0.bias")),
                    l2w=R(g(b + ".norm2.weight")), l2b=R(g(b + ".norm2.bias")),
                    q2=R(q2w),
                    o2w=R(packing.pack_linear_weight(g(b + ".attn2.to_out.0.weight"))), o2b=R(g(b + ".attn2.to_out.0.bias")),
                    l3w=R(g(b + ".norm3.weight")), l3b=R(g(b + ".norm3.bias")),
                    ff1w=R(ff1w), ff1b=R(ff1b), ff1n=ff1.shape[0], ff1tile=tile,
                    ff2w=R(packing.pack_linear_weight(g(b + ".ff.net.2.weight"))), ff2b=R(g(b + ".ff.net.2.bias")),
                    kv_off=self._kv_total, **fold)
                kv_w.append(torch.cat([g(b + ".attn2.to_k.weight"), g(b + ".attn2.to_v.weight")], 0))
                self._kv_total += 2 * ch
                t["blocks"].append(blk)
            return t

        lpb = _t(cfg.get("layers_per_block", 2), n)
        tlpb = _t(cfg.get("transformer_layers_per_block", 1), n)
        self.down = []
        out_ch = boc[0]
        skip_ch = [boc[0]]
        for i, bt in enumerate(cfg["down_block_types"]):
            p = f"down_blocks.{i}"
            blk = dict(res=[], attn=[], down=None)
            for j in range(lpb[i]):
                blk["res"].append(resnet(f"{p}.resnets.{j}"))
                out_ch = boc[i]
                if bt == "CrossAttnDownBlock2D":
                    blk["attn"].append(transformer(f"{p}.attentions.{j}", out_ch, _t(tlpb[i], lpb[i])[j], heads[i]))
                skip_ch.append(out_ch)
            if i != n - 1:
                blk["down"] = dict(w=R(packing.pack_conv_weight(g(f"{p}.downsamplers.0.conv.weight"))),
                                   b=R(g(f"{p}.downsamplers.0.conv.bias")), n=out_ch)
                skip_ch.append(out_ch)
            self.down.append(blk)
        self.mid = dict(res=[resnet("mid_block.resnets.0"), resnet("mid_block.resnets.1")],
                        attn=[transformer("mid_block.attentions.0", boc[-1], _t(tlpb[-1], 1)[0], heads[-1])])
        self.up = []
        rboc, rlpb, rtl, rheads = boc[::-1], lpb[::-1], tlpb[::-1], heads[::-1]
        cur = rboc[0]
        for i, bt in enumerate(cfg["up_block_types"]):
            p = f"up_blocks.{i}"
            blk = dict(res=[], attn=[], up=None)
            nl = rlpb[i] + 1
            for j in range(nl):
                sk = skip_ch.pop()
                blk["res"].append(resnet(f"{p}.resnets.{j}", split=(cur, sk)))
                cur = rboc[i]
                if bt == "CrossAttnUpBlock2D":
                    blk["attn"].append(transformer(f"{p}.attentions.{j}", cur, _t(rtl[i], nl)[j], rheads[i]))
            if i != n - 1:
                # the 3x3 weight is kept for reference_state_dict (the parity fold sums taps: not invertible); the forward runs
                # the four 2x2 parity weights (ops.upsample2x_conv) when the channel count allows the TMA-store epilogue
                uw = g(f"{p}.upsamplers.0.conv.weight")
                blk["up"] = dict(w=R(packing.pack_conv_weight(uw)), b=R(g(f"{p}.upsamplers.0.conv.bias")), n=cur,
                                 w4=[R(t) for t in packing.pack_upsample_conv(uw)] if cur % 32 == 0 else None)
            self.up.append(blk)
        self.norm_out = dict(w=R(g("conv_norm_out.weight")), b=R(g("conv_norm_out.bias")))
        self.conv_out = dict(w=R(packing.pack_conv_weight(g("conv_out.weight"))), b=R(g("conv_out.bias")),
                             n=cfg["out_channels"])
        self.temb_all = dict(w=R(torch.cat(temb_w, 0)), b=R(torch.cat(temb_b, 0)))
        self.kv_all = dict(w=R(packing.pack_linear_weight(torch.cat(kv_w, 0)))) if kv_w else None
        del dt

    # ------------------------------------------------------------------ packed buffers -> reference parameters
    def reference_state_dict(self):
        """The reference's `state_dict()` (same names, shapes; values in the model dtype) rebuilt from the packed buffers:
        the exact inverse of `_build`, so `save_pretrained` writes a checkpoint the unmodified reference loads."""
        cfg = self.config
        spec = specs.unet2d_condition_params(dict(cfg))
        W = lambda n: self._buffers[n].detach().cpu()  # noqa: E731
        out = {}
        boc = tuple(cfg["block_out_channels"])
        cross = cfg["cross_attention_dim"]
        cross = cross[0] if isinstance(cross, (list, tuple)) else cross
        temb_w, temb_b = W(self.temb_all["w"]), W(self.temb_all["b"])
        kv = packing.unpack_linear_weight(W(self.kv_all["w"]), cross) if self.kv_all is not None else None

        def lin_small(p, d):
            out[p + ".weight"], out[p + ".bias"] = W(d["w"]), W(d["b"])

        lin_small("time_embedding.linear_1", self.time_embedding[0])
        lin_small("time_embedding.linear_2", self.time_embedding[1])
        if self.add_emb is not None:
            lin_small("add_embedding.linear_1", self.add_emb[0])
            lin_small("add_embedding.linear_2", self.add_emb[1])
        out["conv_in.weight"] = packing.unpack_conv_weight(W(self.conv_in["w"]), self.in_pad, 3)[:, :cfg["in_channels"]].contiguous()
        out["conv_in.bias"] = W(self.conv_in["b"])

        def resnet(p, r):
            cin, cout, split = r["cin"], r["cout"], r["split"]
            out[p + ".norm1.weight"], out[p + ".norm1.bias"] = W(r["n1w"]), W(r["n1b"])
            out[p + ".conv1.weight"] = packing.unpack_conv_weight(W(r["c1w"]), cin, 3, split)
            out[p + ".conv1.bias"] = W(r["c1b"])
            o = r["temb_off"]
            out[p + ".time_emb_proj.weight"], out[p + ".time_emb_proj.bias"] = temb_w[o:o + cout].contiguous(), temb_b[o:o + cout].contiguous()
            out[p + ".norm2.weight"], out[p + ".norm2.bias"] = W(r["n2w"]), W(r["n2b"])
            out[p + ".conv2.weight"] = packing.unpack_conv_weight(W(r["c2w"]), cout, 3)
            out[p + ".conv2.bias"] = W(r["c2b"])
            if "scw" in r:
                out[p + ".conv_shortcut.weight"] = packing.unpack_conv_weight(W(r["scw"]), cin, 1, split)
                out[p + ".conv_shortcut.bias"] = W(r["scb"])

        def transformer(p, t):
            ch = t["ch"]
            out[p + ".norm.weight"], out[p + ".norm.bias"] = W(t["nw"]), W(t["nb"])
            out[p + ".proj_in.weight"] = packing.unpack_linear_weight(W(t["piw"]), ch)   # reshaped to the spec below
            out[p + ".proj_in.bias"] = W(t["pib"])
            out[p + ".proj_out.weight"] = packing.unpack_linear_weight(W(t["pow"]), ch)
            out[p + ".proj_out.bias"] = W(t["pob"])
            for k, blk in enumerate(t["blocks"]):
                b = f"{p}.transformer_blocks.{k}"
                for nm, (w_, b_) in (("norm1", ("l1w", "l1b")), ("norm2", ("l2w", "l2b")), ("norm3", ("l3w", "l3b"))):
                    out[f"{b}.{nm}.weight"], out[f"{b}.{nm}.bias"] = W(blk[w_]), W(blk[b_])
                fold = self._fold_norms
                qkv = packing.unpack_linear_weight(W(blk["qkv"]), ch)
                if fold:
                    qkv = packing.unfold_layer_norm(qkv, W(blk["l1w"]), W(blk["qkv_sh"]))
                out[b + ".attn1.to_q.weight"], out[b + ".attn1.to_k.weight"], out[b + ".attn1.to_v.weight"] = (
                    qkv[:ch].contiguous(), qkv[ch:2 * ch].contiguous(), qkv[2 * ch:].contiguous())
                out[b + ".attn1.to_out.0.weight"], out[b + ".attn1.to_out.0.bias"] = packing.unpack_linear_weight(W(blk["ow"]), ch), W(blk["ob"])
                q2 = packing.unpack_linear_weight(W(blk["q2"]), ch)
                out[b + ".attn2.to_q.weight"] = packing.unfold_layer_norm(q2, W(blk["l2w"]), W(blk["q2_sh"])) if fold else q2
                o = blk["kv_off"]
                out[b + ".attn2.to_k.weight"], out[b + ".attn2.to_v.weight"] = kv[o:o + ch].contiguous(), kv[o + ch:o + 2 * ch].contiguous()
                out[b + ".attn2.to_out.0.weight"], out[b + ".attn2.to_out.0.bias"] = packing.unpack_linear_weight(W(blk["o2w"]), ch), W(blk["o2b"])
                ffw, ffb = packing.unpack_geglu(W(blk["ff1w"]), W(blk["ff1b"]), ch, blk["ff1tile"])
                if fold:
                    # the stored bias is b + W beta: b = bias - W beta with the recovered W (the row means `ff1_sh` are in the
                    # unpacked row order: they were taken before pack_geglu interleaved the rows)
                    ffw = packing.unfold_layer_norm(ffw, W(blk["l3w"]), W(blk["ff1_sh"]))
                    ffb = (ffb.to(torch.float32) - ffw.to(torch.float32) @ W(blk["l3b"]).to(torch.float32)).to(ffw.dtype)
                out[b + ".ff.net.0.proj.weight"], out[b + ".ff.net.0.proj.bias"] = ffw, ffb
                out[b + ".ff.net.2.weight"] = packing.unpack_linear_weight(W(blk["ff2w"]), blk["ff1n"] // 2)
                out[b + ".ff.net.2.bias"] = W(blk["ff2b"])

        def sampler(p, d, cin):
            out[p + ".conv.weight"], out[p + ".conv.bias"] = packing.unpack_conv_weight(W(d["w"]), cin, 3), W(d["b"])

        for i, blk in enumerate(self.down):
            for j, r in enumerate(blk["res"]):
                resnet(f"down_blocks.{i}.resnets.{j}", r)
            for j, t in enumerate(blk["attn"]):
                transformer(f"down_blocks.{i}.attentions.{j}", t)
            if blk["down"] is not None:
                sampler(f"down_blocks.{i}.downsamplers.0", blk["down"], blk["down"]["n"])
        resnet("mid_block.resnets.0", self.mid["res"][0])
        resnet("mid_block.resnets.1", self.mid["res"][1])
        transformer("mid_block.attentions.0", self.mid["attn"][0])
        for i, blk in enumerate(self.up):
            for j, r in enumerate(blk["res"]):
                resnet(f"up_blocks.{i}.resnets.{j}", r)
            for j, t in enumerate(blk["attn"]):
                transformer(f"up_blocks.{i}.attentions.{j}", t)
            if blk["up"] is not None:
                sampler(f"up_blocks.{i}.upsamplers.0", blk["up"], blk["up"]["n"])
        out["conv_norm_out.weight"], out["conv_norm_out.bias"] = W(self.norm_out["w"]), W(self.norm_out["b"])
        out["conv_out.weight"] = packing.unpack_conv_weight(W(self.conv_out["w"]), boc[0], 3)
        out["conv_out.bias"] = W(self.conv_out["b"])
        missing = [k for k in spec if k not in out]
        if missing or len(out) != len(spec):
            raise RuntimeError(f"reference_state_dict: {len(missing)} parameters not reconstructed, e.g. {missing[:3]}")
        return {k: out[k].reshape(spec[k]).contiguous() for k in spec}  # spec order; 1x1-conv projections get their 4-D shape back

    # ------------------------------------------------------------------ nn.Module-ish surface
    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._buffers["w0"].device

    def W(self, name):
        return self._buffers[name]

    @classmethod
    def from_state_dict(cls, config, state_dict, dtype=torch.bfloat16, device="cuda"):
        return cls(config, state_dict, dtype=dtype, device=device)

    @classmethod
    def random_init(cls, config=None, seed=0, dtype=torch.bfloat16, device="cuda"):
        cfg = dict(specs.SDXL_UNET_CONFIG)
        cfg.update(config or {})
        # weights drawn on the target device (seconds for 11.9 B parameters on a GPU); device="cpu" keeps the CPU stream
        sd = specs.random_state_dict(specs.unet2d_condition_params(cfg), seed=seed, dtype=dtype, device=device if torch.device(device).type == "cuda" else "cpu")
        return cls(cfg, sd, dtype=dtype, device=device)

    def _reset_stateful_cache(self):
        self._kv_key = None
        self._kv = None

    # ------------------------------------------------------------------ blocks
    def _resnet(self, r, x, x2, temb_all, B, H, W):
        hw = H * W
        G, eps = self.config["norm_num_groups"], self.config.get("norm_eps", 1e-5)
        n1 = ops.group_norm(x, x2=x2, batch=B, hw=hw, groups=G, eps=eps, gamma=self.W(r["n1w"]), beta=self.W(r["n1b"]), silu=True)
        tslice = temb_all[:, r["temb_off"]:r["temb_off"] + r["cout"]]
        h = ops.conv_gemm(n1, self.W(r["c1w"]), r["cout"], batch=B, H=H, W=W, ksize=3, bias=self.W(r["c1b"]),
                          rowvec=tslice, rows_per_group=hw)
        n2 = ops.group_norm(h, batch=B, hw=hw, groups=G, eps=eps, gamma=self.W(r["n2w"]), beta=self.W(r["n2b"]), silu=True)
        if "scw" in r:
            sc = ops.conv_gemm(x, self.W(r["scw"]), r["cout"], batch=B, H=H, W=W, ksize=1, x2=x2, bias=self.W(r["scb"]))
        else:
            sc = x
        return ops.conv_gemm(n2, self.W(r["c2w"]), r["cout"], batch=B, H=H, W=W, ksize=3, bias=self.W(r["c2b"]), residual=sc)

    def _transformer(self, t, x, kv, B, H, W, S_txt):
        hw = H * W
        C, nh = t["ch"], t["heads"]
        hd = C // nh
        n = ops.group_norm(x, batch=B, hw=hw, groups=self.config["norm_num_groups"], eps=1e-6, gamma=self.W(t["nw"]),
                           beta=self.W(t["nb"]), silu=False)
        if not self._fold_norms:
            h = ops.linear(n, self.W(t["piw"]), C, bias=self.W(t["pib"]))
            for blk in t["blocks"]:
                n = ops.layer_norm(h, eps=1e-5, gamma=self.W(blk["l1w"]), beta=self.W(blk["l1b"]))
                qkv = ops.linear(n, self.W(blk["qkv"]), 3 * C).view(B, hw, 3 * C)
                o = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads=nh, head_dim=hd)
                h = ops.linear(o.view(B * hw, C), self.W(blk["ow"]), C, bias=self.W(blk["ob"]), residual=h)
                n = ops.layer_norm(h, eps=1e-5, gamma=self.W(blk["l2w"]), beta=self.W(blk["l2b"]))
                q = ops.linear(n, self.W(blk["q2"]), C).view(B, hw, C)
                kvb = kv[:, :, blk["kv_off"]:blk["kv_off"] + 2 * C]
                o = ops.attention(q, kvb[:, :, :C], kvb[:, :, C:], heads=nh, head_dim=hd)
                h = ops.linear(o.view(B * hw, C), self.W(blk["o2w"]), C, bias=self.W(blk["o2b"]), residual=h)
                n = ops.layer_norm(h, eps=1e-5, gamma=self.W(blk["l3w"]), beta=self.W(blk["l3b"]))
                gg = ops.linear(n, self.W(blk["ff1w"]), blk["ff1n"], bias=self.W(blk["ff1b"]), geglu=True, tile_n=blk["ff1tile"])
                h = ops.linear(gg, self.W(blk["ff2w"]), C, bias=self.W(blk["ff2b"]), residual=h)
            return ops.linear(h, self.W(t["pow"]), C, bias=self.W(t["pob"]), residual=x)
        # folded LayerNorms: every GEMM that writes the residual stream h also writes its per-row (sum, sum of squares); the
        # GEMM that would read LN(h) reads h itself and normalises in its epilogue
        FL = ops.FoldedLayerNorm
        h, st = ops.linear(n, self.W(t["piw"]), C, bias=self.W(t["pib"]), row_stats=True)
        last = len(t["blocks"]) - 1
        for i, blk in enumerate(t["blocks"]):
            qkv = ops.linear(h, self.W(blk["qkv"]), 3 * C, bias=self.W(blk["qkv_b"]), ln=FL(st, 1e-5)).view(B, hw, 3 * C)
            o = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads=nh, head_dim=hd)
            h, st = ops.linear(o.view(B * hw, C), self.W(blk["ow"]), C, bias=self.W(blk["ob"]), residual=h, row_stats=True)
            q = ops.linear(h, self.W(blk["q2"]), C, bias=self.W(blk["q2_b"]), ln=FL(st, 1e-5)).view(B, hw, C)
            kvb = kv[:, :, blk["kv_off"]:blk["kv_off"] + 2 * C]
            o = ops.attention(q, kvb[:, :, :C], kvb[:, :, C:], heads=nh, head_dim=hd)
            h, st = ops.linear(o.view(B * hw, C), self.W(blk["o2w"]), C, bias=self.W(blk["o2b"]), residual=h, row_stats=True)
            gg = ops.linear(h, self.W(blk["ff1w"]), blk["ff1n"], bias=self.W(blk["ff1b"]), geglu=True, tile_n=blk["ff1tile"], ln=FL(st, 1e-5))
            if i == last:
                h = ops.linear(gg, self.W(blk["ff2w"]), C, bias=self.W(blk["ff2b"]), residual=h)
            else:
                h, st = ops.linear(gg, self.W(blk["ff2w"]), C, bias=self.W(blk["ff2b"]), residual=h, row_stats=True)
        return ops.linear(h, self.W(t["pow"]), C, bias=self.W(t["pob"]), residual=x)

    def _text_kv(self, ehs):
        """K/V projections of the text states for every cross-attention layer: one GEMM, cached per prompt
        (they do not depend on the timestep; the reference recomputes them 140x per step)."""
        if self.kv_all is None:
            return None
        # cached on the tensor OBJECT (kept alive here) and its version counter: an address-based key would hand a new
        # prompt the previous prompt's K/V whenever the allocator reuses the freed buffer
        src = self._kv_key
        if src is None or src[0] is not ehs or src[1] != ehs._version or self._kv is None:
            B, S, D = ehs.shape
            flat = ehs.to(self._dtype).contiguous().view(B * S, D)
            kv = ops.linear(flat, self.W(self.kv_all["w"]), self._kv_total)
            self._kv = kv.view(B, S, self._kv_total)
            self._kv_key = (ehs, ehs._version)
        return self._kv

    # ------------------------------------------------------------------ forward
    def _embeddings(self, B, timestep, added_cond_kwargs, device):
        cfg = self.config
        dt = self._dtype
        if not torch.is_tensor(timestep):
            t = torch.full((B,), float(timestep), dtype=torch.float32, device=device)
        else:
            t = timestep.to(device=device, dtype=torch.float32).reshape(-1).expand(B).contiguous()
        t_emb = ops.timestep_embedding(t, self.time_dim, dtype=dt, flip_sin_to_cos=cfg.get("flip_sin_to_cos", True),
                                       downscale_freq_shift=float(cfg.get("freq_shift", 0)))
        te = self.time_embedding
        e = ops.small_linear(t_emb, self.W(te[0]["w"]), bias=self.W(te[0]["b"]), act_out=ACT_SILU)
        if self.add_emb is None:
            return ops.small_linear(e, self.W(te[1]["w"]), bias=self.W(te[1]["b"]))
        emb = ops.small_linear(e, self.W(te[1]["w"]), bias=self.W(te[1]["b"]))
        if added_cond_kwargs is None or "text_embeds" not in added_cond_kwargs or "time_ids" not in added_cond_kwargs:
            raise ValueError("addition_embed_type='text_time' requires added_cond_kwargs with `text_embeds` and `time_ids`")
        text_embeds = added_cond_kwargs["text_embeds"]
        time_ids = added_cond_kwargs["time_ids"].to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        ad = cfg["addition_time_embed_dim"]
        add_in = torch.empty((B, text_embeds.shape[-1] + ad * (time_ids.numel() // B)), dtype=dt, device=device)
        add_in[:, :text_embeds.shape[-1]] = text_embeds
        tid = ops.timestep_embedding(time_ids, ad, dtype=dt, flip_sin_to_cos=cfg.get("flip_sin_to_cos", True),
                                     downscale_freq_shift=float(cfg.get("freq_shift", 0)))
        add_in[:, text_embeds.shape[-1]:] = tid.view(B, -1)
        ae = self.add_emb
        a1 = ops.small_linear(add_in, self.W(ae[0]["w"]), bias=self.W(ae[0]["b"]), act_out=ACT_SILU)
        # emb = emb + aug_emb  (aug rounded to 16 bit first, as in the reference)
        return ops.small_linear(a1, self.W(ae[1]["w"]), bias=self.W(ae[1]["b"]), addend=emb)

    @ops.prefetching_forward
    def _forward_nhwc(self, x_in, B, H, W, timestep, kv, added_cond_kwargs):
        """x_in: [B*H*W, in_pad] NHWC; kv: cached text K/V [B, S_txt, kv_total].  Returns the NHWC prediction
        [B*H*W, out_channels]."""
        dev = x_in.device
        emb = self._embeddings(B, timestep, added_cond_kwargs, dev)
        temb_all = ops.small_linear(emb, self.W(self.temb_all["w"]), bias=self.W(self.temb_all["b"]), act_in=ACT_SILU)
        S_txt = kv.shape[1] if kv is not None else 0

        x = ops.conv_gemm(x_in, self.W(self.conv_in["w"]), self.conv_in["n"], batch=B, H=H, W=W, ksize=3,
                          bias=self.W(self.conv_in["b"]))
        skips = [(x, H, W)]
        for blk in self.down:
            for j, r in enumerate(blk["res"]):
                x = self._resnet(r, x, None, temb_all, B, H, W)
                if blk["attn"]:
                    x = self._transformer(blk["attn"][j], x, kv, B, H, W, S_txt)
                skips.append((x, H, W))
            if blk["down"] is not None:
                d = blk["down"]
                x = ops.conv_gemm(x, self.W(d["w"]), d["n"], batch=B, H=H, W=W, ksize=3, stride=2, bias=self.W(d["b"]))
                H, W = H // 2, W // 2
                skips.append((x, H, W))
        x = self._resnet(self.mid["res"][0], x, None, temb_all, B, H, W)
        x = self._transformer(self.mid["attn"][0], x, kv, B, H, W, S_txt)
        x = self._resnet(self.mid["res"][1], x, None, temb_all, B, H, W)
        for blk in self.up:
            for j, r in enumerate(blk["res"]):
                sk, sh, sw = skips.pop()
                assert (sh, sw) == (H, W)
                x = self._resnet(r, x, sk, temb_all, B, H, W)
                if blk["attn"]:
                    x = self._transformer(blk["attn"][j], x, kv, B, H, W, S_txt)
            if blk["up"] is not None:
                u = blk["up"]
                if u["w4"] is not None:
                    x = ops.upsample2x_conv(x, [self.W(t) for t in u["w4"]], u["n"], batch=B, H=H, W=W, bias=self.W(u["b"]))
                    H, W = 2 * H, 2 * W
                else:
                    xu = ops.upsample_nearest2x(x, batch=B, H=H, W=W)
                    H, W = 2 * H, 2 * W
                    x = ops.conv_gemm(xu, self.W(u["w"]), u["n"], batch=B, H=H, W=W, ksize=3, bias=self.W(u["b"]))
        n = ops.group_norm(x, batch=B, hw=H * W, groups=self.config["norm_num_groups"], eps=self.config.get("norm_eps", 1e-5),
                           gamma=self.W(self.norm_out["w"]), beta=self.W(self.norm_out["b"]), silu=True)
        return ops.conv_gemm(n, self.W(self.conv_out["w"]), self.conv_out["n"], batch=B, H=H, W=W, ksize=3,
                             bias=self.W(self.conv_out["b"]))

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None, attention_mask=None,
                cross_attention_kwargs=None, added_cond_kwargs=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, down_intrablock_additional_residuals=None,
                encoder_attention_mask=None, return_dict=True):
        for name, v in (("class_labels", class_labels), ("timestep_cond", timestep_cond), ("attention_mask", attention_mask),
                        ("down_block_additional_residuals", down_block_additional_residuals),
                        ("mid_block_additional_residual", mid_block_additional_residual),
                        ("down_intrablock_additional_residuals", down_intrablock_additional_residuals),
                        ("encoder_attention_mask", encoder_attention_mask)):
            if v is not None:
                raise NotImplementedError(f"{name} is outside the accelerated hot path (SURVEY.md §8)")
        if not sample.is_cuda:
            raise ops.B200Error("UNet2DConditionModel (B200) needs CUDA tensors: there is no CPU fallback")
        B, C, H, W = sample.shape
        if H % (2 ** (len(self.config["block_out_channels"]) - 1)) or W % (2 ** (len(self.config["block_out_channels"]) - 1)):
            raise ValueError("sample height/width must be divisible by the total downsampling factor")
        sample = sample.to(self._dtype)
        if self.use_cuda_graph:
            out = self._forward_graph(sample, timestep, encoder_hidden_states, added_cond_kwargs)
        else:
            x_in = ops.nchw_to_nhwc(sample, c_pad=self.in_pad)
            y = self._forward_nhwc(x_in, B, H, W, timestep, self._text_kv(encoder_hidden_states), added_cond_kwargs)
            out = ops.nhwc_to_nchw(y, batch=B, C_out=self.config["out_channels"], H=H, W=W)
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(out)

    # ------------------------------------------------------------------ CUDA graph replay of a whole forward
    def enable_cuda_graph(self, enabled=True):
        self.use_cuda_graph = enabled
        if not enabled:
            self._graphs.clear()

    def _forward_graph(self, sample, timestep, ehs, added):
        B, C, H, W = sample.shape
        key = (B, C, H, W, tuple(ehs.shape), tuple(added["text_embeds"].shape) if added else None,
               tuple(added["time_ids"].shape) if added else None)
        st = self._graphs.get(key)
        dev = sample.device
        if st is None:
            st = dict(sample=torch.empty_like(sample), t=torch.empty((), dtype=torch.float32, device=dev),
                      kv=torch.empty((ehs.shape[0], ehs.shape[1], self._kv_total), dtype=self._dtype, device=dev),
                      kv_src=None,
                      added=None if not added else dict(text_embeds=torch.empty_like(added["text_embeds"], dtype=self._dtype),
                                                        time_ids=torch.empty_like(added["time_ids"])))
            self._graphs[key] = st

        def load():
            st["sample"].copy_(sample)
            st["t"].copy_(timestep if torch.is_tensor(timestep) else torch.tensor(float(timestep)))
            kv = self._text_kv(ehs)  # eager, cached per prompt: stays outside the per-step graph
            if st["kv_src"] is not kv:
                st["kv"].copy_(kv)
                st["kv_src"] = kv
            if added:
                st["added"]["text_embeds"].copy_(added["text_embeds"])
                st["added"]["time_ids"].copy_(added["time_ids"])

        def run():
            x_in = ops.nchw_to_nhwc(st["sample"], c_pad=self.in_pad)
            y = self._forward_nhwc(x_in, B, H, W, st["t"], st["kv"], st["added"])
            return ops.nhwc_to_nchw(y, batch=B, C_out=self.config["out_channels"], H=H, W=W)

        load()
        if "graph" not in st:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                run()  # warm-up: workspaces, smem attributes, lazy module load
            torch.cuda.current_stream().wait_stream(s)
            gph = torch.cuda.CUDAGraph()
            n0 = ops.launches()
            with torch.cuda.graph(gph):
                st["out"] = run()
            st["graph"] = gph
            st["launches"] = ops.launches() - n0
        st["graph"].replay()
        ops._count(st["launches"])
        return st["out"].clone()
