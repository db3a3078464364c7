"""Drop-in UNet2DModel (unconditional, DDPM family - BASELINE.json config 0) on libb200diff.so.

`forward(sample, timestep, return_dict) -> .sample`, `.config`, `.dtype`, `.device` as the reference
(models/unets/unet_2d.py:39,249).  Blocks: DownBlock2D / AttnDownBlock2D / UNetMidBlock2D / AttnUpBlock2D / UpBlock2D
(unet_2d_blocks.py:1294,1018,589,2185,2474); the legacy `Attention` inside them (group_norm, biased q/k/v, residual,
attention_processor.py:2725-2789) runs as GroupNorm kernel -> fused-QKV GEMM -> tcgen05 attention -> out-proj GEMM with
the residual in its epilogue.  attention_head_dim must be 64 or 128 (the fused kernel's head sizes).
"""
import torch

from . import ops, packing, specs
from .checkpoint import FromPretrainedMixin
from .config import FrozenConfig
from .ops import ACT_SILU


class UNet2DOutput:
    def __init__(self, sample):
        self.sample = sample


class UNet2DModel(torch.nn.Module, FromPretrainedMixin):
    _ref_class_names = ("UNet2DModel",)

    @classmethod
    def _param_spec(cls, cfg):
        full = dict(specs.DDPM_TINY_CONFIG)
        full.update(cfg)
        return specs.unet2d_params(full)

    def __init__(self, config, state_dict, dtype=torch.bfloat16, device="cuda"):
        super().__init__()
        cfg = dict(specs.DDPM_TINY_CONFIG)
        cfg.update(config)
        self.config = FrozenConfig(cfg)
        self._dtype = dtype
        self._n = 0
        if cfg.get("time_embedding_type", "positional") != "positional":
            raise NotImplementedError("only positional time embeddings")
        # reference options (models/unets/unet_2d.py:95-125) whose non-default values are not implemented: refuse, never ignore
        only = dict(act_fn=("silu", "swish"), mid_block_type=("UNetMidBlock2D",), downsample_type=("conv",), upsample_type=("conv",),
                    resnet_time_scale_shift=("default",), center_input_sample=(False,), attn_norm_num_groups=(None,),
                    class_embed_type=(None,), num_class_embeds=(None,), time_embedding_dim=(None,), downsample_padding=(1,),
                    mid_block_scale_factor=(1, 1.0))
        for k, allowed in only.items():
            if k in cfg and cfg[k] not in allowed:
                raise NotImplementedError(f"UNet2DModel option {k}={cfg[k]!r} is outside the accelerated hot path (supported: {allowed})")
        for t in tuple(cfg["down_block_types"]) + tuple(cfg["up_block_types"]):
            if t not in ("DownBlock2D", "AttnDownBlock2D", "UpBlock2D", "AttnUpBlock2D"):
                raise NotImplementedError(f"block type {t} is outside the accelerated hot path")
        hd = cfg.get("attention_head_dim", 8)
        if any("Attn" in t for t in tuple(cfg["down_block_types"]) + tuple(cfg["up_block_types"])) or cfg.get("add_attention", True):
            if hd not in (64, 128):
                raise NotImplementedError(f"attention_head_dim={hd}: the tcgen05 attention kernel supports 64 and 128")
        spec = specs.unet2d_params(cfg)
        for k, shp in spec.items():
            if k not in state_dict or tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"state_dict entry {k} missing or wrong shape (expected {tuple(shp)})")
        self._build(state_dict, torch.device(device))

    def _reg(self, t, device):
        name = f"w{self._n}"
        self._n += 1
        self.register_buffer(name, t.to(device=device, dtype=self._dtype).contiguous(), persistent=False)
        return name

    def W(self, name):
        return self._buffers[name]

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._buffers["w0"].device

    @classmethod
    def random_init(cls, config=None, seed=0, dtype=torch.bfloat16, device="cuda"):
        cfg = dict(specs.DDPM_TINY_CONFIG)
        cfg.update(config or {})
        return cls(cfg, specs.random_state_dict(specs.unet2d_params(cfg), seed=seed, dtype=dtype), dtype=dtype, device=device)

    def _build(self, sd, device):
        cfg = self.config
        R = lambda t: self._reg(t, device)  # noqa: E731
        g = lambda k: sd[k].to(torch.float32)  # noqa: E731
        boc = tuple(cfg["block_out_channels"])
        self.in_pad = packing.rup(cfg["in_channels"], 8)
        self.time_dim = boc[0]
        self.t1 = dict(w=R(g("time_embedding.linear_1.weight")), b=R(g("time_embedding.linear_1.bias")))
        self.t2 = dict(w=R(g("time_embedding.linear_2.weight")), b=R(g("time_embedding.linear_2.bias")))
        w_in = torch.nn.functional.pad(g("conv_in.weight"), (0, 0, 0, 0, 0, self.in_pad - cfg["in_channels"]))
        self.conv_in = dict(w=R(packing.pack_conv_weight(w_in)), b=R(g("conv_in.bias")), n=boc[0])
        temb_w, temb_b = [], []
        self._temb_total = 0

        def resnet(p, split=None):
            w1 = g(p + ".conv1.weight")
            r = dict(cout=w1.shape[0], n1w=R(g(p + ".norm1.weight")), n1b=R(g(p + ".norm1.bias")),
                     c1w=R(packing.pack_conv_weight(w1, split)), c1b=R(g(p + ".conv1.bias")),
                     n2w=R(g(p + ".norm2.weight")), n2b=R(g(p + ".norm2.bias")),
                     c2w=R(packing.pack_conv_weight(g(p + ".conv2.weight"))), c2b=R(g(p + ".conv2.bias")), temb_off=self._temb_total)
            temb_w.append(g(p + ".time_emb_proj.weight"))
            temb_b.append(g(p + ".time_emb_proj.bias"))
            self._temb_total += w1.shape[0]
            if (p + ".conv_shortcut.weight") in sd:
                r["scw"] = R(packing.pack_conv_weight(g(p + ".conv_shortcut.weight"), split))
                r["scb"] = R(g(p + ".conv_shortcut.bias"))
            elif split is not None:
                raise NotImplementedError("two-source resnet without conv_shortcut")
            return r

        def attn(p):
            C = sd[p + ".to_q.weight"].shape[0]
            return dict(C=C, gw=R(g(p + ".group_norm.weight")), gb=R(g(p + ".group_norm.bias")),
                        qkv=R(packing.pack_linear_weight(torch.cat([g(p + ".to_q.weight"), g(p + ".to_k.weight"), g(p + ".to_v.weight")], 0))),
                        qkvb=R(torch.cat([g(p + ".to_q.bias"), g(p + ".to_k.bias"), g(p + ".to_v.bias")], 0)),
                        ow=R(packing.pack_linear_weight(g(p + ".to_out.0.weight"))), ob=R(g(p + ".to_out.0.bias")))

        def conv(p):
            w = g(p + ".weight")
            return dict(w=R(packing.pack_conv_weight(w)), b=R(g(p + ".bias")), n=w.shape[0])

        self.down, skip_ch, cur = [], [boc[0]], boc[0]
        n = len(boc)
        lpb = cfg.get("layers_per_block", 2)
        for i, bt in enumerate(cfg["down_block_types"]):
            p = f"down_blocks.{i}"
            blk = dict(res=[], attn=[], down=None)
            for j in range(lpb):
                blk["res"].append(resnet(f"{p}.resnets.{j}"))
                cur = boc[i]
                if bt == "AttnDownBlock2D":
                    blk["attn"].append(attn(f"{p}.attentions.{j}"))
                skip_ch.append(cur)
            if i != n - 1:
                blk["down"] = conv(f"{p}.downsamplers.0.conv")
                skip_ch.append(cur)
            self.down.append(blk)
        self.mid = dict(res=[resnet("mid_block.resnets.0"), resnet("mid_block.resnets.1")],
                        attn=attn("mid_block.attentions.0") if "mid_block.attentions.0.to_q.weight" in sd else None)
        self.up = []
        rboc = boc[::-1]
        for i, bt in enumerate(cfg["up_block_types"]):
            p = f"up_blocks.{i}"
            blk = dict(res=[], attn=[], up=None)
            for j in range(lpb + 1):
                sk = skip_ch.pop()
                blk["res"].append(resnet(f"{p}.resnets.{j}", split=(cur, sk)))
                cur = rboc[i]
                if bt == "AttnUpBlock2D":
                    blk["attn"].append(attn(f"{p}.attentions.{j}"))
            if i != n - 1:
                blk["up"] = conv(f"{p}.upsamplers.0.conv")
            self.up.append(blk)
        self.norm_out = dict(w=R(g("conv_norm_out.weight")), b=R(g("conv_norm_out.bias")))
        self.conv_out = conv("conv_out")
        self.temb_all = dict(w=R(torch.cat(temb_w, 0)), b=R(torch.cat(temb_b, 0)))

    def reference_state_dict(self):
        """The reference's `state_dict()` rebuilt from the packed buffers (exact inverse of `_build`; the channel bookkeeping
        mirrors it)."""
        cfg = self.config
        spec = specs.unet2d_params(dict(cfg))
        W = lambda n: self._buffers[n].detach().cpu()  # noqa: E731
        out = {}
        boc = tuple(cfg["block_out_channels"])
        temb_w, temb_b = W(self.temb_all["w"]), W(self.temb_all["b"])
        out["time_embedding.linear_1.weight"], out["time_embedding.linear_1.bias"] = W(self.t1["w"]), W(self.t1["b"])
        out["time_embedding.linear_2.weight"], out["time_embedding.linear_2.bias"] = W(self.t2["w"]), W(self.t2["b"])
        out["conv_in.weight"] = packing.unpack_conv_weight(W(self.conv_in["w"]), self.in_pad, 3)[:, :cfg["in_channels"]].contiguous()
        out["conv_in.bias"] = W(self.conv_in["b"])

        def resnet(p, r, cin, split=None):
            cout, o = r["cout"], r["temb_off"]
            out[p + ".norm1.weight"], out[p + ".norm1.bias"] = W(r["n1w"]), W(r["n1b"])
            out[p + ".conv1.weight"], out[p + ".conv1.bias"] = packing.unpack_conv_weight(W(r["c1w"]), cin, 3, split), W(r["c1b"])
            out[p + ".time_emb_proj.weight"], out[p + ".time_emb_proj.bias"] = temb_w[o:o + cout].contiguous(), temb_b[o:o + cout].contiguous()
            out[p + ".norm2.weight"], out[p + ".norm2.bias"] = W(r["n2w"]), W(r["n2b"])
            out[p + ".conv2.weight"], out[p + ".conv2.bias"] = packing.unpack_conv_weight(W(r["c2w"]), cout, 3), W(r["c2b"])
            if "scw" in r:
                out[p + ".conv_shortcut.weight"] = packing.unpack_conv_weight(W(r["scw"]), cin, 1, split)
                out[p + ".conv_shortcut.bias"] = W(r["scb"])
            return cout

        def attn(p, a):
            C = a["C"]
            out[p + ".group_norm.weight"], out[p + ".group_norm.bias"] = W(a["gw"]), W(a["gb"])
            qkv, qkvb = packing.unpack_linear_weight(W(a["qkv"]), C), W(a["qkvb"])
            for i, nm in enumerate(("to_q", "to_k", "to_v")):
                out[f"{p}.{nm}.weight"], out[f"{p}.{nm}.bias"] = qkv[i * C:(i + 1) * C].contiguous(), qkvb[i * C:(i + 1) * C].contiguous()
            out[p + ".to_out.0.weight"], out[p + ".to_out.0.bias"] = packing.unpack_linear_weight(W(a["ow"]), C), W(a["ob"])

        def conv(p, c, cin):
            out[p + ".weight"], out[p + ".bias"] = packing.unpack_conv_weight(W(c["w"]), cin, 3), W(c["b"])

        skip_ch, cur = [boc[0]], boc[0]
        for i, blk in enumerate(self.down):
            for j, r in enumerate(blk["res"]):
                cur = resnet(f"down_blocks.{i}.resnets.{j}", r, cur)
                skip_ch.append(cur)
            for j, a in enumerate(blk["attn"]):
                attn(f"down_blocks.{i}.attentions.{j}", a)
            if blk["down"] is not None:
                conv(f"down_blocks.{i}.downsamplers.0.conv", blk["down"], cur)
                skip_ch.append(cur)
        cur = resnet("mid_block.resnets.0", self.mid["res"][0], cur)
        if self.mid["attn"] is not None:
            attn("mid_block.attentions.0", self.mid["attn"])
        cur = resnet("mid_block.resnets.1", self.mid["res"][1], cur)
        for i, blk in enumerate(self.up):
            for j, r in enumerate(blk["res"]):
                sk = skip_ch.pop()
                cur = resnet(f"up_blocks.{i}.resnets.{j}", r, cur + sk, split=(cur, sk))
            for j, a in enumerate(blk["attn"]):
                attn(f"up_blocks.{i}.attentions.{j}", a)
            if blk["up"] is not None:
                conv(f"up_blocks.{i}.upsamplers.0.conv", blk["up"], cur)
        out["conv_norm_out.weight"], out["conv_norm_out.bias"] = W(self.norm_out["w"]), W(self.norm_out["b"])
        conv("conv_out", self.conv_out, boc[0])
        missing = [k for k in spec if k not in out]
        if missing or len(out) != len(spec):
            raise RuntimeError(f"reference_state_dict: {len(missing)} parameters not reconstructed, e.g. {missing[:3]}")
        return {k: out[k].reshape(spec[k]).contiguous() for k in spec}

    # ------------------------------------------------------------------
    def _gn(self, x, w, b, B, hw, silu, x2=None):
        return ops.group_norm(x, x2=x2, batch=B, hw=hw, groups=self.config["norm_num_groups"], eps=self.config.get("norm_eps", 1e-5),
                              gamma=self.W(w), beta=self.W(b), silu=silu)

    def _resnet(self, r, x, x2, temb_all, B, H, W):
        n1 = self._gn(x, r["n1w"], r["n1b"], B, H * W, True, x2)
        h = ops.conv_gemm(n1, self.W(r["c1w"]), r["cout"], batch=B, H=H, W=W, ksize=3, bias=self.W(r["c1b"]),
                          rowvec=temb_all[:, r["temb_off"]:r["temb_off"] + r["cout"]], rows_per_group=H * W)
        n2 = self._gn(h, r["n2w"], r["n2b"], B, H * W, True)
        sc = ops.conv_gemm(x, self.W(r["scw"]), r["cout"], batch=B, H=H, W=W, ksize=1, x2=x2, bias=self.W(r["scb"])) if "scw" in r else x
        return ops.conv_gemm(n2, self.W(r["c2w"]), r["cout"], batch=B, H=H, W=W, ksize=3, bias=self.W(r["c2b"]), residual=sc)

    def _attn(self, a, x, B, H, W):
        hw, C = H * W, a["C"]
        hd = self.config["attention_head_dim"]
        n = self._gn(x, a["gw"], a["gb"], B, hw, False)
        qkv = ops.linear(n, self.W(a["qkv"]), 3 * C, bias=self.W(a["qkvb"])).view(B, hw, 3 * C)
        o = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads=C // hd, head_dim=hd)
        return ops.linear(o.view(B * hw, C), self.W(a["ow"]), C, bias=self.W(a["ob"]), residual=x)

    @torch.no_grad()
    @ops.prefetching_forward
    def forward(self, sample, timestep, class_labels=None, return_dict=True):
        if class_labels is not None:
            raise NotImplementedError("class conditioning is outside the hot path")
        if not sample.is_cuda:
            raise ops.B200Error("UNet2DModel (B200) needs CUDA tensors: there is no CPU fallback")
        B, C, H, W = sample.shape
        dev = sample.device
        x_in = ops.nchw_to_nhwc(sample.to(self._dtype), c_pad=self.in_pad)
        if not torch.is_tensor(timestep):
            t = torch.full((B,), float(timestep), dtype=torch.float32, device=dev)
        else:
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1).expand(B).contiguous()
        cfg = self.config
        t_emb = ops.timestep_embedding(t, self.time_dim, dtype=self._dtype, flip_sin_to_cos=cfg.get("flip_sin_to_cos", True),
                                       downscale_freq_shift=float(cfg.get("freq_shift", 0)))
        e = ops.small_linear(t_emb, self.W(self.t1["w"]), bias=self.W(self.t1["b"]), act_out=ACT_SILU)
        emb = ops.small_linear(e, self.W(self.t2["w"]), bias=self.W(self.t2["b"]))
        temb_all = ops.small_linear(emb, self.W(self.temb_all["w"]), bias=self.W(self.temb_all["b"]), act_in=ACT_SILU)
        x = ops.conv_gemm(x_in, self.W(self.conv_in["w"]), self.conv_in["n"], batch=B, H=H, W=W, ksize=3, bias=self.W(self.conv_in["b"]))
        skips = [x]
        for blk in self.down:
            for j, r in enumerate(blk["res"]):
                x = self._resnet(r, x, None, temb_all, B, H, W)
                if blk["attn"]:
                    x = self._attn(blk["attn"][j], x, B, H, W)
                skips.append(x)
            if blk["down"] is not None:
                d = blk["down"]
                x = ops.conv_gemm(x, self.W(d["w"]), d["n"], batch=B, H=H, W=W, ksize=3, stride=2, bias=self.W(d["b"]))
                H, W = H // 2, W // 2
                skips.append(x)
        x = self._resnet(self.mid["res"][0], x, None, temb_all, B, H, W)
        if self.mid["attn"] is not None:
            x = self._attn(self.mid["attn"], x, B, H, W)
        x = self._resnet(self.mid["res"][1], x, None, temb_all, B, H, W)
        for blk in self.up:
            for j, r in enumerate(blk["res"]):
                x = self._resnet(r, x, skips.pop(), temb_all, B, H, W)
                if blk["attn"]:
                    x = self._attn(blk["attn"][j], x, B, H, W)
            if blk["up"] is not None:
                u = blk["up"]
                xu = ops.upsample_nearest2x(x, batch=B, H=H, W=W)
                H, W = 2 * H, 2 * W
                x = ops.conv_gemm(xu, self.W(u["w"]), u["n"], batch=B, H=H, W=W, ksize=3, bias=self.W(u["b"]))
        n = self._gn(x, self.norm_out["w"], self.norm_out["b"], B, H * W, True)
        y = ops.conv_gemm(n, self.W(self.conv_out["w"]), self.conv_out["n"], batch=B, H=H, W=W, ksize=3, bias=self.W(self.conv_out["b"]))
        out = ops.nhwc_to_nchw(y, batch=B, C_out=cfg["out_channels"], H=H, W=W)
        if not return_dict:
            return (out,)
        return UNet2DOutput(out)
