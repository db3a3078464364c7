"""Drop-in AutoencoderKL on libb200diff.so.

`decode(z, return_dict=False) -> (image,)`, `.config` (block_out_channels, scaling_factor, shift_factor,
force_upcast, latents_mean/std), `.dtype`, `.post_quant_conv` - what the SDXL / Flux pipelines touch
(pipeline_stable_diffusion_xl.py:1262-1287, pipeline_flux.py:959-961).  Reference: models/autoencoders/
autoencoder_kl.py:199-233, autoencoders/vae.py:180-310.
`encode(x).latent_dist` (img2img / inpaint: autoencoder_kl.py:158-197, Encoder autoencoders/vae.py:59-177,
DiagonalGaussianDistribution :687) when the state_dict carries the encoder half.  Tiling / slicing are not built.

NHWC activations; GroupNorm+SiLU is one fused pass feeding the implicit-GEMM conv; the single head_dim-512
attention of the mid block runs unfused (see ops.attention_unfused).
"""
import types

import torch

from . import ops, packing, specs
from .checkpoint import FromPretrainedMixin
from .config import FrozenConfig


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class AutoencoderKLOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist


class DiagonalGaussianDistribution:
    """autoencoders/vae.py:687-735.  The eight-channel moments come from the CUDA encoder; what is left is a clamp, two
    exponentials and one multiply-add on a latent-sized tensor, written with the reference's own op sequence and dtype so
    that `sample(generator)` consumes the caller's RNG stream and rounds exactly like the reference."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean, device=self.parameters.device, dtype=self.parameters.dtype)

    def sample(self, generator=None):
        from .pipelines import randn_tensor
        noise = randn_tensor(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.0])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0 - self.logvar + other.logvar,
                               dim=[1, 2, 3])


class AutoencoderKL(torch.nn.Module, FromPretrainedMixin):
    _ref_class_names = ("AutoencoderKL",)

    @classmethod
    def _param_spec(cls, cfg):
        full = dict(specs.SDXL_VAE_CONFIG)
        full.update(cfg)
        return specs.vae_decoder_params(full)  # what every pipeline needs; the encoder half is optional (below)

    @classmethod
    def _optional_param_spec(cls, cfg):
        full = dict(specs.SDXL_VAE_CONFIG)
        full.update(cfg)
        return specs.vae_encoder_params(full)

    @staticmethod
    def _fix_keys(sd):
        from .checkpoint import convert_deprecated_attention_keys
        return convert_deprecated_attention_keys(sd)

    def __init__(self, config, state_dict, dtype=torch.bfloat16, device="cuda"):
        super().__init__()
        cfg = dict(specs.SDXL_VAE_CONFIG)
        cfg.update(config)
        self.config = FrozenConfig(cfg)
        # reference options (models/autoencoders/autoencoder_kl.py:75-100) this decoder does not implement must fail loudly
        if cfg.get("act_fn", "silu") not in ("silu", "swish"):
            raise NotImplementedError(f"AutoencoderKL act_fn={cfg['act_fn']!r} is outside the accelerated hot path")
        for t in tuple(cfg["up_block_types"]):
            if t != "UpDecoderBlock2D":
                raise NotImplementedError(f"decoder block type {t} is outside the accelerated hot path")
        if cfg.get("norm_num_groups") is None:
            raise NotImplementedError("norm_num_groups=None is outside the hot path")
        self._dtype = dtype
        self._n = 0
        spec = specs.vae_decoder_params(cfg)
        for k, shp in spec.items():
            if k not in state_dict:
                raise ValueError(f"state_dict is missing {k}")
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: expected shape {tuple(shp)}, got {tuple(state_dict[k].shape)}")
        enc_spec = specs.vae_encoder_params(cfg)
        have = [k for k in enc_spec if k in state_dict]
        if have and len(have) != len(enc_spec):
            raise ValueError(f"state_dict holds only {len(have)} of the {len(enc_spec)} encoder tensors, e.g. missing "
                             f"{[k for k in enc_spec if k not in state_dict][:3]}")
        for k in have:
            if tuple(state_dict[k].shape) != tuple(enc_spec[k]):
                raise ValueError(f"{k}: expected shape {tuple(enc_spec[k])}, got {tuple(state_dict[k].shape)}")
        self._build(state_dict, torch.device(device))
        self.enc = self._build_encoder(state_dict, torch.device(device)) if have else None
        self.use_slicing = False
        self.use_tiling = False
        # the SDXL pipeline only reads post_quant_conv.parameters() to pick a dtype when up-casting fp16 VAEs
        self.post_quant_conv = types.SimpleNamespace(parameters=lambda: iter([self._buffers["w0"]]))

    def _reg(self, t, device):
        name = f"w{self._n}"
        self._n += 1
        self.register_buffer(name, t.to(device=device, dtype=self._dtype).contiguous(), persistent=False)
        return name

    def W(self, name):
        return self._buffers[name]

    def _build(self, sd, device):
        cfg = self.config
        R = lambda t: self._reg(t, device)  # noqa: E731
        g = lambda k: sd[k].to(torch.float32)  # noqa: E731
        lc = cfg["latent_channels"]
        self.lat_pad = packing.rup(lc, 8)

        def conv(p, pad_in=None, ks=None):
            w = g(p + ".weight")
            if pad_in is not None and w.shape[1] < pad_in:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, pad_in - w.shape[1]))
            return dict(w=R(packing.pack_conv_weight(w)), b=R(g(p + ".bias")), n=w.shape[0], k=w.shape[-1])

        def resnet(p):
            r = dict(n1w=R(g(p + ".norm1.weight")), n1b=R(g(p + ".norm1.bias")), c1=conv(p + ".conv1"),
                     n2w=R(g(p + ".norm2.weight")), n2b=R(g(p + ".norm2.bias")), c2=conv(p + ".conv2"))
            if (p + ".conv_shortcut.weight") in sd:
                r["sc"] = conv(p + ".conv_shortcut")
            return r

        self.pq = conv("post_quant_conv", pad_in=self.lat_pad) if "post_quant_conv.weight" in sd else None
        d = "decoder"
        self.conv_in = conv(d + ".conv_in", pad_in=self.lat_pad)
        self.mid_res = [resnet(d + ".mid_block.resnets.0"), resnet(d + ".mid_block.resnets.1")]
        self.mid_attn = None
        a = d + ".mid_block.attentions.0"
        if (a + ".to_q.weight") in sd:
            C = sd[a + ".to_q.weight"].shape[0]
            qkv_w = torch.cat([g(a + ".to_q.weight"), g(a + ".to_k.weight"), g(a + ".to_v.weight")], 0)
            qkv_b = torch.cat([g(a + ".to_q.bias"), g(a + ".to_k.bias"), g(a + ".to_v.bias")], 0)
            self.mid_attn = dict(C=C, gw=R(g(a + ".group_norm.weight")), gb=R(g(a + ".group_norm.bias")),
                                 qkv=R(packing.pack_linear_weight(qkv_w)), qkvb=R(qkv_b),
                                 ow=R(packing.pack_linear_weight(g(a + ".to_out.0.weight"))), ob=R(g(a + ".to_out.0.bias")))
        self.up = []
        i = 0
        while f"{d}.up_blocks.{i}.resnets.0.norm1.weight" in sd:
            p = f"{d}.up_blocks.{i}"
            blk = dict(res=[], up=None)
            j = 0
            while f"{p}.resnets.{j}.norm1.weight" in sd:
                blk["res"].append(resnet(f"{p}.resnets.{j}"))
                j += 1
            if f"{p}.upsamplers.0.conv.weight" in sd:
                # 3x3 weight kept for reference_state_dict; the forward runs the four 2x2 parity weights (ops.upsample2x_conv)
                blk["up"] = conv(f"{p}.upsamplers.0.conv")
                uw = g(f"{p}.upsamplers.0.conv.weight")
                blk["up"]["w4"] = [R(t) for t in packing.pack_upsample_conv(uw)] if uw.shape[0] % 32 == 0 else None
            self.up.append(blk)
            i += 1
        self.norm_out = dict(w=R(g(d + ".conv_norm_out.weight")), b=R(g(d + ".conv_norm_out.bias")))
        self.conv_out = conv(d + ".conv_out")

    def _build_encoder(self, sd, device):
        cfg = self.config
        R = lambda t: self._reg(t, device)  # noqa: E731
        g = lambda k: sd[k].to(torch.float32)  # noqa: E731
        self.img_pad = packing.rup(cfg["in_channels"], 8)
        self.mom_pad = packing.rup(2 * cfg["latent_channels"], 8)

        def conv(p, pad_in=None):
            w = g(p + ".weight")
            if pad_in is not None and w.shape[1] < pad_in:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, pad_in - w.shape[1]))
            return dict(w=R(packing.pack_conv_weight(w)), b=R(g(p + ".bias")), n=w.shape[0], k=w.shape[-1])

        def resnet(p):
            r = dict(n1w=R(g(p + ".norm1.weight")), n1b=R(g(p + ".norm1.bias")), c1=conv(p + ".conv1"),
                     n2w=R(g(p + ".norm2.weight")), n2b=R(g(p + ".norm2.bias")), c2=conv(p + ".conv2"))
            if (p + ".conv_shortcut.weight") in sd:
                r["sc"] = conv(p + ".conv_shortcut")
            return r

        e = "encoder"
        enc = dict(conv_in=conv(e + ".conv_in", pad_in=self.img_pad), down=[])
        i = 0
        while f"{e}.down_blocks.{i}.resnets.0.norm1.weight" in sd:
            p = f"{e}.down_blocks.{i}"
            blk = dict(res=[], down=None)
            j = 0
            while f"{p}.resnets.{j}.norm1.weight" in sd:
                blk["res"].append(resnet(f"{p}.resnets.{j}"))
                j += 1
            if f"{p}.downsamplers.0.conv.weight" in sd:
                blk["down"] = conv(f"{p}.downsamplers.0.conv")
            enc["down"].append(blk)
            i += 1
        enc["mid_res"] = [resnet(e + ".mid_block.resnets.0"), resnet(e + ".mid_block.resnets.1")]
        enc["mid_attn"] = None
        a = e + ".mid_block.attentions.0"
        if (a + ".to_q.weight") in sd:
            C = sd[a + ".to_q.weight"].shape[0]
            qkv_w = torch.cat([g(a + ".to_q.weight"), g(a + ".to_k.weight"), g(a + ".to_v.weight")], 0)
            qkv_b = torch.cat([g(a + ".to_q.bias"), g(a + ".to_k.bias"), g(a + ".to_v.bias")], 0)
            enc["mid_attn"] = dict(C=C, gw=R(g(a + ".group_norm.weight")), gb=R(g(a + ".group_norm.bias")),
                                   qkv=R(packing.pack_linear_weight(qkv_w)), qkvb=R(qkv_b),
                                   ow=R(packing.pack_linear_weight(g(a + ".to_out.0.weight"))), ob=R(g(a + ".to_out.0.bias")))
        enc["norm_out"] = dict(w=R(g(e + ".conv_norm_out.weight")), b=R(g(e + ".conv_norm_out.bias")))
        enc["conv_out"] = conv(e + ".conv_out")
        enc["quant"] = conv("quant_conv", pad_in=self.mom_pad) if "quant_conv.weight" in sd else None
        return enc

    def reference_state_dict(self):
        """The decoder half (+ post_quant_conv), and the encoder half (+ quant_conv) when it was loaded, of the reference's
        `state_dict()`, rebuilt from the packed buffers: the exact inverse of `_build` / `_build_encoder`."""
        spec = specs.vae_decoder_params(dict(self.config))
        if self.enc is not None:
            spec.update(specs.vae_encoder_params(dict(self.config)))
        W = lambda n: self._buffers[n].detach().cpu()  # noqa: E731
        out = {}

        def conv(p, c, cin, padded_from=None):
            w = packing.unpack_conv_weight(W(c["w"]), padded_from or cin, c["k"])
            out[p + ".weight"], out[p + ".bias"] = w[:, :cin].contiguous(), W(c["b"])

        def resnet(p, r, cin):
            cout = r["c1"]["n"]
            out[p + ".norm1.weight"], out[p + ".norm1.bias"] = W(r["n1w"]), W(r["n1b"])
            conv(p + ".conv1", r["c1"], cin)
            out[p + ".norm2.weight"], out[p + ".norm2.bias"] = W(r["n2w"]), W(r["n2b"])
            conv(p + ".conv2", r["c2"], cout)
            if "sc" in r:
                conv(p + ".conv_shortcut", r["sc"], cin)
            return cout

        lc = self.config["latent_channels"]
        if self.pq is not None:
            conv("post_quant_conv", self.pq, lc, padded_from=self.lat_pad)
        d = "decoder"
        conv(d + ".conv_in", self.conv_in, lc, padded_from=self.lat_pad)
        ch = self.conv_in["n"]
        ch = resnet(d + ".mid_block.resnets.0", self.mid_res[0], ch)
        if self.mid_attn is not None:
            a, m = d + ".mid_block.attentions.0", self.mid_attn
            C = m["C"]
            out[a + ".group_norm.weight"], out[a + ".group_norm.bias"] = W(m["gw"]), W(m["gb"])
            qkv, qkvb = packing.unpack_linear_weight(W(m["qkv"]), C), W(m["qkvb"])
            for i, nm in enumerate(("to_q", "to_k", "to_v")):
                out[f"{a}.{nm}.weight"], out[f"{a}.{nm}.bias"] = qkv[i * C:(i + 1) * C].contiguous(), qkvb[i * C:(i + 1) * C].contiguous()
            out[a + ".to_out.0.weight"], out[a + ".to_out.0.bias"] = packing.unpack_linear_weight(W(m["ow"]), C), W(m["ob"])
        ch = resnet(d + ".mid_block.resnets.1", self.mid_res[1], ch)
        for i, blk in enumerate(self.up):
            for j, r in enumerate(blk["res"]):
                ch = resnet(f"{d}.up_blocks.{i}.resnets.{j}", r, ch)
            if blk["up"] is not None:
                conv(f"{d}.up_blocks.{i}.upsamplers.0.conv", blk["up"], ch)
        out[d + ".conv_norm_out.weight"], out[d + ".conv_norm_out.bias"] = W(self.norm_out["w"]), W(self.norm_out["b"])
        conv(d + ".conv_out", self.conv_out, ch)
        if self.enc is not None:
            e, en = "encoder", self.enc
            cin = self.config["in_channels"]
            conv(e + ".conv_in", en["conv_in"], cin, padded_from=self.img_pad)
            ch = en["conv_in"]["n"]
            for i, blk in enumerate(en["down"]):
                for j, r in enumerate(blk["res"]):
                    ch = resnet(f"{e}.down_blocks.{i}.resnets.{j}", r, ch)
                if blk["down"] is not None:
                    conv(f"{e}.down_blocks.{i}.downsamplers.0.conv", blk["down"], ch)
            ch = resnet(e + ".mid_block.resnets.0", en["mid_res"][0], ch)
            if en["mid_attn"] is not None:
                a, m = e + ".mid_block.attentions.0", en["mid_attn"]
                C = m["C"]
                out[a + ".group_norm.weight"], out[a + ".group_norm.bias"] = W(m["gw"]), W(m["gb"])
                qkv, qkvb = packing.unpack_linear_weight(W(m["qkv"]), C), W(m["qkvb"])
                for i, nm in enumerate(("to_q", "to_k", "to_v")):
                    out[f"{a}.{nm}.weight"], out[f"{a}.{nm}.bias"] = qkv[i * C:(i + 1) * C].contiguous(), qkvb[i * C:(i + 1) * C].contiguous()
                out[a + ".to_out.0.weight"], out[a + ".to_out.0.bias"] = packing.unpack_linear_weight(W(m["ow"]), C), W(m["ob"])
            ch = resnet(e + ".mid_block.resnets.1", en["mid_res"][1], ch)
            out[e + ".conv_norm_out.weight"], out[e + ".conv_norm_out.bias"] = W(en["norm_out"]["w"]), W(en["norm_out"]["b"])
            conv(e + ".conv_out", en["conv_out"], ch)
            if en["quant"] is not None:
                conv("quant_conv", en["quant"], 2 * lc, padded_from=self.mom_pad)
        missing = [k for k in spec if k not in out]
        if missing or len(out) != len(spec):
            raise RuntimeError(f"reference_state_dict: {len(missing)} parameters not reconstructed, e.g. {missing[:3]}")
        return {k: out[k].reshape(spec[k]).contiguous() for k in spec}

    def save_pretrained(self, save_directory, *args, **kwargs):
        if self.enc is None:
            raise NotImplementedError("this AutoencoderKL (B200) was built from the decoder half of a checkpoint only: it cannot write a "
                                      "complete reference checkpoint (reference_state_dict() returns the decoder parameters)")
        return FromPretrainedMixin.save_pretrained(self, save_directory, *args, **kwargs)

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._buffers["w0"].device

    @classmethod
    def random_init(cls, config=None, seed=0, dtype=torch.bfloat16, device="cuda", encoder=False):
        cfg = dict(specs.SDXL_VAE_CONFIG)
        cfg.update(config or {})
        # weights drawn on the target device (seconds for 11.9 B parameters on a GPU); device="cpu" keeps the CPU stream
        spec = specs.vae_params(cfg) if encoder else specs.vae_decoder_params(cfg)
        sd = specs.random_state_dict(spec, seed=seed, dtype=dtype, device=device if torch.device(device).type == "cuda" else "cpu")
        return cls(cfg, sd, dtype=dtype, device=device)

    # ------------------------------------------------------------------
    def _conv(self, c, x, B, H, W, residual=None, out=None, stride=1, pad_after_only=False):
        return ops.conv_gemm(x, self.W(c["w"]), c["n"], batch=B, H=H, W=W, ksize=c["k"], stride=stride, bias=self.W(c["b"]),
                             residual=residual, out=out, pad_after_only=pad_after_only)

    def _gn(self, x, w, b, B, hw, silu):
        return ops.group_norm(x, batch=B, hw=hw, groups=self.config["norm_num_groups"], eps=1e-6, gamma=self.W(w),
                              beta=self.W(b), silu=silu)

    def _resnet(self, r, x, B, H, W):
        n1 = self._gn(x, r["n1w"], r["n1b"], B, H * W, True)
        h = self._conv(r["c1"], n1, B, H, W)
        n2 = self._gn(h, r["n2w"], r["n2b"], B, H * W, True)
        sc = self._conv(r["sc"], x, B, H, W) if "sc" in r else x
        return self._conv(r["c2"], n2, B, H, W, residual=sc)

    def _attention(self, a, x, B, H, W):
        hw, C = H * W, a["C"]
        n = self._gn(x, a["gw"], a["gb"], B, hw, False)
        qkv = ops.linear(n, self.W(a["qkv"]), 3 * C, bias=self.W(a["qkvb"]))
        if C in (64, 128):
            q3 = qkv.view(B, hw, 3 * C)
            o = ops.attention(q3[:, :, :C], q3[:, :, C:2 * C], q3[:, :, 2 * C:], heads=1, head_dim=C).view(B * hw, C)
        else:
            o = torch.empty((B * hw, C), dtype=x.dtype, device=x.device)
            for b in range(B):
                rows = qkv[b * hw:(b + 1) * hw]
                o[b * hw:(b + 1) * hw] = ops.attention_unfused(rows[:, :C], rows[:, C:2 * C], rows[:, 2 * C:],
                                                              scale=C ** -0.5)
        return ops.linear(o, self.W(a["ow"]), C, bias=self.W(a["ob"]), residual=x)

    @ops.prefetching_forward
    def _decode_nhwc(self, z_nhwc, B, H, W):
        x = z_nhwc
        if self.pq is not None:
            y = torch.zeros((B * H * W, self.lat_pad), dtype=self._dtype, device=x.device)
            x = self._conv(self.pq, x, B, H, W, out=y)
        x = self._conv(self.conv_in, x, B, H, W)
        x = self._resnet(self.mid_res[0], x, B, H, W)
        if self.mid_attn is not None:
            x = self._attention(self.mid_attn, x, B, H, W)
        x = self._resnet(self.mid_res[1], x, B, H, W)
        for blk in self.up:
            for r in blk["res"]:
                x = self._resnet(r, x, B, H, W)
            if blk["up"] is not None:
                u = blk["up"]
                if u["w4"] is not None:
                    x = ops.upsample2x_conv(x, [self.W(t) for t in u["w4"]], u["n"], batch=B, H=H, W=W, bias=self.W(u["b"]))
                    H, W = 2 * H, 2 * W
                else:
                    xu = ops.upsample_nearest2x(x, batch=B, H=H, W=W)
                    H, W = 2 * H, 2 * W
                    x = self._conv(u, xu, B, H, W)
        n = self._gn(x, self.norm_out["w"], self.norm_out["b"], B, H * W, True)
        y = self._conv(self.conv_out, n, B, H, W)
        return y, H, W

    @torch.no_grad()
    def decode(self, z, return_dict=True, generator=None, max_batch=None):
        if not z.is_cuda:
            raise ops.B200Error("AutoencoderKL (B200) needs CUDA tensors: there is no CPU fallback")
        z = z.to(self._dtype)
        B, C, H, W = z.shape
        # 1024^2 activations are 268 MB per 128-channel tensor per image: decode in sub-batches (HBM footprint)
        mb = max_batch or max(1, min(B, (8 * 128 * 128) // (H * W) or 1))
        outs = []
        for b0 in range(0, B, mb):
            zb = z[b0:b0 + mb]
            nb = zb.shape[0]
            x = ops.nchw_to_nhwc(zb, c_pad=self.lat_pad)
            y, Ho, Wo = self._decode_nhwc(x, nb, H, W)
            outs.append(ops.nhwc_to_nchw(y, batch=nb, C_out=self.config["out_channels"], H=Ho, W=Wo))
        out = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        if not return_dict:
            return (out,)
        return DecoderOutput(out)

    @ops.prefetching_forward
    def _encode_nhwc(self, x, B, H, W):
        en = self.enc
        x = self._conv(en["conv_in"], x, B, H, W)
        for blk in en["down"]:
            for r in blk["res"]:
                x = self._resnet(r, x, B, H, W)
            if blk["down"] is not None:
                # Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) + stride-2 conv, as tap offsets of the implicit GEMM (no padded copy)
                x = self._conv(blk["down"], x, B, H, W, stride=2, pad_after_only=True)
                H, W = H // 2, W // 2
        x = self._resnet(en["mid_res"][0], x, B, H, W)
        if en["mid_attn"] is not None:
            x = self._attention(en["mid_attn"], x, B, H, W)
        x = self._resnet(en["mid_res"][1], x, B, H, W)
        n = self._gn(x, en["norm_out"]["w"], en["norm_out"]["b"], B, H * W, True)
        y = torch.zeros((B * H * W, self.mom_pad), dtype=self._dtype, device=x.device) if en["conv_out"]["n"] != self.mom_pad else None
        y = self._conv(en["conv_out"], n, B, H, W, out=y)
        if en["quant"] is not None:
            y2 = torch.zeros((B * H * W, self.mom_pad), dtype=self._dtype, device=x.device) if en["quant"]["n"] != self.mom_pad else None
            y = self._conv(en["quant"], y, B, H, W, out=y2)
        return y, H, W

    @torch.no_grad()
    def encode(self, x, return_dict=True, max_batch=None):
        """AutoencoderKL.encode (autoencoder_kl.py:169-197): images [B, C, H, W] -> AutoencoderKLOutput(latent_dist)."""
        if self.enc is None:
            raise NotImplementedError("this AutoencoderKL was built without the encoder half of the checkpoint (encoder.* / quant_conv)")
        if not x.is_cuda:
            raise ops.B200Error("AutoencoderKL (B200) needs CUDA tensors: there is no CPU fallback")
        x = x.to(self._dtype)
        B, C, H, W = x.shape
        f = 2 ** (len(self.enc["down"]) - 1)
        if H % f or W % f:
            raise NotImplementedError(f"encode: height and width must be multiples of {f}")
        mb = max_batch or max(1, min(B, (8 * 1024 * 1024) // (H * W) or 1))
        outs = []
        for b0 in range(0, B, mb):
            xb = x[b0:b0 + mb]
            nb = xb.shape[0]
            y, Ho, Wo = self._encode_nhwc(ops.nchw_to_nhwc(xb, c_pad=self.img_pad), nb, H, W)
            outs.append(ops.nhwc_to_nchw(y, batch=nb, C_out=2 * self.config["latent_channels"], H=Ho, W=Wo))
        moments = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    def forward(self, sample, sample_posterior=False, return_dict=True, generator=None):
        """AutoencoderKL.forward (autoencoder_kl.py:407-432): encode -> sample / mode -> decode."""
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z).sample
        if not return_dict:
            return (dec,)
        return DecoderOutput(dec)
