"""Drop-in FluxTransformer2DModel on libb200diff.so.

Same `forward(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
joint_attention_kwargs, return_dict)` signature, `.config`, `.dtype`, `cache_context()` as the reference
(models/transformers/transformer_flux.py:529,671; what FluxPipeline touches: SURVEY.md §8b).

B200-first structure (not a port):
  * all 115 AdaLN modulation projections (Linear(SiLU(temb)), M = batch) are ONE weight-streaming launch per
    forward; LayerNorm + (1+scale)*x+shift is one pass; gates and residual adds live in GEMM epilogues;
  * q/k/v (+ added q/k/v) GEMMs write straight into one joint [text | image] QKV buffer; RMSNorm + RoPE run
    in place on it; the tcgen05 attention kernel reads q/k/v as strided views (no cat, no permute);
  * single blocks never build cat([attn, mlp]): proj_out is a two-source-K GEMM;
  * context_embedder(text) and the RoPE tables are step-invariant and cached.
"""
import contextlib
import os

import torch

from . import ops, packing, specs
from .checkpoint import FromPretrainedMixin
from .config import FrozenConfig
from .ops import ACT_GELU_TANH, ACT_SILU


class Transformer2DModelOutput:
    def __init__(self, sample):
        self.sample = sample


class FluxTransformer2DModel(torch.nn.Module, FromPretrainedMixin):
    _ref_class_names = ("FluxTransformer2DModel",)

    @classmethod
    def _param_spec(cls, cfg):
        full = dict(specs.FLUX_DEV_CONFIG)
        full.update(cfg)
        return specs.flux_params(full)

    def __init__(self, config, state_dict, dtype=torch.bfloat16, device="cuda"):
        super().__init__()
        cfg = dict(specs.FLUX_DEV_CONFIG)
        cfg.update(config)
        if cfg.get("out_channels") is None:
            cfg["out_channels"] = cfg["in_channels"]
        self.config = FrozenConfig(cfg)
        self._dtype = dtype
        self._n = 0
        spec = specs.flux_params(cfg)
        for k, shp in spec.items():
            if k not in state_dict:
                raise ValueError(f"state_dict is missing {k}")
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: expected shape {tuple(shp)}, got {tuple(state_dict[k].shape)}")
        if cfg["attention_head_dim"] not in (64, 128):
            raise NotImplementedError("attention_head_dim must be 64 or 128 for the tcgen05 attention kernel")
        if sum(cfg["axes_dims_rope"]) != cfg["attention_head_dim"]:
            raise ValueError("sum(axes_dims_rope) must equal attention_head_dim")
        self._build(state_dict, torch.device(device))
        self._ctx_key = None
        self._ctx = None
        self._rope_key = None
        self._rope = None
        self._cp = None  # context parallelism (enable_parallelism): dict(world, rank, group, plans, peers)
        # per-head RMSNorm + rotary embedding of q / k in the EPILOGUE of the fused QKV GEMM (the north star's "QKV + RoPE" fusion;
        # B200_FLUX_FUSED_QKROPE=0: the separate in-place b200_qk_norm_rope pass of round 1, kept for A/B runs and as its test oracle)
        self.fused_qk_rope = os.environ.get("B200_FLUX_FUSED_QKROPE", "1") != "0"
        self._ropeT_key = None
        self._ropeT = None

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
        D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
        self.D = D

        def small(p):
            return dict(w=R(g(p + ".weight")), b=R(g(p + ".bias")))

        def lin(p, split=None):
            w = g(p + ".weight")
            return dict(w=R(packing.pack_linear_weight(w, split)), b=R(g(p + ".bias")), n=w.shape[0])

        def lin_cat(ps):
            w = torch.cat([g(p + ".weight") for p in ps], 0)
            b = torch.cat([g(p + ".bias") for p in ps], 0)
            return dict(w=R(packing.pack_linear_weight(w)), b=R(b), n=w.shape[0])

        te = "time_text_embed"
        self.t_emb = [small(te + ".timestep_embedder.linear_1"), small(te + ".timestep_embedder.linear_2")]
        self.g_emb = None
        if cfg.get("guidance_embeds", False):
            self.g_emb = [small(te + ".guidance_embedder.linear_1"), small(te + ".guidance_embedder.linear_2")]
        self.p_emb = [small(te + ".text_embedder.linear_1"), small(te + ".text_embedder.linear_2")]
        self.x_embedder = lin("x_embedder")
        self.context_embedder = lin("context_embedder")

        mod_w, mod_b = [], []
        self._mod_total = 0

        def mod(p):
            off = self._mod_total
            w = g(p + ".weight")
            mod_w.append(w)
            mod_b.append(g(p + ".bias"))
            self._mod_total += w.shape[0]
            return off

        self.double = []
        for i in range(cfg["num_layers"]):
            p = f"transformer_blocks.{i}"
            a = p + ".attn"
            self.double.append(dict(
                mod=mod(p + ".norm1.linear"), cmod=mod(p + ".norm1_context.linear"),
                qkv=lin_cat([a + ".to_q", a + ".to_k", a + ".to_v"]),
                aqkv=lin_cat([a + ".add_q_proj", a + ".add_k_proj", a + ".add_v_proj"]),
                nq=R(g(a + ".norm_q.weight")), nk=R(g(a + ".norm_k.weight")),
                naq=R(g(a + ".norm_added_q.weight")), nak=R(g(a + ".norm_added_k.weight")),
                nqk=R(torch.stack([g(a + ".norm_q.weight"), g(a + ".norm_k.weight")])),
                naqk=R(torch.stack([g(a + ".norm_added_q.weight"), g(a + ".norm_added_k.weight")])),
                out=lin(a + ".to_out.0"), aout=lin(a + ".to_add_out"),
                ff1=lin(p + ".ff.net.0.proj"), ff2=lin(p + ".ff.net.2"),
                cff1=lin(p + ".ff_context.net.0.proj"), cff2=lin(p + ".ff_context.net.2")))
        self.single = []
        for i in range(cfg["num_single_layers"]):
            p = f"single_transformer_blocks.{i}"
            a = p + ".attn"
            self.single.append(dict(
                mod=mod(p + ".norm.linear"),
                qkv=lin_cat([a + ".to_q", a + ".to_k", a + ".to_v"]),
                nq=R(g(a + ".norm_q.weight")), nk=R(g(a + ".norm_k.weight")),
                nqk=R(torch.stack([g(a + ".norm_q.weight"), g(a + ".norm_k.weight")])),
                mlp=lin(p + ".proj_mlp"), out=lin(p + ".proj_out", split=(D, 4 * D))))
        self.mod_out = mod("norm_out.linear")
        self.proj_out = lin("proj_out")
        self.mod_all = dict(w=R(torch.cat(mod_w, 0)), b=R(torch.cat(mod_b, 0)))

    def reference_state_dict(self):
        """The reference's `state_dict()` rebuilt from the packed buffers (exact inverse of `_build`)."""
        if self._cp is not None:
            raise RuntimeError("reference_state_dict: the fused QKV weights were permuted for context parallelism (enable_parallelism); "
                               "export the checkpoint from a model without it")
        cfg = self.config
        spec = specs.flux_params(dict(cfg))
        W = lambda n: self._buffers[n].detach().cpu()  # noqa: E731
        out = {}
        D = self.D
        mod_w, mod_b = W(self.mod_all["w"]), W(self.mod_all["b"])

        def small(p, d):
            out[p + ".weight"], out[p + ".bias"] = W(d["w"]), W(d["b"])

        def lin(p, d, split=None):
            out[p + ".weight"] = packing.unpack_linear_weight(W(d["w"]), spec[p + ".weight"][1], split)
            out[p + ".bias"] = W(d["b"])

        def lin_cat(ps, d):
            w, b = packing.unpack_linear_weight(W(d["w"]), spec[ps[0] + ".weight"][1]), W(d["b"])
            o = 0
            for q in ps:
                n = spec[q + ".weight"][0]
                out[q + ".weight"], out[q + ".bias"] = w[o:o + n].contiguous(), b[o:o + n].contiguous()
                o += n

        def mod(p, off):
            n = spec[p + ".weight"][0]
            out[p + ".weight"], out[p + ".bias"] = mod_w[off:off + n].contiguous(), mod_b[off:off + n].contiguous()

        te = "time_text_embed"
        small(te + ".timestep_embedder.linear_1", self.t_emb[0])
        small(te + ".timestep_embedder.linear_2", self.t_emb[1])
        if self.g_emb is not None:
            small(te + ".guidance_embedder.linear_1", self.g_emb[0])
            small(te + ".guidance_embedder.linear_2", self.g_emb[1])
        small(te + ".text_embedder.linear_1", self.p_emb[0])
        small(te + ".text_embedder.linear_2", self.p_emb[1])
        lin("x_embedder", self.x_embedder)
        lin("context_embedder", self.context_embedder)
        for i, blk in enumerate(self.double):
            p = f"transformer_blocks.{i}"
            a = p + ".attn"
            mod(p + ".norm1.linear", blk["mod"])
            mod(p + ".norm1_context.linear", blk["cmod"])
            lin_cat([a + ".to_q", a + ".to_k", a + ".to_v"], blk["qkv"])
            lin_cat([a + ".add_q_proj", a + ".add_k_proj", a + ".add_v_proj"], blk["aqkv"])
            out[a + ".norm_q.weight"], out[a + ".norm_k.weight"] = W(blk["nq"]), W(blk["nk"])
            out[a + ".norm_added_q.weight"], out[a + ".norm_added_k.weight"] = W(blk["naq"]), W(blk["nak"])
            lin(a + ".to_out.0", blk["out"])
            lin(a + ".to_add_out", blk["aout"])
            lin(p + ".ff.net.0.proj", blk["ff1"])
            lin(p + ".ff.net.2", blk["ff2"])
            lin(p + ".ff_context.net.0.proj", blk["cff1"])
            lin(p + ".ff_context.net.2", blk["cff2"])
        for i, blk in enumerate(self.single):
            p = f"single_transformer_blocks.{i}"
            a = p + ".attn"
            mod(p + ".norm.linear", blk["mod"])
            lin_cat([a + ".to_q", a + ".to_k", a + ".to_v"], blk["qkv"])
            out[a + ".norm_q.weight"], out[a + ".norm_k.weight"] = W(blk["nq"]), W(blk["nk"])
            lin(p + ".proj_mlp", blk["mlp"])
            lin(p + ".proj_out", blk["out"], split=(D, 4 * D))
        mod("norm_out.linear", self.mod_out)
        lin("proj_out", self.proj_out)
        missing = [k for k in spec if k not in out]
        if missing or len(out) != len(spec):
            raise RuntimeError(f"reference_state_dict: {len(missing)} parameters not reconstructed, e.g. {missing[:3]}")
        return {k: out[k].reshape(spec[k]).contiguous() for k in spec}

    # ------------------------------------------------------------------ reference-facing surface
    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._buffers["w0"].device

    @contextlib.contextmanager
    def cache_context(self, name):
        """CacheMixin.cache_context (models/cache_utils.py:155): the pipeline wraps every call in it."""
        yield

    def _reset_stateful_cache(self):
        self._ctx_key = self._ctx = self._rope_key = self._rope = self._ropeT_key = self._ropeT = None

    @classmethod
    def random_init(cls, config=None, seed=0, dtype=torch.bfloat16, device="cuda"):
        cfg = dict(specs.FLUX_DEV_CONFIG)
        cfg.update(config or {})
        # weights drawn on the target device (seconds for 11.9 B parameters on a GPU); device="cpu" keeps the CPU stream
        sd = specs.random_state_dict(specs.flux_params(cfg), seed=seed, dtype=dtype, device=device if torch.device(device).type == "cuda" else "cpu")
        return cls(cfg, sd, dtype=dtype, device=device)

    # ------------------------------------------------------------------ step-invariant pieces
    def _rope_tables(self, txt_ids, img_ids, owners=None):
        # keyed on the tensor OBJECTS the caller passed (kept alive by the key) + version counters: an address-based key
        # goes stale when the allocator hands a freed buffer to a new tensor.  `owners` = the caller's original arguments
        # when txt_ids / img_ids are per-call views of them (deprecated 3-D ids)
        ko = owners or (txt_ids, img_ids)
        k = self._rope_key
        if k is None or k[0] is not ko[0] or k[1] is not ko[1] or k[2] != (ko[0]._version, ko[1]._version):
            ids = torch.cat((txt_ids, img_ids), dim=0)
            pos = ids.float()
            cos_out, sin_out = [], []
            for i, dim in enumerate(self.config["axes_dims_rope"]):
                # FluxPosEmbed.forward / get_1d_rotary_pos_embed: fp64 frequencies, fp32 tables (host-side glue, once)
                freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2, dtype=torch.float64, device=pos.device) / dim))
                freqs = torch.outer(pos[:, i], freqs)
                cos_out.append(freqs.cos().repeat_interleave(2, dim=1, output_size=freqs.shape[1] * 2).float())
                sin_out.append(freqs.sin().repeat_interleave(2, dim=1, output_size=freqs.shape[1] * 2).float())
            self._rope = (torch.cat(cos_out, dim=-1).contiguous(), torch.cat(sin_out, dim=-1).contiguous())
            self._rope_key = (ko[0], ko[1], (ko[0]._version, ko[1]._version))
        return self._rope

    def _rope_transposed(self, rope):
        if self._ropeT_key is not rope[0]:
            self._ropeT = ops.rope_tables_transposed(rope[0], rope[1])
            self._ropeT_key = rope[0]
        return self._ropeT

    def _context(self, ehs):
        k = self._ctx_key
        if k is None or k[0] is not ehs or k[1] != ehs._version:
            B, T, Dj = ehs.shape
            ce = self.context_embedder
            self._ctx = ops.linear(ehs.to(self._dtype).contiguous().view(B * T, Dj), self.W(ce["w"]), ce["n"],
                                   bias=self.W(ce["b"])).view(B, T, self.D)
            self._ctx_key = (ehs, ehs._version)
        return self._ctx

    def _temb(self, timestep, guidance, pooled):
        dt = self._dtype
        B = pooled.shape[0]
        dev = pooled.device
        # timestep.to(hidden_states.dtype) * 1000 (transformer_flux.py:725): rounded to 16 bit like the reference
        t = (timestep.to(device=dev).to(dt) * 1000).to(torch.float32).reshape(-1).expand(B).contiguous()
        tp = ops.timestep_embedding(t, 256, dtype=dt, flip_sin_to_cos=True, downscale_freq_shift=0.0)
        e = ops.small_linear(tp, self.W(self.t_emb[0]["w"]), bias=self.W(self.t_emb[0]["b"]), act_out=ACT_SILU)
        emb = ops.small_linear(e, self.W(self.t_emb[1]["w"]), bias=self.W(self.t_emb[1]["b"]))
        if self.g_emb is not None:
            if guidance is None:
                raise ValueError("guidance_embeds=True requires `guidance`")
            gd = (guidance.to(device=dev).to(dt) * 1000).to(torch.float32).reshape(-1).expand(B).contiguous()
            gp = ops.timestep_embedding(gd, 256, dtype=dt, flip_sin_to_cos=True, downscale_freq_shift=0.0)
            e = ops.small_linear(gp, self.W(self.g_emb[0]["w"]), bias=self.W(self.g_emb[0]["b"]), act_out=ACT_SILU)
            emb = ops.small_linear(e, self.W(self.g_emb[1]["w"]), bias=self.W(self.g_emb[1]["b"]), addend=emb)
        e = ops.small_linear(pooled.to(dt), self.W(self.p_emb[0]["w"]), bias=self.W(self.p_emb[0]["b"]), act_out=ACT_SILU)
        return ops.small_linear(e, self.W(self.p_emb[1]["w"]), bias=self.W(self.p_emb[1]["b"]), addend=emb)

    # ------------------------------------------------------------------ context parallelism (Ulysses)
    def enable_parallelism(self, *, config, cp_plan=None, group=None):
        """ModelMixin.enable_parallelism (models/modeling_utils.py:1607) for `ContextParallelConfig(ulysses_degree=N)`: the
        token sequence of ONE sample is sharded over the N ranks of the (default) process group, every rank calls forward
        with the full inputs and receives the full output, like the reference's `_cp_plan` (transformer_flux.py:573: inputs
        split on the sequence dim, `proj_out` gathered).  See context_parallel.py for how the two all-to-alls of
        TemplatedUlyssesAttention (attention_dispatch.py:2504) are folded into the QKV GEMM and the attention kernel."""
        import torch.distributed as dist

        from .context_parallel import ContextParallelConfig, UlyssesPlan
        if cp_plan is not None:
            raise NotImplementedError("custom cp_plan: the built-in plan of FluxTransformer2DModel is the only one")
        if not isinstance(config, ContextParallelConfig):
            cpc = getattr(config, "context_parallel_config", None)  # the reference's ParallelConfig wrapper
            if cpc is None:
                raise NotImplementedError("only context parallelism (ContextParallelConfig) is built")
            config = ContextParallelConfig(ring_degree=cpc.ring_degree, ulysses_degree=cpc.ulysses_degree) if not isinstance(cpc, ContextParallelConfig) else cpc
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("torch.distributed must be initialized before calling `enable_parallelism`.")
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if config.ulysses_degree != world:
            raise ValueError(f"ulysses_degree ({config.ulysses_degree}) must equal the size of the process group ({world})")
        if self._cp is not None:
            raise RuntimeError("enable_parallelism was already called on this model")
        if world == 1:
            return
        cfg = self.config
        probe = UlyssesPlan(world, rank, world * 8, world * 8, cfg["num_attention_heads"], cfg["attention_head_dim"])  # validates heads % world
        perm = probe.qkv_row_permutation(probe.send_order()).to(self.device)  # block i = the rows of destination send_order()[i]
        for blk in self.double + self.single:
            for key in ("qkv", "aqkv"):
                if key in blk:
                    for part in ("w", "b"):
                        name = blk[key][part]
                        self._buffers[name] = self._buffers[name].index_select(0, perm).contiguous()
        self.__dict__.pop("_weight_prefetch_plan", None)  # the launch order (and the weight addresses) changed
        self._cp = dict(world=world, rank=rank, group=group, plans={}, peers={}, rope_key=None, rope=None)

    def _cp_state(self, T, S, out_ch):
        """(plan, peer buffers) for a joint sequence of T text + S image tokens; the peer-mapped buffers are allocated (a
        collective step: every rank reaches it with the same shapes) on first use."""
        from .context_parallel import PeerGroup, UlyssesPlan
        cp = self._cp
        key = (T, S, out_ch)
        if key not in cp["plans"]:
            cfg = self.config
            plan = UlyssesPlan(cp["world"], cp["rank"], T, S, cfg["num_attention_heads"], cfg["attention_head_dim"])
            es = torch.empty((), dtype=self._dtype).element_size()
            rup = lambda n: (n + 255) // 256 * 256  # noqa: E731
            nbytes = rup(plan.L * 3 * plan.Dl * es) + rup(plan.Ll * plan.D * es) + rup(S * out_ch * es)
            pg = PeerGroup(nbytes, cp["group"])
            bufs = dict(pg=pg, J=pg.carve(self._dtype, (plan.L, 3 * plan.Dl)), A=pg.carve(self._dtype, (plan.Ll, plan.D)),
                        O=pg.carve(self._dtype, (S, out_ch)))
            if len(cp["plans"]) >= 4:
                # a serving process that keeps changing resolution must not accumulate peer mappings (cudaMalloc + IPC handles): drop the
                # oldest shape.  Every rank sees the same sequence of shapes, so every rank closes the same buffers (close() is collective)
                old_key = next(iter(cp["plans"]))
                cp["plans"].pop(old_key)[1]["pg"].close()
            cp["plans"][key] = (plan, bufs)
        return cp["plans"][key]

    def _cp_rope(self, plan, rope):
        """(cos, sin) of the whole sequence in rank-major order [positions, head_dim] (destination-side b200_qk_norm_rope), and of
        THIS rank's rows transposed [head_dim / 2, Ll] (source-side fused epilogue)."""
        cp = self._cp
        if cp["rope_key"] is not rope[0]:
            g = plan.joint_to_rank_major().to(rope[0].device)
            full = (rope[0].index_select(0, g).contiguous(), rope[1].index_select(0, g).contiguous())
            b0, b1 = plan.block()
            cp["rope"] = full + ops.rope_tables_transposed(full[0][b0:b1], full[1][b0:b1])
            cp["rope_key"] = rope[0]
        return cp["rope"]

    def _forward_one_cp(self, x_in, ctx, mod, rope):
        """_forward_one with the joint sequence sharded over the ranks: this rank carries text rows [t0, t1) and image
        rows [s0, s1) through every per-token op; only the attention sees the whole sequence, for its own heads."""
        cfg = self.config
        D, hd = self.D, cfg["attention_head_dim"]
        S, T = x_in.shape[0], ctx.shape[0]
        plan, bufs = self._cp_state(T, S, self.proj_out["n"])
        pg, J, A, O = bufs["pg"], bufs["J"], bufs["A"], bufs["O"]
        r, P, Tl, Sl, Ll, L, Dl, hl = plan.rank, plan.world, plan.Tl, plan.Sl, plan.Ll, plan.L, plan.Dl, plan.hl
        (t0, t1), (s0, s1) = plan.text_rows(), plan.image_rows()
        b0 = plan.block()[0]
        cos, sin, cosT, sinT = self._cp_rope(plan, rope)
        fused = self.fused_qk_rope
        dev = x_in.device
        hbuf = torch.empty((Ll, D), dtype=self._dtype, device=dev)  # this rank's slice of the residual stream [text | image]
        c, x = hbuf[:Tl], hbuf[Tl:]
        self._lin(self.x_embedder, x_in[s0:s1], out=x)
        c.copy_(ctx[t0:t1])
        Jl, Al = J[r], A[r]
        J3 = Jl.view(1, L, 3 * Dl)
        o_seg = [A[s_][:, r * Dl:(r + 1) * Dl] for s_ in range(P)]  # rank s_ owns output rows [s_*Ll, (s_+1)*Ll): my heads' columns of its A
        order = plan.send_order()

        def m(off, i):
            return mod[:, off + i * D: off + (i + 1) * D]

        def send_qkv(l, rows_in, row_off, nw):
            # the first all-to-all of Ulysses, folded into the GEMM: destination d receives its heads' [q | k | v] columns of my rows,
            # q / k already normalised and rotated (positions of MY rows: the local transposed tables)
            n = rows_in.shape[0]
            qk = ops.QkRope(self.W(nw), cosT, sinT, row_off, 2 * Dl, hd, 1e-6) if fused else None
            ops.linear(rows_in, self.W(l["w"]), 3 * D, bias=self.W(l["b"]), qk_rope=qk,
                       out_blocks=[J[d][b0 + row_off: b0 + row_off + n] for d in order])  # ONE launch, column block i -> rank order[i]

        def attend(blk, txt_rows):
            pg.barrier()  # every rank's q/k/v tiles have landed in my J
            if not fused:
                ops.qk_norm_rope(Jl, heads=hl, head_dim=hd, k_off=Dl, seq=L, txt_rows=txt_rows, txt_period=Ll, wq=self.W(blk["nq"]),
                                 wk=self.W(blk["nk"]), wq_txt=self.W(blk["naq"]) if "naq" in blk else None,
                                 wk_txt=self.W(blk["nak"]) if "nak" in blk else None, cos=cos, sin=sin, eps=1e-6)
            # the second all-to-all, folded into the attention epilogue: row i goes to the rank that owns it
            ops.attention(J3[:, :, :Dl], J3[:, :, Dl:2 * Dl], J3[:, :, 2 * Dl:], heads=hl, head_dim=hd, o_seg=o_seg, o_seg_rows=Ll)
            pg.barrier()  # every rank's output rows have landed in my A
            return Al

        for blk in self.double:
            o, co = blk["mod"], blk["cmod"]
            nx = ops.layer_norm(x, eps=1e-6, scale=m(o, 1), shift=m(o, 0), rows_per_group=Sl)
            nc = ops.layer_norm(c, eps=1e-6, scale=m(co, 1), shift=m(co, 0), rows_per_group=Tl)
            send_qkv(blk["aqkv"], nc, 0, blk["naqk"])
            send_qkv(blk["qkv"], nx, Tl, blk["nqk"])
            a = attend(blk, Tl)
            self._lin(blk["out"], a[Tl:], gate=m(o, 2), rows_per_group=Sl, residual=x, out=x)
            self._lin(blk["aout"], a[:Tl], gate=m(co, 2), rows_per_group=Tl, residual=c, out=c)
            nx = ops.layer_norm(x, eps=1e-6, scale=m(o, 4), shift=m(o, 3), rows_per_group=Sl)
            h = self._lin(blk["ff1"], nx, act=ACT_GELU_TANH)
            self._lin(blk["ff2"], h, gate=m(o, 5), rows_per_group=Sl, residual=x, out=x)
            nc = ops.layer_norm(c, eps=1e-6, scale=m(co, 4), shift=m(co, 3), rows_per_group=Tl)
            h = self._lin(blk["cff1"], nc, act=ACT_GELU_TANH)
            self._lin(blk["cff2"], h, gate=m(co, 5), rows_per_group=Tl, residual=c, out=c)
        for blk in self.single:
            o = blk["mod"]
            nh_ = ops.layer_norm(hbuf, eps=1e-6, scale=m(o, 1), shift=m(o, 0), rows_per_group=Ll)
            send_qkv(blk["qkv"], nh_, 0, blk["nqk"])
            mlp = self._lin(blk["mlp"], nh_, act=ACT_GELU_TANH)
            a = attend(blk, 0)
            self._lin(blk["out"], a, x2=mlp, gate=m(o, 2), rows_per_group=Ll, residual=hbuf, out=hbuf)
        nx = ops.layer_norm(x, eps=1e-6, scale=m(self.mod_out, 0), shift=m(self.mod_out, 1), rows_per_group=Sl)
        for d in order:  # `proj_out` gathered on every rank (_cp_plan: ContextParallelOutput(gather_dim=1))
            self._lin(self.proj_out, nx, out=O[d][s0:s1])
        pg.barrier()
        return O[r].clone()

    # ------------------------------------------------------------------ one batch element
    def _lin(self, l, x, **kw):
        return ops.linear(x, self.W(l["w"]), l["n"], bias=self.W(l["b"]), **kw)

    def _forward_one(self, x_in, ctx, mod, rope):
        """x_in [S, in_ch], ctx [T, D], mod [1, mod_total] (all AdaLN projections of this sample)."""
        cfg = self.config
        D, nh, hd = self.D, cfg["num_attention_heads"], cfg["attention_head_dim"]
        S, T = x_in.shape[0], ctx.shape[0]
        L = S + T
        cos, sin = rope
        dev = x_in.device
        hbuf = torch.empty((L, D), dtype=self._dtype, device=dev)  # joint residual stream [text | image]
        c, x = hbuf[:T], hbuf[T:]
        self._lin(self.x_embedder, x_in, out=x)
        c.copy_(ctx)
        J = torch.empty((L, 3 * D), dtype=self._dtype, device=dev)  # joint fused QKV
        J3 = J.view(1, L, 3 * D)
        fused = self.fused_qk_rope
        if fused:
            cosT, sinT = self._rope_transposed(rope)

        def m(off, i):
            return mod[:, off + i * D: off + (i + 1) * D]

        def qkr(blk, key, row0):
            # q / k leave the projection normalised and rotated: no second pass over the joint buffer
            return ops.QkRope(self.W(blk[key]), cosT, sinT, row0, 2 * D, hd, 1e-6) if fused else None

        def attend(blk, txt_rows):
            if not fused:
                ops.qk_norm_rope(J, heads=nh, head_dim=hd, k_off=D, seq=L, txt_rows=txt_rows, wq=self.W(blk["nq"]),
                                 wk=self.W(blk["nk"]), wq_txt=self.W(blk["naq"]) if "naq" in blk else None,
                                 wk_txt=self.W(blk["nak"]) if "nak" in blk else None, cos=cos, sin=sin, eps=1e-6)
            return ops.attention(J3[:, :, :D], J3[:, :, D:2 * D], J3[:, :, 2 * D:], heads=nh, head_dim=hd).view(L, D)

        for blk in self.double:
            o, co = blk["mod"], blk["cmod"]
            nx = ops.layer_norm(x, eps=1e-6, scale=m(o, 1), shift=m(o, 0), rows_per_group=S)
            nc = ops.layer_norm(c, eps=1e-6, scale=m(co, 1), shift=m(co, 0), rows_per_group=T)
            self._lin(blk["aqkv"], nc, out=J[:T], qk_rope=qkr(blk, "naqk", 0))
            self._lin(blk["qkv"], nx, out=J[T:], qk_rope=qkr(blk, "nqk", T))
            a = attend(blk, T)
            self._lin(blk["out"], a[T:], gate=m(o, 2), rows_per_group=S, residual=x, out=x)
            self._lin(blk["aout"], a[:T], gate=m(co, 2), rows_per_group=T, residual=c, out=c)
            nx = ops.layer_norm(x, eps=1e-6, scale=m(o, 4), shift=m(o, 3), rows_per_group=S)
            h = self._lin(blk["ff1"], nx, act=ACT_GELU_TANH)
            self._lin(blk["ff2"], h, gate=m(o, 5), rows_per_group=S, residual=x, out=x)
            nc = ops.layer_norm(c, eps=1e-6, scale=m(co, 4), shift=m(co, 3), rows_per_group=T)
            h = self._lin(blk["cff1"], nc, act=ACT_GELU_TANH)
            self._lin(blk["cff2"], h, gate=m(co, 5), rows_per_group=T, residual=c, out=c)
        for blk in self.single:
            o = blk["mod"]
            nh_ = ops.layer_norm(hbuf, eps=1e-6, scale=m(o, 1), shift=m(o, 0), rows_per_group=L)
            self._lin(blk["qkv"], nh_, out=J, qk_rope=qkr(blk, "nqk", 0))
            mlp = self._lin(blk["mlp"], nh_, act=ACT_GELU_TANH)
            a = attend(blk, 0)
            self._lin(blk["out"], a, x2=mlp, gate=m(o, 2), rows_per_group=L, residual=hbuf, out=hbuf)
        # AdaLayerNormContinuous: scale, shift = chunk(2)  (normalization.py:349)
        nx = ops.layer_norm(x, eps=1e-6, scale=m(self.mod_out, 0), shift=m(self.mod_out, 1), rows_per_group=S)
        return self._lin(self.proj_out, nx)

    @torch.no_grad()
    @ops.prefetching_forward
    def forward(self, hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None, img_ids=None,
                txt_ids=None, guidance=None, joint_attention_kwargs=None, controlnet_block_samples=None,
                controlnet_single_block_samples=None, return_dict=True, controlnet_blocks_repeat=False):
        if controlnet_block_samples is not None or controlnet_single_block_samples is not None:
            raise NotImplementedError("controlnet residuals are outside the accelerated hot path")
        if joint_attention_kwargs:
            raise NotImplementedError("joint_attention_kwargs (IP-adapter / LoRA scale) are outside the hot path")
        if not hidden_states.is_cuda:
            raise ops.B200Error("FluxTransformer2DModel (B200) needs CUDA tensors: there is no CPU fallback")
        id_owners = (txt_ids, img_ids)
        if txt_ids.ndim == 3:
            txt_ids = txt_ids[0]
        if img_ids.ndim == 3:
            img_ids = img_ids[0]
        B, S, Cin = hidden_states.shape
        hs = hidden_states.to(self._dtype).contiguous()
        temb = self._temb(timestep, guidance, pooled_projections)
        mod = ops.small_linear(temb, self.W(self.mod_all["w"]), bias=self.W(self.mod_all["b"]), act_in=ACT_SILU)
        ctx = self._context(encoder_hidden_states)
        rope = self._rope_tables(txt_ids, img_ids, owners=id_owners)
        out = torch.empty((B, S, self.proj_out["n"]), dtype=self._dtype, device=hs.device)
        one = self._forward_one if self._cp is None else self._forward_one_cp
        for b in range(B):
            out[b] = one(hs[b], ctx[b], mod[b:b + 1], rope)
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(out)
