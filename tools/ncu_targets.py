"""One launch of each hot kernel at its dominant SDXL / Flux / VAE shape between cudaProfilerStart/Stop, for
  ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r2_targets python tools/ncu_targets.py
Weights are cold (L2 flushed, as in a real forward), activations warm."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusers_b200 import ops, packing

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev) * sc).bfloat16()  # noqa: E731

M, N, K = 2048, 1280, 1280
x, r, b = rnd(M, K), rnd(M, N), rnd(N)
w = packing.pack_linear_weight(rnd(N, K, sc=K ** -0.5))
NF = 10240
wf_raw, bf = rnd(NF, K, sc=K ** -0.5), rnd(NF)
tn = ops.pick_tile_n(M, NF, True)
wf, bfp = packing.pack_geglu(wf_raw, bf, tn)
xc = rnd(2 * 128 * 128, 320)
wc = packing.pack_conv_weight(rnd(320, 320, 3, 3, sc=(320 * 9) ** -0.5))
qkv = rnd(2, 4096, 3 * 640)
qkv2 = rnd(2, 1024, 3 * 1280)
kvx = rnd(2, 77, 2 * 1280)
gam, bet = rnd(320), rnd(320)
lng, lnb = rnd(1280), rnd(1280)
# Flux / VAE shapes
fq = rnd(1, 4608, 3 * 3072)
cos = torch.rand(4608, 128, generator=g, device=dev)
sin = torch.rand(4608, 128, generator=g, device=dev)
nw = rnd(128)
temb = rnd(1, 3072)
wmod = rnd(18432 * 8, 3072, sc=3072 ** -0.5)   # 8 of the 115 AdaLN projections (0.9 GB): same per-column work as the full batch
bmod = rnd(18432 * 8)
xv = rnd(1024 * 1024, 128)
wv = packing.pack_conv_weight(rnd(128, 128, 3, 3, sc=(128 * 9) ** -0.5))
gv, bv = rnd(128), rnd(128)
# round 2, second session: fused QKV projection with the q/k RMSNorm + RoPE epilogue (Flux single block), T5-XXL / CLIP text attention
xq = rnd(4608, 3072)
wq3 = packing.pack_linear_weight(rnd(9216, 3072, sc=3072 ** -0.5))
bq3 = rnd(9216)
nw2 = (rnd(2, 128, sc=0.1) + 1).contiguous()
cosT, sinT = ops.rope_tables_transposed(cos, sin)
t5q = rnd(1, 512, 3 * 4096)
t5b = (torch.randn(64, 512, 512, generator=g, device=dev)).contiguous()
clq = rnd(2, 77, 3 * 1280)
flush = torch.empty(64 << 20, dtype=torch.int32, device=dev)


def run():
    ops.linear(x, w, N, bias=b, residual=r)                                            # attn out-proj / to_q: 192 per forward
    ops.linear(x, wf, NF, bias=bfp, geglu=True, tile_n=tn)                             # FF GEGLU: 60 per forward
    ops.conv_gemm(xc, wc, 320, batch=2, H=128, W=128, ksize=3, bias=gam)               # 3x3 conv 320->320 at 128x128
    ops.attention(qkv[..., :640], qkv[..., 640:1280], qkv[..., 1280:], heads=10, head_dim=64)   # self-attn 4096 tokens
    ops.attention(qkv2[..., :1280], qkv2[..., 1280:2560], qkv2[..., 2560:], heads=20, head_dim=64)  # self-attn 1024 tokens
    ops.attention(qkv2[..., :1280], kvx[..., :1280], kvx[..., 1280:], heads=20, head_dim=64)        # cross-attn 1024 x 77
    ops.group_norm(xc, batch=2, hw=128 * 128, groups=32, eps=1e-5, gamma=gam, beta=bet, silu=True)
    ops.layer_norm(x, eps=1e-5, gamma=lng, beta=lnb)
    ops.attention(fq[..., :3072], fq[..., 3072:6144], fq[..., 6144:], heads=24, head_dim=128)       # Flux joint attention
    ops.qk_norm_rope(fq.view(4608, -1), heads=24, head_dim=128, k_off=3072, seq=4608, txt_rows=512, wq=nw, wk=nw, wq_txt=nw, wk_txt=nw,
                     cos=cos, sin=sin)
    ops.small_linear(temb, wmod, bias=bmod, act_in=ops.ACT_SILU)                       # Flux AdaLN modulation GEMV batch
    ops.conv_gemm(xv, wv, 128, batch=1, H=1024, W=1024, ksize=3, bias=gv)              # VAE 128-ch conv at 1024^2
    ops.group_norm(xv, batch=1, hw=1024 * 1024, groups=32, eps=1e-6, gamma=gv, beta=bv, silu=True)
    ops.linear(xq, wq3, 9216, bias=bq3, qk_rope=ops.QkRope(nw2, cosT, sinT, 0, 6144, 128))         # Flux fused QKV + q/k RMSNorm + RoPE epilogue
    ops.linear(xq, wq3, 9216, bias=bq3)                                                              # the same projection, plain epilogue
    ops.text_attention(t5q[..., :4096], t5q[..., 4096:8192], t5q[..., 8192:], heads=64, scale=1.0, bias=t5b)     # T5-XXL self-attention, 512 tokens
    ops.text_attention(clq[..., :1280], clq[..., 1280:2560], clq[..., 2560:], heads=20, scale=0.125, causal=True)  # OpenCLIP bigG, 77 tokens


for _ in range(2):
    run()
torch.cuda.synchronize()
flush.zero_()
for t in (x, r, xc, qkv, qkv2):
    t.float().sum()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
