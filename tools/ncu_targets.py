"""One launch of each hot kernel at its dominant SDXL shape between cudaProfilerStart/Stop, for
  ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r1_targets python tools/ncu_targets.py
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
gam, bet = rnd(320), rnd(320)
lng, lnb = rnd(1280), rnd(1280)
flush = torch.empty(64 << 20, dtype=torch.int32, device=dev)


def run():
    ops.linear(x, w, N, bias=b, residual=r)                                            # attn out-proj / to_q: 192 per forward
    ops.linear(x, wf, NF, bias=bfp, geglu=True, tile_n=tn)                             # FF GEGLU: 60 per forward
    ops.conv_gemm(xc, wc, 320, batch=2, H=128, W=128, ksize=3, bias=gam)               # 3x3 conv 320->320 at 128x128
    ops.attention(qkv[..., :640], qkv[..., 640:1280], qkv[..., 1280:], heads=10, head_dim=64)   # self-attn 4096 tokens
    ops.attention(qkv2[..., :1280], qkv2[..., 1280:2560], qkv2[..., 2560:], heads=20, head_dim=64)  # self-attn 1024 tokens
    ops.group_norm(xc, batch=2, hw=128 * 128, groups=32, eps=1e-5, gamma=gam, beta=bet, silu=True)
    ops.layer_norm(x, eps=1e-5, gamma=lng, beta=lnb)


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
