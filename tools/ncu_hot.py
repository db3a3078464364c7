"""Repeated launches of ONE hot kernel between cudaProfilerStart/Stop, for a high-rate sampling capture:
  ncu --set full --import-source on --clock-control none --profile-from-start off --warp-sampling-interval 0 -c 8 -o gpurun_out/x python tools/ncu_hot.py gemm
Targets: gemm (linear 2048x1280x1280 +bias +residual), gn (GroupNorm+SiLU 2x16384x320), gn64 (2x4096x640, slab kernel),
attn (self-attention 2x10 heads x 4096 tokens, head_dim 64), attn1k (2x20x1024)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusers_b200 import ops, packing

what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev) * sc).bfloat16()  # noqa: E731
if what == "gemm":
    M, N, K = 2048, 1280, 1280
    xs, rs, b = [rnd(M, K) for _ in range(4)], [rnd(M, N) for _ in range(4)], rnd(N)
    ws = [packing.pack_linear_weight(rnd(N, K, sc=K ** -0.5)) for _ in range(4)]
    fn = lambda i: ops.linear(xs[i % 4], ws[i % 4], N, bias=b, residual=rs[i % 4])  # noqa: E731
elif what in ("gn", "gn64"):
    hw, C = (16384, 320) if what == "gn" else (4096, 640)
    xs = [rnd(2 * hw, C) for _ in range(4)]
    gam, bet = rnd(C), rnd(C)
    fn = lambda i: ops.group_norm(xs[i % 4], batch=2, hw=hw, groups=32, eps=1e-5, gamma=gam, beta=bet, silu=True)  # noqa: E731
elif what == "attn128":
    qkvs = [rnd(1, 4608, 3 * 24 * 128) for _ in range(2)]
    C = 24 * 128
    fn = lambda i: ops.attention(qkvs[i % 2][..., :C], qkvs[i % 2][..., C:2 * C], qkvs[i % 2][..., 2 * C:], heads=24, head_dim=128)  # noqa: E731
else:
    S, H = (4096, 10) if what == "attn" else (1024, 20)
    qkvs = [rnd(2, S, 3 * H * 64) for _ in range(2)]
    C = H * 64
    fn = lambda i: ops.attention(qkvs[i % 2][..., :C], qkvs[i % 2][..., C:2 * C], qkvs[i % 2][..., 2 * C:], heads=H, head_dim=64)  # noqa: E731
for i in range(4):
    fn(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(n):
    fn(i)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
