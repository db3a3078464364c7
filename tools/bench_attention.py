"""Device time of b200_attention at the SDXL / Flux shapes (20 launches inside one CUDA graph, L2-resident inputs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusers_b200 import ops

SHAPES = [("sdxl self 4096", 2, 10, 4096, 4096, 64), ("sdxl self 1024", 2, 20, 1024, 1024, 64), ("sdxl cross 4096x77", 2, 10, 4096, 77, 64),
          ("sdxl cross 1024x77", 2, 20, 1024, 77, 64), ("flux joint 4608", 1, 24, 4608, 4608, 128)]
g = torch.Generator(device="cuda").manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] in ("d64", "d128"):
    SHAPES = [s for s in SHAPES if s[5] == int(sys.argv[1][1:])]
for name, B, H, Sq, Sk, D in SHAPES:
    q = torch.randn(B, Sq, H * D, generator=g, device="cuda").bfloat16()
    k = torch.randn(B, Sk, H * D, generator=g, device="cuda").bfloat16()
    v = torch.randn(B, Sk, H * D, generator=g, device="cuda").bfloat16()
    out = torch.empty_like(q)
    for nq in (0, 1, 2):
        for _ in range(3):
            ops.attention(q, k, v, heads=H, head_dim=D, out=out, nq=nq)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        n = 20
        with torch.cuda.graph(gr):
            for _ in range(n):
                ops.attention(q, k, v, heads=H, head_dim=D, out=out, nq=nq)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = 1000 * e0.elapsed_time(e1) / n
        fl = 4.0 * B * H * Sq * Sk * D
        print(f"{name:22s} nq={nq} (0 = heuristic) {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
