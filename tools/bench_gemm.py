"""Microbenchmark of b200_conv_gemm over (tile_n, cluster_m) for the shapes that dominate the SDXL UNet.
Prints us/launch (CUDA events, 20 back-to-back launches after warm-up) and TFLOP/s."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusers_b200 import ops, packing

SHAPES = [  # (M, N, K, residual)
    (2048, 1280, 1280, True), (2048, 3840, 1280, False), (2048, 1280, 5120, True), (2048, 10240, 1280, False),
    (8192, 640, 640, True), (8192, 1920, 640, False), (8192, 640, 2560, True), (32768, 320, 2880, False),
]


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    only = [int(a) for a in sys.argv[1:]]
    for si, (M, N, K, res) in enumerate(SHAPES):
        if only and si not in only:
            continue
        x = torch.randn(M, K, generator=g, device="cuda").bfloat16()
        w = packing.pack_linear_weight((torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).bfloat16())
        b = torch.randn(N, generator=g, device="cuda").bfloat16()
        r = torch.randn(M, N, generator=g, device="cuda").bfloat16() if res else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        print(f"--- M={M} N={N} K={K} residual={res}  (auto tile_n={ops.pick_tile_n(M, N)})")
        cfgs = [(int(os.environ['BN']), int(os.environ.get('CM', 1)))] if 'BN' in os.environ else itertools.product((256, 192, 160, 128, 96, 64), (1, 2))
        for bn, cm in cfgs:
            if (bn // cm) % 8:
                continue
            try:
                ws = [w] + [w.clone() for _ in range(7)]  # rotate weight copies so the weights are not L2-resident
                f = lambda i=0: ops.linear(x, ws[i % 8], N, bias=b, residual=r, out=out, tile_n=bn, cluster_m=cm)  # noqa: E731
                for i in range(3):
                    f(i)
                torch.cuda.synchronize()
                # GPU-bound timing: 24 launches captured in one CUDA graph (host launch cost would otherwise dominate)
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    for i in range(24):
                        f(i)
                gph.replay()
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    gph.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3 / 24)
                ts.sort()
                us = ts[len(ts) // 2]
                print(f"   bn={bn:3d} cm={cm}  {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s  (min {ts[0]:.1f})", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"   bn={bn} cm={cm} ERROR {str(e)[:120]}")


if __name__ == "__main__":
    main()
