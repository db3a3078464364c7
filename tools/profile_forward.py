"""Profiling helper for one SDXL UNet forward (B=2, 1024^2).

  python tools/profile_forward.py gemm     # per-shape table of conv_gemm launches from CUDA events (no profiler)
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_forward.py ncu  # cudaProfilerStart/Stop around exactly one forward
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusers_b200 import ops
from diffusers_b200.unet_2d_condition import UNet2DConditionModel


def inputs(B2=2):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B2, 4, 128, 128, generator=g, device="cuda").bfloat16()
    ehs = torch.randn(B2, 77, 2048, generator=g, device="cuda").bfloat16()
    added = dict(text_embeds=torch.randn(B2, 1280, generator=g, device="cuda").bfloat16(),
                 time_ids=torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B2, device="cuda").bfloat16())
    return x, torch.tensor(981.0, device="cuda"), ehs, added


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    m = UNet2DConditionModel.random_init(seed=0)
    x, t, ehs, added = inputs()
    fwd = lambda: m(x, t, ehs, added_cond_kwargs=added, return_dict=False)  # noqa: E731
    for _ in range(2):
        fwd()
    torch.cuda.synchronize()
    if mode == "ncu":
        torch.cuda.profiler.start()
        fwd()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    ops._PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fwd()
    e1.record()
    torch.cuda.synchronize()
    prof, ops._PROFILE = ops._PROFILE, None
    agg = collections.OrderedDict()
    for a, b, fl, kern, shape in prof:
        if kern != "conv_gemm":
            continue
        k = shape
        d = agg.setdefault(k, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += a.elapsed_time(b)
        d[2] += fl
    tot = sum(v[1] for v in agg.values())
    print(f"forward {e0.elapsed_time(e1):.2f} ms; conv_gemm {tot:.2f} ms in {len(prof)} launches")
    print(f"{'M':>7} {'N':>6} {'K':>6} {'cnt':>4} {'ms':>8} {'%':>5} {'us/launch':>9} {'TFLOP/s':>8} {'tile_n':>6}")
    for (M, N, K), (c, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{M:7d} {N:6d} {K:6d} {c:4d} {ms:8.3f} {100 * ms / tot:5.1f} {1000 * ms / c:9.1f} {fl / ms / 1e9:8.1f} {ops.pick_tile_n(M, N):6d}")


if __name__ == "__main__":
    main()
