"""Per-CTA clock64 timeline of one b200_conv_gemm launch (debug_timestamps): where do the cycles go?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusers_b200 import ops, packing

M, N, K = [int(v) for v in (sys.argv[1:4] or (2048, 1280, 1280))]
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, K, generator=g, device="cuda").bfloat16()
w = packing.pack_linear_weight((torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).bfloat16())
b = torch.randn(N, generator=g, device="cuda").bfloat16()
r = torch.randn(M, N, generator=g, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
dbg = torch.zeros(148, 32, dtype=torch.int64, device="cuda")
use_res = os.environ.get('NORES') is None
if not use_res:
    r = None
for _ in range(3):
    ops.linear(x, w, N, bias=b, residual=r, out=out, tile_n=bn, debug_timestamps=dbg)
torch.cuda.synchronize()
flush = torch.empty(64 << 20, dtype=torch.int32, device="cuda")
mode = os.environ.get("MODE", "cold")  # cold | hotact (activations in L2, weights cold) | pf (+ weights prefetched by the previous launch) | hot
xs = torch.randn(256, 256, device="cuda").bfloat16()
ws = packing.pack_linear_weight(torch.randn(256, 256, device="cuda").bfloat16())
names = ["entry", "prologue done", "first TMA issued", "first data landed", "last MMA committed", "accumulator ready (epi)",
         "epilogue stores issued", "exit", "c0 tmem loaded", "c0 math+sts done", "", "c0 store issued",
         "c1 tmem loaded", "c1 math+sts done", "", "c1 store issued"] + [""] * 8 + ["producer: first tile decoded", "producer: first stage free"]
for rep in range(2):
    if mode != "hot":
        flush.zero_()
    if mode in ("hotact", "pf"):
        x.float().sum()
        if r is not None:
            r.float().sum()
    if mode == "pf":
        ops._PLAN = type("P", (), {"_step": lambda self, w_: (w.data_ptr(), w.numel() * 2)})()
        ops.linear(xs, ws, 256)
        ops._PLAN = None
    dbg.zero_()
    ops.linear(x, w, N, bias=b, residual=r, out=out, tile_n=bn, debug_timestamps=dbg)
    torch.cuda.synchronize()
    d = dbg.cpu()
    d = d[d[:, 7] > 0]
    if rep == 0:
        continue
    span = int(d[:, 7].max() - d[:, 0].min())
    print(f"mode {mode}: {d.shape[0]} CTAs; first entry -> last exit {span} cycles (clock64 is per SM: indicative only); "
          f"medians of (t_i - t_entry) in cycles:")
    for i in range(1, len(names)):
        v = (d[:, i] - d[:, 0]).float()
        v = v[d[:, i] > 0]
        if len(v) and names[i]:
            print(f"   {names[i]:28s} median {int(v.median()):7d}  min {int(v.min()):7d}  max {int(v.max()):7d}")
