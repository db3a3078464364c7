"""Main-loop cycles per 64-wide K chunk of b200_conv_gemm, per tile width / CTA-pair mode (clock64 marks of the kernel).
Usage: python tools/gemm_chunk_rate.py [M N K].  B200_GEMM_DEBUG_MODE=1/2 isolates the MMA / the TMA side."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusers_b200 import ops, packing

M, N, K = [int(v) for v in (sys.argv[1:4] or (2048, 3840, 5120))]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, K, generator=g, device="cuda").bfloat16()
w = packing.pack_linear_weight((torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).bfloat16())
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
dbg = torch.zeros(148, 32, dtype=torch.int64, device="cuda")
print(f"M {M} N {N} K {K} chunks {K // 64} debug_mode {os.environ.get('B200_GEMM_DEBUG_MODE', '0')}")
for cm in (1, 2):
    for bn in (64, 96, 128, 160, 192, 256):
        for _ in range(2):
            dbg.zero_()
            ops.linear(x, w, N, out=out, tile_n=bn, cluster_m=cm, debug_timestamps=dbg)
        torch.cuda.synchronize()
        d = dbg.cpu()
        m_tiles, n_tiles = (M + 127) // 128, (N + bn - 1) // bn
        groups = n_tiles * ((m_tiles + cm - 1) // cm)
        n_cl = min(groups, 148 // cm)
        rates = []
        for cl in range(n_cl):
            row = d[cl * cm]
            if row[4] <= 0 or row[3] <= 0:
                continue
            my_groups = (groups - cl + n_cl - 1) // n_cl
            rates.append(float(row[4] - row[3]) / (my_groups * (K // 64)))
        rates = torch.tensor(rates)
        total = (d[:, 7] - d[:, 0]).float()
        total = total[d[:, 7] > 0]
        print(f"cm {cm} bn {bn:3d}: groups {groups:5d} on {n_cl:3d} clusters: {float(rates.median()):7.1f} cyc/chunk (min {float(rates.min()):.1f} max "
              f"{float(rates.max()):.1f}; MMA floor {bn * 2}), kernel {int(total.median())} cyc")
