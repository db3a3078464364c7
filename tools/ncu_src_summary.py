"""Summarise the source page of an ncu report: top SASS instructions by warp-state samples with their dominant stall reasons,
and samples aggregated per CUDA source line when the report carries -lineinfo.
  python tools/ncu_src_summary.py gpurun_out/x.ncu-rep [launch index] [top n] [--range lo hi]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
# the page holds one table per view (CUDA-C lines when the report has -lineinfo sources, then SASS), each with its own header row
tables, cur = [], None
for r in rows[1:]:
    if r and r[0] in ("Address", "#", "Line") and "# Samples" in r:
        cur = dict(hdr=r, rows=[])
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for t in tables:
    hdr = t["hdr"]
    isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for n, r in enumerate(t["rows"]):
        if len(r) < len(hdr):
            continue
        s_ = int(r[ismp] or 0)
        st = sorted([(int(r[i] or 0), h) for i, h in stall], reverse=True)[:3]
        data.append((n, s_, r[isrc].strip(), int(r[iex] or 0), st))
    tot = sum(d[1] for d in data)
    print(f"--- view keyed by {hdr[0]!r}: {len(data)} rows, {tot} samples")
    agg = {}
    for d in data:
        for c, h in d[4]:
            agg[h] = agg.get(h, 0) + c
    print("stall totals (top-3 per row only):", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for d in sorted(data, key=lambda x: -x[1])[:top]:
        print(f"{d[0]:5d} {d[1]:6d} {100.0 * d[1] / max(tot, 1):5.1f}%  {d[2][:64]:64s} x{d[3]:<8d} {[(h, c) for c, h in d[4] if c]}")
