"""Text summary of an `ncu --set full` report: one block per captured launch with the metrics the roofline discussion uses.
  python tools/ncu_report_summary.py gpurun_out/r2_targets.ncu-rep [label ...] > profiles/r2_ncu_full_hot_kernels.txt
Labels (optional) name the launches in capture order."""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
           "sm__cycles_elapsed.max"]

rep = sys.argv[1]
labels = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
iname = hdr.index("Kernel Name")
print(f"# {rep}: {len(rows) - 2} launches (ncu --set full --clock-control none; cold-cache, serialised: compare shares, not absolutes)")
for n, r in enumerate(rows[2:]):
    print(f"--- {labels[n] if n < len(labels) else 'launch ' + str(n)}")
    print(f"    {r[iname][:110]}")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"    {m:88s} {r[i]:>16s} {units[i]}")
