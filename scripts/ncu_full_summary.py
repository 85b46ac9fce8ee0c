"""Selected metrics of an `ncu --set full` report, one column per launch (same layout as the round-1 summaries).
    python scripts/ncu_full_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Block Size", "Grid Size", "launch__cluster_dim_x", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.avg.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    idx = [i for i, h in enumerate(hdr) if h == w or h.endswith("." + w) or h.endswith(w)]
    if not idx:
        continue
    i = idx[0]
    vals = [r[i].replace("void ", "").split("(")[0] if w == "Kernel Name" else r[i] for r in data]
    print(f"{w} [{units[i]}]: " + ", ".join(vals))
