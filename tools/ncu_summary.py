"""Compact per-launch table from an ncu report: python tools/ncu_summary.py <file.ncu-rep> > summary.csv"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
cols = [c for c in want if c in hdr]
w = csv.writer(sys.stdout)
w.writerow([f"{c} [{units[hdr.index(c)]}]" if units[hdr.index(c)] else c for c in cols])
for r in rows[2:]:
    out = []
    for c in cols:
        v = r[hdr.index(c)]
        if c == "Kernel Name":
            v = v.split("(")[0][:60]
        out.append(v)
    w.writerow(out)
