"""Summarise an .ncu-rep (read here with `ncu -i`): per captured launch duration, DRAM traffic, throughput,
occupancy and the top warp-stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, units = r[0], r[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# {rep}")
for row in r[2:]:
    print("kernel:", row[h.index("Kernel Name")])
    for k in keys:
        if k in h:
            print(f"  {k:70s} {row[h.index(k)]:>16s} {units[h.index(k)]}")
    st = []
    for i, k in enumerate(h):
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                st.append((float(row[i]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    tot = sum(x for x, _ in st) or 1.0
    print("  warp stalls:", ", ".join(f"{k} {100 * x / tot:.1f}%" for x, k in st[:6]))
