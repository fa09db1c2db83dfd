"""Print the headline metrics of every launch in an .ncu-rep (via `ncu --page raw --csv`).
usage: python profiles/ncu_summary.py <report.ncu-rep>"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
for r in rows[2:]:
    print("---", r[hdr.index("Kernel Name")][:60])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k} [{units[i]}] = {r[i]}")
