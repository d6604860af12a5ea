"""One line of the metrics DESIGN.md cites per kernel of an .ncu-rep (read in the authoring container):
usage: python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "gpc__cycles_elapsed.max.per_second": "GHz",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor% elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor% active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed": "hmma% elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "XU%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "SM thr%",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
    "dram__bytes_read.sum": "DRAM rd",
    "dram__bytes_write.sum": "DRAM wr",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "DRAM%",
    "lts__t_sector_hit_rate.pct": "L2 hit%",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall long_sb",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall wait",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall no_inst",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall math_throttle",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall short_sb",
}
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void <unnamed>::", "")
    parts = []
    for i, h in enumerate(hdr):
        if h in WANT:
            v = r[i]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            parts.append(f"{WANT[h]} {v}{(' ' + units[i]) if units[i] and units[i] != '%' else ''}")
    print(name)
    print("    " + " | ".join(parts))
