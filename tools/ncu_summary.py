"""One line per kernel launch from an .ncu-rep (ncu -i REP --page raw --csv), keeping the counters DESIGN.md argues with.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [header text] > profiles/x_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEEP = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
raw = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"]).decode()
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("Kernel Name=" + d.get("Kernel Name", "?") + "; " + "; ".join(f"{k}={d[k]} {u[k]}".strip() for k in KEEP if k in d) + "\n")
