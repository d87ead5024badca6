"""Key raw metrics of an ncu report: python scripts/ncu_summary.py <report.ncu-rep>"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg','smsp__cycles_active.avg','launch__waves_per_multiprocessor','launch__occupancy_limit_registers','smsp__inst_executed.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct']
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")][:80])
    for h, u, v in zip(hdr, units, vals):
        if h in want: print(f"- {h} [{u}] = {v}")
