#!/usr/bin/env python3
"""Print the metrics the profiles/ summaries quote from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys
WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__shared_mem_per_block_dynamic
launch__occupancy_limit_shared_mem launch__occupancy_limit_registers launch__grid_size launch__block_size smsp__inst_executed.sum
sm__throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__thread_inst_executed_per_inst_executed.ratio""".split()
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(f"\n## {r[idx['Kernel Name']].split('(')[0]}\n\n| metric | value |\n|---|---|")
    for w in WANT:
        if w in idx:
            print(f"| {w} | {r[idx[w]]} {units[idx[w]]} |")
