#!/usr/bin/env python
"""Print selected metrics from an .ncu-rep (run in the CPU container: ncu -i ... --page raw --csv)."""
import csv
import subprocess
import sys

WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__grid_size launch__block_size smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum smsp__inst_executed_op_shared_ld.sum
l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
sm__cycles_elapsed.avg lts__t_bytes.sum lts__t_sector_hit_rate.pct
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.sum sm__inst_executed_pipe_fma.sum sm__inst_executed_pipe_xu.sum
smsp__thread_inst_executed_per_inst_executed.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_selected_per_issue_active.ratio""".split()

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("==", name[:100])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:85s} {r[i]:>20s} {units[i]}")
