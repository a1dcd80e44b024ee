#!/usr/bin/env python
"""Key metrics of one kernel from an `ncu --page raw --csv` export:  python profiles/ncu_summary.py X.raw.csv"""
import csv
import sys

KEYS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
dram__bytes_read.sum dram__bytes_write.sum dram__throughput.avg.pct_of_peak_sustained_elapsed
lts__t_bytes.sum lts__throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio smsp__average_warps_issue_stalled_membar_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_drain_per_issue_active.ratio smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio""".split()

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:110], "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"  {k:95s} {d[k]:>16s} {units[hdr.index(k)]}")
