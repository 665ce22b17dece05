#!/usr/bin/env python
"""Digest of an `ncu --page raw --csv` export: one block per captured launch with the metrics that explain it."""
import csv
import sys

KEYS = [
    ("time_us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
    ("occ_lim_regs", "launch__occupancy_limit_registers"), ("occ_lim_smem", "launch__occupancy_limit_shared_mem"),
    ("occ_lim_warps", "launch__occupancy_limit_warps"), ("waves", "launch__waves_per_multiprocessor"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_pipe_pct", "sm__pipe_tensor_subpipe_tmem_cycles_active.avg.pct_of_peak_sustained_active"),
    ("sm_thr_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram_thr_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_thr_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_thr_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram_rd_B", "dram__bytes_read.sum"), ("dram_wr_B", "dram__bytes_write.sum"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_noinst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_membar", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"),
    ("stall_sleep", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("stall_dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
    ("smem_bank_conf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("inst", "smsp__inst_executed.sum"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[ix["Kernel Name"]][:70]
    print("%s grid %s block %s" % (name, r[ix["Grid Size"]], r[ix["Block Size"]]))
    out = []
    for short, k in KEYS:
        if k in ix and r[ix[k]] not in ("", "n/a"):
            try:
                v = float(r[ix[k]].replace(",", ""))
                out.append("%s=%.4g" % (short, v))
            except ValueError:
                pass
    print("   " + " ".join(out))
