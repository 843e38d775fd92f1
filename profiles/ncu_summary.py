#!/usr/bin/env python
"""Prints selected metrics from `ncu -i <rep> --page raw --csv` (one column per launch)."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "local_load", "local_store", "smsp__inst_executed_op_local",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_average_data_bytes_per_sector_mem_global",
    "derived__memory_l1_wavefronts_shared_excessive", "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
]


def main(path, extra):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = KEYS + extra
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, r):
            base = h.split(".", 2)[-1] if h.split(".")[0].isupper() or "Triage" in h else h
            if any(k in h for k in keys):
                print(f"   {h} [{u}] = {v}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
