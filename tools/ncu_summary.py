#!/usr/bin/env python3
"""Summarise an .ncu-rep (run where ncu is installed; no GPU needed): python tools/ncu_summary.py rep [rep...]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_warps", "occ_lim_warps"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ%"),
    ("smsp__cycles_active.avg", "smsp_cycles_active"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock_hz"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_inst%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fma_cycles%"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fmaheavy_cycles%"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "pipe_fmaheavy_cycles%(elapsed)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu%"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"), ("l1tex__t_sector_hit_rate.pct", "l1_hit%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum", "global_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_data_pipe%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall_imc_miss"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall_sleeping"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
    ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall_drain"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        seen = set()
        print("#", rep)
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].split("::")[-1]
            if name in seen:
                continue
            seen.add(name)
            print("##", name)
            for key, label in WANT:
                if key in idx:
                    v = r[idx[key]]
                    try:
                        v = "%.4g" % float(v.replace(",", ""))
                    except ValueError:
                        pass
                    print("  %-24s %s %s" % (label, v, units[idx[key]]))


if __name__ == "__main__":
    main()
