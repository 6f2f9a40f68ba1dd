#!/usr/bin/env python
"""Condense `ncu --page raw --csv` output into one line per launch (the numbers DESIGN.md / bench.py quote).

    python tools/ncu_summary.py gpurun_out/prof_X_raw.csv > profiles/X_summary.csv
"""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem_blocks"),
    ("launch__occupancy_limit_registers", "occ_lim_reg_blocks"),
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow([n + ("[%s]" % units[idx[c]] if c in idx and units[idx[c]] else "") for c, n in COLS])
    for d in data:
        w.writerow([d[idx[c]] if c in idx else "" for c, _ in COLS])


if __name__ == "__main__":
    main(sys.argv[1])
