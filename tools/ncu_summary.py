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


FAMILY = {"k_prior_tiles": "prior_tiles", "k_prior_groups": "prior_tiles", "k_cov_fill": "prior_tiles", "k_leaf_cov_fill": "leaf_q", "k_predict_fused": "predict_fused",
          "k_assemble_A": "assemble_A", "k_leaf_solve": "leaf_solve", "k_leaf_gram": "leaf_gram",
          "k_leaf_chol": "leaf_chol", "k_leaf_upd": "leaf_upd", "k_leaf_trsm": "leaf_trsm", "k_leaf_qobs": "leaf_qobs",
          "k_leaf_q": "leaf_q", "k_leaf_linv": "leaf_linv", "k_leaf_ut2": "leaf_ut",
          "k_fold": "fold", "k_knot_gram": "knot_gram", "k_knot_chol": "knot_chol", "k_knot_vkl": "knot_vkl",
          "k_node_chol": "node_chol", "k_node_gt": "node_gt"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(path, traffic_json=None):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow([n + ("[%s]" % units[idx[c]] if c in idx and units[idx[c]] else "") for c, n in COLS])
    for d in data:
        w.writerow([d[idx[c]] if c in idx else "" for c, _ in COLS])
    if traffic_json:
        # DRAM bytes (read + write) per kernel family over the captured launches of ONE step, for bench.py's
        # roofline.traffic (per launch = total / launches)
        import json
        fam = {}
        for d in data:
            name = d[idx["Kernel Name"]]
            key = next((v for k, v in FAMILY.items() if k in name), None)
            if key is None:
                continue
            if key == "leaf_gram" and fam.get("leaf_gram", {}).get("launches", 0) >= 1:
                key = "leaf_gram_T"
            if key == "leaf_solve" and fam.get("leaf_solve", {}).get("launches", 0) >= 1:
                key = "leaf_solve_Q"
            b = sum(float(d[idx[c]]) * UNIT.get(units[idx[c]], 1.0) for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            e = fam.setdefault(key, {"dram_bytes": 0.0, "launches": 0})
            e["dram_bytes"] += b
            e["launches"] += 1
        json.dump(fam, open(traffic_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
