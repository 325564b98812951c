#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`, no GPU needed) into the per-kernel text blocks
kept under profiles/, and optionally write the per-launch DRAM traffic of each kernel as JSON (bench.py reads
profiles/traffic_<config>.json for roofline.traffic).  Usage: ncu_summary.py rep.ncu-rep [--traffic out.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(hdr)}
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]]
        print("----")
        print("%-80s %s" % ("Kernel Name", name))
        for k in KEYS:
            if k in col:
                print("%-80s %s %s" % (k, r[col[k]], units[col[k]]))
        try:
            rd = float(r[col["dram__bytes_read.sum"]]) * UNIT.get(units[col["dram__bytes_read.sum"]], 1.0)
            wr = float(r[col["dram__bytes_write.sum"]]) * UNIT.get(units[col["dram__bytes_write.sum"]], 1.0)
            short = name.split("<")[0].replace("void ", "").replace("b2::", "")
            key = {"k_smooth": "smooth", "k_chain": "smooth", "k_chain_team": "smooth", "k_collide": "collide", "k_make_constraint": "make_constraint",
                   "k_make_rows": "make_constraint", "k_make_blocks": "make_constraint", "k_solve_rows": "make_constraint", "k_pgs_block": "pgs", "k_pgs_island": "pgs", "k_order_envs": "pgs", "k_integrate": "integrate", "k_dense_minv": "make_constraint", "k_project_tc": "make_constraint"}.get(short, short)
            traffic[key] = traffic.get(key, 0) + int(rd + wr)   # bench.py's make_constraint slot = k_make_rows (+ k_solve_rows) + k_make_blocks
        except Exception:
            pass
    traffic["tick_total"] = sum(v for k, v in traffic.items() if k != "tick_total")
    if "--traffic" in sys.argv:
        with open(sys.argv[sys.argv.index("--traffic") + 1], "w") as f:
            json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main()
