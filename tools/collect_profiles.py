#!/usr/bin/env python
"""Copies what tools/gpurun_prof.sh left under gpurun_out/ into profiles/ (tracked): bench JSON lines, the reference arm,
launch lists, ncu summaries and the per-kernel DRAM traffic.  bench.py reads roofline.traffic from profiles/traffic_<cfg>.json
as it was at run time (the previous capture); the lines stored here get the value of the capture made in the same run."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r01"
SLOT = {"k_pgs_block": "pgs", "k_chain_team<7>": "smooth", "k_chain<7>": "smooth"}


def main():
    for c in ("c2", "c3", "c4"):
        for src, dst in (("traffic_%s.json" % c, "traffic_%s.json" % c), ("launches_%s.csv" % c, "%s_launches_%s.csv" % (ROUND, c)),
                         ("ncu_%s_summary.txt" % c, "%s_ncu_%s_summary.txt" % (ROUND, c))):
            if os.path.exists(os.path.join(SRC, src)):
                shutil.copy(os.path.join(SRC, src), os.path.join(DST, dst))
    for c in ("c2", "c2_256k", "c3", "c4", "c5"):
        p = os.path.join(SRC, "bench_%s.json" % c)
        if not os.path.exists(p):
            continue
        d = json.loads(open(p).read().strip().splitlines()[-1])
        tf = os.path.join(DST, "traffic_%s.json" % c.split("_")[0])
        if c in ("c2", "c3", "c4") and os.path.exists(tf):
            t = json.load(open(tf))
            k = SLOT.get(d["roofline"]["kernel"])
            if k in t:
                d["roofline"]["traffic"] = t[k]
                d["roofline"]["traffic_source"] = ("profiles/traffic_%s.json: ncu --set full capture of the same build (tools/gpurun_prof.sh), "
                                                   "dram__bytes_read.sum + dram__bytes_write.sum per launch" % c)
        else:
            d["roofline"]["traffic"] = None
        with open(os.path.join(DST, "%s_bench_%s.json" % (ROUND, c)), "w") as f:
            f.write(json.dumps(d) + "\n")
        r = d["roofline"]
        print(c, "%.3f M env-steps/s" % (d["value"] / 1e6), "%.4f ms" % d["ms_per_step"], "e2e %.3f M" % (d["e2e"]["value"] / 1e6), "| dominant", r["kernel"],
              "%.4f ms" % r["kernel_ms"], "frac %.4f" % r["frac"], "| cpu %.3f M" % (d.get("cpu_baseline", {}).get("value", 0) / 1e6),
              {k: round(v, 3) for k, v in r["kernel_ms_all"].items() if v > 0.01}, d["clocks"]["reasons"])
    for src, dst in (("bench_ref_c2.json", "%s_bench_reference_arm_c2.json" % ROUND), ("pytest_gpu.log", "%s_pytest_gpu.log" % ROUND)):
        if os.path.exists(os.path.join(SRC, src)):
            shutil.copy(os.path.join(SRC, src), os.path.join(DST, dst))


if __name__ == "__main__":
    main()
