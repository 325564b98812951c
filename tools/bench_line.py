"""Print a short summary of a bench.py JSON line: python tools/bench_line.py gpurun_out/bench_c3.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print(d["config"]["workload"][:40], "| n_gpus", d["n_gpus"], "| %.3f M/s  %.3f ms/tick  e2e %.3f M/s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6))
if r:
    print("  kernels", {k: round(v, 3) for k, v in r["kernel_ms_all"].items() if v > 0}, "| hbm_frac %.5f fp32_frac %.5f traffic %s" % (r["frac"], r["fp32_frac"], r.get("traffic")))
print("  cpu", d.get("cpu_baseline")); print("  clocks", d.get("clocks"))
dr = d.get("drift")
if dr: print("  drift", {k: dr[k] for k in dr if k in ("ticks", "median", "max", "frac_below_1e-4", "first_contact_set_mismatch_tick")})
print("  exch", d.get("obs_exchange"))
for c, v in d.get("configs", {}).items():
    print(" ", c, {k: (round(x, 4) if isinstance(x, float) else x) for k, x in v.items() if k not in ("workload", "kernel_ms_all")})
    if "kernel_ms_all" in v: print("     ", {k: round(x, 3) for k, x in v["kernel_ms_all"].items() if x > 0})
