"""GPU experiment: tensor-core projection (B2_TC_PROJECT) against the FFMA path on C4: results and kernel times."""
import sys, os, subprocess, json, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import mujoco_sim_b200 as b2
    from mujoco_sim_b200 import workloads as w
    cfg = "c4"
    asset, nenv, _ = w.CONFIGS[cfg]
    m = b2.Model(b2.asset(asset))
    bt = b2.Batch(m, nenv)
    w.load_config(cfg, bt)
    hw, ctl, kp, kd = w.control_spec(cfg, m)
    bt.set_controlled(ctl); bt.set_hw_joints(hw); bt.set_pd(kp, kd)
    cmd = w.commands(cfg, m, np.arange(nenv))
    bt.write_commands(np.zeros((hw.size, nenv), np.float32), np.ascontiguousarray(cmd.T.astype(np.float32)))
    for _ in range(155): bt.tick_resident()
    bt.sync()
    K = 10
    bt.profile_begin(K)
    for k in range(K):
        bt.l2_flush(256 << 20); bt.tick_resident()
    bt.sync()
    n, ms = bt.profile_end()
    out = {"ms": {k: v / n for k, v in ms.items()}, "qpos": bt.get("qpos")[:256].tolist(), "force": bt.get("efc_force")[:256].tolist(),
           "qfrc_constraint": bt.get("qfrc_constraint")[:256].tolist(), "nefc": bt.get("nefc")[:256, 0].tolist()}
    json.dump(out, open(sys.argv[2], "w"))
    sys.exit(0)
res = {}
for tag, val in (("ffma", "0"), ("tf32x3", "1"), ("tf32x1", "2")):
    env = dict(os.environ, B2_TC_PROJECT=val)
    subprocess.check_call([sys.executable, __file__, "child", "/tmp/tc_%s.json" % tag], env=env)
    res[tag] = json.load(open("/tmp/tc_%s.json" % tag))
    print(tag, {k: round(v, 3) for k, v in res[tag]["ms"].items() if v > 0}, "tick %.3f ms" % sum(res[tag]["ms"].values()))
ref = res["ffma"]
for tag in ("tf32x3", "tf32x1"):
    r = res[tag]
    same_rows = np.array_equal(ref["nefc"], r["nefc"])
    for f in ("qpos", "qfrc_constraint", "force"):
        a, b = np.array(ref[f]), np.array(r[f])
        print(tag, f, "max |diff| after 165 ticks: %.3e (scale %.2e)" % (np.abs(a - b).max(), np.abs(a).max()), "same row counts:", same_rows)
