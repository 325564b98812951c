"""GPU experiment: what the 100-iteration environments of C3 look like (islands, blocks per island)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
asset, nenv, _ = w.CONFIGS[cfg]
m = b2.Model(b2.asset(asset))
bt = b2.Batch(m, nenv)
w.load_config(cfg, bt)
hw, ctl, kp, kd = w.control_spec(cfg, m)
bt.set_controlled(ctl); bt.set_hw_joints(hw)
cmd = w.commands(cfg, m, np.arange(nenv))
bt.write_commands(np.zeros((hw.size, nenv), np.float32), np.ascontiguousarray(cmd.T.astype(np.float32)))
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 155): bt.tick_resident()
bt.sync()
it = bt.get("solver_iter")[:, 0]; ncon = bt.get("ncon")[:, 0]; nisl = bt.get("nisl")[:, 0]
io = bt.get("isl_off"); ie = bt.get("isl_end")
words = (ie - io)
mx = np.array([words[e, :nisl[e]].max() if nisl[e] > 0 else 0 for e in range(nenv)])
heavy = np.where(it >= 100)[0]
print("envs", nenv, "heavy", heavy.size, "mean ncon all %.1f heavy %.1f" % (ncon.mean(), ncon[heavy].mean()))
print("largest island words: all mean %.0f p50 %d p99 %d max %d | heavy mean %.0f min %d max %d" % (mx.mean(), np.median(mx), np.quantile(mx, .99), mx.max(), mx[heavy].mean(), mx[heavy].min(), mx[heavy].max()))
print("iters hist", np.histogram(it, bins=[0, 1, 5, 10, 20, 30, 50, 75, 99, 101])[0].tolist()); nb_max = (mx + 95) // 96; print("blocks in largest island (approx words/96) hist:", np.bincount(nb_max.astype(int))[:16].tolist(), "heavy:", np.bincount(nb_max[heavy].astype(int))[:16].tolist())
for e in heavy[:8]:
    print(" env", e, "ncon", ncon[e], "nisl", nisl[e], "island words", words[e, :nisl[e]].astype(int).tolist())
bt.close()
