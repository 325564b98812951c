import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
cfg = "c3"
asset, nenv, _ = w.CONFIGS[cfg]
m = b2.Model(b2.asset(asset))
bt = b2.Batch(m, nenv)
w.load_config(cfg, bt)
hw, ctl, kp, kd = w.control_spec(cfg, m)
bt.set_controlled(ctl); bt.set_hw_joints(hw)
cmd = w.commands(cfg, m, np.arange(nenv))
bt.write_commands(np.zeros((hw.size, nenv), np.float32), np.ascontiguousarray(cmd.T.astype(np.float32)))
for _ in range(155): bt.tick_resident()
bt.sync()
it = bt.get("solver_iter")[:, 0]
slow = np.where(it >= 100)[0]
ev = bt.get("efc_vel")
nc = bt.get("ncon")[:, 0]; ne = bt.get("nefc")[:, 0]
ci = bt.get("contact_int"); ncm = m.nconmax
print("slow envs", slow.size)
for e in slow[:12]:
    g1 = ci[e, :nc[e]]; g2 = ci[e, ncm:ncm + nc[e]]
    print(e, "ncon", nc[e], "nefc", ne[e], "impr*scale @8,16,..96:", " ".join("%.1e" % v for v in ev[e, :12]))
    print("    pairs", list(zip(g1.tolist(), g2.tolist())))
fast = np.where((it > 10) & (it < 20))[0][:3]
for e in fast:
    print("fast", e, it[e], " ".join("%.1e" % v for v in ev[e, :3]))
