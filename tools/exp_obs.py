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
_, h = bt.obs_create(1, 0); bt.obs_attach(handles=h)
for on in (False, True, False, True):
    bt.obs_enable(on)
    for _ in range(3): bt.tick_resident()
    bt.sync()
    K = 10
    bt.profile_begin(K)
    for k in range(K):
        bt.l2_flush(256 << 20); bt.tick_resident()
    bt.sync()
    n, ms = bt.profile_end()
    print("obs", on, {k: round(v / n, 3) for k, v in ms.items() if v > 0}, "it", bt.get("solver_iter").mean())
