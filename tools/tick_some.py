"""Settle a configuration and run a few ticks eagerly (for ncu launch lists): tools/tick_some.py c4 [nticks] [settle]"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B2_NO_GRAPH", "1")
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
cfg = sys.argv[1]; nt = int(sys.argv[2]) if len(sys.argv) > 2 else 3; settle = int(sys.argv[3]) if len(sys.argv) > 3 else 150
asset, nenv, _ = w.CONFIGS[cfg]
m = b2.Model(b2.asset(asset))
bt = b2.Batch(m, nenv)
w.load_config(cfg, bt)
if cfg == "c5":
    w.c5_init(bt, 0); tick = lambda: bt.step(1)
else:
    hw, ctl, kp, kd = w.control_spec(cfg, m)
    bt.set_controlled(ctl); bt.set_hw_joints(hw)
    if kp is not None: bt.set_pd(kp, kd)
    cmd = w.commands(cfg, m, np.arange(nenv))
    bt.write_commands(np.zeros((hw.size, nenv), np.float32), np.ascontiguousarray(cmd.T.astype(np.float32)))
    tick = bt.tick_resident
for _ in range(settle + nt): tick()
bt.sync()
print("done", cfg, bt.path_name)
