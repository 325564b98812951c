"""One steady-state tick inside a cudaProfilerStart / Stop range, for `ncu --replay-mode range --cache-control none`: the
DRAM traffic of the WHOLE tick as it runs in sequence (kernel-by-kernel captures flush the caches before every kernel and
count every intermediate array as DRAM traffic; in the real tick most of them are still in the 126 MB L2).
Usage: tick_range.py c3 [settle]"""
import ctypes, sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B2_NO_GRAPH", "1")
import torch
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
cfg = sys.argv[1]; settle = int(sys.argv[2]) if len(sys.argv) > 2 else 150
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
for _ in range(settle): tick()
bt.sync()
bt.l2_flush(256 << 20); bt.sync()      # the tick starts from a cold L2, as in the bench
torch.cuda.profiler.start()
tick()
bt.sync()
torch.cuda.profiler.stop()
print("done")
