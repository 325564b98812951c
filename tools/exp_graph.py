"""GPU experiment: graph-replayed tick time of a config (wall clock around K ticks, L2 flushed before every tick),
right after the settle ticks.  B2_SUBBATCH / B2_SUBBATCH_MIN select the window cut."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
asset, nenv, _ = w.CONFIGS[cfg]
m = b2.Model(b2.asset(asset))
bt = b2.Batch(m, nenv)
w.load_config(cfg, bt)
if cfg == "c5":
    w.c5_init(bt, 0)
    tick = lambda: bt.step(1)
else:
    hw, ctl, kp, kd = w.control_spec(cfg, m)
    bt.set_controlled(ctl); bt.set_hw_joints(hw)
    if kp is not None: bt.set_pd(kp, kd)
    cmd = w.commands(cfg, m, np.arange(nenv))
    bt.write_commands(np.zeros((hw.size, nenv), np.float32), np.ascontiguousarray(cmd.T.astype(np.float32)))
    tick = bt.tick_resident
for _ in range(155): tick()
bt.sync()
res = []
for rep in range(5):
    K = 40
    tot = 0.0
    for k in range(K):
        bt.l2_flush(256 << 20); bt.sync()
        t0 = time.perf_counter(); tick(); bt.sync(); tot += time.perf_counter() - t0
    res.append(tot / K * 1e3)
print("%s subbatch %s: tick %.3f ms (min %.3f) -> %.2f M env-steps/s; mean iters %.1f" % (cfg, os.environ.get("B2_SUBBATCH", "default"), np.median(res), min(res), nenv / np.median(res) / 1e3, bt.get("solver_iter").mean()))
bt.close()
