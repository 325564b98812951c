"""GPU experiment: distribution of PGS work per environment on C3 and how the kernel time depends on the iteration cap."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
it = bt.get("solver_iter")[:, 0]; ne = bt.get("nefc")[:, 0]; nc = bt.get("ncon")[:, 0]; nw = bt.get("efc_nwords")[:, 0]
print(cfg, "iters: mean %.1f median %d p90 %d p99 %d max %d  frac>=100: %.4f" % (it.mean(), np.median(it), np.quantile(it, .9), np.quantile(it, .99), it.max(), (it >= 100).mean()))
print("nefc mean %.1f max %d; ncon mean %.1f max %d; words mean %.0f max %d" % (ne.mean(), ne.max(), nc.mean(), nc.max(), nw.mean(), nw.max()))
print("hist iters", np.histogram(it, bins=[0, 1, 2, 5, 10, 20, 30, 50, 75, 99, 101])[0].tolist())
work = it * (nc + (ne > 0))
print("work = iters*blocks: mean %.0f p99 %.0f max %d" % (work.mean(), np.quantile(work, .99), work.max()))
# warp-level (4 consecutive envs in solver order): sum over warps of max work vs sum of work
order = bt.get("env_order")[:, 0]
wk = work[order[: (nenv // 4) * 4]].reshape(-1, 4)
print("sum of warp-max work / sum of mean work: %.2f" % (wk.max(1).sum() / wk.mean(1).sum()))
if os.environ.get('EXP_DISABLE'): bt.set_option('disableflags', int(os.environ['EXP_DISABLE']))
for cap in [int(x) for x in os.environ.get('EXP_CAPS', '100,20,5').split(',')]:
    bt.set_option("iterations", cap)
    for _ in range(3): tick()
    bt.sync()
    K = 10
    bt.profile_begin(K)
    for k in range(K):
        bt.l2_flush(256 << 20); tick()
    bt.sync()
    n, ms = bt.profile_end()
    it = bt.get("solver_iter")[:, 0]
    print("cap %3d: pgs %.3f ms  mean iters %.1f  (all: %s)" % (cap, ms["pgs"] / n, it.mean(), {k: round(v / n, 3) for k, v in ms.items() if v > 0}))
bt.close()
