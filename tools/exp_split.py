"""GPU experiment: one device, the batch of a config cut into S sub-batches that tick side by side on their own streams
(latency-bound kernels of one sub-batch overlap the solver tail of another).  Wall clock around K ticks, sync on both sides."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
asset, nenv, _ = w.CONFIGS[cfg]
m = b2.Model(b2.asset(asset))
for S in [int(x) for x in os.environ.get("EXP_SPLITS", "1,2,4").split(",")]:
    n = nenv // S
    bts = []
    for s in range(S):
        bt = b2.Batch(m, n)
        w.load_config(cfg, bt, env_offset=s * n)
        if cfg == "c5":
            w.c5_init(bt, s * n)
        else:
            hw, ctl, kp, kd = w.control_spec(cfg, m)
            bt.set_controlled(ctl); bt.set_hw_joints(hw)
            if kp is not None: bt.set_pd(kp, kd)
            cmd = w.commands(cfg, m, np.arange(s * n, (s + 1) * n))
            bt.write_commands(np.zeros((hw.size, n), np.float32), np.ascontiguousarray(cmd.T.astype(np.float32)))
        bts.append(bt)
    tick = (lambda bt: bt.step(1)) if cfg == "c5" else (lambda bt: bt.tick_resident())
    for _ in range(155):
        for bt in bts: tick(bt)
    for bt in bts: bt.sync()
    best = []
    for rep in range(5):
        K = 50
        t0 = time.perf_counter()
        for _ in range(K):
            for bt in bts: tick(bt)
        for bt in bts: bt.sync()
        best.append((time.perf_counter() - t0) / K * 1e3)
    print("%s S=%d x %d envs: tick %.3f ms (min %.3f)  -> %.2f M env-steps/s" % (cfg, S, n, np.median(best), min(best), nenv / np.median(best) / 1e3), flush=True)
    for bt in bts: bt.close()
