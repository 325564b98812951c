import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, mujoco_sim_b200 as b2, torch
from mujoco_sim_b200 import workloads as w
m = b2.Model(b2.asset("ur5_tabletop.xml"))
for prec in [b2.engine.F32, b2.engine.F64]:
    bt = b2.Batch(m, 16384, precision=prec)
    w.load_config("c3", bt)
    bt.step(150); bt.sync()
    it = bt.get("solver_iter")[:,0]; ne = bt.get("nefc")[:,0]
    print("prec", prec, "iters: mean %.1f median %d p90 %d p99 %d max %d frac100 %.3f" % (it.mean(), np.median(it), np.percentile(it,90), np.percentile(it,99), it.max(), (it>=100).mean()))
    print("  nefc mean %.1f max %d; warp-max iters mean %.1f" % (ne.mean(), ne.max(), it.reshape(-1,32).max(axis=1).mean()))
    for iters in [100, 50, 20, 5, 1]:
        bt.set_option("iterations", iters)
        bt.step(3); bt.sync()
        bt.profile_begin(10)
        bt.step(10); bt.sync()
        n, ms = bt.profile_end()
        print("  iterations cap", iters, {k: round(v/n,3) for k,v in ms.items() if v>0})
    bt.close()
