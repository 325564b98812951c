import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
m = b2.Model(b2.asset("ur5_tabletop.xml"))
bt = b2.Batch(m, 16384)
w.load_config("c3", bt, env_offset=0)
for _ in range(160): bt.step(1)
bt.sync()
ne = bt.get("nefc")[:,0]; nc = bt.get("ncon")[:,0]; it = bt.get("solver_iter")[:,0]; nw = bt.get("efc_nwords")[:,0]
print("nefc mean %.1f max %d p99 %d | ncon mean %.1f max %d | nwords mean %.0f max %d p99 %d" % (ne.mean(), ne.max(), np.percentile(ne,99), nc.mean(), nc.max(), nw.mean(), nw.max(), np.percentile(nw,99)))
print("iters: mean %.1f, ==100: %.3f, hist" % (it.mean(), (it==100).mean()), np.histogram(it, bins=[0,1,2,5,10,20,50,99,101])[0])
print("nefc hist", np.histogram(ne, bins=[0,8,16,32,48,64,80,96,128,161])[0])
for iters in (100, 50, 10, 1):
    bt.set_option("iterations", iters)
    bt.step(3); bt.sync()
    bt.profile_begin(5)
    for _ in range(5): bt.step(1)
    bt.sync()
    n, ms = bt.profile_end()
    print("iterations", iters, {k: round(v/n,3) for k,v in ms.items() if v>0})
