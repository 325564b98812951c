import sys, os, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import mujoco_sim_b200 as b2
from oracle import pyoracle as orc
from rewrites import rewrite
txt = open(b2.asset("panda7.xml")).read()
xml = rewrite(txt, "link1", True, (0.2, 0.1, 0.4, 0, 0, 0.5), ("lin_odom_x_joint", "lin_odom_y_joint", "ang_odom_z_joint"))
m = b2.Model(xml=xml, basedir=b2.asset(""))
nenv = 4
rng = np.random.default_rng(3)
q = np.tile(np.array(m.qpos0), (nenv, 1)) + rng.uniform(-0.3, 0.3, (nenv, m.nq)); v = rng.uniform(-1, 1, (nenv, m.nv))
F = ["xpos", "xquat", "subtree_com", "cdof", "qM", "qfrc_passive", "qfrc_bias", "qacc_smooth", "qacc"]
bt = b2.Batch(m, nenv, precision=b2.engine.F64, export_stages=True)
bt.set("qpos", q); bt.set("qvel", v); bt.forward(); bt.sync()
d = b2.Data(m)
for e in range(2):
    d.qpos[:] = q[e]; d.qvel[:] = v[e]; d.qacc[:] = 0; d.qacc_warmstart[:] = 0; d.qfrc_applied[:] = 0
    orc.call("forward", m, d)
    for f in F:
        g = bt.get(f)[e]; r = np.array(d.array(f))
        print(e, f, "max|diff| %.3e" % np.abs(g - r).max(), "at", int(np.abs(g - r).argmax()))
print(bt.path_name)
