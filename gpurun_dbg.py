import sys, os, json, subprocess; sys.path.insert(0,'/root/repo')
import numpy as np, mujoco_sim_b200 as b2
from oracle import pyoracle as orc
model = b2.asset("mobile_arm.xml")
def oracle(nticks):
    m = b2.Model(model); d = b2.Data(m); nv = m.nv; arm=[3,4,5]
    ctl = np.zeros(nv, np.uint8); ctl[arm]=1; ddq=np.zeros(nv); dq=np.zeros(nv); eff=np.zeros(3)
    for t in range(nticks):
        orc.call("step1", m, d); orc.controller(m, d, ddq, dq, ctl); orc.call("inverse", m, d)
        eff = np.array(d.qfrc_inverse)[arm]; q, qd = np.array(d.qpos)[arm], np.array(d.qvel)[arm]
        for i in range(3):
            vcmd = 0.2 if (i == 1 and t % 10 == 3) else 0.0
            if abs(vcmd) > 1e-15: dq[arm[i]] = vcmd
            else: ddq[arm[i]] = 20.0 * (0.3 * (i + 1) - q[i]) - 4.0 * qd[i]
        orc.call("step2", m, d)
        orc.set_odom_vels(m, d, [0, 1, -1], [-1, -1, 2], [-1, -1, 2], [0.4, -0.1, 0, 0, 0, 0.3])
    return eff, np.array(d.qacc), np.array(d.qfrc_passive)[arm], np.array(d.qfrc_bias)[arm]
for n in [1, 2, 3, 4, 5, 6, 10]:
    res = subprocess.run(["tests/_compat/compat_tick", model, str(n)], capture_output=True, text=True, env=dict(os.environ, B2_PRECISION="8"))
    got = json.loads(res.stdout.strip().splitlines()[-1])
    e, qa, pas, bias = oracle(n)
    print(n, "gpu", np.round(got["effort"], 6), "oracle", np.round(e, 6))
