"""debug: find convex contacts where the fp64 batch and the oracle disagree; dump inputs"""
import ctypes as C, json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import mujoco_sim_b200 as b2
from oracle import pyoracle as orc
import test_gpu_parity as tp
m = b2.Model(xml=tp.ZOO)
nenv = 192
rng = np.random.default_rng(77)
qpos = np.tile(np.array(m.qpos0), (nenv, 1)).reshape(nenv, 7, 7)
qpos[:, :, 0] = rng.uniform(-0.22, 0.22, (nenv, 7)); qpos[:, :, 1] = rng.uniform(-0.22, 0.22, (nenv, 7)); qpos[:, :, 2] = rng.uniform(0.05, 0.3, (nenv, 7))
q = rng.normal(size=(nenv, 7, 4)); qpos[:, :, 3:] = q / np.linalg.norm(q, axis=2, keepdims=True)
qpos = qpos.reshape(nenv, 49)
bt = b2.Batch(m, nenv, precision=b2.engine.F64, export_stages=True)
bt.set("qpos", qpos); bt.tick(b2.engine.TICK_NOSOLVE); bt.sync()
ncon = bt.get("ncon")[:, 0]; ci = bt.get("contact_int"); cf = bt.get("contact"); ncm = m.nconmax
gx = bt.get("geom_xpos"); gm = bt.get("geom_xmat")
fn = orc.olib.omj_convex_pair; fn.restype = C.c_int
arr = lambda a: (C.c_double * len(a))(*[float(x) for x in a])
d = b2.Data(m)
out = []
for e in range(nenv):
    d.qpos[:] = qpos[e]
    orc.call("kinematics", m, d); orc.call("collision", m, d)
    if d.ncon != ncon[e]: continue
    for c in range(d.ncon):
        k = tp.contact_of(b2, d, c)
        pt = (int(m.geom_type[k["geom1"]]), int(m.geom_type[k["geom2"]]))
        if pt not in tp.CONVEX_PAIRS: continue
        gn = np.array([cf[e, (4 + i) * ncm + c] for i in range(3)])
        if abs(gn - k["frame"][:3]).max() < 1e-6: continue
        g1, g2 = k["geom1"], k["geom2"]
        o7 = (C.c_double * 7)()
        n = fn(pt[0], arr(gx[e, 3*g1:3*g1+3]), arr(gm[e, 9*g1:9*g1+9]), arr(np.array(m.geom_size).reshape(-1, 3)[g1]), pt[1], arr(gx[e, 3*g2:3*g2+3]), arr(gm[e, 9*g2:9*g2+9]), arr(np.array(m.geom_size).reshape(-1, 3)[g2]), C.c_double(0.0), o7)
        out.append(dict(env=e, c=c, pair=pt, gpu_dist=float(cf[e, c]), gpu_n=gn.tolist(), orc_dist=float(k["dist"]), orc_n=k["frame"][:3].tolist(),
                        orc_on_gpu_inputs=[n] + list(o7), pos1=gx[e, 3*g1:3*g1+3].tolist(), mat1=gm[e, 9*g1:9*g1+9].tolist(), size1=np.array(m.geom_size).reshape(-1, 3)[g1].tolist(),
                        pos2=gx[e, 3*g2:3*g2+3].tolist(), mat2=gm[e, 9*g2:9*g2+9].tolist(), size2=np.array(m.geom_size).reshape(-1, 3)[g2].tolist(),
                        dpos=float(abs(np.array(d.geom_xpos)[3*g1:3*g1+3] - gx[e, 3*g1:3*g1+3]).max()), dmat=float(abs(np.array(d.geom_xmat)[9*g1:9*g1+9] - gm[e, 9*g1:9*g1+9]).max())))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/mpr_cases.json", "w"), indent=1)
print(len(out), "mismatching convex contacts")
for o in out[:6]:
    print(o["pair"], o["gpu_dist"], o["orc_dist"], o["gpu_n"], o["orc_n"], o["orc_on_gpu_inputs"][:2], o["dpos"], o["dmat"])
