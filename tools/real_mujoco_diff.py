#!/usr/bin/env python
"""Diff of the fp64 oracle (oracle/liboracle.so) against REAL MuJoCo when one is importable (SURVEY.md section 8c:
"if a libmujoco is ever found, dlopen it and diff").  The build container and the GPU boxes of this project have no
MuJoCo (the reference downloads libmujoco 2.3.7 at build time, Makefile:3-13), so this normally reports absence; it is
wired into bench.py (`real_mujoco_diff` key of the JSON line) and tests/test_real_mujoco_pin.py so that the first
environment that does carry `import mujoco` pins the oracle without further work.

  python tools/real_mujoco_diff.py [c2|c3|c4|pendulum]

What is compared, per environment and tick, on the same MJCF and the same seeded states (solver forced to PGS with the
model's iteration count, pyramidal cones, Euler integrator, as the engine runs): qpos, qvel, qacc after every tick, the
contact count and the contact geom-id lists, nefc, and efc_force.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def diff_model(xml_path, qpos, qvel, frc, ticks=20):
    import mujoco
    import mujoco_sim_b200 as b2
    from oracle import pyoracle as orc
    mm = mujoco.MjModel.from_xml_path(xml_path)
    mm.opt.solver = mujoco.mjtSolver.mjSOL_PGS
    mm.opt.cone = mujoco.mjtCone.mjCONE_PYRAMIDAL
    mm.opt.integrator = mujoco.mjtIntegrator.mjINT_EULER
    md = mujoco.MjData(mm)
    m = b2.Model(xml_path)
    mm.opt.iterations = int(m.int("opt.iterations"))
    d = b2.Data(m)
    out = {"envs": int(qpos.shape[0]), "ticks": ticks, "model_nq_nv_nbody_ngeom": [int(m.nq), int(m.nv), int(m.nbody), int(m.ngeom)],
           "mujoco_nq_nv_nbody_ngeom": [int(mm.nq), int(mm.nv), int(mm.nbody), int(mm.ngeom)]}
    if out["model_nq_nv_nbody_ngeom"] != out["mujoco_nq_nv_nbody_ngeom"]:
        out["error"] = "compiled sizes differ"
        return out
    # model constants the compiler infers
    consts = {}
    for name in ("body_mass", "body_inertia", "body_ipos", "body_iquat", "body_invweight0", "dof_invweight0", "geom_rbound", "body_subtreemass"):
        a, r = np.array(getattr(m, name)).ravel(), np.array(getattr(mm, name)).ravel()
        consts[name] = float(np.abs(a - r).max()) if a.size == r.size and a.size else None
    out["model_const_max_abs_diff"] = consts
    worst = {"qpos": 0.0, "qvel": 0.0, "qacc": 0.0, "efc_force": 0.0}
    ncon_mismatch = nefc_mismatch = geom_mismatch = 0
    for e in range(qpos.shape[0]):
        mujoco.mj_resetData(mm, md)
        md.qpos[:] = qpos[e]; md.qvel[:] = qvel[e]; md.qfrc_applied[:] = frc[e]
        d.qpos[:] = qpos[e]; d.qvel[:] = qvel[e]; d.qfrc_applied[:] = frc[e]; d.qacc[:] = 0; d.qacc_warmstart[:] = 0
        for _ in range(ticks):
            mujoco.mj_step1(mm, md); mujoco.mj_step2(mm, md)
            orc.call("step", m, d)
            if int(md.ncon) != int(d.ncon):
                ncon_mismatch += 1
                break
            g1, g2, _ = b2.data_contacts(d)
            if md.ncon and (not np.array_equal(g1, md.contact.geom1[:md.ncon]) or not np.array_equal(g2, md.contact.geom2[:md.ncon])):
                geom_mismatch += 1
                break
            if int(md.nefc) != int(d.nefc):
                nefc_mismatch += 1
                break
            worst["qpos"] = max(worst["qpos"], float(np.abs(md.qpos - d.qpos).max()))
            worst["qvel"] = max(worst["qvel"], float(np.abs(md.qvel - d.qvel).max()))
            worst["qacc"] = max(worst["qacc"], float(np.abs(md.qacc - d.qacc).max()))
            if md.nefc:
                worst["efc_force"] = max(worst["efc_force"], float(np.abs(md.efc_force[:md.nefc] - np.array(d.efc_force)[:md.nefc]).max()))
    out["max_abs_diff"] = worst
    out["envs_with_ncon_mismatch"], out["envs_with_geom_id_mismatch"], out["envs_with_nefc_mismatch"] = ncon_mismatch, geom_mismatch, nefc_mismatch
    return out


def diff_config(cfg, nenv=16, ticks=20):
    import mujoco_sim_b200 as b2
    from mujoco_sim_b200 import workloads as w
    asset = {"pendulum": "pendulum_world.xml"}.get(cfg) or w.CONFIGS[cfg][0]
    if asset.endswith(".urdf"):
        asset = "panda7.xml"   # the MJCF the importer writes for the same arm
    m = b2.Model(b2.asset(asset))
    c = cfg if cfg in w.CONFIGS else "c1"
    qpos, qvel, frc = w.config_state(c, m, np.arange(nenv))
    return diff_model(b2.asset(asset), np.asarray(qpos), np.asarray(qvel), np.asarray(frc), ticks)


if __name__ == "__main__":
    try:
        import mujoco  # noqa: F401
    except Exception as e:
        print(json.dumps({"present": False, "why": repr(e)[:120]}))
        sys.exit(0)
    print(json.dumps(diff_config(sys.argv[1] if len(sys.argv) > 1 else "c3")))
