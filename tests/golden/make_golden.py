#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ from the fp64 oracle (oracle/liboracle.so).

The reference ships no golden vectors and its arithmetic (MuJoCo 2.3.7) is not available, so these
fixtures pin the ORACLE ITSELF: they were produced once by this script and are committed; test_golden.py checks that the
oracle still reproduces them on CPU (a regression pin for the checker) and that the fp64 / fp32 CUDA batches reach them
on the GPU.  Each fixture: seeded states of a workload config, the state after `ticks` control ticks (controller +
inverse + PD where the config has it), contact / row counts of the final tick.  Re-run only when the oracle is changed on
purpose:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [("c2", 256, 50), ("c3", 256, 60), ("c4", 256, 40)]


def rollout(cfg, nenv, ticks):
    import mujoco_sim_b200 as b2
    from mujoco_sim_b200 import workloads as w
    from oracle import pyoracle as orc
    m = b2.Model(b2.asset(w.CONFIGS[cfg][0]))
    q0, v0, _ = w.config_state(cfg, m, np.arange(nenv))
    hw, ctl, kp, kd = w.control_spec(cfg, m)
    dadr = np.array(m.jnt_dofadr)[hw]
    ddq = np.zeros((nenv, m.nv)); ddq[:, dadr] = w.commands(cfg, m, np.arange(nenv))
    pd = {}
    if kp is not None:
        kpv, kdv = np.zeros(m.nv), np.zeros(m.nv); kpv[dadr], kdv[dadr] = kp, kd
        pd = {"pd_kp": kpv, "pd_kd": kdv}
    q, v = np.ascontiguousarray(q0).copy(), np.ascontiguousarray(v0).copy()
    ws = np.zeros((nenv, m.nv)); finv = np.zeros((nenv, m.nv))
    orc.tick_batch(m, [b2.Data(m) for _ in range(min(8, os.cpu_count() or 1))], ticks, q, v, ws, None, ddq, np.zeros((nenv, m.nv)), ctl, True, finv, **pd)
    ncon, nefc = np.zeros(nenv, np.int32), np.zeros(nenv, np.int32)
    d = b2.Data(m)
    for e in range(nenv):
        d.qpos[:] = q[e]; d.qvel[:] = v[e]
        orc.call("fwdPosition", m, d)
        ncon[e], nefc[e] = d.ncon, d.nefc
    return dict(qpos0=q0, qvel0=v0, qpos=q, qvel=v, qfrc_inverse=finv, ncon=ncon, nefc=nefc, ticks=np.int32(ticks))


def main():
    for cfg, nenv, ticks in CASES:
        out = os.path.join(HERE, "%s_tick.npz" % cfg)
        np.savez_compressed(out, **rollout(cfg, nenv, ticks))
        print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
