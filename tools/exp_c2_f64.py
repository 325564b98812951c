import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
from oracle import pyoracle as orc
m = b2.Model(b2.asset(w.CONFIGS["c2"][0]))
for prec in (b2.engine.F32, b2.engine.F64):
    nenv = 4096
    bt = b2.Batch(m, nenv, precision=prec)
    w.load_config("c2", bt)
    hw, ctl, kp, kd = w.control_spec("c2", m)
    bt.set_controlled(ctl); bt.set_hw_joints(hw)
    cmd = w.commands("c2", m, np.arange(nenv))
    bt.write_commands(np.zeros((hw.size, nenv), np.float32), np.ascontiguousarray(cmd.T.astype(np.float32)))
    for _ in range(20): bt.tick_resident()
    bt.sync()
    stream = torch.cuda.ExternalStream(bt.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    e0.record(stream)
    for _ in range(K): bt.tick_resident()
    e1.record(stream); bt.sync(); torch.cuda.synchronize()
    print("prec", prec, bt.path_name, "us/tick %.2f" % (e0.elapsed_time(e1) / K * 1e3), "env-steps/s %.1f M" % (nenv * K / e0.elapsed_time(e1) / 1e3))
    bt.close()
    # drift over 1000 ticks, 64 envs, plain step with applied torques
    bt = b2.Batch(m, 64, precision=prec)
    q, v, f, _ = w.load_config("c2", bt)
    rq, rv, rf = (np.ascontiguousarray(x, np.float64).copy() for x in (q, v, f))
    pool = [b2.Data(m) for _ in range(8)]
    ws = np.zeros((64, m.nv))
    bt.step(1000); bt.sync()
    orc.tick_batch(m, pool, 1000, rq, rv, ws, rf)
    rel = np.linalg.norm(bt.get("qpos") - rq, axis=1) / np.maximum(np.linalg.norm(rq, axis=1), 1e-12)
    print("   drift@1000: median %.2e max %.2e frac<1e-4 %.3f" % (np.median(rel), rel.max(), (rel < 1e-4).mean()))
    bt.close()
