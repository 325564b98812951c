"""where does the C2 end-to-end tick go?  resident+sync vs zero-copy host exchange vs staged copies"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w

def run(tag):
    m = b2.Model(b2.asset(w.CONFIGS["c2"][0]))
    nenv = 4096
    bt = b2.Batch(m, nenv)
    w.load_config("c2", bt)
    hw, ctl, kp, kd = w.control_spec("c2", m)
    bt.set_controlled(ctl); bt.set_hw_joints(hw)
    cmd = w.commands("c2", m, np.arange(nenv))
    eff = torch.from_numpy(np.ascontiguousarray(cmd.T.astype(np.float32))).pin_memory()
    vel = torch.zeros((hw.size, nenv), dtype=torch.float32).pin_memory()
    outs = [torch.empty((hw.size, nenv), dtype=torch.float32).pin_memory() for _ in range(3)]
    args = (vel.data_ptr(), eff.data_ptr(), *[o.data_ptr() for o in outs])
    bt.write_commands(vel.numpy(), eff.numpy())
    for _ in range(50):
        bt.tick_host_raw(*args)
    K = 2000
    t0 = time.perf_counter()
    for _ in range(K):
        bt.tick_host_raw(*args)
    t_host = (time.perf_counter() - t0) / K
    for _ in range(50):
        bt.tick_resident(); bt.sync()
    t0 = time.perf_counter()
    for _ in range(K):
        bt.tick_resident(); bt.sync()
    t_res = (time.perf_counter() - t0) / K
    t0 = time.perf_counter()
    for _ in range(K):
        bt.tick_resident()
    bt.sync()
    t_async = (time.perf_counter() - t0) / K
    t0 = time.perf_counter()
    for _ in range(K):
        bt.sync()
    t_sync = (time.perf_counter() - t0) / K
    print("%s: host-exchange tick %.1f us | resident tick + sync %.1f us | resident back-to-back %.1f us | empty sync call %.2f us" % (tag, t_host * 1e6, t_res * 1e6, t_async * 1e6, t_sync * 1e6))
    bt.close()

run("zero-copy" if os.environ.get("B2_NO_ZEROCOPY") != "1" else "staged copies")
