import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch, mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
for nenv in [4096, 262144]:
    m = b2.Model(b2.asset("panda7.xml"))
    bt = b2.Batch(m, nenv)
    w.load_config("c2", bt)
    hw = np.arange(7, dtype=np.int32); bt.set_controlled(np.ones(7, np.uint8)); bt.set_hw_joints(hw)
    vel = np.zeros((7, nenv), np.float32); eff = np.zeros((7, nenv), np.float32)
    pos = np.empty_like(vel); v2 = np.empty_like(vel); e2 = np.empty_like(vel)
    bt.tick_host_raw(vel.ctypes.data, eff.ctypes.data, pos.ctypes.data, v2.ctypes.data, e2.ctypes.data)
    stream = torch.cuda.ExternalStream(bt.stream)
    buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    big = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    def flush_fill():
        with torch.cuda.stream(stream): buf.fill_(1)
    def flush_memset(): bt.l2_flush(256 << 20)
    def flush_read():
        with torch.cuda.stream(stream): big.sum()
    def flush_none(): pass
    for name, fl in [("none", flush_none), ("torch_fill", flush_fill), ("memset", flush_memset), ("read_sum", flush_read), ("torch_fill", flush_fill)]:
        for _ in range(5): fl(); bt.tick_resident()
        bt.sync()
        K = 40
        st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for k in range(K):
            fl(); st[k].record(stream); bt.tick_resident(); en[k].record(stream)
        bt.sync(); torch.cuda.synchronize()
        ts = np.array([s.elapsed_time(e) for s, e in zip(st, en)])
        print(nenv, name, "mean %.4f ms  median %.4f  min %.4f  max %.4f" % (ts.mean(), np.median(ts), ts.min(), ts.max()), flush=True)
