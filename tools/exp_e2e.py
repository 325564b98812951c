"""GPU experiment: the control tick through host buffers (b2_tick_host: zero-copy on pinned memory, or staged copies with
B2_NO_ZEROCOPY=1) against the resident tick, interleaved block by block so that both see the same phase of the simulation."""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
asset, nenv, _ = w.CONFIGS[cfg]
m = b2.Model(b2.asset(asset))
bt = b2.Batch(m, nenv)
w.load_config(cfg, bt)
hw, ctl, kp, kd = w.control_spec(cfg, m)
bt.set_controlled(ctl); bt.set_hw_joints(hw)
if kp is not None: bt.set_pd(kp, kd)
cmd = w.commands(cfg, m, np.arange(nenv))
eff = torch.from_numpy(np.ascontiguousarray(cmd.T.astype(np.float32))).pin_memory()
vel = torch.zeros((hw.size, nenv), dtype=torch.float32).pin_memory()
outs = [torch.empty((hw.size, nenv), dtype=torch.float32).pin_memory() for _ in range(3)]
args = (vel.data_ptr(), eff.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr())
bt.write_commands(vel.numpy(), eff.numpy())
for _ in range(155): bt.tick_resident()
bt.sync()
K = 25
res, e2e = [], []
for rep in range(8):
    t = 0.0
    for k in range(K):
        bt.l2_flush(256 << 20); bt.sync()
        t0 = time.perf_counter(); bt.tick_resident(); bt.sync(); t += time.perf_counter() - t0
    res.append(t / K * 1e3)
    t = 0.0
    for k in range(K):
        bt.l2_flush(256 << 20); bt.sync()
        t0 = time.perf_counter(); bt.tick_host_raw(*args); t += time.perf_counter() - t0
    e2e.append(t / K * 1e3)
print("%s zero-copy %s: resident %.3f ms, host buffers %.3f ms (+%.0f us) per tick; blocks %s / %s" % (cfg, "off" if os.environ.get("B2_NO_ZEROCOPY") else "on", np.median(res), np.median(e2e), 1e3 * (np.median(e2e) - np.median(res)), [round(x, 3) for x in res], [round(x, 3) for x in e2e]))
bt.close()
