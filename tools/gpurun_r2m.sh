ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_chain -s 30 -c 5 --csv --log-file gpurun_out/c2_kernel.csv python tools/tick_some.py c2 10 30 > /dev/null 2>&1
grep -v "^==" gpurun_out/c2_kernel.csv | tail -5 | cut -c1-300
python - <<'PY'
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import torch
import mujoco_sim_b200 as b2
from mujoco_sim_b200 import workloads as w
m = b2.Model(b2.asset(w.CONFIGS["c2"][0]))
nenv = 4096
bt = b2.Batch(m, nenv)
w.load_config("c2", bt)
hw, ctl, kp, kd = w.control_spec("c2", m)
bt.set_controlled(ctl); bt.set_hw_joints(hw)
cmd = w.commands("c2", m, np.arange(nenv))
bt.write_commands(np.zeros((hw.size, nenv), np.float32), np.ascontiguousarray(cmd.T.astype(np.float32)))
for _ in range(20): bt.tick_resident()
bt.sync()
stream = torch.cuda.ExternalStream(bt.stream)
K = 50
st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
for flush in (False, True):
    for k in range(K):
        if flush: bt.l2_flush(256 << 20)
        st[k].record(stream); bt.tick_resident(); en[k].record(stream)
    bt.sync(); torch.cuda.synchronize()
    ts = [s.elapsed_time(e) * 1e3 for s, e in zip(st, en)]
    print("flush", flush, "per-step event pairs: median %.1f us min %.1f us" % (np.median(ts), min(ts)))
t0 = time.perf_counter()
for k in range(2000): bt.tick_resident()
bt.sync()
print("wall per tick back-to-back: %.1f us" % ((time.perf_counter() - t0) / 2000 * 1e6))
PY
