"""World-size-2 CPU test (gloo) of the N > 1 path: contiguous environment shards, counter-based workload generation
and the max-over-ranks reduction bench.py uses.  The data path has no collective."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import mujoco_sim_b200 as b2
    from mujoco_sim_b200 import workloads as w
    from oracle import pyoracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = b2.Model(b2.asset("ur5_tabletop.xml"))
    nenv_total, steps = 16, 5
    per = nenv_total // world
    lo = rank * per
    qpos, qvel, frc = w.config_state("c3", m, np.arange(lo, lo + per))
    pool = [b2.Data(m)]
    qpos = np.ascontiguousarray(qpos); qvel = np.ascontiguousarray(qvel)
    orc.tick_batch(m, pool, steps, qpos, qvel, qfrc_applied=np.ascontiguousarray(frc))
    # the bench reduction: device time -> max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros(per, m.nq, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(torch.from_numpy(qpos), gathered, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "sharded.npy"), torch.cat(gathered).numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_the_single_batch(b2, orc, tmp_path):
    import torch.multiprocessing as mp
    from mujoco_sim_b200 import workloads as w
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sharded = np.load(tmp_path / "sharded.npy")
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0
    m = b2.Model(b2.asset("ur5_tabletop.xml"))
    qpos, qvel, frc = w.config_state("c3", m, np.arange(16))
    qpos = np.ascontiguousarray(qpos); qvel = np.ascontiguousarray(qvel)
    orc.tick_batch(m, [b2.Data(m)], 5, qpos, qvel, qfrc_applied=np.ascontiguousarray(frc))
    # environment e is bit-identical whether it was generated and stepped in a shard or in the whole batch
    assert np.array_equal(sharded, qpos)


def test_workload_generator_is_shard_invariant(b2):
    from mujoco_sim_b200 import workloads as w
    m = b2.Model(b2.asset("panda7.xml"))
    full = w.config_state("c2", m, np.arange(64))
    part = w.config_state("c2", m, np.arange(32, 64))
    for a, b in zip(full, part):
        assert np.array_equal(a[32:], b)
    lo, hi = m.jnt_range.reshape(-1, 2).T
    assert np.all(full[0] > lo) and np.all(full[0] < hi)
