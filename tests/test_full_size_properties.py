"""Size-independent properties at BASELINE.json's FULL sizes (the oracle only finishes small samples in seconds, so the full
batches are checked through invariants of the domain): inverse dynamics round trip, unilateral contact forces inside the
friction pyramid, unit quaternions, no cap overflow, determinism, shard invariance.  All through the C ABI, fp32 batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_full_size_inverse_dynamics_round_trip_and_determinism(b2):
    """C2 (4096 envs, contact-free): mj_inverse(mj_forward(tau)) == tau.  The tick computes qfrc_inverse from the PREVIOUS
    tick's qacc at the current state (MjHWInterface::read, mj_hw_interface.cpp:61), so after forward() at a fixed state
    a second forward() must return the applied force on every dof of every environment."""
    from mujoco_sim_b200 import workloads as w
    m = b2.Model(b2.asset(w.CONFIGS["c2"][0]))
    nenv = w.CONFIGS["c2"][1]
    assert nenv == 4096

    def run():
        bt = b2.Batch(m, nenv)
        qpos, qvel, frc, _ = w.load_config("c2", bt)
        bt.tick(0)                                     # mj_forward: leaves qacc
        bt.tick(b2.engine.TICK_INVERSE); bt.sync()     # forward again + mj_inverse on the acceleration of the first call
        finv = bt.get("qfrc_inverse")
        bt.step(100); bt.sync()
        out = bt.get("qpos", dtype=np.float32), bt.get("qvel", dtype=np.float32)
        bt.close()
        return np.asarray(frc), finv, out
    frc, finv, a = run()
    scale = np.maximum(1.0, np.abs(frc))
    assert np.max(np.abs(finv - frc) / scale) < 5e-4, np.max(np.abs(finv - frc) / scale)   # fp32: M (cond ~1e3) times M^-1
    _, _, b = run()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])      # bit-identical rerun
    assert np.isfinite(a[0]).all()


def test_c3_full_size_contact_invariants_and_shard_invariance(b2):
    """C3 (16384 envs, UR5 + props, PGS): after the settle phase every environment satisfies the constraint model's
    invariants, nothing overflowed, and a half-size shard reproduces its environments bit for bit."""
    from mujoco_sim_b200 import workloads as w
    m = b2.Model(b2.asset(w.CONFIGS["c3"][0]))
    nenv = w.CONFIGS["c3"][1]
    assert nenv == 16384

    def run(n, off):
        bt = b2.Batch(m, n)
        w.load_config("c3", bt, env_offset=off)
        bt.step(60); bt.sync()
        return bt
    bt = run(nenv, 0)
    qpos = bt.get("qpos"); status = bt.get("status")[:, 0]; ncon = bt.get("ncon")[:, 0]; nefc = bt.get("nefc")[:, 0]
    assert np.isfinite(qpos).all()
    assert not (status & 7).any(), np.unique(status & 7)                   # no contact / row cap hit, no state reset
    assert ncon.max() <= m.nconmax and nefc.max() <= m.njmax and ncon.mean() > 5
    # free bodies keep unit quaternions under the quaternion-aware integrator
    jt = np.array(m.jnt_type); qa = np.array(m.jnt_qposadr)
    for j in np.where(jt == 0)[0]:
        qn = np.linalg.norm(qpos[:, qa[j] + 3:qa[j] + 7], axis=1)
        assert np.abs(qn - 1).max() < 1e-4
        assert qpos[:, qa[j] + 2].min() > -0.02                            # nothing tunnels through the floor
    # unilateral forces: every pyramid / limit row force is >= 0 (PGS projects onto the cone's facets)
    force = bt.get("efc_force"); etype = bt.get("efc_type")
    rows = np.arange(force.shape[1])[None, :] < nefc[:, None]
    unilateral = rows & (etype >= 2)                                       # friction-loss (1) and equality (0) rows are bilateral
    assert force[unilateral].min() >= 0
    assert (force[rows] > 0).mean() > 0.05                                 # and contacts do carry load
    # penetration stays shallow once settled
    dist = bt.get("contact")[:, :m.nconmax]
    live = np.arange(m.nconmax)[None, :] < ncon[:, None]
    assert dist[live].min() > -0.03, dist[live].min()
    full_q = bt.get("qpos", dtype=np.float32)
    bt.close()
    half = run(nenv // 2, nenv // 2)
    hq = half.get("qpos", dtype=np.float32)
    diff = np.where((hq != full_q[nenv // 2:]).any(axis=1))[0]
    assert diff.size == 0, (diff.size, diff[:8], np.abs(hq - full_q[nenv // 2:]).max())
    half.close()
