"""Shared helpers for the tests: model snippets, seeded state generation, oracle drivers."""
import numpy as np

MASK64 = (1 << 64) - 1


def splitmix(seed, env, idx):
    """Counter-based uniform in [0,1): value depends only on (seed, env, idx) -> shard-invariant."""
    z = (np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (np.asarray(env, np.uint64) * np.uint64(1000003) + np.asarray(idx, np.uint64) + np.uint64(1)))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def random_state(model, nenv, seed, env_offset=0, vmax=1.0, fmax=10.0, free_xy=0.3, free_z=(0.02, 0.25)):
    """qpos / qvel / qfrc_applied for environments [env_offset, env_offset+nenv): scalar joints uniform in their range
    (or +-pi / +-0.3 when unlimited), free bodies displaced from qpos0, qvel ~ U(-vmax, vmax) on scalar dofs."""
    nq, nv, njnt = model.nq, model.nv, model.njnt
    env = np.arange(env_offset, env_offset + nenv)[:, None]
    with np.errstate(over="ignore"):
        uq = splitmix(seed, env, np.arange(nq)[None, :])
        uv = splitmix(seed + 1, env, np.arange(nv)[None, :])
        uf = splitmix(seed + 2, env, np.arange(nv)[None, :])
    qpos = np.tile(np.array(model.qpos0), (nenv, 1))
    qvel = np.zeros((nenv, nv))
    frc = np.zeros((nenv, nv))
    jt, qa, da = model.jnt_type, model.jnt_qposadr, model.jnt_dofadr
    rng, lim = model.jnt_range.reshape(-1, 2), model.jnt_limited
    for j in range(njnt):
        a, d = qa[j], da[j]
        if jt[j] == 0:  # free
            qpos[:, a] += (2 * uq[:, a] - 1) * free_xy
            qpos[:, a + 1] += (2 * uq[:, a + 1] - 1) * free_xy
            qpos[:, a + 2] += free_z[0] + uq[:, a + 2] * (free_z[1] - free_z[0])
        elif jt[j] == 1:  # ball: small random rotation
            v = (2 * uq[:, a + 1:a + 4] - 1) * 0.5
            ang = np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
            qpos[:, a] = np.cos(ang[:, 0] / 2)
            qpos[:, a + 1:a + 4] = v / ang * np.sin(ang / 2)
            qvel[:, d:d + 3] = (2 * uv[:, d:d + 3] - 1) * vmax
            frc[:, d:d + 3] = (2 * uf[:, d:d + 3] - 1) * fmax
        else:
            lo, hi = (rng[j] if lim[j] else ((-np.pi, np.pi) if jt[j] == 3 else (-0.3, 0.3)))
            # stay 2 % inside the range so that the initial state has no active limit
            span = hi - lo
            qpos[:, a] = lo + 0.02 * span + uq[:, a] * 0.96 * span
            qvel[:, d] = (2 * uv[:, d] - 1) * vmax
            frc[:, d] = (2 * uf[:, d] - 1) * fmax
    return qpos, qvel, frc


def oracle_rollout(orc, model, data, qpos, qvel, frc, nsteps, record=None):
    """Step one environment with the oracle; returns final (qpos, qvel)."""
    data.qpos[:] = qpos
    data.qvel[:] = qvel
    data.qacc[:] = 0
    data.qacc_warmstart[:] = 0
    data.qfrc_applied[:] = frc
    data.array("time")[0] = 0
    for s in range(nsteps):
        orc.call("step", model, data)
        if record is not None:
            record(s, data)
    return np.array(data.qpos), np.array(data.qvel)
