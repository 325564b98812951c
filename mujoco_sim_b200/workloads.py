"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d): seeded, shard-invariant state generation.

Environment e always draws the same numbers for a given seed, whatever batch or shard it lives in (counter-based
hash of (seed, env, index)), so sharding across GPUs does not change any environment's trajectory."""
import numpy as np

from . import engine

CONFIGS = {
    # name: (model asset, default nenv, description)
    "c1": ("pendulum_world.xml", 1, "pendulum world, 1 env (plumbing)"),
    "c2": ("panda7.urdf", 4096, "Franka-Panda-like 7-DoF arm (URDF import), contact-free forward dynamics"),
    "c3": ("ur5_tabletop.xml", 16384, "UR5-like arm + tabletop objects with contacts, PGS"),
    "c4": ("pr2_real.mjb", 8192, "PR2 of the reference (model/test/pr2/pr2.xml: 49 dofs, 45 bodies, 37 mesh geoms as convex hulls, 6 mimic equalities, 105 excludes -> 1284 candidate pairs) on the floor of world/empty.xml + PD computed-torque control of the 14 arm joints"),
    "c5": ("multi_world.xml", 8192, "multi-robot world: 3 pendulum bobs + 20 object slots, run-time spawn / destroy as slot activation"),
}

PR2_ARM_JOINTS = ["%s_%s_joint" % (s, j) for s in ("l", "r") for j in
                  ("shoulder_pan", "shoulder_lift", "upper_arm_roll", "elbow_flex", "forearm_roll", "wrist_flex", "wrist_roll")]


def control_spec(cfg, model):
    """(hardware joint ids, per-dof controlled mask, kp, kd) of a config's control tick.  c2 / c3: every scalar joint is
    a ros_control joint commanded with a desired acceleration.  c4: the 14 arm joints under PD computed-torque control,
    ddq = 200 (q* - q) - 50 qdot (gains of the reference's PID config, model/ontology/box/box.yaml:5-13, i = 0)."""
    jt = np.array(model.jnt_type)
    if cfg == "c4":
        hw = np.array([model.name2id(engine.OBJ_JOINT, n) for n in PR2_ARM_JOINTS], np.int32)
        kp, kd = np.full(hw.size, 200.0, np.float32), np.full(hw.size, 50.0, np.float32)
    else:
        hw = np.where(jt >= 2)[0].astype(np.int32)
        kp = kd = None
    ctl = np.zeros(model.nv, np.uint8)
    ctl[np.array(model.jnt_dofadr)[hw]] = 1
    return hw, ctl, kp, kd


def commands(cfg, model, envs):
    """Effort-command buffer [nenv][nhw] of the config: desired accelerations 0.1 * (the config's random torques) for
    c2 / c3, position targets (mid-range +- 0.3 rad) for the PD-controlled arms of c4."""
    hw, _, kp, _ = control_spec(cfg, model)
    dadr = np.array(model.jnt_dofadr)[hw]
    if kp is None:
        _, _, frc = config_state(cfg, model, envs)
        return 0.1 * frc[:, dadr]
    envs = np.asarray(envs)
    rng = np.array(model.jnt_range).reshape(-1, 2)[hw]
    lim = np.array(model.jnt_limited)[hw] != 0
    mid = np.where(lim, 0.5 * (rng[:, 0] + rng[:, 1]), 0.0)
    u = uniform(0xB204 + 99, envs[:, None], np.arange(hw.size)[None, :])
    return mid[None, :] + (2 * u - 1) * 0.3


def uniform(seed, env, idx):
    """Counter-based uniform in [0, 1) from (seed, env, idx) (splitmix64 finaliser)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (np.asarray(env, np.uint64) * np.uint64(1000003) + np.asarray(idx, np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def random_state(model, envs, seed, vmax=1.0, fmax=10.0, free_xy=0.3, free_z=(0.02, 0.25), ranges=None):
    """qpos / qvel / qfrc_applied for the environment ids in `envs`.  Scalar joints: uniform inside their range
    (`ranges[joint_name]` overrides; unlimited hinges +-pi, slides +-0.3), free bodies displaced from qpos0 by
    U(-free_xy, free_xy) in x, y and U(free_z) in z with identity rotation and zero velocity."""
    envs = np.asarray(envs)
    nenv = envs.size
    nq, nv, njnt = model.nq, model.nv, model.njnt
    env = envs[:, None]
    uq = uniform(seed, env, np.arange(nq)[None, :])
    uv = uniform(seed + 1, env, np.arange(nv)[None, :])
    uf = uniform(seed + 2, env, np.arange(nv)[None, :])
    qpos = np.tile(np.array(model.qpos0), (nenv, 1))
    qvel = np.zeros((nenv, nv))
    frc = np.zeros((nenv, nv))
    jt, qa, da = model.jnt_type, model.jnt_qposadr, model.jnt_dofadr
    rng, lim = model.jnt_range.reshape(-1, 2), model.jnt_limited
    for j in range(njnt):
        a, d = int(qa[j]), int(da[j])
        if jt[j] == 0:
            qpos[:, a] += (2 * uq[:, a] - 1) * free_xy
            qpos[:, a + 1] += (2 * uq[:, a + 1] - 1) * free_xy
            qpos[:, a + 2] += free_z[0] + uq[:, a + 2] * (free_z[1] - free_z[0])
        elif jt[j] == 1:
            v = (2 * uq[:, a + 1:a + 4] - 1) * 0.5
            ang = np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
            qpos[:, a] = np.cos(ang[:, 0] / 2)
            qpos[:, a + 1:a + 4] = v / ang * np.sin(ang / 2)
            qvel[:, d:d + 3] = (2 * uv[:, d:d + 3] - 1) * vmax
            frc[:, d:d + 3] = (2 * uf[:, d:d + 3] - 1) * fmax
        else:
            name = model.id2name(engine.OBJ_JOINT, j)
            if ranges and name in ranges:
                lo, hi = ranges[name]
            elif lim[j]:
                lo, hi = rng[j]
            else:
                lo, hi = (-np.pi, np.pi) if jt[j] == 3 else (-0.3, 0.3)
            span = hi - lo
            qpos[:, a] = lo + 0.02 * span + uq[:, a] * 0.96 * span
            qvel[:, d] = (2 * uv[:, d] - 1) * vmax
            frc[:, d] = (2 * uf[:, d] - 1) * fmax
    return qpos, qvel, frc


# joint sub-ranges that keep the UR5-like arm above the table for most draws (C3)
UR5_RANGES = {"shoulder_pan_joint": (-1.2, 1.2), "shoulder_lift_joint": (-2.0, -0.4), "elbow_joint": (-0.8, 1.6),
              "wrist_1_joint": (-1.5, 1.5), "wrist_2_joint": (-1.5, 1.5), "wrist_3_joint": (-3.0, 3.0)}


def config_state(cfg, model, envs, seed=None):
    seed = (0xB200 + int(cfg[1])) if seed is None else seed
    if cfg == "c3":
        # props spread over the table (+-0.12 around their authored places, which are >= 0.2 m apart), dropped from
        # 5 mm of penetration to 8 cm of hover; arm torques are gentle so that it settles onto / around the table
        return random_state(model, envs, seed, vmax=0.5, fmax=5.0, free_xy=0.06, free_z=(-0.005, 0.08), ranges=UR5_RANGES)
    if cfg == "c4":
        # the robot is dropped onto its wheels from 0-25 cm at a random place / heading; arms near mid-range, everything
        # else at its authored pose (the gripper couplings are satisfied there); at rest
        envs = np.asarray(envs)
        nenv = envs.size
        q = np.tile(np.array(model.qpos0), (nenv, 1))
        u = uniform(seed, envs[:, None], np.arange(model.nq)[None, :])
        q[:, 0] += (2 * u[:, 0] - 1); q[:, 1] += (2 * u[:, 1] - 1); q[:, 2] += 0.25 * u[:, 2]
        yaw = (2 * u[:, 3] - 1) * np.pi
        q[:, 3] = np.cos(yaw / 2); q[:, 4:6] = 0; q[:, 6] = np.sin(yaw / 2)
        rngs = np.array(model.jnt_range).reshape(-1, 2)
        for n in PR2_ARM_JOINTS:
            j = model.name2id(engine.OBJ_JOINT, n)
            a = int(model.jnt_qposadr[j])
            mid = 0.5 * (rngs[j, 0] + rngs[j, 1]) if model.jnt_limited[j] else 0.0
            q[:, a] = mid + (2 * u[:, a] - 1) * 0.2
        return q, np.zeros((nenv, model.nv)), np.zeros((nenv, model.nv))
    if cfg == "c5":
        # bobs slightly off their authored pose (they swing), slots parked (c5_init spawns them)
        envs = np.asarray(envs)
        q = np.tile(np.array(model.qpos0), (envs.size, 1))
        v = np.zeros((envs.size, model.nv))
        v[:, :9] = (2 * uniform(seed, envs[:, None], np.arange(9)[None, :]) - 1) * 0.5
        return q, v, np.zeros((envs.size, model.nv))
    if cfg == "c1":
        q = np.tile(np.array(model.qpos0), (np.asarray(envs).size, 1))
        return q, np.zeros((q.shape[0], model.nv)), np.zeros((q.shape[0], model.nv))
    return random_state(model, envs, seed)


# ---- C5: object slots ----
NSLOT_C5, C5_INITIAL = 20, 8


def c5_slot_bodies(model):
    return np.array([model.name2id(engine.OBJ_BODY, "slot_%02d" % s) for s in range(NSLOT_C5)], np.int32)


def park_pose(slot):
    """Where an inactive slot rests (the engine's parking place: csrc/batch.cu park_slot)."""
    return np.array([3.0 * slot, 0.0, 1000.0 + 3.0 * slot, 1.0, 0.0, 0.0, 0.0])


def c5_spawn_pose(envs, rnd):
    """Spawn pose of round `rnd` for each environment: on a ring r ~ U(0.3, 1.2) around the pendulum anchor, z ~ U(0.3, 1.0),
    random heading (test/test_spawn_and_destroy.py:28-54 spawns on a ring at z = 5; lower here so that objects land and pile
    up within a benchmark run)."""
    envs = np.asarray(envs)
    u = uniform(0xB205 + 31 * rnd, envs[:, None], np.arange(5)[None, :])
    r, al, yaw = 0.3 + 0.9 * u[:, 0], (2 * u[:, 1] - 1) * np.pi, (2 * u[:, 3] - 1) * np.pi
    pose = np.zeros((envs.size, 7), np.float32)
    pose[:, 0], pose[:, 1], pose[:, 2] = r * np.sin(al), r * np.cos(al), 0.3 + 0.7 * u[:, 2]
    pose[:, 3], pose[:, 6] = np.cos(yaw / 2), np.sin(yaw / 2)
    return pose


def c5_init(batch, env_offset=0):
    """All slots destroyed, then C5_INITIAL of them spawned per environment (slot = (env + 3 k) mod 20)."""
    model = batch.model
    batch.set_slots(c5_slot_bodies(model))
    envs = np.arange(batch.nenv)
    for s in range(NSLOT_C5):
        batch.destroy_slots(envs, np.full(envs.size, s))
    for k in range(C5_INITIAL):
        batch.spawn(envs, (envs + env_offset + 3 * k) % NSLOT_C5, c5_spawn_pose(envs + env_offset, k))
    return C5_INITIAL


def c5_churn(batch, rnd, env_offset=0):
    """One destroy + one spawn per environment (the reference test does this every 0.3 s: test_spawn_and_destroy.py:86-94):
    the oldest live slot goes, the next free one in the cycle comes."""
    envs = np.arange(batch.nenv)
    g = envs + env_offset
    batch.destroy_slots(envs, (g + 3 * rnd) % NSLOT_C5)
    batch.spawn(envs, (g + 3 * (rnd + C5_INITIAL)) % NSLOT_C5, c5_spawn_pose(g, rnd + C5_INITIAL))


def load_states(batch, gen, max_depth=0.01, max_rounds=8):
    """Fill `batch` from gen(env_ids, round) -> (qpos, qvel, qfrc_applied).  Draws whose initial contacts penetrate
    deeper than max_depth (unphysical starts) are redrawn with round + 1, using the engine's own collision pass for
    the check.  Returns (qpos, qvel, frc, rounds_used)."""
    model = batch.model
    envs = np.arange(batch.nenv)
    qpos, qvel, frc = gen(envs, 0)
    batch.set("qpos", qpos); batch.set("qvel", qvel); batch.set("qfrc_applied", frc)
    if model.npair == 0:
        return qpos, qvel, frc, 0
    ncm = model.nconmax
    for rnd in range(1, max_rounds + 1):
        batch.forward()
        batch.sync()
        ncon = batch.get("ncon")[:, 0]
        dist = batch.get("contact", dtype=np.float32)[:, :ncm]
        mask = np.arange(ncm)[None, :] < ncon[:, None]
        bad = np.where(((dist < -max_depth) & mask).any(axis=1))[0]
        if bad.size == 0:
            return qpos, qvel, frc, rnd - 1
        q2, v2, f2 = gen(envs[bad], rnd)
        qpos[bad], qvel[bad], frc[bad] = q2, v2, f2
        batch.set("qpos", qpos); batch.set("qvel", qvel); batch.set("qfrc_applied", frc)
    return qpos, qvel, frc, max_rounds


def load_config(cfg, batch, env_offset=0, max_depth=0.01):
    """Synthetic state of config `cfg` for global environments [env_offset, env_offset + nenv)."""
    base = 0xB200 + int(cfg[1])
    return load_states(batch, lambda envs, rnd: config_state(cfg, batch.model, envs + env_offset, seed=base + 7919 * rnd), max_depth)
