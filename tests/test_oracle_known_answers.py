"""Known-answer tests that pin the fp64 oracle (the reference ships no golden vectors and libmujoco is not
available: SURVEY.md section 8c).  Every expectation here is analytic or a cross-check between two independent
formulations, so a wrong restatement of MuJoCo's published algorithm fails here before any GPU parity test runs."""
import numpy as np
import pytest

XML_BALL = """<mujoco><option timestep="0.005" gravity="0 0 -9.81"/>
<worldbody><geom name="floor" type="plane" size="0 0 .05" condim="4" friction="2 0.05 0.01"/>
<body name="ball" pos="0 0 1"><freejoint/><geom type="sphere" size="0.1"/></body></worldbody></mujoco>"""

XML_BOX = """<mujoco><option timestep="0.005" gravity="0 0 -9.81"/>
<worldbody><geom name="floor" type="plane" size="0 0 .05"/>
<body name="box" pos="0 0 0.2"><freejoint/><geom type="box" size="0.1 0.15 0.2"/></body></worldbody></mujoco>"""

XML_PEND = """<mujoco><compiler angle="radian"/><option timestep="0.001" gravity="0 0 -9.81"/>
<worldbody><body name="bob" pos="0 0 2"><joint name="h" type="hinge" axis="0 1 0" pos="0 0 0"/>
<geom type="sphere" pos="0 0 -1" size="0.05"/></body></worldbody></mujoco>"""

XML_LIMIT = """<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 -9.81"/>
<worldbody><body name="arm" pos="0 0 1"><joint name="h" type="hinge" axis="0 1 0" limited="true" range="-0.5 0.5"/>
<geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.03"/></body></worldbody></mujoco>"""

XML_MIMIC = """<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 -9.81"/>
<worldbody>
<body name="a" pos="0 0 1"><joint name="ja" type="hinge" axis="0 1 0" damping="0.2"/><geom type="capsule" fromto="0 0 0 0.3 0 0" size="0.02"/></body>
<body name="b" pos="0 0.5 1"><joint name="jb" type="hinge" axis="0 1 0" damping="0.2"/><geom type="capsule" fromto="0 0 0 0.3 0 0" size="0.02"/></body>
</worldbody>
<equality><joint joint1="jb" joint2="ja" polycoef="0.1 0.5 0 0 0"/></equality></mujoco>"""


def load(b2, xml):
    m = b2.Model(xml=xml)
    return m, b2.Data(m)


def test_free_fall_closed_form(b2, orc):
    """Semi-implicit Euler under constant gravity: z_n = z_0 - g h^2 n (n + 1) / 2 exactly."""
    m, d = load(b2, XML_BALL)
    h, g, n = 0.005, 9.81, 60
    for _ in range(n):
        orc.call("step", m, d)
    assert d.ncon == 0
    assert abs(d.qpos[2] - (1.0 - g * h * h * n * (n + 1) / 2)) < 1e-12
    assert abs(d.qvel[2] + g * h * n) < 1e-12
    assert abs(d.time - n * h) < 1e-12


def test_mass_matrix_equals_rne_columns(b2, orc):
    """CRBA mass matrix == column-wise RNE differences: M e_i = RNE(q, 0, e_i) - RNE(q, 0, 0)."""
    from helpers import random_state
    for name in ["panda7.xml", "ur5_tabletop.xml", "pendulum_world.xml"]:
        m = b2.Model(b2.asset(name))
        d = b2.Data(m)
        qpos, _, _ = random_state(m, 1, 7)
        d.qpos[:] = qpos[0]
        d.qvel[:] = 0
        orc.call("fwdPosition", m, d)
        orc.call("fwdVelocity", m, d)
        M = orc.full_M(m, d)
        base = orc.rne(m, d, 0)
        for i in range(m.nv):
            d.qacc[:] = 0
            d.qacc[i] = 1
            col = orc.rne(m, d, 1) - base
            np.testing.assert_allclose(M[:, i] - np.array(m.dof_armature) * (np.arange(m.nv) == i), col, atol=1e-10, err_msg=name)
        assert np.all(np.linalg.eigvalsh(M) > 0)


def test_sparse_solve_matches_dense(b2, orc):
    from helpers import random_state
    m = b2.Model(b2.asset("ur5_tabletop.xml"))
    d = b2.Data(m)
    qpos, qvel, frc = random_state(m, 1, 3)
    d.qpos[:] = qpos[0]
    orc.call("fwdPosition", m, d)
    M = orc.full_M(m, d)
    x = np.linspace(-1, 1, m.nv).copy()
    ref = np.linalg.solve(M, x)
    orc.olib.omj_solveM(m.ptr, d.ptr, x.ctypes.data, 1)
    np.testing.assert_allclose(x, ref, rtol=1e-9, atol=1e-10)
    y = np.zeros(m.nv)
    v = np.linspace(0.3, -2, m.nv).copy()
    orc.olib.omj_mulM(m.ptr, d.ptr, y.ctypes.data, v.ctypes.data)
    np.testing.assert_allclose(y, M @ v, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["panda7.xml", "ur5_tabletop.xml", "pendulum_world.xml"])
def test_inverse_of_forward_is_applied_force(b2, orc, name):
    """mj_inverse(mj_forward(q, v, tau)) == tau, with or without active constraints."""
    from helpers import random_state
    m = b2.Model(b2.asset(name))
    d = b2.Data(m)
    qpos, qvel, frc = random_state(m, 4, 11, free_z=(-0.03, 0.1))
    for e in range(4):
        d.qpos[:] = qpos[e]; d.qvel[:] = qvel[e]; d.qfrc_applied[:] = frc[e]
        d.qacc_warmstart[:] = 0
        m.set_opt("iterations", 2000)
        m.set_opt("tolerance", 0)
        orc.call("forward", m, d)
        orc.call("inverse", m, d)
        # the PGS force and the primal force law agree only at convergence: 2000 sweeps get there
        np.testing.assert_allclose(d.qfrc_inverse, frc[e], atol=2e-5 * max(1, np.abs(frc[e]).max()), err_msg="%s env %d nefc %d" % (name, e, d.nefc))


def test_pendulum_period_and_energy(b2, orc):
    """Point-ish bob on a 1 m hinge: small-angle period 2 pi sqrt(I / (m g L)); energy drift of the symplectic Euler
    scheme stays bounded."""
    m, d = load(b2, XML_PEND)
    m.set_opt("timestep", 0.001)
    mass = m.body_mass[1]
    I = orc.full_M.__call__  # noqa
    d.qpos[0] = 0.05
    orc.call("fwdPosition", m, d)
    Izz = orc.full_M(m, d)[0, 0]
    T_expected = 2 * np.pi * np.sqrt(Izz / (mass * 9.81 * 1.0))
    zero_cross, prev, t = [], d.qpos[0], 0.0
    e0 = None
    emin, emax = 1e9, -1e9
    for s in range(6000):
        orc.call("step1", m, d)
        orc.call("energy", m, d)
        e = d.energy[0] + d.energy[1]
        emin, emax = min(emin, e), max(emax, e)
        orc.call("step2", m, d)
        if prev > 0 >= d.qpos[0]:
            zero_cross.append(d.time)
        prev = d.qpos[0]
    period = np.mean(np.diff(zero_cross))
    assert abs(period - T_expected) / T_expected < 2e-3  # small-angle + O(h) error
    assert (emax - emin) / abs(mass * 9.81 * 1.0 * (1 - np.cos(0.05))) < 0.05


def test_gravcomp_gives_zero_acceleration_at_rest(b2, orc):
    """gravcomp=1 on every body (the reference's default mode, src/mujoco_sim/mj_sim.cpp:301-310) cancels gravity."""
    xml = open(b2.asset("panda7.xml")).read().replace('<body name="link', '<body gravcomp="1" name="link')
    m = b2.Model(xml=xml)
    d = b2.Data(m)
    from helpers import random_state
    qpos, _, _ = random_state(m, 1, 5)
    d.qpos[:] = qpos[0]
    orc.call("forward", m, d)
    assert np.abs(d.qacc).max() < 1e-9
    np.testing.assert_allclose(d.qfrc_passive, d.qfrc_bias, atol=1e-10)


def test_sphere_rests_on_plane_with_weight_as_normal_force(b2, orc):
    m, d = load(b2, XML_BALL)
    d.qpos[2] = 0.1
    for _ in range(400):
        orc.call("step", m, d)
    assert d.ncon == 1
    c = d.model  # noqa
    assert d.nefc == 6  # condim 4 -> 2 * (4 - 1) pyramidal rows
    mg = m.body_mass[1] * 9.81
    assert abs(np.sum(d.efc_force[:6]) - mg) / mg < 1e-3
    assert abs(d.qvel[2]) < 1e-4 and -0.01 < d.qpos[2] - 0.1 < 0
    # geom ids: plane (geom 0) first, sphere second
    from ctypes import Structure  # noqa


def test_box_rests_on_four_corner_contacts(b2, orc):
    m, d = load(b2, XML_BOX)
    for _ in range(400):
        orc.call("step", m, d)
    assert d.ncon == 4
    assert d.nefc == 16
    mg = m.body_mass[1] * 9.81
    assert abs(np.sum(d.efc_force[:16]) - mg) / mg < 1e-3
    assert abs(d.qpos[2] - 0.2) < 5e-3
    assert np.abs(d.qvel).max() < 1e-3
    np.testing.assert_allclose(d.qpos[3:7], [1, 0, 0, 0], atol=1e-6)


def test_joint_limit_is_one_sided(b2, orc):
    m, d = load(b2, XML_LIMIT)
    d.qpos[0] = 0.0
    orc.call("forward", m, d)
    assert d.nefc == 0
    d.qpos[0] = 0.55  # beyond the upper limit: one row pushing back (negative acceleration contribution)
    d.qvel[0] = 0
    orc.call("forward", m, d)
    assert d.nefc == 1 and d.efc_J[0] == -1 and d.efc_pos[0] < 0 and d.efc_force[0] > 0
    assert d.qacc[0] < d.qacc_smooth[0]
    # falls under gravity onto the upper stop and stays there
    d.qpos[0] = 0.0
    for _ in range(3000):
        orc.call("step", m, d)
    assert 0.5 < d.qpos[0] < 0.52 and abs(d.qvel[0]) < 1e-3


def test_mimic_equality_converges(b2, orc):
    """URDF mimic joints become <equality><joint polycoef> (src/mujoco_compile.cpp:219-248): q_b = 0.1 + 0.5 q_a."""
    m, d = load(b2, XML_MIMIC)
    for _ in range(3000):
        orc.call("step", m, d)
    assert d.nefc == 1
    assert abs(d.qpos[1] - (0.1 + 0.5 * d.qpos[0])) < 2e-3


def test_controller_and_odom_follow_the_reference_semantics(b2, orc):
    """MjSim::controller (mj_sim.cpp:1055-1077) and set_odom_vels (:1079-1153)."""
    from helpers import random_state
    m = b2.Model(b2.asset("panda7.xml"))
    d = b2.Data(m)
    qpos, qvel, _ = random_state(m, 1, 9)
    d.qpos[:] = qpos[0]; d.qvel[:] = qvel[0]
    orc.call("step1", m, d)
    M = orc.full_M(m, d)
    ddq = np.linspace(-1, 1, 7).copy(); dq = np.zeros(7); dq[2] = 0.3
    ctl = np.array([1, 1, 0, 0, 1, 0, 0], np.uint8)
    expect = M @ ddq + np.array(d.qfrc_bias) * ctl
    orc.controller(m, d, ddq, dq, ctl)
    np.testing.assert_allclose(d.qfrc_applied, expect, atol=1e-12)
    assert d.qvel[2] == 0.3 and d.qvel[1] == qvel[0][1]
    assert not ddq.any() and not dq.any()
    # odom: rotation of the commanded twist by ZYX(odom angles); rows of Rz Ry Rx
    x, y, z = 0.1, -0.2, 0.7
    d.qpos[0], d.qpos[1], d.qpos[2] = x, y, z
    v = np.array([0.4, -0.3, 0.2, 0.01, 0.02, 0.03])
    orc.set_odom_vels(m, d, [3, 4, 5], [6, -1, -1], [0, 1, 2], v)
    cx, sx, cy, sy, cz, sz = np.cos(x), np.sin(x), np.cos(y), np.sin(y), np.cos(z), np.sin(z)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    np.testing.assert_allclose(d.qvel[3:6], Rz @ Ry @ Rx @ v[:3], atol=1e-12)
    assert d.qvel[6] == 0.01


def test_oracle_is_deterministic(b2, orc):
    from helpers import random_state, oracle_rollout
    m = b2.Model(b2.asset("ur5_tabletop.xml"))
    qpos, qvel, frc = random_state(m, 1, 21)
    outs = []
    for _ in range(2):
        d = b2.Data(m)
        outs.append(oracle_rollout(orc, m, d, qpos[0], qvel[0], frc[0], 50))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_rne_post_constraint_known_answers(b2, orc):
    """mj_rnePostConstraint (what force / torque sensors read), pinned by statics:
    (1) a two-link arm hanging at rest under gravity: the interaction force at the first joint carries the weight of both
        links, at the second joint the weight of the second link (c-frame translational part, +z);
    (2) a box at rest on the floor: its weight is carried by the contact forces, so the interaction force across its free
        joint vanishes; with the contacts disabled (free fall) it vanishes too (cacc cancels the pseudo-gravity)."""
    xml = """<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 -9.81"/><worldbody>
    <body name="l1" pos="0 0 1"><joint type="hinge" axis="0 1 0" damping="5"/><geom type="capsule" size="0.03" fromto="0 0 0 0 0 -0.4"/>
      <body name="l2" pos="0 0 -0.4"><joint type="hinge" axis="0 1 0" damping="5"/><geom type="capsule" size="0.02" fromto="0 0 0 0 0 -0.3"/></body>
    </body></worldbody></mujoco>"""
    m = b2.Model(xml=xml); d = b2.Data(m)
    orc.call("forward", m, d)
    orc.call("rnePostConstraint", m, d)
    mass = np.array(m.body_mass)
    cf = np.array(d.cfrc_int).reshape(-1, 6)
    assert np.allclose(cf[1, 3:], [0, 0, (mass[1] + mass[2]) * 9.81], atol=1e-9)
    assert np.allclose(cf[2, 3:], [0, 0, mass[2] * 9.81], atol=1e-9)
    assert np.allclose(np.array(d.cacc).reshape(-1, 6)[1:, 3:], [[0, 0, 9.81]] * 2, atol=1e-9)
    xml2 = """<mujoco><option timestep="0.002"/><worldbody><geom type="plane" size="0 0 1"/>
    <body name="box" pos="0 0 0.0995"><freejoint/><geom type="box" size="0.1 0.1 0.1"/></body></worldbody></mujoco>"""
    m2 = b2.Model(xml=xml2); d2 = b2.Data(m2)
    for _ in range(600):
        orc.call("step", m2, d2)
    orc.call("forward", m2, d2)
    orc.call("rnePostConstraint", m2, d2)
    w = float(m2.body_mass[1]) * 9.81
    assert d2.ncon == 4 and abs(np.array(d2.qacc)).max() < 1e-3
    assert np.abs(np.array(d2.cfrc_int).reshape(-1, 6)[1]).max() < 2e-3 * w
    m2.set_opt("disableflags", 16)   # contacts off: free fall
    orc.call("forward", m2, d2)
    orc.call("rnePostConstraint", m2, d2)
    assert np.abs(np.array(d2.cfrc_int).reshape(-1, 6)[1]).max() < 1e-9 * w
    assert np.allclose(np.array(d2.cacc).reshape(-1, 6)[1, 3:], 0, atol=1e-9)
