"""URDF import behind mj_loadXML (reference src/mujoco_compile.cpp:317-405 hands URDF files to
mj_loadXML and writes MJCF back with mj_saveLastXML).  Host-side logic only: runs on CPU, the oracle is the checker."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import random_state


def forward_qacc(orc, m, d, qpos, qvel, frc):
    d.qpos[:] = qpos; d.qvel[:] = qvel; d.qfrc_applied[:] = frc
    d.qacc[:] = 0; d.qacc_warmstart[:] = 0
    orc.call("forward", m, d)
    return np.array(d.qacc), np.array(d.qM)


def test_panda_urdf_has_the_dynamics_of_the_mjcf_model(b2, orc):
    """configs[1] names "(URDF import)": the generated panda7.urdf (tools/make_panda_urdf.py) must compile to the same
    joint space, limits, mass matrix and forward dynamics as assets/panda7.xml."""
    mx = b2.Model(b2.asset("panda7.xml"))
    mu = b2.Model(b2.asset("panda7.urdf"))
    assert (mu.nq, mu.nv, mu.njnt) == (mx.nq, mx.nv, mx.njnt) == (7, 7, 7)
    assert mu.nbody == mx.nbody                          # the URDF root link is the world body (fusestatic)
    assert [mu.id2name(3, j) for j in range(7)] == [mx.id2name(3, j) for j in range(7)]
    assert mu.name2id(1, "panda_link0") == -1 and mu.name2id(1, "link7") == 7
    np.testing.assert_allclose(mu.jnt_range, mx.jnt_range, atol=1e-15)
    assert list(mu.jnt_limited) == list(mx.jnt_limited)
    np.testing.assert_allclose(mu.body_mass, mx.body_mass, rtol=1e-14)
    assert abs(mu.timestep - 0.005) < 1e-15
    dx, du = b2.Data(mx), b2.Data(mu)
    qpos, qvel, frc = random_state(mx, 16, 5)
    for e in range(16):
        ax, Mx = forward_qacc(orc, mx, dx, qpos[e], qvel[e], frc[e])
        au, Mu = forward_qacc(orc, mu, du, qpos[e], qvel[e], frc[e])
        np.testing.assert_allclose(Mu, Mx, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(au, ax, rtol=1e-8, atol=1e-9)


URDF = """<?xml version="1.0"?>
<robot name="cart">
  <mujoco><compiler balanceinertia="true" discardvisual="true" boundmass="0.000001" boundinertia="0.000001" meshdir="%s" strippath="false"/></mujoco>
  <link name="base">
    <inertial><origin xyz="0 0 0.1" rpy="0 0 0"/><mass value="5"/><inertia ixx="0.1" iyy="0.2" izz="0.3" ixy="0" ixz="0" iyz="0"/></inertial>
    <collision><origin xyz="0 0 0.1"/><geometry><box size="0.4 0.2 0.2"/></geometry></collision>
    <visual><geometry><box size="9 9 9"/></geometry></visual>
  </link>
  <link name="slider">
    <inertial><origin xyz="0 0 0" rpy="0.3 -0.2 0.5"/><mass value="1"/><inertia ixx="0.02" iyy="0.025" izz="0.03" ixy="0.001" ixz="-0.002" iyz="0.0005"/></inertial>
    <collision><origin xyz="0 0 0" rpy="1.5707963267948966 0 0"/><geometry><cylinder radius="0.05" length="0.3"/></geometry></collision>
  </link>
  <link name="pole">
    <inertial><origin xyz="0 0 0.25"/><mass value="0.5"/><inertia ixx="0.01" iyy="0.01" izz="0.001" ixy="0" ixz="0" iyz="0"/></inertial>
    <collision><origin xyz="0 0 0.5"/><geometry><sphere radius="0.04"/></geometry></collision>
  </link>
  <link name="tip">
    <inertial><origin xyz="0 0.01 0.05" rpy="0.2 0 0.1"/><mass value="0.1"/><inertia ixx="0.0002" iyy="0.0003" izz="0.0004" ixy="0.00001" ixz="0" iyz="0.00002"/></inertial>
    <collision><origin xyz="0 0 0.05" rpy="0 0.3 0"/><geometry><box size="0.04 0.02 0.06"/></geometry></collision>
  </link>
  <link name="wheel">
    <inertial><mass value="0.2"/><inertia ixx="0.001" iyy="0.001" izz="0.001" ixy="0" ixz="0" iyz="0"/></inertial>
    <collision><geometry><mesh filename="package://cart/meshes/octa.stl" scale="2 2 2"/></geometry></collision>
  </link>
  <joint name="slide" type="prismatic"><origin xyz="0 0 0.25" rpy="0 0 0.5"/><parent link="base"/><child link="slider"/><axis xyz="1 0 0"/>
    <limit lower="-0.5" upper="0.7" effort="10" velocity="1"/><dynamics damping="0.3" friction="0.05"/></joint>
  <joint name="swing" type="continuous"><origin xyz="0 0 0.05" rpy="0.1 0.2 0.3"/><parent link="slider"/><child link="pole"/><axis xyz="0 1 0"/></joint>
  <joint name="tip_fixed" type="fixed"><origin xyz="0 0 0.5"/><parent link="pole"/><child link="tip"/></joint>
  <joint name="spin" type="revolute"><origin xyz="0.2 0 0"/><parent link="base"/><child link="wheel"/><axis xyz="0 0 1"/><limit lower="-1" upper="1"/></joint>
</robot>
"""


def test_urdf_semantics_and_mjcf_round_trip(b2, orc, tmp_path):
    from test_gpu_parity import _octahedron_stl
    os.makedirs(tmp_path / "cart" / "meshes")
    _octahedron_stl(str(tmp_path / "cart" / "meshes" / "octa.stl"))
    path = tmp_path / "cart.urdf"
    path.write_text(URDF % str(tmp_path))
    m = b2.Model(str(path))
    # fusestatic (libmujoco's URDF default): "base" is the world, "tip" (fixed joint) is folded into "pole"
    assert (m.nq, m.nv, m.njnt, m.nbody) == (3, 3, 3, 4)
    assert [m.id2name(1, b) for b in range(1, 4)] == ["slider", "pole", "wheel"]
    assert list(m.jnt_type) == [2, 3, 3] and list(m.jnt_limited) == [1, 0, 1]       # slide, hinge, hinge
    np.testing.assert_allclose(m.jnt_range.reshape(-1, 2)[0], [-0.5, 0.7])
    np.testing.assert_allclose(m.dof_damping, [0.3, 0, 0]); np.testing.assert_allclose(m.dof_frictionloss, [0.05, 0, 0])
    # fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)
    r, p, y = 0.1, 0.2, 0.3
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    w, x, yy, z = m.body_quat.reshape(-1, 4)[2]
    Rq = np.array([[1 - 2 * (yy * yy + z * z), 2 * (x * yy - w * z), 2 * (x * z + w * yy)], [2 * (x * yy + w * z), 1 - 2 * (x * x + z * z), 2 * (yy * z - w * x)],
                   [2 * (x * z - w * yy), 2 * (yy * z + w * x), 1 - 2 * (x * x + yy * yy)]])
    np.testing.assert_allclose(Rq, Rz @ Ry @ Rx, atol=1e-14)
    # box extents and cylinder length are halved; visuals are dropped; mesh scaled; the fused link's box now belongs to "pole"
    gs = m.geom_size.reshape(-1, 3); gt = list(m.geom_type)
    assert gt == [6, 5, 2, 6, 7] and m.geom_bodyid[0] == 0 and list(m.geom_bodyid[2:4]) == [2, 2]
    np.testing.assert_allclose(gs[0], [0.2, 0.1, 0.1]); np.testing.assert_allclose(gs[1][:2], [0.05, 0.15]); assert abs(gs[2][0] - 0.04) < 1e-15
    np.testing.assert_allclose(gs[3], [0.02, 0.01, 0.03]); np.testing.assert_allclose(m.geom_pos.reshape(-1, 3)[3], [0, 0, 0.55], atol=1e-15)
    assert abs(np.abs(m.mesh_vert).max() - 0.3) < 1e-6                      # octahedron z extent 0.15, scaled by 2
    # the full inertia tensor given in a rotated inertial frame keeps its principal moments
    I = np.array([[0.02, 0.001, -0.002], [0.001, 0.025, 0.0005], [-0.002, 0.0005, 0.03]])
    np.testing.assert_allclose(np.sort(m.body_inertia.reshape(-1, 3)[1]), np.sort(np.linalg.eigvalsh(I)), rtol=1e-10)
    assert abs(m.body_mass[2] - 0.6) < 1e-15                                 # pole 0.5 + fused tip 0.1
    # <compiler fusestatic="false"/>: every non-root link stays a body; the two models have the same dynamics, which checks the
    # composed geom transforms and the combined inertial (mass, CoM, parallel-axis terms) of the fused body
    path2 = tmp_path / "cart_unfused.urdf"
    path2.write_text((URDF % str(tmp_path)).replace('<compiler ', '<compiler fusestatic="false" '))
    mu = b2.Model(str(path2))
    # (as in libmujoco, the root link is then a body of its own under the world — it would only be the world body if it were
    #  called "world")
    assert mu.nbody == 6 and [mu.id2name(1, b) for b in range(1, 6)] == ["base", "slider", "pole", "tip", "wheel"] and mu.body_dofnum[4] == 0
    assert mu.body_dofnum[1] == 0 and mu.body_parentid[1] == 0 and mu.geom_bodyid[0] == 1
    np.testing.assert_allclose(mu.geom_size, m.geom_size)
    df, du = b2.Data(m), b2.Data(mu)
    for q, v, f in [([0.1, 0.4, -0.3], [0.2, -0.5, 0.7], [1.0, -0.2, 0.05]), ([-0.3, 2.0, 0.5], [-1.0, 1.5, 0.2], [0.0, 0.3, -0.1])]:
        af, Mf = forward_qacc(orc, m, df, np.array(q), np.array(v), np.array(f))
        au, Mu = forward_qacc(orc, mu, du, np.array(q), np.array(v), np.array(f))
        np.testing.assert_allclose(Mf, Mu, rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(af, au, rtol=1e-9, atol=1e-10)
        orc.call("kinematics", m, df); orc.call("kinematics", mu, du)
        np.testing.assert_allclose(np.array(df.geom_xpos).reshape(-1, 3)[3], np.array(du.geom_xpos).reshape(-1, 3)[3], atol=1e-13)
    # mj_saveLastXML writes MJCF (src/mujoco_compile.cpp:470); loading that file reproduces the model
    out = tmp_path / "cart.xml"
    err = C.create_string_buffer(1000)
    assert b2.lib.mj_saveLastXML(str(out).encode(), m.ptr, err, 1000) == 1, err.value
    text = out.read_text()
    assert text.lstrip().startswith("<mujoco") and "<robot" not in text
    m2 = b2.Model(str(out))
    for name in ("body_pos", "body_quat", "body_mass", "body_inertia", "body_ipos", "jnt_axis", "jnt_range", "geom_size", "geom_pos", "dof_damping"):
        np.testing.assert_array_equal(getattr(m2, name), getattr(m, name))
    d, d2 = b2.Data(m), b2.Data(m2)
    q = np.array([0.1, 0.4, -0.3]); v = np.array([0.2, -0.5, 0.7]); f = np.array([1.0, -0.2, 0.05])
    a1, _ = forward_qacc(orc, m, d, q, v, f)
    a2, _ = forward_qacc(orc, m2, d2, q, v, f)
    np.testing.assert_array_equal(a1, a2)
    assert np.all(np.isfinite(a1))


def test_urdf_errors_follow_the_loader_contract(b2, tmp_path):
    """mj_loadXML returns NULL and fills the error text (include/mujoco_sim/mj_util.h:187-192)."""
    for body, msg in [('<link name="a"/><joint name="j" type="revolute"><parent link="a"/><child link="zz"/></joint>', "unknown link"),
                      ('<link name="a"/><link name="b"/><joint name="j" type="planar"><parent link="a"/><child link="b"/></joint>', "unsupported type"),
                      ("", "no <link>")]:
        with pytest.raises(b2.B2Error, match=msg):
            b2.Model(xml="<robot name='r'>%s</robot>" % body)


@pytest.mark.gpu
def test_urdf_imported_arm_takes_the_single_kernel_chain_path(b2):
    """The C2 workload is loaded through the URDF importer (workloads.CONFIGS): the imported model must be recognised as a
    limit-only serial chain (one kernel per tick) and step like the MJCF model it was generated from."""
    mx, mu = b2.Model(b2.asset("panda7.xml")), b2.Model(b2.asset("panda7.urdf"))
    nenv = 256
    qpos, qvel, frc = random_state(mx, nenv, 3)
    got = []
    for m in (mx, mu):
        bt = b2.Batch(m, nenv)
        assert "k_chain" in bt.path_name, bt.path_name
        bt.set("qpos", qpos); bt.set("qvel", qvel); bt.set("qfrc_applied", frc)
        l0 = bt.launch_count
        bt.step(50); bt.sync()
        assert bt.launch_count - l0 == 50
        got.append(bt.get("qpos"))
        bt.close()
    np.testing.assert_allclose(got[1], got[0], atol=2e-5)
