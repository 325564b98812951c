"""CPU-side tests: the C-ABI library loads and exports every symbol the headers declare, the host-side model compiler
and name / error contracts behave as the reference relies on them.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if n.startswith(("mj_", "mju_", "b2_"))))


@pytest.mark.parametrize("header", ["mujoco/mujoco.h", "b2_batch.h"])
def test_library_exports_every_declared_symbol(b2, header):
    names = declared_functions(header)
    assert len(names) > 15
    missing = [n for n in names if not hasattr(b2.lib, n)]
    assert not missing, missing
    # the global control callback is a data symbol (reference: mjcb_control = controller, src/mj_main.cpp:196)
    C.c_void_p.in_dll(b2.lib, "mjcb_control")


def test_no_cpu_fallback(b2):
    """Without a CUDA device b2_create must fail loudly; with one it must succeed."""
    m = b2.Model(b2.asset("panda7.xml"))
    if b2.lib.b2_device_count() == 0:
        with pytest.raises(b2.B2Error, match="no usable CUDA device"):
            b2.Batch(m, 4)
    else:
        b2.Batch(m, 4).close()


def test_loader_error_contract(b2, tmp_path):
    """mj_loadXML returns NULL and fills the error buffer (include/mujoco_sim/mj_util.h:187-192)."""
    err = C.create_string_buffer(500)
    assert not b2.lib.mj_loadXML(b"/nonexistent/model.xml", None, err, 500)
    assert b"cannot open" in err.value
    bad = tmp_path / "bad.xml"
    bad.write_text("<mujoco><worldbody><body><joint type='bogus'/></body></worldbody></mujoco>")
    assert not b2.lib.mj_loadXML(str(bad).encode(), None, err, 500)
    assert b"unknown joint type" in err.value
    bad.write_text("<mujoco><worldbody><body></worldbody></mujoco>")
    assert not b2.lib.mj_loadXML(str(bad).encode(), None, err, 500)
    assert b"XML parse error" in err.value


def test_name_lookup_contract(b2):
    """-1 for unknown names, NULL past the end / for unnamed objects, readable jnt_qposadr[-1]."""
    m = b2.Model(b2.asset("ur5_tabletop.xml"))
    assert m.name2id(b2.engine.OBJ_JOINT, "elbow_joint") == 2
    assert m.name2id(b2.engine.OBJ_JOINT, "robot_ang_odom_x_joint") == -1
    assert m.name2id(b2.engine.OBJ_BODY, "world") == 0
    assert m.id2name(b2.engine.OBJ_BODY, 1) == "shoulder_link"
    assert m.id2name(b2.engine.OBJ_BODY, m.nbody) is None      # loop terminator in src/mujoco_sim/mj_sim.cpp:473-477
    assert m.id2name(b2.engine.OBJ_GEOM, 2) is None            # unnamed geom
    qadr = m.array("jnt_qposadr")
    addr = qadr.ctypes.data - 4
    assert C.c_int.from_address(addr).value == 0               # set_odom_vels reads jnt_qposadr[-1] (mj_sim.cpp:1083-1091)


def test_compiler_matches_hand_computed_model_constants(b2):
    m = b2.Model(b2.asset("pendulum_world.xml"))
    assert (m.nq, m.nv, m.nbody, m.njnt, m.ngeom, m.nM) == (12, 9, 4, 3, 4, 18)
    # geom-inferred inertia at density 1000 (no <inertial> in the file)
    np.testing.assert_allclose(m.body_mass[1:], [4000 * np.pi * 0.1 ** 3 / 3, 8.0, 1000 * np.pi * 0.1 ** 2 * 0.2], rtol=1e-12)
    np.testing.assert_allclose(m.body_inertia[3:6], [0.4 * m.body_mass[1] * 0.01] * 3, rtol=1e-12)
    assert list(m.jnt_type) == [1, 1, 1] and list(m.jnt_qposadr) == [0, 4, 8] and list(m.jnt_dofadr) == [0, 3, 6]
    assert list(m.dof_parentid) == [-1, 0, 1, -1, 3, 4, -1, 6, 7]
    assert m.timestep == 0.005 and m.array("opt.gravity")[2] == -0.1
    assert m.int("opt.integrator") == 1                        # RK4 requested; step1/step2 run Euler
    # candidate pairs: three bob-floor pairs and three bob-bob pairs; the anchor-sharing bobs are not parent/child
    assert m.npair == 6
    m2 = b2.Model(b2.asset("ur5_tabletop.xml"))
    g1, g2, gb = m2.pair_geom1, m2.pair_geom2, m2.geom_bodyid
    for a, c in zip(g1, g2):
        ba, bc = gb[a], gb[c]
        assert ba != bc
        # parent-child filter: consecutive arm links never collide; links whose parent is the world still see the table
        if ba > 0 and bc > 0 and ba <= 6 and bc <= 6:
            assert abs(ba - bc) > 1
    types = m2.geom_type
    assert all(types[a] <= types[c] for a, c in zip(g1, g2))   # geom1 has the lower type; list sorted by type pair


def test_exclude_and_equality_parsing(b2):
    xml = """<mujoco><compiler angle="radian"/><worldbody>
      <body name="a"><joint name="ja" type="hinge"/><geom size="0.1"/>
        <body name="b" pos="0.5 0 0"><joint name="jb" type="hinge"/><geom size="0.1"/>
          <body name="c" pos="0.5 0 0"><joint name="jc" type="slide"/><geom size="0.1"/></body></body></body>
      <body name="d" pos="0 2 0"><freejoint/><geom type="box" size="0.1 0.1 0.1"/></body></worldbody>
      <contact><exclude body1="a" body2="c"/></contact>
      <equality><joint joint1="jc" joint2="ja" polycoef="0.1 2 0 0 0"/><weld body1="d" body2="a" torquescale="0.9"/></equality></mujoco>"""
    m = b2.Model(xml=xml)
    assert m.neq == 2 and list(m.eq_type) == [2, 1]
    np.testing.assert_allclose(m.eq_data[:5], [0.1, 2, 0, 0, 0])
    assert m.eq_data[11 + 10] == 0.9
    pairs = set(zip(m.pair_geom1.tolist(), m.pair_geom2.tolist()))
    assert (0, 2) not in pairs and (0, 1) not in pairs and (1, 2) not in pairs   # excluded / parent-child
    assert {(0, 3), (1, 3), (2, 3)} <= pairs


def test_save_last_xml_roundtrip(b2, tmp_path):
    """mj_saveLastXML (include/mujoco_sim/mj_util.h:207) re-emits a loadable model."""
    m = b2.Model(b2.asset("ur5_tabletop.xml"))
    out = tmp_path / "saved.xml"
    err = C.create_string_buffer(200)
    assert b2.lib.mj_saveLastXML(str(out).encode(), m.ptr, err, 200) == 1
    m2 = b2.Model(str(out))
    assert (m2.nq, m2.nv, m2.nbody, m2.npair) == (m.nq, m.nv, m.nbody, m.npair)
    np.testing.assert_array_equal(m2.body_mass, m.body_mass)


def test_mulM_on_host_mirror_matches_oracle(b2, orc):
    """mj_mulM of the shim works on the mirrored sparse qM (reference: tau = M ddq, mj_sim.cpp:1057)."""
    m = b2.Model(b2.asset("ur5_tabletop.xml"))
    d = b2.Data(m)
    orc.call("fwdPosition", m, d)
    v = np.linspace(-1, 2, m.nv).copy()
    y = np.zeros(m.nv); yr = np.zeros(m.nv)
    b2.lib.mj_mulM(m.ptr, d.ptr, y.ctypes.data, v.ctypes.data)
    orc.olib.omj_mulM(m.ptr, d.ptr, yr.ctypes.data, v.ctypes.data)
    np.testing.assert_allclose(y, yr, rtol=1e-14)
    np.testing.assert_allclose(y, orc.full_M(m, d) @ v, rtol=1e-12)


REF = "/root/reference/model"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout exists only in the build container")
def test_every_shipped_model_of_the_reference_compiles(b2, orc):
    """Row f1: mj_loadXML (include/mujoco_sim/mj_util.h:190) must accept the MJCF the reference ships — meshes, defaults,
    equalities, excludes, includes.  Sizes are the ones SURVEY.md Appendix B lists; the PR2 (18 STL meshes, 1229 candidate
    pairs) also steps under the oracle with its mesh geoms going through the general convex routine."""
    import glob
    files = sorted(glob.glob(REF + "/**/*.xml", recursive=True))
    assert len(files) >= 16
    sizes = {}
    for f in files:
        m = b2.Model(f)
        sizes[os.path.relpath(f, REF)] = (m.nq, m.nv, m.nbody, m.ngeom)
    assert sizes["test/pr2/pr2.xml"] == (50, 49, 45, 55)
    assert sizes["test/pendulum.xml"][:2] == (12, 9)
    assert sizes["test/ridgeback_panda/ridgeback_panda.xml"][:2] == (21, 20)
    m = b2.Model(REF + "/test/pr2/pr2.xml")
    assert m.nmesh == 18 and m.neq == 6 and m.nmeshvert > 1000
    d = b2.Data(m)
    d.qpos[:] = np.array(m.qpos0)
    for _ in range(20):
        orc.call("step", m, d)
    assert np.isfinite(np.array(d.qpos)).all() and np.isfinite(np.array(d.qacc)).all()
