#!/usr/bin/env python
"""Generates mujoco_sim_b200/assets/pr2_like.xml — the C4 workload (BASELINE.json configs[3]).

A PR2-shaped mobile dual-arm robot with the structure of the reference's model/test/pr2/pr2.xml as it is spawned by
test/test_spawn_and_destroy_pr2.py: free base, 4 casters x (rotation + 2 wheels), torso lift, head pan / tilt, laser
tilt, two 7-joint arms, two grippers whose four finger joints are coupled to a driver joint by <equality><joint>
(mimic -> polycoef, src/mujoco_compile.cpp:219-314): nq = 50, nv = 49, 45 bodies, 8 joint equalities (the reference has
6), limited joints, body-pair excludes between neighbouring links.  The reference's meshes are replaced by primitives
(box / capsule / cylinder / sphere): convex-mesh collision is not built yet.  Geometry is
invented here; nothing is copied from the reference's assets."""
import os

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mujoco_sim_b200", "assets", "pr2_like.xml")

lines = []
excludes = []
equalities = []


def emit(depth, s):
    lines.append("  " * depth + s)


def body(depth, name, pos, joint=None, geoms=(), children=(), quat=None, mass=None):
    q = ' quat="%s"' % quat if quat else ""
    # the reference compensates gravity on robot bodies by default (mj_sim.cpp:301-310); here the upper body is
    # compensated while the base, casters and wheels carry their weight onto the floor
    gc = "0" if (name == "base_link" or "caster" in name or "wheel" in name) else "1"
    emit(depth, '<body name="%s" pos="%s"%s gravcomp="%s">' % (name, pos, q, gc))
    if joint:
        emit(depth + 1, joint)
    for g in geoms:
        emit(depth + 1, g)
    for c in children:
        c(depth + 1)
    emit(depth, "</body>")


def hinge(name, axis, rng=None, damping=0.5, armature=0.01):
    r = ' range="%s"' % rng if rng else ""
    return '<joint name="%s" type="hinge" axis="%s"%s damping="%g" armature="%g"/>' % (name, axis, r, damping, armature)


def slide(name, axis, rng, damping=20.0):
    return '<joint name="%s" type="slide" axis="%s" range="%s" damping="%g" armature="0.1"/>' % (name, axis, rng, damping)


def caster(prefix, x, y):
    def wheels(depth):
        for side, yy in (("l", 0.049), ("r", -0.049)):
            body(depth, "%s_%s_wheel_link" % (prefix, side), "0 %g 0" % yy,
                 hinge("%s_%s_wheel_joint" % (prefix, side), "0 1 0", damping=0.2),
                 ['<geom type="cylinder" size="0.074 0.017" quat="0.7071067811865476 0.7071067811865476 0 0" friction="1.5 0.005 0.0001"/>'])
            excludes.append(("%s_%s_wheel_link" % (prefix, side), "base_link"))

    def c(depth):
        body(depth, "%s_rotation_link" % prefix, "%g %g 0.0792" % (x, y), hinge("%s_rotation_joint" % prefix, "0 0 1", damping=0.3),
             ['<geom type="box" size="0.04 0.03 0.03" pos="0 0 0.04"/>'], [wheels])
    return c


def gripper(side):
    """driver slide + 4 finger hinges, all four coupled to the driver by joint equalities (mimic joints)."""
    drv = "%s_gripper_joint" % side

    def finger(depth, which, yy):
        jn = "%s_gripper_%s_finger_joint" % (side, which)
        tn = "%s_gripper_%s_finger_tip_joint" % (side, which)
        sgn = 1 if which == "l" else -1
        equalities.append((jn, drv, "0 %g 0 0 0" % (6.0,)))
        equalities.append((tn, drv, "0 %g 0 0 0" % (-6.0,)))

        def tip(d2):
            body(d2, "%s_gripper_%s_finger_tip_link" % (side, which), "0.09 %g 0" % (0.005 * sgn), hinge(tn, "0 0 %d" % -sgn, "-0.6 0.6", damping=0.05, armature=0.005),
                 ['<geom type="box" size="0.025 0.008 0.012" pos="0.025 0 0"/>'])
        body(depth, "%s_gripper_%s_finger_link" % (side, which), "0.077 %g 0" % yy, hinge(jn, "0 0 %d" % sgn, "0 0.55", damping=0.05, armature=0.005),
             ['<geom type="box" size="0.045 0.01 0.012" pos="0.045 0 0"/>'], [tip])
        excludes.append(("%s_gripper_%s_finger_tip_link" % (side, which), "%s_gripper_palm_link" % side))

    def g(depth):
        def parts(d2):
            body(d2, "%s_gripper_motor_slider_link" % side, "0.05 0 0", slide(drv, "1 0 0", "0 0.088", damping=5.0),
                 ['<geom type="sphere" size="0.01" contype="0" conaffinity="0" mass="0.05"/>'])
            body(d2, "%s_gripper_motor_screw_link" % side, "0.02 0 0", hinge("%s_gripper_motor_screw_joint" % side, "0 1 0", damping=0.01, armature=0.001),
                 ['<geom type="sphere" size="0.008" contype="0" conaffinity="0" mass="0.02"/>'])
            finger(d2, "l", 0.01)
            finger(d2, "r", -0.01)
        body(depth, "%s_gripper_palm_link" % side, "0 0 0", None, ['<geom type="box" size="0.04 0.04 0.025" pos="0.035 0 0"/>'], [parts])
        excludes.append(("%s_gripper_l_finger_link" % side, "%s_gripper_r_finger_link" % side))
        excludes.append(("%s_gripper_l_finger_tip_link" % side, "%s_gripper_r_finger_tip_link" % side))
        excludes.append(("%s_gripper_l_finger_link" % side, "%s_gripper_r_finger_tip_link" % side))
        excludes.append(("%s_gripper_r_finger_link" % side, "%s_gripper_l_finger_tip_link" % side))
    return g


def arm(side, y):
    p = side

    def wrist_roll(depth):
        def palm_holder(d2):
            gripper(p)(d2)
        body(depth, "%s_wrist_roll_link" % p, "0 0 0", hinge("%s_wrist_roll_joint" % p, "1 0 0", damping=0.1),
             ['<geom type="cylinder" size="0.03 0.02" pos="0.03 0 0" quat="0.7071067811865476 0 0.7071067811865476 0"/>'], [palm_holder])
        excludes.append(("%s_wrist_roll_link" % p, "%s_forearm_link" % p))
        excludes.append(("%s_gripper_palm_link" % p, "%s_wrist_flex_link" % p))

    def wrist_flex(depth):
        body(depth, "%s_wrist_flex_link" % p, "0.321 0 0", hinge("%s_wrist_flex_joint" % p, "0 1 0", "-2.18 0", damping=0.1),
             ['<geom type="sphere" size="0.04"/>'], [wrist_roll])
        excludes.append(("%s_wrist_flex_link" % p, "%s_forearm_roll_link" % p))

    def forearm(depth):
        body(depth, "%s_forearm_link" % p, "0 0 0", None, ['<geom type="capsule" fromto="0.05 0 0 0.27 0 0" size="0.045"/>'], [wrist_flex])

    def forearm_roll(depth):
        body(depth, "%s_forearm_roll_link" % p, "0 0 0", hinge("%s_forearm_roll_joint" % p, "1 0 0", damping=0.2),
             ['<geom type="sphere" size="0.03" contype="0" conaffinity="0" mass="0.3"/>'], [forearm])

    def elbow(depth):
        body(depth, "%s_elbow_flex_link" % p, "0.4 0 0", hinge("%s_elbow_flex_joint" % p, "0 1 0", "-2.32 0", damping=1.0),
             ['<geom type="sphere" size="0.055"/>'], [forearm_roll])
        excludes.append(("%s_elbow_flex_link" % p, "%s_upper_arm_link" % p))
        excludes.append(("%s_forearm_link" % p, "%s_upper_arm_link" % p))

    def upper_arm(depth):
        body(depth, "%s_upper_arm_link" % p, "0 0 0", None, ['<geom type="capsule" fromto="0.06 0 0 0.34 0 0" size="0.06"/>'], [elbow])

    def upper_arm_roll(depth):
        body(depth, "%s_upper_arm_roll_link" % p, "0 0 0", hinge("%s_upper_arm_roll_joint" % p, "1 0 0", "-3.9 0.8" if side == "r" else "-0.8 3.9", damping=0.5),
             ['<geom type="sphere" size="0.03" contype="0" conaffinity="0" mass="0.5"/>'], [upper_arm])

    def shoulder_lift(depth):
        body(depth, "%s_shoulder_lift_link" % p, "0.1 0 0", hinge("%s_shoulder_lift_joint" % p, "0 1 0", "-0.52 1.39", damping=3.0),
             ['<geom type="sphere" size="0.08"/>'], [upper_arm_roll])
        excludes.append(("%s_shoulder_lift_link" % p, "torso_lift_link"))
        excludes.append(("%s_upper_arm_link" % p, "%s_shoulder_pan_link" % p))
        excludes.append(("%s_upper_arm_link" % p, "torso_lift_link"))

    def a(depth):
        body(depth, "%s_shoulder_pan_link" % p, "0 %g 0" % y, hinge("%s_shoulder_pan_joint" % p, "0 0 1", "-2.28 0.71" if side == "r" else "-0.71 2.28", damping=3.0),
             ['<geom type="cylinder" size="0.09 0.12" pos="0 0 -0.1"/>'], [shoulder_lift])
        excludes.append(("%s_shoulder_pan_link" % p, "base_link"))
    return a


def head(depth):
    def tilt(d2):
        body(d2, "head_tilt_link", "0.068 0 0", hinge("head_tilt_joint", "0 1 0", "-0.47 1.39", damping=1.0),
             ['<geom type="box" size="0.08 0.14 0.06" pos="0.03 0 0.09"/>'])
    body(depth, "head_pan_link", "-0.017 0 0.38", hinge("head_pan_joint", "0 0 1", "-3.0 3.0", damping=1.0),
         ['<geom type="cylinder" size="0.06 0.03"/>'], [tilt])
    excludes.append(("head_tilt_link", "torso_lift_link"))


def laser(depth):
    body(depth, "laser_tilt_mount_link", "0.099 0 0.23", hinge("laser_tilt_mount_joint", "0 1 0", "-0.78 1.48", damping=0.5),
         ['<geom type="box" size="0.03 0.04 0.04"/>'])
    excludes.append(("laser_tilt_mount_link", "l_shoulder_pan_link"))
    excludes.append(("laser_tilt_mount_link", "r_shoulder_pan_link"))


def screw(depth):
    body(depth, "torso_lift_motor_screw_link", "-0.15 0 -0.2", hinge("torso_lift_motor_screw_joint", "0 0 1", damping=0.01, armature=0.001),
         ['<geom type="sphere" size="0.01" contype="0" conaffinity="0" mass="0.05"/>'])


def torso(depth):
    body(depth, "torso_lift_link", "-0.05 0 0.74", slide("torso_lift_joint", "0 0 1", "0 0.33", damping=200.0),
         ['<geom type="box" size="0.12 0.17 0.28" pos="-0.08 0 0.1"/>'], [screw, head, laser, arm("l", 0.188), arm("r", -0.188)])


def main():
    emit(0, "<!-- C4 workload (BASELINE.json configs[3]): PR2-shaped robot, generated by tools/make_pr2_like.py (see there). -->")
    emit(0, '<mujoco model="pr2_like">')
    emit(1, '<compiler angle="radian" autolimits="true"/>')
    emit(1, '<option timestep="0.005" gravity="0 0 -9.81"/>')
    emit(1, '<size nconmax="32" njmax="200"/>')
    emit(1, '<default><geom density="400"/></default>')
    emit(1, "<worldbody>")
    emit(2, '<geom name="floor" type="plane" size="0 0 0.05" condim="4" friction="2 0.05 0.01"/>')
    body(2, "base_link", "0 0 0.06", "<freejoint/>", ['<geom type="box" size="0.33 0.33 0.1" pos="0 0 0.19" density="600"/>'],
         [caster("fl_caster", 0.2246, 0.2246), caster("fr_caster", 0.2246, -0.2246), caster("bl_caster", -0.2246, 0.2246),
          caster("br_caster", -0.2246, -0.2246), torso])
    excludes.append(("torso_lift_link", "base_link"))
    emit(1, "</worldbody>")
    emit(1, "<contact>")
    seen = set()
    for a, b in excludes:
        key = tuple(sorted((a, b)))
        if key in seen:
            continue
        seen.add(key)
        emit(2, '<exclude body1="%s" body2="%s"/>' % (a, b))
    emit(1, "</contact>")
    emit(1, "<equality>")
    for j1, j2, poly in equalities:
        emit(2, '<joint joint1="%s" joint2="%s" polycoef="%s"/>' % (j1, j2, poly))
    emit(1, "</equality>")
    emit(0, "</mujoco>")
    with open(OUT, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
