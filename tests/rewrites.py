"""The XML rewrites the reference applies to a robot model before it is loaded, restated with
xml.etree for the tests — the reference does them with tinyxml2 on its own side of the boundary:

  init_tmp            src/mujoco_sim/mj_sim.cpp:296-335  gravcomp = 1 (disable_gravity) / 0 on every body under <worldbody>,
                                                         pose_init (x y z roll pitch yaw) on the robot's root body
                      src/mujoco_sim/mj_sim.cpp:338-420  odom joints <robot>_{lin,ang}_odom_{x,y,z}_joint appended to the
                                                         robot's root body (slide x / y / z, hinge x / y / z)
  init_references     src/mujoco_sim/mj_sim.cpp:847-960  for a received body: a joint-less mocap clone "<body>_ref" in a
                                                         new <worldbody>, <weld body1 body2 torquescale="0.9">, and
                                                         <exclude> of the clone against every body
"""
import copy
import math
import xml.etree.ElementTree as ET


def _each_body(e):
    for b in e.iter("body"):
        yield b


def rewrite(xml_text, robot, disable_gravity=True, pose_init=None, odom=("lin_odom_x_joint", "lin_odom_y_joint", "ang_odom_z_joint"),
            receive=()):
    root = ET.fromstring(xml_text)
    for wb in root.findall("worldbody"):
        for b in _each_body(wb):                                         # mj_sim.cpp:301-310
            b.set("gravcomp", "1" if disable_gravity else "0")
        for rb in wb.findall("body"):
            if rb.get("name") != robot:
                continue
            if pose_init is not None:                                     # mj_sim.cpp:312-335
                x, y, z, r, p, yw = pose_init
                rb.set("pos", "%f %f %f" % (x, y, z))
                cr, sr, cp, sp, cy, sy = math.cos(r / 2), math.sin(r / 2), math.cos(p / 2), math.sin(p / 2), math.cos(yw / 2), math.sin(yw / 2)
                rb.set("quat", "%f %f %f %f" % (cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy))
            want = set(odom)
            x_, y_, z_ = "lin_odom_x_joint" in want, "lin_odom_y_joint" in want, "lin_odom_z_joint" in want
            ax, ay, az = "ang_odom_x_joint" in want, "ang_odom_y_joint" in want, "ang_odom_z_joint" in want

            def add(name, typ, axis):
                ET.SubElement(rb, "joint", {"name": "%s_%s" % (robot, name), "type": typ, "axis": axis})
            # mj_sim.cpp:352-420 (x and y slides imply each other when a yaw hinge is requested)
            if x_ or (y_ and az):
                add("lin_odom_x_joint", "slide", "1 0 0")
            if y_ or (x_ and az):
                add("lin_odom_y_joint", "slide", "0 1 0")
            if z_ or (x_ and ay):
                add("lin_odom_z_joint", "slide", "0 0 1")
            if ax:
                add("ang_odom_x_joint", "hinge", "1 0 0")
            if ay:
                add("ang_odom_y_joint", "hinge", "0 1 0")
            if az:
                add("ang_odom_z_joint", "hinge", "0 0 1")
    if receive:                                                           # mj_sim.cpp:847-960
        names = [b.get("name") for wb in root.findall("worldbody") for b in _each_body(wb) if b.get("name")]
        eq = ET.SubElement(root, "equality")
        con = ET.SubElement(root, "contact")
        wnew = ET.SubElement(root, "worldbody")
        for body_name in receive:
            src = next(b for wb in root.findall("worldbody") for b in wb.findall("body") if b.get("name") == body_name)
            ref = copy.deepcopy(src)
            ref.set("name", body_name + "_ref")
            ref.set("mocap", "true")
            for j in list(ref.findall("joint")) + list(ref.findall("freejoint")):
                ref.remove(j)
            for ch in list(ref.findall("body")):                         # (the clone of a single free object has no children)
                ref.remove(ch)
            for g in ref.findall("geom"):
                g.set("rgba", ".5 .5 .5 1")
                g.attrib.pop("name", None)
            wnew.append(ref)
            ET.SubElement(eq, "weld", {"body1": body_name, "body2": body_name + "_ref", "torquescale": "0.9"})
            for n in names:
                ET.SubElement(con, "exclude", {"body1": n, "body2": body_name + "_ref"})
    return ET.tostring(root, encoding="unicode")
