#!/usr/bin/env python
"""Builds mujoco_sim_b200/assets/pr2_real.mjb — the C4 workload (BASELINE.json configs[3]) on the reference's own robot.

Input (read at build time from the read-only reference mount, never copied): model/test/pr2/pr2.xml with its 18 STL
meshes, and the floor of model/world/empty.xml:12 (plane, condim 4, friction 2 / 0.05 / 0.01) — the scene
test/test_spawn_and_destroy_pr2.py:25-42 creates by spawning the PR2 into the empty world.

What is written: a binary image of the COMPILED model (mj_saveModel): topology, inertials, joint ranges, the 6 mimic
equalities, the 105 excludes (as the filtered candidate-pair list) and, for every mesh, only the vertices of its convex
hull (scipy.spatial.ConvexHull) — what MuJoCo collides for a mesh geom.  No MJCF text and no STL data travel.
PD control as in the reference's PID config (model/ontology/box/box.yaml:5-13) is applied by the workload, not here."""
import os
import shutil
import sys
import tempfile
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from independent_model import read_stl  # noqa: E402  (a plain STL reader)

REF = "/root/reference/model"
OUT = os.path.join(ROOT, "mujoco_sim_b200", "assets", "pr2_real.mjb")


def main():
    import mujoco_sim_b200 as b2
    src = os.path.join(REF, "test", "pr2", "pr2.xml")
    tree = ET.parse(src)
    root = tree.getroot()
    tmp = tempfile.mkdtemp(prefix="pr2_real_")
    nv_in = nv_out = 0
    for me in root.iter("mesh"):
        f = os.path.join(os.path.dirname(src), me.get("file"))
        tris = read_stl(f)
        sc = np.array([float(x) for x in me.get("scale", "1 1 1").split()])
        pts = np.unique((tris * sc).reshape(-1, 3), axis=0)
        hull = ConvexHull(pts)
        hv = pts[hull.vertices]
        remap = {int(v): i for i, v in enumerate(hull.vertices)}
        c = hv.mean(0)
        name = me.get("name")
        with open(os.path.join(tmp, name + ".obj"), "w") as o:
            for v in hv:
                o.write("v %.9g %.9g %.9g\n" % tuple(v))
            for s in hull.simplices:
                a, b, d = (pts[i] for i in s)
                if np.dot(np.cross(b - a, d - a), a - c) < 0:      # outward-facing triangles (volume / centroid)
                    s = s[[0, 2, 1]]
                o.write("f %d %d %d\n" % tuple(remap[int(i)] + 1 for i in s))
        me.set("file", name + ".obj")
        me.attrib.pop("scale", None)
        nv_in += pts.shape[0]; nv_out += hv.shape[0]
    for c in root.findall("compiler"):
        c.set("meshdir", tmp)
    # the world the robot is spawned into (model/world/empty.xml): floor, gravity, timestep
    wb = root.find("worldbody")
    floor = ET.Element("geom", {"name": "floor", "size": "0 0 .05", "type": "plane", "condim": "4", "friction": "2 0.05 0.01"})
    wb.insert(0, floor)
    opt = ET.Element("option", {"timestep": "0.005", "gravity": "0 0 -9.81"})
    root.insert(0, opt)
    xml_path = os.path.join(tmp, "pr2_world.xml")
    tree.write(xml_path)
    m = b2.Model(xml_path)
    m.save(OUT)
    m2 = b2.Model(OUT)
    print("pr2_real.mjb: nq %d nv %d nbody %d ngeom %d nmesh %d neq %d npair %d; mesh vertices %d -> %d (hull); %d bytes"
          % (m2.nq, m2.nv, m2.nbody, m2.ngeom, m2.nmesh, m2.neq, m2.npair, nv_in, nv_out, os.path.getsize(OUT)))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
