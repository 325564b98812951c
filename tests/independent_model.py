"""An INDEPENDENT recomputation of the model constants the engine and the oracle both take from one compiler
(mujoco_sim_b200/csrc/mjcf_compile.cpp): written from MuJoCo's XML reference and Computation chapter alone, in numpy, and
sharing no code with the C++ compiler.  It reads the MJCF subset the shipped models use (xml.etree), places the bodies at
qpos0, and computes

  * geom-inferred body mass, centre of mass and inertia tensor (density 1000 unless given), explicit <inertial> otherwise;
  * mesh volume, volume centroid (signed tetrahedra) and convex-hull vertex set (scipy.spatial.ConvexHull);
  * the joint-space inertia matrix at qpos0 as  M = sum_b  m_b Jp_b^T Jp_b + Jr_b^T I_b Jr_b + diag(armature)
    (no CRBA), hence dof_invweight0 = diag(M^-1) (averaged over the 3 dofs of ball / free rotations and translations) and
    body_invweight0 = (tr(Jp M^-1 Jp^T) / 3, tr(Jr M^-1 Jr^T) / 3) at the body's centre of mass;
  * the candidate geom-pair list after MuJoCo's filters (same body, both bodies static, parent-child unless the parent
    is static, contype / conaffinity, <exclude>).

TEST INFRASTRUCTURE (a pin for the compiler that stands outside it): used by
tests/test_compiler_independent.py only.
"""
import math
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np


def quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat_mat(q):
    q = np.asarray(q, float) / np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def axis_angle_quat(axis, ang):
    axis = np.asarray(axis, float)
    n = np.linalg.norm(axis)
    if n < 1e-14:
        return np.array([1.0, 0, 0, 0])
    axis = axis / n
    return np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * axis])


def zaxis_quat(v):
    """Rotation taking +z to v (MuJoCo's quatZ2Vec), used by fromto / zaxis."""
    v = np.asarray(v, float)
    v = v / np.linalg.norm(v)
    z = np.array([0.0, 0, 1])
    ax = np.cross(z, v)
    s = np.linalg.norm(ax)
    if s < 1e-10:
        return np.array([1.0, 0, 0, 0]) if v[2] > 0 else np.array([0.0, 1, 0, 0])
    return axis_angle_quat(ax / s, math.atan2(s, float(np.dot(z, v))))


def read_stl(path):
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack_from("<I", data, 80)[0]
    if 84 + 50 * n == len(data):
        tri = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n, offset=84)
        return np.asarray(tri["v"], float)
    pts = []
    for line in data.decode("ascii", "ignore").splitlines():
        t = line.split()
        if len(t) == 4 and t[0] == "vertex":
            pts.append([float(x) for x in t[1:]])
    return np.array(pts, float).reshape(-1, 3, 3)


def read_obj(path):
    vs, fs = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            vs.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            idx = [int(x.split("/")[0]) - 1 for x in t[1:]]
            for k in range(1, len(idx) - 1):
                fs.append([idx[0], idx[k], idx[k + 1]])
    vs = np.array(vs, float)
    return vs[np.array(fs, int)]


def mesh_props(tris):
    """Volume, volume centroid and inertia tensor about the centroid (unit density) of a closed triangle surface."""
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    v6 = np.einsum("ij,ij->i", a, np.cross(b, c))          # 6 x signed volume of the tetrahedron (0, a, b, c)
    vol = v6.sum() / 6.0
    cen = ((a + b + c) / 4.0 * (v6 / 6.0)[:, None]).sum(0) / vol
    # second moments of a tetrahedron with one vertex at the origin
    C = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            s = (a[:, i] * a[:, j] + b[:, i] * b[:, j] + c[:, i] * c[:, j]) * 2 \
                + a[:, i] * b[:, j] + a[:, j] * b[:, i] + a[:, i] * c[:, j] + a[:, j] * c[:, i] + b[:, i] * c[:, j] + b[:, j] * c[:, i]
            C[i, j] = (s * v6).sum() / 120.0
    if vol < 0:
        vol, C = -vol, -C
    C -= vol * np.outer(cen, cen)                            # about the centroid
    inertia = np.trace(C) * np.eye(3) - C
    return vol, cen, inertia


class Body:
    pass


class Geom:
    pass


class IndependentModel:
    GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4, "cylinder": 5, "box": 6, "mesh": 7}

    def __init__(self, path):
        self.dir = os.path.dirname(os.path.abspath(path))
        root = ET.parse(path).getroot()
        self.root = root
        comp = {}
        for c in root.findall("compiler"):      # several <compiler> elements accumulate
            comp.update(c.attrib)
        self.degree = comp.get("angle", "degree") == "degree"
        self.eulerseq = comp.get("eulerseq", "xyz")
        self.meshdir = comp.get("meshdir", "")
        self.boundmass = float(comp.get("boundmass", 0))
        self.boundinertia = float(comp.get("boundinertia", 0))
        self.inertiafromgeom = comp.get("inertiafromgeom", "auto")
        self.classes = {"main": ({}, None)}
        for d in root.findall("default"):
            self._defaults(d, None, True)
        self.meshes = {}
        for asset in root.findall("asset"):
            for me in asset.findall("mesh"):
                f = os.path.join(self.dir, self.meshdir, me.get("file"))
                tris = read_obj(f) if f.lower().endswith(".obj") else read_stl(f)
                sc = np.array([float(x) for x in me.get("scale", "1 1 1").split()])
                name = me.get("name") or os.path.splitext(os.path.basename(me.get("file")))[0]
                self.meshes[name] = tris * sc
        self.bodies, self.geoms, self.joints = [], [], []
        w = Body()
        w.name, w.parent, w.pos, w.quat, w.elem, w.id = "world", -1, np.zeros(3), np.array([1.0, 0, 0, 0]), None, 0
        w.joints, w.inertial = [], None
        self.bodies.append(w)
        for wb in root.findall("worldbody"):     # several <worldbody> elements are merged
            self._children(wb, 0, "")
        self.excludes = set()
        for c in root.findall("contact"):
            for e in c.findall("exclude"):
                self.excludes.add(frozenset((e.get("body1"), e.get("body2"))))
        self._place()
        self._inertia()

    # ---- defaults ----
    def _defaults(self, d, parent, top):
        name = d.get("class") or ("main" if top else None)
        attrs = {ch.tag: dict(ch.attrib) for ch in d if ch.tag != "default"}
        if name in self.classes and name == "main":
            self.classes["main"][0].update(attrs)
        else:
            self.classes[name] = (attrs, parent)
        for ch in d.findall("default"):
            self._defaults(ch, name, False)

    def _attr(self, e, kind, childclass, key, default=None):
        if e.get(key) is not None:
            return e.get(key)
        cname = e.get("class") or childclass or "main"
        while cname is not None:
            attrs, parent = self.classes[cname]
            if key in attrs.get(kind, {}):
                return attrs[kind][key]
            cname = parent if parent is not None else (None if cname == "main" else "main")
        return default

    def _vec(self, s, n=None):
        v = np.array([float(x) for x in s.split()])
        return v

    def _orient(self, e, kind, cc):
        q = self._attr(e, kind, cc, "quat")
        if q is not None:
            q = self._vec(q)
            n = np.linalg.norm(q)
            return q / n if n > 1e-14 else np.array([1.0, 0, 0, 0])   # a zero quaternion normalises to the identity
        eu = self._attr(e, kind, cc, "euler")
        if eu is not None:
            eu = self._vec(eu) * (math.pi / 180 if self.degree else 1)
            q = np.array([1.0, 0, 0, 0])
            for ch, ang in zip(self.eulerseq, eu):
                ax = {"x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1]}[ch.lower()]
                r = axis_angle_quat(ax, ang)
                q = quat_mul(q, r) if ch.islower() else quat_mul(r, q)
            return q
        aa = self._attr(e, kind, cc, "axisangle")
        if aa is not None:
            aa = self._vec(aa)
            return axis_angle_quat(aa[:3], aa[3] * (math.pi / 180 if self.degree else 1))
        za = self._attr(e, kind, cc, "zaxis")
        if za is not None:
            return zaxis_quat(self._vec(za))
        return np.array([1.0, 0, 0, 0])

    def _children(self, e, parent, cc):
        for ch in e:
            if ch.tag == "geom":
                self._geom(ch, parent, cc)
            elif ch.tag == "body":
                b = Body()
                b.id = len(self.bodies)
                b.name, b.parent, b.elem = ch.get("name"), parent, ch
                b.pos = self._vec(ch.get("pos", "0 0 0"))
                b.quat = self._orient(ch, "body", cc)
                b.joints, b.inertial = [], None
                self.bodies.append(b)
                bcc = ch.get("childclass", cc)
                for j in ch:
                    if j.tag in ("joint", "freejoint"):
                        jt = "free" if j.tag == "freejoint" else self._attr(j, "joint", bcc, "type", "hinge")
                        J = {"type": jt, "body": b.id, "name": j.get("name"),
                             "pos": self._vec(self._attr(j, "joint", bcc, "pos", "0 0 0")),
                             "axis": self._vec(self._attr(j, "joint", bcc, "axis", "0 0 1")),
                             "armature": float(self._attr(j, "joint", bcc, "armature", "0") if jt != "free" else 0)}
                        b.joints.append(J)
                        self.joints.append(J)
                    elif j.tag == "inertial":
                        b.inertial = j
                self._children(ch, b.id, bcc)

    def _geom(self, e, body, cc):
        g = Geom()
        g.id, g.body, g.name = len(self.geoms), body, e.get("name")
        g.type = self._attr(e, "geom", cc, "type", "sphere")
        g.mesh = self._attr(e, "geom", cc, "mesh")
        if g.mesh is not None:
            g.type = "mesh"
        size = self._attr(e, "geom", cc, "size")
        g.size = np.zeros(3)
        if size is not None:
            s = self._vec(size)
            g.size[:len(s)] = s
        g.pos = self._vec(self._attr(e, "geom", cc, "pos", "0 0 0"))
        g.quat = self._orient(e, "geom", cc)
        ft = self._attr(e, "geom", cc, "fromto")
        if ft is not None:
            ft = self._vec(ft)
            d = ft[3:] - ft[:3]
            g.pos = 0.5 * (ft[:3] + ft[3:])
            g.quat = zaxis_quat(d)
            half = 0.5 * np.linalg.norm(d)
            if g.type in ("capsule", "cylinder"):
                g.size[1] = half
            elif g.type in ("box", "ellipsoid"):
                g.size[2] = half
        g.density = float(self._attr(e, "geom", cc, "density", "1000"))
        m = self._attr(e, "geom", cc, "mass")
        g.mass = float(m) if m is not None else None
        g.contype = int(self._attr(e, "geom", cc, "contype", "1"))
        g.conaffinity = int(self._attr(e, "geom", cc, "conaffinity", "1"))
        self.geoms.append(g)

    # ---- placement at qpos0 ----
    def _place(self):
        for b in self.bodies:
            if b.parent < 0:
                b.xpos, b.xquat = np.zeros(3), np.array([1.0, 0, 0, 0])
            else:
                p = self.bodies[b.parent]
                b.xpos = p.xpos + quat_mat(p.xquat) @ b.pos
                b.xquat = quat_mul(p.xquat, b.quat)
            b.xmat = quat_mat(b.xquat)
        # dofs: per joint, (kind, axis in world, anchor in world, body)
        self.dofs = []
        for j in self.joints:
            b = self.bodies[j["body"]]
            anchor = b.xpos + b.xmat @ j["pos"]
            if j["type"] == "free":
                for k in range(3):
                    self.dofs.append(("lin", np.eye(3)[k], b.xpos, b.id, "free_t", 0.0))
                for k in range(3):
                    self.dofs.append(("rot", b.xmat[:, k], b.xpos, b.id, "free_r", 0.0))
            elif j["type"] == "ball":
                for k in range(3):
                    self.dofs.append(("rot", b.xmat[:, k], anchor, b.id, "ball", j["armature"]))
            else:
                ax = b.xmat @ (j["axis"] / np.linalg.norm(j["axis"]))
                self.dofs.append(("lin" if j["type"] == "slide" else "rot", ax, anchor, b.id, j["type"], j["armature"]))
        self.nv = len(self.dofs)

    def _ancestors(self, b):
        out = set()
        while b >= 0:
            out.add(b)
            b = self.bodies[b].parent
        return out

    # ---- inertia ----
    def _geom_inertia(self, g):
        """(mass, centre in geom frame, inertia tensor about that centre in geom axes)."""
        t, s, rho = g.type, g.size, g.density
        cen = np.zeros(3)
        if t == "sphere":
            vol = 4 / 3 * math.pi * s[0] ** 3
            I = np.eye(3) * 0.4 * s[0] ** 2 * vol
        elif t == "box":
            vol = 8 * s[0] * s[1] * s[2]
            I = np.diag([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2]) * vol / 3
        elif t == "cylinder":
            r, h = s[0], s[1]
            vol = math.pi * r * r * 2 * h
            I = np.diag([(3 * r * r + (2 * h) ** 2) / 12, (3 * r * r + (2 * h) ** 2) / 12, r * r / 2]) * vol
        elif t == "capsule":
            r, h = s[0], s[1]
            vc, vs = math.pi * r * r * 2 * h, 4 / 3 * math.pi * r ** 3
            vol = vc + vs
            ixx = vc * (3 * r * r + 4 * h * h) / 12 + vs * (0.4 * r * r + h * h + 0.75 * r * h)
            izz = vc * r * r / 2 + vs * 0.4 * r * r
            I = np.diag([ixx, ixx, izz])
        elif t == "ellipsoid":
            vol = 4 / 3 * math.pi * s[0] * s[1] * s[2]
            I = np.diag([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2]) * vol / 5
        elif t == "mesh":
            vol, cen, I = mesh_props(self.meshes[g.mesh])
        else:
            return 0.0, cen, np.zeros((3, 3))
        mass = vol * rho
        I = I * rho
        if g.mass is not None and vol > 0:
            I = I * (g.mass / mass)
            mass = g.mass
        return mass, cen, I

    def _inertia(self):
        for b in self.bodies:
            b.mass, b.ipos, b.I = 0.0, np.zeros(3), np.zeros((3, 3))
        for b in self.bodies[1:]:
            ine = b.inertial
            use_geoms = self.inertiafromgeom == "true" or (self.inertiafromgeom == "auto" and ine is None)
            if not use_geoms and ine is not None:
                b.mass = float(ine.get("mass"))
                b.ipos = self._vec(ine.get("pos", "0 0 0"))
                R = quat_mat(self._orient(ine, "inertial", ""))
                if ine.get("fullinertia") is not None:
                    f = self._vec(ine.get("fullinertia"))
                    Il = np.array([[f[0], f[3], f[4]], [f[3], f[1], f[5]], [f[4], f[5], f[2]]])
                else:
                    Il = np.diag(self._vec(ine.get("diaginertia", "0 0 0")))
                b.I = R @ Il @ R.T
            elif use_geoms:
                parts = []
                for g in self.geoms:
                    if g.body != b.id:
                        continue
                    mass, cen, I = self._geom_inertia(g)
                    if mass <= 0:
                        continue
                    R = quat_mat(g.quat)
                    parts.append((mass, g.pos + R @ cen, R @ I @ R.T))
                M = sum(p[0] for p in parts)
                if M > 0:
                    com = sum(p[0] * p[1] for p in parts) / M
                    I = np.zeros((3, 3))
                    for mass, c, Ig in parts:
                        d = c - com
                        I += Ig + mass * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
                    b.mass, b.ipos, b.I = M, com, I
            if b.mass < self.boundmass:
                b.mass = self.boundmass
            if self.boundinertia > 0:
                ev, evec = np.linalg.eigh(b.I)
                b.I = evec @ np.diag(np.maximum(ev, self.boundinertia)) @ evec.T

    # ---- derived quantities ----
    def mass_matrix(self):
        M = np.zeros((self.nv, self.nv))
        self._jac = {}
        for b in self.bodies[1:]:
            com = b.xpos + b.xmat @ b.ipos
            anc = self._ancestors(b.id)
            Jp, Jr = np.zeros((3, self.nv)), np.zeros((3, self.nv))
            for i, (kind, ax, anchor, body, _, _) in enumerate(self.dofs):
                if body not in anc:
                    continue
                if kind == "lin":
                    Jp[:, i] = ax
                else:
                    Jr[:, i] = ax
                    Jp[:, i] = np.cross(ax, com - anchor)
            self._jac[b.id] = (Jp, Jr)
            Iw = b.xmat @ b.I @ b.xmat.T
            M += b.mass * Jp.T @ Jp + Jr.T @ Iw @ Jr
        M += np.diag([d[5] for d in self.dofs])
        return M

    def invweights(self):
        M = self.mass_matrix()
        if self.nv == 0:
            return np.zeros(0), np.zeros((len(self.bodies), 2))
        Mi = np.linalg.inv(M)
        dw = np.diag(Mi).copy()
        i = 0
        while i < self.nv:
            tag = self.dofs[i][4]
            if tag in ("free_t", "free_r", "ball"):
                dw[i:i + 3] = dw[i:i + 3].mean()
                i += 3
            else:
                i += 1
        bw = np.zeros((len(self.bodies), 2))
        for b in self.bodies[1:]:
            Jp, Jr = self._jac[b.id]
            if not Jp.any() and not Jr.any():
                continue
            bw[b.id, 0] = np.trace(Jp @ Mi @ Jp.T) / 3
            bw[b.id, 1] = np.trace(Jr @ Mi @ Jr.T) / 3
        return dw, bw

    def weld_id(self, b):
        """Topmost body this body is rigidly attached to (bodies without joints are welded to their parent)."""
        while b > 0 and not self.bodies[b].joints:
            b = self.bodies[b].parent
        return b

    def pairs(self):
        out = []
        nb = len(self.bodies)
        weld = [self.weld_id(b) for b in range(nb)]
        wparent = [weld[self.bodies[weld[b]].parent] if weld[b] > 0 else -1 for b in range(nb)]
        by_body = {}
        for g in self.geoms:
            by_body.setdefault(g.body, []).append(g)
        for b1 in range(nb):
            for b2 in range(b1 + 1, nb):
                if b1 not in by_body or b2 not in by_body:
                    continue
                w1, w2 = weld[b1], weld[b2]
                if w1 == w2:                                   # same rigid group (includes static-static)
                    continue
                # parent-child filter on the welded bodies, unless the parent is the static world group
                if (wparent[b1] == w2 and w2 != 0) or (wparent[b2] == w1 and w1 != 0):
                    continue
                if frozenset((self.bodies[b1].name, self.bodies[b2].name)) in self.excludes:
                    continue
                for ga in by_body[b1]:
                    for gb in by_body[b2]:
                        if not ((ga.contype & gb.conaffinity) or (gb.contype & ga.conaffinity)):
                            continue
                        out.append((ga.id, gb.id))
        return out
