// urdf_import.cpp — URDF -> MJCF translation behind mj_loadXML.
//
// The reference's importer (src/mujoco_compile.cpp:317-405) copies the URDF next to its meshes, injects a
// <mujoco><compiler balanceinertia discardvisual boundmass boundinertia meshdir strippath/></mujoco> element
// (:116-195), hands the file to mj_loadXML (:404) — libmujoco parses URDF natively — and writes the result back as
// MJCF with mj_saveLastXML (:470).  This file is that native URDF reader: <robot> is translated to an MJCF document
// (kept as the model's source text, so mj_saveLastXML emits MJCF exactly as the reference expects) and compiled by the
// ordinary MJCF compiler.  Conventions follow MuJoCo's documented URDF semantics:
//   link -> body; joint origin xyz / rpy (fixed-axis roll-pitch-yaw, R = Rz(y) Ry(p) Rx(r)) -> body pos / quat in the
//   parent link frame; revolute -> limited hinge, continuous -> hinge, prismatic -> limited slide, floating -> free,
//   fixed -> no joint; <dynamics damping friction> -> damping / frictionloss; <limit lower upper> -> range;
//   <inertial> origin + mass + ixx..izz -> inertial pos / mass / fullinertia rotated into the link frame;
//   <collision> box / cylinder / sphere / mesh -> geoms (box extents and cylinder length halved); angles in radians;
//   <visual> elements are dropped (the reference always sets discardvisual, src/mujoco_compile.cpp:158).
// fusestatic (libmujoco's default for URDF; `<compiler fusestatic="false"/>` in the <mujoco> extension turns it off): a link
// behind a fixed joint is folded into the body of its parent link — its collision geoms move there with the composed
// transform, its inertial is combined with the parent's about the common CoM, its movable children attach to the
// parent — and the root link is the world body (the reference wraps the result in a named body afterwards,
// src/mujoco_compile.cpp:197-218).  With fusestatic off every non-root link stays a body of its own.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "hostmath.h"
#include "xml_lite.h"

namespace b2 {

namespace {

using namespace hm;

[[noreturn]] void ufail(const std::string& msg) { throw std::runtime_error("URDF: " + msg); }

int nums(const char* s, double* out, int maxn) {
  int n = 0;
  if (!s) return 0;
  char* end = nullptr;
  while (n < maxn) {
    const double v = std::strtod(s, &end);
    if (end == s) break;
    out[n++] = v;
    s = end;
  }
  return n;
}

std::string fmt(const double* v, int n) {
  std::string r;
  char buf[40];
  for (int i = 0; i < n; i++) {
    std::snprintf(buf, sizeof(buf), "%.17g", v[i]);
    if (i) r += ' ';
    r += buf;
  }
  return r;
}
std::string fmt1(double v) { return fmt(&v, 1); }

std::string esc(const std::string& s) {
  std::string r;
  for (char ch : s) {
    if (ch == '&') r += "&amp;";
    else if (ch == '<') r += "&lt;";
    else if (ch == '"') r += "&quot;";
    else r += ch;
  }
  return r;
}

// <origin xyz rpy>: position and quaternion (identity when absent)
void origin_of(const XmlElem* e, double* pos, double* quat) {
  pos[0] = pos[1] = pos[2] = 0;
  quat[0] = 1; quat[1] = quat[2] = quat[3] = 0;
  const XmlElem* o = e ? e->child("origin") : nullptr;
  if (!o) return;
  if (const char* s = o->attr("xyz")) if (nums(s, pos, 3) != 3) ufail("origin xyz needs 3 numbers");
  if (const char* s = o->attr("rpy")) {
    double rpy[3];
    if (nums(s, rpy, 3) != 3) ufail("origin rpy needs 3 numbers");
    const double ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1};
    double qx[4], qy[4], qz[4], t[4];
    axis_angle2quat(qx, ax, rpy[0]);
    axis_angle2quat(qy, ay, rpy[1]);
    axis_angle2quat(qz, az, rpy[2]);
    mul_quat(t, qz, qy);
    mul_quat(quat, t, qx);
    normalize4(quat);
  }
}

struct Frame { double p[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0}; };   // pose of a link frame in the frame of the body that owns it
inline bool is_identity(const Frame& f) { return f.p[0] == 0 && f.p[1] == 0 && f.p[2] == 0 && f.q[0] == 1 && f.q[1] == 0 && f.q[2] == 0 && f.q[3] == 0; }
// a o b: b expressed in a's parent
Frame compose(const Frame& a, const double* bp, const double* bq) {
  if (is_identity(a)) { Frame r; copy3(r.p, bp); copy4(r.q, bq); return r; }   // (keeps the numbers of the unfused case bit for bit)
  Frame r;
  double t[3];
  rot_vec_quat(t, bp, a.q);
  for (int k = 0; k < 3; k++) r.p[k] = a.p[k] + t[k];
  mul_quat(r.q, a.q, bq);
  normalize4(r.q);
  return r;
}

struct Joint {
  const XmlElem* e;
  std::string name, type, parent, child;
};

struct Importer {
  const XmlElem& robot;
  std::map<std::string, const XmlElem*> links;
  std::map<std::string, std::vector<const Joint*>> children;  // by parent link, in document order
  std::vector<Joint> joints;
  std::map<std::string, std::string> mesh_names;  // "file|scale" -> asset name
  std::vector<std::string> mesh_assets;           // <mesh .../> lines
  bool strippath = true;
  bool fusestatic = true;   // libmujoco's default for URDF: links behind fixed joints are folded into their parent body
  std::ostringstream out;

  explicit Importer(const XmlElem& r) : robot(r) {}

  std::string mesh_asset(const XmlElem& mesh) {
    const char* fn = mesh.attr("filename");
    if (!fn) ufail("<mesh> without filename");
    std::string file = fn;
    const std::string pk = "package://";
    if (file.compare(0, pk.size(), pk) == 0) file = file.substr(pk.size());   // resolved against meshdir
    if (strippath) {
      const size_t sl = file.find_last_of("/\\");
      if (sl != std::string::npos) file = file.substr(sl + 1);
    }
    const std::string scale = mesh.attr("scale") ? mesh.attr("scale") : "";
    const std::string key = file + "|" + scale;
    auto it = mesh_names.find(key);
    if (it != mesh_names.end()) return it->second;
    std::string base = file;
    const size_t sl = base.find_last_of("/\\");
    if (sl != std::string::npos) base = base.substr(sl + 1);
    const size_t dot = base.find_last_of('.');
    if (dot != std::string::npos) base = base.substr(0, dot);
    std::string name = base;
    std::set<std::string> used;
    for (auto& kv : mesh_names) used.insert(kv.second);
    for (int k = 1; used.count(name); k++) name = base + "_" + std::to_string(k);
    mesh_names[key] = name;
    std::string line = "    <mesh name=\"" + esc(name) + "\" file=\"" + esc(file) + "\"";
    if (!scale.empty()) line += " scale=\"" + esc(scale) + "\"";
    mesh_assets.push_back(line + "/>");
    return name;
  }

  void emit_geoms(const XmlElem& link, const std::string& ind, const Frame& fr = Frame()) {
    for (auto& ch : link.children) {
      if (ch->name != "collision") continue;
      const XmlElem* geo = ch->child("geometry");
      if (!geo || geo->children.empty()) continue;
      double pos[3], quat[4];
      origin_of(ch.get(), pos, quat);
      { const Frame g = compose(fr, pos, quat); copy3(pos, g.p); copy4(quat, g.q); }
      const XmlElem& sh = *geo->children[0];
      std::string a;
      if (sh.name == "box") {
        double s[3];
        if (nums(sh.attr("size"), s, 3) != 3) ufail("box needs size=\"x y z\"");
        for (double& x : s) x *= 0.5;
        a = "type=\"box\" size=\"" + fmt(s, 3) + "\"";
      } else if (sh.name == "cylinder") {
        double s[2] = {0, 0};
        if (nums(sh.attr("radius"), &s[0], 1) != 1 || nums(sh.attr("length"), &s[1], 1) != 1) ufail("cylinder needs radius and length");
        s[1] *= 0.5;
        a = "type=\"cylinder\" size=\"" + fmt(s, 2) + "\"";
      } else if (sh.name == "sphere") {
        double r;
        if (nums(sh.attr("radius"), &r, 1) != 1) ufail("sphere needs radius");
        a = "type=\"sphere\" size=\"" + fmt1(r) + "\"";
      } else if (sh.name == "capsule") {   // not in the URDF standard, accepted by several toolchains
        double s[2] = {0, 0};
        if (nums(sh.attr("radius"), &s[0], 1) != 1 || nums(sh.attr("length"), &s[1], 1) != 1) ufail("capsule needs radius and length");
        s[1] *= 0.5;
        a = "type=\"capsule\" size=\"" + fmt(s, 2) + "\"";
      } else if (sh.name == "mesh") {
        a = "type=\"mesh\" mesh=\"" + esc(mesh_asset(sh)) + "\"";
      } else {
        ufail("unsupported collision geometry <" + sh.name + ">");
      }
      out << ind << "<geom " << a << " pos=\"" << fmt(pos, 3) << "\" quat=\"" << fmt(quat, 4) << "\"";
      if (const char* n = ch->attr("name")) out << " name=\"" << esc(n) << "\"";
      out << "/>\n";
    }
  }

  struct Inert { double m = 0, c[3] = {0, 0, 0}, I[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; bool any = false; };   // about its own CoM, body axes

  // <inertial> of a link, expressed in the frame of the body that owns the link (fr = pose of the link frame in it)
  Inert read_inertial(const XmlElem& link, const Frame& fr) {
    Inert r;
    const XmlElem* in = link.child("inertial");
    if (!in) return r;
    r.any = true;
    double pos[3], quat[4];
    origin_of(in, pos, quat);
    const Frame f = compose(fr, pos, quat);
    copy3(r.c, f.p);
    if (const XmlElem* me = in->child("mass")) nums(me->attr("value"), &r.m, 1);
    double I[6] = {0, 0, 0, 0, 0, 0};  // ixx iyy izz ixy ixz iyz
    if (const XmlElem* ie = in->child("inertia")) {
      const char* keys[6] = {"ixx", "iyy", "izz", "ixy", "ixz", "iyz"};
      for (int k = 0; k < 6; k++) nums(ie->attr(keys[k]), &I[k], 1);
    }
    // rotate the tensor from the inertial frame into the body frame: R I R^T
    double R[9], A[9] = {I[0], I[3], I[4], I[3], I[1], I[5], I[4], I[5], I[2]}, T[9];
    quat2mat(R, f.q);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += R[3 * i + k] * A[3 * k + j];
        T[3 * i + j] = s;
      }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += T[3 * i + k] * R[3 * j + k];
        r.I[3 * i + j] = s;
      }
    return r;
  }

  // one <inertial> for the body: a single part as it is, several (fused links) combined about the common CoM
  void emit_inertial(const std::vector<Inert>& parts, const std::string& ind) {
    std::vector<const Inert*> have;
    for (auto& p : parts) if (p.any) have.push_back(&p);
    if (have.empty()) return;
    Inert t = *have[0];
    if (have.size() > 1) {
      t = Inert();
      for (auto* p : have) { t.m += p->m; for (int k = 0; k < 3; k++) t.c[k] += p->m * p->c[k]; }
      if (t.m > 0) for (int k = 0; k < 3; k++) t.c[k] /= t.m;
      else for (int k = 0; k < 3; k++) t.c[k] = have[0]->c[k];
      for (auto* p : have) {
        const double d[3] = {p->c[0] - t.c[0], p->c[1] - t.c[1], p->c[2] - t.c[2]};
        const double d2 = dot3(d, d);
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) t.I[3 * i + j] += p->I[3 * i + j] + p->m * ((i == j ? d2 : 0.0) - d[i] * d[j]);
      }
    }
    const double full[6] = {t.I[0], t.I[4], t.I[8], t.I[1], t.I[2], t.I[5]};
    out << ind << "<inertial pos=\"" << fmt(t.c, 3) << "\" mass=\"" << fmt1(t.m) << "\" fullinertia=\"" << fmt(full, 6) << "\"/>\n";
  }

  void emit_joint(const Joint& j, const std::string& ind) {
    if (j.type == "fixed") return;
    if (j.type == "floating") { out << ind << "<freejoint name=\"" << esc(j.name) << "\"/>\n"; return; }
    std::string type;
    bool limited = false;
    if (j.type == "revolute") { type = "hinge"; limited = true; }
    else if (j.type == "continuous") type = "hinge";
    else if (j.type == "prismatic") { type = "slide"; limited = true; }
    else ufail("joint '" + j.name + "': unsupported type '" + j.type + "'");
    double axis[3] = {1, 0, 0};
    if (const XmlElem* ax = j.e->child("axis")) if (nums(ax->attr("xyz"), axis, 3) != 3) ufail("joint '" + j.name + "': axis needs xyz");
    out << ind << "<joint name=\"" << esc(j.name) << "\" type=\"" << type << "\" pos=\"0 0 0\" axis=\"" << fmt(axis, 3) << "\"";
    const XmlElem* lim = j.e->child("limit");
    if (limited) {
      double r[2] = {0, 0};
      if (lim) { nums(lim->attr("lower"), &r[0], 1); nums(lim->attr("upper"), &r[1], 1); }
      if (r[0] < r[1]) out << " limited=\"true\" range=\"" << fmt(r, 2) << "\"";
      else out << " limited=\"false\"";
    } else {
      out << " limited=\"false\"";
    }
    if (const XmlElem* dyn = j.e->child("dynamics")) {
      double v;
      if (nums(dyn->attr("damping"), &v, 1) == 1) out << " damping=\"" << fmt1(v) << "\"";
      if (nums(dyn->attr("friction"), &v, 1) == 1) out << " frictionloss=\"" << fmt1(v) << "\"";
    }
    out << "/>\n";
  }

  struct Part { const XmlElem* link; Frame fr; };
  struct Kid { const Joint* j; Frame fr; };   // fr: pose of the joint's parent link frame in the owning body

  // the link and, with fusestatic, every link welded to it by fixed joints; the movable joints leaving the group
  void gather(const std::string& name, const Frame& fr, const Joint* via, std::vector<Part>& parts, std::vector<Kid>& kids, std::set<std::string>& open) {
    auto it = links.find(name);
    if (it == links.end()) ufail("joint '" + (via ? via->name : std::string("?")) + "' refers to unknown link '" + name + "'");
    if (!open.insert(name).second) ufail("kinematic loop through link '" + name + "'");
    parts.push_back({it->second, fr});
    for (const Joint* c : children[name]) {
      if (fusestatic && c->type == "fixed") {
        double p[3], q[4];
        origin_of(c->e, p, q);
        gather(c->child, compose(fr, p, q), c, parts, kids, open);
      } else {
        kids.push_back({c, fr});
      }
    }
  }

  // body of link `name` reached through joint `via` (null: the root link, which IS the world body, as with libmujoco's
  // fusestatic: its collision geoms become world geoms, its inertial is irrelevant, its children hang off the world)
  void emit_link(const std::string& name, const Joint* via, const Frame& at, int depth, std::set<std::string>& open) {
    const std::string ind(4 + 2 * depth, ' ');
    std::vector<Part> parts;
    std::vector<Kid> kids;
    std::set<std::string> mine;
    gather(name, Frame(), via, parts, kids, mine);
    for (auto& n : mine) if (!open.insert(n).second) ufail("kinematic loop through link '" + n + "'");
    // libmujoco keeps the root link as a body of its own under the world unless it is called "world"; only fusestatic
    // folds it into the world body
    const bool as_body = via || (!fusestatic && name != "world");
    const std::string in2 = as_body ? ind + "  " : ind;
    if (as_body) {
      out << ind << "<body name=\"" << esc(name) << "\" pos=\"" << fmt(at.p, 3) << "\" quat=\"" << fmt(at.q, 4) << "\">\n";
      std::vector<Inert> inert;
      for (auto& p : parts) inert.push_back(read_inertial(*p.link, p.fr));
      emit_inertial(inert, in2);
      if (via) emit_joint(*via, in2);
    }
    for (auto& p : parts) emit_geoms(*p.link, in2, p.fr);
    for (auto& k : kids) {
      double p[3], q[4];
      origin_of(k.j->e, p, q);
      emit_link(k.j->child, k.j, compose(k.fr, p, q), as_body ? depth + 1 : depth, open);
    }
    if (as_body) out << ind << "</body>\n";
    for (auto& n : mine) open.erase(n);
  }

  std::string run() {
    for (auto& ch : robot.children) {
      if (ch->name == "link") {
        const char* n = ch->attr("name");
        if (!n) ufail("<link> without name");
        if (!links.emplace(n, ch.get()).second) ufail(std::string("duplicate link '") + n + "'");
      }
    }
    if (links.empty()) ufail("no <link> in <robot>");
    for (auto& ch : robot.children) {
      if (ch->name != "joint") continue;
      Joint j;
      j.e = ch.get();
      j.name = ch->attr("name") ? ch->attr("name") : "";
      j.type = ch->attr("type") ? ch->attr("type") : "";
      const XmlElem *p = ch->child("parent"), *c = ch->child("child");
      if (!p || !c || !p->attr("link") || !c->attr("link")) ufail("joint '" + j.name + "' needs <parent link> and <child link>");
      j.parent = p->attr("link"); j.child = c->attr("link");
      joints.push_back(j);
    }
    std::set<std::string> has_parent;
    for (const Joint& j : joints) {
      if (!has_parent.insert(j.child).second) ufail("link '" + j.child + "' has two parent joints");
      children[j.parent].push_back(&j);
    }
    std::string root;
    for (auto& ch : robot.children)   // the root is the (first) link that is nobody's child, in document order
      if (ch->name == "link" && !has_parent.count(ch->attr("name"))) { root = ch->attr("name"); break; }
    if (root.empty()) ufail("no root link (every link is some joint's child)");

    // the <mujoco> extension element: compiler / option / size / ... children are carried over
    const XmlElem* ext = robot.child("mujoco");
    std::vector<std::pair<std::string, std::string>> comp = {{"angle", "radian"}};
    std::ostringstream extra;
    if (ext) {
      for (auto& ch : ext->children) {
        if (ch->name == "compiler") {
          for (auto& kv : ch->attrs) {
            if (kv.first == "strippath") { strippath = kv.second == "true"; continue; }
            if (kv.first == "fusestatic") { fusestatic = kv.second == "true"; continue; }
            if (kv.first == "discardvisual" || kv.first == "angle") continue;
            comp.emplace_back(kv.first, kv.second);
          }
        } else {
          extra << "  <" << ch->name;
          for (auto& kv : ch->attrs) extra << " " << kv.first << "=\"" << esc(kv.second) << "\"";
          extra << "/>\n";   // extension children with nested content are not used by the reference
        }
      }
    }
    std::set<std::string> open;
    emit_link(root, nullptr, Frame(), 0, open);
    const std::string bodies = out.str();

    std::ostringstream doc;
    doc << "<mujoco model=\"" << esc(robot.attr("name") ? robot.attr("name") : "robot") << "\">\n";
    doc << "  <compiler";
    for (auto& kv : comp) doc << " " << kv.first << "=\"" << esc(kv.second) << "\"";
    doc << "/>\n" << extra.str();
    if (!mesh_assets.empty()) {
      doc << "  <asset>\n";
      for (auto& l : mesh_assets) doc << l << "\n";
      doc << "  </asset>\n";
    }
    doc << "  <worldbody>\n" << bodies << "  </worldbody>\n</mujoco>\n";
    return doc.str();
  }
};

}  // namespace

std::string urdf_to_mjcf(const XmlElem& robot) {
  if (robot.name != "robot") ufail("root element must be <robot>");
  Importer imp(robot);
  return imp.run();
}

}  // namespace b2
