// mjcf_compile.cpp — MJCF (subset) -> mjModel compiler for the batched engine.
//
// Covers the tags/attributes the reference's shipped models and run-time XML rewrites use
// (SURVEY.md Appendix B tally; reference src/mujoco_sim/mj_sim.cpp:185-457 for gravcomp / odom joints /
// pose_init rewrites, src/mujoco_compile.cpp:146-165,219-314 for compiler flags, mimic equalities and excludes).
// Semantics follow MuJoCo's public "XML reference" / "Modeling" chapters: depth-first body numbering,
// geom-inferred inertia at density 1000, angle="degree" by default, eulerseq "xyz", and so on.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

#include "hostmath.h"
#include "model_store.h"
#include "xml_lite.h"

namespace b2 {
using namespace hm;

void set_const(ModelStore& S);  // set0.cpp: invweight0, subtreemass, meaninertia
std::string urdf_to_mjcf(const XmlElem& robot);  // urdf_import.cpp

namespace {

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error(msg); }

std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) fail("cannot open file '" + path + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
std::string dir_of(const std::string& path) {
  size_t k = path.find_last_of('/');
  return k == std::string::npos ? std::string(".") : path.substr(0, k);
}
std::string join_path(const std::string& dir, const std::string& file) {
  if (!file.empty() && file[0] == '/') return file;
  if (dir.empty()) return file;
  return dir + "/" + file;
}

int parse_nums(const char* s, double* out, int maxn) {
  int n = 0;
  const char* p = s;
  while (*p && n < maxn) {
    char* end;
    double v = std::strtod(p, &end);
    if (end == p) {
      if (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == ',') { p++; continue; }
      fail(std::string("bad number in attribute value '") + s + "'");
    }
    out[n++] = v;
    p = end;
  }
  return n;
}

bool parse_bool(const char* s, const char* what) {
  if (!std::strcmp(s, "true") || !std::strcmp(s, "enable")) return true;
  if (!std::strcmp(s, "false") || !std::strcmp(s, "disable")) return false;
  fail(std::string("bad boolean '") + s + "' for " + what);
}

// One <default> class: attribute dictionaries per element kind, with a parent chain.
struct DefClass {
  const DefClass* parent = nullptr;
  std::map<std::string, std::map<std::string, std::string>> kinds;
  const char* find(const char* kind, const char* key) const {
    for (const DefClass* c = this; c; c = c->parent) {
      auto k = c->kinds.find(kind);
      if (k == c->kinds.end()) continue;
      auto a = k->second.find(key);
      if (a != k->second.end()) return a->second.c_str();
    }
    return nullptr;
  }
};

struct MeshAsset {
  std::string name;
  std::vector<double> vert;   // unique vertices, xyz
  std::vector<int> face;      // triangle indices (may be empty for point clouds)
  double center[3] = {0, 0, 0};  // what was subtracted from the file's vertices (see recenter_mesh)
};

struct Compiler {
  bool degree = true;
  std::string eulerseq = "xyz";
  std::string meshdir;
  bool autolimits = false;
  double boundmass = 0, boundinertia = 0;
  bool balanceinertia = false;
  int inertiafromgeom = 2;  // 0 false, 1 true, 2 auto
};

struct Ctx {
  ModelStore& S;
  Compiler comp;
  std::string basedir;
  std::map<std::string, DefClass> classes;  // "main" always present
  std::vector<MeshAsset> meshes;
  std::vector<std::string> body_names, jnt_names, geom_names;
  double default_density = 1000.0;
  explicit Ctx(ModelStore& s) : S(s) { classes["main"]; }
};

// attribute lookup: element first, then the default chain of its class
const char* lookup(const Ctx& c, const XmlElem& e, const char* kind, const std::string& childclass, const char* key) {
  if (const char* v = e.attr(key)) return v;
  const char* cls = e.attr("class");
  std::string cname = cls ? cls : (childclass.empty() ? "main" : childclass);
  auto it = c.classes.find(cname);
  if (it == c.classes.end()) fail("unknown default class '" + cname + "'");
  return it->second.find(kind, key);
}

void parse_defaults(Ctx& c, const XmlElem& d, const DefClass* parent, bool top) {
  std::string cname = d.attr("class") ? d.attr("class") : (top ? "main" : "");
  if (cname.empty()) fail("nested <default> needs a class name");
  DefClass& dc = c.classes[cname];
  if (cname != "main" || parent) dc.parent = parent ? parent : nullptr;
  for (auto& ch : d.children) {
    if (ch->name == "default") continue;
    auto& dict = dc.kinds[ch->name];
    for (auto& kv : ch->attrs) dict[kv.first] = kv.second;
  }
  for (auto& ch : d.children)
    if (ch->name == "default") parse_defaults(c, *ch, &c.classes[cname], false);
}

void resolve_includes(XmlElem& e, const std::string& dir, int depth) {
  if (depth > 16) fail("<include> nesting too deep");
  for (size_t i = 0; i < e.children.size();) {
    XmlElem& ch = *e.children[i];
    if (ch.name == "include") {
      const char* f = ch.attr("file");
      if (!f) fail("<include> without file");
      std::string path = join_path(dir, f);
      std::string text = read_file(path);
      auto root = XmlParser(text).parse();
      resolve_includes(*root, dir, depth + 1);
      std::vector<std::unique_ptr<XmlElem>> kids = std::move(root->children);
      e.children.erase(e.children.begin() + i);
      for (size_t k = 0; k < kids.size(); k++) e.children.insert(e.children.begin() + i + k, std::move(kids[k]));
      i += kids.size();
    } else {
      resolve_includes(ch, dir, depth);
      i++;
    }
  }
}

void euler2quat(const Ctx& c, const double* e_in, double* q) {
  double e[3] = {e_in[0], e_in[1], e_in[2]};
  if (c.comp.degree) for (double& x : e) x *= mjPI / 180.0;
  q[0] = 1; q[1] = q[2] = q[3] = 0;
  for (int i = 0; i < 3; i++) {
    char ax = c.comp.eulerseq[i];
    double a[3] = {0, 0, 0};
    int k = (ax == 'x' || ax == 'X') ? 0 : (ax == 'y' || ax == 'Y') ? 1 : 2;
    a[k] = 1;
    double r[4], t[4];
    axis_angle2quat(r, a, e[i]);
    // lower-case: rotate about the moving (intrinsic) axes -> post-multiply; upper-case: fixed axes
    if (ax >= 'a') mul_quat(t, q, r); else mul_quat(t, r, q);
    copy4(q, t);
  }
  normalize4(q);
}

// orientation of an element from quat / euler / axisangle / xyaxes / zaxis
void parse_orientation(const Ctx& c, const XmlElem& e, const char* kind, const std::string& cc, double* q) {
  q[0] = 1; q[1] = q[2] = q[3] = 0;
  double v[6];
  if (const char* s = lookup(c, e, kind, cc, "quat")) {
    if (parse_nums(s, v, 4) != 4) fail("quat needs 4 numbers");
    copy4(q, v);
    if (normalize4(q) < 1e-15) { q[0] = 1; q[1] = q[2] = q[3] = 0; }  // cat.xml:7 authors an all-zero quat
  } else if (const char* s = lookup(c, e, kind, cc, "euler")) {
    if (parse_nums(s, v, 3) != 3) fail("euler needs 3 numbers");
    euler2quat(c, v, q);
  } else if (const char* s = lookup(c, e, kind, cc, "axisangle")) {
    if (parse_nums(s, v, 4) != 4) fail("axisangle needs 4 numbers");
    double ang = c.comp.degree ? v[3] * mjPI / 180.0 : v[3];
    normalize3(v);
    axis_angle2quat(q, v, ang);
  } else if (const char* s = lookup(c, e, kind, cc, "zaxis")) {
    if (parse_nums(s, v, 3) != 3) fail("zaxis needs 3 numbers");
    normalize3(v);
    quat_z2vec(q, v);
  } else if (const char* s = lookup(c, e, kind, cc, "xyaxes")) {
    if (parse_nums(s, v, 6) != 6) fail("xyaxes needs 6 numbers");
    double x[3] = {v[0], v[1], v[2]}, y[3] = {v[3], v[4], v[5]}, z[3];
    normalize3(x);
    double d = dot3(x, y);
    for (int k = 0; k < 3; k++) y[k] -= d * x[k];
    normalize3(y);
    cross(z, x, y);
    double m[9] = {x[0], y[0], z[0], x[1], y[1], z[1], x[2], y[2], z[2]};
    mat2quat(q, m);
  }
}

// ---------- mesh loading ----------
struct V3Less {
  bool operator()(const std::array<float, 3>& a, const std::array<float, 3>& b) const { return a < b; }
};

void load_stl(const std::string& path, MeshAsset& ma) {
  std::string buf = read_file(path);
  std::vector<std::array<float, 3>> tri_verts;
  bool ascii = false;
  if (buf.size() >= 84) {
    uint32_t ntri;
    std::memcpy(&ntri, buf.data() + 80, 4);
    if (84 + (size_t)ntri * 50 != buf.size()) ascii = buf.compare(0, 5, "solid") == 0;
    if (!ascii) {
      if (84 + (size_t)ntri * 50 > buf.size()) fail("truncated binary STL '" + path + "'");
      tri_verts.reserve((size_t)ntri * 3);
      for (uint32_t t = 0; t < ntri; t++) {
        const char* p = buf.data() + 84 + (size_t)t * 50 + 12;
        for (int k = 0; k < 3; k++) {
          std::array<float, 3> v;
          std::memcpy(v.data(), p + 12 * k, 12);
          tri_verts.push_back(v);
        }
      }
    }
  } else {
    ascii = true;
  }
  if (ascii) {
    std::istringstream ss(buf);
    std::string tok;
    while (ss >> tok)
      if (tok == "vertex") {
        std::array<float, 3> v;
        ss >> v[0] >> v[1] >> v[2];
        tri_verts.push_back(v);
      }
  }
  std::map<std::array<float, 3>, int, V3Less> uniq;
  for (auto& v : tri_verts) {
    auto it = uniq.find(v);
    int id;
    if (it == uniq.end()) {
      id = (int)uniq.size();
      uniq.emplace(v, id);
      ma.vert.push_back(v[0]); ma.vert.push_back(v[1]); ma.vert.push_back(v[2]);
    } else {
      id = it->second;
    }
    ma.face.push_back(id);
  }
}

void load_obj(const std::string& path, MeshAsset& ma) {
  std::istringstream ss(read_file(path));
  std::string line;
  while (std::getline(ss, line)) {
    if (line.size() > 2 && line[0] == 'v' && line[1] == ' ') {
      double v[3];
      if (parse_nums(line.c_str() + 2, v, 3) == 3) { ma.vert.push_back(v[0]); ma.vert.push_back(v[1]); ma.vert.push_back(v[2]); }
    } else if (line.size() > 2 && line[0] == 'f' && line[1] == ' ') {
      std::istringstream ls(line.substr(2));
      std::string tok;
      std::vector<int> idx;
      while (ls >> tok) idx.push_back(std::atoi(tok.c_str()) - 1);
      for (size_t k = 1; k + 1 < idx.size(); k++) { ma.face.push_back(idx[0]); ma.face.push_back(idx[k]); ma.face.push_back(idx[k + 1]); }
    }
  }
}

// MuJoCo stores a mesh about its own centre (its CoM) and moves the geom frame accordingly.  That is not cosmetic for
// collision: geom_xpos is the interior point the convex routine (MPR) starts its portal from and the centre of the
// bounding sphere, and a link mesh authored far from its file origin would have neither.  Centre = volume centroid of the
// closed surface; vertex mean when the surface encloses no volume (point clouds, open sheets).
void recenter_mesh(MeshAsset& ma) {
  const size_t nvt = ma.vert.size() / 3;
  if (nvt == 0) return;
  double c[3] = {0, 0, 0}, vol = 0;
  for (size_t f = 0; f + 2 < ma.face.size(); f += 3) {
    const double *a = &ma.vert[3 * ma.face[f]], *b = &ma.vert[3 * ma.face[f + 1]], *d = &ma.vert[3 * ma.face[f + 2]];
    const double v6 = a[0] * (b[1] * d[2] - b[2] * d[1]) - a[1] * (b[0] * d[2] - b[2] * d[0]) + a[2] * (b[0] * d[1] - b[1] * d[0]);
    vol += v6;
    for (int k = 0; k < 3; k++) c[k] += v6 * (a[k] + b[k] + d[k]) * 0.25;
  }
  double ext = 0;
  for (double x : ma.vert) ext = std::max(ext, std::fabs(x));
  if (std::fabs(vol) > 1e-9 * ext * ext * ext && ext > 0) {
    for (int k = 0; k < 3; k++) c[k] /= vol;
  } else {
    c[0] = c[1] = c[2] = 0;
    for (size_t i = 0; i < nvt; i++) for (int k = 0; k < 3; k++) c[k] += ma.vert[3 * i + k] / (double)nvt;
  }
  for (size_t i = 0; i < nvt; i++) for (int k = 0; k < 3; k++) ma.vert[3 * i + k] -= c[k];
  for (int k = 0; k < 3; k++) ma.center[k] = c[k];
}

// volume, centre of mass and inertia tensor (about the CoM, unit density) of a closed triangle mesh
void mesh_mass_props(const MeshAsset& ma, double& vol, double* com, double* I) {
  vol = 0;
  zero3(com);
  double P[6] = {0, 0, 0, 0, 0, 0};  // integrals of xx, yy, zz, xy, xz, yz
  for (size_t f = 0; f + 2 < ma.face.size(); f += 3) {
    const double* a = &ma.vert[3 * ma.face[f]];
    const double* b = &ma.vert[3 * ma.face[f + 1]];
    const double* c = &ma.vert[3 * ma.face[f + 2]];
    double bc[3];
    cross(bc, b, c);
    double v = dot3(a, bc) / 6.0;  // signed tetra volume with the origin
    vol += v;
    for (int k = 0; k < 3; k++) com[k] += v * (a[k] + b[k] + c[k]) / 4.0;
    auto second = [&](int i, int j) {
      return v / 20.0 * (2 * a[i] * a[j] + 2 * b[i] * b[j] + 2 * c[i] * c[j] + a[i] * b[j] + a[j] * b[i] + a[i] * c[j] +
                         a[j] * c[i] + b[i] * c[j] + b[j] * c[i]);
    };
    P[0] += second(0, 0); P[1] += second(1, 1); P[2] += second(2, 2);
    P[3] += second(0, 1); P[4] += second(0, 2); P[5] += second(1, 2);
  }
  if (std::fabs(vol) < 1e-15) fail("mesh '" + ma.name + "' has zero volume; give the body an <inertial>");
  if (vol < 0) { vol = -vol; for (int k = 0; k < 3; k++) com[k] = -com[k]; for (double& x : P) x = -x; }
  for (int k = 0; k < 3; k++) com[k] /= vol;
  double xx = P[0] - vol * com[0] * com[0], yy = P[1] - vol * com[1] * com[1], zz = P[2] - vol * com[2] * com[2];
  double xy = P[3] - vol * com[0] * com[1], xz = P[4] - vol * com[0] * com[2], yz = P[5] - vol * com[1] * com[2];
  I[0] = yy + zz; I[4] = xx + zz; I[8] = xx + yy;
  I[1] = I[3] = -xy; I[2] = I[6] = -xz; I[5] = I[7] = -yz;
}

// ---------- geoms ----------
struct GeomTmp {
  int type, contype, conaffinity, condim, priority, dataid;
  double size[3], pos[3], quat[4], friction[3], solmix, solref[2], solimp[5], margin, gap, density, mass;
  bool has_mass;
  float rgba[4];
  std::string name;
};

int geom_type_from(const char* s) {
  if (!s) return mjGEOM_SPHERE;
  static const char* names[] = {"plane", "hfield", "sphere", "capsule", "ellipsoid", "cylinder", "box", "mesh"};
  for (int i = 0; i < 8; i++)
    if (!std::strcmp(s, names[i])) return i;
  fail(std::string("unknown geom type '") + s + "'");
}

// mass (for unit density: volume) and diagonal inertia in the geom frame
void geom_volume_inertia(const Ctx& c, const GeomTmp& g, double& vol, double* diag, double* com_local, double* Ifull) {
  const double r = g.size[0], h = g.size[1];
  zero3(com_local);
  for (int k = 0; k < 9; k++) Ifull[k] = 0;
  switch (g.type) {
    case mjGEOM_SPHERE:
      vol = 4.0 / 3.0 * mjPI * r * r * r;
      diag[0] = diag[1] = diag[2] = 0.4 * vol * r * r;
      break;
    case mjGEOM_CAPSULE: {
      double height = 2 * h;
      vol = mjPI * (r * r * height + 4.0 / 3.0 * r * r * r);
      double ms = vol * 4 * r / (4 * r + 3 * height), mc = vol - ms;
      diag[0] = diag[1] = mc * (3 * r * r + height * height) / 12.0;
      diag[2] = mc * r * r / 2.0;
      double si = 2 * ms * r * r / 5.0;
      diag[0] += si + ms * height * (3 * r + 2 * height) / 8.0;
      diag[1] = diag[0];
      diag[2] += si;
      break;
    }
    case mjGEOM_CYLINDER: {
      double height = 2 * h;
      vol = mjPI * r * r * height;
      diag[0] = diag[1] = vol * (3 * r * r + height * height) / 12.0;
      diag[2] = vol * r * r / 2.0;
      break;
    }
    case mjGEOM_BOX:
      vol = 8 * g.size[0] * g.size[1] * g.size[2];
      diag[0] = vol * (g.size[1] * g.size[1] + g.size[2] * g.size[2]) / 3.0;
      diag[1] = vol * (g.size[0] * g.size[0] + g.size[2] * g.size[2]) / 3.0;
      diag[2] = vol * (g.size[0] * g.size[0] + g.size[1] * g.size[1]) / 3.0;
      break;
    case mjGEOM_ELLIPSOID:
      vol = 4.0 / 3.0 * mjPI * g.size[0] * g.size[1] * g.size[2];
      diag[0] = vol * (g.size[1] * g.size[1] + g.size[2] * g.size[2]) / 5.0;
      diag[1] = vol * (g.size[0] * g.size[0] + g.size[2] * g.size[2]) / 5.0;
      diag[2] = vol * (g.size[0] * g.size[0] + g.size[1] * g.size[1]) / 5.0;
      break;
    case mjGEOM_MESH: {
      mesh_mass_props(c.meshes[g.dataid], vol, com_local, Ifull);
      diag[0] = diag[1] = diag[2] = -1;  // use Ifull
      return;
    }
    default:
      vol = 0;
      diag[0] = diag[1] = diag[2] = 0;
  }
  Ifull[0] = diag[0]; Ifull[4] = diag[1]; Ifull[8] = diag[2];
}

double geom_rbound(const Ctx& c, const GeomTmp& g) {
  switch (g.type) {
    case mjGEOM_SPHERE: return g.size[0];
    case mjGEOM_CAPSULE: return g.size[0] + g.size[1];
    case mjGEOM_CYLINDER: return std::sqrt(g.size[0] * g.size[0] + g.size[1] * g.size[1]);
    case mjGEOM_BOX: return norm3(g.size);
    case mjGEOM_ELLIPSOID: return std::max(g.size[0], std::max(g.size[1], g.size[2]));
    case mjGEOM_MESH: {
      double r2 = 0;
      const auto& v = c.meshes[g.dataid].vert;
      for (size_t i = 0; i + 2 < v.size(); i += 3) r2 = std::max(r2, v[i] * v[i] + v[i + 1] * v[i + 1] + v[i + 2] * v[i + 2]);
      return std::sqrt(r2);
    }
    default: return 0;  // plane / hfield: unbounded
  }
}

GeomTmp parse_geom(Ctx& c, const XmlElem& e, const std::string& cc) {
  GeomTmp g{};
  auto L = [&](const char* key) { return lookup(c, e, "geom", cc, key); };
  g.name = e.attr("name") ? e.attr("name") : "";
  const char* mesh = L("mesh");
  const char* type = L("type");
  g.type = type ? geom_type_from(type) : (mesh ? mjGEOM_MESH : mjGEOM_SPHERE);
  g.dataid = -1;
  if (g.type == mjGEOM_MESH) {
    if (!mesh) fail("mesh geom without mesh attribute");
    for (size_t i = 0; i < c.meshes.size(); i++)
      if (c.meshes[i].name == mesh) g.dataid = (int)i;
    if (g.dataid < 0) fail(std::string("unknown mesh '") + mesh + "'");
  }
  double v[8];
  g.size[0] = g.size[1] = g.size[2] = 0;
  if (const char* s = L("size")) {
    int n = parse_nums(s, v, 3);
    for (int k = 0; k < n; k++) g.size[k] = v[k];
  }
  zero3(g.pos);
  if (const char* s = L("pos")) { if (parse_nums(s, g.pos, 3) != 3) fail("geom pos needs 3 numbers"); }
  parse_orientation(c, e, "geom", cc, g.quat);
  if (g.type == mjGEOM_MESH) {   // the mesh was re-centred: the geom frame moves with it
    double off[3];
    rot_vec_quat(off, c.meshes[g.dataid].center, g.quat);
    for (int k = 0; k < 3; k++) g.pos[k] += off[k];
  }
  if (const char* s = L("fromto")) {
    if (parse_nums(s, v, 6) != 6) fail("fromto needs 6 numbers");
    double d[3] = {v[3] - v[0], v[4] - v[1], v[5] - v[2]};
    for (int k = 0; k < 3; k++) g.pos[k] = 0.5 * (v[k] + v[k + 3]);
    double len = normalize3(d);
    quat_z2vec(g.quat, d);
    if (g.type == mjGEOM_BOX || g.type == mjGEOM_ELLIPSOID) g.size[2] = 0.5 * len; else g.size[1] = 0.5 * len;
  }
  g.contype = 1; g.conaffinity = 1; g.condim = 3; g.priority = 0;
  if (const char* s = L("contype")) g.contype = std::atoi(s);
  if (const char* s = L("conaffinity")) g.conaffinity = std::atoi(s);
  if (const char* s = L("condim")) g.condim = std::atoi(s);
  if (const char* s = L("priority")) g.priority = std::atoi(s);
  if (g.condim != 1 && g.condim != 3 && g.condim != 4 && g.condim != 6) fail("condim must be 1, 3, 4 or 6");
  g.friction[0] = 1; g.friction[1] = 0.005; g.friction[2] = 0.0001;
  if (const char* s = L("friction")) { int n = parse_nums(s, v, 3); for (int k = 0; k < n; k++) g.friction[k] = v[k]; }
  g.solmix = 1; g.margin = 0; g.gap = 0;
  if (const char* s = L("solmix")) g.solmix = std::atof(s);
  if (const char* s = L("margin")) g.margin = std::atof(s);
  if (const char* s = L("gap")) g.gap = std::atof(s);
  g.solref[0] = 0.02; g.solref[1] = 1;
  if (const char* s = L("solref")) { int n = parse_nums(s, v, 2); for (int k = 0; k < n; k++) g.solref[k] = v[k]; }
  const double simp[5] = {0.9, 0.95, 0.001, 0.5, 2};
  for (int k = 0; k < 5; k++) g.solimp[k] = simp[k];
  if (const char* s = L("solimp")) { int n = parse_nums(s, v, 5); for (int k = 0; k < n; k++) g.solimp[k] = v[k]; }
  g.density = c.default_density;
  if (const char* s = L("density")) g.density = std::atof(s);
  g.has_mass = false;
  if (const char* s = L("mass")) { g.mass = std::atof(s); g.has_mass = true; }
  g.rgba[0] = g.rgba[1] = g.rgba[2] = 0.5f; g.rgba[3] = 1.0f;
  if (const char* s = L("rgba")) { int n = parse_nums(s, v, 4); for (int k = 0; k < n; k++) g.rgba[k] = (float)v[k]; }
  if (g.type == mjGEOM_MESH) {
    // bounding box half-sizes (used for visualisation only by the reference, mj_ros.cpp marker code)
    const auto& mv = c.meshes[g.dataid].vert;
    double mx[3] = {0, 0, 0};
    for (size_t i = 0; i + 2 < mv.size(); i += 3) for (int k = 0; k < 3; k++) mx[k] = std::max(mx[k], std::fabs(mv[i + k]));
    copy3(g.size, mx);
  }
  return g;
}

// ---------- bodies ----------
struct BodyBuild {
  bool has_inertial = false;
};

void add_name(std::vector<int>& adr, std::vector<char>& names, const std::string& n) {
  adr.push_back((int)names.size());
  names.insert(names.end(), n.begin(), n.end());
  names.push_back('\0');
}

void parse_body(Ctx& c, const XmlElem& e, int parent, std::string childclass, bool is_world) {
  ModelStore& S = c.S;
  int id;
  if (is_world) {
    id = 0;
  } else {
    id = (int)S.body_parentid.size();
    if (const char* cc = e.attr("childclass")) childclass = cc;
    S.body_parentid.push_back(parent);
    double pos[3] = {0, 0, 0}, quat[4];
    if (const char* s = e.attr("pos")) { if (parse_nums(s, pos, 3) != 3) fail("body pos needs 3 numbers"); }
    parse_orientation(c, e, "body", childclass, quat);
    S.body_pos.insert(S.body_pos.end(), pos, pos + 3);
    S.body_quat.insert(S.body_quat.end(), quat, quat + 4);
    S.body_gravcomp.push_back(e.attr("gravcomp") ? std::atof(e.attr("gravcomp")) : 0.0);
    bool mocap = e.attr("mocap") && parse_bool(e.attr("mocap"), "mocap");
    S.body_mocapid.push_back(mocap ? S.view.nmocap++ : -1);
    S.body_jntnum.push_back(0); S.body_jntadr.push_back(-1);
    S.body_dofnum.push_back(0); S.body_dofadr.push_back(-1);
    S.body_geomnum.push_back(0); S.body_geomadr.push_back(-1);
    S.body_mass.push_back(0);
    for (int k = 0; k < 3; k++) { S.body_ipos.push_back(0); S.body_inertia.push_back(0); }
    S.body_iquat.push_back(1); S.body_iquat.push_back(0); S.body_iquat.push_back(0); S.body_iquat.push_back(0);
    c.body_names.push_back(e.attr("name") ? e.attr("name") : "");
  }

  // joints (world body cannot have joints)
  for (auto& chp : e.children) {
    const XmlElem& ch = *chp;
    bool fj = ch.name == "freejoint";
    if (!fj && ch.name != "joint") continue;
    if (is_world) fail("joints cannot be attached to the world body");
    auto L = [&](const char* key) { return fj ? ch.attr(key) : lookup(c, ch, "joint", childclass, key); };
    int type = mjJNT_HINGE;
    if (fj) type = mjJNT_FREE;
    else if (const char* s = L("type")) {
      if (!std::strcmp(s, "free")) type = mjJNT_FREE;
      else if (!std::strcmp(s, "ball")) type = mjJNT_BALL;
      else if (!std::strcmp(s, "slide")) type = mjJNT_SLIDE;
      else if (!std::strcmp(s, "hinge")) type = mjJNT_HINGE;
      else fail(std::string("unknown joint type '") + s + "'");
    }
    int jid = (int)S.jnt_type.size();
    if (S.body_jntnum[id] == 0) { S.body_jntadr[id] = jid; S.body_dofadr[id] = (int)S.dof_bodyid.size(); }
    S.body_jntnum[id]++;
    S.jnt_type.push_back(type);
    S.jnt_bodyid.push_back(id);
    S.jnt_qposadr.push_back((int)S.qpos0.size());
    S.jnt_dofadr.push_back((int)S.dof_bodyid.size());
    c.jnt_names.push_back(ch.attr("name") ? ch.attr("name") : "");
    double v[8];
    double jpos[3] = {0, 0, 0}, axis[3] = {0, 0, 1};
    if (const char* s = L("pos")) { if (parse_nums(s, jpos, 3) != 3) fail("joint pos needs 3 numbers"); }
    if (const char* s = L("axis")) { if (parse_nums(s, axis, 3) != 3) fail("joint axis needs 3 numbers"); }
    if (type == mjJNT_FREE) { zero3(jpos); axis[0] = 0; axis[1] = 0; axis[2] = 1; }
    normalize3(axis);
    S.jnt_pos.insert(S.jnt_pos.end(), jpos, jpos + 3);
    S.jnt_axis.insert(S.jnt_axis.end(), axis, axis + 3);
    double range[2] = {0, 0};
    bool has_range = false;
    const bool angular = (type == mjJNT_HINGE || type == mjJNT_BALL);
    if (const char* s = L("range")) {
      if (parse_nums(s, range, 2) != 2) fail("range needs 2 numbers");
      has_range = true;
      if (angular && c.comp.degree) { range[0] *= mjPI / 180.0; range[1] *= mjPI / 180.0; }
    }
    bool limited = false;
    const char* lim = L("limited");
    if (lim && std::strcmp(lim, "auto")) limited = parse_bool(lim, "limited");
    else limited = c.comp.autolimits && has_range;
    if (type == mjJNT_FREE) limited = false;
    S.jnt_limited.push_back(limited ? 1 : 0);
    S.jnt_range.push_back(range[0]); S.jnt_range.push_back(range[1]);
    S.jnt_margin.push_back(L("margin") ? std::atof(L("margin")) : 0.0);
    S.jnt_stiffness.push_back(L("stiffness") ? std::atof(L("stiffness")) : 0.0);
    double sr[2] = {0.02, 1}, si[5] = {0.9, 0.95, 0.001, 0.5, 2};
    if (const char* s = L("solreflimit")) { int n = parse_nums(s, v, 2); for (int k = 0; k < n; k++) sr[k] = v[k]; }
    if (const char* s = L("solimplimit")) { int n = parse_nums(s, v, 5); for (int k = 0; k < n; k++) si[k] = v[k]; }
    S.jnt_solref.insert(S.jnt_solref.end(), sr, sr + 2);
    S.jnt_solimp.insert(S.jnt_solimp.end(), si, si + 5);
    double ref = L("ref") ? std::atof(L("ref")) : 0.0;
    double springref = L("springref") ? std::atof(L("springref")) : 0.0;
    if (angular && c.comp.degree) { ref *= mjPI / 180.0; springref *= mjPI / 180.0; }
    double damping = L("damping") ? std::atof(L("damping")) : 0.0;
    double armature = L("armature") ? std::atof(L("armature")) : 0.0;
    double frictionloss = L("frictionloss") ? std::atof(L("frictionloss")) : 0.0;
    double fsr[2] = {0.02, 1}, fsi[5] = {0.9, 0.95, 0.001, 0.5, 2};
    if (const char* s = L("solreffriction")) { int n = parse_nums(s, v, 2); for (int k = 0; k < n; k++) fsr[k] = v[k]; }
    if (const char* s = L("solimpfriction")) { int n = parse_nums(s, v, 5); for (int k = 0; k < n; k++) fsi[k] = v[k]; }
    int nq = 1, nd = 1;
    if (type == mjJNT_FREE) { nq = 7; nd = 6; }
    if (type == mjJNT_BALL) { nq = 4; nd = 3; }
    if (type == mjJNT_FREE) {
      // qpos0 of a free joint is the body's authored pose (body frame relative to the world)
      if (parent != 0) fail("free joint only allowed on a child of the world body");
      for (int k = 0; k < 3; k++) { S.qpos0.push_back(S.body_pos[3 * id + k]); S.qpos_spring.push_back(S.body_pos[3 * id + k]); }
      for (int k = 0; k < 4; k++) { S.qpos0.push_back(S.body_quat[4 * id + k]); S.qpos_spring.push_back(S.body_quat[4 * id + k]); }
    } else if (type == mjJNT_BALL) {
      const double q1[4] = {1, 0, 0, 0};
      S.qpos0.insert(S.qpos0.end(), q1, q1 + 4);
      S.qpos_spring.insert(S.qpos_spring.end(), q1, q1 + 4);
    } else {
      S.qpos0.push_back(ref);
      S.qpos_spring.push_back(springref);
    }
    for (int k = 0; k < nd; k++) {
      S.dof_bodyid.push_back(id);
      S.dof_jntid.push_back(jid);
      S.dof_damping.push_back(damping);
      S.dof_armature.push_back(armature);
      S.dof_frictionloss.push_back(frictionloss);
      S.dof_solref.insert(S.dof_solref.end(), fsr, fsr + 2);
      S.dof_solimp.insert(S.dof_solimp.end(), fsi, fsi + 5);
    }
    S.body_dofnum[id] += nd;
    (void)nq;
  }

  // geoms
  std::vector<GeomTmp> geoms;
  for (auto& chp : e.children)
    if (chp->name == "geom") geoms.push_back(parse_geom(c, *chp, childclass));
  if (!geoms.empty()) S.body_geomadr[id] = (int)S.geom_type.size();
  S.body_geomnum[id] = (int)geoms.size();
  for (auto& g : geoms) {
    S.geom_type.push_back(g.type); S.geom_contype.push_back(g.contype); S.geom_conaffinity.push_back(g.conaffinity);
    S.geom_condim.push_back(g.condim); S.geom_bodyid.push_back(id); S.geom_dataid.push_back(g.dataid);
    S.geom_priority.push_back(g.priority);
    S.geom_size.insert(S.geom_size.end(), g.size, g.size + 3);
    S.geom_rbound.push_back(geom_rbound(c, g));
    S.geom_pos.insert(S.geom_pos.end(), g.pos, g.pos + 3);
    S.geom_quat.insert(S.geom_quat.end(), g.quat, g.quat + 4);
    S.geom_friction.insert(S.geom_friction.end(), g.friction, g.friction + 3);
    S.geom_solmix.push_back(g.solmix);
    S.geom_solref.insert(S.geom_solref.end(), g.solref, g.solref + 2);
    S.geom_solimp.insert(S.geom_solimp.end(), g.solimp, g.solimp + 5);
    S.geom_margin.push_back(g.margin); S.geom_gap.push_back(g.gap);
    S.geom_rgba.insert(S.geom_rgba.end(), g.rgba, g.rgba + 4);
    c.geom_names.push_back(g.name);
  }

  // inertial: explicit, else inferred from geoms
  if (!is_world) {
    const XmlElem* in = e.child("inertial");
    bool use_geoms = (c.comp.inertiafromgeom == 1) || (c.comp.inertiafromgeom == 2 && !in);
    double mass = 0, ipos[3] = {0, 0, 0}, iquat[4] = {1, 0, 0, 0}, inertia[3] = {0, 0, 0};
    if (!use_geoms && in) {
      double v[6];
      if (const char* s = in->attr("pos")) parse_nums(s, ipos, 3);
      mass = in->attr("mass") ? std::atof(in->attr("mass")) : 0.0;
      parse_orientation(c, *in, "inertial", childclass, iquat);
      if (const char* s = in->attr("diaginertia")) {
        if (parse_nums(s, inertia, 3) != 3) fail("diaginertia needs 3 numbers");
      } else if (const char* s = in->attr("fullinertia")) {
        if (parse_nums(s, v, 6) != 6) fail("fullinertia needs 6 numbers");
        double A[9] = {v[0], v[3], v[4], v[3], v[1], v[5], v[4], v[5], v[2]}, V[9];
        eig3(A, inertia, V);
        mat2quat(iquat, V);
      }
    } else if (!geoms.empty()) {
      // accumulate in the body frame
      double com[3] = {0, 0, 0};
      struct Part { double m, p[3], I[9]; };
      std::vector<Part> parts;
      for (auto& g : geoms) {
        if (g.type == mjGEOM_PLANE || g.type == mjGEOM_HFIELD) continue;
        double vol, diag[3], cl[3], If[9];
        geom_volume_inertia(c, g, vol, diag, cl, If);
        double gm = g.has_mass ? g.mass : g.density * vol;
        double scale = vol > 0 ? gm / vol : 0;
        Part p;
        p.m = gm;
        double R[9], off[3];
        quat2mat(R, g.quat);
        mul_mat_vec3(off, R, cl);
        for (int k = 0; k < 3; k++) p.p[k] = g.pos[k] + off[k];
        // I_body = R * If * R^T * scale
        double T[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
          double s = 0;
          for (int k = 0; k < 3; k++) s += R[3 * i + k] * If[3 * k + j];
          T[3 * i + j] = s;
        }
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
          double s = 0;
          for (int k = 0; k < 3; k++) s += T[3 * i + k] * R[3 * j + k];
          p.I[3 * i + j] = s * scale;
        }
        parts.push_back(p);
        mass += gm;
        for (int k = 0; k < 3; k++) com[k] += gm * p.p[k];
      }
      if (mass > 0) {
        for (int k = 0; k < 3; k++) com[k] /= mass;
        double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (auto& p : parts) {
          double d[3] = {p.p[0] - com[0], p.p[1] - com[1], p.p[2] - com[2]};
          double d2 = dot3(d, d);
          for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
            A[3 * i + j] += p.I[3 * i + j] + p.m * ((i == j ? d2 : 0) - d[i] * d[j]);
        }
        double V[9];
        eig3(A, inertia, V);
        mat2quat(iquat, V);
        copy3(ipos, com);
      }
    }
    if (c.comp.boundmass > 0 && mass < c.comp.boundmass) mass = c.comp.boundmass;
    if (c.comp.boundinertia > 0) for (double& x : inertia) x = std::max(x, c.comp.boundinertia);
    if (c.comp.balanceinertia) {
      if (inertia[0] + inertia[1] < inertia[2] || inertia[0] + inertia[2] < inertia[1] || inertia[1] + inertia[2] < inertia[0]) {
        double mean = (inertia[0] + inertia[1] + inertia[2]) / 3;
        inertia[0] = inertia[1] = inertia[2] = mean;
      }
    }
    S.body_mass[id] = mass;
    for (int k = 0; k < 3; k++) { S.body_ipos[3 * id + k] = ipos[k]; S.body_inertia[3 * id + k] = inertia[k]; }
    for (int k = 0; k < 4; k++) S.body_iquat[4 * id + k] = iquat[k];
  }

  for (auto& chp : e.children)
    if (chp->name == "body") parse_body(c, *chp, id, childclass, false);
}

int find_name(const std::vector<std::string>& names, const char* n, int offset = 0) {
  if (!n) return -1;
  for (size_t i = 0; i < names.size(); i++)
    if (names[i] == n) return (int)i + offset;
  return -1;
}

// Static candidate geom pairs after every compile-time filter, in canonical order:
// sorted by (type1, type2, geom1, geom2) with type1 <= type2, so equal narrow-phase functions are adjacent.
void build_pairs(Ctx& c) {
  ModelStore& S = c.S;
  const int ng = (int)S.geom_type.size();
  std::set<int> excl(S.exclude_signature.begin(), S.exclude_signature.end());
  const bool filterparent = !(S.view.opt.disableflags & mjDSBL_FILTERPARENT);
  struct Pr { int t1, t2, g1, g2, ba = 0, bb = 0, ga = 0, gb = 0; };
  std::vector<Pr> prs;
  for (int a = 0; a < ng; a++)
    for (int b = a + 1; b < ng; b++) {
      int b1 = S.geom_bodyid[a], b2 = S.geom_bodyid[b];
      if (b1 == b2) continue;
      int w1 = S.body_weldid[b1], w2 = S.body_weldid[b2];
      if (w1 == w2) continue;  // includes static-static
      if (filterparent && w1 != 0 && w2 != 0) {
        int wp1 = S.body_weldid[S.body_parentid[w1]], wp2 = S.body_weldid[S.body_parentid[w2]];
        if (wp1 == w2 || wp2 == w1) continue;
      }
      if (!((S.geom_contype[a] & S.geom_conaffinity[b]) || (S.geom_contype[b] & S.geom_conaffinity[a]))) continue;
      int lo = std::min(b1, b2), hi = std::max(b1, b2);
      if (excl.count((lo << 16) | hi)) continue;
      int t1 = S.geom_type[a], t2 = S.geom_type[b];
      int g1 = a, g2 = b;
      if (t1 > t2) { std::swap(t1, t2); std::swap(g1, g2); }
      if (t1 == mjGEOM_PLANE && (t2 == mjGEOM_PLANE || t2 == mjGEOM_HFIELD)) continue;
      // sort key = MuJoCo's pair-generation order: body pair (lower body id first), then the geoms of the
      // first body in order, then the geoms of the second; inside a pair geom1 is the one with the lower geom TYPE
      Pr pr{t1, t2, g1, g2};
      pr.ba = lo; pr.bb = hi; pr.ga = b1 <= b2 ? a : b; pr.gb = b1 <= b2 ? b : a;
      prs.push_back(pr);
    }
  std::sort(prs.begin(), prs.end(), [](const Pr& x, const Pr& y) {
    if (x.ba != y.ba) return x.ba < y.ba;
    if (x.bb != y.bb) return x.bb < y.bb;
    if (x.ga != y.ga) return x.ga < y.ga;
    return x.gb < y.gb;
  });
  for (auto& p : prs) { S.pair_geom1.push_back(p.g1); S.pair_geom2.push_back(p.g2); }
  S.view.npair = (int)prs.size();
}

mjModel* compile_root(std::unique_ptr<XmlElem> root, const std::string& basedir, const std::string& src_text) {
  if (root->name != "mujoco") fail("root element must be <mujoco>");
  auto store = std::make_unique<ModelStore>();
  ModelStore& S = *store;
  S.source_xml = src_text;
  S.source_dir = basedir;
  Ctx c(S);
  c.basedir = basedir;
  resolve_includes(*root, basedir, 0);

  mjOption& o = S.view.opt;
  o.timestep = 0.002; o.impratio = 1; o.tolerance = 1e-8; o.noslip_tolerance = 1e-6;
  o.gravity[0] = 0; o.gravity[1] = 0; o.gravity[2] = -9.81;
  o.integrator = mjINT_EULER; o.cone = mjCONE_PYRAMIDAL; o.solver = mjSOL_NEWTON; o.iterations = 100;   // MuJoCo 2.3.7's defaults (the engine's substitution is reported by mj_loadXML)
  o.noslip_iterations = 0; o.disableflags = 0; o.enableflags = 0;
  int nconmax = -1, njmax = -1;

  // pass 1: compiler / option / size / default (all occurrences, later ones override)
  for (auto& chp : root->children) {
    const XmlElem& e = *chp;
    if (e.name == "compiler") {
      if (const char* s = e.attr("angle")) c.comp.degree = std::strcmp(s, "radian") != 0;
      if (const char* s = e.attr("eulerseq")) { c.comp.eulerseq = s; if (c.comp.eulerseq.size() != 3) fail("eulerseq needs 3 characters"); }
      if (const char* s = e.attr("meshdir")) c.comp.meshdir = s;
      if (const char* s = e.attr("autolimits")) c.comp.autolimits = parse_bool(s, "autolimits");
      if (const char* s = e.attr("boundmass")) c.comp.boundmass = std::atof(s);
      if (const char* s = e.attr("boundinertia")) c.comp.boundinertia = std::atof(s);
      if (const char* s = e.attr("balanceinertia")) c.comp.balanceinertia = parse_bool(s, "balanceinertia");
      if (const char* s = e.attr("inertiafromgeom")) c.comp.inertiafromgeom = !std::strcmp(s, "auto") ? 2 : parse_bool(s, "inertiafromgeom");
    } else if (e.name == "option") {
      double v[3];
      if (const char* s = e.attr("timestep")) o.timestep = std::atof(s);
      if (const char* s = e.attr("gravity")) { if (parse_nums(s, v, 3) != 3) fail("gravity needs 3 numbers"); copy3(o.gravity, v); }
      if (const char* s = e.attr("impratio")) o.impratio = std::atof(s);
      if (const char* s = e.attr("tolerance")) o.tolerance = std::atof(s);
      if (const char* s = e.attr("iterations")) o.iterations = std::atoi(s);
      if (const char* s = e.attr("noslip_iterations")) o.noslip_iterations = std::atoi(s);
      if (const char* s = e.attr("noslip_tolerance")) o.noslip_tolerance = std::atof(s);
      if (const char* s = e.attr("integrator")) {
        if (!std::strcmp(s, "Euler")) o.integrator = mjINT_EULER;
        else if (!std::strcmp(s, "RK4")) o.integrator = mjINT_RK4;  // step1/step2 run Euler regardless
        else if (!std::strcmp(s, "implicit")) o.integrator = mjINT_IMPLICIT;
        else fail(std::string("unknown integrator '") + s + "'");
      }
      if (const char* s = e.attr("cone")) {
        if (!std::strcmp(s, "elliptic")) fail("cone=\"elliptic\" is not supported by the batched PGS engine");
      }
      if (const char* s = e.attr("solver")) {
        o.solver = !std::strcmp(s, "PGS") ? mjSOL_PGS : !std::strcmp(s, "CG") ? mjSOL_CG : mjSOL_NEWTON;
      }
      if (const XmlElem* f = e.child("flag")) {
        struct { const char* n; int bit; } dis[] = {
            {"constraint", mjDSBL_CONSTRAINT}, {"equality", mjDSBL_EQUALITY}, {"frictionloss", mjDSBL_FRICTIONLOSS},
            {"limit", mjDSBL_LIMIT}, {"contact", mjDSBL_CONTACT}, {"passive", mjDSBL_PASSIVE}, {"gravity", mjDSBL_GRAVITY},
            {"warmstart", mjDSBL_WARMSTART}, {"filterparent", mjDSBL_FILTERPARENT}, {"refsafe", mjDSBL_REFSAFE},
            {"eulerdamp", mjDSBL_EULERDAMP}};
        for (auto& d : dis)
          if (const char* s = f->attr(d.n)) { if (!parse_bool(s, d.n)) o.disableflags |= d.bit; else o.disableflags &= ~d.bit; }
        if (const char* s = f->attr("energy")) { if (parse_bool(s, "energy")) o.enableflags |= mjENBL_ENERGY; }
      }
    } else if (e.name == "size") {
      if (const char* s = e.attr("nconmax")) nconmax = std::atoi(s);
      if (const char* s = e.attr("njmax")) njmax = std::atoi(s);
    } else if (e.name == "default") {
      parse_defaults(c, e, nullptr, true);
    }
  }

  // pass 2: assets
  for (auto& chp : root->children) {
    if (chp->name != "asset") continue;
    for (auto& ap : chp->children) {
      const XmlElem& a = *ap;
      if (a.name != "mesh") continue;
      MeshAsset ma;
      const char* file = a.attr("file");
      if (!file) fail("<mesh> without file");
      std::string fname = file;
      ma.name = a.attr("name") ? a.attr("name") : "";
      if (ma.name.empty()) {
        size_t sl = fname.find_last_of('/'), dot = fname.find_last_of('.');
        ma.name = fname.substr(sl == std::string::npos ? 0 : sl + 1, dot == std::string::npos ? std::string::npos : dot - (sl == std::string::npos ? 0 : sl + 1));
      }
      std::string path = join_path(join_path(basedir, c.comp.meshdir), fname);
      std::string ext = fname.size() > 4 ? fname.substr(fname.size() - 4) : "";
      for (char& ch : ext) ch = (char)std::tolower(ch);
      if (ext == ".stl") load_stl(path, ma);
      else if (ext == ".obj") load_obj(path, ma);
      else fail("unsupported mesh format '" + fname + "'");
      double sc[3] = {1, 1, 1};
      if (const char* s = a.attr("scale")) parse_nums(s, sc, 3);
      for (size_t i = 0; i + 2 < ma.vert.size(); i += 3) for (int k = 0; k < 3; k++) ma.vert[i + k] *= sc[k];
      if (sc[0] * sc[1] * sc[2] < 0)
        for (size_t f = 0; f + 2 < ma.face.size(); f += 3) std::swap(ma.face[f + 1], ma.face[f + 2]);
      recenter_mesh(ma);
      c.meshes.push_back(std::move(ma));
    }
  }

  // world body (id 0)
  S.body_parentid.push_back(0);
  const double z3[3] = {0, 0, 0}, q1[4] = {1, 0, 0, 0};
  S.body_pos.insert(S.body_pos.end(), z3, z3 + 3);
  S.body_quat.insert(S.body_quat.end(), q1, q1 + 4);
  S.body_ipos.insert(S.body_ipos.end(), z3, z3 + 3);
  S.body_iquat.insert(S.body_iquat.end(), q1, q1 + 4);
  S.body_inertia.insert(S.body_inertia.end(), z3, z3 + 3);
  S.body_mass.push_back(0); S.body_gravcomp.push_back(0); S.body_mocapid.push_back(-1);
  S.body_jntnum.push_back(0); S.body_jntadr.push_back(-1); S.body_dofnum.push_back(0); S.body_dofadr.push_back(-1);
  S.body_geomnum.push_back(0); S.body_geomadr.push_back(-1);
  c.body_names.push_back("world");

  // pass 3: every <worldbody> in document order. Geoms of the world body must stay contiguous, so world
  // geoms of all <worldbody> elements are gathered first, then child bodies.
  {
    XmlElem merged;
    merged.name = "worldbody";
    std::vector<const XmlElem*> wbs;
    for (auto& chp : root->children)
      if (chp->name == "worldbody") wbs.push_back(chp.get());
    std::vector<GeomTmp> wgeoms;
    for (auto* wb : wbs)
      for (auto& ch : wb->children)
        if (ch->name == "geom") wgeoms.push_back(parse_geom(c, *ch, ""));
    if (!wgeoms.empty()) S.body_geomadr[0] = 0;
    S.body_geomnum[0] = (int)wgeoms.size();
    for (auto& g : wgeoms) {
      S.geom_type.push_back(g.type); S.geom_contype.push_back(g.contype); S.geom_conaffinity.push_back(g.conaffinity);
      S.geom_condim.push_back(g.condim); S.geom_bodyid.push_back(0); S.geom_dataid.push_back(g.dataid);
      S.geom_priority.push_back(g.priority);
      S.geom_size.insert(S.geom_size.end(), g.size, g.size + 3);
      S.geom_rbound.push_back(geom_rbound(c, g));
      S.geom_pos.insert(S.geom_pos.end(), g.pos, g.pos + 3);
      S.geom_quat.insert(S.geom_quat.end(), g.quat, g.quat + 4);
      S.geom_friction.insert(S.geom_friction.end(), g.friction, g.friction + 3);
      S.geom_solmix.push_back(g.solmix);
      S.geom_solref.insert(S.geom_solref.end(), g.solref, g.solref + 2);
      S.geom_solimp.insert(S.geom_solimp.end(), g.solimp, g.solimp + 5);
      S.geom_margin.push_back(g.margin); S.geom_gap.push_back(g.gap);
      S.geom_rgba.insert(S.geom_rgba.end(), g.rgba, g.rgba + 4);
      c.geom_names.push_back(g.name);
    }
    for (auto* wb : wbs)
      for (auto& ch : wb->children)
        if (ch->name == "body") parse_body(c, *ch, 0, "", false);
  }

  // sizes and tree tables
  mjModel& v = S.view;
  v.nbody = (int)S.body_parentid.size();
  v.njnt = (int)S.jnt_type.size();
  v.nq = (int)S.qpos0.size();
  v.nv = (int)S.dof_bodyid.size();
  v.ngeom = (int)S.geom_type.size();
  v.nmesh = (int)c.meshes.size();
  S.body_rootid.assign(v.nbody, 0);
  S.body_weldid.assign(v.nbody, 0);
  for (int b = 1; b < v.nbody; b++) {
    int p = S.body_parentid[b];
    S.body_rootid[b] = p == 0 ? b : S.body_rootid[p];
    S.body_weldid[b] = S.body_jntnum[b] == 0 ? S.body_weldid[p] : b;
    if (S.body_mocapid[b] >= 0 && (p != 0 || S.body_jntnum[b] != 0)) fail("mocap body must be a joint-less child of the world");
  }
  // dof_parentid: previous dof within the body, else last dof of the nearest ancestor that has dofs
  S.dof_parentid.assign(v.nv, -1);
  S.dof_Madr.assign(v.nv, 0);
  for (int d = 0; d < v.nv; d++) {
    int b = S.dof_bodyid[d];
    if (d > S.body_dofadr[b]) { S.dof_parentid[d] = d - 1; continue; }
    int p = S.body_parentid[b];
    while (p > 0 && S.body_dofnum[p] == 0) p = S.body_parentid[p];
    S.dof_parentid[d] = p > 0 ? S.body_dofadr[p] + S.body_dofnum[p] - 1 : -1;
  }
  int nM = 0;
  for (int d = 0; d < v.nv; d++) {
    S.dof_Madr[d] = nM;
    for (int j = d; j >= 0; j = S.dof_parentid[j]) nM++;
  }
  v.nM = nM;

  // meshes into flat arrays
  for (auto& ma : c.meshes) {
    S.mesh_vertadr.push_back((int)S.mesh_vert.size() / 3);
    S.mesh_vertnum.push_back((int)ma.vert.size() / 3);
    S.mesh_vert.insert(S.mesh_vert.end(), ma.vert.begin(), ma.vert.end());
  }
  v.nmeshvert = (int)S.mesh_vert.size() / 3;

  // contact excludes and equality constraints
  for (auto& chp : root->children) {
    if (chp->name == "contact") {
      for (auto& ep : chp->children) {
        if (ep->name != "exclude") continue;
        int b1 = find_name(c.body_names, ep->attr("body1")), b2 = find_name(c.body_names, ep->attr("body2"));
        if (b1 < 0 || b2 < 0) fail(std::string("<exclude> refers to unknown body '") + (ep->attr(b1 < 0 ? "body1" : "body2") ? ep->attr(b1 < 0 ? "body1" : "body2") : "") + "'");
        S.exclude_signature.push_back((std::min(b1, b2) << 16) | std::max(b1, b2));
      }
    } else if (chp->name == "equality") {
      for (auto& ep : chp->children) {
        const XmlElem& q = *ep;
        double data[mjNEQDATA] = {0};
        int type, o1 = -1, o2 = -1;
        double vv[8];
        if (q.name == "joint") {
          type = mjEQ_JOINT;
          o1 = find_name(c.jnt_names, q.attr("joint1"));
          o2 = find_name(c.jnt_names, q.attr("joint2"));
          if (o1 < 0) fail("<equality><joint> joint1 not found");
          if (q.attr("joint2") && o2 < 0) fail("<equality><joint> joint2 not found");
          data[1] = 1;  // default polycoef "0 1 0 0 0"
          if (const char* s = q.attr("polycoef")) { for (int k = 0; k < 5; k++) data[k] = 0; int n = parse_nums(s, vv, 5); for (int k = 0; k < n; k++) data[k] = vv[k]; }
          for (int o : {o1, o2})
            if (o >= 0 && (S.jnt_type[o] == mjJNT_FREE || S.jnt_type[o] == mjJNT_BALL)) fail("joint equality needs scalar joints");
        } else if (q.name == "weld" || q.name == "connect") {
          type = q.name == "weld" ? mjEQ_WELD : mjEQ_CONNECT;
          o1 = find_name(c.body_names, q.attr("body1"));
          o2 = q.attr("body2") ? find_name(c.body_names, q.attr("body2")) : 0;
          if (o1 < 0 || o2 < 0) fail("<equality> body not found");
          if (const char* s = q.attr("anchor")) parse_nums(s, data, 3);
          if (type == mjEQ_WELD) {
            data[6] = 1;  // relpose quat filled in set_const from qpos0
            data[10] = q.attr("torquescale") ? std::atof(q.attr("torquescale")) : 1.0;
            if (const char* s = q.attr("relpose")) parse_nums(s, data + 3, 7);
            else data[3] = NAN;  // marker: compute from qpos0
          }
        } else {
          fail("unsupported equality type <" + q.name + ">");
        }
        S.eq_type.push_back(type); S.eq_obj1id.push_back(o1); S.eq_obj2id.push_back(o2);
        S.eq_active.push_back(q.attr("active") ? (parse_bool(q.attr("active"), "active") ? 1 : 0) : 1);
        double sr[2] = {0.02, 1}, si[5] = {0.9, 0.95, 0.001, 0.5, 2};
        if (const char* s = q.attr("solref")) { int n = parse_nums(s, vv, 2); for (int k = 0; k < n; k++) sr[k] = vv[k]; }
        if (const char* s = q.attr("solimp")) { int n = parse_nums(s, vv, 5); for (int k = 0; k < n; k++) si[k] = vv[k]; }
        S.eq_solref.insert(S.eq_solref.end(), sr, sr + 2);
        S.eq_solimp.insert(S.eq_solimp.end(), si, si + 5);
        S.eq_data.insert(S.eq_data.end(), data, data + mjNEQDATA);
      }
    }
  }
  v.neq = (int)S.eq_type.size();
  v.nexclude = (int)S.exclude_signature.size();

  // names
  for (auto& n : c.body_names) add_name(S.name_bodyadr, S.names, n);
  for (auto& n : c.jnt_names) add_name(S.name_jntadr, S.names, n);
  for (auto& n : c.geom_names) add_name(S.name_geomadr, S.names, n);
  for (auto& ma : c.meshes) add_name(S.name_meshadr, S.names, ma.name);
  v.nnames = (int)S.names.size();

  S.body_subtreemass.assign(v.nbody, 0);
  S.body_invweight0.assign(2 * (size_t)v.nbody, 0);
  S.dof_invweight0.assign(v.nv, 0);

  build_pairs(c);

  // caps: every candidate pair may emit up to its type's maximum; bounded for memory
  if (nconmax < 0) {
    long est = 0;
    for (int p = 0; p < v.npair; p++) {
      int t1 = S.geom_type[S.pair_geom1[p]], t2 = S.geom_type[S.pair_geom2[p]];
      int mx = 1;
      if (t1 == mjGEOM_PLANE && (t2 == mjGEOM_BOX || t2 == mjGEOM_CYLINDER || t2 == mjGEOM_MESH)) mx = 4;
      else if (t1 == mjGEOM_PLANE && t2 == mjGEOM_CAPSULE) mx = 2;
      else if (t1 == mjGEOM_CAPSULE && (t2 == mjGEOM_CAPSULE || t2 == mjGEOM_BOX)) mx = 2;
      else if (t1 == mjGEOM_BOX && t2 == mjGEOM_BOX) mx = 8;
      est += mx;
    }
    nconmax = (int)std::min<long>(est, 128);
  }
  int nlim = 0, nfl = 0, neqrow = 0;
  for (int j = 0; j < v.njnt; j++) nlim += S.jnt_limited[j] ? 1 : 0;
  for (int d = 0; d < v.nv; d++) nfl += S.dof_frictionloss[d] > 0 ? 1 : 0;
  for (int q = 0; q < v.neq; q++) neqrow += S.eq_type[q] == mjEQ_JOINT ? 1 : S.eq_type[q] == mjEQ_CONNECT ? 3 : 6;
  if (njmax < 0) {
    int maxdim = 1;
    for (int g = 0; g < v.ngeom; g++) maxdim = std::max(maxdim, S.geom_condim[g]);
    int rows_per = maxdim == 1 ? 1 : 2 * (maxdim - 1);
    njmax = neqrow + nfl + nlim + std::min(nconmax * rows_per, 4 * nconmax + 64);
    njmax = std::min(njmax, 384);
  }
  v.nconmax = nconmax;
  v.njmax = njmax;

  S.finalize();
  set_const(S);
  return &store.release()->view;
}

}  // namespace

// A <robot> root is URDF (the reference's importer hands URDF files to mj_loadXML, src/mujoco_compile.cpp:404): it is
// translated to MJCF first, and the translation becomes the model's source text (what mj_saveLastXML writes).
mjModel* compile_text(const std::string& text, const std::string& basedir) {
  auto root = XmlParser(text).parse();
  if (root->name == "robot") {
    const std::string mjcf = urdf_to_mjcf(*root);
    auto r2 = XmlParser(mjcf).parse();
    return compile_root(std::move(r2), basedir, mjcf);
  }
  return compile_root(std::move(root), basedir, text);
}

mjModel* compile_mjcf_string(const std::string& xml, const std::string& basedir) { return compile_text(xml, basedir); }

mjModel* compile_mjcf_file(const std::string& path) { return compile_text(read_file(path), dir_of(path)); }

}  // namespace b2
