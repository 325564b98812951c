// shim_host.cpp — host-only part of the MuJoCo-named C API (include/mujoco/mujoco.h): model/data lifecycle,
// name lookup, printing, mju_* utilities. The stepping entry points live in shim_step.cpp (GPU-backed).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>

#include "hostmath.h"
#include "model_store.h"

namespace b2 {
mjModel* compile_mjcf_string(const std::string& xml, const std::string& basedir);
mjModel* compile_mjcf_file(const std::string& path);
void shim_forget(const mjModel* m, const mjData* d);  // shim_step.cpp: drop GPU batches bound to these objects
}  // namespace b2

namespace {
std::mutex g_last_mtx;
std::string g_last_xml;  // text of the most recently loaded model, for mj_saveLastXML

void set_err(char* error, int error_sz, const std::string& msg) {
  if (error && error_sz > 0) {
    std::snprintf(error, (size_t)error_sz, "%s", msg.c_str());
  }
}
}  // namespace

extern "C" {

mjfGeneric mjcb_control = nullptr;

namespace b2 {
// What this engine does differently from the options a model asks for, said once per load instead of silently:
// m->opt keeps the authored values (MuJoCo's default solver is Newton), the constraint solve is always PGS with
// opt.iterations sweeps, and the noslip post-pass and RK4 are not run (BASELINE.json configs name PGS).
// B2_QUIET=1 silences the notes.
void report_substitutions(const mjModel* m) {
  if (!m || std::getenv("B2_QUIET")) return;
  static bool said[3] = {false, false, false};   // once per process and kind
  if (m->opt.solver != mjSOL_PGS && !said[0] && (said[0] = true))
    std::fprintf(stderr, "b2: note: solver %s requested (or defaulted); this engine solves constraints with PGS (%d iterations, tolerance %g)\n",
                 m->opt.solver == mjSOL_CG ? "CG" : "Newton", m->opt.iterations, m->opt.tolerance);
  if (m->opt.noslip_iterations > 0 && !said[1] && (said[1] = true))
    std::fprintf(stderr, "b2: note: noslip_iterations = %d requested; the noslip post-pass is not run\n", m->opt.noslip_iterations);
  if (m->opt.integrator != mjINT_EULER && !said[2] && (said[2] = true))
    std::fprintf(stderr, "b2: note: integrator %d requested; mj_step1 / mj_step2 integrate with semi-implicit Euler (as MuJoCo's split step does)\n", m->opt.integrator);
}
}  // namespace b2

// Replaces libmujoco's mj_loadXML as called from include/mujoco_sim/mj_util.h:190 and
// src/mujoco_compile.cpp:404. Error contract kept: NULL + message in `error`.
mjModel* mj_loadXML(const char* filename, const mjVFS*, char* error, int error_sz) {
  if (error && error_sz > 0) error[0] = 0;
  try {
    mjModel* m = b2::compile_mjcf_file(filename ? filename : "");
    b2::report_substitutions(m);
    std::lock_guard<std::mutex> lk(g_last_mtx);
    g_last_xml = static_cast<b2::ModelStore*>(m->owner_)->source_xml;
    return m;
  } catch (const std::exception& e) {
    set_err(error, error_sz, e.what());
    return nullptr;
  }
}

// MuJoCo's binary model I/O (mj_saveModel / mj_loadModel): a compiled model travels without its MJCF and mesh files
// (the C4 workload ships as such an image of the reference's pr2.xml, its meshes reduced to their hull vertices).
void mj_saveModel(const mjModel* m, const char* filename, void* buffer, int buffer_sz) {
  (void)buffer; (void)buffer_sz;
  if (!m || !m->owner_ || !filename) return;
  try { static_cast<const b2::ModelStore*>(m->owner_)->save(filename); }
  catch (const std::exception& e) { mju_warning(e.what()); }
}
mjModel* mj_loadModel(const char* filename, const mjVFS*) {
  try {
    b2::ModelStore* s = b2::ModelStore::load(filename ? filename : "");
    return &s->view;
  } catch (const std::exception& e) {
    mju_warning(e.what());
    return nullptr;
  }
}

mjModel* mj_loadXMLString(const char* xml, const char* basedir, char* error, int error_sz) {
  if (error && error_sz > 0) error[0] = 0;
  try {
    mjModel* m = b2::compile_mjcf_string(xml ? xml : "", basedir ? basedir : ".");
    b2::report_substitutions(m);
    std::lock_guard<std::mutex> lk(g_last_mtx);
    g_last_xml = static_cast<b2::ModelStore*>(m->owner_)->source_xml;
    return m;
  } catch (const std::exception& e) {
    set_err(error, error_sz, e.what());
    return nullptr;
  }
}

// include/mujoco_sim/mj_util.h:207. Writes back the (include-free) text of the model that was loaded last,
// or of `m` when given. Returns 1 on success like MuJoCo.
int mj_saveLastXML(const char* filename, const mjModel* m, char* error, int error_sz) {
  std::string text;
  if (m && m->owner_) text = static_cast<b2::ModelStore*>(m->owner_)->source_xml;
  else { std::lock_guard<std::mutex> lk(g_last_mtx); text = g_last_xml; }
  if (text.empty()) { set_err(error, error_sz, "no model has been loaded"); return 0; }
  std::ofstream f(filename ? filename : "");
  if (!f) { set_err(error, error_sz, std::string("cannot write '") + (filename ? filename : "") + "'"); return 0; }
  f << text;
  return 1;
}

mjData* mj_makeData(const mjModel* m) { return m ? b2::make_data(m) : nullptr; }

void mj_resetData(const mjModel* m, mjData* d) { if (m && d) b2::reset_data(m, d); }

void mj_deleteData(mjData* d) {
  if (!d) return;
  b2::shim_forget(nullptr, d);
  delete static_cast<b2::DataStore*>(d->owner_);
}

void mj_deleteModel(mjModel* m) {
  if (!m) return;
  b2::shim_forget(m, nullptr);
  delete static_cast<b2::ModelStore*>(m->owner_);
}

static bool name_table(const mjModel* m, int type, const int** adr, int* n) {
  switch (type) {
    case mjOBJ_BODY: case mjOBJ_XBODY: *adr = m->name_bodyadr; *n = m->nbody; return true;
    case mjOBJ_JOINT: *adr = m->name_jntadr; *n = m->njnt; return true;
    case mjOBJ_GEOM: *adr = m->name_geomadr; *n = m->ngeom; return true;
    case mjOBJ_MESH: *adr = m->name_meshadr; *n = m->nmesh; return true;
    default: *adr = nullptr; *n = 0; return false;
  }
}

// -1 when absent (contract relied on by src/mujoco_sim/mj_sim.cpp:1083-1146).
int mj_name2id(const mjModel* m, int type, const char* name) {
  const int* adr; int n;
  if (!m || !name || !name_table(m, type, &adr, &n)) return -1;
  for (int i = 0; i < n; i++)
    if (!std::strcmp(m->names + adr[i], name)) return i;
  return -1;
}

// NULL for an invalid id or an unnamed object (loop terminator in src/mujoco_sim/mj_sim.cpp:473-477).
const char* mj_id2name(const mjModel* m, int type, int id) {
  const int* adr; int n;
  if (!m || !name_table(m, type, &adr, &n) || id < 0 || id >= n) return nullptr;
  const char* s = m->names + adr[id];
  return s[0] ? s : nullptr;
}

void mj_printModel(const mjModel* m, const char* filename) {
  FILE* f = std::fopen(filename, "w");
  if (!f) return;
  std::fprintf(f, "B200 batched engine model (MuJoCo 2.3.7 field names)\n");
  std::fprintf(f, "nq %d\nnv %d\nnbody %d\njnt %d\nngeom %d\nnmesh %d\nneq %d\nnM %d\nnpair %d\nnconmax %d\nnjmax %d\n",
               m->nq, m->nv, m->nbody, m->njnt, m->ngeom, m->nmesh, m->neq, m->nM, m->npair, m->nconmax, m->njmax);
  std::fprintf(f, "timestep %.9g\ngravity %.9g %.9g %.9g\niterations %d\ntolerance %.3g\nmeaninertia %.9g\n", m->opt.timestep,
               m->opt.gravity[0], m->opt.gravity[1], m->opt.gravity[2], m->opt.iterations, m->opt.tolerance, m->stat.meaninertia);
  for (int b = 0; b < m->nbody; b++) {
    const char* nm = mj_id2name(m, mjOBJ_BODY, b);
    std::fprintf(f, "BODY %d: name %s parent %d root %d weld %d mass %.9g pos %.9g %.9g %.9g inertia %.9g %.9g %.9g invweight0 %.6g %.6g\n", b,
                 nm ? nm : "", m->body_parentid[b], m->body_rootid[b], m->body_weldid[b], m->body_mass[b], m->body_pos[3 * b],
                 m->body_pos[3 * b + 1], m->body_pos[3 * b + 2], m->body_inertia[3 * b], m->body_inertia[3 * b + 1],
                 m->body_inertia[3 * b + 2], m->body_invweight0[2 * b], m->body_invweight0[2 * b + 1]);
  }
  for (int j = 0; j < m->njnt; j++) {
    const char* nm = mj_id2name(m, mjOBJ_JOINT, j);
    std::fprintf(f, "JOINT %d: name %s type %d body %d qposadr %d dofadr %d limited %d range %.9g %.9g\n", j, nm ? nm : "",
                 m->jnt_type[j], m->jnt_bodyid[j], m->jnt_qposadr[j], m->jnt_dofadr[j], (int)m->jnt_limited[j], m->jnt_range[2 * j],
                 m->jnt_range[2 * j + 1]);
  }
  for (int g = 0; g < m->ngeom; g++)
    std::fprintf(f, "GEOM %d: type %d body %d size %.9g %.9g %.9g condim %d\n", g, m->geom_type[g], m->geom_bodyid[g],
                 m->geom_size[3 * g], m->geom_size[3 * g + 1], m->geom_size[3 * g + 2], m->geom_condim[g]);
  std::fclose(f);
}

void mj_printData(const mjModel* m, mjData* d, const char* filename) {
  FILE* f = std::fopen(filename, "w");
  if (!f) return;
  auto dump = [&](const char* name, const mjtNum* a, int n) {
    std::fprintf(f, "%s", name);
    for (int i = 0; i < n; i++) std::fprintf(f, " %.12g", a[i]);
    std::fprintf(f, "\n");
  };
  std::fprintf(f, "time %.12g\nncon %d\nnefc %d\n", d->time, d->ncon, d->nefc);
  dump("qpos", d->qpos, m->nq); dump("qvel", d->qvel, m->nv); dump("qacc", d->qacc, m->nv);
  dump("qacc_warmstart", d->qacc_warmstart, m->nv); dump("qfrc_applied", d->qfrc_applied, m->nv);
  dump("qfrc_bias", d->qfrc_bias, m->nv); dump("qfrc_inverse", d->qfrc_inverse, m->nv);
  dump("xpos", d->xpos, 3 * m->nbody); dump("xquat", d->xquat, 4 * m->nbody);
  std::fclose(f);
}

void* mju_malloc(unsigned long size) { return std::malloc(size ? size : 1); }
void mju_free(void* ptr) { std::free(ptr); }
void mju_zero(mjtNum* res, int n) { if (n > 0) std::memset(res, 0, sizeof(mjtNum) * (size_t)n); }
void mju_copy(mjtNum* res, const mjtNum* data, int n) { if (n > 0) std::memcpy(res, data, sizeof(mjtNum) * (size_t)n); }
void mju_addTo3(mjtNum res[3], const mjtNum vec[3]) { res[0] += vec[0]; res[1] += vec[1]; res[2] += vec[2]; }
void mju_mulQuat(mjtNum res[4], const mjtNum a[4], const mjtNum b[4]) { b2::hm::mul_quat(res, a, b); }
void mju_rotVecQuat(mjtNum res[3], const mjtNum vec[3], const mjtNum quat[4]) { b2::hm::rot_vec_quat(res, vec, quat); }
void mju_mat2Quat(mjtNum quat[4], const mjtNum mat[9]) { b2::hm::mat2quat(quat, mat); }
void mju_quat2Mat(mjtNum mat[9], const mjtNum quat[4]) { b2::hm::quat2mat(mat, quat); }
void mju_warning(const char* msg, ...) {
  va_list ap;
  va_start(ap, msg);
  std::fprintf(stderr, "WARNING: ");
  std::vfprintf(stderr, msg, ap);
  std::fprintf(stderr, "\n");
  va_end(ap);
}
void mju_error(const char* msg, ...) {
  va_list ap;
  va_start(ap, msg);
  std::fprintf(stderr, "ERROR: ");
  std::vfprintf(stderr, msg, ap);
  std::fprintf(stderr, "\n");
  va_end(ap);
  std::abort();
}
mjtNum mju_abs(mjtNum x) { return std::fabs(x); }
mjtNum mju_sin(mjtNum x) { return std::sin(x); }
mjtNum mju_cos(mjtNum x) { return std::cos(x); }
mjtNum mju_sqrt(mjtNum x) { return std::sqrt(x); }
mjtNum mju_ceil(mjtNum x) { return std::ceil(x); }

/* string-keyed access used by the Python host mirror (mujoco_sim_b200/engine.py) and the tests */
int b2_model_int(const mjModel* m, const char* name, int* out) { return b2::model_int(m, name, out); }
int b2_model_array(const mjModel* m, const char* name, const void** ptr, int* kind) { return b2::model_array(m, name, ptr, kind); }
int b2_data_array(const mjModel* m, mjData* d, const char* name, void** ptr, int* kind) { return b2::data_array(m, d, name, ptr, kind); }
int b2_model_set_opt(mjModel* m, const char* name, double value) {
  if (!std::strcmp(name, "timestep")) m->opt.timestep = value;
  else if (!std::strcmp(name, "iterations")) m->opt.iterations = (int)value;
  else if (!std::strcmp(name, "tolerance")) m->opt.tolerance = value;
  else if (!std::strcmp(name, "disableflags")) m->opt.disableflags = (int)value;
  else if (!std::strcmp(name, "gravity_x")) m->opt.gravity[0] = value;
  else if (!std::strcmp(name, "gravity_y")) m->opt.gravity[1] = value;
  else if (!std::strcmp(name, "gravity_z")) m->opt.gravity[2] = value;
  else return -1;
  return 0;
}
/* bulk view of the contact list of an mjData: geom ids and distances of the first min(n, ncon) contacts; returns ncon */
int b2_data_contacts(const mjData* d, int n, int* geom1, int* geom2, double* dist) {
  for (int i = 0; i < n && i < d->ncon; i++) {
    if (geom1) geom1[i] = d->contact[i].geom1;
    if (geom2) geom2[i] = d->contact[i].geom2;
    if (dist) dist[i] = d->contact[i].dist;
  }
  return d->ncon;
}
int b2_data_contact(const mjData* d, int i, mjContact* out) {
  if (i < 0 || i >= d->ncon) return -1;
  *out = d->contact[i];
  return 0;
}

}  // extern "C"
