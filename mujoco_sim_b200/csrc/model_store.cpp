// model_store.cpp — storage, mjData allocation and string-keyed field access.
#include "model_store.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace b2 {

void ModelStore::finalize() {
  mjModel& v = view;
  v.owner_ = this;
#define P(name) v.name = name.empty() ? nullptr : name.data()
  P(qpos0); P(qpos_spring);
  P(body_parentid); P(body_rootid); P(body_weldid); P(body_mocapid); P(body_jntnum); P(body_jntadr);
  P(body_dofnum); P(body_dofadr); P(body_geomnum); P(body_geomadr);
  P(body_pos); P(body_quat); P(body_ipos); P(body_iquat); P(body_mass); P(body_subtreemass); P(body_inertia);
  P(body_invweight0); P(body_gravcomp);
  P(jnt_type); P(jnt_bodyid); P(jnt_limited); P(jnt_solref); P(jnt_solimp); P(jnt_pos); P(jnt_axis);
  P(jnt_stiffness); P(jnt_range); P(jnt_margin);
  P(dof_bodyid); P(dof_jntid); P(dof_parentid); P(dof_Madr); P(dof_solref); P(dof_solimp); P(dof_frictionloss);
  P(dof_armature); P(dof_damping); P(dof_invweight0);
  P(geom_type); P(geom_contype); P(geom_conaffinity); P(geom_condim); P(geom_bodyid); P(geom_dataid);
  P(geom_priority); P(geom_size); P(geom_rbound); P(geom_pos); P(geom_quat); P(geom_friction); P(geom_solmix);
  P(geom_solref); P(geom_solimp); P(geom_margin); P(geom_gap); P(geom_rgba);
  P(mesh_vertadr); P(mesh_vertnum); P(mesh_vert);
  P(eq_type); P(eq_obj1id); P(eq_obj2id); P(eq_active); P(eq_solref); P(eq_solimp); P(eq_data);
  P(pair_geom1); P(pair_geom2);
  P(sensor_type); P(sensor_objid); P(sensor_adr);
  P(name_bodyadr); P(name_jntadr); P(name_geomadr); P(name_meshadr); P(names);
#undef P
  jnt_qposadr_padded.assign(jnt_qposadr.size() + 1, 0);
  jnt_dofadr_padded.assign(jnt_dofadr.size() + 1, 0);
  for (size_t i = 0; i < jnt_qposadr.size(); i++) {
    jnt_qposadr_padded[i + 1] = jnt_qposadr[i];
    jnt_dofadr_padded[i + 1] = jnt_dofadr[i];
  }
  v.jnt_qposadr = jnt_qposadr_padded.data() + 1;
  v.jnt_dofadr = jnt_dofadr_padded.data() + 1;
}

// ---- binary model image ------------------------------------------------------------------------------------------
// [magic "B2MJB001"][sizeof(mjModel)][mjModel bytes (pointers are rebuilt by finalize)] then, in one fixed order, every
// vector as [element size][count][bytes], then the source text and directory.
#define B2_MODEL_VECTORS(X)                                                                                            \
  X(qpos0) X(qpos_spring) X(body_parentid) X(body_rootid) X(body_weldid) X(body_mocapid) X(body_jntnum) X(body_jntadr) \
  X(body_dofnum) X(body_dofadr) X(body_geomnum) X(body_geomadr) X(body_pos) X(body_quat) X(body_ipos) X(body_iquat)    \
  X(body_mass) X(body_subtreemass) X(body_inertia) X(body_invweight0) X(body_gravcomp) X(jnt_type) X(jnt_qposadr)      \
  X(jnt_dofadr) X(jnt_bodyid) X(jnt_limited) X(jnt_solref) X(jnt_solimp) X(jnt_pos) X(jnt_axis) X(jnt_stiffness)       \
  X(jnt_range) X(jnt_margin) X(dof_bodyid) X(dof_jntid) X(dof_parentid) X(dof_Madr) X(dof_solref) X(dof_solimp)        \
  X(dof_frictionloss) X(dof_armature) X(dof_damping) X(dof_invweight0) X(geom_type) X(geom_contype)                    \
  X(geom_conaffinity) X(geom_condim) X(geom_bodyid) X(geom_dataid) X(geom_priority) X(geom_size) X(geom_rbound)        \
  X(geom_pos) X(geom_quat) X(geom_friction) X(geom_solmix) X(geom_solref) X(geom_solimp) X(geom_margin) X(geom_gap)    \
  X(geom_rgba) X(mesh_vertadr) X(mesh_vertnum) X(mesh_vert) X(eq_type) X(eq_obj1id) X(eq_obj2id) X(eq_active)          \
  X(eq_solref) X(eq_solimp) X(eq_data) X(pair_geom1) X(pair_geom2) X(sensor_type) X(sensor_objid) X(sensor_adr)        \
  X(name_bodyadr) X(name_jntadr) X(name_geomadr) X(name_meshadr) X(names) X(exclude_signature)

static const char kMagic[8] = {'B', '2', 'M', 'J', 'B', '0', '0', '1'};

void ModelStore::save(const std::string& path) const {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write '" + path + "'");
  auto W = [&](const void* p, size_t n) { if (n && std::fwrite(p, 1, n, f) != n) { std::fclose(f); throw std::runtime_error("short write to '" + path + "'"); } };
  W(kMagic, 8);
  const uint64_t vs = sizeof(mjModel);
  W(&vs, 8);
  W(&view, sizeof(mjModel));
#define X(v) { const uint64_t es = sizeof(v[0]), n = v.size(); W(&es, 8); W(&n, 8); W(v.data(), (size_t)(es * n)); }
  B2_MODEL_VECTORS(X)
#undef X
  // the MJCF source text travels only on request (B2_MJB_WITH_SOURCE=1): an image is a compiled artefact
  const std::string none;
  const bool with_src = std::getenv("B2_MJB_WITH_SOURCE") != nullptr;
  for (const std::string* t : {with_src ? &source_xml : &none, with_src ? &source_dir : &none}) { const uint64_t n = t->size(); W(&n, 8); W(t->data(), (size_t)n); }
  std::fclose(f);
}

ModelStore* ModelStore::load(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open '" + path + "'");
  auto* s = new ModelStore();
  auto bad = [&](const char* why) { std::fclose(f); delete s; throw std::runtime_error("'" + path + "': " + why); };
  auto R = [&](void* p, size_t n) { if (n && std::fread(p, 1, n, f) != n) bad("truncated model image"); };
  char magic[8];
  uint64_t vs = 0;
  R(magic, 8);
  if (std::memcmp(magic, kMagic, 8)) bad("not a model image of this library (magic)");
  R(&vs, 8);
  if (vs != sizeof(mjModel)) bad("model image of a different library version (mjModel size)");
  R(&s->view, sizeof(mjModel));
#define X(v) { uint64_t es = 0, n = 0; R(&es, 8); R(&n, 8); if (es != sizeof(s->v[0]) || n > (1ull << 32)) bad("corrupt array header"); s->v.resize((size_t)n); R(s->v.data(), (size_t)(es * n)); }
  B2_MODEL_VECTORS(X)
#undef X
  for (std::string* t : {&s->source_xml, &s->source_dir}) { uint64_t n = 0; R(&n, 8); if (n > (1ull << 32)) bad("corrupt text header"); t->resize((size_t)n); R(&(*t)[0], (size_t)n); }
  std::fclose(f);
  s->finalize();
  return s;
}

namespace {
struct Carver {
  std::vector<std::pair<double**, size_t>> items;
  size_t total = 0;
  void add(double** p, size_t n) { items.emplace_back(p, n); total += n; }
  void carve(std::vector<double>& buf) {
    buf.assign(total + 1, 0.0);
    size_t off = 0;
    for (auto& it : items) { *it.first = buf.data() + off; off += it.second; }
  }
};
}  // namespace

mjData* make_data(const mjModel* m) {
  auto* s = new DataStore();
  mjData& d = s->view;
  d.owner_ = s;
  const size_t nq = m->nq, nv = m->nv, nb = m->nbody, nj = m->njnt, ng = m->ngeom, nM = m->nM;
  const size_t njmax = m->njmax, ncm = m->nconmax;
  Carver c;
  c.add(&d.qpos, nq); c.add(&d.qvel, nv); c.add(&d.qacc, nv); c.add(&d.qacc_warmstart, nv);
  c.add(&d.qfrc_applied, nv); c.add(&d.xfrc_applied, 6 * nb);
  c.add(&d.mocap_pos, 3 * (size_t)m->nmocap); c.add(&d.mocap_quat, 4 * (size_t)m->nmocap);
  c.add(&d.sensordata, (size_t)m->nsensordata);
  c.add(&d.xpos, 3 * nb); c.add(&d.xquat, 4 * nb); c.add(&d.xmat, 9 * nb); c.add(&d.xipos, 3 * nb);
  c.add(&d.ximat, 9 * nb); c.add(&d.xanchor, 3 * nj); c.add(&d.xaxis, 3 * nj);
  c.add(&d.geom_xpos, 3 * ng); c.add(&d.geom_xmat, 9 * ng);
  c.add(&d.subtree_com, 3 * nb); c.add(&d.cinert, 10 * nb); c.add(&d.crb, 10 * nb); c.add(&d.cdof, 6 * nv);
  c.add(&d.qM, nM); c.add(&d.qLD, nM); c.add(&d.qLDiagInv, nv);
  c.add(&d.cvel, 6 * nb); c.add(&d.cdof_dot, 6 * nv); c.add(&d.qfrc_bias, nv); c.add(&d.qfrc_passive, nv);
  c.add(&d.cacc, 6 * nb); c.add(&d.cfrc_int, 6 * nb);
  c.add(&d.qfrc_smooth, nv); c.add(&d.qacc_smooth, nv); c.add(&d.qfrc_constraint, nv); c.add(&d.qfrc_inverse, nv);
  c.add(&d.efc_J, njmax * nv); c.add(&d.efc_pos, njmax); c.add(&d.efc_margin, njmax);
  c.add(&d.efc_frictionloss, njmax); c.add(&d.efc_diagApprox, njmax); c.add(&d.efc_KBIP, 4 * njmax);
  c.add(&d.efc_D, njmax); c.add(&d.efc_R, njmax); c.add(&d.efc_vel, njmax); c.add(&d.efc_aref, njmax);
  c.add(&d.efc_b, njmax); c.add(&d.efc_force, njmax); c.add(&d.efc_AR, njmax * njmax);
  c.carve(s->buf);
  s->ibuf.assign(2 * njmax + 2, 0);
  d.efc_type = s->ibuf.data();
  d.efc_id = s->ibuf.data() + njmax;
  s->contacts.assign(ncm + 1, mjContact{});
  d.contact = s->contacts.data();
  reset_data(m, &d);
  return &d;
}

void reset_data(const mjModel* m, mjData* d) {
  auto* s = static_cast<DataStore*>(d->owner_);
  std::fill(s->buf.begin(), s->buf.end(), 0.0);
  std::fill(s->ibuf.begin(), s->ibuf.end(), 0);
  d->ncon = d->nefc = d->ne = d->nf = 0;
  d->time = 0;
  d->energy[0] = d->energy[1] = 0;
  d->solver_iter = 0;
  for (int i = 0; i < m->nq; i++) d->qpos[i] = m->qpos0[i];
  for (int b = 0; b < m->nbody; b++) {
    int mid = m->body_mocapid[b];
    if (mid >= 0) {
      for (int k = 0; k < 3; k++) d->mocap_pos[3 * mid + k] = m->body_pos[3 * b + k];
      for (int k = 0; k < 4; k++) d->mocap_quat[4 * mid + k] = m->body_quat[4 * b + k];
    }
  }
}

int model_int(const mjModel* m, const char* name, int* out) {
#define I(f) if (!std::strcmp(name, #f)) { *out = m->f; return 1; }
  I(nq) I(nv) I(nu) I(na) I(nbody) I(njnt) I(ngeom) I(nmesh) I(nmeshvert) I(neq) I(nexclude) I(nM) I(nmocap)
  I(nsensor) I(nsensordata) I(nnames) I(npair) I(nconmax) I(njmax)
#undef I
  if (!std::strcmp(name, "opt.iterations")) { *out = m->opt.iterations; return 1; }
  if (!std::strcmp(name, "opt.integrator")) { *out = m->opt.integrator; return 1; }
  if (!std::strcmp(name, "opt.disableflags")) { *out = m->opt.disableflags; return 1; }
  if (!std::strcmp(name, "opt.enableflags")) { *out = m->opt.enableflags; return 1; }
  return -1;
}

int model_array(const mjModel* m, const char* name, const void** ptr, int* is_int) {
  const int nq = m->nq, nv = m->nv, nb = m->nbody, nj = m->njnt, ng = m->ngeom;
#define D(f, n) if (!std::strcmp(name, #f)) { *ptr = m->f; *is_int = 0; return (n); }
#define N(f, n) if (!std::strcmp(name, #f)) { *ptr = m->f; *is_int = 1; return (n); }
#define B(f, n) if (!std::strcmp(name, #f)) { *ptr = m->f; *is_int = 2; return (n); }
  D(qpos0, nq) D(qpos_spring, nq)
  N(body_parentid, nb) N(body_rootid, nb) N(body_weldid, nb) N(body_mocapid, nb) N(body_jntnum, nb)
  N(body_jntadr, nb) N(body_dofnum, nb) N(body_dofadr, nb) N(body_geomnum, nb) N(body_geomadr, nb)
  D(body_pos, 3 * nb) D(body_quat, 4 * nb) D(body_ipos, 3 * nb) D(body_iquat, 4 * nb) D(body_mass, nb)
  D(body_subtreemass, nb) D(body_inertia, 3 * nb) D(body_invweight0, 2 * nb) D(body_gravcomp, nb)
  N(jnt_type, nj) N(jnt_qposadr, nj) N(jnt_dofadr, nj) N(jnt_bodyid, nj) B(jnt_limited, nj)
  D(jnt_solref, 2 * nj) D(jnt_solimp, 5 * nj) D(jnt_pos, 3 * nj) D(jnt_axis, 3 * nj) D(jnt_stiffness, nj)
  D(jnt_range, 2 * nj) D(jnt_margin, nj)
  N(dof_bodyid, nv) N(dof_jntid, nv) N(dof_parentid, nv) N(dof_Madr, nv) D(dof_solref, 2 * nv)
  D(dof_solimp, 5 * nv) D(dof_frictionloss, nv) D(dof_armature, nv) D(dof_damping, nv) D(dof_invweight0, nv)
  N(geom_type, ng) N(geom_contype, ng) N(geom_conaffinity, ng) N(geom_condim, ng) N(geom_bodyid, ng)
  N(geom_dataid, ng) N(geom_priority, ng) D(geom_size, 3 * ng) D(geom_rbound, ng) D(geom_pos, 3 * ng)
  D(geom_quat, 4 * ng) D(geom_friction, 3 * ng) D(geom_solmix, ng) D(geom_solref, 2 * ng)
  D(geom_solimp, 5 * ng) D(geom_margin, ng) D(geom_gap, ng)
  N(mesh_vertadr, m->nmesh) N(mesh_vertnum, m->nmesh) D(mesh_vert, 3 * m->nmeshvert)
  N(eq_type, m->neq) N(eq_obj1id, m->neq) N(eq_obj2id, m->neq) B(eq_active, m->neq)
  D(eq_solref, 2 * m->neq) D(eq_solimp, 5 * m->neq) D(eq_data, mjNEQDATA * m->neq)
  N(pair_geom1, m->npair) N(pair_geom2, m->npair)
#undef D
#undef N
#undef B
  if (!std::strcmp(name, "opt.gravity")) { *ptr = m->opt.gravity; *is_int = 0; return 3; }
  if (!std::strcmp(name, "opt.timestep")) { *ptr = &m->opt.timestep; *is_int = 0; return 1; }
  if (!std::strcmp(name, "opt.tolerance")) { *ptr = &m->opt.tolerance; *is_int = 0; return 1; }
  if (!std::strcmp(name, "opt.impratio")) { *ptr = &m->opt.impratio; *is_int = 0; return 1; }
  if (!std::strcmp(name, "stat.meaninertia")) { *ptr = &m->stat.meaninertia; *is_int = 0; return 1; }
  if (!std::strcmp(name, "geom_rgba")) { *ptr = m->geom_rgba; *is_int = 3; return 4 * ng; }
  return -1;
}

int data_array(const mjModel* m, mjData* d, const char* name, void** ptr, int* is_int) {
  const int nq = m->nq, nv = m->nv, nb = m->nbody, nj = m->njnt, ng = m->ngeom, nM = m->nM;
  const int njmax = m->njmax;
#define D(f, n) if (!std::strcmp(name, #f)) { *ptr = d->f; *is_int = 0; return (n); }
#define N(f, n) if (!std::strcmp(name, #f)) { *ptr = d->f; *is_int = 1; return (n); }
  D(qpos, nq) D(qvel, nv) D(qacc, nv) D(qacc_warmstart, nv) D(qfrc_applied, nv) D(xfrc_applied, 6 * nb)
  D(mocap_pos, 3 * m->nmocap) D(mocap_quat, 4 * m->nmocap) D(sensordata, m->nsensordata)
  D(xpos, 3 * nb) D(xquat, 4 * nb) D(xmat, 9 * nb) D(xipos, 3 * nb) D(ximat, 9 * nb) D(xanchor, 3 * nj)
  D(xaxis, 3 * nj) D(geom_xpos, 3 * ng) D(geom_xmat, 9 * ng) D(subtree_com, 3 * nb) D(cinert, 10 * nb)
  D(crb, 10 * nb) D(cdof, 6 * nv) D(qM, nM) D(qLD, nM) D(qLDiagInv, nv) D(cvel, 6 * nb) D(cdof_dot, 6 * nv)
  D(qfrc_bias, nv) D(qfrc_passive, nv) D(cacc, 6 * nb) D(cfrc_int, 6 * nb) D(qfrc_smooth, nv)
  D(qacc_smooth, nv) D(qfrc_constraint, nv) D(qfrc_inverse, nv)
  D(efc_J, njmax * nv) D(efc_pos, njmax) D(efc_margin, njmax) D(efc_frictionloss, njmax)
  D(efc_diagApprox, njmax) D(efc_KBIP, 4 * njmax) D(efc_D, njmax) D(efc_R, njmax) D(efc_vel, njmax)
  D(efc_aref, njmax) D(efc_b, njmax) D(efc_force, njmax) D(efc_AR, njmax * njmax)
  N(efc_type, njmax) N(efc_id, njmax)
#undef D
#undef N
  if (!std::strcmp(name, "time")) { *ptr = &d->time; *is_int = 0; return 1; }
  if (!std::strcmp(name, "energy")) { *ptr = d->energy; *is_int = 0; return 2; }
  if (!std::strcmp(name, "ncon")) { *ptr = &d->ncon; *is_int = 1; return 1; }
  if (!std::strcmp(name, "nefc")) { *ptr = &d->nefc; *is_int = 1; return 1; }
  if (!std::strcmp(name, "solver_iter")) { *ptr = &d->solver_iter; *is_int = 1; return 1; }
  return -1;
}

}  // namespace b2
