// batch.cu — host side of the batched engine: model packing, HBM layout, kernel launch pipeline and the extern "C"
// ABI declared in include/b2_batch.h.  One b2_batch = nenv environments of one model on one GPU (one process per
// GPU; shards are contiguous environment ranges).  No CPU fallback: every entry point fails
// loudly when CUDA is not usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "b2_batch.h"
#include "batch_internal.h"
#include "k_args.h"
#include "k_collide.cuh"
#include "k_project_tc.cuh"
namespace b2 { void mirror_post(const mjModel* m, mjData* d, const double* xfrc_applied); }
#include "k_common.cuh"
#include "k_constraint.cuh"
#include "k_smooth.cuh"
#include "model_store.h"

namespace b2 {
thread_local std::string g_err;
int set_error(const std::string& msg) { g_err = msg; return -1; }
}  // namespace b2

namespace {
using b2::Field;
int fail(const std::string& msg) { return b2::set_error(msg); }
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)
}  // namespace

#ifndef PGS_ISL_MINB
#define PGS_ISL_MINB 8
#endif
#ifndef PGS_MINB
#define PGS_MINB 16
#endif
enum { SLOT_HW_WRITE = 0, SLOT_SMOOTH, SLOT_COLLIDE, SLOT_MAKE, SLOT_PROJECT, SLOT_PGS, SLOT_INTEGRATE, SLOT_HW_READ, B2_NSLOT };
// record the boundary event that precedes kernel slot `slot` of the tick being profiled
static inline void prof_mark(b2_batch* b, int slot) {
  if (!b->prof_on || b->prof_tick_open < 0) return;
  cudaEventRecord(b->prof_ev[(size_t)b->prof_tick_open * (B2_NSLOT + 1) + slot], b->stream);
  b->prof_mask[b->prof_tick_open] |= 1u << slot;
}

namespace {
using namespace b2;

// last dof on the chain from body i to the world (-1: the body is welded to the world)
int lastdof_of(const mjModel* m, int i) {
  while (i > 0 && m->body_dofnum[i] == 0) i = m->body_parentid[i];
  return i > 0 ? m->body_dofadr[i] + m->body_dofnum[i] - 1 : -1;
}

// ---- model packing ----
int pack_model(b2_batch* b) {
  const mjModel* m = b->m;
  DModel& h = b->hdr;
  std::memset(&h, 0, sizeof(h));
  const int rb = b->prec;
  h.nq = m->nq; h.nv = m->nv; h.nbody = m->nbody; h.njnt = m->njnt; h.ngeom = m->ngeom; h.nM = m->nM; h.neq = m->neq;
  h.npair = m->npair; h.nconmax = std::max(1, m->nconmax); h.njmax = std::max(1, m->njmax); h.nmocap = m->nmocap;
  h.nodom = (int)b->odom_qpos.size() / 3;
  // kinematic trees: children of the world that carry dofs somewhere below them; their dofs are contiguous
  std::vector<int> body_tree(m->nbody, -1), dof_tree(m->nv, -1), tree_adr, tree_num;
  {
    std::vector<int> root_tree(m->nbody, -1);
    for (int d = 0; d < m->nv; d++) {
      const int root = m->body_rootid[m->dof_bodyid[d]];
      if (root_tree[root] < 0) { root_tree[root] = (int)tree_adr.size(); tree_adr.push_back(d); tree_num.push_back(0); }
      dof_tree[d] = root_tree[root];
      tree_num[root_tree[root]] = d - tree_adr[root_tree[root]] + 1;
    }
    for (int i = 1; i < m->nbody; i++) body_tree[i] = lastdof_of(m, i) >= 0 ? root_tree[m->body_rootid[i]] : -1;
    std::vector<int> sz(tree_num);
    std::sort(sz.begin(), sz.end(), [](int x, int y) { return x > y; });
    h.ntree = (int)tree_adr.size();
    h.wmax = std::max(1, (sz.size() > 0 ? sz[0] : 0) + (sz.size() > 1 ? sz[1] : 0));
  }
  const int ntree = h.ntree;
  h.disableflags = b->opt_disableflags; h.enableflags = m->opt.enableflags; h.iterations = b->opt_iterations;
  for (int k = 0; k < 3; k++) h.gravity[k] = (float)m->opt.gravity[k];
  h.tolerance = (float)b->opt_tolerance; h.meaninertia = (float)m->stat.meaninertia; h.impratio = (float)m->opt.impratio;
  h.real_bytes = rb;
  const int nq = h.nq, nv = h.nv, nbody = h.nbody, njnt = h.njnt, ngeom = h.ngeom, nM = h.nM, neq = h.neq, npair = h.npair,
            nodom = h.nodom;
  (void)nq; (void)nM;
  for (int i = 0; i < nv; i++) {
    h.has_damping |= m->dof_damping[i] > 0;
    h.has_frictionloss |= m->dof_frictionloss[i] > 0;
  }
  for (int i = 0; i < nbody; i++) h.has_gravcomp |= m->body_gravcomp[i] != 0;
  for (int j = 0; j < njnt; j++) { h.has_stiffness |= m->jnt_stiffness[j] != 0; h.has_limits |= m->jnt_limited[j] != 0; }
  for (unsigned char c : b->controlled) h.has_controlled |= c != 0;

  // offsets
  const int header_words = (int)((sizeof(DModel) + 15) / 16 * 4);
  int off = header_words;
  auto place = [&](int n, bool real) {
    if (real && rb == 8) off = (off + 1) & ~1;
    const int o = off;
    off += n * (real && rb == 8 ? 2 : 1);
    return o;
  };
#define X(name, kind, count) h.o_##name = place((count), #kind[0] == 'F');
  B2_MODEL_ARRAYS(X)
#undef X
  h.nwords = (off + 3) & ~3;
  // mesh vertices live behind the staged part of the blob (HBM only; k_collide's support function reads them in place)
  h.o_mesh_vert = h.nwords;
  h.nmeshvert = m->nmeshvert;
  const int tail_words = ((3 * m->nmeshvert * (rb == 8 ? 2 : 1)) + 3) & ~3;
  int ws = 0;
#define X(name, count) h.w_##name = ws; ws += (count);
  B2_WS_ARRAYS(X)
#undef X
  h.ws_slots = ws;

  b->blob.assign(h.nwords + tail_words, 0u);
  std::memcpy(b->blob.data(), &h, sizeof(h));
  // derived tables
  std::vector<int> lastdof(nbody, -1);
  for (int i = 1; i < nbody; i++) lastdof[i] = m->body_dofnum[i] > 0 ? m->body_dofadr[i] + m->body_dofnum[i] - 1 : lastdof[m->body_parentid[i]];
  std::vector<int> ctl(nv, 0);
  for (int i = 0; i < nv && i < (int)b->controlled.size(); i++) ctl[i] = b->controlled[i] ? 1 : 0;

  auto put_int = [&](int o, const int* src, int n) { for (int i = 0; i < n; i++) b->blob[o + i] = (uint32_t)src[i]; };
  auto put_real = [&](int o, const double* src, int n) {
    if (rb == 8) std::memcpy(&b->blob[o], src, sizeof(double) * (size_t)n);
    else for (int i = 0; i < n; i++) { float f = (float)src[i]; std::memcpy(&b->blob[o + i], &f, 4); }
  };
  auto fill = [&](const char* name, int o, int n, bool real) -> int {
    if (n == 0) return 0;
    if (!std::strcmp(name, "body_lastdof")) { put_int(o, lastdof.data(), n); return 0; }
    if (!std::strcmp(name, "dof_controlled")) { put_int(o, ctl.data(), n); return 0; }
    if (!std::strcmp(name, "odom_dof")) { put_int(o, b->odom_dof.data(), n); return 0; }
    if (!std::strcmp(name, "odom_qpos")) { put_int(o, b->odom_qpos.data(), n); return 0; }
    if (!std::strcmp(name, "body_treeid")) { put_int(o, body_tree.data(), n); return 0; }
    if (!std::strcmp(name, "dof_treeid")) { put_int(o, dof_tree.data(), n); return 0; }
    if (!std::strcmp(name, "tree_dofadr")) { put_int(o, tree_adr.data(), n); return 0; }
    if (!std::strcmp(name, "tree_dofnum")) { put_int(o, tree_num.data(), n); return 0; }
    if (!std::strncmp(name, "lane_", 5)) {
      // per-lane item lists of the tree-parallel kernels: tree t belongs to lane t % L, static bodies (and the geoms on
      // them) to lane 0; ascending inside a lane, the lanes one after the other; lane_off[9 * kind + lane] = first position
      const int L = std::max(1, b->tree_lanes);
      // trees to lanes by longest-processing-time-first on a cost proxy (dofs + 2 bodies of the tree: an articulated arm
      // costs about twice a free body of the same dof count), so that the lanes of an environment finish together
      std::vector<int> tree_lane(std::max(1, ntree), 0);
      {
        std::vector<long long> w(std::max(1, ntree), 0), load(L, 0);
        for (int t = 0; t < ntree; t++) w[t] = tree_num[t];
        for (int i = 1; i < nbody; i++) if (body_tree[i] >= 0) w[body_tree[i]] += 2;
        std::vector<int> order(ntree);
        for (int t = 0; t < ntree; t++) order[t] = t;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return w[x] > w[y]; });
        for (int t : order) {
          int best = 0;
          for (int l = 1; l < L; l++) if (load[l] < load[best]) best = l;
          tree_lane[t] = best; load[best] += w[t];
        }
      }
      auto lane_of_body = [&](int i) { return body_tree[i] < 0 ? 0 : tree_lane[body_tree[i]]; };
      std::vector<std::vector<int>> lb(L), ld(L), lj(L), lg(L);
      for (int i = 1; i < nbody; i++) lb[lane_of_body(i)].push_back(i);
      for (int i = 0; i < nv; i++) ld[tree_lane[dof_tree[i]]].push_back(i);
      for (int j = 0; j < njnt; j++) lj[lane_of_body(m->jnt_bodyid[j])].push_back(j);
      for (int g = 0; g < ngeom; g++) lg[lane_of_body(m->geom_bodyid[g])].push_back(g);
      std::vector<int> offs(36, 0), flat;
      const std::vector<std::vector<int>>* kinds[4] = {&lb, &ld, &lj, &lg};
      const int which = !std::strcmp(name, "lane_body") ? 0 : !std::strcmp(name, "lane_dof") ? 1 : !std::strcmp(name, "lane_jnt") ? 2 : !std::strcmp(name, "lane_geom") ? 3 : -1;
      for (int k = 0; k < 4; k++) {
        int pos = k == 0 ? 1 : 0;   // (positions of the body list start at 1: position 0 is the world body)
        std::vector<int> fl(pos, 0);
        for (int l = 0; l < 8; l++) {
          offs[9 * k + l] = pos;
          if (l < L) { for (int v : (*kinds[k])[l]) fl.push_back(v); pos += (int)(*kinds[k])[l].size(); }
        }
        offs[9 * k + 8] = pos;
        if (k == which) flat = fl;
      }
      if (which < 0) put_int(o, offs.data(), 36);
      else put_int(o, flat.data(), std::min(n, (int)flat.size()));
      return 0;
    }
    if (!std::strcmp(name, "dof_Mcnt") || !std::strcmp(name, "dof_anc")) {
      // flattened ancestor lists in the layout of qM: the tree recursions index them instead of chasing dof_parentid,
      // which turns a chain of dependent table loads per hop into independent, pipelinable ones
      const bool cnt = name[5] == 'c';
      for (int i = 0; i < nv; i++) {
        int a = 0;
        for (int j = i; j >= 0; j = m->dof_parentid[j], a++)
          if (!cnt) b->blob[o + m->dof_Madr[i] + a] = (uint32_t)j;
        if (cnt) b->blob[o + i] = (uint32_t)a;
      }
      return 0;
    }
    if (!std::strcmp(name, "geom_vertadr") || !std::strcmp(name, "geom_vertnum")) {
      const bool adr = name[9] == 'a';
      for (int g = 0; g < n; g++) {
        const int id = m->geom_type[g] == mjGEOM_MESH ? m->geom_dataid[g] : -1;
        b->blob[o + g] = (uint32_t)(id < 0 ? 0 : (adr ? m->mesh_vertadr[id] : m->mesh_vertnum[id]));
      }
      return 0;
    }
    if (!std::strcmp(name, "opt_real")) {
      const double v[8] = {m->opt.gravity[0], m->opt.gravity[1], m->opt.gravity[2], b->opt_tolerance, m->stat.meaninertia, m->opt.impratio, 0, 0};
      put_real(o, v, 8);
      return 0;
    }
    const void* p = nullptr;
    int kind = 0;
    const int cnt = b2::model_array(m, name, &p, &kind);
    if (cnt < n || !p) return fail(std::string("pack_model: no source for '") + name + "'");
    if (kind == 0) { if (!real) return fail(std::string("pack_model: kind mismatch for ") + name); put_real(o, (const double*)p, n); }
    else if (kind == 1) { if (real) return fail(std::string("pack_model: kind mismatch for ") + name); put_int(o, (const int*)p, n); }
    else if (kind == 2) { const unsigned char* s = (const unsigned char*)p; for (int i = 0; i < n; i++) b->blob[o + i] = s[i]; }
    else return fail(std::string("pack_model: unsupported kind for ") + name);
    return 0;
  };
#define X(name, kind, count) if (fill(#name, h.o_##name, (count), #kind[0] == 'F') < 0) return -1;
  B2_MODEL_ARRAYS(X)
#undef X
  if (m->nmeshvert > 0) put_real(h.o_mesh_vert, m->mesh_vert, 3 * m->nmeshvert);
  (void)neq; (void)npair; (void)nodom; (void)ngeom; (void)ntree;
  return 0;
}

void drop_graphs(b2_batch* b);
int tick_eager(b2_batch* b, int flags);
int upload_model(b2_batch* b) {
  drop_graphs(b);  // kernel arguments / constants may change
  if (pack_model(b) < 0) return -1;
  if (!b->blob_dev) {
    CK(cudaMalloc(&b->blob_dev, b->blob.size() * 4 + 64));
  }
  CK(cudaMemcpyAsync(b->blob_dev, b->blob.data(), b->blob.size() * 4, cudaMemcpyHostToDevice, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return 0;
}

int alloc_field(b2_batch* b, const char* name, long long count, int kind, void** out) {
  drop_graphs(b);
  b->args_valid = false;
  const size_t esz = kind == 1 ? 4 : (size_t)b->prec;
  const size_t bytes = std::max<size_t>(16, (size_t)count * b->nenvp * esz);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(std::string("cudaMalloc(") + name + ", " + std::to_string(bytes) + " B): " + cudaGetErrorString(e));
  CK(cudaMemsetAsync(p, 0, bytes, b->stream));
  b->allocs.push_back(p);
  Field f;
  f.ptr = p; f.count = (int)count; f.kind = kind;
  b->fields[name] = f;
  if (out) *out = p;
  return 0;
}

template <typename T>
KArgs<T> build_args(b2_batch* b) {
  KArgs<T> a;
  std::memset(&a, 0, sizeof(a));
  a.model = b->blob_dev;
  a.model_words = b->hdr.nwords;
  a.nenv = b->nenv; a.nenvp = b->nenvp; a.ncount = b->nenvp; a.env_base = 0;
  auto R = [&](const char* n) { auto it = b->fields.find(n); return it == b->fields.end() ? (T*)nullptr : (T*)it->second.ptr; };
  auto I = [&](const char* n) { auto it = b->fields.find(n); return it == b->fields.end() ? (int*)nullptr : (int*)it->second.ptr; };
  a.qpos = R("qpos"); a.qvel = R("qvel"); a.qacc = R("qacc"); a.qacc_warmstart = R("qacc_warmstart");
  a.qfrc_applied = R("qfrc_applied"); a.xfrc_applied = R("xfrc_applied"); a.mocap_pos = R("mocap_pos"); a.mocap_quat = R("mocap_quat");
  a.ddq = R("ddq"); a.dq = R("dq"); a.odom_vels = R("odom_vels"); a.time = R("time");
  a.qfrc_bias = R("qfrc_bias"); a.qfrc_inverse = R("qfrc_inverse"); a.xpos = R("xpos"); a.xquat = R("xquat");
  a.xmat = R("xmat"); a.geom_xpos = R("geom_xpos"); a.geom_xmat = R("geom_xmat"); a.subtree_com = R("subtree_com");
  a.cdof = R("cdof"); a.qM = R("qM"); a.qLD = R("qLD"); a.qLDiagInv = R("qLDiagInv"); a.qfrc_passive = R("qfrc_passive");
  a.qfrc_smooth = R("qfrc_smooth"); a.qacc_smooth = R("qacc_smooth"); a.qfrc_constraint = R("qfrc_constraint");
  a.ws = R("_ws");
  a.con = R("contact"); a.coni = I("contact_int"); a.ncon = I("ncon"); a.nefc = I("nefc"); a.efc_type = I("efc_type");
  a.efc_id = I("efc_id"); a.efc_tree = I("efc_tree"); a.efc_J = R("efc_J"); a.efc_pos = R("efc_pos"); a.efc_margin = R("efc_margin");
  a.efc_frictionloss = R("efc_frictionloss"); a.efc_diagApprox = R("efc_diagApprox"); a.efc_R = R("efc_R"); a.efc_D = R("efc_D");
  a.efc_KBI = R("efc_KBI"); a.efc_vel = R("efc_vel"); a.efc_aref = R("efc_aref"); a.efc_b = R("efc_b"); a.efc_force = R("efc_force"); a.efc_finv = R("efc_finv");
  a.efc_ARdiag = R("efc_ARdiag"); a.efc_B = R("efc_B"); a.efc_blocks = R("efc_blocks"); a.efc_nwords = I("efc_nwords"); a.env_order = I("env_order");
  a.blk_row0 = I("blk_row0"); a.blk_off = I("blk_off"); a.nblk = I("nblk"); a.maxblk = I("_maxblk"); a.solver_iter = I("solver_iter"); a.status = I("status");
  a.pending = I("_pending");
  a.isl_off = I("isl_off"); a.isl_end = I("isl_end"); a.nisl = I("nisl");
  a.efc_Jem = R("efc_Jem"); a.efc_Bem = R("efc_Bem"); a.minv_em = R("minv_em");
  return a;
}
// the field pointers are looked up by name once per (re)allocation, not once per tick
template <typename T>
KArgs<T> make_args(b2_batch* b, int flags) {
  if (!b->args_valid) {
    b->args_f = build_args<float>(b);
    b->args_d = build_args<double>(b);
    b->args_valid = true;
  }
  KArgs<T> a;
  if constexpr (sizeof(T) == 4) a = b->args_f; else a = b->args_d;
  a.model = b->blob_dev; a.model_words = b->hdr.nwords;
  a.flags = flags; a.ws_block = b->smooth_block; a.h = (T)b->h; a.wp = b->wp;
  a.hp = b->h_dev ? (sizeof(T) == 8 ? (const T*)b->h_dev : (const T*)((const char*)b->h_dev + 8)) : nullptr;
  a.ld_extra = b->ld_extra;
  a.block_capw = b->block_capw; a.block_npar = b->block_npar; a.stage_cap = b->isl_cap ? b->isl_stage : b->stage_cap; a.row_nb = b->row_nb; a.isl_cap = b->isl_cap; a.em_rows = b->tc_rows;
  a.obs_peers = b->obs_peers_dev; a.obs_world = b->obs_world; a.obs_rank = b->obs_rank; a.obs_nenv = b->nenv;
  a.hw_vel = b->io_in[0]; a.hw_eff = b->io_in[1];
  a.hw_pos = b->io_out[0]; a.hw_velo = b->io_out[1]; a.hw_effo = b->io_out[2];
  a.hw_kp = b->hw_kp; a.hw_kd = b->hw_kd;
  return a;
}

// Sub-batch window [lo, lo + n) of the batch: every pointer advanced to the window's first environment (SoA arrays keep
// the batch's environment stride nenvp, environment-major slabs move by whole slabs), counters get the window's own
// word.  The kernels index environments from 0 and cover `ncount` of them, so a window is launched exactly like a batch.
template <typename T>
KArgs<T> window_args(KArgs<T> a, int s, int lo, int n) {
  auto adv = [&](auto*& p, long long per_env = 1) { if (p) p += (long long)lo * per_env; };
  adv(a.qpos); adv(a.qvel); adv(a.qacc); adv(a.qacc_warmstart); adv(a.qfrc_applied); adv(a.xfrc_applied); adv(a.mocap_pos); adv(a.mocap_quat);
  adv(a.ddq); adv(a.dq); adv(a.odom_vels); adv(a.time); adv(a.qfrc_bias); adv(a.qfrc_inverse); adv(a.xpos); adv(a.xquat);
  adv(a.xmat); adv(a.geom_xpos); adv(a.geom_xmat); adv(a.subtree_com); adv(a.cdof); adv(a.qM); adv(a.qLD); adv(a.qLDiagInv);
  adv(a.qfrc_passive); adv(a.qfrc_smooth); adv(a.qacc_smooth); adv(a.qfrc_constraint); adv(a.ws);
  adv(a.con); adv(a.coni); adv(a.ncon); adv(a.nefc); adv(a.efc_type); adv(a.efc_id); adv(a.efc_tree); adv(a.efc_J);
  adv(a.efc_pos); adv(a.efc_margin); adv(a.efc_frictionloss); adv(a.efc_diagApprox); adv(a.efc_R); adv(a.efc_D); adv(a.efc_KBI);
  adv(a.efc_vel); adv(a.efc_aref); adv(a.efc_b); adv(a.efc_force); adv(a.efc_finv); adv(a.efc_ARdiag); adv(a.efc_B);
  adv(a.efc_Jem, (long long)a.em_rows * 64); adv(a.efc_Bem, (long long)a.em_rows * 64); adv(a.minv_em, 64 * 64);
  adv(a.efc_blocks, a.block_capw);
  adv(a.efc_nwords); adv(a.env_order); adv(a.blk_row0); adv(a.blk_off); adv(a.nblk); adv(a.isl_off); adv(a.isl_end); adv(a.nisl);
  adv(a.solver_iter); adv(a.status);
  if (a.maxblk) a.maxblk += s;
  if (a.pending) a.pending += s;
  a.env_base = lo;
  a.ncount = n;
  a.nenv = std::max(0, std::min(n, a.nenv - lo));
  return a;
}

// shared-memory budgets of the constraint-pipeline kernels (they stage the same model blob; row assembly and the solver
// add their own vectors); redone whenever the blob changes size
int configure_constraint_kernels(b2_batch* b) {
  if (b->fused) return 0;
    // the constraint-pipeline kernels stage the same blob; row assembly and the solver add their shared vectors
    const int need1 = (int)b->blob_smem;
    const int need1i = (int)(b->blob_smem + b->ld_smem);
    // k_make_rows: a shared-memory column per thread for the base rows of the contact being assembled
    {
      int nbm = 1;
      for (int g = 0; g < b->m->ngeom; g++) nbm = std::max(nbm, b->m->geom_condim[g]);
      nbm = std::min(nbm, 6);
      const size_t rs = (size_t)nbm * b->hdr.wmax * 128 * b->prec;
      b->row_nb = (!getenv("B2_NO_ROW_SMEM") && b->blob_smem + rs <= 180 * 1024) ? nbm : 0;
      b->row_smem = b->row_nb ? rs : 0;
    }
    // k_make_rows with island ordering: two int columns of ntree entries per thread behind the row column
    b->isl_smem = b->isl_cap ? (size_t)2 * b->hdr.ntree * 128 * sizeof(int) : 0;
    const int need1r = (int)(b->blob_smem + b->row_smem + b->isl_smem);
    if (b->isl_cap) {
      // k_pgs_island: EPB environments per one-warp CTA, each with its vectors and its staged records; the stage is sized
      // for eight resident CTAs per SM
      const int epbi = 32 / b->pgs_isl;
      const long long fixedi = ((long long)2 * ((b->hdr.nv + 7) & ~3) + (long long)2 * ((b->hdr.njmax + 3) & ~3)) * b->prec + (long long)3 * b->isl_cap * 4;
      // what fits eight resident CTAs per SM, raised — down to four CTAs — towards the record volume a busy environment
      // has (a quarter of the contact cap, each a largest-shape record): measured on C5, whose environments carry 2200
      // words on average: 2304 staged words 1.84 ms, 1232 (eight CTAs) 2.34 ms, 3328 2.07 ms
      int nbmx = 1;
      for (int g = 0; g < b->m->ngeom; g++) nbmx = std::max(nbmx, b->m->geom_condim[g]);
      const long long fit8 = (long long)((220 * 1024 / 8) / epbi - fixedi) / b->prec, fit4 = (long long)((220 * 1024 / 4) / epbi - fixedi) / b->prec;
      const long long busy = (long long)(b->hdr.nconmax / 4) * block_max_words(std::min(nbmx, 6), b->hdr.wmax);
      long long capi = getenv("B2_PGS_STAGE") ? atoi(getenv("B2_PGS_STAGE")) : std::max(fit8, std::min(busy, fit4));
      capi = std::max(0LL, std::min<long long>(capi, b->block_capw)) & ~3LL;
      b->isl_stage = (int)capi;
    }
    if (need1 > 227 * 1024 || need1i > 227 * 1024 || need1r > 227 * 1024) return fail("model too large for the constraint kernels' shared memory");
    // row assembly: one record column per thread; 128-thread CTAs when that fits, else 32
    // (k_make_blocks reads the model from HBM: its shared memory is the record columns only)
    b->make_block = (size_t)b->rec_max * 129 * b->prec <= 56 * 1024 ? 128 : 32;
    const int need2 = (int)((size_t)b->rec_max * (b->make_block + 1) * b->prec);
    // solver: per environment 2 ((b->hdr.nv + 7) & ~3) + njmax words of vectors plus the staged records; sized for ~4 CTAs per SM
    const int epb = 128 / b->pgs_lanes;
    const long long fixed = (long long)b->blob_smem + ((long long)2 * ((b->hdr.nv + 7) & ~3) + ((b->hdr.njmax + 3) & ~3)) * epb * b->prec;
    // staging an environment's records in shared memory lost to plain L1-cached streaming once the records became compact
    // (profiles/r01_pgs_variants.txt): off unless asked for
    long long cap = getenv("B2_PGS_STAGE") ? atoi(getenv("B2_PGS_STAGE")) : 0;
    cap = std::max(0LL, std::min<long long>(cap, b->block_capw)) & ~3LL;
    if (fixed + cap * epb * b->prec > 227 * 1024) cap = 0;
    b->stage_cap = (int)cap;
    const int need3 = (int)(fixed + cap * epb * b->prec);
    if (need2 > 227 * 1024 || need3 > 227 * 1024) return fail("model too large for the constraint kernels' shared memory");
    b->pgs_ctas_per_sm = std::max(1, std::min(16, (int)(227 * 1024 / std::max(1, need3 + 1024))));
    // k_solve_rows (wide trees): factor tile + one right-hand-side column per (row-warp, environment)
    b->solve_rows = 0; b->solve_smem = 0;
    if (b->make_block == 32 && b->fields.count("efc_B") && !getenv("B2_NO_SOLVE_ROWS")) {
      for (int rows : {16, 8, 4}) {
        const size_t need = ((size_t)b->hdr.nM + b->hdr.nv + (size_t)rows * b->hdr.wmax) * 32 * b->prec;
        if (need <= 200 * 1024) { b->solve_rows = rows; b->solve_smem = need; break; }
      }
    }
    // The opt-in limit is a property of (function, device), not of a batch: raise every kernel to the hardware maximum
    // once per device and never lower it, so that batches of different models can coexist in one process.
    static bool attr_done[2][64] = {{false}};
    bool ok = true;
    auto SA = [&](const void* fn) { ok &= cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess; };
    const int pi = b->prec == 8 ? 1 : 0, di = b->device & 63;
    if (!attr_done[pi][di]) {
    if (b->prec == 8) {
      SA((const void*)k_collide<double, 128, 8>); SA((const void*)k_integrate<double, 128>);
      SA((const void*)k_integrate<double, 64>); SA((const void*)k_integrate<double, 32>);
      SA((const void*)k_integrate<double, 128, 8>); SA((const void*)k_integrate<double, 128, 4>); SA((const void*)k_integrate<double, 128, 2>);
      SA((const void*)k_make_rows<double, 128, 8>); SA((const void*)k_make_blocks<double, 128>); SA((const void*)k_make_blocks<double, 32>);
      SA((const void*)k_pgs_block<double, 4, 32, PGS_MINB>); SA((const void*)k_pgs_block<double, 8, 32, PGS_MINB>);
      SA((const void*)k_pgs_block<double, 16, 32, PGS_MINB>); SA((const void*)k_pgs_block<double, 32, 32, PGS_MINB>);
      SA((const void*)k_solve_rows<double, 16>); SA((const void*)k_solve_rows<double, 8>); SA((const void*)k_solve_rows<double, 4>);
      SA((const void*)k_pgs_island<double, 4, PGS_ISL_MINB>); SA((const void*)k_pgs_island<double, 8, PGS_ISL_MINB>); SA((const void*)k_pgs_island<double, 16, PGS_ISL_MINB>);
    } else {
      SA((const void*)k_collide<float, 128, 8>); SA((const void*)k_integrate<float, 128>);
      SA((const void*)k_integrate<float, 64>); SA((const void*)k_integrate<float, 32>);
      SA((const void*)k_integrate<float, 128, 8>); SA((const void*)k_integrate<float, 128, 4>); SA((const void*)k_integrate<float, 128, 2>);
      SA((const void*)k_make_rows<float, 128, 8>); SA((const void*)k_make_blocks<float, 128>); SA((const void*)k_make_blocks<float, 32>);
      SA((const void*)k_pgs_block<float, 4, 32, PGS_MINB>); SA((const void*)k_pgs_block<float, 8, 32, PGS_MINB>);
      SA((const void*)k_pgs_block<float, 16, 32, PGS_MINB>); SA((const void*)k_pgs_block<float, 32, 32, PGS_MINB>);
      SA((const void*)k_solve_rows<float, 16>); SA((const void*)k_solve_rows<float, 8>); SA((const void*)k_solve_rows<float, 4>);
      SA((const void*)k_pgs_island<float, 4, PGS_ISL_MINB>); SA((const void*)k_pgs_island<float, 8, PGS_ISL_MINB>); SA((const void*)k_pgs_island<float, 16, PGS_ISL_MINB>);
    }
    if (b->prec == 4) {
      SA((const void*)k_dense_minv<float, 8>);
      // (this kernel also has static shared memory: ask for what it uses, not for the whole SM)
      ok &= cudaFuncSetAttribute((const void*)k_project_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * tc::A_BYTES + 2 * tc::B_BYTES)) == cudaSuccess;
    }
    attr_done[pi][di] = ok;
    }
    if (!ok) return fail("cudaFuncSetAttribute failed");
  return 0;
}

template <typename T, int BLOCK, typename P, int L = 1>
int launch_smooth(b2_batch* b, const KArgs<T>& a, int grid, cudaStream_t st) {
  static bool attr_set[8] = {false};
  int dev = b->device & 7;
  if (!attr_set[dev]) {
    constexpr int MB = L > 1 ? 2 : 1;   // tree-parallel form: two CTAs per SM (255 registers: no spills; one wave at the BASELINE batch sizes)
    CK(cudaFuncSetAttribute(k_smooth<T, BLOCK, P, MB, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev] = true;
  }
  k_smooth<T, BLOCK, P, (L > 1 ? 2 : 1), L><<<grid, BLOCK, b->smooth_smem, st>>>(a);
  b->launches++;
  return 0;
}

template <typename T>
int launch_chain(b2_batch* b, const KArgs<T>& a, int grid) {
  if constexpr (sizeof(T) == 4) return launch_chain_f32(b, a, grid);
  else return launch_chain_f64(b, a, grid);
}

void io_default(b2_batch* b);
void drop_aliases(b2_batch* b);
int hw_write_async(b2_batch* b, int flags = 0, bool gather = false);  // k_hw_write from b->io_in (+ pre-integration pos / vel gather)
int hw_read_async(b2_batch* b, bool post = true, cudaStream_t st = nullptr);   // k_hw_read into b->io_out (st: default the batch's stream)
enum { B2_TICK_HW = 1 << 20 };  // internal: run k_hw_write / k_hw_read around the tick kernels

template <typename D>
__global__ void k_hold_slots(D* qpos, D* qvel, D* qacc, D* qws, const unsigned char* active, const int* qadr, const int* dadr, int nslot, int nenv, int nenvp);
template <typename T>
int hold_slots(b2_batch* b) {
  if (!b->nslot) return 0;
  const long long tot = (long long)b->nslot * b->nenv;
  const int th = 256, bl = (int)((tot + th - 1) / th);
  k_hold_slots<T><<<bl, th, 0, b->stream>>>((T*)b->fields["qpos"].ptr, (T*)b->fields["qvel"].ptr, (T*)b->fields["qacc"].ptr,
                                            (T*)b->fields["qacc_warmstart"].ptr, b->slot_active, b->slot_qadr, b->slot_dadr, b->nslot, b->nenv, b->nenvp);
  b->launches++;
  return 0;
}

int ensure_sub_streams(b2_batch* b, int nsub) {
  if (!b->sub_fork) CK(cudaEventCreateWithFlags(&b->sub_fork, cudaEventDisableTiming));
  while ((int)b->sub_stream.size() < nsub - 1) {
    cudaStream_t st; cudaEvent_t ev;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    b->sub_stream.push_back(st); b->sub_join.push_back(ev);
  }
  return 0;
}

template <typename T>
int run_tick(b2_batch* b, int flags) {
  int kf = 0;
  if (flags & B2_TICK_CONTROLLER) kf |= B2F_CONTROLLER;
  if (flags & B2_TICK_INVERSE) kf |= B2F_INVERSE;
  if (flags & B2_TICK_INTEGRATE) kf |= B2F_INTEGRATE;
  if ((flags & B2_TICK_ODOM) && b->hdr.nodom > 0) kf |= B2F_ODOM;
  if (b->fused) kf |= B2F_FUSED;
  if (flags & B2_TICK_NOSOLVE) kf |= B2F_NOSOLVE;
  if (flags & B2_TICK_READ_POST) kf |= B2F_READ_POST;
  if (b->ws_global) kf |= B2F_WS_GLOBAL;
  if (b->ws_global && b->ld_smem) kf |= B2F_LD_SMEM;
  if (b->fusable) kf |= B2F_FUSABLE;
  if (b->tick_flags & (1 << 30)) kf |= B2F_XFRC;  // set once xfrc_applied has been written
  if (b->export_stages) kf |= B2F_EXPORT;
  // observation exchange: fused into k_integrate's epilogue when the tick ends there for every environment; a tick
  // that ends elsewhere (single-kernel chain, self-integrating smooth kernel, slot parking after the integration)
  // publishes with one small kernel at its end
  const bool obs = b->obs_on && b->obs_peers_dev && (flags & B2_TICK_INTEGRATE);
  const bool obs_fused = obs && !b->fused && !b->fusable && !b->nslot && !(flags & B2_TICK_NOSOLVE) && !(b->chain_single && !(kf & B2F_XFRC));
  if (obs_fused) kf |= B2F_OBS;
  const bool single = b->chain_single && !(kf & B2F_XFRC);
  // k_chain does the hardware-interface exchange itself when the hardware joints are exactly the chain's dofs
  const bool hwio = single && (flags & B2_TICK_HW) && b->hw_identity && (kf & B2F_CONTROLLER) && !(kf & B2F_ODOM) && !getenv("B2_NO_HWIO");
  if (hwio) kf |= B2F_HWIO;
  KArgs<T> a = make_args<T>(b, kf);
  int per_sm = (int)std::max<size_t>(1, (227 * 1024) / std::max<size_t>(1, b->smooth_smem));
  if (b->chain_n == 0 && b->tree_lanes > 1) per_sm = std::min(per_sm, 2);   // (its launch bounds: two resident CTAs per SM)
  if (getenv("B2_SMOOTH_CTAS_PER_SM")) per_sm = std::max(1, atoi(getenv("B2_SMOOTH_CTAS_PER_SM")));
  auto smooth_grid = [&](int n) {
    const int ntiles = n / (b->smooth_block / (b->chain_n > 0 ? 1 : b->tree_lanes));
    return std::max(1, std::min(ntiles, b->nsm * per_sm));
  };
  int rc;
  const bool read_post = (flags & B2_TICK_READ_POST) != 0;
  if ((flags & B2_TICK_HW) && !hwio) { if (hw_write_async(b, flags, !read_post) < 0) return -1; }
  if (single) {
    // limit-only serial chain: one kernel does the whole tick (k_chain.cuh)
    const int grid = smooth_grid(b->nenvp);
    prof_mark(b, SLOT_SMOOTH);
    if (b->chain_team) { if constexpr (sizeof(T) == 4) rc = launch_chain_team_f32(b, a); else rc = launch_chain_team_f64(b, a); }
    else if constexpr (sizeof(T) == 4) rc = launch_chain1_f32(b, a, grid); else rc = launch_chain1_f64(b, a, grid);
    if (rc < 0) return rc;
    if (obs) { k_publish_obs<T><<<(b->nenv + 127) / 128, 128, 0, b->stream>>>(a); b->launches++; }
    prof_mark(b, SLOT_HW_READ);
    if ((flags & B2_TICK_HW) && !hwio) { if (hw_read_async(b, read_post) < 0) return -1; }
    CK(cudaGetLastError());
    return 0;
  }
  // The kernels of the pipeline for the environments of one window, on one stream.  (prof_mark records on the batch's
  // stream: per-kernel profiling runs the batch as a single window.)
  bool hw_read_early = false;
  int nsub_now = b->nsub;
  if (b->prof_on || b->chain_n > 0 || b->fused || b->tc_rows > 0) nsub_now = 1;
  while (nsub_now > 1 && (b->nenvp % (128 * nsub_now) != 0 || b->nenvp / nsub_now < b->sub_min_envs)) nsub_now /= 2;
  hw_read_early = (flags & B2_TICK_HW) && !read_post && !(flags & B2_TICK_NOSOLVE) && !b->fused && nsub_now <= 1 && !b->prof_on &&
                  !getenv("B2_NO_ORDER_FORK");
  if (hw_read_early && ensure_sub_streams(b, 2) < 0) return -1;
  auto pipeline = [&](const KArgs<T>& a, cudaStream_t st, int n) -> int {
  if (b->fusable) CK(cudaMemsetAsync(a.pending, 0, sizeof(int), st));
  prof_mark(b, SLOT_SMOOTH);
  const int grid = smooth_grid(n);
  if (b->chain_n > 0) rc = launch_chain<T>(b, a, grid);
  else switch (b->smooth_block) {
    case 128:
      if (b->tree_lanes == 8) rc = launch_smooth<T, 128, GenericP, 8>(b, a, grid, st);
      else if (b->tree_lanes == 4) rc = launch_smooth<T, 128, GenericP, 4>(b, a, grid, st);
      else if (b->tree_lanes == 2) rc = launch_smooth<T, 128, GenericP, 2>(b, a, grid, st);
      else rc = launch_smooth<T, 128, GenericP>(b, a, grid, st);
      break;
    case 64: rc = launch_smooth<T, 64, GenericP>(b, a, grid, st); break;
    default: rc = launch_smooth<T, 32, GenericP>(b, a, grid, st); break;
  }
  if (rc < 0) return rc;
  if (!b->fused) {
    constexpr int BL = 128;
    const int nt = n / BL;
    const int g2 = std::max(1, std::min(nt, b->nsm * 4));
    const size_t sm = b->blob_smem;
    prof_mark(b, SLOT_COLLIDE);
    if (b->m->npair > 0) {
      // a team of 8 lanes per environment: 16 environments per CTA, the candidate list of each behind the model blob
      constexpr int CL = 8;
      const size_t smc = 16 + (((size_t)b->hdr.nwords * 4 + 15) & ~(size_t)15) + (size_t)(BL / CL) * ((b->m->npair + 1) & ~1) * sizeof(uint16_t);
      // (measured and dropped: staging the tile's geom frames in shared memory for the cull — the pair walk re-reads them
      //  from L1 already: C3 0.140 -> 0.154 ms, C5 0.180 -> 0.191 ms)
      const int gc = std::max(1, std::min(n / (BL / CL), b->nsm * std::max(1, std::min(8, (int)(200 * 1024 / smc)))));
      k_collide<T, BL, CL><<<gc, BL, smc, st>>>(a);
      b->launches++;
    }
    prof_mark(b, SLOT_MAKE);
    CK(cudaMemsetAsync(a.maxblk, 0, sizeof(int), st));
    {
      constexpr int RL = 8;   // lanes per environment
      // persistent CTAs: as many as are resident at once (registers or shared memory, whichever binds)
      static int occ_cache[2] = {0, 0};
      static size_t occ_smem[2] = {0, 0};
      const int oi = sizeof(T) == 8;
      // (+ the row table of the epilogue, one word per row and environment, when it fits)
      const size_t rtb = (size_t)(BL / RL) * b->hdr.njmax * sizeof(int);
      const bool rtab = !getenv("B2_NO_ROW_TAB") && b->hdr.ntree < 4095 && sm + b->row_smem + b->isl_smem + rtb <= 100 * 1024;
      const size_t smr = sm + b->row_smem + b->isl_smem + (rtab ? rtb : 0);
      KArgs<T> ar = a;
      ar.row_tab = rtab ? 1 : 0;
      if (!occ_cache[oi] || occ_smem[oi] != smr) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_make_rows<T, BL, RL>, BL, smr) != cudaSuccess || occ < 1) { occ = 1; cudaGetLastError(); }
        occ_cache[oi] = occ; occ_smem[oi] = smr;
      }
      const int gr = std::max(1, std::min(n / (BL / RL), b->nsm * occ_cache[oi]));
      k_make_rows<T, BL, RL><<<gr, BL, smr, st>>>(ar);
    }
    // the solver's visit order needs only the row counts (k_make_rows) and last tick's iteration counts: its one-CTA sort
    // runs on a side stream next to k_make_blocks (forked here, joined in front of the solver; also inside the graph)
    cudaStream_t ost = st;
    const bool fork_order = !(flags & B2_TICK_NOSOLVE) && nsub_now <= 1 && !b->prof_on && !getenv("B2_NO_ORDER_FORK");
    if (fork_order) {
      if (ensure_sub_streams(b, 2) < 0) return -1;
      ost = b->sub_stream[0];
      CK(cudaEventRecord(b->sub_fork, st));
      CK(cudaStreamWaitEvent(ost, b->sub_fork, 0));
    }
    if (b->tc_rows > 0) {
      // one-tree model, fp32: dense M^-1 per environment, then B = J M^-1 on the tensor cores (k_project_tc.cuh)
      if constexpr (sizeof(T) == 4) {
        constexpr int MR = 8;
        const size_t smm = 16 + (((size_t)b->hdr.nwords * 4 + 15) & ~(size_t)15) + ((size_t)2 * b->hdr.nM + b->hdr.nv + (size_t)MR * b->hdr.nv) * 32 * sizeof(T);
        const int gm = std::max(1, std::min(n / 32, b->nsm * std::max(1, (int)(227 * 1024 / (smm + 1024)))));
        k_dense_minv<T, MR><<<gm, 32 * MR, smm, st>>>(a);
        const size_t smt = 2 * tc::A_BYTES + 2 * tc::B_BYTES;
        k_project_tc<<<std::min(n, 2 * b->nsm), 128, smt, st>>>(a.efc_Jem, a.minv_em, a.efc_Bem, a.nefc, n, b->tc_rows / tc::TM, b->tc_rows, b->tc_passes);
        b->launches += 2;
      }
    } else if (b->solve_rows > 0) {
      // wide trees: M^-1 J^T of all rows up front, the tile's factor shared through shared memory (k_solve_rows)
      const int gs = std::max(1, std::min(n / 32, b->nsm * std::max(1, (int)(227 * 1024 / (b->solve_smem + 1024)))));
      if (b->solve_rows == 16) k_solve_rows<T, 16><<<gs, 512, b->solve_smem, st>>>(a);
      else if (b->solve_rows == 8) k_solve_rows<T, 8><<<gs, 256, b->solve_smem, st>>>(a);
      else k_solve_rows<T, 4><<<gs, 128, b->solve_smem, st>>>(a);
      b->launches++;
    }
    {
      // one thread per (block, environment); CTAs of block ordinals beyond this tick's largest count exit at once
      const int mb = b->make_block;
      // (block ordinals beyond grid.y are covered by a loop inside the kernel: a grid of njmax rows was mostly CTAs that exit at once)
      const dim3 gb(std::max(1, std::min(n / mb, b->nsm * 8)), std::min(b->hdr.njmax, 32));
      if (mb == 128) k_make_blocks<T, 128><<<gb, 128, (size_t)b->rec_max * 129 * sizeof(T), st>>>(a);
      else k_make_blocks<T, 32><<<gb, 32, (size_t)b->rec_max * 33 * sizeof(T), st>>>(a);
    }
    b->launches += 2;
    if ((flags & B2_TICK_NOSOLVE) && (kf & B2F_INVERSE)) {   // mj_inverse through the shim: no solver pass to fold this into
      k_inverse_rows<T><<<(n + 127) / 128, 128, 0, st>>>(a);
      b->launches += 1;
    }
    if (!(flags & B2_TICK_NOSOLVE)) {
      prof_mark(b, SLOT_PGS);
      k_order_envs<256, 1024><<<1, 1024, 0, ost>>>(a.nefc, a.efc_nwords, a.env_order, n, std::max(4, b->block_capw / 256), a.pending, b->fusable ? 1 : 0,
                                              (b->isl_cap && !getenv("B2_ORDER_BY_WORDS")) ? a.solver_iter : nullptr, 1);
      if (fork_order) { CK(cudaEventRecord(b->sub_join[0], ost)); CK(cudaStreamWaitEvent(st, b->sub_join[0], 0)); }
      // (measured: the one-environment-per-team block solver is fastest with the volume-only order — PR2 5.66 ms against
      //  6.2 / 6.4 ms with last tick's iterations in the key)
      b->launches += 1;
      // one warp per CTA and one CTA per group of environments: the hardware scheduler balances the very uneven
      // per-environment work (contact counts) dynamically
      constexpr int PB = 32;
      if (b->isl_cap) {
        // several small trees: one lane per constraint island (k_pgs_island)
        const int epbi = 32 / b->pgs_isl;
        const size_t smi = ((size_t)2 * ((b->hdr.nv + 7) & ~3) + (size_t)2 * ((b->hdr.njmax + 3) & ~3) + b->isl_stage) * epbi * sizeof(T) + (size_t)epbi * 3 * b->isl_cap * sizeof(int);
        const int gi = n / epbi;
        if (b->pgs_isl == 4) k_pgs_island<T, 4, PGS_ISL_MINB><<<gi, 32, smi, st>>>(a);
        else if (b->pgs_isl == 16) k_pgs_island<T, 16, PGS_ISL_MINB><<<gi, 32, smi, st>>>(a);
        else k_pgs_island<T, 8, PGS_ISL_MINB><<<gi, 32, smi, st>>>(a);
      } else {
      const int epb = PB / b->pgs_lanes;
      const size_t smp = ((size_t)2 * ((b->hdr.nv + 7) & ~3) + ((b->hdr.njmax + 3) & ~3) + b->stage_cap) * epb * sizeof(T);
      const int g3 = n / epb;
      if (b->pgs_lanes == 4) k_pgs_block<T, 4, PB, PGS_MINB><<<g3, PB, smp, st>>>(a);
      else if (b->pgs_lanes == 16) k_pgs_block<T, 16, PB, PGS_MINB><<<g3, PB, smp, st>>>(a);
      else if (b->pgs_lanes == 32) k_pgs_block<T, 32, PB, PGS_MINB><<<g3, PB, smp, st>>>(a);
      else k_pgs_block<T, 8, PB, PGS_MINB><<<g3, PB, smp, st>>>(a);
      }
      // the joint read-back of the reference's order (effort = qfrc_inverse, complete once the solver has subtracted its
      // J^T f; positions / velocities were gathered before the step) does not wait for the integration: side stream
      if (hw_read_early) {
        cudaStream_t rst = b->sub_stream[0];
        CK(cudaEventRecord(b->sub_fork, st));
        CK(cudaStreamWaitEvent(rst, b->sub_fork, 0));
        if (hw_read_async(b, false, rst) < 0) return -1;
        CK(cudaEventRecord(b->sub_join[0], rst));
      }
      prof_mark(b, SLOT_INTEGRATE);
      const int TL = b->chain_n > 0 ? 1 : b->tree_lanes;
      if (TL > 1) {
        // tree-parallel: 128 threads = 128 / TL environments per CTA
        const size_t smi = sm + ((kf & B2F_LD_SMEM) ? b->ld_smem : 0);
        const int gi = std::max(1, std::min(n / (128 / TL), b->nsm * 8));
        if (TL == 8) k_integrate<T, 128, 8><<<gi, 128, smi, st>>>(a);
        else if (TL == 4) k_integrate<T, 128, 4><<<gi, 128, smi, st>>>(a);
        else k_integrate<T, 128, 2><<<gi, 128, smi, st>>>(a);
      } else if (kf & B2F_LD_SMEM) {
        const size_t smi = sm + b->ld_smem;
        const int bi = b->smooth_block, gi = std::max(1, std::min(n / bi, b->nsm * 8));
        if (bi == 128) k_integrate<T, 128><<<gi, 128, smi, st>>>(a);
        else if (bi == 64) k_integrate<T, 64><<<gi, 64, smi, st>>>(a);
        else k_integrate<T, 32><<<gi, 32, smi, st>>>(a);
      } else {
        k_integrate<T, BL><<<g2, BL, sm, st>>>(a);
      }
      b->launches += 2;
    }
  }
  return 0;
  };
  // Sub-batches (b2_set_option "subbatches"; one window by default): the batch is cut into nsub windows of whole
  // 128-environment tiles that run the pipeline side by side on their own streams (forked from and joined back into the
  // batch's stream, also inside a graph capture): windows in different stages fill each other's gaps.  Results do not
  // depend on the cut (an environment never looks at another one).  Measured: +2-3 % when it was built, nothing since the
  // solver visits environments by last tick's iteration count.
  const int nsub = nsub_now;
  if (nsub <= 1) {
    if (pipeline(a, b->stream, b->nenvp) < 0) return -1;
  } else {
    if (ensure_sub_streams(b, nsub) < 0) return -1;
    const int n = b->nenvp / nsub;
    CK(cudaEventRecord(b->sub_fork, b->stream));
    for (int s = 0; s < nsub; s++) {
      cudaStream_t st = s == 0 ? b->stream : b->sub_stream[s - 1];
      if (s > 0) CK(cudaStreamWaitEvent(st, b->sub_fork, 0));
      if (pipeline(window_args(a, s, s * n, n), st, n) < 0) return -1;
      if (s > 0) { CK(cudaEventRecord(b->sub_join[s - 1], st)); CK(cudaStreamWaitEvent(b->stream, b->sub_join[s - 1], 0)); }
    }
  }
  if (flags & B2_TICK_INTEGRATE) hold_slots<T>(b);   // inactive object slots go back to their parking place
  if (obs && !obs_fused) { k_publish_obs<T><<<(b->nenv + 127) / 128, 128, 0, b->stream>>>(a); b->launches++; }
  prof_mark(b, SLOT_HW_READ);
  if (hw_read_early) CK(cudaStreamWaitEvent(b->stream, b->sub_join[0], 0));
  else if (flags & B2_TICK_HW) { if (hw_read_async(b, read_post) < 0) return -1; }
  CK(cudaGetLastError());
  return 0;
}

// a profiled tick records B2_NSLOT + 1 boundary events; slots that do not run collapse to zero duration
static bool prof_open(b2_batch* b) {
  if (!b->prof_on || b->prof_tick_open >= 0 || b->prof_n >= b->prof_max) return false;
  b->prof_tick_open = b->prof_n;
  return true;
}
static void prof_close(b2_batch* b) {
  if (b->prof_tick_open < 0) return;
  cudaEventRecord(b->prof_ev[(size_t)b->prof_tick_open * (B2_NSLOT + 1) + B2_NSLOT], b->stream);
  b->prof_mask[b->prof_tick_open] |= 1u << B2_NSLOT;
  b->prof_tick_open = -1;
  b->prof_n++;
}

int tick_eager(b2_batch* b, int flags) { return b->prec == 8 ? run_tick<double>(b, flags) : run_tick<float>(b, flags); }

// The kernel sequence of one tick is launch-bound for small batches (seven launches of a few microseconds each), so it
// is captured once per (flags, timestep) into a CUDA graph and replayed with a single launch.  The first tick of a key
// runs eagerly (it also sets the function attributes), the second is captured, later ones replay.  Per-kernel event
// profiling (b2_profile_begin) and B2_NO_GRAPH=1 use the eager path.
int tick_dispatch(b2_batch* b, int flags) {
  CK(cudaSetDevice(b->device));
  if (b->prof_on || !b->use_graph) {
    const bool mine = prof_open(b);
    if (mine) prof_mark(b, SLOT_HW_WRITE);
    const int rc = tick_eager(b, flags);
    if (mine) prof_close(b);
    return rc;
  }
  const uint32_t hb = 0;   // (the timestep is read from device memory: not part of the graph key)
  // a tick that is a single kernel gains nothing from a graph
  if ((flags & B2_TICK_HW) && b->chain_single && b->hw_identity && !(b->tick_flags & (1 << 30))) return tick_eager(b, flags);
  unsigned long long ph = 0;  // the exchange buffers are baked into the captured kernel arguments
  if (flags & B2_TICK_HW)
    for (const void* p : {(const void*)b->io_in[0], (const void*)b->io_in[1], (const void*)b->io_out[0], (const void*)b->io_out[1], (const void*)b->io_out[2]})
      ph = ph * 0x9E3779B97F4A7C15ull + (unsigned long long)(uintptr_t)p;
  const std::pair<unsigned long long, unsigned long long> key{((unsigned long long)(unsigned)(flags | (b->tick_flags & (1 << 30))) << 32) | hb, ph};
  if (b->graphs.size() > 32) drop_graphs(b);
  auto it = b->graphs.find(key);
  if (it == b->graphs.end()) {  // first use: eager
    b->graphs[key] = b2_batch::GraphEntry{};
    return tick_eager(b, flags);
  }
  b2_batch::GraphEntry& g = it->second;
  if (!g.exec) {  // second use: capture
    const long long l0 = b->launches;
    CK(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = tick_eager(b, flags);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(b->stream, &graph);
    if (rc < 0 || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      b->use_graph = false;  // capture is not possible here: stay on the eager path
      b->launches = l0;
      return tick_eager(b, flags);
    }
    g.kernels = (int)(b->launches - l0);
    b->launches = l0;
    const cudaError_t e2 = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) { g.exec = nullptr; b->use_graph = false; cudaGetLastError(); return tick_eager(b, flags); }
  }
  CK(cudaGraphLaunch(g.exec, b->stream));
  b->launches += g.kernels;
  return 0;
}

void drop_graphs(b2_batch* b) {
  for (auto& kv : b->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  b->graphs.clear();
}

// ---- host <-> device field transfer with layout / precision conversion ----
template <typename H, typename D>
__global__ void k_scatter_field(D* dev, const H* host, int count, int nenvp, int env_lo, int n, int layout) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)count * n) return;
  const int i = (int)(idx / n), e = (int)(idx % n);  // consecutive threads -> consecutive environments
  const H v = layout == B2_NATIVE ? host[(long long)i * n + e] : host[(long long)e * count + i];
  dev[(long long)i * nenvp + env_lo + e] = (D)v;
}
template <typename H, typename D>
__global__ void k_gather_field(const D* dev, H* host, int count, int nenvp, int env_lo, int n, int layout) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)count * n) return;
  const int i = (int)(idx / n), e = (int)(idx % n);
  const H v = (H)dev[(long long)i * nenvp + env_lo + e];
  if (layout == B2_NATIVE) host[(long long)i * n + e] = v; else host[(long long)e * count + i] = v;
}

int ensure_stage(b2_batch* b, size_t bytes) {
  if (bytes <= b->stage_bytes) return 0;
  if (b->stage_dev) cudaFree(b->stage_dev);
  b->stage_dev = nullptr;
  b->stage_bytes = 0;
  CK(cudaMalloc(&b->stage_dev, bytes));
  b->stage_bytes = bytes;
  return 0;
}

template <typename H>
int set_field(b2_batch* b, const char* name, const H* host, int lo, int hi, int layout) {
  if (!b) return fail("null batch");
  CK(cudaSetDevice(b->device));
  auto it = b->fields.find(name);
  if (it == b->fields.end()) return fail(std::string("unknown field '") + name + "'");
  const Field& f = it->second;
  if (lo < 0 || hi > b->nenv || lo >= hi) return fail("bad environment range");
  if (f.kind != 0) return fail(std::string("field '") + name + "' is an int field");
  const int n = hi - lo;
  const size_t bytes = (size_t)f.count * n * sizeof(H);
  if (bytes == 0) return f.count;
  if (ensure_stage(b, bytes) < 0) return -1;
  CK(cudaMemcpyAsync(b->stage_dev, host, bytes, cudaMemcpyHostToDevice, b->stream));
  const long long tot = (long long)f.count * n;
  const int th = 256, bl = (int)((tot + th - 1) / th);
  if (b->prec == 8) k_scatter_field<H, double><<<bl, th, 0, b->stream>>>((double*)f.ptr, (const H*)b->stage_dev, f.count, b->nenvp, lo, n, layout);
  else k_scatter_field<H, float><<<bl, th, 0, b->stream>>>((float*)f.ptr, (const H*)b->stage_dev, f.count, b->nenvp, lo, n, layout);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(b->stream));
  if (!std::strcmp(name, "xfrc_applied")) b->tick_flags |= (1 << 30);
  return f.count;
}

// legacy dense views (mjData.efc_J is [nefc][nv], efc_AR is [nefc][njmax]) are expanded on demand from the compact rows
template <typename T>
int expand_dense(b2_batch* b, const char* name) {
  if (b->fused) return fail(std::string("field '") + name + "' does not exist for a model without constraints");
  KArgs<T> a = make_args<T>(b, 0);
  const long long njmax = b->hdr.njmax, nv = b->hdr.nv;
  auto ensure = [&](const char* n, long long count) -> T* {
    auto it = b->fields.find(n);
    if (it == b->fields.end()) { if (alloc_field(b, n, count, 0, nullptr) < 0) return nullptr; it = b->fields.find(n); }
    return (T*)it->second.ptr;
  };
  T* Jd = ensure("efc_J_dense", njmax * nv);
  if (!Jd) return -1;
  const int th = 128, bl = (b->nenvp + th - 1) / th;
  k_expand_rows<T><<<bl, th, 0, b->stream>>>(a, 0, Jd);
  if (!std::strcmp(name, "efc_AR") || !std::strcmp(name, "efc_B_dense")) {
    T* Bd = ensure("efc_B_dense", njmax * nv);
    if (!Bd) return -1;
    k_expand_rows<T><<<bl, th, 0, b->stream>>>(a, 1, Bd);
    if (!std::strcmp(name, "efc_AR")) {
      T* AR = ensure("efc_AR", njmax * njmax);
      if (!AR) return -1;
      k_dense_AR<T><<<bl, th, 0, b->stream>>>(a, Jd, Bd, AR);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

template <typename H>
int get_field(b2_batch* b, const char* name_in, H* host, int lo, int hi, int layout, int want_kind) {
  if (!b) return fail("null batch");
  CK(cudaSetDevice(b->device));
  const char* name = name_in;
  if (!std::strcmp(name, "efc_J") || !std::strcmp(name, "efc_AR") || !std::strcmp(name, "efc_B_dense")) {
    if ((b->prec == 8 ? expand_dense<double>(b, name) : expand_dense<float>(b, name)) < 0) return -1;
    if (!std::strcmp(name, "efc_J")) name = "efc_J_dense";
  }
  auto it = b->fields.find(name);
  if (it == b->fields.end()) return fail(std::string("unknown field '") + name + "'");
  const Field& f = it->second;
  if (lo < 0 || hi > b->nenv || lo >= hi) return fail("bad environment range");
  if (f.kind != want_kind) return fail(std::string("field '") + name + "' has a different element kind");
  const int n = hi - lo;
  const size_t bytes = (size_t)f.count * n * sizeof(H);
  if (bytes == 0) return f.count;
  if (ensure_stage(b, bytes) < 0) return -1;
  const long long tot = (long long)f.count * n;
  const int th = 256, bl = (int)((tot + th - 1) / th);
  if (f.kind == 1) k_gather_field<H, int><<<bl, th, 0, b->stream>>>((const int*)f.ptr, (H*)b->stage_dev, f.count, b->nenvp, lo, n, layout);
  else if (b->prec == 8) k_gather_field<H, double><<<bl, th, 0, b->stream>>>((const double*)f.ptr, (H*)b->stage_dev, f.count, b->nenvp, lo, n, layout);
  else k_gather_field<H, float><<<bl, th, 0, b->stream>>>((const float*)f.ptr, (H*)b->stage_dev, f.count, b->nenvp, lo, n, layout);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(host, b->stage_dev, bytes, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return f.count;
}

template <typename D>
__global__ void k_reset(D* qpos, D* qvel, D* qacc, D* qws, D* qapp, D* time, const uint32_t* model, int nenvp, int lo, int hi) {
  const DModel* h = reinterpret_cast<const DModel*>(model);
  const int e = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= hi) return;
  const D* q0 = reinterpret_cast<const D*>(model + h->o_qpos0);
  for (int i = 0; i < h->nq; i++) qpos[(long long)i * nenvp + e] = q0[i];
  for (int i = 0; i < h->nv; i++) {
    qvel[(long long)i * nenvp + e] = 0; qacc[(long long)i * nenvp + e] = 0; qws[(long long)i * nenvp + e] = 0; qapp[(long long)i * nenvp + e] = 0;
  }
  time[e] = 0;
}

// ---- object slots: an inactive slot rests at its parking place (x = 3 slot, y = 0, z = 1000 + 3 slot), far from the
// scene and from the other parked slots, and is put back there after every tick ----
template <typename D>
__device__ __forceinline__ void park_slot(D* qpos, D* qvel, D* qacc, D* qws, int qa, int da, int slot, long long S, int e) {
  const D pose[7] = {D(3 * slot), D(0), D(1000 + 3 * slot), D(1), D(0), D(0), D(0)};
  for (int k = 0; k < 7; k++) qpos[(long long)(qa + k) * S + e] = pose[k];
  for (int k = 0; k < 6; k++) { qvel[(long long)(da + k) * S + e] = 0; qacc[(long long)(da + k) * S + e] = 0; qws[(long long)(da + k) * S + e] = 0; }
}
template <typename D>
__global__ void k_hold_slots(D* qpos, D* qvel, D* qacc, D* qws, const unsigned char* active, const int* qadr, const int* dadr, int nslot, int nenv, int nenvp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nslot * nenv) return;
  const int s = (int)(idx / nenv), e = (int)(idx % nenv);
  if (!active[(long long)s * nenvp + e]) park_slot(qpos, qvel, qacc, qws, qadr[s], dadr[s], s, nenvp, e);
}
// apply n spawn (pose != nullptr) or destroy requests, one thread per request.  "In order" semantics are established on
// the host: slot_requests() keeps only the LAST request per (environment, slot) of a call (env[i] < 0 marks a dropped
// duplicate), so the surviving requests touch disjoint state and run in parallel deterministically.
template <typename D>
__global__ void k_slot_requests(D* qpos, D* qvel, D* qacc, D* qws, unsigned char* active, const int* qadr, const int* dadr, int nenvp,
                                int n, const int* env, const int* slot, const float* pose7, const float* twist6) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = env[i], s = slot[i];
    if (e < 0) continue;   // superseded by a later request of the same call
    if (pose7) {
      const int qa = qadr[s], da = dadr[s];
      D q[4] = {(D)pose7[7 * i + 3], (D)pose7[7 * i + 4], (D)pose7[7 * i + 5], (D)pose7[7 * i + 6]};
      D nrm = sqrt((double)(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]));
      if (!(nrm > 1e-12)) { q[0] = 1; q[1] = q[2] = q[3] = 0; nrm = 1; }
      for (int k = 0; k < 3; k++) qpos[(long long)(qa + k) * nenvp + e] = (D)pose7[7 * i + k];
      for (int k = 0; k < 4; k++) qpos[(long long)(qa + 3 + k) * nenvp + e] = q[k] / nrm;
      for (int k = 0; k < 6; k++) {
        qvel[(long long)(da + k) * nenvp + e] = twist6 ? (D)twist6[6 * i + k] : D(0);
        qacc[(long long)(da + k) * nenvp + e] = 0; qws[(long long)(da + k) * nenvp + e] = 0;
      }
      active[(long long)s * nenvp + e] = 1;
    } else {
      active[(long long)s * nenvp + e] = 0;
      park_slot(qpos, qvel, qacc, qws, qadr[s], dadr[s], s, nenvp, e);
    }
  }
}

// observation pack for the per-tick all-gather: dst[(i) * nenv + e] = (i < nq ? qpos[i] : qvel[i - nq]) of environment e
template <typename D>
__global__ void k_pack_obs(const D* __restrict__ qpos, const D* __restrict__ qvel, float* __restrict__ dst, int nq, int nv, int nenv, int nenvp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(nq + nv) * nenv) return;
  const int i = (int)(idx / nenv), e = (int)(idx % nenv);
  dst[idx] = (float)(i < nq ? qpos[(long long)i * nenvp + e] : qvel[(long long)(i - nq) * nenvp + e]);
}

// MjHWInterface::write (src/mujoco_sim/mj_hw_interface.cpp:73-91) for every environment
template <typename D>
__global__ void k_hw_write(D* ddq, D* dq, const float* vel_cmd, const float* eff_cmd, const int* dadr, const int* ctl, int nhw,
                           int nenv, int nenvp, const D* qpos, const D* qvel, const int* qadr, const float* kp, const float* kd,
                           float* pos_out, float* vel_out, int controller) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nhw * nenv) return;
  const int j = (int)(idx / nenv), e = (int)(idx % nenv);
  const int d = dadr[j];
  const D qj = qpos[(long long)qadr[j] * nenvp + e], vj = qvel[(long long)d * nenvp + e];
  if (ctl[j]) {
    const float v = vel_cmd[idx];
    if (fabsf(v) > 1e-15f) dq[(long long)d * nenvp + e] = (D)v;
    else {
      D cmd = (D)eff_cmd[idx];
      if (kp) cmd = (D)kp[j] * (cmd - qj) - (D)kd[j] * vj;  // PD stage (b2_set_pd)
      ddq[(long long)d * nenvp + e] = cmd;
    }
  }
  // MjHWInterface::read runs between mj_step1 and mj_step2 (src/mj_main.cpp:91-108): it sees this tick's starting
  // position and the velocity AFTER the controller's override by a pending velocity command (mj_sim.cpp:1066-1070)
  if (pos_out) {
    const D dqv = dq[(long long)d * nenvp + e];
    pos_out[idx] = (float)qj;
    vel_out[idx] = (float)((controller && (dqv < 0 ? -dqv : dqv) > D(1e-15)) ? dqv : vj);
  }
}
// MjHWInterface::read gathers (src/mujoco_sim/mj_hw_interface.cpp:62-70)
template <typename D>
__global__ void k_hw_read(const D* qpos, const D* qvel, const D* qfrc_inverse, float* pos, float* vel, float* eff, const int* qadr,
                          const int* dadr, int nhw, int nenv, int nenvp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nhw * nenv) return;
  const int j = (int)(idx / nenv), e = (int)(idx % nenv);
  if (pos) {   // B2_TICK_READ_POST; otherwise k_hw_write has gathered them at the start of the tick
    pos[idx] = (float)qpos[(long long)qadr[j] * nenvp + e];
    vel[idx] = (float)qvel[(long long)dadr[j] * nenvp + e];
  }
  eff[idx] = (float)qfrc_inverse[(long long)dadr[j] * nenvp + e];
}

}  // namespace

extern "C" {

const char* b2_last_error(void) { return b2::g_err.c_str(); }

int b2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

b2_batch* b2_create(const mjModel* m, int nenv, int device, int precision) {
  if (!m) { fail("b2_create: null model"); return nullptr; }
  if (nenv < 1) { fail("b2_create: nenv must be >= 1"); return nullptr; }
  const bool export_stages = (precision & B2_EXPORT_STAGES) != 0;
  precision &= ~B2_EXPORT_STAGES;
  if (precision != B2_F32 && precision != B2_F64) { fail("b2_create: precision must be B2_F32 or B2_F64"); return nullptr; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fail(std::string("b2_create: no usable CUDA device (") + cudaGetErrorString(e) + "); this engine has no CPU fallback");
    cudaGetLastError();
    return nullptr;
  }
  if (device < 0 || device >= ndev) { fail("b2_create: bad device index"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { fail("b2_create: cudaSetDevice failed"); return nullptr; }
  auto* b = new b2_batch();
  b->m = m; b->nenv = nenv; b->nenvp = (nenv + 127) / 128 * 128; b->device = device; b->prec = precision;
  b->h = m->opt.timestep;
  b->opt_iterations = m->opt.iterations; b->opt_tolerance = m->opt.tolerance; b->opt_disableflags = m->opt.disableflags;
  b->controlled.assign(m->nv, 0);
  b->export_stages = export_stages;
  b->use_graph = !getenv("B2_NO_GRAPH");
  if (getenv("B2_SUBBATCH")) b->nsub = std::max(1, std::min(8, atoi(getenv("B2_SUBBATCH"))));
  if (getenv("B2_SUBBATCH_MIN")) b->sub_min_envs = std::max(128, atoi(getenv("B2_SUBBATCH_MIN")));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) b->nsm = prop.multiProcessorCount;
  auto bail = [&](const char* what) -> b2_batch* {
    if (b2::g_err.empty()) fail(what);
    b2_destroy(b);
    return nullptr;
  };
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate failed");
  {
    if (cudaMalloc(&b->h_dev, 16) != cudaSuccess) return bail("cudaMalloc(h_dev) failed");
    struct { double d; float f; float pad; } hv{b->h, (float)b->h, 0.f};
    if (cudaMemcpyAsync(b->h_dev, &hv, 16, cudaMemcpyHostToDevice, b->stream) != cudaSuccess) return bail("timestep upload failed");
  }
  if (upload_model(b) < 0) return bail("upload_model failed");

  const int nq = m->nq, nv = m->nv, nb = m->nbody, ng = m->ngeom, nM = m->nM;
  // does the model have any constraint source?  If not the smooth kernel integrates by itself (one launch per tick).
  bool any_pair = false;
  for (int p = 0; p < m->npair; p++)
    any_pair |= pair_supported(m->geom_type[m->pair_geom1[p]], m->geom_type[m->pair_geom2[p]]);
  bool any_lim = false, any_fl = false;
  for (int j = 0; j < m->njnt; j++) any_lim |= m->jnt_limited[j] != 0;
  for (int i = 0; i < nv; i++) any_fl |= m->dof_frictionloss[i] > 0;
  b->fused = (m->opt.disableflags & mjDSBL_CONSTRAINT) || !(any_pair || any_lim || any_fl || m->neq > 0);
  bool ball_lim = false;
  for (int j = 0; j < m->njnt; j++) ball_lim |= m->jnt_limited[j] && m->jnt_type[j] == mjJNT_BALL;
  b->fusable = !b->fused && !any_pair && !any_fl && m->neq == 0 && !ball_lim;
  // serial chain of scalar joints?
  {
    bool chain = m->nbody >= 2 && m->nv == m->nbody - 1 && m->njnt == m->nbody - 1 && m->nq == m->nv && m->nmocap == 0;
    for (int i = 1; i < m->nbody && chain; i++)
      chain = m->body_parentid[i] == i - 1 && m->body_jntnum[i] == 1 && m->body_jntadr[i] == i - 1 && m->body_dofadr[i] == i - 1 &&
              (m->jnt_type[i - 1] == mjJNT_HINGE || m->jnt_type[i - 1] == mjJNT_SLIDE);
    const int n = m->nbody - 1;
    const bool have = have_chain_kernel(n, precision);
    b->chain_n = (chain && have && !getenv("B2_NO_CHAIN")) ? n : 0;
    // large batches: 128-thread CTAs, 2 per SM at 255 registers (variant 1) beat 4 per SM at 128 registers
    // (profiles/r01_chain_variants.txt)
    b->chain_single = b->chain_n > 0 && (b->fusable || b->fused) && !b->export_stages && !getenv("B2_NO_CHAIN1");
    b->chain_variant = getenv("B2_CHAIN_VARIANT") ? atoi(getenv("B2_CHAIN_VARIANT")) : 1;
    // below ~16 k environments one thread per environment leaves the SMs latency-bound on a single instruction stream:
    // an 8-lane team per environment shortens the dependent chain (k_chain_team.cuh)
    b->chain_team = b->chain_single && (n == 6 || (n == 7 && precision == 4) || n == 7) &&
                    (getenv("B2_CHAIN_TEAM") ? atoi(getenv("B2_CHAIN_TEAM")) != 0 : b->nenvp <= 16384);
  }

  b->epl = 2; b->wp = 16;  // (legacy fields of the row-slab solver; unused)
  {
    // block records (k_constraint.cuh): slab capacity per environment and the largest single record
    int nbmax = 1;
    for (int g = 0; g < m->ngeom; g++) nbmax = std::max(nbmax, m->geom_condim[g]);
    nbmax = std::min(nbmax, 6);
    b->block_capw = block_capacity(b->hdr.njmax, b->hdr.wmax);
    b->block_npar = block_max_params(nbmax);
    b->rec_max = b->block_npar + 2 * ((b->hdr.wmax + 3) & ~3);
    // team width of the solver: the narrowest of 8 / 16 / 32 lanes that holds the widest compact row in two elements per
    // lane (the exact-shape visit's condition); PR2-sized trees (49 dofs) get a whole warp per environment
    const int wm = b->hdr.wmax;
    b->pgs_lanes = getenv("B2_PGS_LANES") ? atoi(getenv("B2_PGS_LANES")) : (wm <= 16 ? 8 : (wm <= 32 ? 16 : 32));
    if (b->pgs_lanes != 4 && b->pgs_lanes != 16 && b->pgs_lanes != 32) b->pgs_lanes = 8;
    // models made of several small kinematic trees (arm + free objects, object slots): island-parallel solver
    b->isl_cap = (b->hdr.ntree >= 2 && b->hdr.ntree <= 32 && wm <= 16 && !getenv("B2_NO_ISLANDS")) ? b->hdr.ntree : 0;
    b->pgs_isl = getenv("B2_PGS_ISL") ? atoi(getenv("B2_PGS_ISL")) : 8;
    if (b->pgs_isl != 4 && b->pgs_isl != 16) b->pgs_isl = 8;
  }
  struct Spec { const char* name; long long count; int kind; };
  std::vector<Spec> specs = {
      {"qpos", nq, 0}, {"qvel", nv, 0}, {"qacc", nv, 0}, {"qacc_warmstart", nv, 0}, {"qfrc_applied", nv, 0},
      {"xfrc_applied", 6 * nb, 0}, {"mocap_pos", 3 * std::max(1, m->nmocap), 0}, {"mocap_quat", 4 * std::max(1, m->nmocap), 0},
      {"ddq", nv, 0}, {"dq", nv, 0}, {"odom_vels", 6, 0}, {"time", 1, 0}, {"qfrc_bias", nv, 0}, {"qfrc_inverse", nv, 0},
      {"xpos", 3 * nb, 0}, {"xquat", 4 * nb, 0}, {"_pending", 1, 1}, {"status", 1, 1}, {"solver_iter", 1, 1}, {"ncon", 1, 1}, {"nefc", 1, 1}};
  if (!b->fused || b->export_stages) {
    std::vector<Spec> more = {
        {"xmat", 9 * nb, 0}, {"geom_xpos", 3 * std::max(1, ng), 0}, {"geom_xmat", 9 * std::max(1, ng), 0}, {"subtree_com", 3 * nb, 0},
        {"cdof", 6 * nv, 0}, {"qM", nM, 0}, {"qLD", nM, 0}, {"qLDiagInv", nv, 0}, {"qfrc_passive", nv, 0}, {"qfrc_smooth", nv, 0},
        {"qacc_smooth", nv, 0}, {"qfrc_constraint", nv, 0}};
    specs.insert(specs.end(), more.begin(), more.end());
  }
  if (!b->fused) {
    const long long njmax = b->hdr.njmax, ncm = b->hdr.nconmax;
    std::vector<Spec> more = {
        {"contact", CF_NFLOAT * ncm, 0}, {"contact_int", CI_NINT * ncm, 1},
        {"efc_type", njmax, 1}, {"efc_id", njmax, 1}, {"efc_tree", 2 * njmax, 1}, {"efc_J", njmax * b->hdr.wmax, 0}, {"efc_pos", njmax, 0}, {"efc_margin", njmax, 0},
        {"efc_frictionloss", njmax, 0}, {"efc_diagApprox", njmax, 0}, {"efc_R", njmax, 0}, {"efc_D", njmax, 0}, {"efc_KBI", 3 * njmax, 0},
        {"efc_vel", njmax, 0}, {"efc_aref", njmax, 0}, {"efc_b", njmax, 0}, {"efc_force", njmax, 0}, {"efc_finv", njmax, 0}, {"efc_ARdiag", njmax, 0},
        {"efc_blocks", b->block_capw, 0}, {"efc_nwords", 1, 1}, {"env_order", 1, 1}, {"blk_row0", njmax, 1}, {"blk_off", njmax, 1}, {"nblk", 1, 1}, {"_maxblk", 1, 1}};
    // tensor-core projection: one tree whose compact row is every dof, fp32, nv <= 56 (the GEMM's padded K), opt-in
    b->tc_rows = 0;
    if (getenv("B2_TC_PROJECT") && atoi(getenv("B2_TC_PROJECT")) > 0 && precision == 4 && b->hdr.ntree == 1 && b->hdr.wmax == nv && nv <= tc::TK &&
        (size_t)b->rec_max * 129 * b->prec > 56 * 1024) {
      b->tc_rows = (int)((njmax + tc::TM - 1) / tc::TM) * tc::TM;
      b->tc_passes = atoi(getenv("B2_TC_PROJECT")) == 1 ? 3 : 1;   // B2_TC_PROJECT=1: 3xTF32, =2: single-pass TF32 (accuracy experiment)
      more.push_back({"efc_Jem", (long long)b->tc_rows * 64, 0}); more.push_back({"efc_Bem", (long long)b->tc_rows * 64, 0}); more.push_back({"minv_em", 64 * 64, 0});
    }
    if (b->isl_cap) { more.push_back({"isl_off", b->isl_cap, 1}); more.push_back({"isl_end", b->isl_cap, 1}); more.push_back({"nisl", 1, 1}); }
    // wide trees (the criterion of make_block == 32): M^-1 J^T of every row is produced by k_solve_rows into efc_B
    if ((size_t)b->rec_max * 129 * b->prec > 56 * 1024 && !getenv("B2_NO_SOLVE_ROWS")) more.push_back({"efc_B", njmax * b->hdr.wmax, 0});
    specs.insert(specs.end(), more.begin(), more.end());
  }
  for (auto& s : specs)
    if (alloc_field(b, s.name, s.count, s.kind, nullptr) < 0) return bail("alloc failed");

  // shared-memory budget of the smooth kernel: 16 B barrier + model blob + workspace[ws_slots][BLOCK]
  b->blob_smem = 16 + (size_t)b->hdr.nwords * 4;
  const size_t budget = 200 * 1024;
  // tree-parallel form of k_smooth / k_integrate: lanes per environment from the number of kinematic trees (an arm and four
  // free props: 8 lanes, one tree each); single-tree models and register-resident chains stay one thread per environment
  b->tree_lanes = 1;
  if (b->chain_n == 0 && !getenv("B2_NO_TREE_LANES")) {
    // lanes: enough to bring the heaviest lane down to about the heaviest tree (same cost proxy as the assignment in
    // pack_model), a power of two <= 8
    const int nt = b->hdr.ntree;
    std::vector<long long> w(std::max(1, nt), 0);
    long long tot = 0, mx = 1;
    {
      std::vector<int> root_tree(m->nbody, -1);
      int ntr = 0;
      for (int d = 0; d < m->nv; d++) { const int root = m->body_rootid[m->dof_bodyid[d]]; if (root_tree[root] < 0) root_tree[root] = ntr++; w[root_tree[root]] += 1; }
      for (int i = 1; i < m->nbody; i++) if (lastdof_of(m, i) >= 0) w[root_tree[m->body_rootid[i]]] += 2;
      for (int t = 0; t < nt; t++) { tot += w[t]; mx = std::max(mx, w[t]); }
    }
    const long long want = (tot + mx - 1) / mx;
    // measured (B200, tools/gpurun_r2l.sh): the smooth kernel is register-bound (three 128-thread CTAs per SM), so more
    // lanes mean more waves; 2 lanes for an arm + a few props (C3: smooth + integrate 0.49 -> 0.38 ms), 4 for a field of
    // free bodies (C5: 1.95 -> 0.84 ms), never 8
    b->tree_lanes = want >= 8 ? 4 : (want >= 2 ? 2 : 1);
    if (getenv("B2_TREE_LANES")) { const int v = atoi(getenv("B2_TREE_LANES")); if (v == 1 || v == 2 || v == 4 || v == 8) b->tree_lanes = v; }
  }
  // (a kinematic tree standing on a static pedestal body would read a frame another lane computes: one lane then)
  for (int i = 1; i < m->nbody && b->tree_lanes > 1; i++) {
    const int pa = m->body_parentid[i];
    if (pa > 0 && lastdof_of(m, i) >= 0 && lastdof_of(m, pa) < 0) b->tree_lanes = 1;
  }
  if (b->tree_lanes > 1 && upload_model(b) < 0) return bail("upload_model failed");   // the blob carries the per-lane item lists
  const int TL = b->tree_lanes;
  int block = 0;
  if (TL > 1) {
    // 128 threads = 128 / TL environments per CTA: the workspace shrinks with it
    const size_t need = b->blob_smem + (size_t)b->hdr.ws_slots * (128 / TL) * precision;
    if (need <= budget) block = 128;
  } else
  // a shared-memory workspace pays off only when at least two warps fit on an SM; below that the L2-resident HBM
  // workspace with four warps per SM is faster (measured on C3: 0.63 ms vs 0.95 ms, profiles/r01_ncu_c3_summary.txt)
  for (int cand : {128, 64}) {
    const size_t need = b->blob_smem + (size_t)b->hdr.ws_slots * cand * precision;
    if (need > budget) continue;
    if (!block) block = cand;
    // prefer a smaller CTA when the batch cannot fill the SMs with the larger one
    if (b->nenvp / block < b->nsm && cand < block) block = cand;
  }
  if (b->chain_n > 0) {
    // register-resident chain kernel: shared memory holds the model blob only
    b->smooth_block = (precision == 8 || b->nenvp / 128 < 2 * b->nsm) ? 32 : 128;
    b->smooth_smem = b->blob_smem;
    b->ws_global = false;
  } else if (block) {
    b->smooth_block = block;
    b->smooth_smem = b->blob_smem + (size_t)b->hdr.ws_slots * (block / TL) * precision;
    b->ws_global = false;
  } else {
    b->smooth_block = 128;
    b->smooth_smem = b->blob_smem;
    b->ws_global = true;
    // factor scratch in shared memory (B2F_LD_SMEM): (nM + 3 nv) words per environment (factor, 1 / D, two vectors), the widest CTA that leaves two CTAs per SM
    b->ld_smem = 0;
    if (!getenv("B2_NO_LD_SMEM"))
      for (int cand : {128, 64, 32}) {
        if (TL > 1 && cand != 128) break;
        // (the fourth vector — the second product of the fused M x pass — only when it does not cost the CTA width)
        const size_t need3 = b->blob_smem + (size_t)(b->hdr.nM + 3 * b->hdr.nv) * (cand / TL) * precision;
        const size_t need2 = b->blob_smem + (size_t)(b->hdr.nM + 2 * b->hdr.nv) * (cand / TL) * precision;
        if (need2 > 110 * 1024) continue;
        b->ld_extra = need3 <= 110 * 1024 ? 1 : 0;
        const size_t need = b->ld_extra ? need3 : need2;
        b->smooth_block = cand; b->ld_smem = need - b->blob_smem; b->smooth_smem = need;
        break;
      }
    if (alloc_field(b, "_ws", b->hdr.ws_slots, 0, nullptr) < 0) return bail("alloc failed");
  }
  if (configure_constraint_kernels(b) < 0) return bail("constraint kernel configuration failed");
  if (b2_reset(b, 0, nenv) < 0) return bail("reset failed");
  // padded environments also start from qpos0 so that they stay finite
  if (b->nenvp > nenv) {
    const int th = 128, n = b->nenvp - nenv;
    if (precision == 8)
      k_reset<double><<<(n + th - 1) / th, th, 0, b->stream>>>((double*)b->fields["qpos"].ptr, (double*)b->fields["qvel"].ptr, (double*)b->fields["qacc"].ptr,
          (double*)b->fields["qacc_warmstart"].ptr, (double*)b->fields["qfrc_applied"].ptr, (double*)b->fields["time"].ptr, b->blob_dev, b->nenvp, nenv, b->nenvp);
    else
      k_reset<float><<<(n + th - 1) / th, th, 0, b->stream>>>((float*)b->fields["qpos"].ptr, (float*)b->fields["qvel"].ptr, (float*)b->fields["qacc"].ptr,
          (float*)b->fields["qacc_warmstart"].ptr, (float*)b->fields["qfrc_applied"].ptr, (float*)b->fields["time"].ptr, b->blob_dev, b->nenvp, nenv, b->nenvp);
  }
  // mocap bodies start at their authored pose
  if (m->nmocap > 0) {
    std::vector<double> mp(3 * (size_t)m->nmocap), mq(4 * (size_t)m->nmocap);
    for (int i = 0; i < m->nbody; i++) {
      const int id = m->body_mocapid[i];
      if (id < 0) continue;
      for (int k = 0; k < 3; k++) mp[3 * id + k] = m->body_pos[3 * i + k];
      for (int k = 0; k < 4; k++) mq[4 * id + k] = m->body_quat[4 * i + k];
    }
    std::vector<double> hp((size_t)nenv * mp.size()), hq((size_t)nenv * mq.size());
    for (int e2 = 0; e2 < nenv; e2++) {
      std::copy(mp.begin(), mp.end(), hp.begin() + (size_t)e2 * mp.size());
      std::copy(mq.begin(), mq.end(), hq.begin() + (size_t)e2 * mq.size());
    }
    if (set_field<double>(b, "mocap_pos", hp.data(), 0, nenv, B2_ENV_MAJOR) < 0) return bail("mocap init failed");
    if (set_field<double>(b, "mocap_quat", hq.data(), 0, nenv, B2_ENV_MAJOR) < 0) return bail("mocap init failed");
  }
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) return bail("stream sync failed");
  b->tick_flags = B2_TICK_INTEGRATE;
  return b;
}

void b2_destroy(b2_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (void* p : b->allocs) cudaFree(p);
  if (b->blob_dev) cudaFree(b->blob_dev);
  if (b->stage_dev) cudaFree(b->stage_dev);
  if (b->hw_qadr) cudaFree(b->hw_qadr);
  if (b->hw_dadr) cudaFree(b->hw_dadr);
  if (b->hw_ctl) cudaFree(b->hw_ctl);
  if (b->hw_buf) cudaFree(b->hw_buf);
  if (b->hw_kp) cudaFree(b->hw_kp);
  if (b->hw_kd) cudaFree(b->hw_kd);
  if (b->slot_qadr) cudaFree(b->slot_qadr);
  if (b->slot_dadr) cudaFree(b->slot_dadr);
  if (b->slot_active) cudaFree(b->slot_active);
  if (b->flush_buf) cudaFree(b->flush_buf);
  if (b->h_dev) cudaFree(b->h_dev);
  for (cudaStream_t st : b->sub_stream) cudaStreamDestroy(st);
  for (cudaEvent_t ev : b->sub_join) cudaEventDestroy(ev);
  if (b->sub_fork) cudaEventDestroy(b->sub_fork);
  for (size_t p = 0; p < b->obs_peers_host.size(); p++)
    if (b->obs_peer_ipc[p] && b->obs_peers_host[p]) cudaIpcCloseMemHandle(b->obs_peers_host[p]);
  if (b->obs_peers_dev) cudaFree(b->obs_peers_dev);
  if (b->obs_buf) cudaFree(b->obs_buf);
  for (auto& kv : b->registered) cudaHostUnregister(const_cast<void*>(kv.first));
  for (cudaEvent_t e : b->prof_ev) cudaEventDestroy(e);
  drop_graphs(b);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

// which kernels a tick of this batch launches (for reports): "k_chain<N>", "k_smooth<ChainP<N>>+pipeline",
// "k_smooth<GenericP>" (no constraint source) or "k_smooth<GenericP>+pipeline"
const char* b2_path_name(const b2_batch* b) {
  static thread_local char buf[64];
  if (!b) return "";
  if (b->chain_single) std::snprintf(buf, sizeof(buf), b->chain_team ? "k_chain_team<%d>" : "k_chain<%d>", b->chain_n);
  else if (b->chain_n > 0) std::snprintf(buf, sizeof(buf), "k_smooth<ChainP<%d>>%s", b->chain_n, b->fused ? "" : "+pipeline");
  else std::snprintf(buf, sizeof(buf), "k_smooth<GenericP>%s", b->fused ? "" : "+pipeline");
  return buf;
}

int b2_nenv(const b2_batch* b) { return b ? b->nenv : -1; }
int b2_nenv_padded(const b2_batch* b) { return b ? b->nenvp : -1; }
int b2_precision(const b2_batch* b) { return b ? b->prec : -1; }

int b2_set_controlled(b2_batch* b, const unsigned char* mask) {
  if (!b || !mask) return fail("b2_set_controlled: null argument");
  CK(cudaSetDevice(b->device));
  b->controlled.assign(mask, mask + b->m->nv);
  if (upload_model(b) < 0) return -1;
  if (b->nhw > 0) {  // refresh the per-joint controlled flags
    std::vector<int> dadr(b->nhw), ctl(b->nhw);
    CK(cudaMemcpy(dadr.data(), b->hw_dadr, sizeof(int) * b->nhw, cudaMemcpyDeviceToHost));
    for (int j = 0; j < b->nhw; j++) ctl[j] = b->controlled[dadr[j]] ? 1 : 0;
    CK(cudaMemcpy(b->hw_ctl, ctl.data(), sizeof(int) * b->nhw, cudaMemcpyHostToDevice));
  }
  return 0;
}

int b2_set_odom(b2_batch* b, int nrobot, const int* dof, const int* qposadr) {
  if (!b || nrobot < 0 || (nrobot > 0 && (!dof || !qposadr))) return fail("b2_set_odom: bad argument");
  CK(cudaSetDevice(b->device));
  b->odom_dof.assign(dof, dof + 6 * nrobot);
  b->odom_qpos.assign(qposadr, qposadr + 3 * nrobot);
  // the blob grows: reallocate it
  if (b->blob_dev) { cudaFree(b->blob_dev); b->blob_dev = nullptr; }
  if (upload_model(b) < 0) return -1;
  // odom_vels holds 6 numbers per robot
  auto it = b->fields.find("odom_vels");
  if (it != b->fields.end() && it->second.count < 6 * nrobot) {
    if (alloc_field(b, "odom_vels", 6 * nrobot, 0, nullptr) < 0) return -1;
  }
  b->blob_smem = 16 + (size_t)b->hdr.nwords * 4;
  if (!b->ws_global) b->smooth_smem = b->blob_smem + (size_t)b->hdr.ws_slots * b->smooth_block * b->prec;
  else b->smooth_smem = b->blob_smem + b->ld_smem;
  if (configure_constraint_kernels(b) < 0) return -1;
  return 0;
}

int b2_set_timestep(b2_batch* b, double h) {
  if (!b || !(h > 0)) return fail("b2_set_timestep: bad argument");
  if (h == b->h && b->h_dev) return 0;
  b->h = h;
  // the kernels read the timestep from device memory (KArgs::hp), so captured CUDA graphs follow it: the reference
  // changes m->opt.timestep on every tick (src/mj_main.cpp:150-163) and a graph per value would never be reused
  CK(cudaSetDevice(b->device));
  if (!b->h_dev) CK(cudaMalloc(&b->h_dev, 16));
  struct { double d; float f; float pad; } hv{h, (float)h, 0.f};
  CK(cudaMemcpyAsync(b->h_dev, &hv, 16, cudaMemcpyHostToDevice, b->stream));   // pageable source: staged before the call returns
  return 0;
}

int b2_set_option(b2_batch* b, const char* name, double value) {
  if (!b || !name) return fail("b2_set_option: null argument");
  CK(cudaSetDevice(b->device));
  if (!std::strcmp(name, "subbatches") || !std::strcmp(name, "subbatch_min")) {
    // scheduling only (run_tick): how many windows of the batch run the pipeline side by side, and their smallest size
    if (value < 1) return fail("b2_set_option: subbatches / subbatch_min must be >= 1");
    CK(cudaStreamSynchronize(b->stream));
    drop_graphs(b);
    if (name[8] == 'e') b->nsub = std::min(8, (int)value); else b->sub_min_envs = (int)value;
    return 0;
  }
  if (!std::strcmp(name, "iterations")) b->opt_iterations = (int)value;
  else if (!std::strcmp(name, "tolerance")) b->opt_tolerance = value;
  else if (!std::strcmp(name, "disableflags")) b->opt_disableflags = (int)value;
  else return fail(std::string("b2_set_option: unknown option '") + name + "'");
  return upload_model(b);
}

int b2_field_size(const b2_batch* b, const char* field) {
  if (!b || !field) return fail("b2_field_size: null argument");
  if (!b->fused && !std::strcmp(field, "efc_J")) return b->hdr.njmax * b->hdr.nv;            // dense legacy view
  if (!b->fused && !std::strcmp(field, "efc_AR")) return b->hdr.njmax * b->hdr.njmax;
  auto it = b->fields.find(field);
  if (it == b->fields.end()) return fail(std::string("unknown field '") + field + "'");
  return it->second.count;
}
int b2_set_field_f32(b2_batch* b, const char* f, const float* h, int lo, int hi, int layout) { return set_field<float>(b, f, h, lo, hi, layout); }
int b2_set_field_f64(b2_batch* b, const char* f, const double* h, int lo, int hi, int layout) { return set_field<double>(b, f, h, lo, hi, layout); }
int b2_get_field_f32(b2_batch* b, const char* f, float* h, int lo, int hi, int layout) { return get_field<float>(b, f, h, lo, hi, layout, 0); }
int b2_get_field_f64(b2_batch* b, const char* f, double* h, int lo, int hi, int layout) { return get_field<double>(b, f, h, lo, hi, layout, 0); }
int b2_get_field_i32(b2_batch* b, const char* f, int* h, int lo, int hi, int layout) { return get_field<int>(b, f, h, lo, hi, layout, 1); }

void* b2_device_ptr(b2_batch* b, const char* field) {
  if (!b || !field) { fail("b2_device_ptr: null argument"); return nullptr; }
  auto it = b->fields.find(field);
  if (it == b->fields.end()) { fail(std::string("unknown field '") + field + "'"); return nullptr; }
  return it->second.ptr;
}

int b2_reset(b2_batch* b, int lo, int hi) {
  if (!b || lo < 0 || hi > b->nenv || lo >= hi) return fail("b2_reset: bad range");
  CK(cudaSetDevice(b->device));
  const int th = 128, n = hi - lo;
  auto P = [&](const char* nm) { return b->fields[nm].ptr; };
  if (b->prec == 8)
    k_reset<double><<<(n + th - 1) / th, th, 0, b->stream>>>((double*)P("qpos"), (double*)P("qvel"), (double*)P("qacc"), (double*)P("qacc_warmstart"),
                                                            (double*)P("qfrc_applied"), (double*)P("time"), b->blob_dev, b->nenvp, lo, hi);
  else
    k_reset<float><<<(n + th - 1) / th, th, 0, b->stream>>>((float*)P("qpos"), (float*)P("qvel"), (float*)P("qacc"), (float*)P("qacc_warmstart"),
                                                           (float*)P("qfrc_applied"), (float*)P("time"), b->blob_dev, b->nenvp, lo, hi);
  CK(cudaGetLastError());
  return 0;
}

int b2_tick(b2_batch* b, int flags) {
  if (!b) return fail("b2_tick: null batch");
  return tick_dispatch(b, flags);
}
int b2_set_tick_flags(b2_batch* b, int flags) {
  if (!b) return fail("null batch");
  b->tick_flags = (b->tick_flags & (1 << 30)) | (flags & ~(1 << 30));
  return 0;
}
int b2_step(b2_batch* b, int nsteps) {
  if (!b) return fail("b2_step: null batch");
  for (int s = 0; s < nsteps; s++)
    if (tick_dispatch(b, (b->tick_flags & ~(1 << 30)) | B2_TICK_INTEGRATE) < 0) return -1;
  return 0;
}
int b2_forward(b2_batch* b) {
  if (!b) return fail("b2_forward: null batch");
  return tick_dispatch(b, 0);
}
int b2_sync(b2_batch* b) {
  if (!b) return fail("b2_sync: null batch");
  CK(cudaSetDevice(b->device));
  CK(cudaStreamSynchronize(b->stream));
  return 0;
}
void* b2_stream(b2_batch* b) { return b ? (void*)b->stream : nullptr; }
long long b2_launch_count(const b2_batch* b) { return b ? b->launches : -1; }

int b2_set_hw_joints(b2_batch* b, int njoint, const int* jnt_ids) {
  if (!b || njoint < 1 || !jnt_ids) return fail("b2_set_hw_joints: bad argument");
  CK(cudaSetDevice(b->device));
  drop_graphs(b);
  std::vector<int> qadr(njoint), dadr(njoint), ctl(njoint);
  for (int j = 0; j < njoint; j++) {
    const int id = jnt_ids[j];
    if (id < 0 || id >= b->m->njnt) return fail("b2_set_hw_joints: joint id out of range");
    if (b->m->jnt_type[id] != mjJNT_HINGE && b->m->jnt_type[id] != mjJNT_SLIDE) return fail("b2_set_hw_joints: only scalar joints");
    qadr[j] = b->m->jnt_qposadr[id];
    dadr[j] = b->m->jnt_dofadr[id];
    ctl[j] = b->controlled[dadr[j]] ? 1 : 0;
  }
  for (int** p : {&b->hw_qadr, &b->hw_dadr, &b->hw_ctl}) { if (*p) cudaFree(*p); *p = nullptr; }
  if (b->hw_buf) { cudaFree(b->hw_buf); b->hw_buf = nullptr; }
  if (b->hw_kp) { cudaFree(b->hw_kp); b->hw_kp = nullptr; }
  if (b->hw_kd) { cudaFree(b->hw_kd); b->hw_kd = nullptr; }
  CK(cudaMalloc(&b->hw_qadr, sizeof(int) * njoint));
  CK(cudaMalloc(&b->hw_dadr, sizeof(int) * njoint));
  CK(cudaMalloc(&b->hw_ctl, sizeof(int) * njoint));
  CK(cudaMalloc(&b->hw_buf, sizeof(float) * 5 * (size_t)njoint * b->nenv));
  CK(cudaMemcpy(b->hw_qadr, qadr.data(), sizeof(int) * njoint, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b->hw_dadr, dadr.data(), sizeof(int) * njoint, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b->hw_ctl, ctl.data(), sizeof(int) * njoint, cudaMemcpyHostToDevice));
  b->nhw = njoint;
  drop_aliases(b);   // the exchange size changed: every cached alias is re-validated against the new size
  b->hw_identity = njoint == b->m->nv;
  for (int j = 0; j < njoint && b->hw_identity; j++) b->hw_identity = qadr[j] == j && dadr[j] == j;
  io_default(b);
  return 0;
}

}  // extern "C"
namespace {
void io_default(b2_batch* b) {  // the exchange runs on the HBM staging buffers
  const size_t n = (size_t)b->nhw * b->nenv;
  b->io_in[0] = b->hw_buf; b->io_in[1] = b->hw_buf + n;
  b->io_out[0] = b->hw_buf + 2 * n; b->io_out[1] = b->hw_buf + 3 * n; b->io_out[2] = b->hw_buf + 4 * n;
}
int hw_write_async(b2_batch* b, int flags, bool gather) {
  if (!b->nhw) return fail("b2_write_commands: call b2_set_hw_joints first");
  const size_t n = (size_t)b->nhw * b->nenv;
  const int th = 256, bl = (int)((n + th - 1) / th);
  float* po = gather ? b->io_out[0] : nullptr; float* vo = gather ? b->io_out[1] : nullptr;
  const int ctlr = (flags & B2_TICK_CONTROLLER) ? 1 : 0;
  if (b->prec == 8) k_hw_write<double><<<bl, th, 0, b->stream>>>((double*)b->fields["ddq"].ptr, (double*)b->fields["dq"].ptr, b->io_in[0], b->io_in[1], b->hw_dadr, b->hw_ctl, b->nhw, b->nenv, b->nenvp, (const double*)b->fields["qpos"].ptr, (const double*)b->fields["qvel"].ptr, b->hw_qadr, b->hw_kp, b->hw_kd, po, vo, ctlr);
  else k_hw_write<float><<<bl, th, 0, b->stream>>>((float*)b->fields["ddq"].ptr, (float*)b->fields["dq"].ptr, b->io_in[0], b->io_in[1], b->hw_dadr, b->hw_ctl, b->nhw, b->nenv, b->nenvp, (const float*)b->fields["qpos"].ptr, (const float*)b->fields["qvel"].ptr, b->hw_qadr, b->hw_kp, b->hw_kd, po, vo, ctlr);
  b->launches++;
  CK(cudaGetLastError());
  return 0;
}
int hw_read_async(b2_batch* b, bool post, cudaStream_t st) {
  if (!st) st = b->stream;
  if (!b->nhw) return fail("b2_read_joints: call b2_set_hw_joints first");
  const size_t n = (size_t)b->nhw * b->nenv;
  const int th = 256, bl = (int)((n + th - 1) / th);
  float* po = post ? b->io_out[0] : nullptr; float* vo = post ? b->io_out[1] : nullptr;
  if (b->prec == 8) k_hw_read<double><<<bl, th, 0, st>>>((const double*)b->fields["qpos"].ptr, (const double*)b->fields["qvel"].ptr, (const double*)b->fields["qfrc_inverse"].ptr, po, vo, b->io_out[2], b->hw_qadr, b->hw_dadr, b->nhw, b->nenv, b->nenvp);
  else k_hw_read<float><<<bl, th, 0, st>>>((const float*)b->fields["qpos"].ptr, (const float*)b->fields["qvel"].ptr, (const float*)b->fields["qfrc_inverse"].ptr, po, vo, b->io_out[2], b->hw_qadr, b->hw_dadr, b->nhw, b->nenv, b->nenvp);
  b->launches++;
  CK(cudaGetLastError());
  return 0;
}

// Device-accessible alias of a caller buffer, or nullptr when it has to be staged through hw_buf.  Device memory and
// host memory the CALLER pinned (cudaHostAlloc / cudaHostRegister, or b2_register_host) are used in place: the kernels
// read the commands and write the joint states over PCIe themselves, which removes five staging copies from the control
// tick.  Pageable memory is never pinned behind the caller's back (its lifetime is the caller's): it is staged.
// (B2_NO_ZEROCOPY=1 stages everything.)
float* device_alias(b2_batch* b, const void* host, size_t bytes) {
  static const bool off = getenv("B2_NO_ZEROCOPY") != nullptr;
  (void)b; (void)bytes;
  if (off || !host) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return (float*)const_cast<void*>(host);
  if (at.type == cudaMemoryTypeHost) {
    // a registered range shorter than the exchange (registered for an earlier, smaller joint set) cannot be used in place
    auto it = b->registered.find(host);
    if (it != b->registered.end() && it->second < bytes) return nullptr;
    return (float*)at.devicePointer;
  }
  return nullptr;
}
void drop_aliases(b2_batch* b) {
  for (int k = 0; k < 5; k++) { b->alias_host[k] = nullptr; b->alias_dev[k] = nullptr; b->alias_bytes[k] = 0; }
}

}  // namespace
extern "C" {

// Explicit pinning of caller-owned pageable buffers for the in-place (zero-copy) exchange of b2_tick_host.  The caller
// keeps the buffer alive until b2_unregister_host (or b2_destroy) and must unregister BEFORE freeing it.
int b2_register_host(b2_batch* b, void* host, long long bytes) {
  if (!b || !host || bytes <= 0) return fail("b2_register_host: bad argument");
  CK(cudaSetDevice(b->device));
  auto it = b->registered.find(host);
  if (it != b->registered.end()) {
    if (it->second >= (size_t)bytes) return 0;
    cudaHostUnregister(host);
    b->registered.erase(it);
  }
  drop_aliases(b);
  drop_graphs(b);
  const cudaError_t e = cudaHostRegister(host, (size_t)bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(std::string("b2_register_host: cudaHostRegister: ") + cudaGetErrorString(e)); }
  b->registered[host] = (size_t)bytes;
  return 0;
}
int b2_unregister_host(b2_batch* b, void* host) {
  if (!b || !host) return fail("b2_unregister_host: bad argument");
  CK(cudaSetDevice(b->device));
  auto it = b->registered.find(host);
  if (it == b->registered.end()) return fail("b2_unregister_host: not registered through b2_register_host");
  CK(cudaStreamSynchronize(b->stream));   // no kernel may still be reading / writing the range
  drop_aliases(b);
  drop_graphs(b);                          // captured ticks hold the device alias in their kernel arguments
  cudaHostUnregister(host);
  b->registered.erase(it);
  return 0;
}

int b2_set_pd(b2_batch* b, const float* kp, const float* kd) {
  if (!b) return fail("b2_set_pd: null batch");
  CK(cudaSetDevice(b->device));
  if (!b->nhw) return fail("b2_set_pd: call b2_set_hw_joints first");
  drop_graphs(b);
  CK(cudaStreamSynchronize(b->stream));
  if (b->hw_kp) { cudaFree(b->hw_kp); b->hw_kp = nullptr; }
  if (b->hw_kd) { cudaFree(b->hw_kd); b->hw_kd = nullptr; }
  if (!kp && !kd) return 0;
  if (!kp || !kd) return fail("b2_set_pd: kp and kd must both be given (or both NULL)");
  CK(cudaMalloc(&b->hw_kp, sizeof(float) * b->nhw));
  CK(cudaMalloc(&b->hw_kd, sizeof(float) * b->nhw));
  CK(cudaMemcpy(b->hw_kp, kp, sizeof(float) * b->nhw, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b->hw_kd, kd, sizeof(float) * b->nhw, cudaMemcpyHostToDevice));
  return 0;
}

// upload the command buffers into the HBM staging area and apply them (MjHWInterface::write for every environment)
int b2_write_commands(b2_batch* b, const float* vel, const float* eff) {
  if (!b || !vel || !eff) return fail("b2_write_commands: null argument");
  CK(cudaSetDevice(b->device));
  if (!b->nhw) return fail("b2_write_commands: call b2_set_hw_joints first");
  const size_t n = (size_t)b->nhw * b->nenv;
  io_default(b);
  CK(cudaMemcpyAsync(b->hw_buf, vel, n * 4, cudaMemcpyDefault, b->stream));
  CK(cudaMemcpyAsync(b->hw_buf + n, eff, n * 4, cudaMemcpyDefault, b->stream));
  if (hw_write_async(b) < 0) return -1;
  CK(cudaStreamSynchronize(b->stream));
  return 0;
}
int b2_read_joints(b2_batch* b, float* pos, float* vel, float* eff) {
  if (!b) return fail("b2_read_joints: null batch");
  CK(cudaSetDevice(b->device));
  if (!b->nhw) return fail("b2_read_joints: call b2_set_hw_joints first");
  const size_t n = (size_t)b->nhw * b->nenv;
  io_default(b);
  if (hw_read_async(b) < 0) return -1;
  if (pos) CK(cudaMemcpyAsync(pos, b->io_out[0], n * 4, cudaMemcpyDefault, b->stream));
  if (vel) CK(cudaMemcpyAsync(vel, b->io_out[1], n * 4, cudaMemcpyDefault, b->stream));
  if (eff) CK(cudaMemcpyAsync(eff, b->io_out[2], n * 4, cudaMemcpyDefault, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return 0;
}

// One control tick as the reference's loop body sees it (src/mj_main.cpp:82-112):
// write(commands) -> step1 + controller -> read (mj_inverse) -> step2 -> odom; joint states out.
// Order: the reference reads the joint state between mj_step1 and mj_step2 (qpos of the tick's start, qvel after the
// controller's override, qfrc_inverse of this tick) and so does this tick; B2_TICK_READ_POST (b2_set_tick_flags) returns
// the post-integration qpos / qvel instead, i.e. what read() of the NEXT tick would see.
static int tick_hw(b2_batch* b, const float* vel, const float* eff, float* pos, float* velo, float* effo, bool sync) {
  CK(cudaSetDevice(b->device));
  if (!b->nhw) return fail("b2_tick_host: call b2_set_hw_joints first");
  const size_t n = (size_t)b->nhw * b->nenv;
  const int flags = (b->tick_flags & ~(1 << 30)) | B2_TICK_INTEGRATE | B2_TICK_CONTROLLER | B2_TICK_INVERSE | B2_TICK_HW;
  io_default(b);
  // buffers that are not device-accessible in place go through the HBM staging area
  const float* in[2] = {vel, eff};
  float* out[3] = {pos, velo, effo};
  float* stage_out[3] = {nullptr, nullptr, nullptr};
  auto alias = [&](int slot, const void* host) -> float* {
    if (b->alias_host[slot] != host || b->alias_bytes[slot] != n * 4) {
      b->alias_host[slot] = host; b->alias_bytes[slot] = n * 4; b->alias_dev[slot] = device_alias(b, host, n * 4);
    }
    return b->alias_dev[slot];
  };
  for (int k = 0; k < 2; k++) {
    if (!in[k]) continue;  // resident: the staging buffer of the last upload is re-issued
    if (const float* d = alias(k, in[k])) b->io_in[k] = d;
    else CK(cudaMemcpyAsync(b->hw_buf + k * n, in[k], n * 4, cudaMemcpyHostToDevice, b->stream));
  }
  for (int k = 0; k < 3; k++) {
    if (!out[k]) continue;
    if (float* d = alias(2 + k, out[k])) b->io_out[k] = d;
    else stage_out[k] = b->hw_buf + (2 + k) * n;
  }
  if (tick_dispatch(b, flags) < 0) return -1;
  for (int k = 0; k < 3; k++)
    if (stage_out[k]) CK(cudaMemcpyAsync(out[k], stage_out[k], n * 4, cudaMemcpyDeviceToHost, b->stream));
  if (sync) CK(cudaStreamSynchronize(b->stream));
  return 0;
}
// host buffers in, host buffers out, synchronised (the end-to-end path)
int b2_tick_host(b2_batch* b, const float* vel, const float* eff, float* pos, float* velo, float* effo) {
  if (!b || !vel || !eff) return fail("b2_tick_host: null argument");
  return tick_hw(b, vel, eff, pos, velo, effo, true);
}
// same tick with the commands already resident in HBM (b2_write_commands) and the joint states left in HBM (asynchronous)
int b2_tick_resident(b2_batch* b) {
  if (!b) return fail("b2_tick_resident: null batch");
  return tick_hw(b, nullptr, nullptr, nullptr, nullptr, nullptr, false);
}

int b2_set_slots(b2_batch* b, int nslot, const int* body_ids) {
  if (!b || nslot < 0 || (nslot > 0 && !body_ids)) return fail("b2_set_slots: bad argument");
  CK(cudaSetDevice(b->device));
  drop_graphs(b);
  CK(cudaStreamSynchronize(b->stream));
  for (int** p : {&b->slot_qadr, &b->slot_dadr}) { if (*p) cudaFree(*p); *p = nullptr; }
  if (b->slot_active) { cudaFree(b->slot_active); b->slot_active = nullptr; }
  b->nslot = 0;
  if (nslot == 0) return 0;
  const mjModel* m = b->m;
  std::vector<int> qadr(nslot), dadr(nslot);
  for (int s = 0; s < nslot; s++) {
    const int id = body_ids[s];
    if (id < 1 || id >= m->nbody) return fail("b2_set_slots: body id out of range");
    if (m->body_jntnum[id] != 1 || m->jnt_type[m->body_jntadr[id]] != mjJNT_FREE) return fail("b2_set_slots: a slot body must carry exactly one free joint");
    qadr[s] = m->jnt_qposadr[m->body_jntadr[id]];
    dadr[s] = m->jnt_dofadr[m->body_jntadr[id]];
  }
  CK(cudaMalloc(&b->slot_qadr, sizeof(int) * nslot));
  CK(cudaMalloc(&b->slot_dadr, sizeof(int) * nslot));
  CK(cudaMalloc(&b->slot_active, (size_t)nslot * b->nenvp));
  CK(cudaMemcpy(b->slot_qadr, qadr.data(), sizeof(int) * nslot, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b->slot_dadr, dadr.data(), sizeof(int) * nslot, cudaMemcpyHostToDevice));
  CK(cudaMemset(b->slot_active, 1, (size_t)nslot * b->nenvp));
  b->nslot = nslot;
  return 0;
}

static int slot_requests(b2_batch* b, int n, const int* env, const int* slot, const float* pose7, const float* twist6, bool spawn) {
  if (!b || n < 0 || (n > 0 && (!env || !slot)) || (spawn && n > 0 && !pose7)) return fail("b2_spawn / b2_destroy_slots: bad argument");
  CK(cudaSetDevice(b->device));
  if (!b->nslot) return fail("b2_spawn / b2_destroy_slots: call b2_set_slots first");
  if (n == 0) return 0;
  for (int i = 0; i < n; i++)
    if (env[i] < 0 || env[i] >= b->nenv || slot[i] < 0 || slot[i] >= b->nslot) return fail("b2_spawn / b2_destroy_slots: environment or slot out of range");
  // requests are applied in order: of several requests for one (environment, slot) only the last one has an effect
  std::vector<int> env_last(env, env + n);
  {
    std::unordered_map<long long, int> last;
    last.reserve((size_t)n * 2);
    for (int i = 0; i < n; i++) {
      const long long key = (long long)env[i] * b->nslot + slot[i];
      auto it = last.find(key);
      if (it != last.end()) { env_last[it->second] = -1; it->second = i; } else last.emplace(key, i);
    }
  }
  env = env_last.data();
  // one staging block: env | slot | pose7 | twist6
  const size_t bytes = (size_t)n * (2 * sizeof(int) + 13 * sizeof(float));
  if (ensure_stage(b, bytes) < 0) return -1;
  char* base = (char*)b->stage_dev;
  int* d_env = (int*)base; int* d_slot = d_env + n;
  float* d_pose = (float*)(d_slot + n); float* d_tw = d_pose + (size_t)7 * n;
  CK(cudaMemcpyAsync(d_env, env, sizeof(int) * n, cudaMemcpyHostToDevice, b->stream));
  CK(cudaMemcpyAsync(d_slot, slot, sizeof(int) * n, cudaMemcpyHostToDevice, b->stream));
  if (spawn) CK(cudaMemcpyAsync(d_pose, pose7, sizeof(float) * 7 * n, cudaMemcpyHostToDevice, b->stream));
  if (spawn && twist6) CK(cudaMemcpyAsync(d_tw, twist6, sizeof(float) * 6 * n, cudaMemcpyHostToDevice, b->stream));
  auto P = [&](const char* nm) { return b->fields[nm].ptr; };
  if (b->prec == 8)
    k_slot_requests<double><<<std::min(64, (n + 255) / 256), 256, 0, b->stream>>>((double*)P("qpos"), (double*)P("qvel"), (double*)P("qacc"), (double*)P("qacc_warmstart"), b->slot_active,
                                                      b->slot_qadr, b->slot_dadr, b->nenvp, n, d_env, d_slot, spawn ? d_pose : nullptr, spawn && twist6 ? d_tw : nullptr);
  else
    k_slot_requests<float><<<std::min(64, (n + 255) / 256), 256, 0, b->stream>>>((float*)P("qpos"), (float*)P("qvel"), (float*)P("qacc"), (float*)P("qacc_warmstart"), b->slot_active,
                                                     b->slot_qadr, b->slot_dadr, b->nenvp, n, d_env, d_slot, spawn ? d_pose : nullptr, spawn && twist6 ? d_tw : nullptr);
  b->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(b->stream));   // the host arrays may be reused by the caller
  return 0;
}
int b2_spawn(b2_batch* b, int n, const int* env, const int* slot, const float* pose7, const float* twist6) {
  return slot_requests(b, n, env, slot, pose7, twist6, true);
}
int b2_destroy_slots(b2_batch* b, int n, const int* env, const int* slot) { return slot_requests(b, n, env, slot, nullptr, nullptr, false); }
int b2_slot_active(b2_batch* b, unsigned char* active, int lo, int hi) {
  if (!b || !active || lo < 0 || hi > b->nenv || lo >= hi) return fail("b2_slot_active: bad argument");
  CK(cudaSetDevice(b->device));
  if (!b->nslot) return fail("b2_slot_active: call b2_set_slots first");
  CK(cudaStreamSynchronize(b->stream));
  std::vector<unsigned char> tmp((size_t)b->nslot * b->nenvp);
  CK(cudaMemcpy(tmp.data(), b->slot_active, tmp.size(), cudaMemcpyDeviceToHost));
  for (int e = lo; e < hi; e++)
    for (int s = 0; s < b->nslot; s++) active[(size_t)(e - lo) * b->nslot + s] = tmp[(size_t)s * b->nenvp + e];
  return 0;
}

int b2_pack_obs(b2_batch* b, float* obs_dev) {
  if (!b || !obs_dev) return fail("b2_pack_obs: null argument");
  CK(cudaSetDevice(b->device));
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, obs_dev) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
    cudaGetLastError();
    return fail("b2_pack_obs: obs_dev must be device memory");
  }
  const int nq = b->m->nq, nv = b->m->nv;
  const long long tot = (long long)(nq + nv) * b->nenv;
  const int th = 256, bl = (int)((tot + th - 1) / th);
  if (b->prec == 8) k_pack_obs<double><<<bl, th, 0, b->stream>>>((const double*)b->fields["qpos"].ptr, (const double*)b->fields["qvel"].ptr, obs_dev, nq, nv, b->nenv, b->nenvp);
  else k_pack_obs<float><<<bl, th, 0, b->stream>>>((const float*)b->fields["qpos"].ptr, (const float*)b->fields["qvel"].ptr, obs_dev, nq, nv, b->nenv, b->nenvp);
  b->launches++;
  CK(cudaGetLastError());
  return 0;
}

// the reference's add_old_state (src/mujoco_sim/mj_sim.cpp:465-558) for whole batches
int b2_transfer_state(b2_batch* src, b2_batch* dst) {
  if (!src || !dst || src == dst) return fail("b2_transfer_state: bad argument");
  if (src->nenv != dst->nenv || src->prec != dst->prec || src->device != dst->device) return fail("b2_transfer_state: the batches must agree in environments, precision and device");
  CK(cudaSetDevice(dst->device));
  CK(cudaStreamSynchronize(src->stream));
  const mjModel* ma = src->m;
  const mjModel* mb = dst->m;
  const size_t row = (size_t)dst->prec, na = src->nenvp, nb_ = dst->nenvp;
  auto copy_rows = [&](const char* field, int ia, int ib, int n) -> int {
    auto fa = src->fields.find(field), fb = dst->fields.find(field);
    if (fa == src->fields.end() || fb == dst->fields.end()) return 0;
    for (int k = 0; k < n; k++)
      CK(cudaMemcpyAsync((char*)fb->second.ptr + (size_t)(ib + k) * nb_ * row, (const char*)fa->second.ptr + (size_t)(ia + k) * na * row,
                         (size_t)src->nenv * row, cudaMemcpyDeviceToDevice, dst->stream));
    return 0;
  };
  int carried = 0;
  for (int i = 1; i < ma->nbody; i++) {
    const char* name = mj_id2name(ma, mjOBJ_BODY, i);
    if (!name || !*name) continue;
    const int j = mj_name2id(mb, mjOBJ_BODY, name);
    if (j < 0) continue;
    if (ma->body_jntnum[i] != mb->body_jntnum[j] || ma->body_dofnum[i] != mb->body_dofnum[j]) continue;   // (the reference warns and skips)
    if (ma->body_jntnum[i] > 0) {
      // qpos width of the body's joints (free 7, ball 4, scalar 1); the joint types must agree as well
      int wa = 0, wb = 0;
      bool same = true;
      for (int k = 0; k < ma->body_jntnum[i]; k++) {
        const int ta = ma->jnt_type[ma->body_jntadr[i] + k], tb = mb->jnt_type[mb->body_jntadr[j] + k];
        same &= ta == tb;
        wa += ta == mjJNT_FREE ? 7 : (ta == mjJNT_BALL ? 4 : 1);
        wb += tb == mjJNT_FREE ? 7 : (tb == mjJNT_BALL ? 4 : 1);
      }
      if (!same || wa != wb) continue;
      const int qa = ma->jnt_qposadr[ma->body_jntadr[i]], qb = mb->jnt_qposadr[mb->body_jntadr[j]];
      const int da = ma->body_dofadr[i], db = mb->body_dofadr[j], nd = ma->body_dofnum[i];
      if (copy_rows("qpos", qa, qb, wa) < 0) return -1;
      for (const char* f : {"qvel", "qacc", "qacc_warmstart", "qfrc_applied"})
        if (copy_rows(f, da, db, nd) < 0) return -1;
    }
    carried++;
  }
  if (copy_rows("time", 0, 0, 1) < 0) return -1;
  CK(cudaStreamSynchronize(dst->stream));
  return carried;
}

// ---- fused observation exchange (include/b2_batch.h) ----
float* b2_obs_create(b2_batch* b, int world, int rank) {
  if (!b || world < 1 || rank < 0 || rank >= world) { fail("b2_obs_create: bad argument"); return nullptr; }
  if (cudaSetDevice(b->device) != cudaSuccess) { fail("b2_obs_create: cudaSetDevice"); return nullptr; }
  if (b->obs_buf) { fail("b2_obs_create: already created"); return nullptr; }
  const size_t bytes = (size_t)world * (b->m->nq + b->m->nv) * b->nenv * sizeof(float);
  if (cudaMalloc(&b->obs_buf, bytes) != cudaSuccess) { fail("b2_obs_create: cudaMalloc"); return nullptr; }
  cudaMemsetAsync(b->obs_buf, 0, bytes, b->stream);
  b->obs_world = world; b->obs_rank = rank;
  return b->obs_buf;
}
int b2_obs_handle(b2_batch* b, void* handle64) {
  if (!b || !b->obs_buf || !handle64) return fail("b2_obs_handle: call b2_obs_create first");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  CK(cudaSetDevice(b->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, b->obs_buf));
  std::memcpy(handle64, &h, 64);
  return 0;
}
int b2_obs_attach(b2_batch* b, const void* handles, float* const* ptrs) {
  if (!b || !b->obs_buf || (!handles && !ptrs)) return fail("b2_obs_attach: call b2_obs_create first and pass handles or pointers");
  CK(cudaSetDevice(b->device));
  drop_graphs(b);
  b->obs_peers_host.assign(b->obs_world, nullptr);
  b->obs_peer_ipc.assign(b->obs_world, false);
  for (int p = 0; p < b->obs_world; p++) {
    if (p == b->obs_rank) { b->obs_peers_host[p] = b->obs_buf; continue; }
    if (ptrs) {
      cudaPointerAttributes at;
      CK(cudaPointerGetAttributes(&at, ptrs[p]));
      if (at.device != b->device) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(std::string("b2_obs_attach: no peer access to device ") + std::to_string(at.device));
        cudaGetLastError();
      }
      b->obs_peers_host[p] = ptrs[p];
    } else {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, (const char*)handles + (size_t)64 * p, 64);
      void* q = nullptr;
      CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      b->obs_peers_host[p] = (float*)q;
      b->obs_peer_ipc[p] = true;
    }
  }
  if (!b->obs_peers_dev) CK(cudaMalloc(&b->obs_peers_dev, sizeof(float*) * b->obs_world));
  CK(cudaMemcpy(b->obs_peers_dev, b->obs_peers_host.data(), sizeof(float*) * b->obs_world, cudaMemcpyHostToDevice));
  b->obs_on = true;
  return 0;
}
int b2_obs_read(b2_batch* b, float* host) {
  if (!b || !b->obs_buf || !host) return fail("b2_obs_read: call b2_obs_create first");
  CK(cudaSetDevice(b->device));
  CK(cudaStreamSynchronize(b->stream));
  CK(cudaMemcpy(host, b->obs_buf, (size_t)b->obs_world * (b->m->nq + b->m->nv) * b->nenv * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}
int b2_obs_enable(b2_batch* b, int on) {
  if (!b) return fail("b2_obs_enable: null batch");
  if (on && !b->obs_peers_dev) return fail("b2_obs_enable: call b2_obs_attach first");
  if (b->obs_on != (on != 0)) drop_graphs(b);   // the flag is baked into the captured kernel arguments
  b->obs_on = on != 0;
  return 0;
}

// ---- one process, several devices ----
struct b2_multi { std::vector<b2_batch*> shard; };
b2_multi* b2_create_multi(const mjModel* m, int nenv, const int* devices, int ndev, int precision, int obs) {
  if (!m || nenv < 1 || !devices || ndev < 1 || nenv % ndev) { fail("b2_create_multi: bad argument (nenv must divide by ndev)"); return nullptr; }
  auto* mb = new b2_multi();
  for (int i = 0; i < ndev; i++) {
    b2_batch* b = b2_create(m, nenv / ndev, devices[i], precision);
    if (!b) { b2_multi_destroy(mb); return nullptr; }
    mb->shard.push_back(b);
  }
  if (obs) {
    std::vector<float*> ptrs(ndev, nullptr);
    for (int i = 0; i < ndev; i++) { ptrs[i] = b2_obs_create(mb->shard[i], ndev, i); if (!ptrs[i]) { b2_multi_destroy(mb); return nullptr; } }
    for (int i = 0; i < ndev; i++) if (b2_obs_attach(mb->shard[i], nullptr, ptrs.data()) < 0) { b2_multi_destroy(mb); return nullptr; }
  }
  return mb;
}
void b2_multi_destroy(b2_multi* mb) {
  if (!mb) return;
  for (b2_batch* b : mb->shard) b2_sync(b);   // nobody writes into a buffer that is about to be freed
  for (b2_batch* b : mb->shard) b2_destroy(b);
  delete mb;
}
int b2_multi_count(const b2_multi* mb) { return mb ? (int)mb->shard.size() : 0; }
b2_batch* b2_multi_shard(b2_multi* mb, int i) { return (mb && i >= 0 && i < (int)mb->shard.size()) ? mb->shard[i] : nullptr; }
int b2_multi_tick(b2_multi* mb, int flags) {
  if (!mb) return fail("b2_multi_tick: null");
  for (b2_batch* b : mb->shard) if (b2_tick(b, flags) < 0) return -1;   // asynchronous launches: the devices run side by side
  return 0;
}
int b2_multi_sync(b2_multi* mb) {
  if (!mb) return fail("b2_multi_sync: null");
  for (b2_batch* b : mb->shard) if (b2_sync(b) < 0) return -1;
  return 0;
}

// write `bytes` of scratch on the batch's stream: evicts the batch state from the 126 MB L2 between timed steps
int b2_l2_flush(b2_batch* b, long long bytes) {
  if (!b || bytes <= 0) return fail("b2_l2_flush: bad argument");
  CK(cudaSetDevice(b->device));
  if ((size_t)bytes > b->flush_bytes) {
    if (b->flush_buf) cudaFree(b->flush_buf);
    b->flush_buf = nullptr; b->flush_bytes = 0;
    CK(cudaMalloc(&b->flush_buf, (size_t)bytes));
    b->flush_bytes = (size_t)bytes;
  }
  CK(cudaMemsetAsync(b->flush_buf, 1, (size_t)bytes, b->stream));
  return 0;
}

int b2_profile_begin(b2_batch* b, int max_ticks) {
  if (!b || max_ticks < 1) return fail("b2_profile_begin: bad argument");
  CK(cudaSetDevice(b->device));
  const size_t need = (size_t)max_ticks * (B2_NSLOT + 1);
  while (b->prof_ev.size() < need) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    b->prof_ev.push_back(e);
  }
  b->prof_mask.assign(max_ticks, 0u);
  b->prof_max = max_ticks; b->prof_n = 0; b->prof_on = true; b->prof_tick_open = -1;
  return 0;
}
// ms[slot] = summed device time of kernel slot over the profiled ticks (slots: hw_write, smooth, collide,
// make_constraint, project, pgs, integrate, hw_read); returns the number of ticks profiled
int b2_profile_end(b2_batch* b, double* ms, int nslot) {
  if (!b || !ms) return fail("b2_profile_end: null argument");
  CK(cudaSetDevice(b->device));
  CK(cudaStreamSynchronize(b->stream));
  b->prof_on = false;
  for (int s = 0; s < nslot; s++) ms[s] = 0;
  for (int t = 0; t < b->prof_n; t++) {
    // boundary events that were not recorded this tick are skipped by walking to the next recorded one
    const size_t base = (size_t)t * (B2_NSLOT + 1);
    int prev = -1;
    for (int s = 0; s <= B2_NSLOT; s++) {
      if (!(b->prof_mask[t] & (1u << s))) continue;
      if (prev >= 0 && prev < nslot) {
        float dt = 0;
        if (cudaEventElapsedTime(&dt, b->prof_ev[base + prev], b->prof_ev[base + s]) == cudaSuccess) ms[prev] += dt; else cudaGetLastError();
      }
      prev = s;
    }
  }
  return b->prof_n;
}

int b2_mirror_env(b2_batch* b, int env, mjData* d) {
  if (!b || !d || env < 0 || env >= b->nenv) return fail("b2_mirror_env: bad argument");
  const mjModel* m = b->m;
  auto G = [&](const char* name, double* dst, int n) -> int {
    if (n <= 0 || !dst) return 0;
    auto it = b->fields.find(name);
    if (it == b->fields.end()) return 0;
    if (it->second.count < n) return fail(std::string("b2_mirror_env: size mismatch for ") + name);
    std::vector<double> tmp(it->second.count);
    if (get_field<double>(b, name, tmp.data(), env, env + 1, B2_ENV_MAJOR, 0) < 0) return -1;
    std::copy(tmp.begin(), tmp.begin() + n, dst);
    return 0;
  };
  const int nq = m->nq, nv = m->nv, nb = m->nbody, ng = m->ngeom;
  if (G("qpos", d->qpos, nq) < 0 || G("qvel", d->qvel, nv) < 0 || G("qacc", d->qacc, nv) < 0 ||
      G("qacc_warmstart", d->qacc_warmstart, nv) < 0 || G("qfrc_applied", d->qfrc_applied, nv) < 0 ||
      G("qfrc_bias", d->qfrc_bias, nv) < 0 || G("qfrc_inverse", d->qfrc_inverse, nv) < 0 || G("xpos", d->xpos, 3 * nb) < 0 ||
      G("xquat", d->xquat, 4 * nb) < 0 || G("xmat", d->xmat, 9 * nb) < 0 || G("geom_xpos", d->geom_xpos, 3 * ng) < 0 ||
      G("geom_xmat", d->geom_xmat, 9 * ng) < 0 || G("qM", d->qM, m->nM) < 0 || G("qfrc_passive", d->qfrc_passive, nv) < 0 ||
      G("qfrc_constraint", d->qfrc_constraint, nv) < 0 || G("qacc_smooth", d->qacc_smooth, nv) < 0 ||
      G("qfrc_smooth", d->qfrc_smooth, nv) < 0 || G("subtree_com", d->subtree_com, 3 * nb) < 0 || G("cdof", d->cdof, 6 * nv) < 0)
    return -1;
  double t = 0;
  if (get_field<double>(b, "time", &t, env, env + 1, B2_ENV_MAJOR, 0) < 0) return -1;
  d->time = t;
  int v = 0;
  if (get_field<int>(b, "ncon", &v, env, env + 1, B2_ENV_MAJOR, 1) < 0) return -1;
  d->ncon = b->fused ? 0 : v;
  if (get_field<int>(b, "nefc", &v, env, env + 1, B2_ENV_MAJOR, 1) < 0) return -1;
  d->nefc = b->fused ? 0 : v;
  if (get_field<int>(b, "solver_iter", &v, env, env + 1, B2_ENV_MAJOR, 1) < 0) return -1;
  d->solver_iter = v;
  if (!b->fused && d->ncon > 0) {
    std::vector<double> cf((size_t)CF_NFLOAT * b->hdr.nconmax);
    std::vector<int> ci((size_t)CI_NINT * b->hdr.nconmax);
    if (get_field<double>(b, "contact", cf.data(), env, env + 1, B2_ENV_MAJOR, 0) < 0) return -1;
    if (get_field<int>(b, "contact_int", ci.data(), env, env + 1, B2_ENV_MAJOR, 1) < 0) return -1;
    const int ncm = b->hdr.nconmax;
    for (int c = 0; c < d->ncon && c < m->nconmax; c++) {
      mjContact& k = d->contact[c];
      k.dist = cf[(size_t)CF_DIST * ncm + c];
      for (int i = 0; i < 3; i++) k.pos[i] = cf[(size_t)(CF_POS + i) * ncm + c];
      for (int i = 0; i < 9; i++) k.frame[i] = cf[(size_t)(CF_FRAME + i) * ncm + c];
      k.includemargin = cf[(size_t)CF_INCLUDEMARGIN * ncm + c];
      for (int i = 0; i < 5; i++) k.friction[i] = cf[(size_t)(CF_FRICTION + i) * ncm + c];
      for (int i = 0; i < 2; i++) k.solref[i] = cf[(size_t)(CF_SOLREF + i) * ncm + c];
      for (int i = 0; i < 5; i++) k.solimp[i] = cf[(size_t)(CF_SOLIMP + i) * ncm + c];
      k.mu = k.friction[0];
      k.geom1 = ci[(size_t)CI_GEOM1 * ncm + c]; k.geom2 = ci[(size_t)CI_GEOM2 * ncm + c]; k.dim = ci[(size_t)CI_DIM * ncm + c];
      k.pair = ci[(size_t)CI_PAIR * ncm + c]; k.efc_address = ci[(size_t)CI_EFC * ncm + c]; k.exclude = 0;
    }
  }
  if (!b->fused && d->nefc > 0) {
    const int njm = b->hdr.njmax;
    std::vector<double> tmp((size_t)njm * std::max(njm, nv));
    auto E = [&](const char* name, double* dst) -> int {
      if (get_field<double>(b, name, tmp.data(), env, env + 1, B2_ENV_MAJOR, 0) < 0) return -1;
      std::copy(tmp.begin(), tmp.begin() + d->nefc, dst);
      return 0;
    };
    if (E("efc_pos", d->efc_pos) < 0 || E("efc_margin", d->efc_margin) < 0 || E("efc_frictionloss", d->efc_frictionloss) < 0 ||
        E("efc_diagApprox", d->efc_diagApprox) < 0 || E("efc_R", d->efc_R) < 0 || E("efc_D", d->efc_D) < 0 || E("efc_vel", d->efc_vel) < 0 ||
        E("efc_aref", d->efc_aref) < 0 || E("efc_b", d->efc_b) < 0 || E("efc_force", d->efc_force) < 0)
      return -1;
    if (get_field<double>(b, "efc_J", tmp.data(), env, env + 1, B2_ENV_MAJOR, 0) < 0) return -1;
    std::copy(tmp.begin(), tmp.begin() + (size_t)d->nefc * nv, d->efc_J);
    std::vector<int> it((size_t)njm);
    if (get_field<int>(b, "efc_type", it.data(), env, env + 1, B2_ENV_MAJOR, 1) < 0) return -1;
    std::copy(it.begin(), it.begin() + d->nefc, d->efc_type);
    if (get_field<int>(b, "efc_id", it.data(), env, env + 1, B2_ENV_MAJOR, 1) < 0) return -1;
    std::copy(it.begin(), it.begin() + d->nefc, d->efc_id);
  }
  // derived quantities the publishers / viewer read (cacc, cfrc_int, energy, ...): computed on the host for this one
  // environment (mirror_post.cpp)
  if (b->fields.count("subtree_com") && b->fields.count("cdof") && b->fields.count("qM")) {
    std::vector<double> xf((size_t)6 * nb, 0.0);
    if (b->tick_flags & (1 << 30)) { if (G("xfrc_applied", xf.data(), 6 * nb) < 0) return -1; }
    b2::mirror_post(m, d, xf.data());
  }
  return 0;
}

int b2_load_env(b2_batch* b, int env, const mjData* d) {
  if (!b || !d || env < 0 || env >= b->nenv) return fail("b2_load_env: bad argument");
  const mjModel* m = b->m;
  auto S = [&](const char* name, const double* src, int n) -> int {
    if (n <= 0 || !src) return 0;
    std::vector<double> tmp(b->fields[name].count, 0.0);
    std::copy(src, src + n, tmp.begin());
    return set_field<double>(b, name, tmp.data(), env, env + 1, B2_ENV_MAJOR) < 0 ? -1 : 0;
  };
  if (S("qpos", d->qpos, m->nq) < 0 || S("qvel", d->qvel, m->nv) < 0 || S("qacc", d->qacc, m->nv) < 0 ||
      S("qacc_warmstart", d->qacc_warmstart, m->nv) < 0 || S("qfrc_applied", d->qfrc_applied, m->nv) < 0)
    return -1;
  bool any = false;
  for (int i = 0; i < 6 * m->nbody; i++) any |= d->xfrc_applied[i] != 0;
  if (any || (b->tick_flags & (1 << 30))) if (S("xfrc_applied", d->xfrc_applied, 6 * m->nbody) < 0) return -1;
  if (m->nmocap > 0) if (S("mocap_pos", d->mocap_pos, 3 * m->nmocap) < 0 || S("mocap_quat", d->mocap_quat, 4 * m->nmocap) < 0) return -1;
  double t = d->time;
  return set_field<double>(b, "time", &t, env, env + 1, B2_ENV_MAJOR) < 0 ? -1 : 0;
}

}  // extern "C"
