// shim_step.cpp — the stepping half of the MuJoCo-named C API (include/mujoco/mujoco.h), backed by the batched
// CUDA engine.  (m, d) is bound to environment 0 of a b2_batch of B2_NUM_ENVS environments (default 1) that is created
// on first use; the other environments are stepped in lockstep.  Replaces the libmujoco calls made by the reference's
// hot loop: mj_step1 / mj_step2 (src/mj_main.cpp:83,108), mj_inverse (src/mujoco_sim/mj_hw_interface.cpp:61),
// mj_forward (src/mujoco_sim/mj_ros.cpp:608,1421), mj_mulM (src/mujoco_sim/mj_sim.cpp:1057).
//
// Semantics note (documented in DESIGN.md): every call re-evaluates the pipeline from the state found in `d`
// (qpos, qvel, qacc, qacc_warmstart, qfrc_applied, xfrc_applied, mocap, time), so mj_step2 uses velocity-stage
// quantities that are consistent with the qvel the controller callback may have just overridden — which is what the
// reference gets anyway because MjHWInterface::read() calls mj_inverse between mj_step1 and mj_step2.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "b2_batch.h"
#include "hostmath.h"
#include "model_store.h"

namespace b2 {

namespace {
struct Bound { const mjModel* m; mjData* d; b2_batch* b; };
std::vector<Bound> g_bound;
std::mutex g_mtx;

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}

b2_batch* batch_for(const mjModel* m, mjData* d) {
  std::lock_guard<std::mutex> lk(g_mtx);
  for (auto& x : g_bound)
    if (x.m == m && x.d == d) return x.b;
  const int nenv = std::max(1, env_int("B2_NUM_ENVS", 1));
  const int dev = env_int("B2_DEVICE", 0);
  const int prec = env_int("B2_PRECISION", 4) == 8 ? B2_F64 : B2_F32;
  b2_batch* b = b2_create(m, nenv, dev, prec | B2_EXPORT_STAGES);
  if (!b) mju_error("b2 shim: cannot create the GPU batch: %s", b2_last_error());
  g_bound.push_back({m, d, b});
  return b;
}

void run(const mjModel* m, mjData* d, int flags) {
  b2_batch* b = batch_for(m, d);
  // the reference mutates m->opt.timestep at run time (src/mj_main.cpp:150-163)
  if (b2_set_timestep(b, m->opt.timestep) < 0 || b2_load_env(b, 0, d) < 0 || b2_tick(b, flags) < 0 || b2_mirror_env(b, 0, d) < 0)
    mju_error("b2 shim: %s", b2_last_error());
}
}  // namespace

void shim_forget(const mjModel* m, const mjData* d) {
  std::lock_guard<std::mutex> lk(g_mtx);
  for (size_t i = 0; i < g_bound.size();) {
    if ((m && g_bound[i].m == m) || (d && g_bound[i].d == d)) {
      b2_destroy(g_bound[i].b);
      g_bound.erase(g_bound.begin() + i);
    } else {
      i++;
    }
  }
}

}  // namespace b2

extern "C" {

// batch bound to (m, d): lets a host that started with the MuJoCo-named API reach the batched one
b2_batch* b2_shim_batch(const mjModel* m, mjData* d) { return b2::batch_for(m, d); }

void mj_step1(const mjModel* m, mjData* d) {
  b2::run(m, d, B2_TICK_NOSOLVE);
  if (mjcb_control) mjcb_control(m, d);
}

void mj_step2(const mjModel* m, mjData* d) { b2::run(m, d, B2_TICK_INTEGRATE); }

void mj_step(const mjModel* m, mjData* d) {
  mj_step1(m, d);
  mj_step2(m, d);
}

void mj_forward(const mjModel* m, mjData* d) {
  if (mjcb_control) {
    b2::run(m, d, B2_TICK_NOSOLVE);
    mjcb_control(m, d);
  }
  b2::run(m, d, 0);
}

void mj_inverse(const mjModel* m, mjData* d) { b2::run(m, d, B2_TICK_INVERSE | B2_TICK_NOSOLVE); }

// res = M vec on the mirrored sparse inertia matrix (reference: tau = M ddq, src/mujoco_sim/mj_sim.cpp:1057)
void mj_mulM(const mjModel* m, const mjData* d, mjtNum* res, const mjtNum* vec) {
  const int nv = m->nv;
  for (int i = 0; i < nv; i++) res[i] = 0;
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    res[i] += d->qM[adr] * vec[i];
    adr++;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j], adr++) {
      res[i] += d->qM[adr] * vec[j];
      res[j] += d->qM[adr] * vec[i];
    }
  }
}

}  // extern "C"
