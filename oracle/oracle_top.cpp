// oracle_top.cpp — stage orchestration of the fp64 CPU oracle: mj_step1 / mj_step2 / mj_forward / mj_inverse and the
// reference's tick (src/mj_main.cpp:82-112).  TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED.
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <thread>

#include "oracle.h"
#include "oracle_util.h"

using namespace omath;

extern "C" void omj_invConstraint(const mjModel* m, mjData* d);

void omj_fwdPosition(const mjModel* m, mjData* d) {
  omj_kinematics(m, d);
  omj_comPos(m, d);
  omj_crb(m, d);
  omj_factorM(m, d);
  omj_collision(m, d);
  omj_makeConstraint(m, d);
  omj_projectConstraint(m, d);
}

// mj_checkPos / mj_checkVel / mj_checkAcc: a non-finite or huge value resets the environment to qpos0
static bool bad(const mjtNum* x, int n) {
  for (int i = 0; i < n; i++)
    if (!std::isfinite(x[i]) || std::fabs(x[i]) > mjMAXVAL) return true;
  return false;
}
static void reset_state(const mjModel* m, mjData* d) {
  copy(d->qpos, m->qpos0, m->nq);
  zero(d->qvel, m->nv);
  zero(d->qacc, m->nv);
  zero(d->qacc_warmstart, m->nv);
  zero(d->qfrc_applied, m->nv);
  d->time = 0;
}

// src/mj_main.cpp:83. The control callback (mjcb_control) is NOT invoked here: omj_tick applies omj_controller.
void omj_step1(const mjModel* m, mjData* d) {
  if (bad(d->qpos, m->nq) || bad(d->qvel, m->nv)) reset_state(m, d);
  omj_fwdPosition(m, d);
  omj_fwdVelocity(m, d);
  if (m->opt.enableflags & mjENBL_ENERGY) omj_energy(m, d);
}

// src/mj_main.cpp:108. RK4 is not available in the split API: always semi-implicit Euler.
void omj_step2(const mjModel* m, mjData* d) {
  omj_fwdAcceleration(m, d);
  omj_fwdConstraint(m, d);
  if (bad(d->qacc, m->nv)) { reset_state(m, d); return; }
  omj_Euler(m, d);
}

void omj_forward(const mjModel* m, mjData* d) {
  omj_step1(m, d);
  omj_fwdAcceleration(m, d);
  omj_fwdConstraint(m, d);
}

void omj_step(const mjModel* m, mjData* d) {
  omj_step1(m, d);
  omj_step2(m, d);
}

// src/mujoco_sim/mj_hw_interface.cpp:61. Recomputes position + velocity stages for the current (qpos, qvel), the
// constraint force implied by the stored qacc, and qfrc_inverse = RNE(q, v, qacc) + armature qacc - passive - constraint.
void omj_inverse(const mjModel* m, mjData* d) {
  omj_kinematics(m, d);
  omj_comPos(m, d);
  omj_crb(m, d);
  omj_factorM(m, d);
  omj_collision(m, d);
  omj_makeConstraint(m, d);
  omj_projectConstraint(m, d);  // harmless: keeps efc_AR valid for the following mj_step2
  omj_fwdVelocity(m, d);
  omj_invConstraint(m, d);
  omj_rne(m, d, 1, d->qfrc_inverse);
  for (int i = 0; i < m->nv; i++)
    d->qfrc_inverse[i] += m->dof_armature[i] * d->qacc[i] - d->qfrc_passive[i] - d->qfrc_constraint[i];
}

// One tick in the order of src/mj_main.cpp:82-112 (commands in ddq/dq were written by MjHWInterface::write on the
// previous tick and are consumed + zeroed by the controller here).
void omj_tick(const mjModel* m, mjData* d, mjtNum* ddq, mjtNum* dq, const mjtByte* controlled, int do_inverse) {
  omj_step1(m, d);
  if (ddq && dq) omj_controller(m, d, ddq, dq, controlled);  // mjcb_control at the end of mj_step1
  if (do_inverse) omj_inverse(m, d);                           // MjHWInterface::read()
  omj_step2(m, d);
}

// pd_kp / pd_kd ([nv], optional): dofs with a non-zero gain take ddq as a position TARGET and are commanded
// kp (target - q) - kd qdot every tick — what the reference's ros_control PID controllers feed into write()
// (gains: model/ontology/box/box.yaml:5-13); mirrors b2_set_pd of the CUDA engine.
int omj_tick_batch_pd(const mjModel* m, mjData** pool, int npool, int nenv, int nsteps, mjtNum* qpos, mjtNum* qvel,
                      mjtNum* qacc_warmstart, const mjtNum* qfrc_applied, const mjtNum* ddq, const mjtNum* dq,
                      const mjtByte* controlled, int do_inverse, mjtNum* qfrc_inverse_out, const mjtNum* pd_kp, const mjtNum* pd_kd) {
  const int nq = m->nq, nv = m->nv;
  const int used = std::max(1, std::min(npool, nenv));
  auto work = [&](int tid) {
    // static split: thread tid owns environments [lo, hi)
    const int lo = (int)((long long)nenv * tid / used), hi = (int)((long long)nenv * (tid + 1) / used);
    mjData* d = pool[tid];
    std::vector<mjtNum> c_ddq(nv, 0.0), c_dq(nv, 0.0);
    for (int e = lo; e < hi; e++) {
      copy(d->qpos, qpos + (size_t)e * nq, nq);
      copy(d->qvel, qvel + (size_t)e * nv, nv);
      if (qacc_warmstart) copy(d->qacc_warmstart, qacc_warmstart + (size_t)e * nv, nv); else zero(d->qacc_warmstart, nv);
      zero(d->qacc, nv);
      if (qfrc_applied) copy(d->qfrc_applied, qfrc_applied + (size_t)e * nv, nv); else zero(d->qfrc_applied, nv);
      d->time = 0;
      for (int s = 0; s < nsteps; s++) {
        const bool ctl = ddq && dq;
        if (ctl) {  // the same command is re-issued every tick (write() runs every tick in the reference)
          copy(c_ddq.data(), ddq + (size_t)e * nv, nv);
          copy(c_dq.data(), dq + (size_t)e * nv, nv);
          if (pd_kp && pd_kd)
            for (int i = 0; i < nv; i++) {
              if (pd_kp[i] == 0 && pd_kd[i] == 0) continue;
              const int j = m->dof_jntid[i];
              c_ddq[i] = pd_kp[i] * (c_ddq[i] - d->qpos[m->jnt_qposadr[j]]) - pd_kd[i] * d->qvel[i];
            }
        }
        omj_tick(m, d, ctl ? c_ddq.data() : nullptr, ctl ? c_dq.data() : nullptr, controlled, do_inverse);
      }
      copy(qpos + (size_t)e * nq, d->qpos, nq);
      copy(qvel + (size_t)e * nv, d->qvel, nv);
      if (qacc_warmstart) copy(qacc_warmstart + (size_t)e * nv, d->qacc_warmstart, nv);
      if (qfrc_inverse_out) copy(qfrc_inverse_out + (size_t)e * nv, d->qfrc_inverse, nv);
    }
  };
  if (used == 1) { work(0); return 1; }
  std::vector<std::thread> th;
  for (int t = 0; t < used; t++) th.emplace_back(work, t);
  for (auto& t : th) t.join();
  return used;
}

int omj_tick_batch(const mjModel* m, mjData** pool, int npool, int nenv, int nsteps, mjtNum* qpos, mjtNum* qvel,
                   mjtNum* qacc_warmstart, const mjtNum* qfrc_applied, const mjtNum* ddq, const mjtNum* dq,
                   const mjtByte* controlled, int do_inverse, mjtNum* qfrc_inverse_out) {
  return omj_tick_batch_pd(m, pool, npool, nenv, nsteps, qpos, qvel, qacc_warmstart, qfrc_applied, ddq, dq, controlled, do_inverse,
                           qfrc_inverse_out, nullptr, nullptr);
}
