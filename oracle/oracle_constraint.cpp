// oracle_constraint.cpp — fp64 CPU restatement of constraint assembly, impedance, projection, the PGS solve
// and the inverse-dynamics constraint force.  TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED:
// restates MuJoCo 2.3.7's published soft-constraint model ("Computation" chapter: Constraint model, Solver
// parameters, PGS) because libmujoco is not available; reached by the reference only through
// src/mj_main.cpp:83,108 (mj_step1/mj_step2) and src/mujoco_sim/mj_hw_interface.cpp:61 (mj_inverse).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"
#include "oracle_util.h"

using namespace omath;

namespace {

struct RowAdder {
  const mjModel* m;
  mjData* d;
  bool overflow = false;
  // append one row; returns the row index or -1 when njmax is exhausted
  int add(int type, int id, const mjtNum* J, mjtNum pos, mjtNum margin, mjtNum frictionloss) {
    if (d->nefc >= m->njmax) { overflow = true; return -1; }
    const int r = d->nefc++;
    d->efc_type[r] = type;
    d->efc_id[r] = id;
    copy(d->efc_J + (size_t)r * m->nv, J, m->nv);
    d->efc_pos[r] = pos;
    d->efc_margin[r] = margin;
    d->efc_frictionloss[r] = frictionloss;
    return r;
  }
};

// difference of the point Jacobians of two bodies at a common world point: (body2 - body1)
void jac_dif_pair(const mjModel* m, const mjData* d, const mjtNum* point, int b1, int b2, mjtNum* jp, mjtNum* jr) {
  const int nv = m->nv;
  std::vector<mjtNum> p1(3 * (size_t)nv), r1(3 * (size_t)nv);
  omj_jac(m, d, p1.data(), r1.data(), point, b1);
  omj_jac(m, d, jp, jr, point, b2);
  for (int i = 0; i < 3 * nv; i++) { jp[i] -= p1[i]; jr[i] -= r1[i]; }
}

// impedance d(r) and its role in R (MuJoCo docs, "Solver parameters": solimp = dmin dmax width midpoint power)
mjtNum impedance(const mjtNum* solimp, mjtNum pos, mjtNum margin) {
  mjtNum dmin = std::min(mjMAXIMP, std::max(mjMINIMP, solimp[0]));
  mjtNum dmax = std::min(mjMAXIMP, std::max(mjMINIMP, solimp[1]));
  const mjtNum width = solimp[2];
  const mjtNum mid = std::min(mjMAXIMP, std::max(mjMINIMP, solimp[3]));
  const mjtNum power = std::max((mjtNum)1, solimp[4]);
  if (dmin == dmax || width <= mjMINVAL) return 0.5 * (dmin + dmax);
  const mjtNum x = std::fabs(pos - margin) / width;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  mjtNum y;
  if (power == 1) y = x;
  else if (x <= mid) y = std::pow(x, power) / std::pow(mid, power - 1);
  else y = 1 - std::pow(1 - x, power) / std::pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}

}  // namespace

// ---- constraint rows in MuJoCo order: equality, friction loss, limits, contacts ----
void omj_makeConstraint(const mjModel* m, mjData* d) {
  const int nv = m->nv;
  d->nefc = d->ne = d->nf = 0;
  for (int c = 0; c < d->ncon; c++) d->contact[c].efc_address = -1;
  if (m->opt.disableflags & mjDSBL_CONSTRAINT) return;
  RowAdder A{m, d};
  std::vector<mjtNum> J(nv), jp(3 * (size_t)nv), jr(3 * (size_t)nv), jp2(3 * (size_t)nv), jr2(3 * (size_t)nv);

  // equality
  if (!(m->opt.disableflags & mjDSBL_EQUALITY)) {
    for (int q = 0; q < m->neq; q++) {
      if (!m->eq_active[q]) continue;
      const mjtNum* data = m->eq_data + (size_t)q * mjNEQDATA;
      const int o1 = m->eq_obj1id[q], o2 = m->eq_obj2id[q];
      if (m->eq_type[q] == mjEQ_JOINT) {
        // q1 - q1_0 = poly(q2 - q2_0); produced by the reference's URDF mimic conversion (src/mujoco_compile.cpp:219-248)
        zero(J.data(), nv);
        const int qa1 = m->jnt_qposadr[o1], da1 = m->jnt_dofadr[o1];
        mjtNum pos = d->qpos[qa1] - m->qpos0[qa1], ref = data[0];
        J[da1] = 1;
        if (o2 >= 0) {
          const int qa2 = m->jnt_qposadr[o2], da2 = m->jnt_dofadr[o2];
          const mjtNum dif = d->qpos[qa2] - m->qpos0[qa2];
          ref = data[0] + dif * (data[1] + dif * (data[2] + dif * (data[3] + dif * data[4])));
          const mjtNum deriv = data[1] + dif * (2 * data[2] + dif * (3 * data[3] + dif * 4 * data[4]));
          J[da2] = -deriv;
        }
        if (A.add(mjCNSTR_EQUALITY, q, J.data(), pos - ref, 0, 0) >= 0) d->ne++;
      } else if (m->eq_type[q] == mjEQ_CONNECT || m->eq_type[q] == mjEQ_WELD) {
        // anchor on body1 = data[0:3] (connect) or relpose position data[3:6] (weld); anchor on body2 from qpos0
        const bool weld = m->eq_type[q] == mjEQ_WELD;
        mjtNum a1[3], a2[3], p1[3], p2[3];
        if (weld) { copy(a1, data + 3, 3); copy(a2, data, 3); }
        else { copy(a1, data, 3); copy(a2, data + 3, 3); }
        mulMatVec3(p1, d->xmat + 9 * o1, a1);
        mulMatVec3(p2, d->xmat + 9 * o2, a2);
        for (int k = 0; k < 3; k++) { p1[k] += d->xpos[3 * o1 + k]; p2[k] += d->xpos[3 * o2 + k]; }
        omj_jac(m, d, jp.data(), jr.data(), p1, o1);
        omj_jac(m, d, jp2.data(), jr2.data(), p2, o2);
        for (int k = 0; k < 3; k++) {
          for (int i = 0; i < nv; i++) J[i] = jp[k * nv + i] - jp2[k * nv + i];
          if (A.add(mjCNSTR_EQUALITY, q, J.data(), p1[k] - p2[k], 0, 0) >= 0) d->ne++;
        }
        if (weld) {
          // orientation residual: imaginary part of q2^-1 * q1 * relpose, scaled by torquescale (data[10]);
          // Jacobian = torquescale * 0.5 * (jacr1 - jacr2) expressed through the same quaternion product
          const mjtNum ts = data[10];
          mjtNum q1r[4], q2n[4] = {d->xquat[4 * o2], -d->xquat[4 * o2 + 1], -d->xquat[4 * o2 + 2], -d->xquat[4 * o2 + 3]}, qe[4];
          mulQuat(q1r, d->xquat + 4 * o1, data + 6);
          mulQuat(qe, q2n, q1r);
          for (int k = 0; k < 3; k++) {
            for (int i = 0; i < nv; i++) {
              // d(qe)/dt for angular velocity difference w (world frame): 0.5 * q2^-1 * (0,w) * q1r
              const mjtNum w[4] = {0, jr[0 * nv + i] - jr2[0 * nv + i], jr[1 * nv + i] - jr2[1 * nv + i], jr[2 * nv + i] - jr2[2 * nv + i]};
              mjtNum t1[4], t2[4];
              mulQuat(t1, q2n, w);
              mulQuat(t2, t1, q1r);
              J[i] = 0.5 * ts * t2[1 + k];
            }
            if (A.add(mjCNSTR_EQUALITY, q, J.data(), ts * qe[1 + k], 0, 0) >= 0) d->ne++;
          }
        }
      }
    }
  }

  // dof friction loss (tiago.xml: 16 joints with frictionloss=1)
  if (!(m->opt.disableflags & mjDSBL_FRICTIONLOSS)) {
    for (int i = 0; i < nv; i++) {
      if (m->dof_frictionloss[i] <= 0) continue;
      zero(J.data(), nv);
      J[i] = 1;
      if (A.add(mjCNSTR_FRICTION_DOF, i, J.data(), 0, 0, m->dof_frictionloss[i]) >= 0) d->nf++;
    }
  }

  // joint limits: one row per side that is within margin
  if (!(m->opt.disableflags & mjDSBL_LIMIT)) {
    for (int j = 0; j < m->njnt; j++) {
      if (!m->jnt_limited[j]) continue;
      const int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
      const mjtNum margin = m->jnt_margin[j];
      if (m->jnt_type[j] == mjJNT_SLIDE || m->jnt_type[j] == mjJNT_HINGE) {
        const mjtNum value = d->qpos[qa];
        for (int side = -1; side <= 1; side += 2) {
          const mjtNum dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - value);
          if (dist < margin) {
            zero(J.data(), nv);
            J[da] = -side;
            A.add(mjCNSTR_LIMIT_JOINT, j, J.data(), dist, margin, 0);
          }
        }
      } else if (m->jnt_type[j] == mjJNT_BALL) {
        mjtNum q[4], axis[3];
        copy(q, d->qpos + qa, 4);
        normalize4(q);
        copy(axis, q + 1, 3);
        const mjtNum s = normalize3(axis);
        mjtNum angle = 2 * std::atan2(s, q[0]);
        if (angle > mjPI) angle -= 2 * mjPI;
        if (angle < 0) { angle = -angle; for (mjtNum& a : axis) a = -a; }
        const mjtNum dist = std::max(m->jnt_range[2 * j], m->jnt_range[2 * j + 1]) - angle;
        if (dist < margin) {
          zero(J.data(), nv);
          for (int k = 0; k < 3; k++) J[da + k] = -axis[k];
          A.add(mjCNSTR_LIMIT_JOINT, j, J.data(), dist, margin, 0);
        }
      }
    }
  }

  // contacts: pyramidal friction cone, 2*(dim-1) rows per contact (1 row when frictionless)
  if (!(m->opt.disableflags & mjDSBL_CONTACT)) {
    for (int c = 0; c < d->ncon; c++) {
      mjContact* con = d->contact + c;
      if (con->exclude) continue;
      const int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
      jac_dif_pair(m, d, con->pos, b1, b2, jp.data(), jr.data());
      // rotate into the contact frame: rows of `frame` are normal, tangent1, tangent2
      std::vector<mjtNum> fp(3 * (size_t)nv), fr(3 * (size_t)nv);
      for (int r = 0; r < 3; r++)
        for (int i = 0; i < nv; i++) {
          fp[r * nv + i] = con->frame[3 * r] * jp[i] + con->frame[3 * r + 1] * jp[nv + i] + con->frame[3 * r + 2] * jp[2 * nv + i];
          fr[r * nv + i] = con->frame[3 * r] * jr[i] + con->frame[3 * r + 1] * jr[nv + i] + con->frame[3 * r + 2] * jr[2 * nv + i];
        }
      const int first = d->nefc;
      const int nrow = con->dim == 1 ? 1 : 2 * (con->dim - 1);
      if (first + nrow > m->njmax) { A.overflow = true; continue; }  // a contact is kept whole or dropped whole
      if (con->dim == 1) {
        A.add(mjCNSTR_CONTACT_FRICTIONLESS, c, fp.data(), con->dist, con->includemargin, 0);
      } else {
        for (int k = 1; k < con->dim; k++) {
          const mjtNum* dir = k < 3 ? &fp[k * nv] : &fr[(k - 3) * nv];
          const mjtNum mu = con->friction[k - 1];
          for (int sgn = 1; sgn >= -1; sgn -= 2) {
            for (int i = 0; i < nv; i++) J[i] = fp[i] + sgn * mu * dir[i];
            A.add(mjCNSTR_CONTACT_PYRAMIDAL, c, J.data(), con->dist, con->includemargin, 0);
          }
        }
      }
      if (d->nefc > first) con->efc_address = first;
    }
  }

  // diagApprox, impedance -> R, D, KBIP
  const mjtNum h = m->opt.timestep;
  for (int r = 0; r < d->nefc; r++) {
    const int id = d->efc_id[r];
    const mjtNum *solref, *solimp;
    mjtNum diag;
    switch (d->efc_type[r]) {
      case mjCNSTR_EQUALITY:
        solref = m->eq_solref + 2 * id; solimp = m->eq_solimp + 5 * id;
        if (m->eq_type[id] == mjEQ_JOINT) {
          diag = m->dof_invweight0[m->jnt_dofadr[m->eq_obj1id[id]]];
          if (m->eq_obj2id[id] >= 0) diag += m->dof_invweight0[m->jnt_dofadr[m->eq_obj2id[id]]];
        } else {
          // rows 0-2 translational, 3-5 rotational (weld)
          int k = 0;
          for (int rr = r - 1; rr >= 0 && d->efc_type[rr] == mjCNSTR_EQUALITY && d->efc_id[rr] == id; rr--) k++;
          const int b1 = m->eq_obj1id[id], b2 = m->eq_obj2id[id];
          diag = m->body_invweight0[2 * b1 + (k >= 3)] + m->body_invweight0[2 * b2 + (k >= 3)];
        }
        break;
      case mjCNSTR_FRICTION_DOF:
        solref = m->dof_solref + 2 * id; solimp = m->dof_solimp + 5 * id;
        diag = m->dof_invweight0[id];
        break;
      case mjCNSTR_LIMIT_JOINT:
        solref = m->jnt_solref + 2 * id; solimp = m->jnt_solimp + 5 * id;
        diag = m->dof_invweight0[m->jnt_dofadr[id]];
        break;
      default: {
        const mjContact* con = d->contact + id;
        solref = con->solref; solimp = con->solimp;
        const int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
        const mjtNum tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
        const mjtNum rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
        if (d->efc_type[r] == mjCNSTR_CONTACT_FRICTIONLESS) diag = tran;
        else {
          const int j = r - con->efc_address;
          const mjtNum fri = con->friction[j / 2];
          diag = tran + fri * fri * (j < 4 ? tran : rot);
        }
      }
    }
    d->efc_diagApprox[r] = diag;
    const mjtNum imp = impedance(solimp, d->efc_pos[r], d->efc_margin[r]);
    d->efc_R[r] = std::max(mjMINVAL, (1 - imp) * diag / imp);
    mjtNum* kbip = d->efc_KBIP + 4 * r;
    const mjtNum dmax = std::min(mjMAXIMP, std::max(mjMINIMP, solimp[1]));
    if (solref[0] > 0) {
      mjtNum tc = solref[0];
      const mjtNum dr = solref[1];
      if (!(m->opt.disableflags & mjDSBL_REFSAFE)) tc = std::max(tc, 2 * h);
      kbip[0] = 1 / std::max(mjMINVAL, dmax * dmax * tc * tc * dr * dr);
      kbip[1] = 2 / std::max(mjMINVAL, dmax * tc);
    } else {
      kbip[0] = -solref[0] / std::max(mjMINVAL, dmax * dmax);
      kbip[1] = -solref[1] / std::max(mjMINVAL, dmax);
    }
    kbip[2] = imp;
    kbip[3] = 0;
    if (d->efc_type[r] == mjCNSTR_FRICTION_DOF) kbip[0] = 0;  // friction loss has no position term
  }
  // pyramidal cones: every row of a contact shares R = 2 mu^2 R_first (mu = friction[0] / sqrt(impratio))
  for (int c = 0; c < d->ncon; c++) {
    const mjContact* con = d->contact + c;
    if (con->efc_address < 0 || con->dim == 1) continue;
    const int a = con->efc_address;
    const mjtNum mu = con->friction[0] / std::sqrt(std::max(mjMINVAL, m->opt.impratio));
    const mjtNum Rpy = std::max(mjMINVAL, 2 * mu * mu * d->efc_R[a]);
    for (int j = 0; j < 2 * (con->dim - 1) && a + j < d->nefc; j++) d->efc_R[a + j] = Rpy;
  }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1 / d->efc_R[r];
}

// AR = J M^-1 J^T + diag(R)
void omj_projectConstraint(const mjModel* m, mjData* d) {
  const int nv = m->nv, ne = d->nefc;
  if (!ne) return;
  std::vector<mjtNum> MiJT((size_t)ne * nv);
  copy(MiJT.data(), d->efc_J, ne * nv);
  omj_solveM(m, d, MiJT.data(), ne);  // row r = M^-1 J_r^T
  for (int r = 0; r < ne; r++)
    for (int c = 0; c <= r; c++) {
      const mjtNum v = dot(d->efc_J + (size_t)r * nv, MiJT.data() + (size_t)c * nv, nv);
      d->efc_AR[(size_t)r * m->njmax + c] = d->efc_AR[(size_t)c * m->njmax + r] = v;
    }
  for (int r = 0; r < ne; r++) d->efc_AR[(size_t)r * m->njmax + r] += d->efc_R[r];
}

// efc_vel = J qvel ; aref = -B vel - K imp (pos - margin)
void omj_referenceConstraint(const mjModel* m, mjData* d) {
  const int nv = m->nv;
  for (int r = 0; r < d->nefc; r++) {
    d->efc_vel[r] = dot(d->efc_J + (size_t)r * nv, d->qvel, nv);
    const mjtNum* k = d->efc_KBIP + 4 * r;
    d->efc_aref[r] = -k[1] * d->efc_vel[r] - k[0] * k[2] * (d->efc_pos[r] - d->efc_margin[r]);
  }
}

// primal force law: force as a function of the constraint-space acceleration residual jar = J qacc - aref
static void constraint_update(const mjModel* m, mjData* d, const mjtNum* jar) {
  (void)m;
  for (int r = 0; r < d->nefc; r++) {
    const mjtNum D = d->efc_D[r], R = d->efc_R[r];
    switch (d->efc_type[r]) {
      case mjCNSTR_EQUALITY:
        d->efc_force[r] = -D * jar[r];
        break;
      case mjCNSTR_FRICTION_DOF: {
        const mjtNum f = d->efc_frictionloss[r];
        if (jar[r] <= -R * f) d->efc_force[r] = f;
        else if (jar[r] >= R * f) d->efc_force[r] = -f;
        else d->efc_force[r] = -D * jar[r];
        break;
      }
      default:
        d->efc_force[r] = jar[r] < 0 ? -D * jar[r] : 0;
    }
  }
}

static void mul_JT(const mjModel* m, const mjData* d, mjtNum* res, const mjtNum* f) {
  const int nv = m->nv;
  zero(res, nv);
  for (int r = 0; r < d->nefc; r++)
    if (f[r] != 0) for (int i = 0; i < nv; i++) res[i] += d->efc_J[(size_t)r * nv + i] * f[r];
}

// ---- PGS in the dual (A.8). Fixed schedule: opt.iterations sweeps, early exit on scaled improvement < tolerance ----
void omj_fwdConstraint(const mjModel* m, mjData* d) {
  const int nv = m->nv, ne = d->nefc, ld = m->njmax;
  d->solver_iter = 0;
  if (!ne) {
    copy(d->qacc, d->qacc_smooth, nv);
    copy(d->qacc_warmstart, d->qacc_smooth, nv);
    zero(d->qfrc_constraint, nv);
    return;
  }
  std::vector<mjtNum> jar(ne);
  for (int r = 0; r < ne; r++) {
    d->efc_b[r] = dot(d->efc_J + (size_t)r * nv, d->qacc_smooth, nv) - d->efc_aref[r];
  }
  // warm start: forces implied by qacc_warmstart, kept only if their dual cost is negative (better than f = 0)
  if (!(m->opt.disableflags & mjDSBL_WARMSTART)) {
    for (int r = 0; r < ne; r++) jar[r] = dot(d->efc_J + (size_t)r * nv, d->qacc_warmstart, nv) - d->efc_aref[r];
    constraint_update(m, d, jar.data());
    mjtNum cost = 0;
    for (int r = 0; r < ne; r++) {
      const mjtNum Af = dot(d->efc_AR + (size_t)r * ld, d->efc_force, ne);
      cost += d->efc_force[r] * (0.5 * Af + d->efc_b[r]);
    }
    if (cost > 0) zero(d->efc_force, ne);
  } else {
    zero(d->efc_force, ne);
  }
  const mjtNum scale = 1 / (m->stat.meaninertia * std::max(1, nv));
  for (int it = 0; it < m->opt.iterations; it++) {
    mjtNum improvement = 0;
    for (int r = 0; r < ne; r++) {
      const mjtNum* Ar = d->efc_AR + (size_t)r * ld;
      const mjtNum res = dot(Ar, d->efc_force, ne) + d->efc_b[r];
      const mjtNum old = d->efc_force[r];
      mjtNum f = old - res / Ar[r];
      switch (d->efc_type[r]) {
        case mjCNSTR_EQUALITY: break;
        case mjCNSTR_FRICTION_DOF: f = std::min(d->efc_frictionloss[r], std::max(-d->efc_frictionloss[r], f)); break;
        default: f = std::max((mjtNum)0, f);
      }
      const mjtNum delta = f - old;
      const mjtNum change = 0.5 * delta * delta * Ar[r] + delta * res;  // dual cost change, <= 0 for a valid step
      if (change > 1e-10) continue;                                     // reject: keep the old force
      d->efc_force[r] = f;
      improvement -= change;
    }
    d->solver_iter = it + 1;
    if (improvement * scale < m->opt.tolerance) break;
  }
  mul_JT(m, d, d->qfrc_constraint, d->efc_force);
  copy(d->qacc, d->qfrc_constraint, nv);
  omj_solveM(m, d, d->qacc, 1);
  for (int i = 0; i < nv; i++) d->qacc[i] += d->qacc_smooth[i];
  copy(d->qacc_warmstart, d->qacc, nv);
}

// constraint force for inverse dynamics: primal law evaluated at the given qacc
extern "C" void omj_invConstraint(const mjModel* m, mjData* d) {
  const int nv = m->nv, ne = d->nefc;
  if (!ne) { zero(d->qfrc_constraint, nv); return; }
  std::vector<mjtNum> jar(ne);
  for (int r = 0; r < ne; r++) jar[r] = dot(d->efc_J + (size_t)r * nv, d->qacc, nv) - d->efc_aref[r];
  constraint_update(m, d, jar.data());
  mul_JT(m, d, d->qfrc_constraint, d->efc_force);
}
