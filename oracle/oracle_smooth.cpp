// oracle_smooth.cpp — fp64 CPU restatement of the smooth-dynamics stages of mj_step1 / mj_step2 / mj_inverse.
// TEST INFRASTRUCTURE ONLY (see oracle.h). PARITY UNPINNED: MuJoCo 2.3.7 is not available; every function restates
// the algorithm published in MuJoCo's "Computation" chapter, reached by the reference only through
// src/mj_main.cpp:83,108, src/mujoco_sim/mj_hw_interface.cpp:61 and src/mujoco_sim/mj_ros.cpp:608,1421.
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"
#include "oracle_util.h"

using namespace omath;

// ---- kinematics ----
void omj_kinematics(const mjModel* m, mjData* d) {
  zero(d->xpos, 3);
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  zero(d->xipos, 3);
  quat2mat(d->ximat, d->xquat);
  for (int i = 1; i < m->nbody; i++) {
    const int pid = m->body_parentid[i];
    const int jntadr = m->body_jntadr[i], jntnum = m->body_jntnum[i];
    mjtNum xpos[3], xquat[4];
    if (jntnum == 1 && m->jnt_type[jntadr] == mjJNT_FREE) {
      const int qadr = m->jnt_qposadr[jntadr];
      copy(xpos, d->qpos + qadr, 3);
      copy(xquat, d->qpos + qadr + 3, 4);
      normalize4(xquat);
      copy(d->qpos + qadr + 3, xquat, 4);  // MuJoCo normalises the stored quaternion in place
      copy(d->xanchor + 3 * jntadr, xpos, 3);
      copy(d->xaxis + 3 * jntadr, m->jnt_axis + 3 * jntadr, 3);
    } else {
      mjtNum bpos[3], bquat[4];
      const int mid = m->body_mocapid[i];
      if (mid >= 0) {
        copy(bpos, d->mocap_pos + 3 * mid, 3);
        copy(bquat, d->mocap_quat + 4 * mid, 4);
        normalize4(bquat);
      } else {
        copy(bpos, m->body_pos + 3 * i, 3);
        copy(bquat, m->body_quat + 4 * i, 4);
      }
      mjtNum vec[3];
      mulMatVec3(vec, d->xmat + 9 * pid, bpos);
      for (int k = 0; k < 3; k++) xpos[k] = d->xpos[3 * pid + k] + vec[k];
      mulQuat(xquat, d->xquat + 4 * pid, bquat);
      for (int j = jntadr; j < jntadr + jntnum; j++) {
        const int qadr = m->jnt_qposadr[j];
        mjtNum* xanchor = d->xanchor + 3 * j;
        mjtNum* xaxis = d->xaxis + 3 * j;
        rotVecQuat(xanchor, m->jnt_pos + 3 * j, xquat);
        for (int k = 0; k < 3; k++) xanchor[k] += xpos[k];
        rotVecQuat(xaxis, m->jnt_axis + 3 * j, xquat);
        switch (m->jnt_type[j]) {
          case mjJNT_SLIDE: {
            const mjtNum dq = d->qpos[qadr] - m->qpos0[qadr];
            for (int k = 0; k < 3; k++) xpos[k] += xaxis[k] * dq;
            break;
          }
          case mjJNT_BALL:
          case mjJNT_HINGE: {
            mjtNum qloc[4], qn[4];
            if (m->jnt_type[j] == mjJNT_BALL) {
              copy(qloc, d->qpos + qadr, 4);
              normalize4(qloc);
              copy(d->qpos + qadr, qloc, 4);
            } else {
              axisAngle2Quat(qloc, m->jnt_axis + 3 * j, d->qpos[qadr] - m->qpos0[qadr]);
            }
            mulQuat(qn, xquat, qloc);
            copy(xquat, qn, 4);
            // keep the anchor fixed: rotation happens about the joint position, not the body origin
            rotVecQuat(vec, m->jnt_pos + 3 * j, xquat);
            for (int k = 0; k < 3; k++) xpos[k] = xanchor[k] - vec[k];
            break;
          }
          default: break;
        }
      }
    }
    normalize4(xquat);
    copy(d->xpos + 3 * i, xpos, 3);
    copy(d->xquat + 4 * i, xquat, 4);
    quat2mat(d->xmat + 9 * i, xquat);
  }
  // inertial frames and geoms
  for (int i = 1; i < m->nbody; i++) {
    mjtNum vec[3], q[4];
    mulMatVec3(vec, d->xmat + 9 * i, m->body_ipos + 3 * i);
    for (int k = 0; k < 3; k++) d->xipos[3 * i + k] = d->xpos[3 * i + k] + vec[k];
    mulQuat(q, d->xquat + 4 * i, m->body_iquat + 4 * i);
    quat2mat(d->ximat + 9 * i, q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    const int b = m->geom_bodyid[g];
    mjtNum vec[3], q[4];
    mulMatVec3(vec, d->xmat + 9 * b, m->geom_pos + 3 * g);
    for (int k = 0; k < 3; k++) d->geom_xpos[3 * g + k] = d->xpos[3 * b + k] + vec[k];
    mulQuat(q, d->xquat + 4 * b, m->geom_quat + 4 * g);
    quat2mat(d->geom_xmat + 9 * g, q);
  }
}

// ---- CoM-based quantities (A.3) ----
void omj_comPos(const mjModel* m, mjData* d) {
  const int nb = m->nbody;
  for (int i = 0; i < nb; i++)
    for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] = m->body_mass[i] * d->xipos[3 * i + k];
  for (int i = nb - 1; i > 0; i--) {
    const int p = m->body_parentid[i];
    for (int k = 0; k < 3; k++) d->subtree_com[3 * p + k] += d->subtree_com[3 * i + k];
  }
  for (int i = 0; i < nb; i++) {
    if (m->body_subtreemass[i] < mjMINVAL) copy(d->subtree_com + 3 * i, d->xipos + 3 * i, 3);
    else for (int k = 0; k < 3; k++) d->subtree_com[3 * i + k] /= m->body_subtreemass[i];
  }
  zero(d->cinert, 10);
  for (int i = 1; i < nb; i++) {
    mjtNum off[3];
    const mjtNum* c = d->subtree_com + 3 * m->body_rootid[i];
    for (int k = 0; k < 3; k++) off[k] = d->xipos[3 * i + k] - c[k];
    inertCom(d->cinert + 10 * i, m->body_inertia + 3 * i, d->ximat + 9 * i, off, m->body_mass[i]);
  }
  for (int j = 0; j < m->njnt; j++) {
    const int bi = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    mjtNum off[3];
    const mjtNum* c = d->subtree_com + 3 * m->body_rootid[bi];
    for (int k = 0; k < 3; k++) off[k] = c[k] - d->xanchor[3 * j + k];
    int skip = 0;
    switch (m->jnt_type[j]) {
      case mjJNT_FREE:
        for (int k = 0; k < 3; k++) {
          zero(d->cdof + 6 * (da + k), 6);
          d->cdof[6 * (da + k) + 3 + k] = 1;
        }
        skip = 3;
        // fall through: rotational dofs of a free joint behave like a ball joint
      case mjJNT_BALL:
        for (int k = 0; k < 3; k++) {
          mjtNum ax[3] = {d->xmat[9 * bi + k], d->xmat[9 * bi + 3 + k], d->xmat[9 * bi + 6 + k]};
          dofCom(d->cdof + 6 * (da + skip + k), ax, off);
        }
        break;
      case mjJNT_SLIDE:
        dofCom(d->cdof + 6 * da, d->xaxis + 3 * j, nullptr);
        break;
      case mjJNT_HINGE:
        dofCom(d->cdof + 6 * da, d->xaxis + 3 * j, off);
        break;
    }
  }
}

// ---- composite rigid body algorithm (A.4) ----
void omj_crb(const mjModel* m, mjData* d) {
  const int nb = m->nbody, nv = m->nv;
  copy(d->crb, d->cinert, 10 * nb);
  for (int i = nb - 1; i > 0; i--) {
    const int p = m->body_parentid[i];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * i + k];
  }
  zero(d->qM, m->nM);
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    mjtNum buf[6];
    mulInertVec(buf, d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
    d->qM[adr] = m->dof_armature[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j]) d->qM[adr++] += dot(d->cdof + 6 * j, buf, 6);
  }
}

static void factor_sparse(const mjModel* m, mjtNum* LD, mjtNum* diaginv) {
  const int nv = m->nv;
  for (int k = nv - 1; k >= 0; k--) {
    const int Mkk = m->dof_Madr[k];
    int Mki = Mkk + 1;
    for (int i = m->dof_parentid[k]; i >= 0; i = m->dof_parentid[i], Mki++) {
      const mjtNum tmp = LD[Mki] / LD[Mkk];
      int Mij = m->dof_Madr[i], Mkj = Mki;
      for (int j = i; j >= 0; j = m->dof_parentid[j]) LD[Mij++] -= LD[Mkj++] * tmp;
      LD[Mki] = tmp;
    }
    diaginv[k] = 1.0 / LD[Mkk];
  }
}

static void solve_sparse(const mjModel* m, const mjtNum* LD, const mjtNum* diaginv, mjtNum* x) {
  const int nv = m->nv;
  for (int i = nv - 1; i >= 0; i--) {  // x <- L^-T x
    if (x[i] == 0) continue;
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= LD[adr++] * x[i];
  }
  for (int i = 0; i < nv; i++) x[i] *= diaginv[i];
  for (int i = 0; i < nv; i++) {  // x <- L^-1 x
    int adr = m->dof_Madr[i] + 1;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[i] -= LD[adr++] * x[j];
  }
}

void omj_factorM(const mjModel* m, mjData* d) {
  copy(d->qLD, d->qM, m->nM);
  factor_sparse(m, d->qLD, d->qLDiagInv);
}

void omj_solveM(const mjModel* m, const mjData* d, mjtNum* x, int n) {
  for (int r = 0; r < n; r++) solve_sparse(m, d->qLD, d->qLDiagInv, x + (size_t)r * m->nv);
}

// symmetric sparse product res = M vec (mj_mulM, called by the reference at src/mujoco_sim/mj_sim.cpp:1057)
void omj_mulM(const mjModel* m, const mjData* d, mjtNum* res, const mjtNum* vec) {
  const int nv = m->nv;
  zero(res, nv);
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

void omj_fullM(const mjModel* m, const mjData* d, mjtNum* dst) {
  const int nv = m->nv;
  zero(dst, nv * nv);
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j], adr++) dst[i * nv + j] = dst[j * nv + i] = d->qM[adr];
  }
}

// translational / rotational Jacobian of a world point rigidly attached to `body`
void omj_jac(const mjModel* m, const mjData* d, mjtNum* jacp, mjtNum* jacr, const mjtNum point[3], int body) {
  const int nv = m->nv;
  if (jacp) zero(jacp, 3 * nv);
  if (jacr) zero(jacr, 3 * nv);
  mjtNum off[3];
  const mjtNum* c = d->subtree_com + 3 * m->body_rootid[body];
  for (int k = 0; k < 3; k++) off[k] = point[k] - c[k];
  while (body > 0 && m->body_dofnum[body] == 0) body = m->body_parentid[body];
  if (body <= 0) return;
  for (int i = m->body_dofadr[body] + m->body_dofnum[body] - 1; i >= 0; i = m->dof_parentid[i]) {
    const mjtNum* cd = d->cdof + 6 * i;
    if (jacr) for (int k = 0; k < 3; k++) jacr[k * nv + i] = cd[k];
    if (jacp) {
      mjtNum t[3];
      cross(t, cd, off);
      for (int k = 0; k < 3; k++) jacp[k * nv + i] = cd[3 + k] + t[k];
    }
  }
}

// qfrc += J_p^T force + J_r^T torque for a wrench applied at `point` on `body`
static void apply_ft(const mjModel* m, const mjData* d, const mjtNum* force, const mjtNum* torque,
                     const mjtNum* point, int body, mjtNum* qfrc) {
  const int nv = m->nv;
  std::vector<mjtNum> jp(3 * (size_t)nv), jr(3 * (size_t)nv);
  omj_jac(m, d, jp.data(), jr.data(), point, body);
  for (int i = 0; i < nv; i++) {
    mjtNum s = 0;
    if (force) for (int k = 0; k < 3; k++) s += jp[k * nv + i] * force[k];
    if (torque) for (int k = 0; k < 3; k++) s += jr[k * nv + i] * torque[k];
    qfrc[i] += s;
  }
}

// ---- velocity stage (A.5) ----
void omj_comVel(const mjModel* m, mjData* d) {
  zero(d->cvel, 6);
  for (int i = 1; i < m->nbody; i++) {
    mjtNum cvel[6];
    copy(cvel, d->cvel + 6 * m->body_parentid[i], 6);
    const int da = m->body_dofadr[i];
    for (int j = m->body_jntadr[i]; j < m->body_jntadr[i] + m->body_jntnum[i]; j++) {
      int dof = m->jnt_dofadr[j];
      (void)da;
      switch (m->jnt_type[j]) {
        case mjJNT_FREE:
          for (int k = 0; k < 3; k++) {
            zero(d->cdof_dot + 6 * (dof + k), 6);
            for (int r = 0; r < 6; r++) cvel[r] += d->cdof[6 * (dof + k) + r] * d->qvel[dof + k];
          }
          dof += 3;
          // fall through
        case mjJNT_BALL:
          for (int k = 0; k < 3; k++) crossMotion(d->cdof_dot + 6 * (dof + k), cvel, d->cdof + 6 * (dof + k));
          for (int k = 0; k < 3; k++)
            for (int r = 0; r < 6; r++) cvel[r] += d->cdof[6 * (dof + k) + r] * d->qvel[dof + k];
          break;
        default:
          crossMotion(d->cdof_dot + 6 * dof, cvel, d->cdof + 6 * dof);
          for (int r = 0; r < 6; r++) cvel[r] += d->cdof[6 * dof + r] * d->qvel[dof];
      }
    }
    copy(d->cvel + 6 * i, cvel, 6);
  }
}

void omj_passive(const mjModel* m, mjData* d) {
  const int nv = m->nv;
  zero(d->qfrc_passive, nv);
  if (m->opt.disableflags & mjDSBL_PASSIVE) return;
  for (int j = 0; j < m->njnt; j++) {
    const mjtNum k = m->jnt_stiffness[j];
    if (k == 0) continue;
    const int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case mjJNT_FREE:
        for (int r = 0; r < 3; r++) d->qfrc_passive[da + r] = -k * (d->qpos[qa + r] - m->qpos_spring[qa + r]);
        // fall through with the rotational part
        {
          mjtNum q[4], dif[3];
          copy(q, d->qpos + qa + 3, 4);
          normalize4(q);
          subQuat(dif, q, m->qpos_spring + qa + 3);
          for (int r = 0; r < 3; r++) d->qfrc_passive[da + 3 + r] = -k * dif[r];
        }
        break;
      case mjJNT_BALL: {
        mjtNum q[4], dif[3];
        copy(q, d->qpos + qa, 4);
        normalize4(q);
        subQuat(dif, q, m->qpos_spring + qa);
        for (int r = 0; r < 3; r++) d->qfrc_passive[da + r] = -k * dif[r];
        break;
      }
      default:
        d->qfrc_passive[da] = -k * (d->qpos[qa] - m->qpos_spring[qa]);
    }
  }
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] -= m->dof_damping[i] * d->qvel[i];
  // gravity compensation: the reference sets gravcomp="1" on every robot body by default
  // (src/mujoco_sim/mj_sim.cpp:301-310, src/config/robot.yaml:19)
  if (!(m->opt.disableflags & mjDSBL_GRAVITY)) {
    for (int i = 1; i < m->nbody; i++) {
      const mjtNum gc = m->body_gravcomp[i];
      if (gc == 0) continue;
      mjtNum f[3];
      for (int k = 0; k < 3; k++) f[k] = -m->opt.gravity[k] * m->body_mass[i] * gc;
      apply_ft(m, d, f, nullptr, d->xipos + 3 * i, i, d->qfrc_passive);
    }
  }
}

// recursive Newton-Euler in the CoM-based frame; flg_acc adds cdof*qacc (inverse dynamics)
void omj_rne(const mjModel* m, mjData* d, int flg_acc, mjtNum* result) {
  const int nb = m->nbody, nv = m->nv;
  std::vector<mjtNum> cacc(6 * (size_t)nb, 0.0), cfrc(6 * (size_t)nb, 0.0);
  if (!(m->opt.disableflags & mjDSBL_GRAVITY)) for (int k = 0; k < 3; k++) cacc[3 + k] = -m->opt.gravity[k];
  for (int i = 1; i < nb; i++) {
    mjtNum* a = &cacc[6 * i];
    copy(a, &cacc[6 * m->body_parentid[i]], 6);
    const int da = m->body_dofadr[i];
    for (int j = 0; j < m->body_dofnum[i]; j++) {
      for (int r = 0; r < 6; r++) a[r] += d->cdof_dot[6 * (da + j) + r] * d->qvel[da + j];
      if (flg_acc) for (int r = 0; r < 6; r++) a[r] += d->cdof[6 * (da + j) + r] * d->qacc[da + j];
    }
    mjtNum Ia[6], Iv[6], vxIv[6];
    mulInertVec(Ia, d->cinert + 10 * i, a);
    mulInertVec(Iv, d->cinert + 10 * i, d->cvel + 6 * i);
    crossForce(vxIv, d->cvel + 6 * i, Iv);
    for (int r = 0; r < 6; r++) cfrc[6 * i + r] = Ia[r] + vxIv[r];
  }
  for (int i = nb - 1; i > 0; i--) {
    const int p = m->body_parentid[i];
    if (p > 0) for (int r = 0; r < 6; r++) cfrc[6 * p + r] += cfrc[6 * i + r];
  }
  for (int i = 0; i < nv; i++) result[i] = dot(d->cdof + 6 * i, &cfrc[6 * m->dof_bodyid[i]], 6);
  copy(d->cacc, cacc.data(), 6 * nb);
  copy(d->cfrc_int, cfrc.data(), 6 * nb);
}

// mj_rnePostConstraint (MuJoCo "Computation": RNE with the final acceleration and all external forces): body
// accelerations cacc and interaction forces cfrc_int, in the c-frame ([rotational; translational] about the subtree CoM of
// the tree root).  External forces: xfrc_applied at the body's inertial frame origin and the contact forces of the
// current solution (pyramid multipliers -> contact-frame force -> world wrench at the contact point; geom2's body
// receives it, geom1's body its opposite).  What force / torque sensors read (reference mj_ros.cpp:1940-1966).
void omj_rnePostConstraint(const mjModel* m, mjData* d) {
  const int nb = m->nbody;
  std::vector<mjtNum> ext(6 * (size_t)nb, 0.0);
  auto wrench = [&](int body, const mjtNum* point, const mjtNum* f, const mjtNum* t, mjtNum sgn) {
    if (body <= 0) return;
    const mjtNum* com = d->subtree_com + 3 * m->body_rootid[body];
    mjtNum arm[3] = {point[0] - com[0], point[1] - com[1], point[2] - com[2]}, mom[3];
    cross(mom, arm, f);
    for (int k = 0; k < 3; k++) { ext[6 * body + k] += sgn * (mom[k] + (t ? t[k] : 0)); ext[6 * body + 3 + k] += sgn * f[k]; }
  };
  for (int i = 1; i < nb; i++) {
    const mjtNum* w = d->xfrc_applied + 6 * i;
    if (w[0] || w[1] || w[2] || w[3] || w[4] || w[5]) wrench(i, d->xipos + 3 * i, w, w + 3, 1);
  }
  for (int c = 0; c < d->ncon; c++) {
    const mjContact& con = d->contact[c];
    if (con.efc_address < 0) continue;
    mjtNum cf[6] = {0, 0, 0, 0, 0, 0};
    const mjtNum* lam = d->efc_force + con.efc_address;
    if (con.dim == 1) cf[0] = lam[0];
    else {
      for (int i = 0; i < 2 * (con.dim - 1); i++) cf[0] += lam[i];
      for (int i = 1; i < con.dim; i++) cf[i] = con.friction[i - 1] * (lam[2 * i - 2] - lam[2 * i - 1]);
    }
    mjtNum fw[3], tw[3];
    for (int k = 0; k < 3; k++) {
      fw[k] = con.frame[k] * cf[0] + con.frame[3 + k] * cf[1] + con.frame[6 + k] * cf[2];
      tw[k] = con.frame[k] * cf[3] + con.frame[3 + k] * cf[4] + con.frame[6 + k] * cf[5];
    }
    wrench(m->geom_bodyid[con.geom1], con.pos, fw, tw, -1);
    wrench(m->geom_bodyid[con.geom2], con.pos, fw, tw, 1);
  }
  zero(d->cacc, 6); zero(d->cfrc_int, 6);
  if (!(m->opt.disableflags & mjDSBL_GRAVITY)) for (int k = 0; k < 3; k++) d->cacc[3 + k] = -m->opt.gravity[k];
  for (int i = 1; i < nb; i++) {
    mjtNum* a = d->cacc + 6 * i;
    copy(a, d->cacc + 6 * m->body_parentid[i], 6);
    for (int j = m->body_dofadr[i]; j < m->body_dofadr[i] + m->body_dofnum[i]; j++)
      for (int r = 0; r < 6; r++) a[r] += d->cdof_dot[6 * j + r] * d->qvel[j] + d->cdof[6 * j + r] * d->qacc[j];
    mjtNum Ia[6], Iv[6], vIv[6];
    mulInertVec(Ia, d->cinert + 10 * i, a);
    mulInertVec(Iv, d->cinert + 10 * i, d->cvel + 6 * i);
    crossForce(vIv, d->cvel + 6 * i, Iv);
    for (int r = 0; r < 6; r++) d->cfrc_int[6 * i + r] = Ia[r] + vIv[r] - ext[6 * (size_t)i + r];
  }
  for (int i = nb - 1; i > 0; i--)
    for (int r = 0; r < 6; r++) d->cfrc_int[6 * m->body_parentid[i] + r] += d->cfrc_int[6 * i + r];
}

void omj_fwdVelocity(const mjModel* m, mjData* d) {
  omj_comVel(m, d);
  omj_passive(m, d);
  omj_referenceConstraint(m, d);
  omj_rne(m, d, 0, d->qfrc_bias);
}

// ---- acceleration stage ----
void omj_fwdAcceleration(const mjModel* m, mjData* d) {
  const int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i];
  for (int b = 1; b < m->nbody; b++) {
    const mjtNum* w = d->xfrc_applied + 6 * b;
    bool any = false;
    for (int k = 0; k < 6; k++) any |= (w[k] != 0);
    if (any) apply_ft(m, d, w, w + 3, d->xipos + 3 * b, b, d->qfrc_smooth);
  }
  copy(d->qacc_smooth, d->qfrc_smooth, nv);
  omj_solveM(m, d, d->qacc_smooth, 1);
}

// semi-implicit Euler with implicit joint damping (A.9)
void omj_Euler(const mjModel* m, mjData* d) {
  const int nv = m->nv;
  const mjtNum h = m->opt.timestep;
  std::vector<mjtNum> qacc(nv);
  bool damp = false;
  if (!(m->opt.disableflags & mjDSBL_EULERDAMP))
    for (int i = 0; i < nv; i++) damp |= m->dof_damping[i] > 0;
  if (!damp) {
    copy(qacc.data(), d->qacc, nv);
  } else {
    std::vector<mjtNum> H(m->nM), Hinv(nv);
    copy(H.data(), d->qM, m->nM);
    for (int i = 0; i < nv; i++) H[m->dof_Madr[i]] += h * m->dof_damping[i];
    factor_sparse(m, H.data(), Hinv.data());
    for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    solve_sparse(m, H.data(), Hinv.data(), qacc.data());
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    const int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case mjJNT_FREE:
        for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
        quatIntegrate(d->qpos + qa + 3, d->qvel + da + 3, h);
        break;
      case mjJNT_BALL:
        quatIntegrate(d->qpos + qa, d->qvel + da, h);
        break;
      default:
        d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time += h;
}

void omj_energy(const mjModel* m, mjData* d) {
  // potential: gravity + springs (hinge/slide only here); kinetic: 0.5 v' M v
  mjtNum pot = 0;
  if (!(m->opt.disableflags & mjDSBL_GRAVITY))
    for (int i = 1; i < m->nbody; i++) pot -= m->body_mass[i] * dot(m->opt.gravity, d->xipos + 3 * i, 3);
  for (int j = 0; j < m->njnt; j++) {
    const mjtNum k = m->jnt_stiffness[j];
    if (k == 0) continue;
    const int qa = m->jnt_qposadr[j];
    if (m->jnt_type[j] == mjJNT_HINGE || m->jnt_type[j] == mjJNT_SLIDE) {
      const mjtNum dq = d->qpos[qa] - m->qpos_spring[qa];
      pot += 0.5 * k * dq * dq;
    }
  }
  std::vector<mjtNum> Mv(m->nv);
  omj_mulM(m, d, Mv.data(), d->qvel);
  d->energy[0] = pot;
  d->energy[1] = 0.5 * dot(Mv.data(), d->qvel, m->nv);
}

// ---- in-tree hot functions of the reference ----
void omj_controller(const mjModel* m, mjData* d, mjtNum* ddq, mjtNum* dq, const mjtByte* controlled) {
  // src/mujoco_sim/mj_sim.cpp:1055-1077
  const int nv = m->nv;
  std::vector<mjtNum> tau(nv);
  omj_mulM(m, d, tau.data(), ddq);                                           // :1057
  for (int i = 0; i < nv; i++) if (controlled && controlled[i]) tau[i] += d->qfrc_bias[i];  // :1058-1063
  copy(d->qfrc_applied, tau.data(), nv);                                     // :1065
  for (int i = 0; i < nv; i++) if (std::fabs(dq[i]) > mjMINVAL) d->qvel[i] = dq[i];       // :1067-1073
  zero(ddq, nv);                                                             // :1075
  zero(dq, nv);                                                              // :1076
}

void omj_set_odom_vels(const mjModel* m, mjData* d, const int lin_dof[3], const int ang_dof[3], const int ang_qpos[3],
                       const mjtNum vels[6]) {
  // src/mujoco_sim/mj_sim.cpp:1079-1153: rows of the ZYX rotation matrix built from the three angular odom joints
  (void)m;
  const mjtNum x = ang_qpos[0] >= 0 ? d->qpos[ang_qpos[0]] : 0, y = ang_qpos[1] >= 0 ? d->qpos[ang_qpos[1]] : 0,
               z = ang_qpos[2] >= 0 ? d->qpos[ang_qpos[2]] : 0;
  const mjtNum cx = std::cos(x), sx = std::sin(x), cy = std::cos(y), sy = std::sin(y), cz = std::cos(z), sz = std::sin(z);
  if (lin_dof[0] >= 0) d->qvel[lin_dof[0]] = vels[0] * cy * cz + vels[1] * (sx * sy * cz - cx * sz) + vels[2] * (cx * sy * cz + sx * sz);
  if (lin_dof[1] >= 0) d->qvel[lin_dof[1]] = vels[0] * cy * sz + vels[1] * (sx * sy * sz + cx * cz) + vels[2] * (cx * sy * sz - sx * cz);
  if (lin_dof[2] >= 0) d->qvel[lin_dof[2]] = -vels[0] * sy + vels[1] * sx * cy + vels[2] * cx * cy;
  for (int k = 0; k < 3; k++) if (ang_dof[k] >= 0) d->qvel[ang_dof[k]] = vels[3 + k];
}
