// set0.cpp — compile-time constants that need the model evaluated at qpos0:
// body_subtreemass, dof_invweight0, body_invweight0, stat.meaninertia, weld relpose.
// Uses a deliberately different formulation from the run-time kernels (dense M = sum_b J_b^T I_b J_b,
// dense Cholesky) so that it doubles as an independent cross-check of CRBA in the tests.
#include <cmath>
#include <vector>

#include "hostmath.h"
#include "model_store.h"

namespace b2 {
using namespace hm;

// Forward kinematics on the host in double precision (MuJoCo "Computation" chapter, kinematics).
void host_kinematics(const mjModel* m, const double* qpos, const double* mocap_pos, const double* mocap_quat,
                     double* xpos, double* xquat, double* xmat, double* xipos, double* ximat, double* xanchor,
                     double* xaxis) {
  zero3(xpos);
  xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
  quat2mat(xmat, xquat);
  zero3(xipos);
  quat2mat(ximat, xquat);
  for (int b = 1; b < m->nbody; b++) {
    const int p = m->body_parentid[b];
    double pos[3], quat[4];
    const int jn = m->body_jntnum[b], ja = m->body_jntadr[b];
    if (jn == 1 && m->jnt_type[ja] == mjJNT_FREE) {
      const int qa = m->jnt_qposadr[ja];
      copy3(pos, qpos + qa);
      copy4(quat, qpos + qa + 3);
      normalize4(quat);
      copy3(xanchor + 3 * ja, pos);
      copy3(xaxis + 3 * ja, m->jnt_axis + 3 * ja);
    } else {
      const int mid = m->body_mocapid[b];
      double bp[3], bq[4];
      if (mid >= 0 && mocap_pos) {
        copy3(bp, mocap_pos + 3 * mid);
        copy4(bq, mocap_quat + 4 * mid);
        normalize4(bq);
      } else {
        copy3(bp, m->body_pos + 3 * b);
        copy4(bq, m->body_quat + 4 * b);
      }
      double r[3];
      mul_mat_vec3(r, xmat + 9 * p, bp);
      for (int k = 0; k < 3; k++) pos[k] = xpos[3 * p + k] + r[k];
      mul_quat(quat, xquat + 4 * p, bq);
      for (int j = ja; j < ja + jn; j++) {
        const int qa = m->jnt_qposadr[j];
        double anchor[3], axis[3];
        rot_vec_quat(anchor, m->jnt_pos + 3 * j, quat);
        for (int k = 0; k < 3; k++) anchor[k] += pos[k];
        rot_vec_quat(axis, m->jnt_axis + 3 * j, quat);
        copy3(xanchor + 3 * j, anchor);
        copy3(xaxis + 3 * j, axis);
        const int t = m->jnt_type[j];
        if (t == mjJNT_SLIDE) {
          const double dq = qpos[qa] - m->qpos0[qa];
          for (int k = 0; k < 3; k++) pos[k] += axis[k] * dq;
        } else if (t == mjJNT_BALL || t == mjJNT_HINGE) {
          double ql[4], qn[4];
          if (t == mjJNT_BALL) { copy4(ql, qpos + qa); normalize4(ql); }
          else axis_angle2quat(ql, m->jnt_axis + 3 * j, qpos[qa] - m->qpos0[qa]);
          mul_quat(qn, quat, ql);
          copy4(quat, qn);
          double off[3];
          rot_vec_quat(off, m->jnt_pos + 3 * j, quat);
          for (int k = 0; k < 3; k++) pos[k] = anchor[k] - off[k];
        }
      }
    }
    normalize4(quat);
    copy3(xpos + 3 * b, pos);
    copy4(xquat + 4 * b, quat);
    quat2mat(xmat + 9 * b, quat);
    double r[3], qi[4];
    mul_mat_vec3(r, xmat + 9 * b, m->body_ipos + 3 * b);
    for (int k = 0; k < 3; k++) xipos[3 * b + k] = pos[k] + r[k];
    mul_quat(qi, quat, m->body_iquat + 4 * b);
    quat2mat(ximat + 9 * b, qi);
  }
}

// 6 x nv Jacobian (rows 0-2 translational at `point`, rows 3-5 rotational) of body b, world frame.
void host_jacobian(const mjModel* m, const double* xmat, const double* xanchor, const double* xaxis, int b,
                   const double* point, double* J /* 6*nv */) {
  const int nv = m->nv;
  for (int i = 0; i < 6 * nv; i++) J[i] = 0;
  while (b > 0) {
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      const int da = m->jnt_dofadr[j], t = m->jnt_type[j];
      const double* anchor = xanchor + 3 * j;
      auto rot_col = [&](int dof, const double* ax) {
        double r[3] = {point[0] - anchor[0], point[1] - anchor[1], point[2] - anchor[2]}, lin[3];
        cross(lin, ax, r);
        for (int k = 0; k < 3; k++) { J[k * nv + dof] = lin[k]; J[(3 + k) * nv + dof] = ax[k]; }
      };
      if (t == mjJNT_FREE) {
        for (int k = 0; k < 3; k++) J[k * nv + da + k] = 1;
        for (int k = 0; k < 3; k++) { double ax[3] = {xmat[9 * b + k], xmat[9 * b + 3 + k], xmat[9 * b + 6 + k]}; rot_col(da + 3 + k, ax); }
      } else if (t == mjJNT_BALL) {
        for (int k = 0; k < 3; k++) { double ax[3] = {xmat[9 * b + k], xmat[9 * b + 3 + k], xmat[9 * b + 6 + k]}; rot_col(da + k, ax); }
      } else if (t == mjJNT_SLIDE) {
        for (int k = 0; k < 3; k++) J[k * nv + da] = xaxis[3 * j + k];
      } else {
        rot_col(da, xaxis + 3 * j);
      }
    }
    b = m->body_parentid[b];
  }
}

// dense joint-space inertia at configuration given by the kinematics arrays
void host_dense_mass(const mjModel* m, const double* xmat, const double* xipos, const double* ximat,
                     const double* xanchor, const double* xaxis, double* M /* nv*nv */) {
  const int nv = m->nv;
  for (int i = 0; i < nv * nv; i++) M[i] = 0;
  std::vector<double> J(6 * (size_t)nv);
  for (int b = 1; b < m->nbody; b++) {
    const double mass = m->body_mass[b];
    host_jacobian(m, xmat, xanchor, xaxis, b, xipos + 3 * b, J.data());
    // world-frame inertia R diag(I) R^T
    const double* R = ximat + 9 * b;
    double Iw[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += R[3 * i + k] * m->body_inertia[3 * b + k] * R[3 * j + k];
        Iw[3 * i + j] = s;
      }
    for (int a = 0; a < nv; a++)
      for (int c = 0; c <= a; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += mass * J[k * nv + a] * J[k * nv + c];
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) s += J[(3 + i) * nv + a] * Iw[3 * i + j] * J[(3 + j) * nv + c];
        M[a * nv + c] += s;
        if (c != a) M[c * nv + a] += s;
      }
  }
  for (int a = 0; a < nv; a++) M[a * nv + a] += m->dof_armature[a];
}

// in-place Cholesky (lower) and inverse; returns false if not positive definite
static bool dense_inverse_spd(std::vector<double>& A, int n, std::vector<double>& inv) {
  std::vector<double> L(A);
  for (int j = 0; j < n; j++) {
    double d = L[j * n + j];
    for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
    if (d <= 0) return false;
    d = std::sqrt(d);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = L[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / d;
    }
  }
  inv.assign((size_t)n * n, 0.0);
  std::vector<double> y(n);
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < n; i++) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
      y[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
      double s = y[i];
      for (int k = i + 1; k < n; k++) s -= L[k * n + i] * inv[k * n + c];
      inv[i * n + c] = s / L[i * n + i];
    }
  }
  return true;
}

void set_const(ModelStore& S) {
  mjModel* m = &S.view;
  const int nb = m->nbody, nv = m->nv, nj = m->njnt;
  // subtree masses
  for (int b = 0; b < nb; b++) S.body_subtreemass[b] = S.body_mass[b];
  for (int b = nb - 1; b > 0; b--) S.body_subtreemass[S.body_parentid[b]] += S.body_subtreemass[b];

  std::vector<double> xpos(3 * nb), xquat(4 * nb), xmat(9 * nb), xipos(3 * nb), ximat(9 * nb), xanchor(3 * nj + 3),
      xaxis(3 * nj + 3);
  host_kinematics(m, m->qpos0, nullptr, nullptr, xpos.data(), xquat.data(), xmat.data(), xipos.data(), ximat.data(),
                  xanchor.data(), xaxis.data());

  // weld relpose from qpos0 when not authored
  for (int q = 0; q < m->neq; q++) {
    double* data = &S.eq_data[(size_t)q * mjNEQDATA];
    if (S.eq_type[q] == mjEQ_CONNECT) {
      // anchor authored in body1 (data[0:3]); the matching anchor in body2 (data[3:6]) comes from qpos0
      const int b1 = S.eq_obj1id[q], b2 = S.eq_obj2id[q];
      double w[3], dd[3];
      mul_mat_vec3(w, &xmat[9 * b1], data);
      for (int k = 0; k < 3; k++) dd[k] = xpos[3 * b1 + k] + w[k] - xpos[3 * b2 + k];
      mul_matT_vec3(data + 3, &xmat[9 * b2], dd);
      continue;
    }
    if (S.eq_type[q] != mjEQ_WELD) continue;
    if (!std::isnan(data[3])) continue;
    const int b1 = S.eq_obj1id[q], b2 = S.eq_obj2id[q];
    // pose of body2 in the frame of body1
    double d[3] = {xpos[3 * b2] - xpos[3 * b1], xpos[3 * b2 + 1] - xpos[3 * b1 + 1], xpos[3 * b2 + 2] - xpos[3 * b1 + 2]};
    mul_matT_vec3(data + 3, &xmat[9 * b1], d);
    double qi[4];
    neg_quat(qi, &xquat[4 * b1]);
    mul_quat(data + 6, qi, &xquat[4 * b2]);
  }
  m->stat.meaninertia = 1;
  m->stat.meanmass = 0;
  {
    double tot = 0; int cnt = 0;
    for (int b = 1; b < nb; b++) if (S.body_mass[b] > 0) { tot += S.body_mass[b]; cnt++; }
    m->stat.meanmass = cnt ? tot / cnt : 0;
  }
  if (nv == 0) return;

  std::vector<double> M((size_t)nv * nv), Minv;
  host_dense_mass(m, xmat.data(), xipos.data(), ximat.data(), xanchor.data(), xaxis.data(), M.data());
  double tr = 0;
  for (int i = 0; i < nv; i++) tr += M[(size_t)i * nv + i];
  m->stat.meaninertia = tr / nv > mjMINVAL ? tr / nv : 1.0;
  if (!dense_inverse_spd(M, nv, Minv)) {
    // singular inertia (e.g. massless chain): leave invweights at zero; R falls back to mjMINVAL
    return;
  }
  // dof_invweight0: diagonal of M^-1, averaged over the dofs of ball joints and over each half of free joints
  for (int j = 0; j < nj; j++) {
    const int da = m->jnt_dofadr[j];
    auto avg = [&](int lo, int n) {
      double s = 0;
      for (int k = 0; k < n; k++) s += Minv[(size_t)(lo + k) * nv + lo + k];
      for (int k = 0; k < n; k++) S.dof_invweight0[lo + k] = s / n;
    };
    switch (m->jnt_type[j]) {
      case mjJNT_FREE: avg(da, 3); avg(da + 3, 3); break;
      case mjJNT_BALL: avg(da, 3); break;
      default: avg(da, 1);
    }
  }
  // body_invweight0: mean diagonal of J M^-1 J^T, translational and rotational blocks, J at the body CoM
  std::vector<double> J(6 * (size_t)nv), JMi(6 * (size_t)nv);
  for (int b = 1; b < nb; b++) {
    if (S.body_weldid[b] == 0) continue;  // static body
    host_jacobian(m, xmat.data(), xanchor.data(), xaxis.data(), b, &xipos[3 * b], J.data());
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < nv; c++) {
        double s = 0;
        for (int k = 0; k < nv; k++) s += J[(size_t)r * nv + k] * Minv[(size_t)k * nv + c];
        JMi[(size_t)r * nv + c] = s;
      }
    double tran = 0, rot = 0;
    for (int r = 0; r < 6; r++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += JMi[(size_t)r * nv + k] * J[(size_t)r * nv + k];
      (r < 3 ? tran : rot) += s / 3.0;
    }
    S.body_invweight0[2 * b] = tran;
    S.body_invweight0[2 * b + 1] = rot;
  }
}

}  // namespace b2
