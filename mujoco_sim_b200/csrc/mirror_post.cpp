// mirror_post.cpp — derived quantities of the legacy mjData view: what the reference's publishers and
// viewer read besides the state (src/mujoco_sim/mj_ros.cpp:1639-1966 sensor / wrench publishers read cfrc_int through
// force / torque sensors, mj_visual.cpp:176 prints d->energy) is computed HERE, on the host, for the one environment that
// is mirrored at publisher rate (<= 60 Hz) — not per tick for every environment on the GPU.
//   xipos, ximat, cinert, cvel, cdof_dot     (position / velocity stage by-products the GPU keeps in registers)
//   cacc, cfrc_ext, cfrc_int                 (mj_rnePostConstraint: body accelerations and interaction forces with the
//                                             applied wrenches and the contact forces of the mirrored tick)
//   energy[0 .. 1]                           (potential incl. joint springs, kinetic 1/2 v^T M v)
// Spatial vectors are [rotational; translational] about the subtree centre of mass of the body's kinematic tree root,
// world axes (MuJoCo's "c-frame").
#include <cmath>
#include <vector>

#include "hostmath.h"
#include "model_store.h"

namespace b2 {
namespace {
using namespace hm;
void mul_inert(double* res, const double* i, const double* v) {
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
void cross_motion(double* res, const double* vel, const double* v) {
  double a[3], b[3], c[3];
  cross(a, vel, v); cross(b, vel, v + 3); cross(c, vel + 3, v);
  for (int k = 0; k < 3; k++) { res[k] = a[k]; res[3 + k] = b[k] + c[k]; }
}
void cross_force(double* res, const double* vel, const double* f) {
  double a[3], b[3], c[3];
  cross(a, vel, f); cross(b, vel + 3, f + 3); cross(c, vel, f + 3);
  for (int k = 0; k < 3; k++) { res[k] = a[k] + b[k]; res[3 + k] = c[k]; }
}
// wrench (force, torque) applied at world point p on body b -> c-frame spatial force accumulated into cfrc_ext
void add_wrench(const mjModel* m, const mjData* d, double* cfrc_ext, int b, const double* p, const double* force, const double* torque, double sign) {
  if (b <= 0) return;
  const double* com = d->subtree_com + 3 * m->body_rootid[b];
  const double off[3] = {p[0] - com[0], p[1] - com[1], p[2] - com[2]};
  double t[3];
  cross(t, off, force);
  for (int k = 0; k < 3; k++) {
    cfrc_ext[6 * b + k] += sign * (t[k] + (torque ? torque[k] : 0.0));
    cfrc_ext[6 * b + 3 + k] += sign * force[k];
  }
}
}  // namespace

void mirror_post(const mjModel* m, mjData* d, const double* xfrc_applied) {
  const int nb = m->nbody, nv = m->nv;
  // ---- inertial frames and c-frame inertias ----
  for (int b = 0; b < nb; b++) {
    double r[3], qi[4];
    mul_mat_vec3(r, d->xmat + 9 * b, m->body_ipos + 3 * b);
    for (int k = 0; k < 3; k++) d->xipos[3 * b + k] = d->xpos[3 * b + k] + r[k];
    mul_quat(qi, d->xquat + 4 * b, m->body_iquat + 4 * b);
    quat2mat(d->ximat + 9 * b, qi);
  }
  for (int k = 0; k < 10; k++) d->cinert[k] = 0;
  for (int b = 1; b < nb; b++) {
    const double* mat = d->ximat + 9 * b;
    const double* in = m->body_inertia + 3 * b;
    const double mass = m->body_mass[b];
    double dif[3];
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[3 * b + k] - d->subtree_com[3 * m->body_rootid[b] + k];
    double* ci = d->cinert + 10 * b;
    auto rit = [&](int r0, int r1) { return mat[3 * r0] * in[0] * mat[3 * r1] + mat[3 * r0 + 1] * in[1] * mat[3 * r1 + 1] + mat[3 * r0 + 2] * in[2] * mat[3 * r1 + 2]; };
    ci[0] = rit(0, 0) + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    ci[1] = rit(1, 1) + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    ci[2] = rit(2, 2) + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    ci[3] = rit(0, 1) - mass * dif[0] * dif[1];
    ci[4] = rit(0, 2) - mass * dif[0] * dif[2];
    ci[5] = rit(1, 2) - mass * dif[1] * dif[2];
    ci[6] = mass * dif[0]; ci[7] = mass * dif[1]; ci[8] = mass * dif[2]; ci[9] = mass;
  }
  // ---- velocities: cvel, cdof_dot (joint by joint; the three rotational dofs of a ball / free joint see the same cvel) ----
  for (int k = 0; k < 6; k++) d->cvel[k] = 0;
  for (int b = 1; b < nb; b++) {
    double cv[6];
    for (int k = 0; k < 6; k++) cv[k] = d->cvel[6 * m->body_parentid[b] + k];
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int dof = m->jnt_dofadr[j];
      const int jt = m->jnt_type[j];
      if (jt == mjJNT_FREE) {
        for (int k = 0; k < 3; k++) {
          for (int r = 0; r < 6; r++) { d->cdof_dot[6 * (dof + k) + r] = 0; cv[r] += d->cdof[6 * (dof + k) + r] * d->qvel[dof + k]; }
        }
        dof += 3;
      }
      const int n = (jt == mjJNT_FREE || jt == mjJNT_BALL) ? 3 : 1;
      for (int k = 0; k < n; k++) cross_motion(d->cdof_dot + 6 * (dof + k), cv, d->cdof + 6 * (dof + k));
      for (int k = 0; k < n; k++)
        for (int r = 0; r < 6; r++) cv[r] += d->cdof[6 * (dof + k) + r] * d->qvel[dof + k];
    }
    for (int k = 0; k < 6; k++) d->cvel[6 * b + k] = cv[k];
  }
  // ---- external forces on the bodies: applied wrenches and the contacts of this tick ----
  std::vector<double> ext(6 * (size_t)nb, 0.0);
  if (xfrc_applied)
    for (int b = 1; b < nb; b++) {
      const double* w = xfrc_applied + 6 * b;
      bool any = false;
      for (int k = 0; k < 6; k++) any |= w[k] != 0;
      if (any) add_wrench(m, d, ext.data(), b, d->xipos + 3 * b, w, w + 3, 1.0);
    }
  for (int c = 0; c < d->ncon; c++) {
    const mjContact& k = d->contact[c];
    if (k.efc_address < 0) continue;
    // contact-frame force from the pyramid multipliers: normal = sum f, direction i = mu_i (f_{2i-2} - f_{2i-1})
    double res[6] = {0, 0, 0, 0, 0, 0};
    const double* f = d->efc_force + k.efc_address;
    if (k.dim == 1) res[0] = f[0];
    else
      for (int i = 0; i < k.dim - 1; i++) {
        res[0] += f[2 * i] + f[2 * i + 1];
        res[1 + i] = (f[2 * i] - f[2 * i + 1]) * k.friction[i];
      }
    double fw[3] = {0, 0, 0}, tw[3] = {0, 0, 0};
    for (int r = 0; r < 3; r++)
      for (int kk = 0; kk < 3; kk++) { fw[kk] += k.frame[3 * r + kk] * res[r]; tw[kk] += k.frame[3 * r + kk] * res[3 + r]; }
    // the contact pushes geom2's body along the normal and geom1's body the opposite way
    add_wrench(m, d, ext.data(), m->geom_bodyid[k.geom1], k.pos, fw, tw, -1.0);
    add_wrench(m, d, ext.data(), m->geom_bodyid[k.geom2], k.pos, fw, tw, 1.0);
  }
  // ---- accelerations and interaction forces ----
  for (int k = 0; k < 6; k++) d->cacc[k] = 0;
  if (!(m->opt.disableflags & mjDSBL_GRAVITY)) for (int k = 0; k < 3; k++) d->cacc[3 + k] = -m->opt.gravity[k];
  for (int k = 0; k < 6; k++) d->cfrc_int[k] = 0;
  for (int b = 1; b < nb; b++) {
    double* a = d->cacc + 6 * b;
    for (int k = 0; k < 6; k++) a[k] = d->cacc[6 * m->body_parentid[b] + k];
    const int da = m->body_dofadr[b];
    for (int j = 0; j < m->body_dofnum[b]; j++)
      for (int r = 0; r < 6; r++) a[r] += d->cdof_dot[6 * (da + j) + r] * d->qvel[da + j] + d->cdof[6 * (da + j) + r] * d->qacc[da + j];
    double Ia[6], Iv[6], vxIv[6];
    mul_inert(Ia, d->cinert + 10 * b, a);
    mul_inert(Iv, d->cinert + 10 * b, d->cvel + 6 * b);
    cross_force(vxIv, d->cvel + 6 * b, Iv);
    for (int r = 0; r < 6; r++) d->cfrc_int[6 * b + r] = Ia[r] + vxIv[r] - ext[6 * (size_t)b + r];
  }
  for (int b = nb - 1; b > 0; b--) {
    const int p = m->body_parentid[b];
    for (int r = 0; r < 6; r++) d->cfrc_int[6 * p + r] += d->cfrc_int[6 * b + r];
  }
  // ---- energy ----
  double pot = 0;
  if (!(m->opt.disableflags & mjDSBL_GRAVITY))
    for (int b = 1; b < nb; b++) pot -= m->body_mass[b] * dot3(m->opt.gravity, d->xipos + 3 * b);
  for (int j = 0; j < m->njnt; j++) {
    const double ks = m->jnt_stiffness[j];
    if (ks == 0) continue;
    const int qa = m->jnt_qposadr[j], jt = m->jnt_type[j];
    if (jt == mjJNT_SLIDE || jt == mjJNT_HINGE) { const double dq = d->qpos[qa] - m->qpos_spring[qa]; pot += 0.5 * ks * dq * dq; }
  }
  double kin = 0;
  for (int i = 0; i < nv; i++) {
    int adr = m->dof_Madr[i];
    kin += 0.5 * d->qM[adr] * d->qvel[i] * d->qvel[i];
    adr++;
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j], adr++) kin += d->qM[adr] * d->qvel[i] * d->qvel[j];
  }
  d->energy[0] = pot;
  d->energy[1] = kin;
}

}  // namespace b2
