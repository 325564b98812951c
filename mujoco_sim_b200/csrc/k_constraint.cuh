// k_constraint.cuh — constraint rows (equality, friction loss, limits, pyramidal contacts), impedance / reference
// acceleration, projection B = M^-1 J^T (+ diag(AR)), the PGS solve fused with the integrate stage.
//
// Row storage is COMPACT: the mass matrix is block diagonal over kinematic trees and a constraint row touches at
// most two trees (contact between two bodies, joint equality, limit, friction loss), so row r keeps only the dofs of
// trees (t1, t2): element k < n1 is dof s1 + k, element k >= n1 is dof s2 + k - n1, width <= DModel.wmax.
// J and B = M^-1 J^T share that layout; [row][k][env] in HBM (env fastest: coalesced for one thread per environment).
// PGS runs in acceleration space: with a = qacc_smooth + sum_r f_r B_r the Gauss-Seidel residual of row r is
// J_r a - aref_r + R_r f_r, identical to AR_r f + b_r of the dual formulation but O(width) instead of O(nefc) per row
// and without ever forming the nefc x nefc matrix (which is still available on demand for the legacy efc_AR field).
// (rows s6, s7(aref), s11, s12, s13, s14 of SURVEY.md section 8a').  Row order and formulas follow MuJoCo's published
// constraint model; the reference reaches them only through mj_step1 / mj_step2 / mj_inverse
// (src/mj_main.cpp:83,108; src/mujoco_sim/mj_hw_interface.cpp:61).
#pragma once
#include "k_args.h"
#include "k_common.cuh"
#include "k_smooth.cuh"

namespace b2 {

enum { CN_EQUALITY = 0, CN_FRICTION_DOF = 1, CN_LIMIT_JOINT = 3, CN_CONTACT_FRICTIONLESS = 5, CN_CONTACT_PYRAMIDAL = 6 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2 };

template <typename T>
__device__ T impedance(const T* solimp, T pos, T margin) {
  const T lo = T(0.0001), hi = T(0.9999);
  const T dmin = t_min(hi, t_max(lo, solimp[0])), dmax = t_min(hi, t_max(lo, solimp[1]));
  const T width = solimp[2], mid = t_min(hi, t_max(lo, solimp[3])), power = t_max(T(1), solimp[4]);
  if (dmin == dmax || width <= Eps<T>::minval()) return T(0.5) * (dmin + dmax);
  const T x = t_abs(pos - margin) / width;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  T y;
  if (power == 1) y = x;
  else if (power == 2) y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
  else if (x <= mid) y = t_pow(x, power) / t_pow(mid, power - 1);
  else y = 1 - t_pow(1 - x, power) / t_pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}


struct Seg { int s1, n1, s2, n2; };

template <typename T>
__device__ __forceinline__ Seg seg_of(const MV<T>& m, int t1, int t2) {
  const DModel& h = *m.h;
  Seg g{0, 0, 0, 0};
  if (t1 >= 0) { g.s1 = m.i(h.o_tree_dofadr, t1); g.n1 = m.i(h.o_tree_dofnum, t1); }
  if (t2 >= 0) { g.s2 = m.i(h.o_tree_dofadr, t2); g.n2 = m.i(h.o_tree_dofnum, t2); }
  return g;
}
// compact position of dof i in a row with segments g (i must lie in one of them)
__device__ __forceinline__ int seg_pos(const Seg& g, int i) { return (i >= g.s1 && i < g.s1 + g.n1) ? i - g.s1 : g.n1 + i - g.s2; }
__device__ __forceinline__ int seg_dof(const Seg& g, int k) { return k < g.n1 ? g.s1 + k : g.s2 + k - g.n1; }

template <typename T>
struct Rows {
  MV<T> m;
  const DModel& h;
  const KArgs<T>& a;
  int env;
  long long S;
  int nefc = 0;
  __device__ Rows(const MV<T>& mv, const KArgs<T>& args, int e) : m(mv), h(*mv.h), a(args), env(e), S(args.nenvp) {}
  __device__ __forceinline__ T& Jc(int r, int k) const { return a.efc_J[((long long)r * h.wmax + k) * S + env]; }
  __device__ __forceinline__ T cdof(int i, int k) const { return a.cdof[(6 * i + k) * S + env]; }

  // start a new zero row over trees (t1, t2); returns its index or -1 when njmax is exhausted
  __device__ int open(int type, int id, T pos, T margin, T frictionloss, int t1, int t2, Seg* gout) {
    if (nefc >= h.njmax) { a.status[env] |= 2; return -1; }
    if (t1 < 0) { t1 = t2; t2 = -1; }
    if (t1 == t2) t2 = -1;
    if (t2 >= 0 && t2 < t1) { const int t = t1; t1 = t2; t2 = t; }
    const int r = nefc++;
    const Seg g = seg_of(m, t1, t2);
    for (int k = 0; k < g.n1 + g.n2; k++) Jc(r, k) = 0;
    a.efc_tree[((long long)2 * r) * S + env] = t1;
    a.efc_tree[((long long)2 * r + 1) * S + env] = t2;
    a.efc_type[(long long)r * S + env] = type;
    a.efc_id[(long long)r * S + env] = id;
    a.efc_pos[(long long)r * S + env] = pos;
    a.efc_margin[(long long)r * S + env] = margin;
    a.efc_frictionloss[(long long)r * S + env] = frictionloss;
    if (gout) *gout = g;
    return r;
  }

  // column i of the 3 x nv translational (at `point`) and rotational Jacobian of a body on whose chain dof i lies
  __device__ __forceinline__ void jac_col(int i, const T* off, T* jp, T* jr) const {
    T cd[6];
    for (int k = 0; k < 6; k++) cd[k] = cdof(i, k);
    T t[3];
    cross3(t, cd, off);
    jp[0] = cd[3] + t[0]; jp[1] = cd[4] + t[1]; jp[2] = cd[5] + t[2];
    jr[0] = cd[0]; jr[1] = cd[1]; jr[2] = cd[2];
  }
  __device__ __forceinline__ void point_off(int b, const T* point, T* off) const {
    const int root = m.i(h.o_body_rootid, b);
    for (int k = 0; k < 3; k++) off[k] = point[k] - a.subtree_com[(3 * root + k) * S + env];
  }

  // rows [r0, r0 + 3) += sign * translational Jacobian of body b at point (world axes)
  __device__ void add_jacp(int r0, const Seg& g, int b, const T* point, T sign) {
    T off[3];
    point_off(b, point, off);
    for (int i = m.i(h.o_body_lastdof, b); i >= 0; i = m.i(h.o_dof_parentid, i)) {
      T jp[3], jr[3];
      jac_col(i, off, jp, jr);
      const int k = seg_pos(g, i);
      for (int c = 0; c < 3; c++) Jc(r0 + c, k) += sign * jp[c];
    }
  }

  __device__ void equality() {
    if (h.disableflags & DSBL_EQUALITY) return;
    for (int q = 0; q < h.neq; q++) {
      if (!m.i(h.o_eq_active, q)) continue;
      const int type = m.i(h.o_eq_type, q), o1 = m.i(h.o_eq_obj1id, q), o2 = m.i(h.o_eq_obj2id, q);
      const T* data = m.fp(h.o_eq_data) + 11 * q;
      Seg g;
      if (type == EQ_JOINT) {
        const int qa1 = m.i(h.o_jnt_qposadr, o1), da1 = m.i(h.o_jnt_dofadr, o1);
        const T pos = a.qpos[qa1 * S + env] - m.f(h.o_qpos0, qa1);
        T ref = data[0], deriv = 0;
        int da2 = -1;
        if (o2 >= 0) {
          const int qa2 = m.i(h.o_jnt_qposadr, o2);
          da2 = m.i(h.o_jnt_dofadr, o2);
          const T dif = a.qpos[qa2 * S + env] - m.f(h.o_qpos0, qa2);
          ref = data[0] + dif * (data[1] + dif * (data[2] + dif * (data[3] + dif * data[4])));
          deriv = data[1] + dif * (2 * data[2] + dif * (3 * data[3] + dif * 4 * data[4]));
        }
        const int r = open(CN_EQUALITY, q, pos - ref, 0, 0, m.i(h.o_dof_treeid, da1), da2 >= 0 ? m.i(h.o_dof_treeid, da2) : -1, &g);
        if (r < 0) continue;
        if (da2 >= 0) Jc(r, seg_pos(g, da2)) = -deriv;
        Jc(r, seg_pos(g, da1)) = 1;
      } else {
        const bool weld = type == EQ_WELD;
        const int t1 = m.i(h.o_body_treeid, o1), t2 = m.i(h.o_body_treeid, o2);
        T a1[3], a2[3], p1[3], p2[3], m1[9], m2[9];
        for (int k = 0; k < 3; k++) { a1[k] = weld ? data[3 + k] : data[k]; a2[k] = weld ? data[k] : data[3 + k]; }
        for (int k = 0; k < 9; k++) { m1[k] = a.xmat[(9 * o1 + k) * S + env]; m2[k] = a.xmat[(9 * o2 + k) * S + env]; }
        mat_vec3(p1, m1, a1);
        mat_vec3(p2, m2, a2);
        for (int k = 0; k < 3; k++) { p1[k] += a.xpos[(3 * o1 + k) * S + env]; p2[k] += a.xpos[(3 * o2 + k) * S + env]; }
        int r0 = -1;
        for (int k = 0; k < 3; k++) { const int r = open(CN_EQUALITY, q, p1[k] - p2[k], 0, 0, t1, t2, &g); if (k == 0) r0 = r; }
        if (r0 < 0 || nefc - r0 < 3) continue;
        add_jacp(r0, g, o1, p1, T(1));
        add_jacp(r0, g, o2, p2, T(-1));
        if (weld) {
          const T ts = data[10];
          T q1[4], q2n[4], q1r[4], qe[4];
          for (int k = 0; k < 4; k++) { q1[k] = a.xquat[(4 * o1 + k) * S + env]; q2n[k] = a.xquat[(4 * o2 + k) * S + env]; }
          q2n[1] = -q2n[1]; q2n[2] = -q2n[2]; q2n[3] = -q2n[3];
          mul_quat(q1r, q1, data + 6);
          mul_quat(qe, q2n, q1r);
          int rr = -1;
          for (int k = 0; k < 3; k++) { const int r = open(CN_EQUALITY, q, ts * qe[1 + k], 0, 0, t1, t2, &g); if (k == 0) rr = r; }
          if (rr < 0 || nefc - rr < 3) continue;
          for (int side = 0; side < 2; side++) {
            const int b = side ? o2 : o1;
            const T sg = side ? T(-1) : T(1);
            for (int i = m.i(h.o_body_lastdof, b); i >= 0; i = m.i(h.o_dof_parentid, i)) {
              const T w[4] = {0, cdof(i, 0), cdof(i, 1), cdof(i, 2)};
              T t1q[4], t2q[4];
              mul_quat(t1q, q2n, w);
              mul_quat(t2q, t1q, q1r);
              const int kk = seg_pos(g, i);
              for (int k = 0; k < 3; k++) Jc(rr + k, kk) += sg * T(0.5) * ts * t2q[1 + k];
            }
          }
        }
      }
    }
  }

  __device__ void friction_loss() {
    if (!h.has_frictionloss || (h.disableflags & DSBL_FRICTIONLOSS)) return;
    for (int i = 0; i < h.nv; i++) {
      const T fl = m.f(h.o_dof_frictionloss, i);
      if (fl <= 0) continue;
      Seg g;
      const int r = open(CN_FRICTION_DOF, i, 0, 0, fl, m.i(h.o_dof_treeid, i), -1, &g);
      if (r >= 0) Jc(r, seg_pos(g, i)) = 1;
    }
  }

  __device__ void limits() {
    if (!h.has_limits || (h.disableflags & DSBL_LIMIT)) return;
    for (int j = 0; j < h.njnt; j++) {
      if (!m.i(h.o_jnt_limited, j)) continue;
      const int qa = m.i(h.o_jnt_qposadr, j), da = m.i(h.o_jnt_dofadr, j), jt = m.i(h.o_jnt_type, j);
      const T margin = m.f(h.o_jnt_margin, j);
      Seg g;
      if (jt == JNT_SLIDE || jt == JNT_HINGE) {
        const T value = a.qpos[qa * S + env];
        for (int side = -1; side <= 1; side += 2) {
          const T dist = side * (m.f(h.o_jnt_range, 2 * j + (side + 1) / 2) - value);
          if (dist < margin) {
            const int r = open(CN_LIMIT_JOINT, j, dist, margin, 0, m.i(h.o_dof_treeid, da), -1, &g);
            if (r >= 0) Jc(r, seg_pos(g, da)) = T(-side);
          }
        }
      } else if (jt == JNT_BALL) {
        T q[4];
        for (int k = 0; k < 4; k++) q[k] = a.qpos[(qa + k) * S + env];
        normalize4(q);
        T ax[3] = {q[1], q[2], q[3]};
        const T s = normalize3(ax);
        T angle = 2 * t_atan2(s, q[0]);
        const T pi = T(3.14159265358979323846);
        if (angle > pi) angle -= 2 * pi;
        if (angle < 0) { angle = -angle; ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; }
        const T dist = t_max(m.f(h.o_jnt_range, 2 * j), m.f(h.o_jnt_range, 2 * j + 1)) - angle;
        if (dist < margin) {
          const int r = open(CN_LIMIT_JOINT, j, dist, margin, 0, m.i(h.o_dof_treeid, da), -1, &g);
          if (r >= 0) for (int k = 0; k < 3; k++) Jc(r, seg_pos(g, da + k)) = -ax[k];
        }
      }
    }
  }

  __device__ void contacts() {
    if (h.disableflags & DSBL_CONTACT) return;
    const int ncon = a.ncon[env];
    for (int c = 0; c < ncon; c++) {
      auto F = [&](int f) -> T { return a.con[((long long)f * h.nconmax + c) * S + env]; };
      auto I = [&](int f) -> int& { return a.coni[((long long)f * h.nconmax + c) * S + env]; };
      const int dim = I(CI_DIM);
      const int b1 = m.i(h.o_geom_bodyid, I(CI_GEOM1)), b2 = m.i(h.o_geom_bodyid, I(CI_GEOM2));
      T pos[3], frame[9], fri[5];
      for (int k = 0; k < 3; k++) pos[k] = F(CF_POS + k);
      for (int k = 0; k < 9; k++) frame[k] = F(CF_FRAME + k);
      for (int k = 0; k < 5; k++) fri[k] = F(CF_FRICTION + k);
      const T dist = F(CF_DIST), im = F(CF_INCLUDEMARGIN);
      const int nrow = dim == 1 ? 1 : 2 * (dim - 1);
      const int first = nefc;
      if (first + nrow > h.njmax) { a.status[env] |= 2; I(CI_EFC) = -1; continue; }  // kept whole or dropped whole
      Seg g;
      const int t1 = m.i(h.o_body_treeid, b1), t2 = m.i(h.o_body_treeid, b2);
      for (int k = 0; k < nrow; k++) open(dim == 1 ? CN_CONTACT_FRICTIONLESS : CN_CONTACT_PYRAMIDAL, c, dist, im, 0, t1, t2, &g);
      I(CI_EFC) = first;
      for (int side = 0; side < 2; side++) {
        const int b = side ? b2 : b1;
        const T sg = side ? T(1) : T(-1);
        T off[3];
        point_off(b, pos, off);
        for (int i = m.i(h.o_body_lastdof, b); i >= 0; i = m.i(h.o_dof_parentid, i)) {
          T jp[3], jr[3], fp[3], fr[3];
          jac_col(i, off, jp, jr);
          mat_vec3(fp, frame, jp);  // rows of frame: normal, tangent1, tangent2
          mat_vec3(fr, frame, jr);
          const int kk = seg_pos(g, i);
          if (dim == 1) { Jc(first, kk) += sg * fp[0]; continue; }
          for (int k = 1; k < dim; k++) {
            const T dir = k < 3 ? fp[k] : fr[k - 3];
            const T mu = fri[k - 1];
            Jc(first + 2 * k - 2, kk) += sg * (fp[0] + mu * dir);
            Jc(first + 2 * k - 1, kk) += sg * (fp[0] - mu * dir);
          }
        }
      }
    }
  }

  // diagApprox, impedance -> R, D; K, B, imp; aref; (vel uses the current, possibly overridden, qvel)
  __device__ void finish() {
    const T hs = a.h;
    for (int r = 0; r < nefc; r++) {
      const int type = a.efc_type[(long long)r * S + env], id = a.efc_id[(long long)r * S + env];
      T solref[2], solimp[5], diag;
      if (type == CN_EQUALITY) {
        for (int k = 0; k < 2; k++) solref[k] = m.f(h.o_eq_solref, 2 * id + k);
        for (int k = 0; k < 5; k++) solimp[k] = m.f(h.o_eq_solimp, 5 * id + k);
        if (m.i(h.o_eq_type, id) == EQ_JOINT) {
          diag = m.f(h.o_dof_invweight0, m.i(h.o_jnt_dofadr, m.i(h.o_eq_obj1id, id)));
          const int o2 = m.i(h.o_eq_obj2id, id);
          if (o2 >= 0) diag += m.f(h.o_dof_invweight0, m.i(h.o_jnt_dofadr, o2));
        } else {
          int k = 0;
          for (int rr = r - 1; rr >= 0 && a.efc_type[(long long)rr * S + env] == CN_EQUALITY && a.efc_id[(long long)rr * S + env] == id; rr--) k++;
          const int rot = k >= 3 ? 1 : 0;
          diag = m.f(h.o_body_invweight0, 2 * m.i(h.o_eq_obj1id, id) + rot) + m.f(h.o_body_invweight0, 2 * m.i(h.o_eq_obj2id, id) + rot);
        }
      } else if (type == CN_FRICTION_DOF) {
        for (int k = 0; k < 2; k++) solref[k] = m.f(h.o_dof_solref, 2 * id + k);
        for (int k = 0; k < 5; k++) solimp[k] = m.f(h.o_dof_solimp, 5 * id + k);
        diag = m.f(h.o_dof_invweight0, id);
      } else if (type == CN_LIMIT_JOINT) {
        for (int k = 0; k < 2; k++) solref[k] = m.f(h.o_jnt_solref, 2 * id + k);
        for (int k = 0; k < 5; k++) solimp[k] = m.f(h.o_jnt_solimp, 5 * id + k);
        diag = m.f(h.o_dof_invweight0, m.i(h.o_jnt_dofadr, id));
      } else {
        auto F = [&](int f) -> T { return a.con[((long long)f * h.nconmax + id) * S + env]; };
        auto I = [&](int f) -> int { return a.coni[((long long)f * h.nconmax + id) * S + env]; };
        for (int k = 0; k < 2; k++) solref[k] = F(CF_SOLREF + k);
        for (int k = 0; k < 5; k++) solimp[k] = F(CF_SOLIMP + k);
        const int b1 = m.i(h.o_geom_bodyid, I(CI_GEOM1)), b2 = m.i(h.o_geom_bodyid, I(CI_GEOM2));
        const T tran = m.f(h.o_body_invweight0, 2 * b1) + m.f(h.o_body_invweight0, 2 * b2);
        const T rot = m.f(h.o_body_invweight0, 2 * b1 + 1) + m.f(h.o_body_invweight0, 2 * b2 + 1);
        if (type == CN_CONTACT_FRICTIONLESS) diag = tran;
        else {
          const int j = r - I(CI_EFC);
          const T fr = F(CF_FRICTION + j / 2);
          diag = tran + fr * fr * (j < 4 ? tran : rot);
        }
      }
      const T pos = a.efc_pos[(long long)r * S + env], margin = a.efc_margin[(long long)r * S + env];
      const T imp = impedance(solimp, pos, margin);
      T R = t_max(Eps<T>::minval(), (1 - imp) * diag / imp);
      a.efc_diagApprox[(long long)r * S + env] = diag;
      a.efc_R[(long long)r * S + env] = R;
      const T dmax = t_min(T(0.9999), t_max(T(0.0001), solimp[1]));
      T K, B;
      if (solref[0] > 0) {
        T tc = solref[0];
        const T dr = solref[1];
        if (!(h.disableflags & DSBL_REFSAFE)) tc = t_max(tc, 2 * hs);
        K = 1 / t_max(Eps<T>::minval(), dmax * dmax * tc * tc * dr * dr);
        B = 2 / t_max(Eps<T>::minval(), dmax * tc);
      } else {
        K = -solref[0] / t_max(Eps<T>::minval(), dmax * dmax);
        B = -solref[1] / t_max(Eps<T>::minval(), dmax);
      }
      if (type == CN_FRICTION_DOF) K = 0;
      a.efc_KBI[((long long)0 * h.njmax + r) * S + env] = K;
      a.efc_KBI[((long long)1 * h.njmax + r) * S + env] = B;
      a.efc_KBI[((long long)2 * h.njmax + r) * S + env] = imp;
    }
    // pyramidal cones share R = 2 mu^2 R_first
    const int ncon = a.ncon[env];
    for (int c = 0; c < ncon; c++) {
      const int adr = a.coni[((long long)CI_EFC * h.nconmax + c) * S + env];
      const int dim = a.coni[((long long)CI_DIM * h.nconmax + c) * S + env];
      if (adr < 0 || dim == 1) continue;
      const T mu = a.con[((long long)CF_FRICTION * h.nconmax + c) * S + env] / t_sqrt(t_max(Eps<T>::minval(), m.f(h.o_opt_real, 5)));
      const T Rpy = t_max(Eps<T>::minval(), 2 * mu * mu * a.efc_R[(long long)adr * S + env]);
      for (int j = 0; j < 2 * (dim - 1); j++) a.efc_R[(long long)(adr + j) * S + env] = Rpy;
    }
  }
};

// primal force law (force from the constraint-space residual jar = J qacc - aref)
template <typename T>
__device__ __forceinline__ T primal_force(int type, T jar, T D, T R, T fl) {
  if (type == CN_EQUALITY) return -D * jar;
  if (type == CN_FRICTION_DOF) {
    if (jar <= -R * fl) return fl;
    if (jar >= R * fl) return -fl;
    return -D * jar;
  }
  return jar < 0 ? -D * jar : T(0);
}

#define B2_KERNEL_PROLOGUE                                                           \
  if ((a.flags & B2F_FUSABLE) && a.pending[0] == 0) return;                 \
  extern __shared__ __align__(16) unsigned char smem_raw[];                          \
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);                             \
  uint32_t* blob = reinterpret_cast<uint32_t*>(smem_raw + 16);                       \
  const int nwords = reinterpret_cast<const DModel*>(a.model)->nwords;               \
  stage_model(blob, a.model, nwords, bar);                                           \
  MV<T> m{reinterpret_cast<const DModel*>(blob), blob};                              \
  const DModel& h = *m.h;                                                            \
  const long long S = a.nenvp;                                                       \
  const int ntiles = a.nenvp / BLOCK;                                                \
  (void)S; (void)h;

// Row slab for the solver, environment-major: row r of environment e is 2 * wp contiguous numbers
// [J compact (wp) | B = M^-1 J^T compact (wp)] at efc_rows[(e * njmax + r) * 2 wp], plus an 8-number record
// {R, aref, diag(AR), frictionloss, type, tree1, tree2, b} at efc_meta[(e * njmax + r) * 8].  One 8-lane team of the
// solver streams these rows with 8/16-byte vector loads; for wp = 16 (C3) a row is exactly one 128-byte line.
enum { META_R = 0, META_AREF, META_DIAG, META_FL, META_TYPE, META_T1, META_T2, META_B, META_N };

// K4 + K5 fused: rows, impedance, then per row: vel, aref, b = J qacc_smooth - aref, B_r = M^-1 J_r^T by sparse
// back-substitution inside the row's tree blocks (in shared memory), diag(AR)_r = J_r B_r + R_r, and (for mj_inverse)
// qfrc_inverse -= J^T f(qacc_prev).  Finished rows leave through a shared-memory transpose so that every global store
// of the slab is a full coalesced line.
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_make_constraint(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  const int WP = a.wp;
  constexpr int LDS = BLOCK + 1;  // +1: conflict-free both for per-thread columns and for the transposed row reads
  T* rowsh = reinterpret_cast<T*>(smem_raw + 16 + (size_t)nwords * 4);  // [2 * WP + META_N][LDS]: J | B | meta
  const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * BLOCK + threadIdx.x;
    Rows<T> rows(m, a, env);
    const bool done = (a.flags & B2F_FUSABLE) && (a.status[env] & 8);  // already integrated by the smooth kernel
    if (!(h.disableflags & DSBL_CONSTRAINT) && !done) {
      rows.equality();
      rows.friction_loss();
      rows.limits();
      rows.contacts();
      rows.finish();
    }
    const int ne = rows.nefc, W = h.wmax;
    if (!done) a.nefc[env] = ne;
    SArr<T> LD{a.qLD + env, S}, dinv{a.qLDiagInv + env, S};
    SArr<T> Jr{rowsh + threadIdx.x, LDS}, Br{rowsh + (size_t)WP * LDS + threadIdx.x, LDS};
    int nemax = ne;
    for (int o = 16; o > 0; o >>= 1) nemax = max(nemax, __shfl_xor_sync(0xffffffffu, nemax, o));
    for (int r = 0; r < nemax; r++) {
      T R = 0, aref = 0, dg = 0, fl = 0, bb = 0;
      int type = 0, t1 = -1, t2 = -1;
      if (r < ne) {
        const long long o = (long long)r * S + env;
        t1 = a.efc_tree[((long long)2 * r) * S + env]; t2 = a.efc_tree[((long long)2 * r + 1) * S + env];
        const Seg g = seg_of(m, t1, t2);
        const int w = g.n1 + g.n2;
        T vel = 0, js = 0, jq = 0;
        for (int k = 0; k < w; k++) {
          const T j = a.efc_J[((long long)r * W + k) * S + env];
          const long long d = (long long)seg_dof(g, k) * S + env;
          vel += j * a.qvel[d];
          js += j * a.qacc_smooth[d];
          jq += j * a.qacc[d];
          Jr[k] = j;
          Br[k] = j;
        }
        for (int k = w; k < WP; k++) { Jr[k] = 0; Br[k] = 0; }
        const T K = a.efc_KBI[((long long)0 * h.njmax + r) * S + env], Bd = a.efc_KBI[((long long)1 * h.njmax + r) * S + env];
        const T imp = a.efc_KBI[((long long)2 * h.njmax + r) * S + env];
        R = a.efc_R[o];
        const T D = 1 / R;
        aref = -Bd * vel - K * imp * (a.efc_pos[o] - a.efc_margin[o]);
        type = a.efc_type[o];
        fl = a.efc_frictionloss[o];
        bb = js - aref;
        a.efc_D[o] = D;
        a.efc_vel[o] = vel;
        a.efc_aref[o] = aref;
        a.efc_b[o] = bb;
        if (a.flags & B2F_INVERSE) {
          const T f = primal_force(type, jq - aref, D, R, fl);
          if (f != 0) for (int k = 0; k < w; k++) a.qfrc_inverse[(long long)seg_dof(g, k) * S + env] -= Jr[k] * f;
        }
        // B_r = M^-1 J_r^T, one tree block at a time (M is block diagonal over trees)
        for (int sgm = 0; sgm < 2; sgm++) {
          const int lo = sgm ? g.s2 : g.s1, n = sgm ? g.n2 : g.n1, base = sgm ? g.n1 : 0;
          if (n == 0) continue;
          for (int i = lo + n - 1; i >= lo; i--) {
            const T xi = Br[base + i - lo];
            if (xi == 0) continue;
            int adr = m.i(h.o_dof_Madr, i) + 1;
            for (int j = m.i(h.o_dof_parentid, i); j >= 0; j = m.i(h.o_dof_parentid, j)) Br[base + j - lo] -= LD[adr++] * xi;
          }
          for (int i = lo; i < lo + n; i++) Br[base + i - lo] *= dinv[i];
          for (int i = lo; i < lo + n; i++) {
            int adr = m.i(h.o_dof_Madr, i) + 1;
            T xi = Br[base + i - lo];
            for (int j = m.i(h.o_dof_parentid, i); j >= 0; j = m.i(h.o_dof_parentid, j)) xi -= LD[adr++] * Br[base + j - lo];
            Br[base + i - lo] = xi;
          }
        }
        dg = R;
        for (int k = 0; k < w; k++) dg += Jr[k] * Br[k];
        a.efc_ARdiag[o] = dg;
      }
      // ---- transpose out: lanes cooperate on one environment's row at a time (coalesced line stores) ----
      {
        SArr<T> Mr{rowsh + (size_t)2 * WP * LDS + threadIdx.x, LDS};
        Mr[META_R] = R; Mr[META_AREF] = aref; Mr[META_DIAG] = dg; Mr[META_FL] = fl;
        Mr[META_TYPE] = (T)type; Mr[META_T1] = (T)t1; Mr[META_T2] = (T)t2; Mr[META_B] = bb;
      }
      __syncwarp();
      const unsigned live = __ballot_sync(0xffffffffu, r < ne);
      for (unsigned rem = live; rem; rem &= rem - 1) {
        const int e = __ffs(rem) - 1;
        const long long env_e = (long long)tile * BLOCK + wbase + e;
        T* dst = a.efc_rows + (env_e * h.njmax + r) * (2 * WP);
        T* dstm = a.efc_meta + (env_e * h.njmax + r) * META_N;
        for (int l = lane; l < 2 * WP + META_N; l += 32) {
          const T v = rowsh[(size_t)l * LDS + wbase + e];
          if (l < 2 * WP) dst[l] = v; else dstm[l - 2 * WP] = v;
        }
      }
      __syncwarp();
    }
  }
}

template <typename T, int N> struct alignas(sizeof(T) * N) VecN { T v[N]; };

// K6: projected Gauss-Seidel (A.8) in acceleration space.  An 8-lane team owns one environment (4 per warp): lane l
// holds elements [l * EPL, (l + 1) * EPL) of the current row of J and B, the running acceleration and the forces live
// in shared memory, rows stream from the slab one ahead of their use.
template <typename T, int EPL, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_pgs_team(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  (void)ntiles;
  constexpr int TEAM = 8, EPB = BLOCK / TEAM;  // environments per CTA
  constexpr int WP = TEAM * EPL;
  const int nv = h.nv, njmax = h.njmax;
  const int nvs = nv + 4;  // stride of the per-environment vectors (skews the teams over the banks)
  T* accsh = reinterpret_cast<T*>(smem_raw + 16 + (size_t)nwords * 4);  // [EPB][nvs] running acceleration
  T* tmpsh = accsh + (size_t)EPB * nvs;                                   // [EPB][nvs] M^-1 J^T f, later qfrc_constraint
  T* fsh = tmpsh + (size_t)EPB * nvs;                                     // [EPB][njmax] forces
  const int team = threadIdx.x / TEAM, l = threadIdx.x % TEAM;
  const unsigned tmask = 0xffu << ((threadIdx.x & 31) & ~7);
  T* acc = accsh + (size_t)team * nvs;
  T* tmp = tmpsh + (size_t)team * nvs;
  T* f = fsh + (size_t)team * njmax;
  const int ngroups = (a.nenvp + EPB - 1) / EPB;
  const T tol = m.f(h.o_opt_real, 3), scale = 1 / (m.f(h.o_opt_real, 4) * T(nv > 1 ? nv : 1));

  auto team_sum = [&](T v) {
    v += __shfl_xor_sync(tmask, v, 4);
    v += __shfl_xor_sync(tmask, v, 2);
    v += __shfl_xor_sync(tmask, v, 1);
    return v;
  };
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long long env = (long long)grp * EPB + team;  // nenvp is a multiple of 128 >= EPB: always in range
    if ((a.flags & B2F_FUSABLE) && (a.status[env] & 8)) continue;  // integrated by the smooth kernel (team-uniform)
    const int ne = a.nefc[env];
    const T* rowp = a.efc_rows + env * njmax * (2 * WP);
    const T* metap = a.efc_meta + env * njmax * META_N;
    int iters = 0;
    // dof index of this lane's elements for a row with trees (t1, t2); -1 for padding
    auto dofs_of = [&](int t1, int t2, int* d) {
      const Seg g = seg_of(m, t1, t2);
#pragma unroll
      for (int k = 0; k < EPL; k++) { const int kk = l * EPL + k; d[k] = kk < g.n1 + g.n2 ? seg_dof(g, kk) : -1; }
    };
    auto load_row = [&](int r, VecN<T, EPL>& J, VecN<T, EPL>& B, VecN<T, META_N>& M) {
      J = *reinterpret_cast<const VecN<T, EPL>*>(rowp + (size_t)r * 2 * WP + l * EPL);
      B = *reinterpret_cast<const VecN<T, EPL>*>(rowp + (size_t)r * 2 * WP + WP + l * EPL);
      M = *reinterpret_cast<const VecN<T, META_N>*>(metap + (size_t)r * META_N);
    };
    for (int i = l; i < nv; i += TEAM) { acc[i] = a.qacc_warmstart[(long long)i * S + env]; tmp[i] = 0; }
    __syncwarp(tmask);
    if (ne > 0) {
      VecN<T, EPL> J, B;
      VecN<T, META_N> M;
      int d[EPL];
      // ---- warm start: forces implied by qacc_warmstart (held in acc), kept only if their dual cost is negative ----
      bool warm = !(h.disableflags & DSBL_WARMSTART);
      if (warm) {
        for (int r = 0; r < ne; r++) {
          load_row(r, J, B, M);
          dofs_of((int)M.v[META_T1], (int)M.v[META_T2], d);
          T p = 0;
#pragma unroll
          for (int k = 0; k < EPL; k++) if (d[k] >= 0) p += J.v[k] * acc[d[k]];
          const T jw = team_sum(p);
          const T R = M.v[META_R];
          const T fr = primal_force((int)M.v[META_TYPE], jw - M.v[META_AREF], 1 / R, R, M.v[META_FL]);
          if (l == 0) f[r] = fr;
          if (fr != 0) {
#pragma unroll
            for (int k = 0; k < EPL; k++) if (d[k] >= 0) tmp[d[k]] += fr * B.v[k];
          }
          __syncwarp(tmask);
        }
        T cost = 0;
        for (int r = 0; r < ne; r++) {
          const T fr = f[r];
          if (fr == 0) continue;
          load_row(r, J, B, M);
          dofs_of((int)M.v[META_T1], (int)M.v[META_T2], d);
          T p = 0;
#pragma unroll
          for (int k = 0; k < EPL; k++) if (d[k] >= 0) p += J.v[k] * tmp[d[k]];
          const T Af = team_sum(p) + M.v[META_R] * fr;
          cost += fr * (T(0.5) * Af + M.v[META_B]);
        }
        if (cost > 0) warm = false;
      }
      __syncwarp(tmask);
      if (warm) {
        for (int i = l; i < nv; i += TEAM) acc[i] = a.qacc_smooth[(long long)i * S + env] + tmp[i];
      } else {
        for (int i = l; i < nv; i += TEAM) acc[i] = a.qacc_smooth[(long long)i * S + env];
        for (int r = l; r < ne; r += TEAM) f[r] = 0;
      }
      __syncwarp(tmask);
      // ---- Gauss-Seidel sweeps; the next row is fetched while the current one is processed ----
      for (int it = 0; it < h.iterations; it++) {
        T improvement = 0;
        VecN<T, EPL> Jn, Bn;
        VecN<T, META_N> Mn;
        load_row(0, Jn, Bn, Mn);
        for (int r = 0; r < ne; r++) {
          J = Jn; B = Bn; M = Mn;
          if (r + 1 < ne) load_row(r + 1, Jn, Bn, Mn);
          dofs_of((int)M.v[META_T1], (int)M.v[META_T2], d);
          T p = 0;
#pragma unroll
          for (int k = 0; k < EPL; k++) if (d[k] >= 0) p += J.v[k] * acc[d[k]];
          const T old = f[r];
          const T res = team_sum(p) + M.v[META_R] * old - M.v[META_AREF];
          const T Arr = M.v[META_DIAG];
          T fn = old - res / Arr;
          const int type = (int)M.v[META_TYPE];
          if (type == CN_FRICTION_DOF) { const T flv = M.v[META_FL]; fn = t_min(flv, t_max(-flv, fn)); }
          else if (type != CN_EQUALITY) fn = t_max(T(0), fn);
          const T delta = fn - old;
          const T change = T(0.5) * delta * delta * Arr + delta * res;
          if (delta != 0 && !(change > T(1e-10))) {
            if (l == 0) f[r] = fn;
            improvement -= change;
#pragma unroll
            for (int k = 0; k < EPL; k++) if (d[k] >= 0) acc[d[k]] += delta * B.v[k];
          }
          __syncwarp(tmask);
        }
        iters = it + 1;
        if (improvement * scale < tol) break;
      }
      // ---- qfrc_constraint = J^T f ----
      for (int i = l; i < nv; i += TEAM) tmp[i] = 0;
      __syncwarp(tmask);
      for (int r = 0; r < ne; r++) {
        const T fr = f[r];
        if (l == 0) a.efc_force[(long long)r * S + env] = fr;
        if (fr == 0) continue;
        load_row(r, J, B, M);
        dofs_of((int)M.v[META_T1], (int)M.v[META_T2], d);
#pragma unroll
        for (int k = 0; k < EPL; k++) if (d[k] >= 0) tmp[d[k]] += J.v[k] * fr;
        __syncwarp(tmask);
      }
    } else {
      for (int i = l; i < nv; i += TEAM) acc[i] = a.qacc_smooth[(long long)i * S + env];
    }
    __syncwarp(tmask);
    for (int i = l; i < nv; i += TEAM) {
      const T v = acc[i];
      a.qacc[(long long)i * S + env] = v;
      a.qacc_warmstart[(long long)i * S + env] = v;
      a.qfrc_constraint[(long long)i * S + env] = tmp[i];
    }
    if (l == 0) a.solver_iter[env] = iters;
    __syncwarp(tmask);
  }
}

// G7: mj_checkAcc, semi-implicit Euler with implicit damping, odom override (one thread per environment)
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_integrate(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * BLOCK + threadIdx.x;
    const int nv = h.nv;
    if ((a.flags & B2F_FUSABLE) && (a.status[env] & 8)) continue;
    SArr<T> qacc{a.qacc + env, S};
    bool bad = false;
    for (int i = 0; i < nv; i++) bad |= !(t_abs(qacc[i]) < T(1e10));
    if (bad) {  // reset instead of integrating garbage
      for (int i = 0; i < h.nq; i++) a.qpos[i * S + env] = m.f(h.o_qpos0, i);
      for (int i = 0; i < nv; i++) { a.qvel[i * S + env] = 0; qacc[i] = 0; a.qacc_warmstart[i * S + env] = 0; a.qfrc_applied[i * S + env] = 0; }
      a.time[env] = 0;
      a.status[env] |= 4;
      continue;
    }
    if (a.flags & B2F_INTEGRATE) {
      SArr<T> qpos{a.qpos + env, S}, qvel{a.qvel + env, S}, qM{a.qM + env, S}, LD{a.qLD + env, S}, dinv{a.qLDiagInv + env, S};
      SArr<T> frc{a.qfrc_smooth + env, S}, xa{a.qacc_smooth + env, S};
      // qfrc_smooth becomes the total force of the implicit-damping solve; qLD / qLDiagInv / qacc_smooth are dead after
      // the solver and serve as scratch for the damped factorisation
      if (h.has_damping) for (int i = 0; i < nv; i++) frc[i] += a.qfrc_constraint[i * S + env];
      euler_step<GenericP>(m, qpos, qvel, qM, qacc, frc, a.h, LD, dinv, xa);
      a.time[env] += a.h;
      if (a.flags & B2F_ODOM) odom_override(m, a, env);
    }
  }
}

// on-demand expansions for the legacy dense fields: efc_J [njmax][nv] and efc_AR [njmax][njmax], one thread per env
template <typename T>
__global__ void k_expand_rows(const KArgs<T> a, int which /* 0: J, 1: B */, T* dst /* [njmax * nv][nenvp] */) {
  const DModel* h = reinterpret_cast<const DModel*>(a.model);
  const uint32_t* w = a.model;
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= a.nenvp) return;
  const long long S = a.nenvp;
  const int ne = a.nefc[env], nv = h->nv, WP = a.wp;
  for (int r = 0; r < h->njmax; r++) {
    for (int i = 0; i < nv; i++) dst[((long long)r * nv + i) * S + env] = 0;
    if (r >= ne) continue;
    const int t1 = a.efc_tree[((long long)2 * r) * S + env], t2 = a.efc_tree[((long long)2 * r + 1) * S + env];
    Seg g{0, 0, 0, 0};
    if (t1 >= 0) { g.s1 = (int)w[h->o_tree_dofadr + t1]; g.n1 = (int)w[h->o_tree_dofnum + t1]; }
    if (t2 >= 0) { g.s2 = (int)w[h->o_tree_dofadr + t2]; g.n2 = (int)w[h->o_tree_dofnum + t2]; }
    const T* row = a.efc_rows + ((long long)env * h->njmax + r) * (2 * WP) + (which ? WP : 0);
    for (int k = 0; k < g.n1 + g.n2; k++) dst[((long long)r * nv + seg_dof(g, k)) * S + env] = row[k];
  }
}
template <typename T>
__global__ void k_dense_AR(const KArgs<T> a, const T* Jd, const T* Bd, T* AR /* [njmax * njmax][nenvp] */) {
  const DModel* h = reinterpret_cast<const DModel*>(a.model);
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= a.nenvp) return;
  const long long S = a.nenvp;
  const int ne = a.nefc[env], nv = h->nv, ld = h->njmax;
  for (int r = 0; r < ne; r++)
    for (int c = 0; c < ne; c++) {
      T v = (r == c) ? a.efc_R[(long long)r * S + env] : T(0);
      for (int i = 0; i < nv; i++) v += Jd[((long long)r * nv + i) * S + env] * Bd[((long long)c * nv + i) * S + env];
      AR[((long long)r * ld + c) * S + env] = v;
    }
}

}  // namespace b2
