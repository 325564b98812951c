// k_constraint.cuh — constraint rows (equality, friction loss, limits, pyramidal contacts), impedance / reference
// acceleration, projection AR = J M^-1 J^T + R, the PGS solve and the final integrate kernel
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

template <typename T>
struct Rows {
  MV<T> m;
  const DModel& h;
  const KArgs<T>& a;
  int env;
  long long S;
  int nefc = 0;
  __device__ Rows(const MV<T>& mv, const KArgs<T>& args, int e) : m(mv), h(*mv.h), a(args), env(e), S(args.nenvp) {}
  __device__ __forceinline__ T& J(int r, int i) const { return a.efc_J[((long long)r * h.nv + i) * S + env]; }
  __device__ __forceinline__ T cdof(int i, int k) const { return a.cdof[(6 * i + k) * S + env]; }

  // start a new zero row; returns its index or -1 when njmax is exhausted
  __device__ int open(int type, int id, T pos, T margin, T frictionloss) {
    if (nefc >= h.njmax) { a.status[env] |= 2; return -1; }
    const int r = nefc++;
    for (int i = 0; i < h.nv; i++) J(r, i) = 0;
    a.efc_type[(long long)r * S + env] = type;
    a.efc_id[(long long)r * S + env] = id;
    a.efc_pos[(long long)r * S + env] = pos;
    a.efc_margin[(long long)r * S + env] = margin;
    a.efc_frictionloss[(long long)r * S + env] = frictionloss;
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
  __device__ void add_jacp(int r0, int b, const T* point, T sign) {
    T off[3];
    point_off(b, point, off);
    for (int i = m.i(h.o_body_lastdof, b); i >= 0; i = m.i(h.o_dof_parentid, i)) {
      T jp[3], jr[3];
      jac_col(i, off, jp, jr);
      for (int k = 0; k < 3; k++) J(r0 + k, i) += sign * jp[k];
    }
  }

  __device__ void equality() {
    if (h.disableflags & DSBL_EQUALITY) return;
    for (int q = 0; q < h.neq; q++) {
      if (!m.i(h.o_eq_active, q)) continue;
      const int type = m.i(h.o_eq_type, q), o1 = m.i(h.o_eq_obj1id, q), o2 = m.i(h.o_eq_obj2id, q);
      const T* data = m.fp(h.o_eq_data) + 11 * q;
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
        const int r = open(CN_EQUALITY, q, pos - ref, 0, 0);
        if (r < 0) continue;
        if (da2 >= 0) J(r, da2) = -deriv;
        J(r, da1) = 1;
      } else {
        const bool weld = type == EQ_WELD;
        T a1[3], a2[3], p1[3], p2[3], m1[9], m2[9];
        for (int k = 0; k < 3; k++) { a1[k] = weld ? data[3 + k] : data[k]; a2[k] = weld ? data[k] : data[3 + k]; }
        for (int k = 0; k < 9; k++) { m1[k] = a.xmat[(9 * o1 + k) * S + env]; m2[k] = a.xmat[(9 * o2 + k) * S + env]; }
        mat_vec3(p1, m1, a1);
        mat_vec3(p2, m2, a2);
        for (int k = 0; k < 3; k++) { p1[k] += a.xpos[(3 * o1 + k) * S + env]; p2[k] += a.xpos[(3 * o2 + k) * S + env]; }
        int r0 = -1;
        for (int k = 0; k < 3; k++) { const int r = open(CN_EQUALITY, q, p1[k] - p2[k], 0, 0); if (k == 0) r0 = r; }
        if (r0 < 0 || nefc - r0 < 3) continue;
        add_jacp(r0, o1, p1, T(1));
        add_jacp(r0, o2, p2, T(-1));
        if (weld) {
          const T ts = data[10];
          T q1[4], q2n[4], q1r[4], qe[4];
          for (int k = 0; k < 4; k++) { q1[k] = a.xquat[(4 * o1 + k) * S + env]; q2n[k] = a.xquat[(4 * o2 + k) * S + env]; }
          q2n[1] = -q2n[1]; q2n[2] = -q2n[2]; q2n[3] = -q2n[3];
          mul_quat(q1r, q1, data + 6);
          mul_quat(qe, q2n, q1r);
          int rr = -1;
          for (int k = 0; k < 3; k++) { const int r = open(CN_EQUALITY, q, ts * qe[1 + k], 0, 0); if (k == 0) rr = r; }
          if (rr < 0 || nefc - rr < 3) continue;
          for (int side = 0; side < 2; side++) {
            const int b = side ? o2 : o1;
            const T sg = side ? T(-1) : T(1);
            for (int i = m.i(h.o_body_lastdof, b); i >= 0; i = m.i(h.o_dof_parentid, i)) {
              const T w[4] = {0, cdof(i, 0), cdof(i, 1), cdof(i, 2)};
              T t1[4], t2[4];
              mul_quat(t1, q2n, w);
              mul_quat(t2, t1, q1r);
              for (int k = 0; k < 3; k++) J(rr + k, i) += sg * T(0.5) * ts * t2[1 + k];
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
      const int r = open(CN_FRICTION_DOF, i, 0, 0, fl);
      if (r >= 0) J(r, i) = 1;
    }
  }

  __device__ void limits() {
    if (!h.has_limits || (h.disableflags & DSBL_LIMIT)) return;
    for (int j = 0; j < h.njnt; j++) {
      if (!m.i(h.o_jnt_limited, j)) continue;
      const int qa = m.i(h.o_jnt_qposadr, j), da = m.i(h.o_jnt_dofadr, j), jt = m.i(h.o_jnt_type, j);
      const T margin = m.f(h.o_jnt_margin, j);
      if (jt == JNT_SLIDE || jt == JNT_HINGE) {
        const T value = a.qpos[qa * S + env];
        for (int side = -1; side <= 1; side += 2) {
          const T dist = side * (m.f(h.o_jnt_range, 2 * j + (side + 1) / 2) - value);
          if (dist < margin) {
            const int r = open(CN_LIMIT_JOINT, j, dist, margin, 0);
            if (r >= 0) J(r, da) = T(-side);
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
          const int r = open(CN_LIMIT_JOINT, j, dist, margin, 0);
          if (r >= 0) for (int k = 0; k < 3; k++) J(r, da + k) = -ax[k];
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
      for (int k = 0; k < nrow; k++) open(dim == 1 ? CN_CONTACT_FRICTIONLESS : CN_CONTACT_PYRAMIDAL, c, dist, im, 0);
      if (nefc - first < nrow) { nefc = first; I(CI_EFC) = -1; continue; }  // did not fit: drop the whole contact
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
          if (dim == 1) { J(first, i) += sg * fp[0]; continue; }
          for (int k = 1; k < dim; k++) {
            const T dir = k < 3 ? fp[k] : fr[k - 3];
            const T mu = fri[k - 1];
            J(first + 2 * k - 2, i) += sg * (fp[0] + mu * dir);
            J(first + 2 * k - 1, i) += sg * (fp[0] - mu * dir);
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
    for (int r = 0; r < nefc; r++) {
      const long long o = (long long)r * S + env;
      a.efc_D[o] = 1 / a.efc_R[o];
      T vel = 0;
      for (int i = 0; i < h.nv; i++) vel += J(r, i) * a.qvel[i * S + env];
      a.efc_vel[o] = vel;
      const T K = a.efc_KBI[((long long)0 * h.njmax + r) * S + env], B = a.efc_KBI[((long long)1 * h.njmax + r) * S + env];
      const T imp = a.efc_KBI[((long long)2 * h.njmax + r) * S + env];
      a.efc_aref[o] = -B * vel - K * imp * (a.efc_pos[o] - a.efc_margin[o]);
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

// K4: rows + impedance + aref, b = J qacc_smooth - aref, and (for mj_inverse) qfrc_inverse -= J^T f(qacc_prev)
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_make_constraint(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * BLOCK + threadIdx.x;
    Rows<T> rows(m, a, env);
    if (!(h.disableflags & DSBL_CONSTRAINT)) {
      rows.equality();
      rows.friction_loss();
      rows.limits();
      rows.contacts();
      rows.finish();
    }
    const int ne = rows.nefc, nv = h.nv;
    a.nefc[env] = ne;
    for (int r = 0; r < ne; r++) {
      const long long o = (long long)r * S + env;
      T js = 0, jq = 0;
      for (int i = 0; i < nv; i++) {
        const T j = rows.J(r, i);
        js += j * a.qacc_smooth[i * S + env];
        jq += j * a.qacc[i * S + env];
      }
      const T aref = a.efc_aref[o];
      a.efc_b[o] = js - aref;
      if (a.flags & B2F_INVERSE) {
        const T f = primal_force(a.efc_type[o], jq - aref, a.efc_D[o], a.efc_R[o], a.efc_frictionloss[o]);
        if (f != 0) for (int i = 0; i < nv; i++) a.qfrc_inverse[i * S + env] -= rows.J(r, i) * f;
      }
    }
  }
}

// K5: rows of M^-1 J^T by sparse back-substitution, then AR = J (M^-1 J^T) + diag(R)   (FFMA version)
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_project(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * BLOCK + threadIdx.x;
    const int ne = a.nefc[env], nv = h.nv, ld = h.njmax;
    SArr<T> LD{a.qLD + env, S}, dinv{a.qLDiagInv + env, S};
    for (int r = 0; r < ne; r++) {
      SArr<T> x{a.efc_MiJT + (long long)r * nv * S + env, S};
      for (int i = 0; i < nv; i++) x[i] = a.efc_J[((long long)r * nv + i) * S + env];
      ld_solve(m, LD, dinv, x);
    }
    for (int r = 0; r < ne; r++)
      for (int c = 0; c <= r; c++) {
        T v = 0;
        for (int i = 0; i < nv; i++) v += a.efc_J[((long long)r * nv + i) * S + env] * a.efc_MiJT[((long long)c * nv + i) * S + env];
        if (c == r) v += a.efc_R[(long long)r * S + env];
        a.efc_AR[((long long)r * ld + c) * S + env] = v;
        a.efc_AR[((long long)c * ld + r) * S + env] = v;
      }
  }
}

// K6: projected Gauss-Seidel on the dual (A.8), warm-started from qacc_warmstart
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_pgs(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * BLOCK + threadIdx.x;
    const int ne = a.nefc[env], nv = h.nv, ld = h.njmax;
    int iters = 0;
    if (ne > 0) {
      auto AR = [&](int r, int c) -> T { return a.efc_AR[((long long)r * ld + c) * S + env]; };
      auto F = [&](int r) -> T& { return a.efc_force[(long long)r * S + env]; };
      if (!(h.disableflags & DSBL_WARMSTART)) {
        for (int r = 0; r < ne; r++) {
          const long long o = (long long)r * S + env;
          T jw = 0;
          for (int i = 0; i < nv; i++) jw += a.efc_J[((long long)r * nv + i) * S + env] * a.qacc_warmstart[i * S + env];
          F(r) = primal_force(a.efc_type[o], jw - a.efc_aref[o], a.efc_D[o], a.efc_R[o], a.efc_frictionloss[o]);
        }
        T cost = 0;
        for (int r = 0; r < ne; r++) {
          T Af = 0;
          for (int c = 0; c < ne; c++) Af += AR(r, c) * F(c);
          cost += F(r) * (T(0.5) * Af + a.efc_b[(long long)r * S + env]);
        }
        if (cost > 0) for (int r = 0; r < ne; r++) F(r) = 0;
      } else {
        for (int r = 0; r < ne; r++) F(r) = 0;
      }
      const T scale = 1 / (m.f(h.o_opt_real, 4) * T(nv > 1 ? nv : 1));
      for (int it = 0; it < h.iterations; it++) {
        T improvement = 0;
        for (int r = 0; r < ne; r++) {
          const long long o = (long long)r * S + env;
          T res = a.efc_b[o];
          for (int c = 0; c < ne; c++) res += AR(r, c) * F(c);
          const T Arr = AR(r, r), old = F(r);
          T f = old - res / Arr;
          const int type = a.efc_type[o];
          if (type == CN_FRICTION_DOF) { const T fl = a.efc_frictionloss[o]; f = t_min(fl, t_max(-fl, f)); }
          else if (type != CN_EQUALITY) f = t_max(T(0), f);
          const T delta = f - old;
          const T change = T(0.5) * delta * delta * Arr + delta * res;
          if (change > T(1e-10)) continue;
          F(r) = f;
          improvement -= change;
        }
        iters = it + 1;
        if (improvement * scale < m.f(h.o_opt_real, 3)) break;
      }
    }
    a.solver_iter[env] = iters;
  }
}

// G7: qfrc_constraint = J^T f, qacc = qacc_smooth + M^-1 qfrc_constraint, warm start, Euler, odom override
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_integrate(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * BLOCK + threadIdx.x;
    const int ne = a.nefc[env], nv = h.nv;
    SArr<T> qfc{a.qfrc_constraint + env, S}, qacc{a.qacc + env, S}, LD{a.qLD + env, S}, dinv{a.qLDiagInv + env, S};
    for (int i = 0; i < nv; i++) qfc[i] = 0;
    for (int r = 0; r < ne; r++) {
      const T f = a.efc_force[(long long)r * S + env];
      if (f == 0) continue;
      for (int i = 0; i < nv; i++) qfc[i] += a.efc_J[((long long)r * nv + i) * S + env] * f;
    }
    if (ne > 0) {
      for (int i = 0; i < nv; i++) qacc[i] = qfc[i];
      ld_solve(m, LD, dinv, qacc);
      for (int i = 0; i < nv; i++) qacc[i] += a.qacc_smooth[i * S + env];
    } else {
      for (int i = 0; i < nv; i++) qacc[i] = a.qacc_smooth[i * S + env];
    }
    bool bad = false;
    for (int i = 0; i < nv; i++) {
      const T v = qacc[i];
      bad |= !(t_abs(v) < T(1e10));
      a.qacc_warmstart[i * S + env] = v;
    }
    if (bad) {  // mj_checkAcc: reset instead of integrating garbage
      for (int i = 0; i < h.nq; i++) a.qpos[i * S + env] = m.f(h.o_qpos0, i);
      for (int i = 0; i < nv; i++) { a.qvel[i * S + env] = 0; qacc[i] = 0; a.qacc_warmstart[i * S + env] = 0; a.qfrc_applied[i * S + env] = 0; }
      a.time[env] = 0;
      a.status[env] |= 4;
      continue;
    }
    if (a.flags & B2F_INTEGRATE) {
      SArr<T> qpos{a.qpos + env, S}, qvel{a.qvel + env, S}, qM{a.qM + env, S}, frc{a.qfrc_smooth + env, S};
      // total force for the implicit-damping solve; qLD / qLDiagInv / efc_MiJT row 0 are free to be reused as scratch
      for (int i = 0; i < nv; i++) frc[i] += qfc[i];
      SArr<T> xa{a.qfrc_passive + env, S};
      euler_step(m, qpos, qvel, qM, qacc, frc, a.h, LD, dinv, xa);
      a.time[env] += a.h;
      if (a.flags & B2F_ODOM) odom_override(m, a, env);
    }
  }
}

}  // namespace b2
