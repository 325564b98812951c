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
#include <type_traits>
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
  T* rowsh = nullptr;   // shared-memory column of this thread for the base rows of the contact being assembled (or null)
  int rowld = 0;        // its stride (the CTA width)
  int* rowtab = nullptr; // shared-memory table of the environment, one word per row: type | nb << 4 | (t1 + 1) << 8 | (t2 + 1) << 20
                         // (what the block table and the island ordering need of a block's first row; or null)
  __device__ Rows(const MV<T>& mv, const KArgs<T>& args, int e) : m(mv), h(*mv.h), a(args), env(e), S(args.nenvp) {}
  // element (r, k) of the compact Jacobian; with the tensor-core projection on, every write also goes to the
  // environment-major copy the GEMM reads
  struct JRef {
    T* p; T* q;
    __device__ __forceinline__ void operator=(T v) const { *p = v; if (q) *q = v; }
    __device__ __forceinline__ void operator+=(T v) const { const T nv = *p + v; *p = nv; if (q) *q = nv; }
  };
  __device__ __forceinline__ JRef Jc(int r, int k) const {
    return JRef{a.efc_J + ((long long)r * h.wmax + k) * S + env,
                a.em_rows ? a.efc_Jem + ((long long)env * a.em_rows + r) * 64 + k : nullptr};
  }
  __device__ __forceinline__ T cdof(int i, int k) const { return a.cdof[(6 * i + k) * S + env]; }

  // start a new zero row over trees (t1, t2); returns its index or -1 when njmax is exhausted
  __device__ int open(int type, int id, T pos, T margin, T frictionloss, int t1, int t2, Seg* gout, bool zero = true) {
    if (nefc >= h.njmax) { a.status[env] |= 2; return -1; }
    return open_at(nefc++, type, id, pos, margin, frictionloss, t1, t2, gout, zero);
  }
  // the same for a row whose index is already known (contacts assembled side by side by the lanes of a team)
  __device__ int open_at(int r, int type, int id, T pos, T margin, T frictionloss, int t1, int t2, Seg* gout, bool zero = true, int nb = 1) {
    if (t1 < 0) { t1 = t2; t2 = -1; }
    if (t1 == t2) t2 = -1;
    if (t2 >= 0 && t2 < t1) { const int t = t1; t1 = t2; t2 = t; }
    const Seg g = seg_of(m, t1, t2);
    if (rowtab) rowtab[r] = type | (nb << 4) | ((t1 + 1) << 8) | ((t2 + 1) << 20);
    if (zero) for (int k = 0; k < g.n1 + g.n2; k++) Jc(r, k) = 0;
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
    for (int j0 = 0; j0 < h.njnt; j0 += 8) {
    // (the positions of eight joints are fetched before the first row is opened: its stores order later loads behind them)
    T val8[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int j = j0 + k;
      const bool scalar = j < h.njnt && m.i(h.o_jnt_limited, j) && (m.i(h.o_jnt_type, j) == JNT_SLIDE || m.i(h.o_jnt_type, j) == JNT_HINGE);
      val8[k] = scalar ? a.qpos[m.i(h.o_jnt_qposadr, j) * S + env] : T(0);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int j = j0 + k;
      if (j >= h.njnt) break;
      if (!m.i(h.o_jnt_limited, j)) continue;
      const int qa = m.i(h.o_jnt_qposadr, j), da = m.i(h.o_jnt_dofadr, j), jt = m.i(h.o_jnt_type, j);
      const T margin = m.f(h.o_jnt_margin, j);
      Seg g;
      if (jt == JNT_SLIDE || jt == JNT_HINGE) {
        const T value = val8[k];
        for (int side = -1; side <= 1; side += 2) {
          const T dist = side * (m.f(h.o_jnt_range, 2 * j + (side + 1) / 2) - value);
          if (dist < margin) {
            const int r = open(CN_LIMIT_JOINT, j, dist, margin, 0, m.i(h.o_dof_treeid, da), -1, &g);
            if (r >= 0) Jc(r, seg_pos(g, da)) = T(-side);
          }
        }
      } else if (jt == JNT_BALL) {
        T q[4];
        for (int k4 = 0; k4 < 4; k4++) q[k4] = a.qpos[(qa + k4) * S + env];
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
          if (r >= 0) for (int k2 = 0; k2 < 3; k2++) Jc(r, seg_pos(g, da + k2)) = -ax[k2];
        }
      }
    }
    }
  }

  // first row of every contact, in contact order (sequential: a contact that does not fit is dropped whole, later and
  // smaller ones may still fit).  Returns the row count after the contacts.
  __device__ int contact_offsets(int ne0) {
    if (h.disableflags & DSBL_CONTACT) return ne0;
    const int ncon = a.ncon[env];
    int r = ne0;
    // (the dims of eight contacts are fetched before the first offset is stored: a store into contact_int orders every
    //  later load of the same array behind it, one L2 round trip per contact)
    for (int c0 = 0; c0 < ncon; c0 += 8) {
      int dims[8];
#pragma unroll
      for (int k = 0; k < 8; k++) dims[k] = c0 + k < ncon ? a.coni[((long long)CI_DIM * h.nconmax + c0 + k) * S + env] : 1;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int c = c0 + k;
        if (c >= ncon) break;
        const int nrow = dims[k] == 1 ? 1 : 2 * (dims[k] - 1);
        int& efc = a.coni[((long long)CI_EFC * h.nconmax + c) * S + env];
        if (r + nrow > h.njmax) { a.status[env] |= 2; efc = -1; }
        else { efc = r; r += nrow; }
      }
    }
    return r;
  }

  // rows of contact c (first row from contact_offsets): headers, the base directions' Jacobians, impedance, cone R
  __device__ void contact_one(int c) {
      auto F = [&](int f) -> T { return a.con[((long long)f * h.nconmax + c) * S + env]; };
      auto I = [&](int f) -> int& { return a.coni[((long long)f * h.nconmax + c) * S + env]; };
      const int first = I(CI_EFC);
      if (first < 0) return;
      const int dim = I(CI_DIM);
      const int b1 = m.i(h.o_geom_bodyid, I(CI_GEOM1)), b2 = m.i(h.o_geom_bodyid, I(CI_GEOM2));
      T pos[3], frame[9];
      for (int k = 0; k < 3; k++) pos[k] = F(CF_POS + k);
      for (int k = 0; k < 9; k++) frame[k] = F(CF_FRAME + k);
      const T dist = F(CF_DIST), im = F(CF_INCLUDEMARGIN);
      const int nrow = dim == 1 ? 1 : 2 * (dim - 1);
      Seg g;
      const int t1 = m.i(h.o_body_treeid, b1), t2 = m.i(h.o_body_treeid, b2);
      // With a shared-memory column the dim base rows are accumulated there and stored once: in HBM every "+=" of a
      // Jacobian entry is a load behind a store (an L2 round trip), ~100 per contact.  Rows dim .. nrow - 1 of a pyramidal
      // contact are index space only (never read: k_solve_rows / k_make_blocks work on the base rows) and are left as is.
      const bool sh = rowsh != nullptr && dim <= a.row_nb;
      for (int k = 0; k < nrow; k++) open_at(first + k, dim == 1 ? CN_CONTACT_FRICTIONLESS : CN_CONTACT_PYRAMIDAL, c, dist, im, 0, t1, t2, &g, !sh, dim);
      const int wrow = g.n1 + g.n2;
      auto RS = [&](int k, int e) -> T& { return rowsh[((long long)k * h.wmax + e) * rowld]; };
      if (sh) {
        for (int k = 0; k < dim; k++)
          for (int e = 0; e < wrow; e++) RS(k, e) = 0;
      }
      for (int side = 0; side < 2; side++) {
        const int b = side ? b2 : b1;
        const T sg = side ? T(1) : T(-1);
        T off[3];
        point_off(b, pos, off);
        // (the spatial axes of six dofs of the chain are fetched before the first one is used: one dof per L2 round trip
        //  otherwise, twelve per arm - prop contact)
        for (int i0 = m.i(h.o_body_lastdof, b); i0 >= 0;) {
          int id6[6];
          T cd6[6][6];
          int nd = 0, inext = i0;
#pragma unroll
          for (int q = 0; q < 6; q++) {
            id6[q] = inext;
            if (inext >= 0) { nd = q + 1; for (int k = 0; k < 6; k++) cd6[q][k] = cdof(inext, k); inext = m.i(h.o_dof_parentid, inext); }
          }
          i0 = inext;
#pragma unroll
          for (int q = 0; q < 6; q++) {
          if (q >= nd) break;
          const int i = id6[q];
          T jp[3], jr[3], fp[3], fr[3], t[3];
          cross3(t, cd6[q], off);
          jp[0] = cd6[q][3] + t[0]; jp[1] = cd6[q][4] + t[1]; jp[2] = cd6[q][5] + t[2];
          jr[0] = cd6[q][0]; jr[1] = cd6[q][1]; jr[2] = cd6[q][2];
          mat_vec3(fp, frame, jp);  // rows of frame: normal, tangent1, tangent2
          mat_vec3(fr, frame, jr);
          const int kk = seg_pos(g, i);
          // rows first .. first + dim - 1 hold the BASE directions of the contact: normal, tangent 1, tangent 2, torsion,
          // rolling 1, rolling 2 (unscaled); pyramid row 2 (k - 1) + s is base 0 +/- friction[k - 1] * base k and is never
          // materialised on the hot path (the legacy efc_J view is expanded on demand)
          if (sh) {
            RS(0, kk) += sg * fp[0];
            for (int k = 1; k < dim; k++) RS(k, kk) += sg * (k < 3 ? fp[k] : fr[k - 3]);
          } else {
            Jc(first, kk) += sg * fp[0];
            for (int k = 1; k < dim; k++) Jc(first + k, kk) += sg * (k < 3 ? fp[k] : fr[k - 3]);
          }
          }
        }
      }
      if (sh)
        for (int k = 0; k < dim; k++)
          for (int e = 0; e < wrow; e++) Jc(first + k, e) = RS(k, e);
      // diagApprox, impedance, R, K, B of the contact's rows (finish_range for one contact): its rows share type, id,
      // pos and margin, all at hand here, so nothing is read back from the row arrays just written (each such load is an
      // L2 round trip behind a store), and the impedance and the reference spring are evaluated once
      {
        T solref[2], solimp[5], fric[5];
        for (int k = 0; k < 2; k++) solref[k] = F(CF_SOLREF + k);
        for (int k = 0; k < 5; k++) solimp[k] = F(CF_SOLIMP + k);
        for (int k = 0; k < 5; k++) fric[k] = F(CF_FRICTION + k);
        const T tran = m.f(h.o_body_invweight0, 2 * b1) + m.f(h.o_body_invweight0, 2 * b2);
        const T rot = m.f(h.o_body_invweight0, 2 * b1 + 1) + m.f(h.o_body_invweight0, 2 * b2 + 1);
        const T imp = impedance(solimp, dist, im);
        T K, B;
        spring(solref, solimp, K, B);
        T R0 = 0;
        for (int j = 0; j < nrow; j++) {
          const int r = first + j;
          T diag = tran;
          if (dim > 1) { const T fr = fric[j / 2]; diag = tran + fr * fr * (j < 4 ? tran : rot); }
          const T R = t_max(Eps<T>::minval(), (1 - imp) * diag / imp);
          if (j == 0) R0 = R;
          a.efc_diagApprox[(long long)r * S + env] = diag;
          a.efc_KBI[((long long)0 * h.njmax + r) * S + env] = K;
          a.efc_KBI[((long long)1 * h.njmax + r) * S + env] = B;
          a.efc_KBI[((long long)2 * h.njmax + r) * S + env] = imp;
          if (dim == 1) a.efc_R[(long long)r * S + env] = R;
        }
        if (dim > 1) {   // pyramidal cones share R = 2 mu^2 R_first
          const T mu = fric[0] / t_sqrt(t_max(Eps<T>::minval(), m.f(h.o_opt_real, 5)));
          const T Rpy = t_max(Eps<T>::minval(), 2 * mu * mu * R0);
          for (int j = 0; j < nrow; j++) a.efc_R[(long long)(first + j) * S + env] = Rpy;
        }
      }
  }

  // reference spring of a row: stiffness K and damping B from solref / solimp (mj_makeImpedance)
  __device__ __forceinline__ void spring(const T* solref, const T* solimp, T& K, T& B) const {
    const T hs = a.dt();
    const T dmax = t_min(T(0.9999), t_max(T(0.0001), solimp[1]));
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
  }

  // diagApprox, impedance -> R, D; K, B, imp; aref; (vel uses the current, possibly overridden, qvel)
  __device__ void finish_range(int r0, int r1, int step) {
    // the rows of one contact share its parameters: fetched once per contact (13 HBM loads), not once per row
    int c_id = -1, c_first = 0;
    T c_solref[2] = {0, 0}, c_solimp[5] = {0, 0, 0, 0, 0}, c_tran = 0, c_rot = 0;
    for (int r = r0; r < r1; r += step) {
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
        if (id != c_id) {
          c_id = id;
          for (int k = 0; k < 2; k++) c_solref[k] = F(CF_SOLREF + k);
          for (int k = 0; k < 5; k++) c_solimp[k] = F(CF_SOLIMP + k);
          const int b1 = m.i(h.o_geom_bodyid, I(CI_GEOM1)), b2 = m.i(h.o_geom_bodyid, I(CI_GEOM2));
          c_tran = m.f(h.o_body_invweight0, 2 * b1) + m.f(h.o_body_invweight0, 2 * b2);
          c_rot = m.f(h.o_body_invweight0, 2 * b1 + 1) + m.f(h.o_body_invweight0, 2 * b2 + 1);
          c_first = I(CI_EFC);
        }
        for (int k = 0; k < 2; k++) solref[k] = c_solref[k];
        for (int k = 0; k < 5; k++) solimp[k] = c_solimp[k];
        const T tran = c_tran, rot = c_rot;
        if (type == CN_CONTACT_FRICTIONLESS) diag = tran;
        else {
          const int j = r - c_first;
          const T fr = F(CF_FRICTION + j / 2);
          diag = tran + fr * fr * (j < 4 ? tran : rot);
        }
      }
      const T pos = a.efc_pos[(long long)r * S + env], margin = a.efc_margin[(long long)r * S + env];
      const T imp = impedance(solimp, pos, margin);
      T R = t_max(Eps<T>::minval(), (1 - imp) * diag / imp);
      a.efc_diagApprox[(long long)r * S + env] = diag;
      a.efc_R[(long long)r * S + env] = R;
      T K, B;
      spring(solref, solimp, K, B);
      if (type == CN_FRICTION_DOF) K = 0;
      a.efc_KBI[((long long)0 * h.njmax + r) * S + env] = K;
      a.efc_KBI[((long long)1 * h.njmax + r) * S + env] = B;
      a.efc_KBI[((long long)2 * h.njmax + r) * S + env] = imp;
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
  const int ntiles = a.ncount / BLOCK;                                               \
  (void)S; (void)h;

template <typename T, int N> struct alignas(sizeof(T) * N) VecN { T v[N]; };

// ---- block records: what the solver streams -----------------------------------------------------------------------
// A BLOCK is one scalar row (equality, friction loss, limit, frictionless contact: nb = 1) or one pyramidal contact
// (nb = condim base directions, nrow = 2 (nb - 1) solver rows J_0 +/- mu_k J_k).  Environment e owns the word range
// efc_blocks[e * capw, e * capw + efc_nwords[e]) (environment-major: a team streams its own environment), block after
// block in row order:
//   [0] code = type + 16 nb + 256 nrow   [1] s1   [2] n1 + 1024 w   [3] s2   [4] R   [5] frictionloss   [6] len   [7] row0
//   aref[nrow]  Arr[nrow]  1/Arr[nrow]  b[nrow]            (Arr = J_r M^-1 J_r^T + R, b = J_r qacc_smooth - aref_r)
//   mu[nb - 1]  ARu[nrow (nrow - 1) / 2]                   (nb > 1: J_r M^-1 J_s^T for the pyramid rows r < s of the contact,
//                                                           packed by rows: what relaxing row r does to the later rows)
//   (pad to a multiple of 4)  J[nb][wq]  B[nb][wq]         (compact over the block's trees, wq = w rounded up to 4)
// Keeping the base directions instead of the pyramid rows makes a contact 2-3x smaller than its rows, needs nb instead
// of 2 (nb - 1) products with M^-1, and lets the solver update a whole contact from nb dot products (k_pgs_block).
enum { BH_CODE = 0, BH_S1, BH_N1W, BH_S2, BH_R, BH_FL, BH_LEN, BH_ROW0, BH_N };
struct BlockShape {
  int type, nb, nrow, s1, n1, s2, w, wq, oAref, oArr, oiA, ob, oMu, oA, oJ, oB, len;
  __host__ __device__ void layout() {
    wq = (w + 3) & ~3;
    oAref = BH_N; oArr = oAref + nrow; oiA = oArr + nrow; ob = oiA + nrow; oMu = ob + nrow;
    oA = oMu + (nb > 1 ? nb - 1 : 0);
    oJ = (oA + (nb > 1 ? nrow * (nrow - 1) / 2 : 0) + 3) & ~3;
    oB = oJ + nb * wq;
    // records are whole 128-byte lines: the lanes of k_pgs_island keep the bank alignment they were staged with while they
    // walk from record to record
    len = (oB + nb * wq + 31) & ~31;
  }
  __host__ __device__ int dof(int e) const { return e < n1 ? s1 + e : s2 + e - n1; }
  // position of (r, s), r < s, in the packed strict upper triangle
  __host__ __device__ int aru(int r, int s) const { return oA + r * (2 * nrow - r - 1) / 2 + (s - r - 1); }
};
// integers ride in the records as bit patterns (fp32: no conversion on the solver's critical path) / exact values (fp64)
__device__ __forceinline__ float enc_int(int v, float) { return __int_as_float(v); }
__device__ __forceinline__ double enc_int(int v, double) { return (double)v; }
__device__ __forceinline__ int dec_int(float v) { return __float_as_int(v); }
__device__ __forceinline__ int dec_int(double v) { return (int)v; }
template <typename T>
__device__ __forceinline__ BlockShape block_shape(const T* rec) {
  BlockShape s;
  const int code = dec_int(rec[BH_CODE]), n1w = dec_int(rec[BH_N1W]);
  s.type = code & 15; s.nb = (code >> 4) & 15; s.nrow = code >> 8;
  s.s1 = dec_int(rec[BH_S1]); s.n1 = n1w & 1023; s.w = n1w >> 10; s.s2 = dec_int(rec[BH_S2]);
  s.layout();
  return s;
}
// words one environment can need: every row as a single-row block is the worst case
__host__ __device__ inline int block_capacity(int njmax, int wmax) { return njmax * ((BH_N + 4 + 2 * ((wmax + 3) & ~3) + 31) & ~31); }
// words of header + row parameters (everything in front of J) of the largest block
__host__ __device__ inline int block_max_params(int nbmax) {
  BlockShape s{};
  s.nb = nbmax; s.nrow = nbmax > 1 ? 2 * (nbmax - 1) : 1; s.w = 4;
  s.layout();
  return s.oJ;
}
__host__ __device__ inline int block_max_words(int nbmax, int wmax) {
  BlockShape s{};
  s.nb = nbmax; s.nrow = nbmax > 1 ? 2 * (nbmax - 1) : 1; s.w = wmax;
  s.layout();
  return s.len;
}

// K4: rows and impedance in MuJoCo's row order, and the block table (first row, word offset in the slab) that lets
// k_make_blocks work on all blocks of all environments at once.  A team of L lanes per environment: lane 0 opens the
// scalar rows (equalities, friction loss, limits) and fixes every contact's first row — the inherently sequential part —
// then the lanes assemble the contacts side by side (headers, base-direction Jacobians, impedance, cone R: a contact is
// independent of the others once its first row is known) and share the impedance pass over the scalar rows; lane 0
// finishes with the block table and the island ordering.
template <typename T, int BLOCK, int L>
__global__ void __launch_bounds__(BLOCK) k_make_rows(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  (void)ntiles;
  constexpr int EPB = BLOCK / L;
  const int nteams = a.ncount / EPB;
  const int team = threadIdx.x / L, l = threadIdx.x % L;
  const int tshift = (threadIdx.x & 31) & ~(L - 1);
  const unsigned tmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << tshift;
  int wmaxblk = 0;
  for (int tile = blockIdx.x; tile < nteams; tile += gridDim.x) {
    const int env = tile * EPB + team;
    Rows<T> rows(m, a, env);
    if (a.row_nb > 0) { rows.rowsh = reinterpret_cast<T*>(smem_raw + 16 + (size_t)nwords * 4) + threadIdx.x; rows.rowld = BLOCK; }
    int* const rowtab = a.row_tab ? reinterpret_cast<int*>(smem_raw + 16 + (size_t)nwords * 4 + (size_t)a.row_nb * h.wmax * BLOCK * sizeof(T) +
                                                           (a.isl_cap > 0 ? (size_t)2 * h.ntree * BLOCK * sizeof(int) : 0)) + (size_t)team * h.njmax : nullptr;
    rows.rowtab = rowtab;
    const bool done = (a.flags & B2F_FUSABLE) && (a.status[env] & 8);  // already integrated by the smooth kernel
    const bool active = !(h.disableflags & DSBL_CONSTRAINT) && !done;
    int ne0 = 0, ne = 0;
    if (active && l == 0) {
      rows.equality();
      rows.friction_loss();
      rows.limits();
      ne0 = rows.nefc;
      ne = rows.contact_offsets(ne0);
    }
    __syncwarp(tmask);
    ne0 = __shfl_sync(tmask, ne0, 0, L);
    ne = __shfl_sync(tmask, ne, 0, L);
    if (active) {
      rows.finish_range(l, ne0, L);
      if (!(h.disableflags & DSBL_CONTACT)) {
        const int ncon = a.ncon[env];
        for (int c = l; c < ncon; c += L) rows.contact_one(c);
      }
    }
    __syncwarp(tmask);
    if (l != 0) continue;
    if (!done) a.nefc[env] = ne;
    int r = 0, woff = 0, nblk = 0;
    // island labels: union-find over the kinematic trees (root = smallest tree id of the component), in a per-thread
    // shared-memory column behind the row column
    const int nt = a.isl_cap > 0 ? h.ntree : 0;
    int* lab = reinterpret_cast<int*>(smem_raw + 16 + (size_t)nwords * 4 + (size_t)a.row_nb * h.wmax * BLOCK * sizeof(T)) + threadIdx.x;
    auto LAB = [&](int t) -> int& { return lab[(size_t)t * BLOCK]; };
    auto CNT = [&](int t) -> int& { return lab[(size_t)(nt + t) * BLOCK]; };
    auto find = [&](int t) { if (t < 0) return 0; while (LAB(t) != t) t = LAB(t); return t; };
    for (int t = 0; t < nt; t++) { LAB(t) = t; CNT(t) = 0; }
    // shape and trees of the block that starts at row r: from the shared-memory row table the lanes filled while they
    // opened the rows, else read back from the row arrays in HBM (an L2 round trip per block and per pass below)
    auto block_at = [&](int r, BlockShape& bs, int& t1, int& t2) {
      if (rowtab) {
        const int pk = rowtab[r];
        bs.type = pk & 15; bs.nb = (pk >> 4) & 15;
        t1 = ((pk >> 8) & 4095) - 1; t2 = ((pk >> 20) & 4095) - 1;
      } else {
        const long long o = (long long)r * S + env;
        bs.type = a.efc_type[o];
        bs.nb = bs.type == CN_CONTACT_PYRAMIDAL ? a.coni[((long long)CI_DIM * h.nconmax + a.efc_id[o]) * S + env] : 1;
        t1 = a.efc_tree[((long long)2 * r) * S + env]; t2 = a.efc_tree[((long long)2 * r + 1) * S + env];
      }
      bs.nrow = bs.type == CN_CONTACT_PYRAMIDAL ? 2 * (bs.nb - 1) : 1;
      const Seg g = seg_of(m, t1, t2);
      bs.s1 = g.s1; bs.n1 = g.n1; bs.s2 = g.s2; bs.w = g.n1 + g.n2;
      bs.layout();
    };
    while (r < ne) {
      BlockShape bs{};
      int t1, t2;
      block_at(r, bs, t1, t2);
      a.blk_row0[(long long)nblk * S + env] = r;
      if (!nt) a.blk_off[(long long)nblk * S + env] = woff;   // (islands: the offsets follow below)
      if (nt && t2 >= 0) {
        const int ra = find(t1), rb = find(t2);
        if (ra != rb) LAB(ra > rb ? ra : rb) = ra < rb ? ra : rb;
      }
      nblk++;
      woff += bs.len; r += bs.nrow;
    }
    if (nt) {
      // words per island, islands numbered by ascending root; then every block's offset inside its island (row order kept)
      for (int r2 = 0; r2 < ne;) {
        BlockShape bs{};
        int t1, t2;
        block_at(r2, bs, t1, t2);
        CNT(find(t1)) += bs.len;
        r2 += bs.nrow;
      }
      int start = 0, ni = 0;
      for (int t = 0; t < nt; t++) {
        const int c = CNT(t);
        if (c <= 0) continue;
        if (ni < a.isl_cap) { a.isl_off[(long long)ni * S + env] = start; a.isl_end[(long long)ni * S + env] = start + c; }
        else a.isl_end[(long long)(a.isl_cap - 1) * S + env] = start + c;   // more islands than slots: the last slot takes the rest (still one contiguous range)
        CNT(t) = start;
        start += c;
        ni++;
      }
      int q = 0;
      for (int r2 = 0; r2 < ne; q++) {
        BlockShape bs{};
        int t1, t2;
        block_at(r2, bs, t1, t2);
        const int root = find(t1);
        a.blk_off[(long long)q * S + env] = CNT(root);
        CNT(root) += bs.len;
        r2 += bs.nrow;
      }
      if (!done) a.nisl[env] = ni < a.isl_cap ? ni : a.isl_cap; else a.nisl[env] = 0;
    }
    if (!done) { a.efc_nwords[env] = woff; a.nblk[env] = nblk; wmaxblk = max(wmaxblk, nblk); }
    else a.nblk[env] = 0;
  }
  for (int o = 16; o > 0; o >>= 1) wmaxblk = max(wmaxblk, __shfl_xor_sync(0xffffffffu, wmaxblk, o));
  if ((threadIdx.x & 31) == 0 && wmaxblk > 0) atomicMax(a.maxblk, wmaxblk);
}

// K5a (wide trees only, e.g. a PR2: one 49-dof tree, nM = 492): B = M^-1 J^T for every row of every environment as a
// kernel of its own.  A CTA owns a tile of 32 environments (lane = environment, so every global access is a coalesced
// 128-byte run) and ROWS warps, one constraint row each per pass.  The tile's factor (qLD, qLDiagInv: what every solve
// reads ~2 nM times) is loaded into shared memory ONCE and shared by all rows; each thread keeps its right-hand side as a
// shared-memory column.  In k_make_blocks the same solve re-read the factor from HBM / L2 for every (row, environment):
// 4.0 of the 12.7 ms of a PR2 tick.  Same loops, same operation order as that solve: results are bit-identical.
template <typename T, int ROWS>
__global__ void __launch_bounds__(32 * ROWS) k_solve_rows(const KArgs<T> a) {
  if ((a.flags & B2F_FUSABLE) && a.pending[0] == 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MV<T> m{reinterpret_cast<const DModel*>(a.model), a.model};
  const DModel& h = *m.h;
  const long long S = a.nenvp;
  const int nM = h.nM, nv = h.nv, W = h.wmax;
  T* LDs = reinterpret_cast<T*>(smem_raw);          // [nM][32]
  T* dis = LDs + (size_t)nM * 32;                   // [nv][32]
  T* Xs = dis + (size_t)nv * 32;                    // [ROWS][W][32]
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int ntiles = a.ncount / 32;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * 32 + lane;
    __syncthreads();   // the previous tile's solves are done with the factor
    for (int i = wrp; i < nM; i += ROWS) LDs[i * 32 + lane] = a.qLD[(long long)i * S + env];
    for (int i = wrp; i < nv; i += ROWS) dis[i * 32 + lane] = a.qLDiagInv[(long long)i * S + env];
    __syncthreads();
    const int ne = a.nefc[env];
    const int nemax = __reduce_max_sync(0xffffffffu, ne);
    SArr<T> LD{LDs + lane, 32}, dinv{dis + lane, 32};
    SArr<T> X{Xs + (size_t)wrp * W * 32 + lane, 32};
    for (int r = wrp; r < nemax; r += ROWS) {
      if (r >= ne) continue;
      if (a.efc_type[(long long)r * S + env] == CN_CONTACT_PYRAMIDAL) {   // only the base rows of a contact carry a Jacobian
        const int c = a.efc_id[(long long)r * S + env];
        if (r - a.coni[((long long)CI_EFC * h.nconmax + c) * S + env] >= a.coni[((long long)CI_DIM * h.nconmax + c) * S + env]) continue;
      }
      const Seg g = seg_of(m, a.efc_tree[((long long)2 * r) * S + env], a.efc_tree[((long long)2 * r + 1) * S + env]);
      const int w = g.n1 + g.n2;
      for (int e = 0; e < w; e++) X[e] = a.efc_J[((long long)r * W + e) * S + env];
      for (int sgm = 0; sgm < 2; sgm++) {
        const int lo = sgm ? g.s2 : g.s1, n = sgm ? g.n2 : g.n1, base = sgm ? g.n1 : 0;
        if (n == 0) continue;
        for (int i = lo + n - 1; i >= lo; i--) {
          const T xi = X[base + i - lo];
          if (xi == 0) continue;
          const int adr = m.i(h.o_dof_Madr, i), cnt = m.i(h.o_dof_Mcnt, i);
          for (int q = 1; q < cnt; q++) X[base + m.i(h.o_dof_anc, adr + q) - lo] -= LD[adr + q] * xi;
        }
        for (int i = lo; i < lo + n; i++) X[base + i - lo] *= dinv[i];
        for (int i = lo; i < lo + n; i++) {
          const int adr = m.i(h.o_dof_Madr, i), cnt = m.i(h.o_dof_Mcnt, i);
          T xi = X[base + i - lo];
          for (int q = 1; q < cnt; q++) xi -= LD[adr + q] * X[base + m.i(h.o_dof_anc, adr + q) - lo];
          X[base + i - lo] = xi;
        }
      }
      for (int e = 0; e < w; e++) a.efc_B[((long long)r * W + e) * S + env] = X[e];
    }
  }
}

template <typename T, int BLOCK> __device__ __forceinline__ void make_blocks_pass(const KArgs<T>& a, const int blk);
// K5: the expensive, embarrassingly parallel part — one thread per (block, environment): vel, aref, b of the block's solver
// rows, the primal force of mj_inverse per row, B = M^-1 J^T of the block's base directions by sparse back-substitution
// inside its trees (in shared memory), the local matrix A, the couplings between the pyramid rows, diag(AR) per row, and
// the finished record, which leaves through a shared-memory transpose so that
// every global store of the slab is a coalesced run.  grid.y = njmax (the most blocks an environment can have): CTAs
// beyond the largest block count of this tick (maxblk, found by k_make_rows) exit at once.
template <typename T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_make_blocks(const KArgs<T> a) {
  if ((a.flags & B2F_FUSABLE) && a.pending[0] == 0) return;
  const int maxblk = a.maxblk[0];
  for (int blk = blockIdx.y; blk < maxblk; blk += gridDim.y) make_blocks_pass<T, BLOCK>(a, blk);
}
template <typename T, int BLOCK>
__device__ __forceinline__ void make_blocks_pass(const KArgs<T>& a, const int blk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // The kernel needs a few integer tables of the model (dof_parentid, dof_Madr, the tree table), warp-uniform or nearly
  // so: they are read from the blob in HBM through L1 instead of staging a private copy per CTA.  For PR2-sized models
  // the copy (29 KB) was more than the CTA's own columns (20 KB) and capped the SM at 4 resident warps.
  const int nwords = 0;
  MV<T> m{reinterpret_cast<const DModel*>(a.model), a.model};
  const DModel& h = *m.h;
  const long long S = a.nenvp;
  const int ntiles = a.ncount / BLOCK;
  constexpr int LDS = BLOCK + 1;  // +1: conflict-free both for per-thread columns and for the transposed reads
  // per thread (column): [0, npar) header + row parameters, then one base direction at a time: J_c [wqmax] | B_c [wqmax].
  // Streaming the base directions keeps the footprint at npar + 2 wq words per thread instead of the whole record
  // (PR2-sized trees: 156 instead of 468 words -> 4x the resident warps).
  T* colsh = reinterpret_cast<T*>(smem_raw + (size_t)nwords * 4);
  const int npar = a.block_npar, wqmax = (h.wmax + 3) & ~3;
  const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
  const int capw = a.block_capw;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * BLOCK + threadIdx.x;
    const int W = h.wmax;
    SArr<T> LD{a.qLD + env, S}, dinv{a.qLDiagInv + env, S};
    SArr<T> rec{colsh + threadIdx.x, LDS};
    SArr<T> Jc{colsh + (size_t)npar * LDS + threadIdx.x, LDS}, Bc{colsh + (size_t)(npar + wqmax) * LDS + threadIdx.x, LDS};
    const bool have = blk < a.nblk[env];
    BlockShape bs{};
    int woff = 0, r = 0;
    Seg g{0, 0, 0, 0};
    T fri[5] = {0, 0, 0, 0, 0};
    if (have) {
      r = a.blk_row0[(long long)blk * S + env];
      woff = a.blk_off[(long long)blk * S + env];
      const long long o = (long long)r * S + env;
      bs.type = a.efc_type[o];
      const int id = a.efc_id[o];
      bs.nb = 1; bs.nrow = 1;
      if (bs.type == CN_CONTACT_PYRAMIDAL) {
        bs.nb = a.coni[((long long)CI_DIM * h.nconmax + id) * S + env];
        bs.nrow = 2 * (bs.nb - 1);
        for (int k = 0; k < 5; k++) fri[k] = a.con[((long long)(CF_FRICTION + k) * h.nconmax + id) * S + env];
      }
      g = seg_of(m, a.efc_tree[((long long)2 * r) * S + env], a.efc_tree[((long long)2 * r + 1) * S + env]);
      bs.s1 = g.s1; bs.n1 = g.n1; bs.s2 = g.s2; bs.w = g.n1 + g.n2;
      bs.layout();
    }
    const int w = bs.w, wq = bs.wq, nb = bs.nb;
    const long long slab_env = ((long long)tile * BLOCK + wbase) * capw;   // slab offset of lane 0's environment
    T vel[6], js[6], jq[6], A[6][6];
    const int nbw = __reduce_max_sync(0xffffffffu, nb);
    // The rows of the factor the block's trees need (off-diagonal entries of L in row order, then 1 / D), fetched ONCE
    // per block into the head of the thread's column — free until the parameters are written at the end — when they fit
    // there: the solves of the block's base directions then read them from shared memory (a row fetched inside the solve
    // is one L2 round trip per dof, per pass and per base direction).
    bool ldc = false;
    if (have && !a.em_rows && !a.efc_B) {
      int tot = g.n1 + g.n2;
      for (int i = g.s1; i < g.s1 + g.n1; i++) tot += m.i(h.o_dof_Mcnt, i) - 1;
      for (int i = g.s2; i < g.s2 + g.n2; i++) tot += m.i(h.o_dof_Mcnt, i) - 1;
      ldc = tot <= npar;
      if (ldc) {
        int k = 0;
        for (int sgm = 0; sgm < 2; sgm++) {
          const int lo = sgm ? g.s2 : g.s1, n = sgm ? g.n2 : g.n1;
          for (int i = lo; i < lo + n; i++) {
            const int adr = m.i(h.o_dof_Madr, i), cnt = m.i(h.o_dof_Mcnt, i);
            for (int q0 = 1; q0 < cnt; q0 += 8) {
              T l[8];
#pragma unroll
              for (int q = 0; q < 8; q++) l[q] = q0 + q < cnt ? LD[adr + q0 + q] : T(0);
#pragma unroll
              for (int q = 0; q < 8; q++) if (q0 + q < cnt) rec[k + q0 - 1 + q] = l[q];
            }
            k += cnt - 1;
          }
          for (int i0 = lo; i0 < lo + n; i0 += 8) {
            T dv[8];
#pragma unroll
            for (int q = 0; q < 8; q++) dv[q] = i0 + q < lo + n ? dinv[i0 + q] : T(0);
#pragma unroll
            for (int q = 0; q < 8; q++) if (i0 + q < lo + n) rec[k + i0 - lo + q] = dv[q];
          }
          k += n;
        }
      }
    }
    for (int c = 0; c < nbw; c++) {
      const bool on = have && c < nb;
      if (on) {
        T v0 = 0, s0 = 0, q0 = 0;
        // (four elements at a time, every load of the chunk in flight before the first use: the row and the three state
        //  vectors come from L2, and a load -> test -> load chain per element was two round trips each)
        const bool inv = (a.flags & B2F_INVERSE) != 0;
        for (int e0 = 0; e0 < w; e0 += 4) {
          T j[4], qv[4], qs[4], qa[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const bool in = e0 + k < w;
            const long long d = (long long)bs.dof(in ? e0 + k : 0) * S + env;
            j[k] = in ? a.efc_J[((long long)(r + c) * W + e0 + k) * S + env] : T(0);
            qv[k] = in ? a.qvel[d] : T(0); qs[k] = in ? a.qacc_smooth[d] : T(0); qa[k] = (in && inv) ? a.qacc[d] : T(0);
          }
#pragma unroll
          for (int k = 0; k < 4; k++) {
            if (e0 + k < w) { Jc[e0 + k] = j[k]; Bc[e0 + k] = j[k]; }
            if (j[k] != 0) { v0 += j[k] * qv[k]; s0 += j[k] * qs[k]; if (inv) q0 += j[k] * qa[k]; }
          }
        }
        vel[c] = v0; js[c] = s0; jq[c] = q0;
        for (int e = w; e < wq; e++) { Jc[e] = 0; Bc[e] = 0; }
        // B_c = M^-1 J_c^T: precomputed by k_solve_rows for wide trees, else solved here one tree at a time (M is block
        // diagonal over trees)
        if (a.em_rows) {
          const T* brow = a.efc_Bem + ((long long)env * a.em_rows + (r + c)) * 64;
          for (int e = 0; e < w; e++) Bc[e] = brow[e];
        } else if (a.efc_B) {
          for (int e = 0; e < w; e++) Bc[e] = a.efc_B[((long long)(r + c) * W + e) * S + env];
        } else if (ldc) {
          // same loops, same operation order as below, the factor's rows from the column head
          int k0 = 0;
          for (int sgm = 0; sgm < 2; sgm++) {
            const int lo = sgm ? g.s2 : g.s1, n = sgm ? g.n2 : g.n1, base = sgm ? g.n1 : 0;
            if (n == 0) continue;
            int nl = 0;
            for (int i = lo; i < lo + n; i++) nl += m.i(h.o_dof_Mcnt, i) - 1;
            int k = k0 + nl;
            for (int i = lo + n - 1; i >= lo; i--) {
              const int adr = m.i(h.o_dof_Madr, i), cnt = m.i(h.o_dof_Mcnt, i);
              k -= cnt - 1;
              const T xi = Bc[base + i - lo];
              if (xi == 0) continue;
              for (int q = 1; q < cnt; q++) Bc[base + m.i(h.o_dof_anc, adr + q) - lo] -= rec[k + q - 1] * xi;
            }
            for (int i = lo; i < lo + n; i++) Bc[base + i - lo] *= rec[k0 + nl + i - lo];
            k = k0;
            for (int i = lo; i < lo + n; i++) {
              const int adr = m.i(h.o_dof_Madr, i), cnt = m.i(h.o_dof_Mcnt, i);
              T xi = Bc[base + i - lo];
              for (int q = 1; q < cnt; q++) xi -= rec[k + q - 1] * Bc[base + m.i(h.o_dof_anc, adr + q) - lo];
              Bc[base + i - lo] = xi;
              k += cnt - 1;
            }
            k0 += nl + n;
          }
        } else
        for (int sgm = 0; sgm < 2; sgm++) {
          const int lo = sgm ? g.s2 : g.s1, n = sgm ? g.n2 : g.n1, base = sgm ? g.n1 : 0;
          if (n == 0) continue;
          // (a row of the factor is fetched eight entries at a time before it is used: the entries come from L2, and one
          //  load per update of the shared-memory column was one round trip per entry)
          for (int i = lo + n - 1; i >= lo; i--) {
            const T xi = Bc[base + i - lo];
            if (xi == 0) continue;
            const int adr = m.i(h.o_dof_Madr, i), cnt = m.i(h.o_dof_Mcnt, i);
            for (int q0 = 1; q0 < cnt; q0 += 8) {
              T l[8]; int an[8];
#pragma unroll
              for (int k = 0; k < 8; k++) { const bool in = q0 + k < cnt; l[k] = in ? LD[adr + q0 + k] : T(0); an[k] = in ? m.i(h.o_dof_anc, adr + q0 + k) : lo; }
#pragma unroll
              for (int k = 0; k < 8; k++) if (q0 + k < cnt) Bc[base + an[k] - lo] -= l[k] * xi;
            }
          }
          for (int i0 = lo; i0 < lo + n; i0 += 8) {
            T dv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) dv[k] = i0 + k < lo + n ? dinv[i0 + k] : T(0);
#pragma unroll
            for (int k = 0; k < 8; k++) if (i0 + k < lo + n) Bc[base + i0 + k - lo] *= dv[k];
          }
          for (int i = lo; i < lo + n; i++) {
            const int adr = m.i(h.o_dof_Madr, i), cnt = m.i(h.o_dof_Mcnt, i);
            T xi = Bc[base + i - lo];
            for (int q0 = 1; q0 < cnt; q0 += 8) {
              T l[8]; int an[8];
#pragma unroll
              for (int k = 0; k < 8; k++) { const bool in = q0 + k < cnt; l[k] = in ? LD[adr + q0 + k] : T(0); an[k] = in ? m.i(h.o_dof_anc, adr + q0 + k) : lo; }
#pragma unroll
              for (int k = 0; k < 8; k++) if (q0 + k < cnt) xi -= l[k] * Bc[base + an[k] - lo];
            }
            Bc[base + i - lo] = xi;
          }
        }
        // column c of the local matrix A = J_base B_base^T (symmetric): rows k < c from efc_J, the diagonal from J_c
        for (int k = 0; k <= c; k++) {
          T s0a = 0;
          if (k == c) { for (int e = 0; e < w; e++) s0a += Jc[e] * Bc[e]; }
          else {
            for (int e0 = 0; e0 < w; e0 += 4) {
              T j[4];
#pragma unroll
              for (int k2 = 0; k2 < 4; k2++) j[k2] = e0 + k2 < w ? a.efc_J[((long long)(r + k) * W + e0 + k2) * S + env] : T(0);
#pragma unroll
              for (int k2 = 0; k2 < 4; k2++) if (j[k2] != 0) s0a += j[k2] * Bc[e0 + k2];
            }
          }
          A[k][c] = s0a; A[c][k] = s0a;
        }
      }
      // ---- J_c | B_c leave as 16-byte stores from the thread's own column: every thread writes a contiguous run of its
      //      environment's slab (whole 32-byte sectors after two stores).  The round-1 form — the warp writing one
      //      environment's run at a time for 128-byte coalescing — spent ~2500 instructions per block on the ballot / shuffle
      //      loop for 32-byte runs; the sector writes are the same ----
      if (on) {
        using V4 = VecN<T, 4>;
        T* dst = a.efc_blocks + slab_env + (long long)lane * capw + woff;
        for (int q = 0; q < wq; q += 4) {
          V4 vj, vb;
#pragma unroll
          for (int c2 = 0; c2 < 4; c2++) { vj.v[c2] = Jc[q + c2]; vb.v[c2] = Bc[q + c2]; }
          *reinterpret_cast<V4*>(dst + bs.oJ + c * wq + q) = vj;
          *reinterpret_cast<V4*>(dst + bs.oB + c * wq + q) = vb;
        }
      }
    }
    if (have) {
      const long long o = (long long)r * S + env;
      if (nb > 1)
        for (int r1 = 0; r1 < bs.nrow; r1++)
          for (int r2 = r1 + 1; r2 < bs.nrow; r2++) {
            const int k1 = r1 / 2 + 1, k2 = r2 / 2 + 1;
            const T m1 = (r1 & 1) ? -fri[k1 - 1] : fri[k1 - 1], m2 = (r2 & 1) ? -fri[k2 - 1] : fri[k2 - 1];
            rec[bs.aru(r1, r2)] = A[0][0] + m2 * A[0][k2] + m1 * (A[k1][0] + m2 * A[k1][k2]);
          }
      for (int k = 1; k < nb; k++) rec[bs.oMu + k - 1] = fri[k - 1];
      const T R = a.efc_R[o], D = 1 / R, fl = a.efc_frictionloss[o];
      for (int rr0 = 0; rr0 < bs.nrow; rr0 += 4) {
      // (the row parameters of four rows are fetched before the first row's results are stored: a store to the row arrays
      //  orders every later load behind it)
      T Kq[4], Bq[4], Iq[4], Pq[4], Mq[4];
#pragma unroll
      for (int k2 = 0; k2 < 4; k2++) {
        const int rr = rr0 + k2 < bs.nrow ? rr0 + k2 : rr0;
        const long long orr = (long long)(r + rr) * S + env;
        Kq[k2] = a.efc_KBI[((long long)0 * h.njmax + r + rr) * S + env]; Bq[k2] = a.efc_KBI[((long long)1 * h.njmax + r + rr) * S + env];
        Iq[k2] = a.efc_KBI[((long long)2 * h.njmax + r + rr) * S + env]; Pq[k2] = a.efc_pos[orr]; Mq[k2] = a.efc_margin[orr];
      }
#pragma unroll
      for (int k2 = 0; k2 < 4; k2++) {
        const int rr = rr0 + k2;
        if (rr >= bs.nrow) break;
        const long long orr = (long long)(r + rr) * S + env;
        const int k = nb > 1 ? rr / 2 + 1 : 0;
        const T sm = nb > 1 ? ((rr & 1) ? -fri[k - 1] : fri[k - 1]) : T(0);
        const T velr = vel[0] + sm * vel[k], jsr = js[0] + sm * js[k];
        const T K = Kq[k2], Bd = Bq[k2], imp = Iq[k2];
        const T aref = -Bd * velr - K * imp * (Pq[k2] - Mq[k2]);
        const T bb = jsr - aref;
        const T Arr = (nb > 1 ? A[0][0] + 2 * sm * A[0][k] + sm * sm * A[k][k] : A[0][0]) + R;
        a.efc_D[orr] = D; a.efc_vel[orr] = velr; a.efc_aref[orr] = aref; a.efc_b[orr] = bb; a.efc_ARdiag[orr] = Arr;
        // force implied by the previous tick's acceleration (mj_inverse); summed into qfrc_inverse later, in row order
        if (a.flags & B2F_INVERSE) a.efc_finv[orr] = primal_force(bs.type, jq[0] + sm * jq[k] - aref, D, R, fl);
        rec[bs.oAref + rr] = aref; rec[bs.oArr + rr] = Arr; rec[bs.oiA + rr] = 1 / Arr; rec[bs.ob + rr] = bb;
      }
      }
      for (int q = bs.oA + (nb > 1 ? bs.nrow * (bs.nrow - 1) / 2 : 0); q < bs.oJ; q++) rec[q] = 0;
      rec[BH_CODE] = enc_int(bs.type + 16 * nb + 256 * bs.nrow, T());
      rec[BH_S1] = enc_int(bs.s1, T()); rec[BH_N1W] = enc_int(bs.n1 + 1024 * w, T()); rec[BH_S2] = enc_int(bs.s2, T());
      rec[BH_R] = R; rec[BH_FL] = fl; rec[BH_LEN] = enc_int(bs.len, T()); rec[BH_ROW0] = enc_int(r, T());
    }
    // ---- header + parameters: 16-byte stores from the thread's column ----
    if (have) {
      using V4 = VecN<T, 4>;
      T* dst = a.efc_blocks + slab_env + (long long)lane * capw + woff;
      for (int q = 0; q < bs.oJ; q += 4) {
        V4 v;
#pragma unroll
        for (int c2 = 0; c2 < 4; c2++) v.v[c2] = rec[q + c2];
        *reinterpret_cast<V4*>(dst + q) = v;
      }
    }
  }
}

// mj_inverse without a solve (the MuJoCo-named shim: B2_TICK_INVERSE | B2_TICK_NOSOLVE): qfrc_inverse -= J^T f from the
// rows' base directions in efc_J and the per-row forces, one thread per environment.
template <typename T>
__global__ void k_inverse_rows(const KArgs<T> a) {
  const DModel* h = reinterpret_cast<const DModel*>(a.model);
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= a.ncount) return;
  if ((a.flags & B2F_FUSABLE) && (a.status[env] & 8)) return;
  const long long S = a.nenvp;
  const int ne = a.nefc[env], W = h->wmax;
  if (ne <= 0) return;
  const int nw = a.efc_nwords[env];
  const T* slab = a.efc_blocks + (long long)env * a.block_capw;
  for (int off = 0; off < nw;) {
    const T* rec = slab + off;
    const BlockShape bs = block_shape(rec);
    const int row0 = dec_int(rec[BH_ROW0]);
    T d[6] = {0, 0, 0, 0, 0, 0};
    for (int rr = 0; rr < bs.nrow; rr++) {
      const T fr = a.efc_finv[(long long)(row0 + rr) * S + env];
      const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
      d[0] += fr;
      if (bs.nb > 1) d[k] += ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) * fr;
    }
    for (int e = 0; e < bs.w; e++) {
      T s0 = 0;
      for (int k = 0; k < bs.nb; k++) s0 += d[k] * a.efc_J[((long long)(row0 + k) * W + e) * S + env];
      if (s0 != 0) a.qfrc_inverse[(long long)bs.dof(e) * S + env] -= s0;
    }
    off += bs.len;
  }
}

// One Gauss-Seidel visit of a block.  NBW is a WARP-uniform upper bound of the block's base directions (the teams of a
// warp relax different blocks at the same time: running all of them through one instruction stream, padded rows
// predicated off, keeps the warp converged where a switch on the block's own shape would serialise the teams).
// Everything the visit needs is fetched up front; after the nb dot products v_r = J_r . acc of the pyramid rows is
// tracked through the relaxation with the packed couplings ARu, so a row costs ~7 dependent operations.
template <typename T, int NBW, int LANES>
__device__ __forceinline__ void pgs_visit(const T* __restrict__ rec, const BlockShape& bs, T* acc, T* f, int l, T& improvement) {
  // called by ALL lanes of the warp (full-mask shuffles of width LANES); a team without a block passes nb = nrow = w = 0
  constexpr int NROWW = NBW > 1 ? 2 * (NBW - 1) : 1;
  const int wq = bs.wq, w = bs.w, nb = bs.nb, nrow = bs.nrow;
  const bool have = nrow > 0;
  const int row0 = have ? dec_int(rec[BH_ROW0]) : 0;
  const T R = have ? rec[BH_R] : T(0);
  // admissible interval of the forces of this block
  const T big = T(3.0e38);
  const T flv = have ? rec[BH_FL] : T(0);
  const T lo = bs.type == CN_EQUALITY ? -big : (bs.type == CN_FRICTION_DOF ? -flv : T(0));
  const T hi = bs.type == CN_FRICTION_DOF ? flv : big;
  T aref[NROWW], Arr[NROWW], iA[NROWW], fo[NROWW], mu[NBW > 1 ? NBW - 1 : 1];
#pragma unroll
  for (int r = 0; r < NROWW; r++) {
    const bool on = r < nrow;
    aref[r] = on ? rec[bs.oAref + r] : T(0); Arr[r] = on ? rec[bs.oArr + r] : T(0); iA[r] = on ? rec[bs.oiA + r] : T(0);
    fo[r] = on ? f[row0 + r] : T(0);
  }
#pragma unroll
  for (int k = 0; k < NBW - 1; k++) mu[k] = k < nb - 1 ? rec[bs.oMu + k] : T(0);
  T cpl[NROWW > 1 ? NROWW * (NROWW - 1) / 2 : 1];   // couplings (r, s), r < s, in this function's own static packing
  if (NBW > 1) {
    int q = 0;
#pragma unroll
    for (int r = 0; r < NROWW; r++)
#pragma unroll
      for (int c = r + 1; c < NROWW; c++, q++) cpl[q] = c < nrow ? rec[bs.aru(r, c)] : T(0);
  }
  T u[NBW];
#pragma unroll
  for (int k = 0; k < NBW; k++) u[k] = 0;
  for (int e = l; e < w; e += LANES) {
    const T xv = acc[bs.dof(e)];
#pragma unroll
    for (int k = 0; k < NBW; k++) if (k < nb) u[k] = t_fma(rec[bs.oJ + k * wq + e], xv, u[k]);
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < NBW; k++) u[k] += __shfl_xor_sync(0xffffffffu, u[k], o, LANES);
  }
  T v[NROWW], dl[NROWW];
  if (NBW == 1) v[0] = u[0];
  else {
#pragma unroll
    for (int r = 0; r < NROWW; r++) v[r] = nb > 1 ? t_fma((r & 1) ? -mu[r / 2] : mu[r / 2], u[r / 2 + 1], u[0]) : u[0];
  }
  bool any = false;
  {
    int q = 0;
#pragma unroll
    for (int r = 0; r < NROWW; r++) {
      // (explicit fma / mul: pgs_visit_exact must round every one of these exactly the same way)
      const T res = v[r] + t_fma(R, fo[r], -aref[r]);
      const T fn = t_min(hi, t_max(lo, t_fma(-res, iA[r], fo[r])));
      T delta = fn - fo[r];
      const T change = t_fma(t_mul(t_mul(T(0.5), delta), delta), Arr[r], t_mul(delta, res));
      const bool ok = r < nrow && delta != 0 && !(change > T(1e-10));
      delta = ok ? delta : T(0);
      improvement -= ok ? change : T(0);
      fo[r] = ok ? fn : fo[r];
      any |= ok;
      dl[r] = delta;
#pragma unroll
      for (int c = r + 1; c < NROWW; c++, q++) v[c] = t_fma(delta, cpl[q], v[c]);
    }
  }
  if (any) {
    if constexpr (LANES == 1) {   // one lane per block (k_pgs_island): it stores every row itself
#pragma unroll
      for (int r = 0; r < NROWW; r++) if (r < nrow) f[row0 + r] = fo[r];
    } else
    {  // lane r stores force r (and r + LANES): select chains, not predicated stores per row (those compile to a jump table)
      T mine = fo[0];
#pragma unroll
      for (int r = 1; r < NROWW && r < LANES; r++) mine = (l == r) ? fo[r] : mine;
      if (l < nrow && l < LANES) f[row0 + l] = mine;
      if (NROWW > LANES) {
        T mine2 = fo[NROWW > LANES ? LANES : 0];
#pragma unroll
        for (int r = LANES + 1; r < NROWW; r++) mine2 = (l == r - LANES) ? fo[r] : mine2;
        if (l + LANES < nrow) f[row0 + l + LANES] = mine2;
      }
    }
    T d[NBW];
    d[0] = dl[0];
#pragma unroll
    for (int r = 1; r < NROWW; r++) d[0] += dl[r];
#pragma unroll
    for (int k = 1; k < NBW; k++) d[k] = mu[k - 1] * (dl[2 * k - 2] - dl[2 * k - 1]);
    for (int e = l; e < w; e += LANES) {
      T s0 = 0;
#pragma unroll
      for (int k = 0; k < NBW; k++) if (k < nb) s0 = t_fma(d[k], rec[bs.oB + k * wq + e], s0);
      acc[bs.dof(e)] += s0;
    }
  }
}

// The visit of k_pgs_island: ONE lane relaxes one block, every lane of the warp a block of its own (different
// environments / islands, possibly different shapes).  NBW is the warp-uniform bound of the base directions (1, 3 or 4).
// Everything is fetched up front with 16-byte loads in fixed-trip, predicated loops (the header, the parameter words, the
// <= 16 elements of the acceleration, J and B), so the loads are in flight together and the dependent chain is
// "dot products -> rows -> one update".  The parameter words sit at offsets that depend on the block's own shape; they
// are picked out of the loaded vectors by select chains over the shapes NBW admits (register indices stay static).
// Same operation order as pgs_visit with LANES = 1.
__host__ __device__ constexpr int lv_nrow(int nb) { return nb > 1 ? 2 * (nb - 1) : 1; }
__host__ __device__ constexpr int lv_arr(int nb, int r) { return r < lv_nrow(nb) ? lv_nrow(nb) + r : 0; }
__host__ __device__ constexpr int lv_ia(int nb, int r) { return r < lv_nrow(nb) ? 2 * lv_nrow(nb) + r : 0; }
__host__ __device__ constexpr int lv_mu(int nb, int k) { return k < nb - 1 ? 4 * lv_nrow(nb) + k : 0; }
__host__ __device__ constexpr int lv_cpl(int nb, int r, int c) {
  return c < lv_nrow(nb) ? 4 * lv_nrow(nb) + nb - 1 + r * (2 * lv_nrow(nb) - r - 1) / 2 + (c - r - 1) : 0;
}
__host__ __device__ constexpr int lv_np(int nb) { return (4 * lv_nrow(nb) + (nb > 1 ? nb - 1 + lv_nrow(nb) * (lv_nrow(nb) - 1) / 2 : 0) + 3) & ~3; }
template <typename T, int NBW, int WCAP>
__device__ __forceinline__ void pgs_visit_lane(const T* __restrict__ rec, bool have, const VecN<T, 4>& h0, const VecN<T, 4>& h1,
                                               T* __restrict__ acc, T* __restrict__ f, T& improvement) {
  static_assert(NBW == 1 || NBW == 3 || NBW == 4, "scalar rows and pyramidal contacts of condim 3 / 4");
  static_assert(WCAP % 4 == 0 && WCAP <= 16, "compact rows of at most 16 elements");
  constexpr int NROWW = lv_nrow(NBW), NPV = lv_np(NBW) / 4;
  using V4 = VecN<T, 4>;
  // (h0 / h1: the record's header words, loaded by the caller — zeros for a lane without a block)
  const int code = dec_int(h0.v[0]), type = code & 15, nb = have ? (code >> 4) & 15 : 0, nrow = have ? code >> 8 : 0;
  const int s1 = dec_int(h0.v[1]), n1w = dec_int(h0.v[2]), n1 = n1w & 1023, w = have ? n1w >> 10 : 0, s2 = dec_int(h0.v[3]);
  const T R = h1.v[0], flv = h1.v[1];
  const int row0 = dec_int(h1.v[3]);
  const int wq = (w + 3) & ~3;
  const int np = nb == 1 ? lv_np(1) : (nb == 3 ? lv_np(3) : lv_np(4));
  const int oJ = BH_N + np, oB = oJ + nb * wq;
  T P[NPV * 4];
#pragma unroll
  for (int q = 0; q < NPV; q++) {
    V4 v;
    if (have && 4 * q < np) v = *reinterpret_cast<const V4*>(rec + BH_N + 4 * q);
    else { v.v[0] = 0; v.v[1] = 0; v.v[2] = 0; v.v[3] = 0; }
#pragma unroll
    for (int c = 0; c < 4; c++) P[4 * q + c] = v.v[c];
  }
  // pick a parameter whose offset depends on the shape: i1 / i3 / i4 are its offsets for nb = 1 / 3 / 4
  auto pick = [&](int i1, int i3, int i4) -> T {
    T v = P[i1 < NPV * 4 ? i1 : 0];
    if (NBW >= 3) v = nb == 3 ? P[i3 < NPV * 4 ? i3 : 0] : v;
    if (NBW >= 4) v = nb == 4 ? P[i4 < NPV * 4 ? i4 : 0] : v;
    return v;
  };
  const T big = T(3.0e38);
  const T lo = type == CN_EQUALITY ? -big : (type == CN_FRICTION_DOF ? -flv : T(0));
  const T hi = type == CN_FRICTION_DOF ? flv : big;
  T fo[NROWW];
#pragma unroll
  for (int r = 0; r < NROWW; r++) fo[r] = r < nrow ? f[row0 + r] : T(0);
  // elements of the running acceleration this block touches
  T x[WCAP];
  int dofs[WCAP];
#pragma unroll
  for (int e = 0; e < WCAP; e++) {
    dofs[e] = e < n1 ? s1 + e : s2 + e - n1;
    x[e] = e < w ? acc[dofs[e]] : T(0);
  }
  T u[NBW];
#pragma unroll
  for (int k = 0; k < NBW; k++) {
    u[k] = 0;
#pragma unroll
    for (int q = 0; q < WCAP / 4; q++) {
      if (k < nb && 4 * q < w) {   // (the pad elements of a row are zero in the record)
        const V4 j = *reinterpret_cast<const V4*>(rec + oJ + k * wq + 4 * q);
#pragma unroll
        for (int c = 0; c < 4; c++) u[k] = t_fma(j.v[c], x[4 * q + c], u[k]);
      }
    }
  }
  T Bv[NBW][WCAP];
#pragma unroll
  for (int k = 0; k < NBW; k++)
#pragma unroll
    for (int q = 0; q < WCAP / 4; q++) {
      V4 b;
      if (k < nb && 4 * q < w) b = *reinterpret_cast<const V4*>(rec + oB + k * wq + 4 * q);
      else { b.v[0] = 0; b.v[1] = 0; b.v[2] = 0; b.v[3] = 0; }
#pragma unroll
      for (int c = 0; c < 4; c++) Bv[k][4 * q + c] = b.v[c];
    }
  T mu[NBW > 1 ? NBW - 1 : 1];
#pragma unroll
  for (int k = 0; k < NBW - 1; k++) mu[k] = k < nb - 1 ? pick(0, lv_mu(3, k), lv_mu(4, k)) : T(0);
  T v[NROWW], dl[NROWW];
  if (NBW == 1) v[0] = u[0];
  else {
#pragma unroll
    for (int r = 0; r < NROWW; r++) v[r] = nb > 1 ? t_fma((r & 1) ? -mu[r / 2] : mu[r / 2], u[r / 2 + 1], u[0]) : u[0];
  }
  bool any = false;
#pragma unroll
  for (int r = 0; r < NROWW; r++) {
    const T aref = P[r], Arr = pick(lv_arr(1, r), lv_arr(3, r), lv_arr(4, r)), iA = pick(lv_ia(1, r), lv_ia(3, r), lv_ia(4, r));
    const T res = v[r] + t_fma(R, fo[r], -aref);
    const T fn = t_min(hi, t_max(lo, t_fma(-res, iA, fo[r])));
    T delta = fn - fo[r];
    const T change = t_fma(t_mul(t_mul(T(0.5), delta), delta), Arr, t_mul(delta, res));
    const bool ok = r < nrow && delta != 0 && !(change > T(1e-10));
    delta = ok ? delta : T(0);
    improvement -= ok ? change : T(0);
    fo[r] = ok ? fn : fo[r];
    any |= ok;
    dl[r] = delta;
#pragma unroll
    for (int c = r + 1; c < NROWW; c++) v[c] = t_fma(delta, c < nrow ? pick(0, lv_cpl(3, r, c), lv_cpl(4, r, c)) : T(0), v[c]);
  }
  if (any) {
#pragma unroll
    for (int r = 0; r < NROWW; r++) if (r < nrow) f[row0 + r] = fo[r];
    T d[NBW];
    d[0] = dl[0];
#pragma unroll
    for (int r = 1; r < NROWW; r++) d[0] += dl[r];
#pragma unroll
    for (int k = 1; k < NBW; k++) d[k] = mu[k - 1] * (dl[2 * k - 2] - dl[2 * k - 1]);
#pragma unroll
    for (int e = 0; e < WCAP; e++) {
      T s0 = 0;
#pragma unroll
      for (int k = 0; k < NBW; k++) s0 = t_fma(d[k], Bv[k][e], s0);
      if (e < w) acc[dofs[e]] = x[e] + s0;
    }
  }
}

// The lane visit when every lane of the warp that has a block has a pyramidal contact with exactly NB base directions
// (the common case: most contacts of a scene share a condim) — record offsets are compile-time constants, no parameter
// picks, no row padding, contact bounds only (f >= 0).  SH: every such record is staged in shared memory (the loads
// become LDS instead of generic loads).  Same operation order as pgs_visit_lane.
template <typename T, int NB, int WCAP, bool SH>
__device__ __forceinline__ void pgs_visit_lane_exact(const T* __restrict__ rec, bool have, const VecN<T, 4>& h0, const VecN<T, 4>& h1,
                                                     T* __restrict__ acc, T* __restrict__ f, T& improvement) {
  static_assert(NB == 3 || NB == 4, "pyramidal contacts of condim 3 / 4");
  constexpr int NROW = 2 * (NB - 1), NP = lv_np(NB), NPV = NP / 4, oJ = BH_N + NP;
  using V4 = VecN<T, 4>;
  if (SH) __builtin_assume(__isShared(rec));
  const int s1 = dec_int(h0.v[1]), n1w = dec_int(h0.v[2]), n1 = n1w & 1023, w = have ? n1w >> 10 : 0, s2 = dec_int(h0.v[3]);
  const T R = h1.v[0];
  const int row0 = dec_int(h1.v[3]);
  const int wq = (w + 3) & ~3, oB = oJ + NB * wq;
  T P[NPV * 4];
#pragma unroll
  for (int q = 0; q < NPV; q++) {
    V4 v;
    if (have) v = *reinterpret_cast<const V4*>(rec + BH_N + 4 * q);
    else { v.v[0] = 0; v.v[1] = 0; v.v[2] = 0; v.v[3] = 0; }
#pragma unroll
    for (int c = 0; c < 4; c++) P[4 * q + c] = v.v[c];
  }
  T fo[NROW];
#pragma unroll
  for (int r = 0; r < NROW; r++) fo[r] = have ? f[row0 + r] : T(0);
  T x[WCAP];
  int dofs[WCAP];
#pragma unroll
  for (int e = 0; e < WCAP; e++) {
    dofs[e] = e < n1 ? s1 + e : s2 + e - n1;
    x[e] = e < w ? acc[dofs[e]] : T(0);
  }
  T u[NB];
#pragma unroll
  for (int k = 0; k < NB; k++) {
    u[k] = 0;
#pragma unroll
    for (int q = 0; q < WCAP / 4; q++) {
      if (4 * q < w) {
        const V4 j = *reinterpret_cast<const V4*>(rec + oJ + k * wq + 4 * q);
#pragma unroll
        for (int c = 0; c < 4; c++) u[k] = t_fma(j.v[c], x[4 * q + c], u[k]);
      }
    }
  }
  T Bv[NB][WCAP];
#pragma unroll
  for (int k = 0; k < NB; k++)
#pragma unroll
    for (int q = 0; q < WCAP / 4; q++) {
      V4 b;
      if (4 * q < w) b = *reinterpret_cast<const V4*>(rec + oB + k * wq + 4 * q);
      else { b.v[0] = 0; b.v[1] = 0; b.v[2] = 0; b.v[3] = 0; }
#pragma unroll
      for (int c = 0; c < 4; c++) Bv[k][4 * q + c] = b.v[c];
    }
  T v[NROW], dl[NROW];
#pragma unroll
  for (int r = 0; r < NROW; r++) v[r] = t_fma((r & 1) ? -P[lv_mu(NB, r / 2)] : P[lv_mu(NB, r / 2)], u[r / 2 + 1], u[0]);
  bool any = false;
#pragma unroll
  for (int r = 0; r < NROW; r++) {
    const T res = v[r] + t_fma(R, fo[r], -P[r]);
    const T fn = t_min(T(3.0e38), t_max(T(0), t_fma(-res, P[lv_ia(NB, r)], fo[r])));
    T delta = fn - fo[r];
    const T change = t_fma(t_mul(t_mul(T(0.5), delta), delta), P[lv_arr(NB, r)], t_mul(delta, res));
    const bool ok = have && delta != 0 && !(change > T(1e-10));
    delta = ok ? delta : T(0);
    improvement -= ok ? change : T(0);
    fo[r] = ok ? fn : fo[r];
    any |= ok;
    dl[r] = delta;
#pragma unroll
    for (int c = r + 1; c < NROW; c++) v[c] = t_fma(delta, P[lv_cpl(NB, r, c)], v[c]);
  }
  if (any) {
#pragma unroll
    for (int r = 0; r < NROW; r++) f[row0 + r] = fo[r];
    T d[NB];
    d[0] = dl[0];
#pragma unroll
    for (int r = 1; r < NROW; r++) d[0] += dl[r];
#pragma unroll
    for (int k = 1; k < NB; k++) d[k] = P[lv_mu(NB, k - 1)] * (dl[2 * k - 2] - dl[2 * k - 1]);
#pragma unroll
    for (int e = 0; e < WCAP; e++) {
      T s0 = 0;
#pragma unroll
      for (int k = 0; k < NB; k++) s0 = t_fma(d[k], Bv[k][e], s0);
      if (e < w) acc[dofs[e]] = x[e] + s0;
    }
  }
}

// The same visit when every team of the warp that has a block has one with exactly NB base directions (the common case:
// most contacts of a scene share a condim): all record offsets are compile-time constants, the parameters arrive as
// 16-byte vector loads and no row is padded or predicated.  `have` masks the teams that only keep the warp company
// (their rec points at valid memory of their own slab; whatever they compute is discarded).
template <typename T, int NB, int LANES>
__device__ __forceinline__ void pgs_visit_exact(const T* __restrict__ rec, bool have, int s1, int n1, int s2, int w, int row0, T R,
                                                T* acc, T* f, int l, T& improvement) {
  static_assert(NB == 3 || NB == 4, "pyramidal contacts of condim 3 / 4");
  constexpr int NROW = 2 * (NB - 1);
  constexpr int oAref = BH_N, oArr = oAref + NROW, oiA = oArr + NROW, oMu = oiA + 2 * NROW, oA = oMu + NB - 1;
  constexpr int NCPL = NROW * (NROW - 1) / 2;
  constexpr int oJ = (oA + NCPL + 3) & ~3;
  using V4 = VecN<T, 4>;
  const int wq = (w + 3) & ~3, oB = oJ + NB * wq;
  // parameters: words [oAref, oJ) as vectors (NB = 3: 24 words, NB = 4: 40 words; both offsets are multiples of 4)
  constexpr int NPV = (oJ - oAref) / 4;
  T P[NPV * 4];
#pragma unroll
  for (int q = 0; q < NPV; q++) {
    const V4 v = *reinterpret_cast<const V4*>(rec + oAref + 4 * q);
#pragma unroll
    for (int c = 0; c < 4; c++) P[4 * q + c] = v.v[c];
  }
  auto PAR = [&](int off) -> T { return P[off - oAref]; };
  T fo[NROW];
#pragma unroll
  for (int r = 0; r < NROW; r++) fo[r] = f[row0 + r];
  // dot products with the running acceleration: lane l owns elements l and l + LANES (w <= 2 LANES on this path)
  const int e0 = l, e1 = l + LANES;
  const bool in0 = e0 < w, in1 = e1 < w;
  const int d0 = e0 < n1 ? s1 + e0 : s2 + e0 - n1, d1 = e1 < n1 ? s1 + e1 : s2 + e1 - n1;
  const T x0 = in0 ? acc[d0] : T(0), x1 = in1 ? acc[d1] : T(0);
  T u[NB];
#pragma unroll
  for (int k = 0; k < NB; k++) {
    const T j0 = in0 ? rec[oJ + k * wq + e0] : T(0), j1 = in1 ? rec[oJ + k * wq + e1] : T(0);
    // rounded exactly like the generic visit's lane-strided loop (u = 0; u = fma(j0, x0, u); u = fma(j1, x1, u)): which of
    // the two visits a block gets depends on the other teams of the warp, and an environment's result must not
    u[k] = t_fma(j1, x1, t_mul(j0, x0));
  }
  T b0[NB], b1[NB];   // this lane's elements of B, fetched early: their latency hides behind the reduction and the rows
#pragma unroll
  for (int k = 0; k < NB; k++) { b0[k] = in0 ? rec[oB + k * wq + e0] : T(0); b1[k] = in1 ? rec[oB + k * wq + e1] : T(0); }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < NB; k++) u[k] += __shfl_xor_sync(0xffffffffu, u[k], o, LANES);
  }
  T v[NROW], dl[NROW];
#pragma unroll
  for (int r = 0; r < NROW; r++) v[r] = t_fma((r & 1) ? -PAR(oMu + r / 2) : PAR(oMu + r / 2), u[r / 2 + 1], u[0]);
  bool any = false;
  {
    int q = 0;
#pragma unroll
    for (int r = 0; r < NROW; r++) {
      const T res = v[r] + t_fma(R, fo[r], -PAR(oAref + r));
      const T fn = t_max(T(0), t_fma(-res, PAR(oiA + r), fo[r]));
      T delta = fn - fo[r];
      const T change = t_fma(t_mul(t_mul(T(0.5), delta), delta), PAR(oArr + r), t_mul(delta, res));
      const bool ok = have && delta != 0 && !(change > T(1e-10));
      delta = ok ? delta : T(0);
      improvement -= ok ? change : T(0);
      fo[r] = ok ? fn : fo[r];
      any |= ok;
      dl[r] = delta;
#pragma unroll
      for (int c = r + 1; c < NROW; c++, q++) v[c] = t_fma(delta, PAR(oA + q), v[c]);
    }
  }
  if (any) {
    {  // lane r stores force r: a select chain instead of NROW predicated stores (which compile to a jump table)
      T mine = fo[0];
#pragma unroll
      for (int r = 1; r < NROW; r++) mine = (l == r) ? fo[r] : mine;
      if (l < NROW) f[row0 + l] = mine;
    }
    T d[NB];
    d[0] = dl[0];
#pragma unroll
    for (int r = 1; r < NROW; r++) d[0] += dl[r];
#pragma unroll
    for (int k = 1; k < NB; k++) d[k] = PAR(oMu + k - 1) * (dl[2 * k - 2] - dl[2 * k - 1]);
    T a0 = 0, a1 = 0;
#pragma unroll
    for (int k = 0; k < NB; k++) { a0 = t_fma(d[k], b0[k], a0); a1 = t_fma(d[k], b1[k], a1); }
    if (in0) acc[d0] = x0 + a0;
    if (in1) acc[d1] = x1 + a1;
  }
}

// Visit order of the solver: environments sorted by descending record volume (a counting sort over NB bins of
// words / BINW, one CTA; the order inside a bin is whatever the atomics give, which affects scheduling only).
template <int NBIN, int THREADS>
__global__ void __launch_bounds__(THREADS) k_order_envs(const int* __restrict__ nefc, const int* __restrict__ nwords, int* __restrict__ order,
                                                        int nenvp, int binw, const int* pending, int use_pending, const int* __restrict__ iters, int) {
  if (use_pending && pending[0] == 0) return;
  __shared__ int hist[NBIN], start[NBIN];
  for (int i = threadIdx.x; i < NBIN; i += THREADS) hist[i] = 0;
  __syncthreads();
  // predicted work (island solver; iters == null: record volume only): the iterations the environment's solve took LAST
  // tick first (contact states persist from tick to tick), the record volume second — the few environments that use all
  // 100 iterations set the kernel's duration: they must start first, and together, because a warp steps its four
  // environments in lockstep
  auto bin_of = [&](int e) {
    const int w = nefc[e] > 0 ? nwords[e] : 0;
    int b = w / binw;
    if (iters) {
      const int it = nefc[e] > 0 ? (iters[e] < 100 ? iters[e] : 100) : 0;
      b = 2 * it + (b / 5 < 54 ? b / 5 : 54);
    }
    return NBIN - 1 - (b < NBIN ? b : NBIN - 1);
  };
  for (int e = threadIdx.x; e < nenvp; e += THREADS) atomicAdd(&hist[bin_of(e)], 1);
  __syncthreads();
  if (threadIdx.x == 0) { int s0 = 0; for (int i = 0; i < NBIN; i++) { start[i] = s0; s0 += hist[i]; } }
  __syncthreads();
  for (int e = threadIdx.x; e < nenvp; e += THREADS) order[atomicAdd(&start[bin_of(e)], 1)] = e;
}

// K6: projected Gauss-Seidel (A.8) in acceleration space, one BLOCK at a time.  A team of LANES lanes owns one
// environment; the running acceleration and the forces live in shared memory, and so do the first `stage_cap` words of
// the environment's records (copied once, read `iterations` times; anything beyond is streamed from the slab).  Per
// visit of a contact: u_k = J_k . acc for its nb base directions (lane-strided products + a team reduction), then the
// contact's 2 (nb - 1) pyramid rows are relaxed in order on the local nb x nb matrix A (u is updated by columns of A
// after every row, which is exactly what the row-by-row sweep over J_r = J_0 +/- mu_k J_k would have seen), and the
// acceleration is corrected once by the accumulated force changes.  Same iterates as the row-by-row dual PGS.
template <typename T, int LANES, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_pgs_block(const KArgs<T> a) {
  if ((a.flags & B2F_FUSABLE) && a.pending[0] == 0) return;
  // the solver needs a handful of scalars of the model, not the staged blob: read them from the header in HBM (cached)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DModel& h = *reinterpret_cast<const DModel*>(a.model);
  MV<T> m{&h, a.model};
  const long long S = a.nenvp;
  constexpr int EPB = BLOCK / LANES;  // environments per CTA
  const int nv = h.nv, njmax = h.njmax, capw = a.block_capw, cap = a.stage_cap;
  const int nvs = (nv + 7) & ~3;  // stride of the per-environment vectors: nv + 4 (skews the teams over the banks), rounded so that the staged records behind them stay 16 / 32-byte aligned
  T* accsh = reinterpret_cast<T*>(smem_raw);                             // [EPB][nvs] running acceleration
  T* tmpsh = accsh + (size_t)EPB * nvs;                                   // [EPB][nvs] M^-1 J^T f, later qfrc_constraint
  T* fsh = tmpsh + (size_t)EPB * nvs;                                     // [EPB][njmax] forces
  T* stsh = fsh + (size_t)EPB * ((njmax + 3) & ~3);                       // [EPB][cap] staged records
  const int team = threadIdx.x / LANES, l = threadIdx.x % LANES;
  const unsigned tmask = (LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u)) << ((threadIdx.x & 31) & ~(LANES - 1));
  T* acc = accsh + (size_t)team * nvs;
  T* tmp = tmpsh + (size_t)team * nvs;
  T* f = fsh + (size_t)team * ((njmax + 3) & ~3);
  T* st = stsh + (size_t)team * cap;
  const int ngroups = (a.ncount + EPB - 1) / EPB;
  const T tol = m.f(h.o_opt_real, 3), scale = 1 / (m.f(h.o_opt_real, 4) * T(nv > 1 ? nv : 1));

  auto team_sum = [&](T v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(tmask, v, o);
    return v;
  };
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    // environments are visited in the order of k_order_envs: most constraint words first, so that the longest serial
    // chains start at once and the teams of a warp carry similar loads (nenvp is a multiple of 128 >= EPB: in range)
    const long long env = a.env_order[(long long)grp * EPB + team];
    // integrated by the smooth kernel already (team-uniform): such a team only keeps the warp's collectives company
    const bool skip = (a.flags & B2F_FUSABLE) && (a.status[env] & 8);
    const int ne = skip ? 0 : a.nefc[env], nw = ne > 0 ? a.efc_nwords[env] : 0;
    const T* slab = a.efc_blocks + env * capw;
    int iters = 0;
    // stage the leading records (16-byte copies: capw, cap and every record length are multiples of 4 words)
    {
      const int nst = nw < cap ? nw : cap;
      constexpr int VW = 16 / (int)sizeof(T);
      using V = VecN<T, VW>;
      for (int q = l * VW; q < nst; q += LANES * VW) *reinterpret_cast<V*>(st + q) = *reinterpret_cast<const V*>(slab + q);
    }
    for (int i = l; i < nv; i += LANES) { acc[i] = a.qacc_warmstart[(long long)i * S + env]; tmp[i] = 0; }
    __syncwarp(tmask);
    // record at word offset off: from shared memory when it was staged whole, else from the slab
    auto rec_at = [&](int off) -> const T* {
      if (off + BH_N <= cap) { const int len = dec_int(st[off + BH_LEN]); if (off + len <= cap) return st + off; }
      return slab + off;
    };
    // u[k] = J_k . x for the nb base directions of a block
    auto base_dots = [&](const T* rec, const BlockShape& bs, const T* x, T* u) {
#pragma unroll
      for (int k = 0; k < 6; k++) u[k] = 0;
      for (int e = l; e < bs.w; e += LANES) {
        const T xv = x[bs.dof(e)];
#pragma unroll
        for (int k = 0; k < 6; k++) if (k < bs.nb) u[k] += rec[bs.oJ + k * bs.wq + e] * xv;
      }
#pragma unroll
      for (int k = 0; k < 6; k++) if (k < bs.nb) u[k] = team_sum(u[k]);
    };
    // x += sum_k d[k] X_k  (X = J or B rows of the block)
    auto base_axpy = [&](const T* rec, const BlockShape& bs, int oX, const T* d, T* x) {
      for (int e = l; e < bs.w; e += LANES) {
        T s0 = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) if (k < bs.nb) s0 += d[k] * rec[oX + k * bs.wq + e];
        x[bs.dof(e)] += s0;
      }
    };
    if (ne > 0) {
      // ---- warm start: forces implied by qacc_warmstart (held in acc), kept only if their dual cost is negative ----
      bool warm = !(h.disableflags & DSBL_WARMSTART);
      if (warm) {
        for (int off = 0; off < nw;) {
          const T* rec = rec_at(off);
          const BlockShape bs = block_shape(rec);
          T u[6], d[6] = {0, 0, 0, 0, 0, 0};
          base_dots(rec, bs, acc, u);
          const int row0 = dec_int(rec[BH_ROW0]);
          const T R = rec[BH_R];
          bool any = false;
          for (int rr = 0; rr < bs.nrow; rr++) {
            const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
            const T sm = bs.nb > 1 ? ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) : T(0);
            const T fr = primal_force(bs.type, u[0] + sm * u[k] - rec[bs.oAref + rr], 1 / R, R, rec[BH_FL]);
            if (l == 0) f[row0 + rr] = fr;
            if (fr != 0) { any = true; d[0] += fr; if (bs.nb > 1) d[k] += sm * fr; }
          }
          if (any) base_axpy(rec, bs, bs.oB, d, tmp);
          __syncwarp(tmask);
          off += bs.len;
        }
        T cost = 0;
        for (int off = 0; off < nw;) {
          const T* rec = rec_at(off);
          const BlockShape bs = block_shape(rec);
          const int row0 = dec_int(rec[BH_ROW0]);
          bool any = false;
          for (int rr = 0; rr < bs.nrow; rr++) any |= f[row0 + rr] != 0;
          if (any) {
            T u[6];
            base_dots(rec, bs, tmp, u);
            for (int rr = 0; rr < bs.nrow; rr++) {
              const T fr = f[row0 + rr];
              if (fr == 0) continue;
              const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
              const T sm = bs.nb > 1 ? ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) : T(0);
              const T Af = u[0] + sm * u[k] + rec[BH_R] * fr;
              cost += fr * (T(0.5) * Af + rec[bs.ob + rr]);
            }
          }
          off += bs.len;
        }
        if (cost > 0) warm = false;
      }
      __syncwarp(tmask);
      if (warm) {
        for (int i = l; i < nv; i += LANES) acc[i] = a.qacc_smooth[(long long)i * S + env] + tmp[i];
      } else {
        for (int i = l; i < nv; i += LANES) acc[i] = a.qacc_smooth[(long long)i * S + env];
        for (int r = l; r < ne; r += LANES) f[r] = 0;
      }
      __syncwarp(tmask);
    }
    // ---- Gauss-Seidel sweeps over the blocks: every lane of the warp takes part in every visit (a team that has run
    //      out of blocks, converged, or has no constraint at all relaxes an empty block), so the loop is warp-uniform,
    //      the shuffles use the full mask and the teams never serialise ----
    {
      bool done = ne <= 0;
      using V4 = VecN<T, 4>;
      V4 nh0 = *reinterpret_cast<const V4*>(slab), nh1 = *reinterpret_cast<const V4*>(slab + 4);
      for (int it = 0; it < h.iterations; it++) {
        if (__all_sync(0xffffffffu, done)) break;
        T improvement = 0;
        int off = 0;
        while (true) {
          const bool have = !done && off < nw;
          if (!__any_sync(0xffffffffu, have)) break;
          const T* rec = slab + (have ? off : 0);
          // header {code, s1, n1 | w, s2} {R, frictionloss, len, row0}: fetched one visit ahead (nh0 / nh1), so its latency
          // is never on the critical path
          const V4 h0 = nh0, h1 = nh1;
          const int code = dec_int(h0.v[0]), nb = have ? (code >> 4) & 15 : 0, len = have ? dec_int(h1.v[2]) : 0;
          const int n1w = dec_int(h0.v[2]), w = have ? n1w >> 10 : 0;   // w = 0: a team without a block touches no element
          if (have) {  // next record (at the end of a sweep: the first again): header into registers, body towards L1
            const int nxt = off + len < nw ? off + len : 0;
            nh0 = *reinterpret_cast<const V4*>(slab + nxt); nh1 = *reinterpret_cast<const V4*>(slab + nxt + 4);
            const int pq = nxt + l * (128 / (int)sizeof(T));
            if (pq < nw && l < 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(slab + pq));
          }
          // one instruction stream for the whole warp: the exact-shape visit when every team that has a block has the same
          // condim and fits two elements per lane, else the padded generic visit
          const bool fits = !have || w <= 2 * LANES;
          const bool all3 = __all_sync(0xffffffffu, (!have || nb == 3) && fits), all4 = __all_sync(0xffffffffu, (!have || nb == 4) && fits);
          if (all3) pgs_visit_exact<T, 3, LANES>(rec, have, dec_int(h0.v[1]), n1w & 1023, dec_int(h0.v[3]), w, have ? dec_int(h1.v[3]) : 0, h1.v[0], acc, f, l, improvement);
          else if (all4) pgs_visit_exact<T, 4, LANES>(rec, have, dec_int(h0.v[1]), n1w & 1023, dec_int(h0.v[3]), w, have ? dec_int(h1.v[3]) : 0, h1.v[0], acc, f, l, improvement);
          else {
            BlockShape bs{};
            if (have) bs = block_shape(rec);
            const int nbw = __reduce_max_sync(0xffffffffu, bs.nb);
            if (nbw <= 1) pgs_visit<T, 1, LANES>(rec, bs, acc, f, l, improvement);
            else if (nbw <= 3) pgs_visit<T, 3, LANES>(rec, bs, acc, f, l, improvement);
            else if (nbw <= 4) pgs_visit<T, 4, LANES>(rec, bs, acc, f, l, improvement);
            else pgs_visit<T, 6, LANES>(rec, bs, acc, f, l, improvement);
          }
          __syncwarp();
          off += len;
        }
        if (!done) { iters = it + 1; if (improvement * scale < tol) done = true; }
      }
    }
    if (ne > 0) {
      // ---- qfrc_constraint = J^T f ----
      for (int i = l; i < nv; i += LANES) tmp[i] = 0;
      __syncwarp(tmask);
      for (int off = 0; off < nw;) {
        const T* rec = rec_at(off);
        const BlockShape bs = block_shape(rec);
        const int row0 = dec_int(rec[BH_ROW0]);
        T d[6] = {0, 0, 0, 0, 0, 0};
        bool any = false;
        for (int rr = 0; rr < bs.nrow; rr++) {
          const T fr = f[row0 + rr];
          if (l == 0) a.efc_force[(long long)(row0 + rr) * S + env] = fr;
          if (fr == 0) continue;
          const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
          any = true;
          d[0] += fr;
          if (bs.nb > 1) d[k] += ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) * fr;
        }
        if (any) base_axpy(rec, bs, bs.oJ, d, tmp);
        __syncwarp(tmask);
        off += bs.len;
      }
    } else {
      for (int i = l; i < nv; i += LANES) acc[i] = a.qacc_smooth[(long long)i * S + env];
    }
    __syncwarp(tmask);
    if (!skip) {
      for (int i = l; i < nv; i += LANES) {
        const T v = acc[i];
        a.qacc[(long long)i * S + env] = v;
        a.qacc_warmstart[(long long)i * S + env] = v;
        a.qfrc_constraint[(long long)i * S + env] = tmp[i];
      }
      if (l == 0) a.solver_iter[env] = iters;
    }
    __syncwarp(tmask);
    // ---- mj_inverse: qfrc_inverse -= J^T f(qacc of the previous tick), forces per row from k_make_blocks ----
    if ((a.flags & B2F_INVERSE) && ne > 0) {
      for (int i = l; i < nv; i += LANES) tmp[i] = 0;
      __syncwarp(tmask);
      for (int off = 0; off < nw;) {
        const T* rec = rec_at(off);
        const BlockShape bs = block_shape(rec);
        const int row0 = dec_int(rec[BH_ROW0]);
        T d[6] = {0, 0, 0, 0, 0, 0};
        bool any = false;
        for (int rr = 0; rr < bs.nrow; rr++) {
          const T fr = a.efc_finv[(long long)(row0 + rr) * S + env];
          if (fr == 0) continue;
          const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
          any = true;
          d[0] += fr;
          if (bs.nb > 1) d[k] += ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) * fr;
        }
        if (any) base_axpy(rec, bs, bs.oJ, d, tmp);
        __syncwarp(tmask);
        off += bs.len;
      }
      for (int i = l; i < nv; i += LANES) if (tmp[i] != 0) a.qfrc_inverse[(long long)i * S + env] -= tmp[i];
      __syncwarp(tmask);
    }
  }
}

// K6': the same solver for models made of many small kinematic trees (an arm and free objects, a field of object slots):
// one LANE per constraint ISLAND.  Blocks that share no tree commute exactly — they read and write disjoint entries of
// the acceleration and of the forces — so relaxing the islands of an environment side by side produces the iterates of
// the row-ordered sweep bit for bit, while the serial chain of an iteration shrinks from "all blocks of the environment"
// to "the blocks of its largest island" (C3: ~14 -> ~4).  A team of ISL lanes owns an environment (islands i = lane,
// lane + ISL, ...); every lane runs the scalar visit (pgs_visit with LANES = 1: no shuffles, no redundant copies of the
// row relaxation in eight lanes), the teams of a warp step through their blocks in lockstep so that one padded
// instruction stream serves all 32 lanes.  The environment's records are staged in shared memory once (16-byte
// coalesced copies) and read from there every iteration; only the iteration's convergence test is a team reduction.
template <typename T, int ISL, int MINB>
__global__ void __launch_bounds__(32, MINB) k_pgs_island(const KArgs<T> a) {
  if ((a.flags & B2F_FUSABLE) && a.pending[0] == 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DModel& h = *reinterpret_cast<const DModel*>(a.model);
  MV<T> m{&h, a.model};
  const long long S = a.nenvp;
  constexpr int EPB = 32 / ISL;
  const int nv = h.nv, njmax = h.njmax, capw = a.block_capw, cap = a.stage_cap;
  const int nvs = (nv + 7) & ~3;   // (nv + 4 rounded up to a multiple of 4: the staged records behind stay aligned)
  T* accsh = reinterpret_cast<T*>(smem_raw);
  T* tmpsh = accsh + (size_t)EPB * nvs;
  T* fsh = tmpsh + (size_t)EPB * nvs;
  T* finvsh = fsh + (size_t)EPB * ((njmax + 3) & ~3);   // per-row forces of mj_inverse (efc_finv), fetched with the prologue
  T* stsh = finvsh + (size_t)EPB * ((njmax + 3) & ~3);
  // [EPB][3][isl_cap]: where island i was staged (-1: not staged), and the island table of the environment (word range
  // of island i in the slab) — read by every walk and every sweep, so kept next to the records
  int* stosh = reinterpret_cast<int*>(stsh + (size_t)EPB * cap);
  const int team = threadIdx.x / ISL, l = threadIdx.x % ISL;
  const unsigned tmask = (ISL == 32 ? 0xffffffffu : ((1u << ISL) - 1u)) << (threadIdx.x & ~(ISL - 1));
  T* acc = accsh + (size_t)team * nvs;
  T* tmp = tmpsh + (size_t)team * nvs;
  T* f = fsh + (size_t)team * ((njmax + 3) & ~3);
  T* finv = finvsh + (size_t)team * ((njmax + 3) & ~3);
  T* st = stsh + (size_t)team * cap;
  int* sto = stosh + (size_t)team * 3 * a.isl_cap;
  int* iso = sto + a.isl_cap;
  int* ise = iso + a.isl_cap;
  const int ngroups = (a.ncount + EPB - 1) / EPB;
  const T tol = m.f(h.o_opt_real, 3), scale = 1 / (m.f(h.o_opt_real, 4) * T(nv > 1 ? nv : 1));
  auto team_sum = [&](T v) {
#pragma unroll
    for (int o = ISL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(tmask, v, o, ISL);
    return v;
  };
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long long env = a.env_order[(long long)grp * EPB + team];
    const bool skip = (a.flags & B2F_FUSABLE) && (a.status[env] & 8);
    const int ne = skip ? 0 : a.nefc[env], nw = ne > 0 ? a.efc_nwords[env] : 0, nisl = ne > 0 ? a.nisl[env] : 0;
    const T* slab = a.efc_blocks + env * capw;
    int iters = 0;
    __syncwarp();   // the previous group's readers are done with the shared vectors
    // everything the solve reads once per environment, in flight together: the island table, the warm-start acceleration
    // and mj_inverse's row forces (one L2 / HBM round trip each when they were fetched where they are used)
    const bool invf = (a.flags & B2F_INVERSE) && ne > 0;
#pragma unroll 2
    for (int i = l; i < nisl; i += ISL) { iso[i] = a.isl_off[(long long)i * S + env]; ise[i] = a.isl_end[(long long)i * S + env]; }
#pragma unroll 4
    for (int i = l; i < nv; i += ISL) { acc[i] = a.qacc_warmstart[(long long)i * S + env]; tmp[i] = 0; }
    if (invf) {
#pragma unroll 4
      for (int r = l; r < ne; r += ISL) finv[r] = a.efc_finv[(long long)r * S + env];
    }
    __syncwarp(tmask);
    {
      // stage the records island by island: island i starts on the bank group of the lane that will walk it (records
      // are whole 128-byte lines, so the lanes of a quarter-warp keep reading disjoint banks while their blocks have
      // the same shape)
      // (asynchronous 16-byte copies, cp.async: no register round trip, every load of the environment in flight at once —
      //  a load -> store loop serialised on the L2 latency, ~40 k cycles per environment group)
      int pos = 0;
      for (int i = 0; i < nisl; i++) {
        const int o = iso[i], len = ise[i] - o;
        const int start = ((pos + 31) & ~31) + 4 * ((team * ISL + (i % ISL)) & 7);
        const bool fits = start + len <= cap;
        if (fits) {
          const char* src = reinterpret_cast<const char*>(slab + o);
          const uint32_t dst = smem_u32(st + start);
          const int bytes = len * (int)sizeof(T);
          for (int q = l * 16; q < bytes; q += ISL * 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)q), "l"(src + q) : "memory");
          pos = start + len;
        }
        if (l == 0) sto[i] = fits ? start : -1;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    // this lane's blocks: islands l, l + ISL, ... one after the other
    struct Cursor { int isl, off, end; const T* base; };   // the record at slab offset off is base + off
    auto cur_load = [&](Cursor& c) {
      c.base = slab;
      if (c.isl < nisl) {
        c.off = iso[c.isl]; c.end = ise[c.isl];
        const int so = sto[c.isl];
        if (so >= 0) c.base = st + so - c.off;
      } else { c.off = 0; c.end = 0; }
    };
    auto cur_init = [&](Cursor& c) { c.isl = l; cur_load(c); };
    auto cur_next = [&](Cursor& c, int len) {
      c.off += len;
      if (c.off >= c.end) { c.isl += ISL; cur_load(c); }
    };
    // (measured: 16-byte loads in fixed-trip, predicated loops made these four per-solve walks slower than the plain
    //  loops — more registers and code for work that runs once per solve, profiles/r02_pgs_analysis.txt)
    auto base_dots = [&](const T* rec, const BlockShape& bs, const T* x, T* u) {
#pragma unroll
      for (int k = 0; k < 6; k++) u[k] = 0;
      for (int e = 0; e < bs.w; e++) {
        const T xv = x[bs.dof(e)];
#pragma unroll
        for (int k = 0; k < 6; k++) if (k < bs.nb) u[k] += rec[bs.oJ + k * bs.wq + e] * xv;
      }
    };
    auto base_axpy = [&](const T* rec, const BlockShape& bs, int oX, const T* d, T* x) {
      for (int e = 0; e < bs.w; e++) {
        T s0 = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) if (k < bs.nb) s0 += d[k] * rec[oX + k * bs.wq + e];
        x[bs.dof(e)] += s0;
      }
    };
    {
      // ---- warm start: forces implied by qacc_warmstart (held in acc), kept only if their dual cost is negative ----
      bool warm = ne > 0 && !(h.disableflags & DSBL_WARMSTART);
      T cost = 0;
      if (warm) {
        Cursor c;
        for (cur_init(c); c.off < c.end;) {
          const T* rec = c.base + c.off;
          const BlockShape bs = block_shape(rec);
          T u[6], d[6] = {0, 0, 0, 0, 0, 0};
          base_dots(rec, bs, acc, u);
          const int row0 = dec_int(rec[BH_ROW0]);
          const T R = rec[BH_R];
          bool any = false;
          for (int rr = 0; rr < bs.nrow; rr++) {
            const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
            const T sm = bs.nb > 1 ? ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) : T(0);
            const T fr = primal_force(bs.type, u[0] + sm * u[k] - rec[bs.oAref + rr], 1 / R, R, rec[BH_FL]);
            f[row0 + rr] = fr;
            if (fr != 0) { any = true; d[0] += fr; if (bs.nb > 1) d[k] += sm * fr; }
          }
          if (any) base_axpy(rec, bs, bs.oB, d, tmp);
          cur_next(c, bs.len);
        }
        // (tmp is complete for this lane's islands: the cost of an island needs only its own entries)
        for (cur_init(c); c.off < c.end;) {
          const T* rec = c.base + c.off;
          const BlockShape bs = block_shape(rec);
          const int row0 = dec_int(rec[BH_ROW0]);
          bool any = false;
          for (int rr = 0; rr < bs.nrow; rr++) any |= f[row0 + rr] != 0;
          if (any) {
            T u[6];
            base_dots(rec, bs, tmp, u);
            for (int rr = 0; rr < bs.nrow; rr++) {
              const T fr = f[row0 + rr];
              if (fr == 0) continue;
              const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
              const T sm = bs.nb > 1 ? ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) : T(0);
              const T Af = u[0] + sm * u[k] + rec[BH_R] * fr;
              cost += fr * (T(0.5) * Af + rec[bs.ob + rr]);
            }
          }
          cur_next(c, bs.len);
        }
      }
      __syncwarp(tmask);
      cost = team_sum(cost);
      if (cost > 0) warm = false;
      if (ne > 0) {
        if (warm) {
#pragma unroll 4
          for (int i = l; i < nv; i += ISL) acc[i] = a.qacc_smooth[(long long)i * S + env] + tmp[i];
        } else {
#pragma unroll 4
          for (int i = l; i < nv; i += ISL) acc[i] = a.qacc_smooth[(long long)i * S + env];
          for (int r = l; r < ne; r += ISL) f[r] = 0;
        }
      }
    }
    __syncwarp();
    // ---- Gauss-Seidel sweeps: the lanes of the warp step through their own blocks in lockstep ----
    {
      bool done = ne <= 0;
      for (int it = 0; it < h.iterations; it++) {
        if (__all_sync(0xffffffffu, done)) break;
        T improvement = 0;
        Cursor c;
        cur_init(c);
        if (done) { c.off = 0; c.end = 0; }
        while (true) {
          const bool have = c.off < c.end;
          const T* rec = have ? c.base + c.off : st;   // (a lane without a block: any valid shared address, never read)
          using V4 = VecN<T, 4>;
          V4 h0, h1;
          if (have) { h0 = *reinterpret_cast<const V4*>(rec); h1 = *reinterpret_cast<const V4*>(rec + 4); }
          else {
#pragma unroll
            for (int q2 = 0; q2 < 4; q2++) { h0.v[q2] = enc_int(0, T()); h1.v[q2] = enc_int(0, T()); }
          }
          const int code = dec_int(h0.v[0]), len = dec_int(h1.v[2]);
          // one instruction stream for the warp, sized by the largest block of this step (base directions, row width):
          // both maxima in one warp reduction
          const int nbw = __reduce_max_sync(0xffffffffu, (code >> 4) & 15);
          if (nbw == 0) break;   // no lane of the warp has a block left in this sweep
          const int ww = __reduce_max_sync(0xffffffffu, dec_int(h0.v[2]) >> 10);
          // every lane with a block has a condim-3 contact staged in shared memory (the common step): exact-shape visit
          const bool ex3 = __all_sync(0xffffffffu, !have || (((code >> 4) & 15) == 3 && c.base != slab));
          // (an exact visit for condim 4 as well was measured and lost: the extra vote per step costs the tail more than
          //  the shorter visit gains where condim-4 steps are the minority, profiles/r02_pgs_analysis.txt)
          // (an exact visit for blocks that lie in one kinematic tree — one base address, immediate offsets — was measured
          //  and lost as well: C5 PGS 1.79 -> 1.98 ms)
          if (ex3 && ww <= 8) pgs_visit_lane_exact<T, 3, 8, true>(rec, have, h0, h1, acc, f, improvement);
          else if (ex3 && ww <= 12) pgs_visit_lane_exact<T, 3, 12, true>(rec, have, h0, h1, acc, f, improvement);
          else if (nbw <= 4) {
            if (ww <= 8) {
              if (nbw <= 1) pgs_visit_lane<T, 1, 8>(rec, have, h0, h1, acc, f, improvement);
              else if (nbw <= 3) pgs_visit_lane<T, 3, 8>(rec, have, h0, h1, acc, f, improvement);
              else pgs_visit_lane<T, 4, 8>(rec, have, h0, h1, acc, f, improvement);
            } else if (ww <= 12) {
              if (nbw <= 1) pgs_visit_lane<T, 1, 12>(rec, have, h0, h1, acc, f, improvement);
              else if (nbw <= 3) pgs_visit_lane<T, 3, 12>(rec, have, h0, h1, acc, f, improvement);
              else pgs_visit_lane<T, 4, 12>(rec, have, h0, h1, acc, f, improvement);
            } else {
              if (nbw <= 1) pgs_visit_lane<T, 1, 16>(rec, have, h0, h1, acc, f, improvement);
              else if (nbw <= 3) pgs_visit_lane<T, 3, 16>(rec, have, h0, h1, acc, f, improvement);
              else pgs_visit_lane<T, 4, 16>(rec, have, h0, h1, acc, f, improvement);
            }
          } else {
            BlockShape bs{};
            if (have) bs = block_shape(rec);
            pgs_visit<T, 6, 1>(rec, bs, acc, f, 0, improvement);
          }
          if (have) cur_next(c, len);
        }
        __syncwarp();
        improvement = team_sum(improvement);
        if (!done) { iters = it + 1; if (improvement * scale < tol) done = true; }
#ifdef B2_DEBUG_PGS
        if (l == 0 && ne > 0 && (it & 7) == 7 && it / 8 < 12) a.efc_vel[(long long)(it / 8) * S + env] = improvement * scale;
#endif
      }
    }
    __syncwarp();
    if (ne > 0) {
      // ---- qfrc_constraint = J^T f ----
      for (int i = l; i < nv; i += ISL) tmp[i] = 0;
      __syncwarp(tmask);
      Cursor c;
      for (cur_init(c); c.off < c.end;) {
        const T* rec = c.base + c.off;
        const BlockShape bs = block_shape(rec);
        const int row0 = dec_int(rec[BH_ROW0]);
        T d[6] = {0, 0, 0, 0, 0, 0};
        bool any = false;
        for (int rr = 0; rr < bs.nrow; rr++) {
          const T fr = f[row0 + rr];
          a.efc_force[(long long)(row0 + rr) * S + env] = fr;
          if (fr == 0) continue;
          const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
          any = true;
          d[0] += fr;
          if (bs.nb > 1) d[k] += ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) * fr;
        }
        if (any) base_axpy(rec, bs, bs.oJ, d, tmp);
        cur_next(c, bs.len);
      }
    } else {
      for (int i = l; i < nv; i += ISL) acc[i] = a.qacc_smooth[(long long)i * S + env];
    }
    __syncwarp();
    if (!skip) {
      for (int i = l; i < nv; i += ISL) {
        const T v = acc[i];
        a.qacc[(long long)i * S + env] = v;
        a.qacc_warmstart[(long long)i * S + env] = v;
        a.qfrc_constraint[(long long)i * S + env] = tmp[i];
      }
      if (l == 0) a.solver_iter[env] = iters;
    }
    __syncwarp();
    // ---- mj_inverse: qfrc_inverse -= J^T f(qacc of the previous tick), forces per row from k_make_blocks ----
    if ((a.flags & B2F_INVERSE) && ne > 0) {
      for (int i = l; i < nv; i += ISL) tmp[i] = 0;
      __syncwarp(tmask);
      Cursor c;
      for (cur_init(c); c.off < c.end;) {
        const T* rec = c.base + c.off;
        const BlockShape bs = block_shape(rec);
        const int row0 = dec_int(rec[BH_ROW0]);
        T d[6] = {0, 0, 0, 0, 0, 0};
        bool any = false;
        for (int rr = 0; rr < bs.nrow; rr++) {
          const T fr = finv[row0 + rr];
          if (fr == 0) continue;
          const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
          any = true;
          d[0] += fr;
          if (bs.nb > 1) d[k] += ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) * fr;
        }
        if (any) base_axpy(rec, bs, bs.oJ, d, tmp);
        cur_next(c, bs.len);
      }
      __syncwarp(tmask);
#pragma unroll 4
      for (int i = l; i < nv; i += ISL) if (tmp[i] != 0) a.qfrc_inverse[(long long)i * S + env] -= tmp[i];
    }
  }
}

template <typename T>
__device__ __forceinline__ void publish_obs(const KArgs<T>& a, int env, int lane = 0, int nlanes = 1) {
  const DModel* hd = reinterpret_cast<const DModel*>(a.model);
  const int nq = hd->nq, nv = hd->nv;
  const long long S = a.nenvp;
  const long long base = (long long)a.obs_rank * (nq + nv) * a.obs_nenv + a.env_base + env;
  for (int i = lane; i < nq + nv; i += nlanes) {
    const float v = (float)(i < nq ? a.qpos[(long long)i * S + env] : a.qvel[(long long)(i - nq) * S + env]);
    for (int p = 0; p < a.obs_world; p++) a.obs_peers[p][base + (long long)i * a.obs_nenv] = v;
  }
}
// the same as a kernel of its own, for the ticks that do not end in k_integrate (single-kernel chains)
template <typename T>
__global__ void k_publish_obs(const KArgs<T> a) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env < a.nenv) publish_obs(a, env);
}

// G7: mj_checkAcc, semi-implicit Euler with implicit damping, odom override, observation publish.  L lanes per
// environment: lane l integrates the kinematic trees of its item lists (the damped factorisation M + h D is block
// diagonal over trees like M itself); L = 1 is the thread-per-environment form.
template <typename T, int BLOCK, int L = 1>
__global__ void __launch_bounds__(BLOCK) k_integrate(const KArgs<T> a) {
  B2_KERNEL_PROLOGUE
  (void)ntiles;
  constexpr int EPB = BLOCK / L;
  const int nteams = a.ncount / EPB;
  const int envl = threadIdx.x / L, lane = threadIdx.x % L;
  const unsigned tmask = (L >= 32 ? 0xffffffffu : ((1u << L) - 1u)) << ((threadIdx.x & 31) & ~(L - 1));
  m.lane = lane; m.nlanes = L;
  for (int tile = blockIdx.x; tile < nteams; tile += gridDim.x) {
    const int env = tile * EPB + envl;
    const int nv = h.nv;
    prefetch_rows(a.qacc, nv, S, tile * EPB, EPB);
    if (a.flags & B2F_INTEGRATE) {
      prefetch_rows(a.qvel, nv, S, tile * EPB, EPB);
      prefetch_rows(a.qpos, h.nq, S, tile * EPB, EPB);
      if (h.has_damping) { prefetch_rows(a.qM, h.nM, S, tile * EPB, EPB); prefetch_rows(a.qfrc_smooth, nv, S, tile * EPB, EPB); prefetch_rows(a.qfrc_constraint, nv, S, tile * EPB, EPB); }
    }
    const bool skip = (a.flags & B2F_FUSABLE) && (a.status[env] & 8);   // integrated by the smooth kernel already
    SArr<T> qacc{a.qacc + env, S};
    bool bad = false;
    if (!skip)   // (every lane looks at the whole vector: one decision per environment; eight loads in flight per step)
      for (int i0 = 0; i0 < nv; i0 += 8) {
        T v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = i0 + q < nv ? qacc[i0 + q] : T(0);
#pragma unroll
        for (int q = 0; q < 8; q++) bad |= !(t_abs(v[q]) < T(1e10));
      }
    if (bad) {  // reset instead of integrating garbage
      if (L > 1) __syncwarp(tmask);   // every lane has read qacc before anybody clears it
      if (lane == 0) {
        for (int i = 0; i < h.nq; i++) a.qpos[i * S + env] = m.f(h.o_qpos0, i);
        for (int i = 0; i < nv; i++) { a.qvel[i * S + env] = 0; qacc[i] = 0; a.qacc_warmstart[i * S + env] = 0; a.qfrc_applied[i * S + env] = 0; }
        a.time[env] = 0;
        a.status[env] |= 4;
      }
    } else if (!skip && (a.flags & B2F_INTEGRATE)) {
      SArr<T> qpos{a.qpos + env, S}, qvel{a.qvel + env, S}, qM{a.qM + env, S}, LD{a.qLD + env, S}, dinv{a.qLDiagInv + env, S};
      SArr<T> frc{a.qfrc_smooth + env, S}, xa{a.qacc_smooth + env, S};
      if (a.flags & B2F_LD_SMEM) {
        // scratch of the damped factorisation (M + h D) in per-environment shared-memory columns instead of HBM arrays
        T* sc = reinterpret_cast<T*>(smem_raw + 16 + (size_t)nwords * 4);
        LD = SArr<T>{sc + envl, EPB};
        dinv = SArr<T>{sc + (size_t)h.nM * EPB + envl, EPB};
        xa = SArr<T>{sc + (size_t)(h.nM + nv) * EPB + envl, EPB};
      }
      // qfrc_smooth becomes the total force of the implicit-damping solve; qLD / qLDiagInv / qacc_smooth are dead after
      // the solver and serve as scratch for the damped factorisation
      if (h.has_damping)
        for (int k0 = GenericP::dof_lo(m); k0 < GenericP::dof_hi(m); k0 += 8) {   // (eight dofs' loads before the first store)
          T f[8], c[8]; int ii[8];
#pragma unroll
          for (int q = 0; q < 8; q++) { ii[q] = GenericP::dof_at(m, k0 + q < GenericP::dof_hi(m) ? k0 + q : k0); f[q] = frc[ii[q]]; c[q] = a.qfrc_constraint[ii[q] * S + env]; }
#pragma unroll
          for (int q = 0; q < 8; q++) if (k0 + q < GenericP::dof_hi(m)) frc[ii[q]] = f[q] + c[q];
        }
      euler_step<GenericP>(m, qpos, qvel, qM, qacc, frc, a.dt(), LD, dinv, xa);
      if (L > 1) __syncwarp(tmask);
      if (lane == 0) {
        a.time[env] += a.dt();
        if (a.flags & B2F_ODOM) odom_override(m, a, env);
      }
    }
    // observation exchange fused into the integrate epilogue: once the CTA's environments are integrated, its threads
    // store their new state into slice `rank` of every GPU's observation buffer — a warp writes one element of 32
    // consecutive environments, i.e. whole 128-byte lines through the NVLink peer mappings (plain posted stores).
    // No pack kernel, no collective call.
    if (a.flags & B2F_OBS) {
      __syncthreads();
      const DModel* hd = reinterpret_cast<const DModel*>(a.model);
      const int nobs = hd->nq + hd->nv, nq = hd->nq;
      const long long base = (long long)a.obs_rank * nobs * a.obs_nenv + a.env_base;
      for (int idx = threadIdx.x; idx < nobs * EPB; idx += BLOCK) {
        const int i = idx / EPB, e = tile * EPB + idx % EPB;
        if (e >= a.nenv) continue;
        const float v = (float)(i < nq ? a.qpos[(long long)i * S + e] : a.qvel[(long long)(i - nq) * S + e]);
        for (int p = 0; p < a.obs_world; p++) a.obs_peers[p][base + (long long)i * a.obs_nenv + e] = v;
      }
    }
  }
}

// on-demand expansions for the legacy dense fields: efc_J [njmax][nv] (pyramid rows rebuilt from the base directions of
// the block records) and efc_AR [njmax][njmax], one thread per env
template <typename T>
__global__ void k_expand_rows(const KArgs<T> a, int which /* 0: J, 1: B */, T* dst /* [njmax * nv][nenvp] */) {
  const DModel* h = reinterpret_cast<const DModel*>(a.model);
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= a.nenvp) return;
  const long long S = a.nenvp;
  const int ne = a.nefc[env], nv = h->nv;
  for (int r = 0; r < h->njmax; r++)
    for (int i = 0; i < nv; i++) dst[((long long)r * nv + i) * S + env] = 0;
  if (ne <= 0) return;
  const int nw = a.efc_nwords[env];
  const T* slab = a.efc_blocks + (long long)env * a.block_capw;
  for (int off = 0; off < nw;) {
    const T* rec = slab + off;
    const BlockShape bs = block_shape(rec);
    const int row0 = dec_int(rec[BH_ROW0]), oX = which ? bs.oB : bs.oJ;
    for (int rr = 0; rr < bs.nrow; rr++) {
      const int k = bs.nb > 1 ? rr / 2 + 1 : 0;
      const T sm = bs.nb > 1 ? ((rr & 1) ? -rec[bs.oMu + k - 1] : rec[bs.oMu + k - 1]) : T(0);
      for (int e = 0; e < bs.w; e++)
        dst[((long long)(row0 + rr) * nv + bs.dof(e)) * S + env] = rec[oX + e] + (bs.nb > 1 ? sm * rec[oX + k * bs.wq + e] : T(0));
    }
    off += bs.len;
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
