// k_smooth.cuh — fused smooth-dynamics kernel (group G1 of SURVEY.md section 8a', plus G7 when no constraint is
// active): forward kinematics, CoM quantities, CRBA, sparse LtDL, CoM velocities, passive forces incl. gravity
// compensation, RNE bias, the reference's PD computed-torque controller (src/mujoco_sim/mj_sim.cpp:1055-1077),
// inverse dynamics for MjHWInterface::read (src/mujoco_sim/mj_hw_interface.cpp:61), smooth acceleration and, for
// environments without active constraints, semi-implicit Euler + odom override.  One thread per environment; the
// model constants are TMA-staged into shared memory once per persistent CTA.  The same source is compiled under the
// two policies of k_policy.cuh (generic tree tables + strided workspace, or compile-time serial chain in registers).
#pragma once
#include "k_args.h"
#include "k_common.cuh"
#include "k_policy.cuh"

namespace b2 {

// Which lane of an environment's team works on a dof / body / joint (see MV::lane): kinematic trees are independent
// in every stage of the smooth dynamics (M is block diagonal over them), so the lanes never exchange data.
template <typename P, typename T>
__device__ __forceinline__ bool own_tree(const MV<T>& m, int t) {
  if constexpr (P::STATIC) return true;
  else return m.nlanes == 1 || (t < 0 ? 0 : t % m.nlanes) == m.lane;
}
template <typename P, typename T>
__device__ __forceinline__ bool own_dof(const MV<T>& m, int i) {
  if constexpr (P::STATIC) return true;
  else return m.nlanes == 1 || own_tree<P>(m, m.i(m.h->o_dof_treeid, i));
}
template <typename P, typename T>
__device__ __forceinline__ bool own_body(const MV<T>& m, int b) {
  if constexpr (P::STATIC) return true;
  else return m.nlanes == 1 || own_tree<P>(m, m.i(m.h->o_body_treeid, b));
}
// forward kinematics only: static bodies (no tree) are placed by EVERY lane, redundantly and identically, because a tree
// rooted on a static pedestal reads its parent's frame and the lanes do not synchronise inside the position stage
template <typename P, typename T>
__device__ __forceinline__ bool own_body_kin(const MV<T>& m, int b) {
  if constexpr (P::STATIC) return true;
  else {
    if (m.nlanes == 1) return true;
    const int t = m.i(m.h->o_body_treeid, b);
    return t < 0 || t % m.nlanes == m.lane;
  }
}
template <typename P, typename T>
__device__ __forceinline__ bool own_jnt(const MV<T>& m, int j) {
  if constexpr (P::STATIC) return true;
  else return m.nlanes == 1 || own_body<P>(m, P::jnt_bodyid(m, j));
}

// in-place sparse LtDL factorisation of the matrix stored in LD (tree order), dinv = 1 / D  (mj_factorM, A.4)
template <typename P, typename T, typename A1, typename A2>
__device__ __forceinline__ void ld_factor(const MV<T>& m, const A1& LD, const A2& dinv) {
  if constexpr (!P::STATIC) {
    // generic trees: the rows of M list a dof and its ancestors in order (dof_anc), so the ancestors of the a-th ancestor
    // of k are the entries a.. of row k: no dof_parentid chasing, the inner loop is a plain axpy of known length
    for (int k_ = P::dof_hi(m) - 1; k_ >= P::dof_lo(m); k_--) {
      const int k = P::dof_at(m, k_);
      const int Mkk = P::dof_Madr(m, k), cnt = P::dof_Mcnt(m, k);
      const T inv = T(1) / LD[Mkk];
      for (int a = 1; a < cnt; a++) {
        const int Mki = Mkk + a, Mij = P::dof_Madr(m, P::dof_anc(m, Mki));
        const T tmp = LD[Mki] * inv;
        for (int t = 0; t < cnt - a; t++) LD[Mij + t] -= LD[Mki + t] * tmp;
        LD[Mki] = tmp;
      }
      dinv[k] = inv;
    }
  } else {
#pragma unroll(P::UNROLL)
  for (int k = P::nv(m) - 1; k >= 0; k--) {
    const int Mkk = P::dof_Madr(m, k);
    const T inv = T(1) / LD[Mkk];
    int Mki = Mkk + 1;
#pragma unroll(P::UNROLL)
    for (int i = P::dof_parentid(m, k); i >= 0; i = P::dof_parentid(m, i), Mki++) {
      const T tmp = LD[Mki] * inv;
      int Mij = P::dof_Madr(m, i), Mkj = Mki;
#pragma unroll(P::UNROLL)
      for (int j = i; j >= 0; j = P::dof_parentid(m, j)) { LD[Mij] -= LD[Mkj] * tmp; Mij++; Mkj++; }
      LD[Mki] = tmp;
    }
    dinv[k] = inv;
  }
  }
}
// x <- M^-1 x by back / forward substitution on the factor (mj_solveM)
template <typename P, typename T, typename A1, typename A2, typename A3>
__device__ __forceinline__ void ld_solve(const MV<T>& m, const A1& LD, const A2& dinv, const A3& x) {
  const int nv = P::nv(m);
  if constexpr (!P::STATIC) {
    for (int k_ = P::dof_hi(m) - 1; k_ >= P::dof_lo(m); k_--) {
      const int i = P::dof_at(m, k_);
      const T xi = x[i];
      if (xi == 0) continue;
      const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
      for (int a = 1; a < cnt; a++) x[P::dof_anc(m, adr + a)] -= LD[adr + a] * xi;
    }
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); x[i] *= dinv[i]; }
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
      const int i = P::dof_at(m, k_);
      const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
      T xi = x[i];
      for (int a = 1; a < cnt; a++) xi -= LD[adr + a] * x[P::dof_anc(m, adr + a)];
      x[i] = xi;
    }
  } else {
#pragma unroll(P::UNROLL)
  for (int i = nv - 1; i >= 0; i--) {
    const T xi = x[i];
    int adr = P::dof_Madr(m, i) + 1;
#pragma unroll(P::UNROLL)
    for (int j = P::dof_parentid(m, i); j >= 0; j = P::dof_parentid(m, j)) x[j] -= LD[adr++] * xi;
  }
#pragma unroll(P::UNROLL)
  for (int i = 0; i < nv; i++) x[i] *= dinv[i];
#pragma unroll(P::UNROLL)
  for (int i = 0; i < nv; i++) {
    int adr = P::dof_Madr(m, i) + 1;
    T xi = x[i];
#pragma unroll(P::UNROLL)
    for (int j = P::dof_parentid(m, i); j >= 0; j = P::dof_parentid(m, j)) xi -= LD[adr++] * x[j];
    x[i] = xi;
  }
  }
}

// semi-implicit Euler with implicit joint damping (A.9): solves (M + h D) a = frc when any dof is damped, else uses
// qacc_in; then qvel += h a, qpos (+)= h qvel (quaternion-aware).  LDtmp / dinvtmp / xa are scratch.
template <typename P, typename T, typename AQ, typename AV, typename AM, typename AA, typename AF, typename AL, typename AD, typename AX>
__device__ __forceinline__ void euler_step(const MV<T>& m, const AQ& qpos, const AV& qvel, const AM& qM, const AA& qacc_in,
                                           const AF& frc, T hs, const AL& LDtmp, const AD& dinvtmp, const AX& xa) {
  const DModel& h = *m.h;
  const int nv = P::nv(m);
  const bool damp = h.has_damping && !(h.disableflags & DSBL_EULERDAMP);
  if constexpr (!P::STATIC) {
    // generic trees: the arrays live in HBM / L2 (or shared-memory scratch), every loop fetches a group of elements before
    // it stores the first result — a store orders the later loads of possibly aliasing arrays behind it, one L2 round
    // trip per element otherwise
    if (damp) {
      for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {   // this lane's rows of M
        const int i = P::dof_at(m, k_);
        const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
        const T dmp = hs * m.f(h.o_dof_damping, i);
        for (int a0 = 0; a0 < cnt; a0 += 8) {
          T v[8];
#pragma unroll
          for (int q = 0; q < 8; q++) v[q] = a0 + q < cnt ? qM[adr + a0 + q] : T(0);
          if (a0 == 0) v[0] += dmp;
#pragma unroll
          for (int q = 0; q < 8; q++) if (a0 + q < cnt) LDtmp[adr + a0 + q] = v[q];
        }
      }
      ld_factor<P>(m, LDtmp, dinvtmp);
    }
    for (int k0 = P::dof_lo(m); k0 < P::dof_hi(m); k0 += 8) {
      T v[8]; int ii[8];
#pragma unroll
      for (int q = 0; q < 8; q++) { ii[q] = P::dof_at(m, k0 + q < P::dof_hi(m) ? k0 + q : k0); v[q] = damp ? frc[ii[q]] : qacc_in[ii[q]]; }
#pragma unroll
      for (int q = 0; q < 8; q++) if (k0 + q < P::dof_hi(m)) xa[ii[q]] = v[q];
    }
    if (damp) ld_solve<P>(m, LDtmp, dinvtmp, xa);
    // velocity and position of a joint together: its dofs' qvel, acceleration and its qpos are fetched at once
    for (int k_ = P::jnt_lo(m); k_ < P::jnt_hi(m); k_++) {
      const int j = P::jnt_at(m, k_);
      const int qa = P::jnt_qposadr(m, j), da = P::jnt_dofadr(m, j), jt = P::jnt_type(m, j);
      if (jt == JNT_FREE) {
        T q[7], v[6], ac[6];
        ld<T, 7>(q, qpos, qa); ld<T, 6>(v, qvel, da); ld<T, 6>(ac, xa, da);
        for (int k = 0; k < 6; k++) v[k] += hs * ac[k];
        for (int k = 0; k < 3; k++) q[k] += hs * v[k];
        quat_integrate(q + 3, v + 3, hs);
        st<T, 6>(qvel, da, v); st<T, 7>(qpos, qa, q);
      } else if (jt == JNT_BALL) {
        T q[4], v[3], ac[3];
        ld<T, 4>(q, qpos, qa); ld<T, 3>(v, qvel, da); ld<T, 3>(ac, xa, da);
        for (int k = 0; k < 3; k++) v[k] += hs * ac[k];
        quat_integrate(q, v, hs);
        st<T, 3>(qvel, da, v); st<T, 4>(qpos, qa, q);
      } else {
        T q = qpos[qa], v = qvel[da];
        const T ac = xa[da];
        v += hs * ac;
        q += hs * v;
        qvel[da] = v; qpos[qa] = q;
      }
    }
    return;
  }
  if (damp) {
#pragma unroll(P::UNROLL)
    for (int i = 0; i < P::nM(m); i++) {
      if constexpr (P::STATIC) LDtmp[i] = qM[i];
      else if (m.nlanes == 1) LDtmp[i] = qM[i];
    }
    if constexpr (!P::STATIC) {
      if (m.nlanes > 1)
        for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {   // this lane's rows of M only
          const int i = P::dof_at(m, k_);
          const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
          for (int a = 0; a < cnt; a++) LDtmp[adr + a] = qM[adr + a];
        }
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); LDtmp[P::dof_Madr(m, i)] += hs * m.f(h.o_dof_damping, i); }
    ld_factor<P>(m, LDtmp, dinvtmp);
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); xa[i] = frc[i]; }
    ld_solve<P>(m, LDtmp, dinvtmp, xa);
  } else {
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); xa[i] = qacc_in[i]; }
  }
#pragma unroll(P::UNROLL)
  for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); qvel[i] += hs * xa[i]; }
#pragma unroll(P::UNROLL)
  for (int k_ = P::jnt_lo(m); k_ < P::jnt_hi(m); k_++) {
    const int j = P::jnt_at(m, k_);
    const int qa = P::jnt_qposadr(m, j), da = P::jnt_dofadr(m, j), jt = P::jnt_type(m, j);
    if (!P::STATIC && (jt == JNT_FREE || jt == JNT_BALL)) {
      int qo = qa, dofo = da;
      if (jt == JNT_FREE) {
        for (int k = 0; k < 3; k++) qpos[qa + k] += hs * qvel[da + k];
        qo += 3; dofo += 3;
      }
      T q[4], w[3];
      ld<T, 4>(q, qpos, qo);
      ld<T, 3>(w, qvel, dofo);
      quat_integrate(q, w, hs);
      st<T, 4>(qpos, qo, q);
    } else {
      qpos[qa] += hs * qvel[da];
    }
  }
}

template <typename T, typename P>
struct Smooth {
  static constexpr int nbody = P::NBODY, njnt = P::NJNT, nv = P::NV, nM = P::NM, nq = P::NQ;
  MV<T> m;
  const DModel& h;
  const KArgs<T>& a;
  // state: views of the HBM arrays (GenericP) or thread-private copies loaded once (ChainP)
  typename P::template Arr<T, nq> qpos;
  typename P::template Arr<T, nv> qvel, qacc, qfrc_applied, qfrc_bias;
  bool aliased = false;  // the exported stage arrays are the working arrays (GenericP with the workspace in HBM)
#define X(name, count) typename P::template Arr<T, (count)> name;
  B2_WS_ARRAYS(X)
#undef X

  __device__ __forceinline__ Smooth(const MV<T>& mv, const KArgs<T>& args, T* wsbase, long long wsstride, int env)
      : m(mv), h(*mv.h), a(args) {
    const long long s = args.nenvp;
    if constexpr (!P::STATIC) {
      qpos = SArr<T>{args.qpos + env, s};
      qvel = SArr<T>{args.qvel + env, s};
      qacc = SArr<T>{args.qacc + env, s};
      qfrc_applied = SArr<T>{args.qfrc_applied + env, s};
      qfrc_bias = SArr<T>{args.qfrc_bias + env, s};
#define X(name, count) name = SArr<T>{wsbase + (long long)h.w_##name * wsstride, wsstride};
      B2_WS_ARRAYS(X)
#undef X
      // workspace in HBM: the stage results the constraint pipeline reads are computed in place in their exported arrays
      // (same [element][env] layout) instead of in scratch slots that are copied out at the end — the copy was half of the
      // kernel's DRAM traffic and a third of the scratch footprint that has to stay L2-resident
      // (only when the constraint pipeline always follows: a kernel that integrates by itself reuses qLD as scratch)
      aliased = (args.flags & B2F_WS_GLOBAL) && !(args.flags & (B2F_FUSED | B2F_FUSABLE)) && args.xmat != nullptr;
      if (aliased) {
        xpos = SArr<T>{args.xpos + env, s}; xquat = SArr<T>{args.xquat + env, s}; xmat = SArr<T>{args.xmat + env, s};
        subtree_com = SArr<T>{args.subtree_com + env, s}; cdof = SArr<T>{args.cdof + env, s};
        qM = SArr<T>{args.qM + env, s}; qLD = SArr<T>{args.qLD + env, s}; qLDiagInv = SArr<T>{args.qLDiagInv + env, s};
        qfrc_passive = SArr<T>{args.qfrc_passive + env, s}; qfrc_smooth = SArr<T>{args.qfrc_smooth + env, s};
        qacc_smooth = SArr<T>{args.qacc_smooth + env, s};
      }
    } else {
#pragma unroll
      for (int i = 0; i < nq; i++) qpos[i] = args.qpos[i * s + env];
#pragma unroll
      for (int i = 0; i < nv; i++) {
        qvel[i] = args.qvel[i * s + env];
        qacc[i] = args.qacc[i * s + env];
        qfrc_applied[i] = args.qfrc_applied[i * s + env];
      }
    }
  }

  // ---- forward kinematics ----
  __device__ __forceinline__ void kinematics(int env) {
    const int nb = P::nbody(m);
    {
      const T z3[3] = {0, 0, 0}, q1[4] = {1, 0, 0, 0}, I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      st<T, 3>(xpos, 0, z3); st<T, 4>(xquat, 0, q1); st<T, 9>(xmat, 0, I); st<T, 3>(xipos, 0, z3); st<T, 9>(ximat, 0, I);
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_lo(m); k_ < P::body_hi(m); k_++) {
      const int b = P::body_at(m, k_);
      const int p = P::body_parentid(m, b), jn = P::body_jntnum(m, b), ja = P::body_jntadr(m, b);
      T pos[3], quat[4];
      if (!P::STATIC && jn == 1 && P::jnt_type(m, ja) == JNT_FREE) {
        const int qa = P::jnt_qposadr(m, ja);
        ld<T, 3>(pos, qpos, qa);
        ld<T, 4>(quat, qpos, qa + 3);
        normalize4(quat);
        st<T, 3>(xanchor, 3 * ja, pos);
        T ax[3];
        ldm<T, 3>(ax, m, h.o_jnt_axis, 3 * ja);
        st<T, 3>(xaxis, 3 * ja, ax);
      } else {
        T bp[3], bq[4], pm[9], pp[3], pq[4], r[3];
        const int mid = P::body_mocapid(m, b);
        if (mid >= 0) {
          SArr<T> mp{a.mocap_pos + env, a.nenvp}, mq{a.mocap_quat + env, a.nenvp};
          ld<T, 3>(bp, mp, 3 * mid);
          ld<T, 4>(bq, mq, 4 * mid);
          normalize4(bq);
        } else {
          ldm<T, 3>(bp, m, h.o_body_pos, 3 * b);
          ldm<T, 4>(bq, m, h.o_body_quat, 4 * b);
        }
        ld<T, 9>(pm, xmat, 9 * p);
        ld<T, 3>(pp, xpos, 3 * p);
        ld<T, 4>(pq, xquat, 4 * p);
        mat_vec3(r, pm, bp);
        pos[0] = pp[0] + r[0]; pos[1] = pp[1] + r[1]; pos[2] = pp[2] + r[2];
        mul_quat(quat, pq, bq);
#pragma unroll(P::UNROLL)
        for (int j = ja; j < ja + jn; j++) {
          const int qa = P::jnt_qposadr(m, j), jt = P::jnt_type(m, j);
          T jp[3], jx[3], anchor[3], axis[3];
          ldm<T, 3>(jp, m, h.o_jnt_pos, 3 * j);
          ldm<T, 3>(jx, m, h.o_jnt_axis, 3 * j);
          rot_vec_quat(anchor, jp, quat);
          anchor[0] += pos[0]; anchor[1] += pos[1]; anchor[2] += pos[2];
          rot_vec_quat(axis, jx, quat);
          st<T, 3>(xanchor, 3 * j, anchor);
          st<T, 3>(xaxis, 3 * j, axis);
          if (jt == JNT_SLIDE) {
            const T dq = qpos[qa] - m.f(h.o_qpos0, qa);
            pos[0] += axis[0] * dq; pos[1] += axis[1] * dq; pos[2] += axis[2] * dq;
          } else {
            T ql[4], qn[4], off[3];
            if (!P::STATIC && jt == JNT_BALL) { ld<T, 4>(ql, qpos, qa); normalize4(ql); }
            else axis_angle2quat(ql, jx, qpos[qa] - m.f(h.o_qpos0, qa));
            mul_quat(qn, quat, ql);
            quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];
            rot_vec_quat(off, jp, quat);
            pos[0] = anchor[0] - off[0]; pos[1] = anchor[1] - off[1]; pos[2] = anchor[2] - off[2];
          }
        }
      }
      normalize4(quat);
      T mat[9], ip[3], iq[4], r[3], qi[4], imat[9];
      quat2mat(mat, quat);
      st<T, 3>(xpos, 3 * b, pos);
      st<T, 4>(xquat, 4 * b, quat);
      st<T, 9>(xmat, 9 * b, mat);
      ldm<T, 3>(ip, m, h.o_body_ipos, 3 * b);
      ldm<T, 4>(iq, m, h.o_body_iquat, 4 * b);
      mat_vec3(r, mat, ip);
      r[0] += pos[0]; r[1] += pos[1]; r[2] += pos[2];
      st<T, 3>(xipos, 3 * b, r);
      if constexpr (P::STATIC) {   // register-resident chains keep the inertial frames; generic trees recompute them in com_pos
        mul_quat(qi, quat, iq);
        quat2mat(imat, qi);
        st<T, 9>(ximat, 9 * b, imat);
      }
    }
  }

  // geom frames go straight to HBM (consumed by the collision kernel and the ROS marker publishers); the body frames
  // are re-read from their exported HBM copies because the body index of a geom is a run-time value
  __device__ void geoms(int env) {
    const long long S = a.nenvp;
    for (int k_ = P::geom_lo(m); k_ < P::geom_hi(m); k_++) {
      const int g = P::geom_at(m, k_);
      const int b = m.i(h.o_geom_bodyid, g);
      T gp[3], gq[4], bm[9], bp[3], bq[4], r[3], q[4], gm[9];
      ldm<T, 3>(gp, m, h.o_geom_pos, 3 * g);
      ldm<T, 4>(gq, m, h.o_geom_quat, 4 * g);
      for (int k = 0; k < 9; k++) bm[k] = a.xmat[(9 * b + k) * S + env];
      for (int k = 0; k < 3; k++) bp[k] = a.xpos[(3 * b + k) * S + env];
      for (int k = 0; k < 4; k++) bq[k] = a.xquat[(4 * b + k) * S + env];
      mat_vec3(r, bm, gp);
      for (int k = 0; k < 3; k++) a.geom_xpos[(3 * g + k) * S + env] = bp[k] + r[k];
      mul_quat(q, bq, gq);
      quat2mat(gm, q);
      for (int k = 0; k < 9; k++) a.geom_xmat[(9 * g + k) * S + env] = gm[k];
    }
  }

  // ---- CoM-based quantities (A.3) ----
  __device__ __forceinline__ void com_pos() {
    const int nb = P::nbody(m);
    const bool team = !P::STATIC && m.nlanes > 1;
    // (tree-parallel form: the world body's total is not needed by any stage and would be a cross-lane sum; the kernel
    //  fills it in afterwards, for the exported array)
    if (!team) { const T mass = m.f(h.o_body_mass, 0); for (int k = 0; k < 3; k++) subtree_com[k] = mass * xipos[k]; }
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_lo(m); k_ < P::body_hi(m); k_++) {
      const int b = P::body_at(m, k_);
      const T mass = m.f(h.o_body_mass, b);
      for (int k = 0; k < 3; k++) subtree_com[3 * b + k] = mass * xipos[3 * b + k];
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_hi(m) - 1; k_ >= P::body_lo(m); k_--) {
      const int b = P::body_at(m, k_);
      const int p = P::body_parentid(m, b);
      if (team && p == 0) continue;
      for (int k = 0; k < 3; k++) subtree_com[3 * p + k] += subtree_com[3 * b + k];
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_lo(m) - (team ? 0 : 1); k_ < P::body_hi(m); k_++) {
      const int b = k_ < P::body_lo(m) ? 0 : P::body_at(m, k_);
      const T sm = m.f(h.o_body_subtreemass, b);
      if (sm < Eps<T>::minval()) { for (int k = 0; k < 3; k++) subtree_com[3 * b + k] = xipos[3 * b + k]; }
      else { const T inv = T(1) / sm; for (int k = 0; k < 3; k++) subtree_com[3 * b + k] *= inv; }
    }
    for (int k = 0; k < 10; k++) cinert[k] = 0;
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_lo(m); k_ < P::body_hi(m); k_++) {
      const int b = P::body_at(m, k_);
      const int root = P::body_rootid(m, b);
      T dif[3], mat[9], inert[3], tmp[6];
      for (int k = 0; k < 3; k++) dif[k] = xipos[3 * b + k] - subtree_com[3 * root + k];
      if constexpr (P::STATIC) {
        ld<T, 9>(mat, ximat, 9 * b);
      } else {
        // inertial frame from the body quaternion (4 loads + ~40 flops) instead of a stored 3 x 3 (9 stores + 9 loads
        // through the HBM / L2-resident workspace): the same operations kinematics() used to do, so the values are identical
        T bq[4], iq[4], qi[4];
        ld<T, 4>(bq, xquat, 4 * b);
        ldm<T, 4>(iq, m, h.o_body_iquat, 4 * b);
        mul_quat(qi, bq, iq);
        quat2mat(mat, qi);
      }
      ldm<T, 3>(inert, m, h.o_body_inertia, 3 * b);
      const T mass = m.f(h.o_body_mass, b);
      // R diag(I) R^T, upper triangle: xx yy zz xy xz yz
      tmp[0] = mat[0] * inert[0] * mat[0] + mat[1] * inert[1] * mat[1] + mat[2] * inert[2] * mat[2];
      tmp[1] = mat[3] * inert[0] * mat[3] + mat[4] * inert[1] * mat[4] + mat[5] * inert[2] * mat[5];
      tmp[2] = mat[6] * inert[0] * mat[6] + mat[7] * inert[1] * mat[7] + mat[8] * inert[2] * mat[8];
      tmp[3] = mat[0] * inert[0] * mat[3] + mat[1] * inert[1] * mat[4] + mat[2] * inert[2] * mat[5];
      tmp[4] = mat[0] * inert[0] * mat[6] + mat[1] * inert[1] * mat[7] + mat[2] * inert[2] * mat[8];
      tmp[5] = mat[3] * inert[0] * mat[6] + mat[4] * inert[1] * mat[7] + mat[5] * inert[2] * mat[8];
      T ci[10];
      ci[0] = tmp[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
      ci[1] = tmp[1] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
      ci[2] = tmp[2] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
      ci[3] = tmp[3] - mass * dif[0] * dif[1];
      ci[4] = tmp[4] - mass * dif[0] * dif[2];
      ci[5] = tmp[5] - mass * dif[1] * dif[2];
      ci[6] = mass * dif[0]; ci[7] = mass * dif[1]; ci[8] = mass * dif[2];
      ci[9] = mass;
      st<T, 10>(cinert, 10 * b, ci);
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::jnt_lo(m); k_ < P::jnt_hi(m); k_++) {
      const int j = P::jnt_at(m, k_);
      const int bi = P::jnt_bodyid(m, j), da = P::jnt_dofadr(m, j), jt = P::jnt_type(m, j);
      const int root = P::body_rootid(m, bi);
      T off[3];
      for (int k = 0; k < 3; k++) off[k] = subtree_com[3 * root + k] - xanchor[3 * j + k];
      if (!P::STATIC && (jt == JNT_FREE || jt == JNT_BALL)) {
        int skip = 0;
        if (jt == JNT_FREE) {
          for (int k = 0; k < 3; k++)
            for (int r = 0; r < 6; r++) cdof[6 * (da + k) + r] = (r == 3 + k) ? T(1) : T(0);
          skip = 3;
        }
        for (int k = 0; k < 3; k++) {
          T ax[3] = {xmat[9 * bi + k], xmat[9 * bi + 3 + k], xmat[9 * bi + 6 + k]}, lin[3];
          cross3(lin, ax, off);
          const int d = da + skip + k;
          cdof[6 * d] = ax[0]; cdof[6 * d + 1] = ax[1]; cdof[6 * d + 2] = ax[2];
          cdof[6 * d + 3] = lin[0]; cdof[6 * d + 4] = lin[1]; cdof[6 * d + 5] = lin[2];
        }
      } else if (jt == JNT_SLIDE) {
        cdof[6 * da] = 0; cdof[6 * da + 1] = 0; cdof[6 * da + 2] = 0;
        for (int k = 0; k < 3; k++) cdof[6 * da + 3 + k] = xaxis[3 * j + k];
      } else {
        T ax[3], lin[3];
        ld<T, 3>(ax, xaxis, 3 * j);
        cross3(lin, ax, off);
        cdof[6 * da] = ax[0]; cdof[6 * da + 1] = ax[1]; cdof[6 * da + 2] = ax[2];
        cdof[6 * da + 3] = lin[0]; cdof[6 * da + 4] = lin[1]; cdof[6 * da + 5] = lin[2];
      }
    }
  }

  // ---- composite rigid body algorithm (A.4) ----
  __device__ __forceinline__ void crb_mass() {
    const int nb = P::nbody(m), nvv = P::nv(m);
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_lo(m) - 1; k_ < P::body_hi(m); k_++) {
      const int b = k_ < P::body_lo(m) ? 0 : P::body_at(m, k_);
      T ci[10];
      ld<T, 10>(ci, cinert, 10 * b);   // (all ten loads before the first store: the arrays may alias for the compiler)
      st<T, 10>(crb, 10 * b, ci);
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_hi(m) - 1; k_ >= P::body_lo(m); k_--) {
      const int b = P::body_at(m, k_);
      const int p = P::body_parentid(m, b);
      if (p > 0) {
        T cp[10], cb[10];
        ld<T, 10>(cp, crb, 10 * p);
        ld<T, 10>(cb, crb, 10 * b);
        for (int k = 0; k < 10; k++) cp[k] += cb[k];
        st<T, 10>(crb, 10 * p, cp);
      }
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
      const int i = P::dof_at(m, k_);
      int adr = P::dof_Madr(m, i);
      T ci[10], cd[6], buf[6];
      ld<T, 10>(ci, crb, 10 * P::dof_bodyid(m, i));
      ld<T, 6>(cd, cdof, 6 * i);
      mul_inert_vec(buf, ci, cd);
      T v = m.f(h.o_dof_armature, i);
      if constexpr (!P::STATIC) {
        const int cnt = P::dof_Mcnt(m, i);
        // row i of M over the flattened ancestor list (no dof_parentid chasing); the axes of four ancestors are fetched
        // before the first entry is stored
        for (int a0 = 0; a0 < cnt; a0 += 4) {
          T cj[4][6];
#pragma unroll
          for (int q = 0; q < 4; q++) if (a0 + q < cnt) ld<T, 6>(cj[q], cdof, 6 * P::dof_anc(m, adr + a0 + q));
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (a0 + q >= cnt) break;
            v += cj[q][0] * buf[0] + cj[q][1] * buf[1] + cj[q][2] * buf[2] + cj[q][3] * buf[3] + cj[q][4] * buf[4] + cj[q][5] * buf[5];
            qM[adr + a0 + q] = v;
            v = 0;
          }
        }
      } else {
#pragma unroll(P::UNROLL)
      for (int j = i; j >= 0; j = P::dof_parentid(m, j)) {
        T cj[6];
        ld<T, 6>(cj, cdof, 6 * j);
        v += cj[0] * buf[0] + cj[1] * buf[1] + cj[2] * buf[2] + cj[3] * buf[3] + cj[4] * buf[4] + cj[5] * buf[5];
        qM[adr++] = v;
        v = 0;
      }
      }
    }
  }

  // res = M vec and res2 = M vec2 in ONE pass over the rows of M (generic trees): the controller's M ddq and mj_inverse's
  // M qacc of the reference tick.  Every row costs one L2 round trip whatever the number of right-hand sides.  Same
  // operation order per product as mul_M.
  template <typename AR, typename AV, typename AR2, typename AV2>
  __device__ __forceinline__ void mul_M2(const AR& res, const AV& vec, const AR2& res2, const AV2& vec2) {
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); res[i] = 0; res2[i] = 0; }
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
      const int i = P::dof_at(m, k_);
      const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
      T vi = 0, ri = 0, vi2 = 0, ri2 = 0;
      for (int a0 = 0; a0 < cnt; a0 += 4) {
        T mij[4], vj[4], rj[4], vj2[4], rj2[4]; int jj[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const bool in = a0 + q < cnt;
          jj[q] = in ? P::dof_anc(m, adr + a0 + q) : i;
          mij[q] = in ? qM[adr + a0 + q] : T(0);
          vj[q] = in ? vec[jj[q]] : T(0); rj[q] = in ? res[jj[q]] : T(0);
          vj2[q] = in ? vec2[jj[q]] : T(0); rj2[q] = in ? res2[jj[q]] : T(0);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (a0 + q >= cnt) break;
          if (a0 + q == 0) { vi = vj[0]; ri = rj[0] + mij[0] * vi; vi2 = vj2[0]; ri2 = rj2[0] + mij[0] * vi2; continue; }
          ri += mij[q] * vj[q]; ri2 += mij[q] * vj2[q];
          res[jj[q]] = rj[q] + mij[q] * vi;
          res2[jj[q]] = rj2[q] + mij[q] * vi2;
        }
      }
      res[i] = ri; res2[i] = ri2;
    }
  }

  // res = M vec (mj_mulM, reference call site src/mujoco_sim/mj_sim.cpp:1057)
  template <typename AR, typename AV>
  __device__ __forceinline__ void mul_M(const AR& res, const AV& vec) {
    const int nvv = P::nv(m);
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); res[i] = 0; }
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
      const int i = P::dof_at(m, k_);
      int adr = P::dof_Madr(m, i);
      if constexpr (!P::STATIC) {
        const int cnt = P::dof_Mcnt(m, i);
        // (the row's entries and the vector elements of the dof and its ancestors, four at a time, before the first
        //  update is stored; entry 0 of a row is the diagonal and its "ancestor" the dof itself)
        T vi = 0, ri = 0;
        for (int a0 = 0; a0 < cnt; a0 += 4) {
          T mij[4], vj[4], rj[4]; int jj[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const bool in = a0 + q < cnt;
            jj[q] = in ? P::dof_anc(m, adr + a0 + q) : i;
            mij[q] = in ? qM[adr + a0 + q] : T(0); vj[q] = in ? vec[jj[q]] : T(0); rj[q] = in ? res[jj[q]] : T(0);
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (a0 + q >= cnt) break;
            if (a0 + q == 0) { vi = vj[0]; ri = rj[0] + mij[0] * vi; continue; }
            ri += mij[q] * vj[q];
            res[jj[q]] = rj[q] + mij[q] * vi;   // (the ancestors of a row are distinct dofs: no entry is updated twice here)
          }
        }
        res[i] = ri;
      } else {
      const T vi = vec[i];
      T ri = res[i] + qM[adr] * vi;
      adr++;
#pragma unroll(P::UNROLL)
      for (int j = P::dof_parentid(m, i); j >= 0; j = P::dof_parentid(m, j), adr++) {
        const T mij = qM[adr];
        ri += mij * vec[j];
        res[j] += mij * vi;
      }
      res[i] = ri;
      }
    }
  }

  // ---- velocity stage (A.5) ----
  __device__ __forceinline__ void com_vel() {
    for (int k = 0; k < 6; k++) cvel[k] = 0;
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_lo(m); k_ < P::body_hi(m); k_++) {
      const int b = P::body_at(m, k_);
      T cv[6];
      ld<T, 6>(cv, cvel, 6 * P::body_parentid(m, b));
      const int ja = P::body_jntadr(m, b), jn = P::body_jntnum(m, b);
#pragma unroll(P::UNROLL)
      for (int j = ja; j < ja + jn; j++) {
        int dof = P::jnt_dofadr(m, j);
        const int jt = P::jnt_type(m, j);
        if (!P::STATIC && (jt == JNT_FREE || jt == JNT_BALL)) {
          if (jt == JNT_FREE) {
            for (int k = 0; k < 3; k++) {
              const T v = qvel[dof + k];
              for (int r = 0; r < 6; r++) { cdof_dot[6 * (dof + k) + r] = 0; cv[r] += cdof[6 * (dof + k) + r] * v; }
            }
            dof += 3;
          }
          // all three axes see the body velocity before this joint's own rotational dofs are added
          for (int k = 0; k < 3; k++) {
            T cd[6], dd[6];
            ld<T, 6>(cd, cdof, 6 * (dof + k));
            cross_motion(dd, cv, cd);
            st<T, 6>(cdof_dot, 6 * (dof + k), dd);
          }
          for (int k = 0; k < 3; k++) {
            const T v = qvel[dof + k];
            for (int r = 0; r < 6; r++) cv[r] += cdof[6 * (dof + k) + r] * v;
          }
        } else {
          T cd[6], dd[6];
          ld<T, 6>(cd, cdof, 6 * dof);
          cross_motion(dd, cv, cd);
          st<T, 6>(cdof_dot, 6 * dof, dd);
          const T v = qvel[dof];
          for (int r = 0; r < 6; r++) cv[r] += cd[r] * v;
        }
      }
      st<T, 6>(cvel, 6 * b, cv);
    }
  }

  // qfrc += J^T [force; torque] for a wrench applied at world point `point` on body b
  template <typename AQ>
  __device__ __forceinline__ void apply_ft(const AQ& qfrc, int b, const T* point, const T* force, const T* torque) {
    const int root = P::body_rootid(m, b);
    T off[3];
    for (int k = 0; k < 3; k++) off[k] = point[k] - subtree_com[3 * root + k];
#pragma unroll(P::UNROLL)
    for (int i = P::body_lastdof(m, b); i >= 0; i = P::dof_parentid(m, i)) {
      T cd[6], t[3];
      ld<T, 6>(cd, cdof, 6 * i);
      cross3(t, cd, off);
      T s = (cd[3] + t[0]) * force[0] + (cd[4] + t[1]) * force[1] + (cd[5] + t[2]) * force[2];
      if (torque) s += cd[0] * torque[0] + cd[1] * torque[1] + cd[2] * torque[2];
      qfrc[i] += s;
    }
  }

  __device__ __forceinline__ void passive() {
    const int nvv = P::nv(m);
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); qfrc_passive[i] = 0; }
    if (h.disableflags & DSBL_PASSIVE) return;
    if (h.has_stiffness) {
#pragma unroll(P::UNROLL)
      for (int k_ = P::jnt_lo(m); k_ < P::jnt_hi(m); k_++) {
        const int j = P::jnt_at(m, k_);
        const T k = m.f(h.o_jnt_stiffness, j);
        if (k == 0) continue;
        const int qa = P::jnt_qposadr(m, j), da = P::jnt_dofadr(m, j), jt = P::jnt_type(m, j);
        if (!P::STATIC && (jt == JNT_FREE || jt == JNT_BALL)) {
          int qo = qa, dofo = da;
          if (jt == JNT_FREE) {
            for (int r = 0; r < 3; r++) qfrc_passive[da + r] = -k * (qpos[qa + r] - m.f(h.o_qpos_spring, qa + r));
            qo += 3; dofo += 3;
          }
          T q[4], qs[4], dif[3];
          ld<T, 4>(q, qpos, qo);
          normalize4(q);
          ldm<T, 4>(qs, m, h.o_qpos_spring, qo);
          sub_quat(dif, q, qs);
          for (int r = 0; r < 3; r++) qfrc_passive[dofo + r] = -k * dif[r];
        } else {
          qfrc_passive[da] = -k * (qpos[qa] - m.f(h.o_qpos_spring, qa));
        }
      }
    }
    if (h.has_damping) {
#pragma unroll(P::UNROLL)
      for (int k0 = P::dof_lo(m); k0 < P::dof_hi(m); k0 += 4) {   // (four dofs' loads before the first store)
        T fp[4], vq[4]; int ii[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          ii[q] = P::dof_at(m, k0 + q < P::dof_hi(m) ? k0 + q : k0);
          fp[q] = qfrc_passive[ii[q]]; vq[q] = qvel[ii[q]];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) if (k0 + q < P::dof_hi(m)) qfrc_passive[ii[q]] = fp[q] - m.f(h.o_dof_damping, ii[q]) * vq[q];
      }
    }
    // gravity compensation: the reference sets gravcomp="1" on every robot body by default
    // (src/mujoco_sim/mj_sim.cpp:301-310, src/config/robot.yaml:19)
    if (h.has_gravcomp && !(h.disableflags & DSBL_GRAVITY)) {
#pragma unroll(P::UNROLL)
      for (int k_ = P::body_lo(m); k_ < P::body_hi(m); k_++) {
        const int b = P::body_at(m, k_);
        const T gc = m.f(h.o_body_gravcomp, b);
        if (gc == 0) continue;
        const T s = -m.f(h.o_body_mass, b) * gc;
        T f[3] = {m.f(h.o_opt_real, 0) * s, m.f(h.o_opt_real, 1) * s, m.f(h.o_opt_real, 2) * s}, pt[3];
        ld<T, 3>(pt, xipos, 3 * b);
        apply_ft(qfrc_passive, b, pt, f, (const T*)nullptr);
      }
    }
  }

  // recursive Newton-Euler in the CoM frame; with_acc adds cdof * qacc (inverse dynamics)
  template <typename AR>
  __device__ __forceinline__ void rne(const AR& result, bool with_acc) {
    const int nb = P::nbody(m), nvv = P::nv(m);
    const bool grav = !(h.disableflags & DSBL_GRAVITY);
    cacc[0] = 0; cacc[1] = 0; cacc[2] = 0;
    cacc[3] = grav ? -m.f(h.o_opt_real, 0) : T(0); cacc[4] = grav ? -m.f(h.o_opt_real, 1) : T(0); cacc[5] = grav ? -m.f(h.o_opt_real, 2) : T(0);
    for (int k = 0; k < 6; k++) cfrc[k] = 0;
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_lo(m); k_ < P::body_hi(m); k_++) {
      const int b = P::body_at(m, k_);
      T ac[6], ci[10], cv[6], Ia[6], Iv[6], x[6];
      ld<T, 6>(ac, cacc, 6 * P::body_parentid(m, b));
      const int da = P::body_dofadr(m, b), dn = P::body_dofnum(m, b);
      // (the body's inertia and velocity are fetched together with its dofs' terms, before cacc is stored)
      ld<T, 10>(ci, cinert, 10 * b);
      ld<T, 6>(cv, cvel, 6 * b);
      if constexpr (P::STATIC) {
#pragma unroll(P::UNROLL)
      for (int j = 0; j < dn; j++) {
        const T v = qvel[da + j];
        for (int r = 0; r < 6; r++) ac[r] += cdof_dot[6 * (da + j) + r] * v;
        if (with_acc) {
          const T q2 = qacc[da + j];
          for (int r = 0; r < 6; r++) ac[r] += cdof[6 * (da + j) + r] * q2;
        }
      }
      } else {
      for (int j0 = 0; j0 < dn; j0 += 3) {
        T v[3], dd[3][6], q2[3] = {0, 0, 0}, cd[3][6];
#pragma unroll
        for (int q = 0; q < 3; q++) {
          if (j0 + q < dn) {
            v[q] = qvel[da + j0 + q];
            ld<T, 6>(dd[q], cdof_dot, 6 * (da + j0 + q));
            if (with_acc) { q2[q] = qacc[da + j0 + q]; ld<T, 6>(cd[q], cdof, 6 * (da + j0 + q)); }
          }
        }
#pragma unroll
        for (int q = 0; q < 3; q++) {
          if (j0 + q >= dn) break;
          for (int r = 0; r < 6; r++) ac[r] += dd[q][r] * v[q];
          if (with_acc) for (int r = 0; r < 6; r++) ac[r] += cd[q][r] * q2[q];
        }
      }
      }
      st<T, 6>(cacc, 6 * b, ac);
      mul_inert_vec(Ia, ci, ac);
      mul_inert_vec(Iv, ci, cv);
      cross_force(x, cv, Iv);
      for (int r = 0; r < 6; r++) cfrc[6 * b + r] = Ia[r] + x[r];
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::body_hi(m) - 1; k_ >= P::body_lo(m); k_--) {
      const int b = P::body_at(m, k_);
      const int p = P::body_parentid(m, b);
      if (p > 0) {
        T fp[6], fb[6];
        ld<T, 6>(fp, cfrc, 6 * p);
        ld<T, 6>(fb, cfrc, 6 * b);
        for (int r = 0; r < 6; r++) fp[r] += fb[r];
        st<T, 6>(cfrc, 6 * p, fp);
      }
    }
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
      const int i = P::dof_at(m, k_);
      const int b = P::dof_bodyid(m, i);
      T s = 0;
      for (int r = 0; r < 6; r++) s += cdof[6 * i + r] * cfrc[6 * b + r];
      result[i] = s;
    }
  }
};

// MjSim::set_odom_vels (src/mujoco_sim/mj_sim.cpp:1079-1153): per robot r, odom_dof[6r..] = dof of lin x,y,z / ang x,y,z
// odom joints (-1 when absent or disabled), odom_qpos[3r..] = qpos address of the three angular odom joints (-1 -> 0).
template <typename T>
__device__ void odom_override(const MV<T>& m, const KArgs<T>& a, int env) {
  const DModel& h = *m.h;
  SArr<T> qpos{a.qpos + env, a.nenvp}, qvel{a.qvel + env, a.nenvp}, ov{a.odom_vels + env, a.nenvp};
  for (int r = 0; r < h.nodom; r++) {
    const int ax = m.i(h.o_odom_qpos, 3 * r), ay = m.i(h.o_odom_qpos, 3 * r + 1), az = m.i(h.o_odom_qpos, 3 * r + 2);
    const T x = ax >= 0 ? qpos[ax] : T(0), y = ay >= 0 ? qpos[ay] : T(0), z = az >= 0 ? qpos[az] : T(0);
    T sx, cx, sy, cy, sz, cz;
    t_sincos(x, &sx, &cx); t_sincos(y, &sy, &cy); t_sincos(z, &sz, &cz);
    const T vx = ov[6 * r], vy = ov[6 * r + 1], vz = ov[6 * r + 2];
    const int d0 = m.i(h.o_odom_dof, 6 * r), d1 = m.i(h.o_odom_dof, 6 * r + 1), d2 = m.i(h.o_odom_dof, 6 * r + 2);
    if (d0 >= 0) qvel[d0] = vx * cy * cz + vy * (sx * sy * cz - cx * sz) + vz * (cx * sy * cz + sx * sz);
    if (d1 >= 0) qvel[d1] = vx * cy * sz + vy * (sx * sy * sz + cx * cz) + vz * (cx * sy * sz - sx * cz);
    if (d2 >= 0) qvel[d2] = -vx * sy + vy * sx * cy + vz * cx * cy;
    for (int k = 0; k < 3; k++) {
      const int dk = m.i(h.o_odom_dof, 6 * r + 3 + k);
      if (dk >= 0) qvel[dk] = ov[6 * r + 3 + k];
    }
  }
}

// ---- the kernel: persistent CTAs; L lanes per environment (L = 1: one thread per environment).  With L > 1 lane l works
//      on the kinematic trees t with t % L == l (own_tree): the trees of an environment are independent in every stage,
//      so an arm and the free objects around it are processed side by side and the per-environment serial chain shrinks
//      to its largest tree ----
template <typename T, int BLOCK, typename P, int MINB = 1, int L = 1>
__global__ void __launch_bounds__(BLOCK, MINB) k_smooth(const KArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* blob = reinterpret_cast<uint32_t*>(smem_raw + 16);
  const int nwords = reinterpret_cast<const DModel*>(a.model)->nwords;
  stage_model(blob, a.model, nwords, bar);
  MV<T> m{reinterpret_cast<const DModel*>(blob), blob};
  const DModel& h = *m.h;
  T* ws_sh = reinterpret_cast<T*>(smem_raw + 16 + (size_t)nwords * 4);
  const int nv = P::nv(m), nb = P::nbody(m), nq = P::nq(m), nM = P::nM(m);
  const long long S = a.nenvp;
  static_assert(L == 1 || !P::STATIC, "the register-resident chain policy is one thread per environment");
  constexpr int EPB = BLOCK / L;
  const int ntiles = a.ncount / EPB;
  const int envl = threadIdx.x / L, lane = threadIdx.x % L;
  const unsigned tmask = (L >= 32 ? 0xffffffffu : ((1u << L) - 1u)) << ((threadIdx.x & 31) & ~(L - 1));
  m.lane = lane; m.nlanes = L;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * EPB + envl;
    if constexpr (!P::STATIC) {
      // the state this tile reads, in flight at once
      prefetch_rows(a.qpos, nq, S, tile * EPB, EPB);
      prefetch_rows(a.qvel, nv, S, tile * EPB, EPB);
      prefetch_rows(a.qfrc_applied, nv, S, tile * EPB, EPB);
      if (a.flags & B2F_INVERSE) prefetch_rows(a.qacc, nv, S, tile * EPB, EPB);
      if (a.flags & B2F_CONTROLLER) { prefetch_rows(a.ddq, nv, S, tile * EPB, EPB); prefetch_rows(a.dq, nv, S, tile * EPB, EPB); }
    }
    T* wsbase = (a.flags & B2F_WS_GLOBAL) ? a.ws + env : ws_sh + envl;
    const long long wss = (a.flags & B2F_WS_GLOBAL) ? S : EPB;
    Smooth<T, P> s(m, a, wsbase, wss, env);
    SArr<T> qfrc_inverse{a.qfrc_inverse + env, S};
    if constexpr (!P::STATIC) {
      // workspace in HBM with the factor scratch in shared memory: its third vector (free in this kernel) is the scratch
      // vector of the controller's and mj_inverse's M x products and, at the end, of the solve for qacc_smooth
      if (a.flags & B2F_LD_SMEM) s.tmpv = SArr<T>{ws_sh + (size_t)(nM + nv) * EPB + envl, EPB};
    }
    // the reference tick needs M ddq (controller) and M qacc (mj_inverse): one pass over M for both, the second product in
    // the scratch's fourth vector
    bool both = false;
    if constexpr (!P::STATIC) both = (a.flags & B2F_LD_SMEM) && a.ld_extra && (a.flags & B2F_CONTROLLER) && (a.flags & B2F_INVERSE);
    auto tmpv2 = s.tmpv;
    if constexpr (!P::STATIC) { if (both) tmpv2 = SArr<T>{ws_sh + (size_t)(nM + 2 * nv) * EPB + envl, EPB}; }

    // mj_checkPos / mj_checkVel: reset an environment whose state went non-finite
    {
      bool bad = false;
      if constexpr (P::STATIC) {
#pragma unroll(P::UNROLL)
        for (int i = 0; i < nq; i++) { const T v = s.qpos[i]; bad |= !(t_abs(v) < T(1e10)); }
#pragma unroll(P::UNROLL)
        for (int i = 0; i < nv; i++) { const T v = s.qvel[i]; bad |= !(t_abs(v) < T(1e10)); }
      } else {
        // (eight loads in flight per step: this is the first touch of the state, an HBM round trip per element otherwise)
        for (int i0 = 0; i0 < nq; i0 += 8) {
          T v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = i0 + k < nq ? s.qpos[i0 + k] : T(0);
#pragma unroll
          for (int k = 0; k < 8; k++) bad |= !(t_abs(v[k]) < T(1e10));
        }
        for (int i0 = 0; i0 < nv; i0 += 8) {
          T v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = i0 + k < nv ? s.qvel[i0 + k] : T(0);
#pragma unroll
          for (int k = 0; k < 8; k++) bad |= !(t_abs(v[k]) < T(1e10));
        }
      }
      if (bad) {
#pragma unroll(P::UNROLL)
        for (int i = 0; i < nq; i++) { const T q0 = m.f(h.o_qpos0, i); s.qpos[i] = q0; a.qpos[i * S + env] = q0; }
#pragma unroll(P::UNROLL)
        for (int i = 0; i < nv; i++) {
          s.qvel[i] = 0; s.qacc[i] = 0; s.qfrc_applied[i] = 0;
          a.qvel[i * S + env] = 0; a.qacc[i * S + env] = 0; a.qacc_warmstart[i * S + env] = 0; a.qfrc_applied[i * S + env] = 0;
        }
        a.time[env] = 0;
        a.status[env] |= 4;
      }
    }

    // position stage
    s.kinematics(env);
    s.com_pos();
    if (L > 1) {
      // the world body's subtree centre of mass (exported only): the one cross-tree sum, by lane 0 once every lane's
      // inertial frame positions are in the workspace
      __syncwarp(tmask);
      if (lane == 0) {
        T c[3] = {0, 0, 0};
        for (int b = 0; b < nb; b++) { const T mass = m.f(h.o_body_mass, b); for (int k2 = 0; k2 < 3; k2++) c[k2] += mass * s.xipos[3 * b + k2]; }
        const T sm = m.f(h.o_body_subtreemass, 0);
        for (int k2 = 0; k2 < 3; k2++) s.subtree_com[k2] = sm < Eps<T>::minval() ? s.xipos[k2] : c[k2] / sm;
      }
    }
    s.crb_mass();
    // L^T D L.  With the workspace in HBM the O(sum depth^2) read-modify-writes of the elimination would each be an L2
    // round trip: the factor is built in a per-thread shared-memory column instead and written out once (B2F_LD_SMEM).
    bool ldsm = false;
    if constexpr (!P::STATIC) ldsm = (a.flags & B2F_LD_SMEM) != 0;
    if constexpr (!P::STATIC) {
      if (ldsm) {
        SArr<T> LDs{ws_sh + envl, EPB}, dis{ws_sh + (size_t)nM * EPB + envl, EPB};
        for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
          const int i = P::dof_at(m, k_);
          const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
          for (int q0 = 0; q0 < cnt; q0 += 8) {   // (eight loads in flight per step)
            T v[8];
#pragma unroll
            for (int q2 = 0; q2 < 8; q2++) v[q2] = q0 + q2 < cnt ? s.qM[adr + q0 + q2] : T(0);
#pragma unroll
            for (int q2 = 0; q2 < 8; q2++) if (q0 + q2 < cnt) LDs[adr + q0 + q2] = v[q2];
          }
        }
        ld_factor<P>(m, LDs, dis);
        for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
          const int i = P::dof_at(m, k_);
          const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
          for (int q2 = 0; q2 < cnt; q2++) s.qLD[adr + q2] = LDs[adr + q2];
          s.qLDiagInv[i] = dis[i];
        }
      }
    }
    if (!ldsm) {
      if constexpr (P::STATIC) {
#pragma unroll(P::UNROLL)
        for (int i = 0; i < nM; i++) s.qLD[i] = s.qM[i];
      } else {
        for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
          const int i = P::dof_at(m, k_);
          const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
          for (int q2 = 0; q2 < cnt; q2++) s.qLD[adr + q2] = s.qM[adr + q2];
        }
      }
      ld_factor<P>(m, s.qLD, s.qLDiagInv);
    }
    // velocity stage
    s.com_vel();
    s.passive();
    s.rne(s.qfrc_bias, false);

    // mjcb_control -> MjSim::controller (src/mujoco_sim/mj_sim.cpp:1055-1077)
    bool overridden = false;
    if (a.flags & B2F_CONTROLLER) {
      typename P::template Arr<T, P::NV> ddq;
      if constexpr (!P::STATIC) ddq = SArr<T>{a.ddq + env, S};
      else {
#pragma unroll
        for (int i = 0; i < P::NV; i++) ddq[i] = a.ddq[i * S + env];
      }
      if constexpr (!P::STATIC) { if (both) s.mul_M2(s.tmpv, ddq, tmpv2, s.qacc); else s.mul_M(s.tmpv, ddq); }
      else s.mul_M(s.tmpv, ddq);
      if constexpr (P::STATIC) {
#pragma unroll(P::UNROLL)
      for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
        const int i = P::dof_at(m, k_);
        T tau = s.tmpv[i];
        if (P::dof_controlled(m, i)) tau += s.qfrc_bias[i];
        s.qfrc_applied[i] = tau;
        if (P::STATIC) a.qfrc_applied[i * S + env] = tau;
        const T v = a.dq[i * S + env];
        if (t_abs(v) > Eps<T>::minval()) { s.qvel[i] = v; overridden = true; }
        a.ddq[i * S + env] = 0;
        a.dq[i * S + env] = 0;
      }
      } else {
      for (int k0 = P::dof_lo(m); k0 < P::dof_hi(m); k0 += 4) {   // (four dofs' loads before the first store)
        T tq[4], bq[4], vq[4]; int ii[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          ii[q] = P::dof_at(m, k0 + q < P::dof_hi(m) ? k0 + q : k0);
          tq[q] = s.tmpv[ii[q]]; bq[q] = s.qfrc_bias[ii[q]]; vq[q] = a.dq[ii[q] * S + env];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (k0 + q >= P::dof_hi(m)) break;
          const int i = ii[q];
          T tau = tq[q];
          if (P::dof_controlled(m, i)) tau += bq[q];
          s.qfrc_applied[i] = tau;
          if (t_abs(vq[q]) > Eps<T>::minval()) { s.qvel[i] = vq[q]; overridden = true; }
          a.ddq[i * S + env] = 0;
          a.dq[i * S + env] = 0;
        }
      }
      }
    }
    // MjHWInterface::read -> mj_inverse: velocity stage again for the overridden qvel, then RNE with the stored qacc
    if (a.flags & B2F_INVERSE) {
      if (overridden) { s.com_vel(); s.passive(); s.rne(s.qfrc_bias, false); }
      // RNE is affine in the acceleration: RNE(q, v, a) + armature a = M a + bias, so the second tree pass of
      // mj_inverse collapses to one sparse mat-vec with the CRBA matrix already at hand
      if (!both) s.mul_M(s.tmpv, s.qacc);
      if constexpr (P::STATIC) {
#pragma unroll(P::UNROLL)
      for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); qfrc_inverse[i] = s.tmpv[i] + s.qfrc_bias[i] - s.qfrc_passive[i]; }
      } else {
      for (int k0 = P::dof_lo(m); k0 < P::dof_hi(m); k0 += 4) {
        T tq[4], bq[4], pq[4]; int ii[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          ii[q] = P::dof_at(m, k0 + q < P::dof_hi(m) ? k0 + q : k0);
          tq[q] = tmpv2[ii[q]]; bq[q] = s.qfrc_bias[ii[q]]; pq[q] = s.qfrc_passive[ii[q]];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) if (k0 + q < P::dof_hi(m)) qfrc_inverse[ii[q]] = tq[q] + bq[q] - pq[q];
      }
      }
    }
    if (P::STATIC) {
#pragma unroll(P::UNROLL)
      for (int i = 0; i < nv; i++) a.qfrc_bias[i * S + env] = s.qfrc_bias[i];
    }

    // smooth acceleration
    bool rhs_ready = false;   // the solve's right-hand side is already in the shared-memory vector
    if constexpr (P::STATIC) {
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); s.qfrc_smooth[i] = s.qfrc_passive[i] - s.qfrc_bias[i] + s.qfrc_applied[i]; }
    } else {
      rhs_ready = ldsm && !(a.flags & B2F_XFRC);
      for (int k0 = P::dof_lo(m); k0 < P::dof_hi(m); k0 += 4) {   // (four dofs' loads before the first store)
        T fp[4], fb[4], fa[4]; int ii[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          ii[q] = P::dof_at(m, k0 + q < P::dof_hi(m) ? k0 + q : k0);
          fp[q] = s.qfrc_passive[ii[q]]; fb[q] = s.qfrc_bias[ii[q]]; fa[q] = s.qfrc_applied[ii[q]];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (k0 + q >= P::dof_hi(m)) break;
          const T v = fp[q] - fb[q] + fa[q];
          s.qfrc_smooth[ii[q]] = v;
          if (rhs_ready) s.tmpv[ii[q]] = v;
        }
      }
    }
    if (a.flags & B2F_XFRC) {
      SArr<T> xf{a.xfrc_applied + env, S};
#pragma unroll(P::UNROLL)
      for (int k_ = P::body_lo(m); k_ < P::body_hi(m); k_++) {
        const int b = P::body_at(m, k_);
        T w[6], pt[3];
        ld<T, 6>(w, xf, 6 * b);
        if (w[0] == 0 && w[1] == 0 && w[2] == 0 && w[3] == 0 && w[4] == 0 && w[5] == 0) continue;
        ld<T, 3>(pt, s.xipos, 3 * b);
        s.apply_ft(s.qfrc_smooth, b, pt, w, w + 3);
      }
    }
    if (!ldsm) {
#pragma unroll(P::UNROLL)
    for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); s.qacc_smooth[i] = s.qfrc_smooth[i]; }
    }
    if constexpr (!P::STATIC) {
      if (ldsm) {   // the factor is still in the shared-memory column; the right-hand side joins it (tmpv)
        SArr<T> LDs{ws_sh + envl, EPB}, dis{ws_sh + (size_t)nM * EPB + envl, EPB};
        if (!rhs_ready) for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); s.tmpv[i] = s.qfrc_smooth[i]; }
        ld_solve<P>(m, LDs, dis, s.tmpv);
        for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) { const int i = P::dof_at(m, k_); s.qacc_smooth[i] = s.tmpv[i]; }
      }
    }
    if (!ldsm) ld_solve<P>(m, s.qLD, s.qLDiagInv, s.qacc_smooth);

    // body poses for the ROS layer (tf / marker publishers read d->xpos, d->xquat: SURVEY.md Appendix C)
    if (!s.aliased) {
#pragma unroll(P::UNROLL)
      for (int k_ = P::body_lo(m) - 1; k_ < P::body_hi(m); k_++) {
        const int b = k_ < P::body_lo(m) ? 0 : P::body_at(m, k_);   // (the world body: written by every lane, identically)
        for (int k = 0; k < 3; k++) a.xpos[(long long)(3 * b + k) * S + env] = s.xpos[3 * b + k];
        for (int k = 0; k < 4; k++) a.xquat[(long long)(4 * b + k) * S + env] = s.xquat[4 * b + k];
      }
    }

    // does this environment need the constraint pipeline this tick?
    bool pipeline = !(a.flags & B2F_FUSED);
    if (a.flags & B2F_FUSABLE) {
      // joint limits are the model's only constraint source: the pipeline runs only for environments with an active one
      pipeline = false;
      if (!(h.disableflags & (DSBL_LIMIT | DSBL_CONSTRAINT))) {
#pragma unroll(P::UNROLL)
        for (int j = 0; j < P::njnt(m); j++) {
          if (!P::jnt_limited(m, j)) continue;
          const T q = s.qpos[P::jnt_qposadr(m, j)], mg = m.f(h.o_jnt_margin, j);
          pipeline |= (q - m.f(h.o_jnt_range, 2 * j) < mg) || (m.f(h.o_jnt_range, 2 * j + 1) - q < mg);
        }
      }
      if (lane == 0) {
        a.nefc[env] = 0;
        a.status[env] = (a.status[env] & 7) | (pipeline ? 0 : 8);
      }
      const unsigned vote = __ballot_sync(__activemask(), pipeline && lane == 0);
      if (vote && (int)(threadIdx.x & 31) == __ffs(vote) - 1) atomicAdd(&a.pending[0], __popc(vote));
    }

    if (pipeline || (a.flags & B2F_EXPORT)) {
      // export the stage results the constraint pipeline (and the legacy mjData mirror) consume
      if (!s.aliased) {
#pragma unroll(P::UNROLL)
        for (int k_ = P::body_lo(m) - 1; k_ < P::body_hi(m); k_++) {
          const int b = k_ < P::body_lo(m) ? 0 : P::body_at(m, k_);
          if (b == 0 && L > 1 && lane != 0) continue;   // the world's subtree_com was summed by lane 0
          for (int k = 0; k < 9; k++) a.xmat[(long long)(9 * b + k) * S + env] = s.xmat[9 * b + k];
          for (int k = 0; k < 3; k++) a.subtree_com[(long long)(3 * b + k) * S + env] = s.subtree_com[3 * b + k];
        }
#pragma unroll(P::UNROLL)
        for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
          const int i = P::dof_at(m, k_);
          for (int k = 0; k < 6; k++) a.cdof[(long long)(6 * i + k) * S + env] = s.cdof[6 * i + k];
        }
        if constexpr (P::STATIC) {
#pragma unroll(P::UNROLL)
          for (int i = 0; i < nM; i++) { a.qM[i * S + env] = s.qM[i]; a.qLD[i * S + env] = s.qLD[i]; }
        } else {
          for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
            const int i = P::dof_at(m, k_);
            const int adr = P::dof_Madr(m, i), cnt = P::dof_Mcnt(m, i);
            for (int q2 = 0; q2 < cnt; q2++) { a.qM[(long long)(adr + q2) * S + env] = s.qM[adr + q2]; a.qLD[(long long)(adr + q2) * S + env] = s.qLD[adr + q2]; }
          }
        }
#pragma unroll(P::UNROLL)
        for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
          const int i = P::dof_at(m, k_);
          a.qLDiagInv[i * S + env] = s.qLDiagInv[i];
          a.qfrc_passive[i * S + env] = s.qfrc_passive[i];
          a.qfrc_smooth[i * S + env] = s.qfrc_smooth[i];
          a.qacc_smooth[i * S + env] = s.qacc_smooth[i];
        }
      }
      s.geoms(env);
      if (P::STATIC && overridden) {
#pragma unroll(P::UNROLL)
        for (int i = 0; i < nv; i++) a.qvel[i * S + env] = s.qvel[i];
      }
    }
    if (!pipeline && (a.flags & B2F_NOSOLVE)) {
      // mj_step1 leaves qacc alone (MjHWInterface::read needs the previous tick's); only a velocity override is kept
      if (P::STATIC && overridden) {
#pragma unroll(P::UNROLL)
        for (int i = 0; i < nv; i++) a.qvel[i * S + env] = s.qvel[i];
      }
    } else if (!pipeline) {
      // no active constraint: qacc = qacc_smooth, integrate right here
#pragma unroll(P::UNROLL)
      for (int k_ = P::dof_lo(m); k_ < P::dof_hi(m); k_++) {
        const int i = P::dof_at(m, k_);
        const T v = s.qacc_smooth[i];
        a.qacc[i * S + env] = v;
        a.qacc_warmstart[i * S + env] = v;
      }
      if (a.flags & B2F_INTEGRATE) {
        // qLD / qLDiagInv are dead by now: reuse them as scratch for the damped factorisation
        euler_step<P>(m, s.qpos, s.qvel, s.qM, s.qacc_smooth, s.qfrc_smooth, a.dt(), s.qLD, s.qLDiagInv, s.tmpv);
        if (L > 1) __syncwarp(tmask);   // the odom override below reads joints of other lanes' trees
        if (P::STATIC) {
#pragma unroll(P::UNROLL)
          for (int i = 0; i < nq; i++) a.qpos[i * S + env] = s.qpos[i];
#pragma unroll(P::UNROLL)
          for (int i = 0; i < nv; i++) a.qvel[i * S + env] = s.qvel[i];
        }
        if (lane == 0) {
          a.time[env] += a.dt();
          if (a.flags & B2F_ODOM) odom_override(m, a, env);
        }
      } else if (P::STATIC && overridden) {
#pragma unroll(P::UNROLL)
        for (int i = 0; i < nv; i++) a.qvel[i * S + env] = s.qvel[i];
      }
    }
  }
}

}  // namespace b2
