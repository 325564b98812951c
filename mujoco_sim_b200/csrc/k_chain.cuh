// k_chain.cuh — the whole tick of a serial chain of N scalar joints in ONE kernel, one thread per environment
// (rows s1-s3, s7-s10, s12-s14 of SURVEY.md section 8a' plus joint-limit rows of s6 / s11 solved inline).
//
// Compared with the policy-generic k_smooth<ChainP<N>> (which keeps every per-body array of MuJoCo's mjData alive and
// spills ~3.8 KB per thread to local memory — 10x the algorithmic HBM traffic at large batches, see
// profiles/r01_ncu_smooth_c2_262144_summary.txt) this kernel is organised around loop-carried frames:
//   forward pass 1  : body frames parent -> child in registers; only xipos, the rotated inertia, the joint anchor and
//                     axis of each body survive the pass; xpos / xquat go straight to HBM
//   CoM             : cinert (10) and cdof (6) per body — the only per-body state the rest of the tick needs
//   forward pass 2  : cvel / cacc carried along the chain, cfrc (6) per body kept for the way back
//   backward pass   : running sums of cfrc (-> qfrc_bias) and of cinert (-> composite inertia -> rows of M)
// Joint limits — the only constraint source such a model can have here — are solved inline with the same
// acceleration-space PGS as k_pgs_team (identical iterates), so the tick needs no other kernel.
#pragma once
#include "k_args.h"
#include "k_common.cuh"
#include "k_constraint.cuh"
#include "k_policy.cuh"
#include "k_smooth.cuh"

namespace b2 {

template <typename T, int N, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_chain(const KArgs<T> a) {
  using P = ChainP<N>;
  constexpr int NM = N * (N + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* blob = reinterpret_cast<uint32_t*>(smem_raw + 16);
  const int nwords = reinterpret_cast<const DModel*>(a.model)->nwords;
  stage_model(blob, a.model, nwords, bar);
  MV<T> m{reinterpret_cast<const DModel*>(blob), blob};
  const DModel& h = *m.h;
  const long long S = a.nenvp;
  const int ntiles = a.nenvp / BLOCK;
  const bool grav = !(h.disableflags & DSBL_GRAVITY);
  const T g3[3] = {grav ? m.f(h.o_opt_real, 0) : T(0), grav ? m.f(h.o_opt_real, 1) : T(0), grav ? m.f(h.o_opt_real, 2) : T(0)};

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * BLOCK + threadIdx.x;
    T q[N], v[N], qa[N], fa[N];
    bool bad = false;
#pragma unroll
    for (int i = 0; i < N; i++) {
      q[i] = a.qpos[i * S + env]; v[i] = a.qvel[i * S + env]; qa[i] = a.qacc[i * S + env]; fa[i] = a.qfrc_applied[i * S + env];
      bad |= !(t_abs(q[i]) < T(1e10)) || !(t_abs(v[i]) < T(1e10));
    }
    // commands of the hardware interface: issued first so that a zero-copy read over PCIe overlaps the position stage
    const bool hwio = (a.flags & B2F_HWIO) && env < a.nenv;
    float hwv[N], hwe[N];
    if (hwio) {
#pragma unroll
      for (int i = 0; i < N; i++) { hwv[i] = a.hw_vel[(long long)i * a.nenv + env]; hwe[i] = a.hw_eff[(long long)i * a.nenv + env]; }
    }
    // MjHWInterface::read gathers (src/mujoco_sim/mj_hw_interface.cpp:62-70)
    auto hw_out = [&](const T* qo, const T* vo, const T* fo) {
      if (!hwio) return;
#pragma unroll
      for (int i = 0; i < N; i++) {
        a.hw_pos[(long long)i * a.nenv + env] = (float)qo[i];
        a.hw_velo[(long long)i * a.nenv + env] = (float)vo[i];
        a.hw_effo[(long long)i * a.nenv + env] = (float)fo[i];
      }
    };
    if (bad) {  // mj_checkPos / mj_checkVel
#pragma unroll
      for (int i = 0; i < N; i++) {
        q[i] = m.f(h.o_qpos0, i); v[i] = 0; qa[i] = 0; fa[i] = 0;
        a.qpos[i * S + env] = q[i]; a.qvel[i * S + env] = 0; a.qacc[i * S + env] = 0; a.qacc_warmstart[i * S + env] = 0; a.qfrc_applied[i * S + env] = 0;
      }
      a.time[env] = 0;
      a.status[env] |= 4;
    }

    // ---- forward pass 1: frames (A.2) ----
    T cin[N][10], cd[N][6];
    {
      T xip[N][3], Ir[N][6], anc[N][3], axs[N][3];
      T ppos[3] = {0, 0, 0}, pquat[4] = {1, 0, 0, 0}, com[3] = {0, 0, 0};
      {
        const T z3[3] = {0, 0, 0}, q1[4] = {1, 0, 0, 0};
        for (int k = 0; k < 3; k++) a.xpos[k * S + env] = z3[k];
        for (int k = 0; k < 4; k++) a.xquat[k * S + env] = q1[k];
      }
#pragma unroll
      for (int b = 1; b <= N; b++) {
        const int j = b - 1;
        T bp[3], bq[4], jp[3], jx[3], pos[3], quat[4], r[3];
        ldm<T, 3>(bp, m, h.o_body_pos, 3 * b);
        ldm<T, 4>(bq, m, h.o_body_quat, 4 * b);
        ldm<T, 3>(jp, m, h.o_jnt_pos, 3 * j);
        ldm<T, 3>(jx, m, h.o_jnt_axis, 3 * j);
        rot_vec_quat(r, bp, pquat);
        pos[0] = ppos[0] + r[0]; pos[1] = ppos[1] + r[1]; pos[2] = ppos[2] + r[2];
        mul_quat(quat, pquat, bq);
        rot_vec_quat(anc[j], jp, quat);
        anc[j][0] += pos[0]; anc[j][1] += pos[1]; anc[j][2] += pos[2];
        rot_vec_quat(axs[j], jx, quat);
        const T dq = q[j] - m.f(h.o_qpos0, j);
        if (m.i(h.o_jnt_type, j) == JNT_SLIDE) {
          pos[0] += axs[j][0] * dq; pos[1] += axs[j][1] * dq; pos[2] += axs[j][2] * dq;
        } else {
          T ql[4], qn[4], off[3];
          axis_angle2quat(ql, jx, dq);
          mul_quat(qn, quat, ql);
          quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];
          rot_vec_quat(off, jp, quat);
          pos[0] = anc[j][0] - off[0]; pos[1] = anc[j][1] - off[1]; pos[2] = anc[j][2] - off[2];
        }
        normalize4(quat);
        for (int k = 0; k < 3; k++) a.xpos[(3 * b + k) * S + env] = pos[k];
        for (int k = 0; k < 4; k++) a.xquat[(4 * b + k) * S + env] = quat[k];
        T ip[3], iq[4], qi[4], mat[9], inert[3];
        ldm<T, 3>(ip, m, h.o_body_ipos, 3 * b);
        ldm<T, 4>(iq, m, h.o_body_iquat, 4 * b);
        ldm<T, 3>(inert, m, h.o_body_inertia, 3 * b);
        rot_vec_quat(r, ip, quat);
        xip[j][0] = pos[0] + r[0]; xip[j][1] = pos[1] + r[1]; xip[j][2] = pos[2] + r[2];
        mul_quat(qi, quat, iq);
        quat2mat(mat, qi);
        // R diag(I) R^T, upper triangle: xx yy zz xy xz yz
        Ir[j][0] = mat[0] * inert[0] * mat[0] + mat[1] * inert[1] * mat[1] + mat[2] * inert[2] * mat[2];
        Ir[j][1] = mat[3] * inert[0] * mat[3] + mat[4] * inert[1] * mat[4] + mat[5] * inert[2] * mat[5];
        Ir[j][2] = mat[6] * inert[0] * mat[6] + mat[7] * inert[1] * mat[7] + mat[8] * inert[2] * mat[8];
        Ir[j][3] = mat[0] * inert[0] * mat[3] + mat[1] * inert[1] * mat[4] + mat[2] * inert[2] * mat[5];
        Ir[j][4] = mat[0] * inert[0] * mat[6] + mat[1] * inert[1] * mat[7] + mat[2] * inert[2] * mat[8];
        Ir[j][5] = mat[3] * inert[0] * mat[6] + mat[4] * inert[1] * mat[7] + mat[5] * inert[2] * mat[8];
        const T mass = m.f(h.o_body_mass, b);
        com[0] += mass * xip[j][0]; com[1] += mass * xip[j][1]; com[2] += mass * xip[j][2];
        for (int k = 0; k < 3; k++) ppos[k] = pos[k];
        for (int k = 0; k < 4; k++) pquat[k] = quat[k];
      }
      // ---- CoM frame quantities (A.3): the chain has a single root, body 1 ----
      {
        const T sm = m.f(h.o_body_subtreemass, 1);
        if (sm < Eps<T>::minval()) { com[0] = xip[0][0]; com[1] = xip[0][1]; com[2] = xip[0][2]; }
        else { const T inv = T(1) / sm; com[0] *= inv; com[1] *= inv; com[2] *= inv; }
      }
#pragma unroll
      for (int j = 0; j < N; j++) {
        const T mass = m.f(h.o_body_mass, j + 1);
        const T dif[3] = {xip[j][0] - com[0], xip[j][1] - com[1], xip[j][2] - com[2]};
        cin[j][0] = Ir[j][0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
        cin[j][1] = Ir[j][1] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
        cin[j][2] = Ir[j][2] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
        cin[j][3] = Ir[j][3] - mass * dif[0] * dif[1];
        cin[j][4] = Ir[j][4] - mass * dif[0] * dif[2];
        cin[j][5] = Ir[j][5] - mass * dif[1] * dif[2];
        cin[j][6] = mass * dif[0]; cin[j][7] = mass * dif[1]; cin[j][8] = mass * dif[2];
        cin[j][9] = mass;
        if (m.i(h.o_jnt_type, j) == JNT_SLIDE) {
          cd[j][0] = 0; cd[j][1] = 0; cd[j][2] = 0;
          cd[j][3] = axs[j][0]; cd[j][4] = axs[j][1]; cd[j][5] = axs[j][2];
        } else {
          const T off[3] = {com[0] - anc[j][0], com[1] - anc[j][1], com[2] - anc[j][2]};
          cd[j][0] = axs[j][0]; cd[j][1] = axs[j][1]; cd[j][2] = axs[j][2];
          cross3(cd[j] + 3, axs[j], off);
        }
      }
    }

    // ---- forward pass 2 + backward pass: cvel / cacc along the chain, cfrc back (A.5), CRBA (A.4) ----
    T bias[N], pas[N];
    LArr<T, NM> qM;
    auto velocity_stage = [&]() {
      T cf[N][6];
      T cv[6] = {0, 0, 0, 0, 0, 0}, ca[6] = {0, 0, 0, -g3[0], -g3[1], -g3[2]};
#pragma unroll
      for (int i = 0; i < N; i++) {
        T dd[6], Ia[6], Iv[6], x[6];
        cross_motion(dd, cv, cd[i]);
        for (int r = 0; r < 6; r++) { cv[r] += cd[i][r] * v[i]; ca[r] += dd[r] * v[i]; }
        mul_inert_vec(Ia, cin[i], ca);
        mul_inert_vec(Iv, cin[i], cv);
        cross_force(x, cv, Iv);
        for (int r = 0; r < 6; r++) cf[i][r] = Ia[r] + x[r];
      }
      T fs[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int i = N - 1; i >= 0; i--) {
        T s = 0;
        for (int r = 0; r < 6; r++) { fs[r] += cf[i][r]; s += cd[i][r] * fs[r]; }
        bias[i] = s;
      }
      // passive forces: springs, dampers, gravity compensation (reference default: gravcomp = 1 on robot bodies)
#pragma unroll
      for (int i = 0; i < N; i++) pas[i] = 0;
      if (!(h.disableflags & DSBL_PASSIVE)) {
        if (h.has_stiffness) {
#pragma unroll
          for (int i = 0; i < N; i++) pas[i] -= m.f(h.o_jnt_stiffness, i) * (q[i] - m.f(h.o_qpos_spring, i));
        }
        if (h.has_damping) {
#pragma unroll
          for (int i = 0; i < N; i++) pas[i] -= m.f(h.o_dof_damping, i) * v[i];
        }
        if (h.has_gravcomp && grav) {
          T wf[3] = {0, 0, 0}, wt[3] = {0, 0, 0};  // wrench of the compensating forces of the bodies below, about the CoM
#pragma unroll
          for (int i = N - 1; i >= 0; i--) {
            const T gc = m.f(h.o_body_gravcomp, i + 1);
            if (gc != 0) {
              const T f[3] = {-g3[0] * cin[i][9] * gc, -g3[1] * cin[i][9] * gc, -g3[2] * cin[i][9] * gc};
              const T mo[3] = {-cin[i][6] * gc, -cin[i][7] * gc, -cin[i][8] * gc}, t[3] = {0, 0, 0};
              T tq[3];
              cross3(tq, mo, g3);  // (mass off) x (-g gc) = off x f
              (void)t;
              wf[0] += f[0]; wf[1] += f[1]; wf[2] += f[2];
              wt[0] += tq[0]; wt[1] += tq[1]; wt[2] += tq[2];
            }
            pas[i] += cd[i][3] * wf[0] + cd[i][4] * wf[1] + cd[i][5] * wf[2] + cd[i][0] * wt[0] + cd[i][1] * wt[1] + cd[i][2] * wt[2];
          }
        }
      }
    };
    velocity_stage();
    {
      T crb[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int i = N - 1; i >= 0; i--) {
        T buf[6];
        for (int k = 0; k < 10; k++) crb[k] += cin[i][k];
        mul_inert_vec(buf, crb, cd[i]);
#pragma unroll
        for (int j = i; j >= 0; j--) {
          T s = (j == i) ? m.f(h.o_dof_armature, i) : T(0);
          for (int r = 0; r < 6; r++) s += cd[j][r] * buf[r];
          qM[i * (i + 1) / 2 + (i - j)] = s;
        }
      }
    }
    auto mul_M = [&](T* res, const T* vec) {
#pragma unroll
      for (int i = 0; i < N; i++) res[i] = 0;
#pragma unroll
      for (int i = 0; i < N; i++) {
        res[i] += qM[i * (i + 1) / 2] * vec[i];
#pragma unroll
        for (int j = i - 1; j >= 0; j--) {
          const T mij = qM[i * (i + 1) / 2 + (i - j)];
          res[i] += mij * vec[j];
          res[j] += mij * vec[i];
        }
      }
    };

    // ---- mjcb_control -> MjSim::controller (src/mujoco_sim/mj_sim.cpp:1055-1077) ----
    bool overridden = false;
    if (a.flags & B2F_CONTROLLER) {
      T ddq[N], dqc[N], tau[N];
#pragma unroll
      for (int i = 0; i < N; i++) { ddq[i] = a.ddq[i * S + env]; dqc[i] = a.dq[i * S + env]; }
      if (hwio) {  // MjHWInterface::write (src/mujoco_sim/mj_hw_interface.cpp:73-91), hardware joint i == dof i
#pragma unroll
        for (int i = 0; i < N; i++) {
          if (!m.i(h.o_dof_controlled, i)) continue;
          if (fabsf(hwv[i]) > 1e-15f) dqc[i] = (T)hwv[i];
          else ddq[i] = a.hw_kp ? (T)a.hw_kp[i] * ((T)hwe[i] - q[i]) - (T)a.hw_kd[i] * v[i] : (T)hwe[i];   // PD stage (b2_set_pd)
        }
      }
      mul_M(tau, ddq);
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (m.i(h.o_dof_controlled, i)) tau[i] += bias[i];
        fa[i] = tau[i];
        a.qfrc_applied[i * S + env] = tau[i];
        const T dv = dqc[i];
        if (t_abs(dv) > Eps<T>::minval()) { v[i] = dv; overridden = true; }
        a.ddq[i * S + env] = 0;
        a.dq[i * S + env] = 0;
      }
    }
    if ((a.flags & B2F_INVERSE) && overridden) velocity_stage();  // read() -> mj_inverse sees the overridden qvel
#pragma unroll
    for (int i = 0; i < N; i++) a.qfrc_bias[i * S + env] = bias[i];
    T finv[N];
#pragma unroll
    for (int i = 0; i < N; i++) finv[i] = 0;
    if (a.flags & B2F_INVERSE) {
      // RNE(q, v, a) + armature a = M a + bias  (see k_smooth.cuh)
      mul_M(finv, qa);
#pragma unroll
      for (int i = 0; i < N; i++) finv[i] += bias[i] - pas[i];
    }

    // ---- smooth acceleration ----
    T fsm[N];
    LArr<T, NM> LD;
    LArr<T, N> dinv, accs;
#pragma unroll
    for (int i = 0; i < N; i++) { fsm[i] = pas[i] - bias[i] + fa[i]; accs[i] = fsm[i]; }
#pragma unroll
    for (int i = 0; i < NM; i++) LD[i] = qM[i];
    ld_factor<P>(m, LD, dinv);
    ld_solve<P>(m, LD, dinv, accs);

    // ---- joint limits: the only constraint rows this model can have (A.7), solved inline (A.8) ----
    T acc[N], qfc[N];
#pragma unroll
    for (int i = 0; i < N; i++) { acc[i] = accs[i]; qfc[i] = 0; }
    unsigned act = 0;  // bit 2 j: lower side of joint j active, bit 2 j + 1: upper side
    if (h.has_limits && !(h.disableflags & (DSBL_LIMIT | DSBL_CONSTRAINT))) {
#pragma unroll
      for (int j = 0; j < N; j++) {
        if (!m.i(h.o_jnt_limited, j)) continue;
        const T mg = m.f(h.o_jnt_margin, j);
        if (q[j] - m.f(h.o_jnt_range, 2 * j) < mg) act |= 1u << (2 * j);
        if (m.f(h.o_jnt_range, 2 * j + 1) - q[j] < mg) act |= 1u << (2 * j + 1);
      }
    }
    int iters = 0;
    if (act) {
      T Bc[N][N];           // column j of M^-1 for joints with an active side
      T R[2 * N], aref[2 * N], f[2 * N];
#pragma unroll
      for (int j = 0; j < N; j++) {
        if (!((act >> (2 * j)) & 3u)) continue;
        LArr<T, N> col;
#pragma unroll
        for (int i = 0; i < N; i++) col[i] = (i == j) ? T(1) : T(0);
        ld_solve<P>(m, LD, dinv, col);
#pragma unroll
        for (int i = 0; i < N; i++) Bc[j][i] = col[i];
        T solref[2], solimp[5];
        ldm<T, 2>(solref, m, h.o_jnt_solref, 2 * j);
        ldm<T, 5>(solimp, m, h.o_jnt_solimp, 5 * j);
        const T mg = m.f(h.o_jnt_margin, j), diag = m.f(h.o_dof_invweight0, j);
        const T dmax = t_min(T(0.9999), t_max(T(0.0001), solimp[1]));
        T K, Bd;
        if (solref[0] > 0) {
          T tc = solref[0];
          if (!(h.disableflags & DSBL_REFSAFE)) tc = t_max(tc, 2 * a.dt());
          K = 1 / t_max(Eps<T>::minval(), dmax * dmax * tc * tc * solref[1] * solref[1]);
          Bd = 2 / t_max(Eps<T>::minval(), dmax * tc);
        } else {
          K = -solref[0] / t_max(Eps<T>::minval(), dmax * dmax);
          Bd = -solref[1] / t_max(Eps<T>::minval(), dmax);
        }
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const int s = 2 * j + side;
          if (!((act >> s) & 1u)) continue;
          const T js = side ? T(-1) : T(1);
          const T pos = side ? m.f(h.o_jnt_range, 2 * j + 1) - q[j] : q[j] - m.f(h.o_jnt_range, 2 * j);
          const T imp = impedance(solimp, pos, mg);
          R[s] = t_max(Eps<T>::minval(), (1 - imp) * diag / imp);
          aref[s] = -Bd * (js * v[j]) - K * imp * (pos - mg);
          if (a.flags & B2F_INVERSE) {
            const T jar = js * qa[j] - aref[s];
            if (jar < 0) finv[j] -= js * (-jar / R[s]);
          }
        }
      }
      if (!(a.flags & B2F_NOSOLVE)) {
        // warm start: forces implied by qacc_warmstart, kept only if their dual cost is negative
        bool warm = !(h.disableflags & DSBL_WARMSTART);
        T aw[N];
#pragma unroll
        for (int i = 0; i < N; i++) aw[i] = 0;
        if (warm) {
#pragma unroll
          for (int s = 0; s < 2 * N; s++) {
            f[s] = 0;
            if (!((act >> s) & 1u)) continue;
            const int j = s / 2;
            const T js = (s & 1) ? T(-1) : T(1);
            const T jw = js * a.qacc_warmstart[j * S + env] - aref[s];
            f[s] = jw < 0 ? -jw / R[s] : T(0);
            if (f[s] != 0) for (int i = 0; i < N; i++) aw[i] += f[s] * js * Bc[j][i];
          }
          T cost = 0;
#pragma unroll
          for (int s = 0; s < 2 * N; s++) {
            if (!((act >> s) & 1u) || f[s] == 0) continue;
            const int j = s / 2;
            const T js = (s & 1) ? T(-1) : T(1);
            const T Af = js * aw[j] + R[s] * f[s];
            cost += f[s] * (T(0.5) * Af + (js * accs[j] - aref[s]));
          }
          if (cost > 0) warm = false;
        }
        if (warm) {
#pragma unroll
          for (int i = 0; i < N; i++) acc[i] = accs[i] + aw[i];
        } else {
#pragma unroll
          for (int s = 0; s < 2 * N; s++) f[s] = 0;
        }
        const T tol = m.f(h.o_opt_real, 3), scale = 1 / (m.f(h.o_opt_real, 4) * T(N));
        for (int it = 0; it < h.iterations; it++) {
          T improvement = 0;
#pragma unroll
          for (int s = 0; s < 2 * N; s++) {
            if (!((act >> s) & 1u)) continue;
            const int j = s / 2;
            const T js = (s & 1) ? T(-1) : T(1);
            const T old = f[s];
            const T res = js * acc[j] + R[s] * old - aref[s];
            const T Arr = Bc[j][j] + R[s];
            const T fn = t_max(T(0), old - res / Arr);
            const T delta = fn - old;
            const T change = T(0.5) * delta * delta * Arr + delta * res;
            if (delta != 0 && !(change > T(1e-10))) {
              f[s] = fn;
              improvement -= change;
              for (int i = 0; i < N; i++) acc[i] += delta * js * Bc[j][i];
            }
          }
          iters = it + 1;
          if (improvement * scale < tol) break;
        }
#pragma unroll
        for (int s = 0; s < 2 * N; s++)
          if ((act >> s) & 1u) qfc[s / 2] += ((s & 1) ? T(-1) : T(1)) * f[s];
      }
    }
    if (a.flags & B2F_INVERSE) {
#pragma unroll
      for (int i = 0; i < N; i++) a.qfrc_inverse[i * S + env] = finv[i];
    }
    a.nefc[env] = __popc(act);
    a.solver_iter[env] = iters;
    a.status[env] = (a.status[env] & 7) | 8;
    if (a.flags & B2F_NOSOLVE) {
      if (overridden) {
#pragma unroll
        for (int i = 0; i < N; i++) a.qvel[i * S + env] = v[i];
      }
      continue;
    }

    // ---- qacc, warm start, mj_checkAcc, semi-implicit Euler (A.9), odom ----
    bool badacc = false;
#pragma unroll
    for (int i = 0; i < N; i++) badacc |= !(t_abs(acc[i]) < T(1e10));
    if (badacc) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        a.qpos[i * S + env] = m.f(h.o_qpos0, i); a.qvel[i * S + env] = 0; a.qacc[i * S + env] = 0; a.qacc_warmstart[i * S + env] = 0;
        a.qfrc_applied[i * S + env] = 0;
      }
      a.time[env] = 0;
      a.status[env] |= 4;
      if (hwio) {
        T q0[N], z[N];
#pragma unroll
        for (int i = 0; i < N; i++) { q0[i] = m.f(h.o_qpos0, i); z[i] = 0; }
        hw_out(q0, z, finv);
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < N; i++) { a.qacc[i * S + env] = acc[i]; a.qacc_warmstart[i * S + env] = acc[i]; }
    // MjHWInterface::read sits between mj_step1 and mj_step2 (src/mj_main.cpp:91-108): joint states before the integration
    if (!(a.flags & B2F_READ_POST)) hw_out(q, v, finv);
    if (a.flags & B2F_INTEGRATE) {
      LArr<T, N> xa;
      if (h.has_damping && !(h.disableflags & DSBL_EULERDAMP)) {
#pragma unroll
        for (int i = 0; i < NM; i++) LD[i] = qM[i];
#pragma unroll
        for (int i = 0; i < N; i++) { LD[i * (i + 1) / 2] += a.dt() * m.f(h.o_dof_damping, i); xa[i] = fsm[i] + qfc[i]; }
        ld_factor<P>(m, LD, dinv);
        ld_solve<P>(m, LD, dinv, xa);
      } else {
#pragma unroll
        for (int i = 0; i < N; i++) xa[i] = acc[i];
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        v[i] += a.dt() * xa[i];
        q[i] += a.dt() * v[i];
        a.qvel[i * S + env] = v[i];
        a.qpos[i * S + env] = q[i];
      }
      a.time[env] += a.dt();
      if (a.flags & B2F_ODOM) odom_override(m, a, env);
    } else if (overridden) {
#pragma unroll
      for (int i = 0; i < N; i++) a.qvel[i * S + env] = v[i];
    }
    if (a.flags & B2F_READ_POST) hw_out(q, v, finv);
  }
}

}  // namespace b2
