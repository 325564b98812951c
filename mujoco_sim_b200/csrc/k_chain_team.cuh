// k_chain_team.cuh — the whole tick of a serial chain of N <= 8 scalar joints in ONE kernel, an 8-lane TEAM per
// environment (lane l owns body l + 1 / joint l / dof l; four environments per warp).
//
// k_chain (one thread per environment) runs ~13.5 k dependent instructions per tick: with a few thousand environments
// there is one warp per SM and the tick costs the latency of that single instruction stream (~26 us on B200).  Here the
// articulated-body recursions become warp-shuffle scans inside the team, so the dependent chain shrinks to a few
// thousand cycles and a 4096-environment batch puts 7 warps on every SM:
//   frames (A.2)      local joint transforms in parallel, world frames by an inclusive scan of rigid transforms (3 rounds)
//   CoM (A.3)         subtree CoM by a butterfly sum, cinert / cdof per lane
//   velocities (A.5)  cvel, cacc: inclusive scans of 6-vectors; cfrc summed back with a reverse scan -> qfrc_bias
//   CRBA (A.4)        composite inertia = reverse scan of cinert; row l of M from broadcast cdof columns
//   L^T D L           right-looking elimination, lane l keeps row l and column l of L; solves are 2 N shuffle steps
//   joint limits      row (joint j, side) lives on lane j; the acceleration-space PGS of k_chain with one broadcast per row
// Same stage order, formulas and iterates as k_chain / the fp64 oracle; sums along the chain are associated pairwise
// instead of left to right, which moves results by a few ulp (covered by the written fp32 tolerance, 1e-9 in B2_F64).
// Selected for small batches (b2_batch::chain_team); large batches are issue-bound and keep k_chain.
#pragma once
#include "k_args.h"
#include "k_common.cuh"
#include "k_constraint.cuh"

namespace b2 {

namespace team8 {
constexpr unsigned FULL = 0xffffffffu;
template <typename T> __device__ __forceinline__ T bc(T v, int src) { return __shfl_sync(FULL, v, src, 8); }
template <typename T> __device__ __forceinline__ T up(T v, int d) { return __shfl_up_sync(FULL, v, d, 8); }
template <typename T> __device__ __forceinline__ T dn(T v, int d) { return __shfl_down_sync(FULL, v, d, 8); }
template <typename T> __device__ __forceinline__ T bfly(T v, int d) { return __shfl_xor_sync(FULL, v, d, 8); }
// inclusive prefix sum over the lanes of a team, K values at once
template <typename T, int K> __device__ __forceinline__ void scan_up(T* x, int l) {
#pragma unroll
  for (int d = 1; d < 8; d *= 2) {
#pragma unroll
    for (int k = 0; k < K; k++) { const T o = up(x[k], d); if (l >= d) x[k] += o; }
  }
}
// inclusive suffix sum (lane l <- sum over lanes >= l)
template <typename T, int K> __device__ __forceinline__ void scan_dn(T* x, int l) {
#pragma unroll
  for (int d = 1; d < 8; d *= 2) {
#pragma unroll
    for (int k = 0; k < K; k++) { const T o = dn(x[k], d); if (l + d < 8) x[k] += o; }
  }
}
__device__ __forceinline__ bool team_any(bool p) {
  const unsigned b = __ballot_sync(FULL, p);
  return ((b >> (threadIdx.x & 24)) & 0xffu) != 0;
}
__device__ __forceinline__ unsigned team_bits(bool p) {
  const unsigned b = __ballot_sync(FULL, p);
  return (b >> (threadIdx.x & 24)) & 0xffu;
}
}  // namespace team8

template <typename T, int N, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_chain_team(const KArgs<T> a) {
  using namespace team8;
  static_assert(N >= 1 && N <= 8, "one lane per joint");
  constexpr int EPB = BLOCK / 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* blob = reinterpret_cast<uint32_t*>(smem_raw + 16);
  const int nwords = a.model_words;
  stage_model_issue(blob, a.model, nwords, bar);   // TMA bulk copy of the model constants; waited for below
  MV<T> m{reinterpret_cast<const DModel*>(blob), blob};
  const DModel& h = *m.h;
  T* Msh = reinterpret_cast<T*>(smem_raw + 16 + (size_t)nwords * 4) + (threadIdx.x >> 3) * 64;  // [8][8] per team
  // hardware-interface staging [5][8][EPB]: the exchange buffers may be mapped host memory, so the CTA touches them with
  // contiguous runs of EPB environments per joint instead of one 16-byte piece per team
  float* hwsh = reinterpret_cast<float*>(smem_raw + 16 + (size_t)nwords * 4 + (size_t)EPB * 64 * sizeof(T));
  const long long S = a.nenvp;
  const int ntiles = a.nenvp / EPB;
  const int l = threadIdx.x & 7;
  const bool on = l < N;
  // the first tile's state goes in flight while the model constants arrive
  T pre[6] = {0, 0, 0, 0, 0, 0};
  int pre_status = 0;
  auto prefetch = [&](int tile) {
    if (tile >= ntiles) return;
    const int env = tile * EPB + (threadIdx.x >> 3);
    const long long at = (long long)l * S + env;
    if (on) {
      pre[0] = a.qpos[at]; pre[1] = a.qvel[at]; pre[2] = a.qacc[at]; pre[3] = a.qfrc_applied[at];
      if (a.flags & B2F_CONTROLLER) { pre[4] = a.ddq[at]; pre[5] = a.dq[at]; }
    }
    if (l == 0) pre_status = a.status[env];
  };
  prefetch(blockIdx.x);
  stage_model_wait(bar);
  const int j = on ? l : N - 1;   // index used for model reads (kept in range on idle lanes)
  const bool grav = !(h.disableflags & DSBL_GRAVITY);
  const T g3[3] = {grav ? m.f(h.o_opt_real, 0) : T(0), grav ? m.f(h.o_opt_real, 1) : T(0), grav ? m.f(h.o_opt_real, 2) : T(0)};

  // per-lane model constants
  T bp[3], bq[4], jp[3], jx[3], ip[3], iq[4], inert[3];
  ldm<T, 3>(bp, m, h.o_body_pos, 3 * (j + 1));
  ldm<T, 4>(bq, m, h.o_body_quat, 4 * (j + 1));
  ldm<T, 3>(jp, m, h.o_jnt_pos, 3 * j);
  ldm<T, 3>(jx, m, h.o_jnt_axis, 3 * j);
  ldm<T, 3>(ip, m, h.o_body_ipos, 3 * (j + 1));
  ldm<T, 4>(iq, m, h.o_body_iquat, 4 * (j + 1));
  ldm<T, 3>(inert, m, h.o_body_inertia, 3 * (j + 1));
  const T mass = on ? m.f(h.o_body_mass, j + 1) : T(0);
  const T q0 = m.f(h.o_qpos0, j), arm = m.f(h.o_dof_armature, j), damp = m.f(h.o_dof_damping, j);
  const bool slide = m.i(h.o_jnt_type, j) == JNT_SLIDE;
  const bool ctl = on && m.i(h.o_dof_controlled, j) != 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * EPB + (threadIdx.x >> 3);
    const long long at = (long long)l * S + env;
    T q = q0, v = 0, qa = 0, fa = 0, ddq_in = 0, dq_in = 0;
    if (on) { q = pre[0]; v = pre[1]; qa = pre[2]; fa = pre[3]; ddq_in = pre[4]; dq_in = pre[5]; }
    const int status_in = pre_status;
    const bool hwio = (a.flags & B2F_HWIO) != 0;   // CTA-uniform
    // commands: issued now (coalesced), parked in registers while the position stage runs, exchanged through shared
    // memory in front of the controller
    float hwc[2][(N * EPB + BLOCK - 1) / BLOCK] = {};
    if (hwio) {
#pragma unroll
      for (int u = 0; u < (N * EPB + BLOCK - 1) / BLOCK; u++) {
        const int idx = threadIdx.x + u * BLOCK, jn = idx / EPB, e2 = tile * EPB + idx % EPB;
        if (idx < N * EPB && e2 < a.nenv) { hwc[0][u] = a.hw_vel[(long long)jn * a.nenv + e2]; hwc[1][u] = a.hw_eff[(long long)jn * a.nenv + e2]; }
      }
    }
    T finv = 0;
    auto hw_out = [&](T qo, T vo, T fo) {   // MjHWInterface::read gathers (src/mujoco_sim/mj_hw_interface.cpp:62-70)
      if (!hwio || !on) return;
      hwsh[(2 * 8 + l) * EPB + (threadIdx.x >> 3)] = (float)qo;
      hwsh[(3 * 8 + l) * EPB + (threadIdx.x >> 3)] = (float)vo;
      hwsh[(4 * 8 + l) * EPB + (threadIdx.x >> 3)] = (float)fo;
    };
    int stat = 0;
    if (team_any(on && (!(t_abs(q) < T(1e10)) || !(t_abs(v) < T(1e10))))) {  // mj_checkPos / mj_checkVel
      q = q0; v = 0; qa = 0; fa = 0;
      if (on) { a.qpos[at] = q; a.qvel[at] = 0; a.qacc[at] = 0; a.qacc_warmstart[at] = 0; a.qfrc_applied[at] = 0; }
      if (l == 0) a.time[env] = 0;
      stat = 4;
    }

    // ---- frames (A.2): local transform of body l + 1 in its parent, then an inclusive scan of rigid transforms ----
    T pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};
    if (on) {
      const T dq = q - q0;
      if (slide) {
        T r[3];
        rot_vec_quat(r, jx, bq);
        for (int k = 0; k < 3; k++) pos[k] = bp[k] + r[k] * dq;
        for (int k = 0; k < 4; k++) quat[k] = bq[k];
      } else {
        T ql[4], r1[3], r2[3];
        axis_angle2quat(ql, jx, dq);
        mul_quat(quat, bq, ql);
        rot_vec_quat(r1, jp, bq);
        rot_vec_quat(r2, jp, quat);
        for (int k = 0; k < 3; k++) pos[k] = bp[k] + r1[k] - r2[k];
      }
    }
#pragma unroll
    for (int d = 1; d < 8; d *= 2) {
      T op[3], oq[4];
#pragma unroll
      for (int k = 0; k < 3; k++) op[k] = up(pos[k], d);
#pragma unroll
      for (int k = 0; k < 4; k++) oq[k] = up(quat[k], d);
      if (l >= d) {
        T r[3], qn[4];
        rot_vec_quat(r, pos, oq);
        mul_quat(qn, oq, quat);
        for (int k = 0; k < 3; k++) pos[k] = op[k] + r[k];
        for (int k = 0; k < 4; k++) quat[k] = qn[k];
      }
    }
    normalize4(quat);
    if (on) {
      for (int k = 0; k < 3; k++) a.xpos[(3 * (l + 1) + k) * S + env] = pos[k];
      for (int k = 0; k < 4; k++) a.xquat[(4 * (l + 1) + k) * S + env] = quat[k];
    }
    if (l == 0) {
      for (int k = 0; k < 3; k++) a.xpos[k * S + env] = 0;
      for (int k = 0; k < 4; k++) a.xquat[k * S + env] = k == 0 ? T(1) : T(0);
    }
    T cin[10], cd[6];
    {
      T anc[3], axs[3], xip[3], qi[4], mat[9], Ir[6], r[3];
      rot_vec_quat(anc, jp, quat);
      for (int k = 0; k < 3; k++) anc[k] += pos[k];
      rot_vec_quat(axs, jx, quat);
      rot_vec_quat(r, ip, quat);
      for (int k = 0; k < 3; k++) xip[k] = pos[k] + r[k];
      mul_quat(qi, quat, iq);
      quat2mat(mat, qi);
      Ir[0] = mat[0] * inert[0] * mat[0] + mat[1] * inert[1] * mat[1] + mat[2] * inert[2] * mat[2];
      Ir[1] = mat[3] * inert[0] * mat[3] + mat[4] * inert[1] * mat[4] + mat[5] * inert[2] * mat[5];
      Ir[2] = mat[6] * inert[0] * mat[6] + mat[7] * inert[1] * mat[7] + mat[8] * inert[2] * mat[8];
      Ir[3] = mat[0] * inert[0] * mat[3] + mat[1] * inert[1] * mat[4] + mat[2] * inert[2] * mat[5];
      Ir[4] = mat[0] * inert[0] * mat[6] + mat[1] * inert[1] * mat[7] + mat[2] * inert[2] * mat[8];
      Ir[5] = mat[3] * inert[0] * mat[6] + mat[4] * inert[1] * mat[7] + mat[5] * inert[2] * mat[8];
      // ---- CoM frame quantities (A.3): single root, body 1 ----
      T com[3] = {mass * xip[0], mass * xip[1], mass * xip[2]};
#pragma unroll
      for (int d = 1; d < 8; d *= 2) {
#pragma unroll
        for (int k = 0; k < 3; k++) com[k] += bfly(com[k], d);
      }
      const T sm = m.f(h.o_body_subtreemass, 1);
      if (sm < Eps<T>::minval()) { for (int k = 0; k < 3; k++) com[k] = bc(xip[k], 0); }
      else { const T inv = T(1) / sm; com[0] *= inv; com[1] *= inv; com[2] *= inv; }
      const T dif[3] = {xip[0] - com[0], xip[1] - com[1], xip[2] - com[2]};
      cin[0] = Ir[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
      cin[1] = Ir[1] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
      cin[2] = Ir[2] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
      cin[3] = Ir[3] - mass * dif[0] * dif[1];
      cin[4] = Ir[4] - mass * dif[0] * dif[2];
      cin[5] = Ir[5] - mass * dif[1] * dif[2];
      cin[6] = mass * dif[0]; cin[7] = mass * dif[1]; cin[8] = mass * dif[2];
      cin[9] = mass;
      if (slide) {
        cd[0] = 0; cd[1] = 0; cd[2] = 0; cd[3] = axs[0]; cd[4] = axs[1]; cd[5] = axs[2];
      } else {
        const T off[3] = {com[0] - anc[0], com[1] - anc[1], com[2] - anc[2]};
        cd[0] = axs[0]; cd[1] = axs[1]; cd[2] = axs[2];
        cross3(cd + 3, axs, off);
      }
      if (!on) {
#pragma unroll
        for (int k = 0; k < 10; k++) cin[k] = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) cd[k] = 0;
      }
    }

    // ---- velocity stage (A.5): cvel / cacc down the chain, cfrc back; passive forces ----
    T bias = 0, pas = 0;
    auto velocity_stage = [&]() {
      T cv[6], cvx[6], dd[6], ca[6], Ia[6], Iv[6], x[6], fs[6];
#pragma unroll
      for (int r = 0; r < 6; r++) cv[r] = cd[r] * v;
      scan_up<T, 6>(cv, l);
#pragma unroll
      for (int r = 0; r < 6; r++) { const T o = up(cv[r], 1); cvx[r] = l ? o : T(0); }
      cross_motion(dd, cvx, cd);
#pragma unroll
      for (int r = 0; r < 6; r++) ca[r] = dd[r] * v;
      scan_up<T, 6>(ca, l);
      ca[3] -= g3[0]; ca[4] -= g3[1]; ca[5] -= g3[2];
      mul_inert_vec(Ia, cin, ca);
      mul_inert_vec(Iv, cin, cv);
      cross_force(x, cv, Iv);
#pragma unroll
      for (int r = 0; r < 6; r++) fs[r] = Ia[r] + x[r];
      scan_dn<T, 6>(fs, l);
      bias = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) bias += cd[r] * fs[r];
      pas = 0;
      if (!(h.disableflags & DSBL_PASSIVE)) {
        if (h.has_stiffness) pas -= m.f(h.o_jnt_stiffness, j) * (q - m.f(h.o_qpos_spring, j));
        if (h.has_damping) pas -= damp * v;
        if (h.has_gravcomp && grav) {
          const T gc = on ? m.f(h.o_body_gravcomp, j + 1) : T(0);
          T w[6] = {0, 0, 0, 0, 0, 0};  // [torque about the CoM; force] of the compensating forces at and below this body
          if (gc != 0) {
            const T mo[3] = {-cin[6] * gc, -cin[7] * gc, -cin[8] * gc};
            cross3(w, mo, g3);
            w[3] = -g3[0] * cin[9] * gc; w[4] = -g3[1] * cin[9] * gc; w[5] = -g3[2] * cin[9] * gc;
          }
          scan_dn<T, 6>(w, l);
          pas += cd[3] * w[3] + cd[4] * w[4] + cd[5] * w[5] + cd[0] * w[0] + cd[1] * w[1] + cd[2] * w[2];
        }
      }
      if (!on) { bias = 0; pas = 0; }
    };
    velocity_stage();

    // ---- CRBA (A.4): lane l gets row l of M (full symmetric row, the upper part through shared memory) ----
    T Mr[N];
    {
      T crb[10], buf[6];
#pragma unroll
      for (int k = 0; k < 10; k++) crb[k] = cin[k];
      scan_dn<T, 10>(crb, l);
      mul_inert_vec(buf, crb, cd);
      __syncwarp();
#pragma unroll
      for (int c = 0; c < N; c++) {
        T s = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) s += bc(cd[r], c) * buf[r];
        if (c == l) s += arm;
        Mr[c] = s;
        if (c <= l) Msh[l * 8 + c] = s;
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < N; c++)
        if (c > l && on) Mr[c] = Msh[c * 8 + l];
    }
    auto mul_M = [&](T x) {
      T r = 0;
#pragma unroll
      for (int c = 0; c < N; c++) r += Mr[c] * bc(x, c);
      return r;
    };
    // L^T D L of W (rows in lanes, lower part), MuJoCo's elimination order; lane l keeps row l (Lr) and column l (Lc) of L
    T Lr[N], Lc[N], dinv = 1;
    auto factor = [&](T* W) {
#pragma unroll
      for (int k = N - 1; k >= 0; k--) {
        T rk[N];
#pragma unroll
        for (int c = 0; c <= k; c++) rk[c] = bc(W[c], k);
        const T inv = T(1) / rk[k];
        T mine = 0;
#pragma unroll
        for (int c = 0; c < k; c++) mine = (c == l) ? rk[c] : mine;
        const T tmp = mine * inv;
        if (l == k) {
          dinv = inv;
#pragma unroll
          for (int c = 0; c < k; c++) Lr[c] = W[c] * inv;
        }
        if (l < k) {
          Lc[k] = tmp;
#pragma unroll
          for (int c = 0; c < k; c++) if (c <= l) W[c] -= rk[c] * tmp;
        }
      }
    };
    auto solve = [&](T x) {
#pragma unroll
      for (int i = N - 1; i >= 1; i--) { const T xi = bc(x, i); if (l < i) x -= Lc[i] * xi; }
      x *= dinv;
#pragma unroll
      for (int c = 0; c < N - 1; c++) { const T xc = bc(x, c); if (l > c && on) x -= Lr[c] * xc; }
      return x;
    };

    // ---- mjcb_control -> MjSim::controller (src/mujoco_sim/mj_sim.cpp:1055-1077) ----
    bool overridden = false;
    if (a.flags & B2F_CONTROLLER) {
      T ddq = ddq_in, dqc = dq_in;
      if (hwio) {  // MjHWInterface::write (src/mujoco_sim/mj_hw_interface.cpp:73-91), hardware joint l == dof l
        __syncthreads();   // the previous tile's joint states have left the staging area
#pragma unroll
        for (int u = 0; u < (N * EPB + BLOCK - 1) / BLOCK; u++) {
          const int idx = threadIdx.x + u * BLOCK;
          if (idx < N * EPB) { hwsh[(idx / EPB) * EPB + idx % EPB] = hwc[0][u]; hwsh[(8 + idx / EPB) * EPB + idx % EPB] = hwc[1][u]; }
        }
        __syncthreads();
        if (ctl && env < a.nenv) {
          const float hwv = hwsh[l * EPB + (threadIdx.x >> 3)], hwe = hwsh[(8 + l) * EPB + (threadIdx.x >> 3)];
          if (fabsf(hwv) > 1e-15f) dqc = (T)hwv;
          else ddq = a.hw_kp ? (T)a.hw_kp[l] * ((T)hwe - q) - (T)a.hw_kd[l] * v : (T)hwe;   // PD stage (b2_set_pd)
        }
      }
      T tau = mul_M(ddq);
      if (ctl) tau += bias;
      fa = tau;
      if (on) {
        a.qfrc_applied[at] = tau;
        if (t_abs(dqc) > Eps<T>::minval()) { v = dqc; overridden = true; }
        a.ddq[at] = 0; a.dq[at] = 0;
      }
    }
    const bool ov_team = team_any(overridden);
    // warp-uniform: recomputing with an unchanged qvel reproduces the same values bit for bit
    if ((a.flags & B2F_INVERSE) && __any_sync(FULL, overridden)) velocity_stage();
    if (on) a.qfrc_bias[at] = bias;
    if (a.flags & B2F_INVERSE) finv = mul_M(qa) + bias - pas;   // RNE(q, v, a) + armature a = M a + bias

    // ---- smooth acceleration ----
    const T fsm = pas - bias + fa;
    T W[N];
#pragma unroll
    for (int c = 0; c < N; c++) W[c] = Mr[c];
    factor(W);
    T accs = solve(on ? fsm : T(0));
    if (!on) accs = 0;
    T acc = accs, qfc = 0;

    // ---- joint limits (A.7), solved inline in acceleration space (A.8); row (joint l, side) lives on lane l ----
    bool act[2] = {false, false};
    if (on && h.has_limits && !(h.disableflags & (DSBL_LIMIT | DSBL_CONSTRAINT)) && m.i(h.o_jnt_limited, j)) {
      const T mg = m.f(h.o_jnt_margin, j);
      act[0] = q - m.f(h.o_jnt_range, 2 * j) < mg;
      act[1] = m.f(h.o_jnt_range, 2 * j + 1) - q < mg;
    }
    const unsigned tact = team_bits(act[0] || act[1]);
    const int nrow = __popc(team_bits(act[0])) + __popc(team_bits(act[1]));
    int iters = 0;
    if (__any_sync(FULL, act[0] || act[1])) {
      unsigned wj = 0;  // joints with an active side anywhere in the warp
#pragma unroll
      for (int c = 0; c < N; c++) wj |= __any_sync(FULL, (act[0] || act[1]) && l == c) ? 1u << c : 0u;
      T Bi[N];          // Bi[c] = (M^-1)[l][c] for those joints
#pragma unroll
      for (int c = 0; c < N; c++) {
        Bi[c] = 0;
        if ((wj >> c) & 1u) Bi[c] = solve(l == c ? T(1) : T(0));
      }
      T diagB = 0;
#pragma unroll
      for (int c = 0; c < N; c++) diagB = (c == l) ? Bi[c] : diagB;
      T R[2] = {1, 1}, aref[2] = {0, 0}, f[2] = {0, 0};
      if (act[0] || act[1]) {
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
          if (!act[side]) continue;
          const T js = side ? T(-1) : T(1);
          const T p = side ? m.f(h.o_jnt_range, 2 * j + 1) - q : q - m.f(h.o_jnt_range, 2 * j);
          const T imp = impedance(solimp, p, mg);
          R[side] = t_max(Eps<T>::minval(), (1 - imp) * diag / imp);
          aref[side] = -Bd * (js * v) - K * imp * (p - mg);
          if (a.flags & B2F_INVERSE) {
            const T jar = js * qa - aref[side];
            if (jar < 0) finv -= js * (-jar / R[side]);
          }
        }
      }
      if (!(a.flags & B2F_NOSOLVE)) {
        unsigned wrow = 0;  // rows active anywhere in the warp
#pragma unroll
        for (int s = 0; s < 2 * N; s++) wrow |= __any_sync(FULL, act[s & 1] && l == (s >> 1)) ? 1u << s : 0u;
        // warm start: forces implied by qacc_warmstart, kept only if their dual cost is negative
        bool warm = !(h.disableflags & DSBL_WARMSTART);
        T aw = 0;
        if (warm) {
          const T ws = on ? a.qacc_warmstart[at] : T(0);
#pragma unroll
          for (int side = 0; side < 2; side++) {
            if (!act[side]) continue;
            const T jw = (side ? -ws : ws) - aref[side];
            f[side] = jw < 0 ? -jw / R[side] : T(0);
          }
#pragma unroll
          for (int s = 0; s < 2 * N; s++) {
            if (!((wrow >> s) & 1u)) continue;
            const T fj = bc((s & 1) ? -f[1] : f[0], s >> 1);
            aw += fj * Bi[s >> 1];
          }
          T cost = 0;
#pragma unroll
          for (int s = 0; s < 2 * N; s++) {
            if (!((wrow >> s) & 1u)) continue;
            const int side = s & 1;
            const T js = side ? T(-1) : T(1);
            T term = 0;
            if (act[side] && f[side] != 0) term = f[side] * (T(0.5) * (js * aw + R[side] * f[side]) + (js * accs - aref[side]));
            cost += bc(term, s >> 1);
          }
          if (cost > 0) warm = false;   // team-uniform: every lane holds the same sum
        }
        if (warm) acc = accs + aw;
        else { f[0] = 0; f[1] = 0; }
        const T tol = m.f(h.o_opt_real, 3), scale = 1 / (m.f(h.o_opt_real, 4) * T(N));
        bool done = tact == 0;
        for (int it = 0; it < h.iterations; it++) {
          T improvement = 0;
#pragma unroll
          for (int s = 0; s < 2 * N; s++) {
            if (!((wrow >> s) & 1u)) continue;
            const int side = s & 1;
            const T js = side ? T(-1) : T(1);
            T dj = 0, ch = 0;
            if (!done && act[side] && l == (s >> 1)) {
              const T old = f[side];
              const T res = js * acc + R[side] * old - aref[side];
              const T Arr = diagB + R[side];
              const T fn = t_max(T(0), old - res / Arr);
              const T delta = fn - old;
              const T change = T(0.5) * delta * delta * Arr + delta * res;
              if (delta != 0 && !(change > T(1e-10))) { f[side] = fn; dj = delta * js; ch = change; }
            }
            dj = bc(dj, s >> 1);
            improvement -= bc(ch, s >> 1);
            acc += dj * Bi[s >> 1];
          }
          if (!done) { iters = it + 1; if (improvement * scale < tol) done = true; }
          if (__all_sync(FULL, done)) break;
        }
        qfc = f[0] - f[1];
      }
    }
    if (on && (a.flags & B2F_INVERSE)) a.qfrc_inverse[at] = finv;
    if (l == 0) {
      a.nefc[env] = nrow;
      a.solver_iter[env] = iters;
      a.status[env] = ((status_in | stat) & 7) | 8;
    }
    if (a.flags & B2F_NOSOLVE) {
      if (ov_team && on) a.qvel[at] = v;
      prefetch(tile + gridDim.x);
      continue;
    }

    // ---- qacc, warm start, mj_checkAcc, semi-implicit Euler (A.9) ----
    if (team_any(on && !(t_abs(acc) < T(1e10)))) {
      if (on) { a.qpos[at] = q0; a.qvel[at] = 0; a.qacc[at] = 0; a.qacc_warmstart[at] = 0; a.qfrc_applied[at] = 0; }
      if (l == 0) { a.time[env] = 0; a.status[env] = ((status_in | stat) & 7) | 12; }
      hw_out(q0, T(0), finv);
      acc = 0;   // keep the lanes of this team finite for the shuffles below
      // fallthrough is not wanted: the other teams of the warp still need this team's lanes in the collectives below,
      // so the reset team computes along with a zero acceleration and skips its stores
      stat |= 16;
    }
    const bool live = !(stat & 16);
    if (on && live) { a.qacc[at] = acc; a.qacc_warmstart[at] = acc; }
    // MjHWInterface::read sits between mj_step1 and mj_step2 (src/mj_main.cpp:91-108): joint states before the integration
    if (live && !(a.flags & B2F_READ_POST)) hw_out(q, v, finv);
    if (a.flags & B2F_INTEGRATE) {
      T xa = acc;
      if (h.has_damping && !(h.disableflags & DSBL_EULERDAMP)) {
#pragma unroll
        for (int c = 0; c < N; c++) W[c] = Mr[c] + ((c == l) ? a.dt() * damp : T(0));
        factor(W);
        xa = solve(on ? fsm + qfc : T(0));
      }
      if (on && live) {
        v += a.dt() * xa;
        q += a.dt() * v;
        a.qvel[at] = v;
        a.qpos[at] = q;
      }
      if (l == 0 && live) a.time[env] += a.dt();
    } else if (ov_team && on && live) {
      a.qvel[at] = v;
    }
    if (live && (a.flags & B2F_READ_POST)) hw_out(q, v, finv);
    prefetch(tile + gridDim.x);
    if (hwio) {
      __syncthreads();
#pragma unroll
      for (int u = 0; u < (N * EPB + BLOCK - 1) / BLOCK; u++) {
        const int idx = threadIdx.x + u * BLOCK, jn = idx / EPB, e2 = tile * EPB + idx % EPB;
        if (idx < N * EPB && e2 < a.nenv) {
          a.hw_pos[(long long)jn * a.nenv + e2] = hwsh[(2 * 8 + jn) * EPB + idx % EPB];
          a.hw_velo[(long long)jn * a.nenv + e2] = hwsh[(3 * 8 + jn) * EPB + idx % EPB];
          a.hw_effo[(long long)jn * a.nenv + e2] = hwsh[(4 * 8 + jn) * EPB + idx % EPB];
        }
      }
    }
  }
}

}  // namespace b2
