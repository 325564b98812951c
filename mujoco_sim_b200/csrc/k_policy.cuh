// k_policy.cuh — compile-time model policies for the smooth-dynamics kernel (see k_smooth.cuh).
//   GenericP   tree tables are read from the TMA-staged model blob in shared memory, per-environment intermediates
//              live in a strided workspace (shared memory when it fits, HBM otherwise).  Any model.
//   ChainP<N>  serial chain of N scalar joints (the Panda / UR-type arms): every tree index is a compile-time
//              constant, all loops unroll completely and the intermediates become registers (spilling to L1-backed
//              local memory where the compiler decides), leaving shared memory to the model constants only.
#pragma once
#include "k_common.cuh"

namespace b2 {

template <typename T, int N>
struct LArr {  // thread-private array; indices are compile-time constants after unrolling
  mutable T v[N > 0 ? N : 1];
  __device__ __forceinline__ T& operator[](int i) const { return v[i]; }
};

struct GenericP {
  static constexpr bool STATIC = false;
  static constexpr int UNROLL = 1;
  static constexpr int NBODY = 0, NJNT = 0, NV = 0, NM = 0, NQ = 0;
  template <typename T, int N> using Arr = SArr<T>;
#define B2_TAB(name) \
  template <typename T> static __device__ __forceinline__ int name(const MV<T>& m, int i) { return m.i(m.h->o_##name, i); }
  B2_TAB(body_parentid) B2_TAB(body_rootid) B2_TAB(body_mocapid) B2_TAB(body_jntnum) B2_TAB(body_jntadr)
  B2_TAB(body_dofnum) B2_TAB(body_dofadr) B2_TAB(body_lastdof) B2_TAB(jnt_type) B2_TAB(jnt_qposadr)
  B2_TAB(jnt_dofadr) B2_TAB(jnt_bodyid) B2_TAB(jnt_limited) B2_TAB(dof_bodyid) B2_TAB(dof_parentid) B2_TAB(dof_Madr)
  B2_TAB(dof_controlled) B2_TAB(dof_Mcnt) B2_TAB(dof_anc)
#undef B2_TAB
  // this lane's bodies (>= 1; static bodies belong to lane 0) / dofs / joints / geoms: positions [lo, hi) of the per-lane
  // lists of the model blob, ascending (parents before children).  One lane per environment: the identity.
#define B2_LANE(kind, idx, first, count, list)                                                                              \
  template <typename T> static __device__ __forceinline__ int kind##_lo(const MV<T>& m) { return m.nlanes == 1 ? (first) : m.i(m.h->o_lane_off, 9 * (idx) + m.lane); } \
  template <typename T> static __device__ __forceinline__ int kind##_hi(const MV<T>& m) { return m.nlanes == 1 ? m.h->count : m.i(m.h->o_lane_off, 9 * (idx) + m.lane + 1); } \
  template <typename T> static __device__ __forceinline__ int kind##_at(const MV<T>& m, int k) { return m.nlanes == 1 ? k : m.i(m.h->o_##list, k); }
  B2_LANE(body, 0, 1, nbody, lane_body) B2_LANE(dof, 1, 0, nv, lane_dof) B2_LANE(jnt, 2, 0, njnt, lane_jnt) B2_LANE(geom, 3, 0, ngeom, lane_geom)
#undef B2_LANE
  template <typename T> static __device__ __forceinline__ int nbody(const MV<T>& m) { return m.h->nbody; }
  template <typename T> static __device__ __forceinline__ int njnt(const MV<T>& m) { return m.h->njnt; }
  template <typename T> static __device__ __forceinline__ int nv(const MV<T>& m) { return m.h->nv; }
  template <typename T> static __device__ __forceinline__ int nM(const MV<T>& m) { return m.h->nM; }
  template <typename T> static __device__ __forceinline__ int nq(const MV<T>& m) { return m.h->nq; }
};

template <int N>
struct ChainP {  // world -> body 1 -> ... -> body N, body b carries scalar joint b - 1 = dof b - 1
  static constexpr bool STATIC = true;
  static constexpr int UNROLL = 64;
  static constexpr int NBODY = N + 1, NJNT = N, NV = N, NM = N * (N + 1) / 2, NQ = N;
  template <typename T, int K> using Arr = LArr<T, K>;
  template <typename T> static __device__ __forceinline__ int body_parentid(const MV<T>&, int b) { return b > 0 ? b - 1 : 0; }
  template <typename T> static __device__ __forceinline__ int body_rootid(const MV<T>&, int b) { return b > 0 ? 1 : 0; }
  template <typename T> static __device__ __forceinline__ int body_mocapid(const MV<T>&, int) { return -1; }
  template <typename T> static __device__ __forceinline__ int body_jntnum(const MV<T>&, int b) { return b > 0 ? 1 : 0; }
  template <typename T> static __device__ __forceinline__ int body_jntadr(const MV<T>&, int b) { return b - 1; }
  template <typename T> static __device__ __forceinline__ int body_dofnum(const MV<T>&, int b) { return b > 0 ? 1 : 0; }
  template <typename T> static __device__ __forceinline__ int body_dofadr(const MV<T>&, int b) { return b - 1; }
  template <typename T> static __device__ __forceinline__ int body_lastdof(const MV<T>&, int b) { return b - 1; }
  template <typename T> static __device__ __forceinline__ int jnt_type(const MV<T>& m, int j) { return m.i(m.h->o_jnt_type, j); }
  template <typename T> static __device__ __forceinline__ int jnt_limited(const MV<T>& m, int j) { return m.i(m.h->o_jnt_limited, j); }
  template <typename T> static __device__ __forceinline__ int jnt_qposadr(const MV<T>&, int j) { return j; }
  template <typename T> static __device__ __forceinline__ int jnt_dofadr(const MV<T>&, int j) { return j; }
  template <typename T> static __device__ __forceinline__ int jnt_bodyid(const MV<T>&, int j) { return j + 1; }
  template <typename T> static __device__ __forceinline__ int dof_bodyid(const MV<T>&, int i) { return i + 1; }
  template <typename T> static __device__ __forceinline__ int dof_parentid(const MV<T>&, int i) { return i - 1; }
  template <typename T> static __device__ __forceinline__ int dof_Madr(const MV<T>&, int i) { return i * (i + 1) / 2; }
  template <typename T> static __device__ __forceinline__ int dof_Mcnt(const MV<T>&, int i) { return i + 1; }
  template <typename T> static __device__ __forceinline__ int dof_controlled(const MV<T>& m, int i) { return m.i(m.h->o_dof_controlled, i); }
  template <typename T> static __device__ __forceinline__ int body_lo(const MV<T>&) { return 1; }
  template <typename T> static __device__ __forceinline__ int body_hi(const MV<T>&) { return NBODY; }
  template <typename T> static __device__ __forceinline__ int body_at(const MV<T>&, int k) { return k; }
  template <typename T> static __device__ __forceinline__ int dof_lo(const MV<T>&) { return 0; }
  template <typename T> static __device__ __forceinline__ int dof_hi(const MV<T>&) { return NV; }
  template <typename T> static __device__ __forceinline__ int dof_at(const MV<T>&, int k) { return k; }
  template <typename T> static __device__ __forceinline__ int jnt_lo(const MV<T>&) { return 0; }
  template <typename T> static __device__ __forceinline__ int jnt_hi(const MV<T>&) { return NJNT; }
  template <typename T> static __device__ __forceinline__ int jnt_at(const MV<T>&, int k) { return k; }
  template <typename T> static __device__ __forceinline__ int geom_lo(const MV<T>&) { return 0; }
  template <typename T> static __device__ __forceinline__ int geom_hi(const MV<T>& m) { return m.h->ngeom; }
  template <typename T> static __device__ __forceinline__ int geom_at(const MV<T>&, int k) { return k; }
  template <typename T> static __device__ __forceinline__ int nbody(const MV<T>&) { return NBODY; }
  template <typename T> static __device__ __forceinline__ int njnt(const MV<T>&) { return NJNT; }
  template <typename T> static __device__ __forceinline__ int nv(const MV<T>&) { return NV; }
  template <typename T> static __device__ __forceinline__ int nM(const MV<T>&) { return NM; }
  template <typename T> static __device__ __forceinline__ int nq(const MV<T>&) { return NQ; }
};

}  // namespace b2
