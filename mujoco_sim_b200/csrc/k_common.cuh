// k_common.cuh — shared device-side plumbing: strided per-environment arrays, the shared-memory model view,
// TMA staging of the model blob, small vector / quaternion / spatial-algebra helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmodel.h"

namespace b2 {

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { DSBL_CONSTRAINT = 1, DSBL_EQUALITY = 2, DSBL_FRICTIONLOSS = 4, DSBL_LIMIT = 8, DSBL_CONTACT = 16, DSBL_PASSIVE = 32,
       DSBL_GRAVITY = 64, DSBL_WARMSTART = 256, DSBL_REFSAFE = 2048, DSBL_EULERDAMP = 16384 };

// Per-environment array element i lives at p[i * s]: s = padded environment count for HBM-resident SoA arrays
// (a warp touches 32 consecutive environments = one 128 B line per element), s = blockDim.x for the shared-memory
// workspace (bank-conflict free: lane is the fastest index).
template <typename T>
struct SArr {
  T* p;
  long long s;
  __device__ __forceinline__ T& operator[](int i) const { return p[(long long)i * s]; }
  __device__ __forceinline__ SArr<T> at(int i) const { return SArr<T>{p + (long long)i * s, s}; }
};

// Model view over the staged blob.
template <typename T>
struct MV {
  const DModel* h;
  const uint32_t* w;
  // tree-parallel kernels (k_smooth / k_integrate with several lanes per environment): this thread works on the
  // kinematic trees t with t % nlanes == lane (static bodies: lane 0); nlanes == 1 is the thread-per-environment form
  int lane = 0, nlanes = 1;
  __device__ __forceinline__ int i(int off, int k) const { return (int)w[off + k]; }
  __device__ __forceinline__ T f(int off, int k) const { return reinterpret_cast<const T*>(w + off)[k]; }
  __device__ __forceinline__ const T* fp(int off) const { return reinterpret_cast<const T*>(w + off); }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One elected thread issues a single TMA bulk copy (cp.async.bulk, SASS: UBLKCP) of the model blob HBM -> shared
// memory and every thread of the CTA waits on the mbarrier transaction count.
// Split form: stage_model_issue starts the copy, stage_model_wait blocks until it has landed, so that a kernel can put
// its first global loads in flight in between.
__device__ __forceinline__ void stage_model_issue(uint32_t* dst, const uint32_t* src, int nwords, uint64_t* bar) {
  const uint32_t bytes = (uint32_t)nwords * 4u;
  const uint32_t bar_a = smem_u32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    // bulk copies are limited in size per instruction only by the smem capacity; the blob is < 64 KB
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(bar_a)
                 : "memory");
  }
}
__device__ __forceinline__ void stage_model_wait(uint64_t* bar) {
  const uint32_t bar_a = smem_u32(bar);
  __syncthreads();   // the barrier initialisation by thread 0 is visible to every waiter
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar_a)
        : "memory");
  }
}
__device__ __forceinline__ void stage_model(uint32_t* dst, const uint32_t* src, int nwords, uint64_t* bar) {
  stage_model_issue(dst, src, nwords, bar);
  stage_model_wait(bar);
}

// The CTA's threads put rows [0, nrows) of an [element][env] array in flight towards L1 for its `nenvs` consecutive
// environments starting at env0 (one prefetch per 128-byte line): the thread-per-environment stages walk these arrays
// element by element inside dependent chains, so without this every first touch is a serialised L2 / HBM round trip.
template <typename T>
__device__ __forceinline__ void prefetch_rows(const T* base, int nrows, long long S, int env0, int nenvs) {
  if (!base) return;
  constexpr int PER_LINE = 128 / (int)sizeof(T);
  const int nlines = (nenvs + PER_LINE - 1) / PER_LINE;
  for (int idx = threadIdx.x; idx < nrows * nlines; idx += blockDim.x) {
    const int r = idx / nlines, ln = idx % nlines;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (long long)r * S + env0 + ln * PER_LINE));
  }
}

// ---------- small math (registers) ----------
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < 0 ? -x : x; }
// explicit fused multiply-add / unfused product: where two code paths must round identically (shard invariance)
__device__ __forceinline__ float t_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double t_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float t_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double t_mul(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T t_min(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T t_max(T a, T b) { return a > b ? a : b; }
__device__ __forceinline__ void t_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void t_sincos(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ float t_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double t_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float t_pow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double t_pow(double x, double y) { return pow(x, y); }

template <typename T> struct Eps;
template <> struct Eps<float> { static __device__ __forceinline__ float minval() { return 1e-15f; } };
template <> struct Eps<double> { static __device__ __forceinline__ double minval() { return 1e-15; } };

template <typename T> __device__ __forceinline__ T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> __device__ __forceinline__ void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ T norm3(const T* a) { return t_sqrt(dot3(a, a)); }
template <typename T> __device__ __forceinline__ T normalize3(T* a) {
  T n = norm3(a);
  if (n < Eps<T>::minval()) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { T inv = T(1) / n; a[0] *= inv; a[1] *= inv; a[2] *= inv; }
  return n;
}
template <typename T> __device__ __forceinline__ void normalize4(T* q) {
  T n = t_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < Eps<T>::minval()) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else { T inv = T(1) / n; q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv; }
}
template <typename T> __device__ __forceinline__ void mul_quat(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <typename T> __device__ __forceinline__ void quat2mat(T* m, const T* q) {
  T q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  T q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = 2 * (q12 - q03);       m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03);       m[4] = q00 - q11 + q22 - q33; m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02);       m[7] = 2 * (q23 + q01);       m[8] = q00 - q11 - q22 + q33;
}
template <typename T> __device__ __forceinline__ void mat_vec3(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  T y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  T z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ void matT_vec3(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  T y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  T z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ void rot_vec_quat(T* r, const T* v, const T* q) {
  // r = v + 2 w (u x v) + 2 u x (u x v), u = q[1:4]
  T t[3], u[3] = {q[1], q[2], q[3]};
  cross3(t, u, v);
  t[0] *= 2; t[1] *= 2; t[2] *= 2;
  T c[3];
  cross3(c, u, t);
  r[0] = v[0] + q[0] * t[0] + c[0];
  r[1] = v[1] + q[0] * t[1] + c[1];
  r[2] = v[2] + q[0] * t[2] + c[2];
}
template <typename T> __device__ __forceinline__ void axis_angle2quat(T* q, const T* axis, T angle) {
  T s, c;
  t_sincos(angle * T(0.5), &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
// q <- normalize(q) * quat(vel * scale), vel in the local frame
template <typename T> __device__ __forceinline__ void quat_integrate(T* q, const T* vel, T scale) {
  normalize4(q);
  T ax[3] = {vel[0], vel[1], vel[2]};
  T speed = norm3(ax);
  if (speed < Eps<T>::minval()) return;
  T inv = T(1) / speed;
  ax[0] *= inv; ax[1] *= inv; ax[2] *= inv;
  T r[4], o[4];
  axis_angle2quat(r, ax, scale * speed);
  mul_quat(o, q, r);
  q[0] = o[0]; q[1] = o[1]; q[2] = o[2]; q[3] = o[3];
}
// rotation vector taking qb to qa, in the frame of qb
template <typename T> __device__ __forceinline__ void sub_quat(T* res, const T* qa, const T* qb) {
  T qn[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, qd[4];
  mul_quat(qd, qn, qa);
  T ax[3] = {qd[1], qd[2], qd[3]};
  T s = norm3(ax);
  if (s < Eps<T>::minval()) { res[0] = res[1] = res[2] = 0; return; }
  T speed = 2 * t_atan2(s, qd[0]);
  const T pi = T(3.14159265358979323846);
  if (speed > pi) speed -= 2 * pi;
  T k = speed / s;
  res[0] = ax[0] * k; res[1] = ax[1] * k; res[2] = ax[2] * k;
}
// spatial inertia (10 numbers) times motion vector [ang; lin]
template <typename T> __device__ __forceinline__ void mul_inert_vec(T* res, const T* i, const T* v) {
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
template <typename T> __device__ __forceinline__ void cross_motion(T* res, const T* vel, const T* v) {
  T a[3], b[3], c[3];
  cross3(a, vel, v);
  cross3(b, vel, v + 3);
  cross3(c, vel + 3, v);
  res[0] = a[0]; res[1] = a[1]; res[2] = a[2];
  res[3] = b[0] + c[0]; res[4] = b[1] + c[1]; res[5] = b[2] + c[2];
}
template <typename T> __device__ __forceinline__ void cross_force(T* res, const T* vel, const T* f) {
  T a[3], b[3], c[3];
  cross3(a, vel, f);
  cross3(b, vel + 3, f + 3);
  cross3(c, vel, f + 3);
  res[0] = a[0] + b[0]; res[1] = a[1] + b[1]; res[2] = a[2] + b[2];
  res[3] = c[0]; res[4] = c[1]; res[5] = c[2];
}

template <typename T, int N, typename A> __device__ __forceinline__ void ld(T* r, const A& a, int base) {
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = a[base + k];
}
template <typename T, int N, typename A> __device__ __forceinline__ void st(const A& a, int base, const T* r) {
#pragma unroll
  for (int k = 0; k < N; k++) a[base + k] = r[k];
}
template <typename T, int N> __device__ __forceinline__ void ldm(T* r, const MV<T>& m, int off, int base) {
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = m.f(off, base + k);
}

}  // namespace b2
