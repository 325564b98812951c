// k_collide.cuh — broad + narrow phase (rows s4 / s5 of SURVEY.md section 8a'): primitive pair functions and the
// general convex path (MPR over support functions: cylinder, ellipsoid, mesh and every pair without a primitive).
// One thread per environment walks the static candidate-pair list in order, so the contact order is the canonical
// (pair index, emission index) order by construction and geom ids are reproducible bit for bit.
// Conventions (MuJoCo docs, mjContact): dist < 0 is penetration, pos is the midpoint, frame row 0 is the normal from
// geom1 to geom2 where geom1 has the lower geom type.
#pragma once
#include "k_args.h"
#include "k_common.cuh"

namespace b2 {

enum { GEOM_PLANE = 0, GEOM_HFIELD = 1, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_ELLIPSOID = 4, GEOM_CYLINDER = 5,
       GEOM_BOX = 6, GEOM_MESH = 7 };
#define B2_MAXCONPAIR 8

template <typename T>
struct RawCon { T dist, pos[3], n[3], tan[3]; };

template <typename T>
struct GeomW {
  T pos[3], mat[9], size[3];
  int type;         // geom type (general convex path)
  const T* vert;    // mesh vertices in the geom frame (HBM, behind the staged part of the model blob)
  int nvert;
};

template <typename T> __device__ __forceinline__ void mcol(T* r, const T* mat, int k) { r[0] = mat[k]; r[1] = mat[3 + k]; r[2] = mat[6 + k]; }

template <typename T>
__device__ __forceinline__ void set_raw(RawCon<T>& c, T dist, const T* pos, const T* n) {
  c.dist = dist;
  for (int k = 0; k < 3; k++) { c.pos[k] = pos[k]; c.n[k] = n[k]; c.tan[k] = 0; }
}

template <typename T>
__device__ int d_sphere_sphere(RawCon<T>* out, const T* c1, T r1, const T* c2, T r2, T margin) {
  T dif[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
  const T cd = norm3(dif), dist = cd - r1 - r2;
  if (dist > margin) return 0;
  T n[3] = {1, 0, 0};
  if (cd >= Eps<T>::minval()) { const T inv = T(1) / cd; n[0] = dif[0] * inv; n[1] = dif[1] * inv; n[2] = dif[2] * inv; }
  const T s = r1 + T(0.5) * dist;
  T p[3] = {c1[0] + s * n[0], c1[1] + s * n[1], c1[2] + s * n[2]};
  set_raw(out[0], dist, p, n);
  return 1;
}

template <typename T>
__device__ int d_plane_sphere(RawCon<T>* out, const T* ppos, const T* n, const T* c, T r, T margin) {
  const T dif[3] = {c[0] - ppos[0], c[1] - ppos[1], c[2] - ppos[2]};
  const T dist = dot3(dif, n) - r;
  if (dist > margin) return 0;
  const T s = -(r + T(0.5) * dist);
  T p[3] = {c[0] + s * n[0], c[1] + s * n[1], c[2] + s * n[2]};
  set_raw(out[0], dist, p, n);
  return 1;
}

template <typename T>
__device__ int d_plane_capsule(RawCon<T>* out, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  T n[3], ax[3];
  mcol(n, a.mat, 2);
  mcol(ax, b.mat, 2);
  int cnt = 0;
  for (int s = 1; s >= -1; s -= 2) {
    const T hl = s * b.size[1];
    T e[3] = {b.pos[0] + hl * ax[0], b.pos[1] + hl * ax[1], b.pos[2] + hl * ax[2]};
    cnt += d_plane_sphere(out + cnt, a.pos, n, e, b.size[0], margin);
  }
  for (int i = 0; i < cnt; i++) { out[i].tan[0] = ax[0]; out[i].tan[1] = ax[1]; out[i].tan[2] = ax[2]; }
  return cnt;
}

template <typename T>
__device__ int d_plane_cylinder(RawCon<T>* out, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  T n[3], ax[3];
  mcol(n, a.mat, 2);
  mcol(ax, b.mat, 2);
  T prjaxis = dot3(n, ax);
  if (prjaxis > 0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; prjaxis = -prjaxis; }
  const T dif[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
  const T dist0 = dot3(dif, n);
  T vec[3] = {ax[0] * prjaxis - n[0], ax[1] * prjaxis - n[1], ax[2] * prjaxis - n[2]};
  const T len = norm3(vec);
  if (len >= T(1e-12)) { const T s = b.size[0] / len; vec[0] *= s; vec[1] *= s; vec[2] *= s; }
  else { mcol(vec, b.mat, 0); vec[0] *= b.size[0]; vec[1] *= b.size[0]; vec[2] *= b.size[0]; }
  const T prjvec = dot3(vec, n);
  const T axs[3] = {ax[0] * b.size[1], ax[1] * b.size[1], ax[2] * b.size[1]};
  prjaxis *= b.size[1];
  int cnt = 0;
  T p[3];
  const T d1 = dist0 + prjaxis + prjvec;
  if (d1 > margin) return 0;
  for (int k = 0; k < 3; k++) p[k] = b.pos[k] + vec[k] + axs[k] - n[k] * d1 * T(0.5);
  set_raw(out[cnt++], d1, p, n);
  const T d2 = dist0 - prjaxis + prjvec;
  if (d2 <= margin) {
    for (int k = 0; k < 3; k++) p[k] = b.pos[k] + vec[k] - axs[k] - n[k] * d2 * T(0.5);
    set_raw(out[cnt++], d2, p, n);
  }
  const T d3 = dist0 + prjaxis - T(0.5) * prjvec;
  if (d3 <= margin) {
    T side[3];
    cross3(side, vec, ax);
    normalize3(side);
    const T sc = b.size[0] * T(0.8660254037844386);
    for (int s = 1; s >= -1; s -= 2) {
      for (int k = 0; k < 3; k++) p[k] = b.pos[k] + s * sc * side[k] + axs[k] - T(0.5) * vec[k] - n[k] * d3 * T(0.5);
      set_raw(out[cnt++], d3, p, n);
    }
  }
  return cnt;
}

template <typename T>
__device__ int d_plane_box(RawCon<T>* out, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  T n[3];
  mcol(n, a.mat, 2);
  const T dif[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
  const T dist = dot3(dif, n);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    const T v[3] = {(i & 1 ? b.size[0] : -b.size[0]), (i & 2 ? b.size[1] : -b.size[1]), (i & 4 ? b.size[2] : -b.size[2])};
    T corner[3];
    mat_vec3(corner, b.mat, v);
    const T ldist = dot3(n, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    const T cd = dist + ldist;
    T p[3];
    for (int k = 0; k < 3; k++) p[k] = b.pos[k] + corner[k] - n[k] * cd * T(0.5);
    set_raw(out[cnt], cd, p, n);
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

// sphere against a convex solid, given (in the solid's frame) the sphere centre c, the solid's nearest point p and,
// for a centre inside the solid, the outward normal / depth of the shallowest exit
template <typename T>
__device__ int d_sphere_solid(RawCon<T>* out, const T* spos, T r, const T* smat, const T* c, const T* p, const T* nin,
                              T depth_in, T margin) {
  T dl[3] = {c[0] - p[0], c[1] - p[1], c[2] - p[2]};
  const T dn = norm3(dl);
  T nloc[3], dist;
  if (dn > T(1e-12)) {
    dist = dn - r;
    const T inv = T(1) / dn;
    nloc[0] = dl[0] * inv; nloc[1] = dl[1] * inv; nloc[2] = dl[2] * inv;
  } else {
    dist = -depth_in - r;
    nloc[0] = nin[0]; nloc[1] = nin[1]; nloc[2] = nin[2];
  }
  if (dist > margin) return 0;
  T nw[3];
  mat_vec3(nw, smat, nloc);
  nw[0] = -nw[0]; nw[1] = -nw[1]; nw[2] = -nw[2];
  const T s = r + T(0.5) * dist;
  T pw[3] = {spos[0] + s * nw[0], spos[1] + s * nw[1], spos[2] + s * nw[2]};
  set_raw(out[0], dist, pw, nw);
  return 1;
}

template <typename T>
__device__ int d_sphere_box(RawCon<T>* out, const T* spos, T r, const GeomW<T>& b, T margin) {
  const T dif[3] = {spos[0] - b.pos[0], spos[1] - b.pos[1], spos[2] - b.pos[2]};
  T c[3], p[3], nin[3] = {0, 0, 0};
  matT_vec3(c, b.mat, dif);
  int kmin = 0;
  T dmin = T(1e30);
  for (int k = 0; k < 3; k++) {
    p[k] = t_min(b.size[k], t_max(-b.size[k], c[k]));
    const T ex = b.size[k] - t_abs(c[k]);
    if (ex < dmin) { dmin = ex; kmin = k; }
  }
  const T sg = c[kmin] >= 0 ? T(1) : T(-1);
  if (kmin == 0) nin[0] = sg; else if (kmin == 1) nin[1] = sg; else nin[2] = sg;
  return d_sphere_solid(out, spos, r, b.mat, c, p, nin, dmin, margin);
}

template <typename T>
__device__ int d_sphere_cylinder(RawCon<T>* out, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  const T dif[3] = {a.pos[0] - b.pos[0], a.pos[1] - b.pos[1], a.pos[2] - b.pos[2]};
  T c[3], p[3], nin[3] = {0, 0, 0};
  matT_vec3(c, b.mat, dif);
  const T R = b.size[0], H = b.size[1];
  const T rho = t_sqrt(c[0] * c[0] + c[1] * c[1]);
  const T s = rho > R ? R / rho : T(1);
  p[0] = c[0] * s; p[1] = c[1] * s;
  p[2] = t_min(H, t_max(-H, c[2]));
  const T ex_r = R - rho, ex_z = H - t_abs(c[2]);
  T depth;
  if (ex_r < ex_z) {
    depth = ex_r;
    if (rho > T(1e-12)) { nin[0] = c[0] / rho; nin[1] = c[1] / rho; } else nin[0] = 1;
  } else {
    depth = ex_z;
    nin[2] = c[2] >= 0 ? T(1) : T(-1);
  }
  return d_sphere_solid(out, a.pos, a.size[0], b.mat, c, p, nin, depth, margin);
}

template <typename T>
__device__ void d_segment_segment(const T* c1, const T* a1, T h1, const T* c2, const T* a2, T h2, T& t1, T& t2, bool& par) {
  const T dif[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
  const T b = dot3(a1, a2), u = -dot3(a1, dif), v = dot3(a2, dif);
  const T det = 1 - b * b;
  par = det < T(1e-10);
  t1 = par ? T(0) : (u + b * v) / det;
  t1 = t_min(h1, t_max(-h1, t1));
  t2 = t_min(h2, t_max(-h2, v + b * t1));
  t1 = t_min(h1, t_max(-h1, u + b * t2));
}

template <typename T>
__device__ int d_capsule_capsule(RawCon<T>* out, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  T a1[3], a2[3], t1, t2;
  mcol(a1, a.mat, 2);
  mcol(a2, b.mat, 2);
  bool par;
  d_segment_segment(a.pos, a1, a.size[1], b.pos, a2, b.size[1], t1, t2, par);
  const T dif[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
  const T mid = dot3(dif, a1);
  const T lo = t_max(-a.size[1], mid - b.size[1]), hi = t_min(a.size[1], mid + b.size[1]);
  if (!par || lo >= hi) {
    T p1[3], p2[3];
    for (int k = 0; k < 3; k++) { p1[k] = a.pos[k] + t1 * a1[k]; p2[k] = b.pos[k] + t2 * a2[k]; }
    return d_sphere_sphere(out, p1, a.size[0], p2, b.size[0], margin);
  }
  const T s = dot3(a1, a2) >= 0 ? T(1) : T(-1);
  int cnt = 0;
  for (int e = 0; e < 2; e++) {
    const T x = e ? hi : lo, y = s * (x - mid);
    T p1[3], p2[3];
    for (int k = 0; k < 3; k++) { p1[k] = a.pos[k] + x * a1[k]; p2[k] = b.pos[k] + y * a2[k]; }
    cnt += d_sphere_sphere(out + cnt, p1, a.size[0], p2, b.size[0], margin);
  }
  return cnt;
}

// signed distance from a local point to the box surface (negative inside)
template <typename T>
__device__ __forceinline__ T box_sdist(const T* s, const T* c) {
  T d2 = 0, ex = T(1e30);
  for (int k = 0; k < 3; k++) {
    const T ac = t_abs(c[k]);
    const T o = ac - s[k];
    if (o > 0) d2 += o * o;
    ex = t_min(ex, -o);
  }
  return d2 > 0 ? t_sqrt(d2) : -ex;
}

template <typename T>
__device__ int d_capsule_box(RawCon<T>* out, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  T axw[3], ax[3], c0[3];
  mcol(axw, a.mat, 2);
  const T dif[3] = {a.pos[0] - b.pos[0], a.pos[1] - b.pos[1], a.pos[2] - b.pos[2]};
  matT_vec3(c0, b.mat, dif);
  matT_vec3(ax, b.mat, axw);
  const T hl = a.size[1];
  const T gr = T(0.6180339887498949);
  T lo = -hl, hi = hl;
  T x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo);
  T c[3];
  for (int k = 0; k < 3; k++) c[k] = c0[k] + x1 * ax[k];
  T f1 = box_sdist(b.size, c);
  for (int k = 0; k < 3; k++) c[k] = c0[k] + x2 * ax[k];
  T f2 = box_sdist(b.size, c);
  for (int it = 0; it < 40; it++) {
    if (f1 <= f2) {
      hi = x2; x2 = x1; f2 = f1; x1 = hi - gr * (hi - lo);
      for (int k = 0; k < 3; k++) c[k] = c0[k] + x1 * ax[k];
      f1 = box_sdist(b.size, c);
    } else {
      lo = x1; x1 = x2; f1 = f2; x2 = lo + gr * (hi - lo);
      for (int k = 0; k < 3; k++) c[k] = c0[k] + x2 * ax[k];
      f2 = box_sdist(b.size, c);
    }
  }
  const T tbest = T(0.5) * (lo + hi);
  int cnt = 0;
  for (int i = 0; i < 3; i++) {
    const T ti = i == 0 ? tbest : (i == 1 ? -hl : hl);
    if (i > 0 && t_abs(ti - tbest) < T(0.1) * hl + T(1e-9)) continue;
    T cw[3] = {a.pos[0] + ti * axw[0], a.pos[1] + ti * axw[1], a.pos[2] + ti * axw[2]};
    cnt += d_sphere_box(out + cnt, cw, a.size[0], b, margin);
  }
  return cnt;
}

// box-box: separating-axis search over the 15 candidate axes, then either clip the incident face against the
// reference face (up to 8 points) or take the closest points of the two supporting edges (1 point)
template <typename T>
__device__ int d_box_box(RawCon<T>* out, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  T A[3][3], B[3][3];
  for (int k = 0; k < 3; k++) { mcol(A[k], a.mat, k); mcol(B[k], b.mat, k); }
  const T dif[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
  T Rabs[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rabs[i][j] = t_abs(dot3(A[i], B[j])) + T(1e-12);
  T best = T(-1e30);
  int code = -1;
  T axis[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) {
    const T rb = b.size[0] * Rabs[i][0] + b.size[1] * Rabs[i][1] + b.size[2] * Rabs[i][2];
    const T t = dot3(dif, A[i]), sep = t_abs(t) - a.size[i] - rb;
    if (sep > margin) return 0;
    if (sep > best + (i ? T(1e-6) : T(0))) { best = sep; code = i; const T s = t >= 0 ? T(1) : T(-1); axis[0] = s * A[i][0]; axis[1] = s * A[i][1]; axis[2] = s * A[i][2]; }
  }
  for (int j = 0; j < 3; j++) {
    const T ra = a.size[0] * Rabs[0][j] + a.size[1] * Rabs[1][j] + a.size[2] * Rabs[2][j];
    const T t = dot3(dif, B[j]), sep = t_abs(t) - ra - b.size[j];
    if (sep > margin) return 0;
    if (sep > best + T(1e-6)) { best = sep; code = 3 + j; const T s = t >= 0 ? T(1) : T(-1); axis[0] = s * B[j][0]; axis[1] = s * B[j][1]; axis[2] = s * B[j][2]; }
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      T L[3];
      cross3(L, A[i], B[j]);
      const T ln = norm3(L);
      if (ln < T(1e-6)) continue;
      const T inv = T(1) / ln;
      L[0] *= inv; L[1] *= inv; L[2] *= inv;
      T ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) { ra += a.size[k] * t_abs(dot3(A[k], L)); rb += b.size[k] * t_abs(dot3(B[k], L)); }
      const T t = dot3(dif, L), sep = t_abs(t) - ra - rb;
      if (sep > margin) return 0;
      if (sep > best + T(1e-6)) { best = sep; code = 6 + 3 * i + j; const T s = t >= 0 ? T(1) : T(-1); axis[0] = s * L[0]; axis[1] = s * L[1]; axis[2] = s * L[2]; }
    }
  if (code < 0) return 0;

  if (code >= 6) {
    const int i = (code - 6) / 3, j = (code - 6) % 3;
    T pa[3] = {a.pos[0], a.pos[1], a.pos[2]}, pb[3] = {b.pos[0], b.pos[1], b.pos[2]};
    for (int k = 0; k < 3; k++) {
      if (k != i) { const T s = (dot3(A[k], axis) >= 0 ? T(1) : T(-1)) * a.size[k]; pa[0] += s * A[k][0]; pa[1] += s * A[k][1]; pa[2] += s * A[k][2]; }
      if (k != j) { const T s = (dot3(B[k], axis) >= 0 ? T(-1) : T(1)) * b.size[k]; pb[0] += s * B[k][0]; pb[1] += s * B[k][1]; pb[2] += s * B[k][2]; }
    }
    T t1, t2;
    bool par;
    d_segment_segment(pa, A[i], a.size[i], pb, B[j], b.size[j], t1, t2, par);
    T p[3];
    for (int k = 0; k < 3; k++) p[k] = T(0.5) * (pa[k] + t1 * A[i][k] + pb[k] + t2 * B[j][k]);
    set_raw(out[0], best, p, axis);
    return 1;
  }

  const bool ref_a = code < 3;
  const GeomW<T>& rf = ref_a ? a : b;
  const GeomW<T>& in = ref_a ? b : a;
  T (*RA)[3] = ref_a ? A : B;
  T (*IA)[3] = ref_a ? B : A;
  const int ri = ref_a ? code : code - 3;
  const T sgn = ref_a ? T(1) : T(-1);
  const T nref[3] = {sgn * axis[0], sgn * axis[1], sgn * axis[2]};
  int ii = 0;
  T mind = T(1e30), isgn = 1;
  for (int k = 0; k < 3; k++) {
    const T dd = dot3(IA[k], nref);
    if (-t_abs(dd) < mind) { mind = -t_abs(dd); ii = k; isgn = dd > 0 ? T(-1) : T(1); }
  }
  const int iu = (ii + 1) % 3, iv = (ii + 2) % 3;
  T poly[8][3], tmp[8][3];
  int np = 4;
  for (int c = 0; c < 4; c++) {
    const T su = (c == 0 || c == 3) ? T(-1) : T(1), sv = c < 2 ? T(-1) : T(1);
    for (int k = 0; k < 3; k++)
      poly[c][k] = in.pos[k] + isgn * in.size[ii] * IA[ii][k] + su * in.size[iu] * IA[iu][k] + sv * in.size[iv] * IA[iv][k];
  }
  const int ru = (ri + 1) % 3, rv = (ri + 2) % 3;
  for (int s = 0; s < 4 && np > 0; s++) {
    const int sa = s < 2 ? ru : rv;
    const T ss = (s & 1) ? T(-1) : T(1);
    const T* ax = RA[sa];
    const T lim = rf.size[sa];
    int nn = 0;
    for (int c = 0; c < np; c++) {
      const T* p0 = poly[c];
      const T* p1 = poly[(c + 1) % np];
      const T d0[3] = {p0[0] - rf.pos[0], p0[1] - rf.pos[1], p0[2] - rf.pos[2]};
      const T d1[3] = {p1[0] - rf.pos[0], p1[1] - rf.pos[1], p1[2] - rf.pos[2]};
      const T e0 = ss * dot3(d0, ax) - lim, e1 = ss * dot3(d1, ax) - lim;
      if (e0 <= 0 && nn < 8) { tmp[nn][0] = p0[0]; tmp[nn][1] = p0[1]; tmp[nn][2] = p0[2]; nn++; }
      if (((e0 < 0 && e1 > 0) || (e0 > 0 && e1 < 0)) && nn < 8) {
        const T t = e0 / (e0 - e1);
        for (int k = 0; k < 3; k++) tmp[nn][k] = p0[k] + t * (p1[k] - p0[k]);
        nn++;
      }
    }
    np = nn;
    for (int c = 0; c < np; c++) { poly[c][0] = tmp[c][0]; poly[c][1] = tmp[c][1]; poly[c][2] = tmp[c][2]; }
  }
  int cnt = 0;
  for (int c = 0; c < np && cnt < B2_MAXCONPAIR; c++) {
    const T dv[3] = {poly[c][0] - rf.pos[0], poly[c][1] - rf.pos[1], poly[c][2] - rf.pos[2]};
    const T depth = dot3(dv, nref) - rf.size[ri];
    if (depth > margin) continue;
    T p[3] = {poly[c][0] - T(0.5) * depth * nref[0], poly[c][1] - T(0.5) * depth * nref[1], poly[c][2] - T(0.5) * depth * nref[2]};
    set_raw(out[cnt++], depth, p, axis);
  }
  return cnt;
}

// ---- general convex pairs: Minkowski Portal Refinement (what MuJoCo's mjc_Convex runs through libccd) ----
// One contact per pair: dist = margin - depth, normal = portal direction (geom1 -> geom2), pos = midpoint of the
// witness points blended by the barycentric coordinates of the origin ray.  mpr_tolerance 1e-6, mpr_iterations 50.
template <typename T> struct CcdEps;
template <> struct CcdEps<float> { static __device__ __forceinline__ float v() { return 1.1920929e-7f; } };
template <> struct CcdEps<double> { static __device__ __forceinline__ double v() { return 2.220446049250313e-16; } };
template <typename T> __device__ __forceinline__ bool ccd_zero(T x) { return t_abs(x) < CcdEps<T>::v(); }
template <typename T> __device__ __forceinline__ bool ccd_eq(T a, T b) {
  const T ab = t_abs(a - b);
  if (ab < CcdEps<T>::v()) return true;
  return ab < CcdEps<T>::v() * t_max(t_abs(a), t_abs(b));
}
template <typename T> __device__ __forceinline__ T t_sgn(T x) { return x > 0 ? T(1) : (x < 0 ? T(-1) : T(0)); }
template <typename T> __device__ __forceinline__ void ccd_normalize(T* v) { const T inv = T(1) / norm3(v); v[0] *= inv; v[1] *= inv; v[2] *= inv; }

// geometry as the convex routine sees it: reals of type T, mesh vertices of the batch's storage type V
template <typename T, typename V>
struct GeomC {
  T pos[3], mat[9], size[3];
  int type;
  const V* vert;
  int nvert;
};

template <typename T, typename G>
__device__ void d_support_geom(T* res, const G& g, const T* dir, T margin) {
  T l[3], r[3] = {0, 0, 0};
  for (int k = 0; k < 3; k++) l[k] = g.mat[k] * dir[0] + g.mat[3 + k] * dir[1] + g.mat[6 + k] * dir[2];
  switch (g.type) {
    case GEOM_SPHERE: for (int k = 0; k < 3; k++) r[k] = l[k] * g.size[0]; break;
    case GEOM_CAPSULE:
      for (int k = 0; k < 3; k++) r[k] = l[k] * g.size[0];
      r[2] += t_sgn(l[2]) * g.size[1];
      break;
    case GEOM_ELLIPSOID: {
      const T t[3] = {l[0] * g.size[0], l[1] * g.size[1], l[2] * g.size[2]};
      const T n = norm3(t);
      if (n >= Eps<T>::minval()) for (int k = 0; k < 3; k++) r[k] = g.size[k] * t[k] / n;
      break;
    }
    case GEOM_CYLINDER: {
      const T t = t_sqrt(l[0] * l[0] + l[1] * l[1]);
      if (t > Eps<T>::minval()) { r[0] = l[0] / t * g.size[0]; r[1] = l[1] / t * g.size[0]; }
      r[2] = t_sgn(l[2]) * g.size[1];
      break;
    }
    case GEOM_BOX: for (int k = 0; k < 3; k++) r[k] = t_sgn(l[k]) * g.size[k]; break;
    case GEOM_MESH: {
      T best = T(-1e30);
      int ib = 0;
      for (int i = 0; i < g.nvert; i++) {
        const T v = (T)g.vert[3 * i] * l[0] + (T)g.vert[3 * i + 1] * l[1] + (T)g.vert[3 * i + 2] * l[2];
        if (v > best) { best = v; ib = i; }
      }
      if (g.nvert) { r[0] = (T)g.vert[3 * ib]; r[1] = (T)g.vert[3 * ib + 1]; r[2] = (T)g.vert[3 * ib + 2]; }
      break;
    }
    default: break;
  }
  for (int k = 0; k < 3; k++)
    res[k] = g.mat[3 * k] * r[0] + g.mat[3 * k + 1] * r[1] + g.mat[3 * k + 2] * r[2] + g.pos[k] + T(0.5) * margin * dir[k];
}

template <typename T> struct MprV { T v[3], v1[3], v2[3]; };

template <typename T, typename G>
__device__ __noinline__ void d_mpr_support(MprV<T>& s, const G& a, const G& b, const T* dir, T margin) {
  const T nd[3] = {-dir[0], -dir[1], -dir[2]};
  d_support_geom(s.v1, a, dir, margin);
  d_support_geom(s.v2, b, nd, margin);
  for (int k = 0; k < 3; k++) s.v[k] = s.v1[k] - s.v2[k];
}

template <typename T> __device__ __forceinline__ void d_sub3(T* r, const T* a, const T* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }

template <typename T>
__device__ __forceinline__ void d_portal_dir(const MprV<T>* P, T* dir) {
  T e1[3], e2[3];
  d_sub3(e1, P[2].v, P[1].v);
  d_sub3(e2, P[3].v, P[1].v);
  cross3(dir, e1, e2);
  ccd_normalize(dir);
}

template <typename T>
__device__ __forceinline__ bool d_portal_reach_tol(const MprV<T>* P, const MprV<T>& v4, const T* dir) {
  const T dv4 = dot3(v4.v, dir);
  T d = dv4 - dot3(P[1].v, dir);
  d = t_min(d, dv4 - dot3(P[2].v, dir));
  d = t_min(d, dv4 - dot3(P[3].v, dir));
  return ccd_eq(d, T(1e-6)) || d < T(1e-6);
}

template <typename T>
__device__ __forceinline__ void d_expand_portal(MprV<T>* P, const MprV<T>& v4) {
  T v4v0[3];
  cross3(v4v0, v4.v, P[0].v);
  int slot;
  if (dot3(P[1].v, v4v0) > 0) slot = dot3(P[2].v, v4v0) > 0 ? 1 : 3;
  else slot = dot3(P[3].v, v4v0) > 0 ? 2 : 1;
  P[slot] = v4;
}

template <typename T>
__device__ T d_point_seg_dist2(const T* x0, const T* b, T* wit) {
  T d[3];
  d_sub3(d, b, x0);
  const T t = -dot3(x0, d) / dot3(d, d);
  if (t < 0 || ccd_zero(t)) { wit[0] = x0[0]; wit[1] = x0[1]; wit[2] = x0[2]; }
  else if (t > 1 || ccd_eq(t, T(1))) { wit[0] = b[0]; wit[1] = b[1]; wit[2] = b[2]; }
  else { for (int k = 0; k < 3; k++) wit[k] = x0[k] + t * d[k]; }
  return dot3(wit, wit);
}

template <typename T>
__device__ T d_point_tri_dist2(const T* x0, const T* B, const T* Cc, T* wit) {
  T d1[3], d2[3];
  d_sub3(d1, B, x0);
  d_sub3(d2, Cc, x0);
  const T v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(x0, d1), q = dot3(x0, d2), r = dot3(d1, d2);
  const T div = w * v - r * r;
  T s = -1, t = -1;
  if (!ccd_zero(div)) { s = (q * r - w * p) / div; t = (-s * r - q) / w; }
  if ((ccd_zero(s) || s > 0) && (ccd_eq(s, T(1)) || s < 1) && (ccd_zero(t) || t > 0) && (ccd_eq(t, T(1)) || t < 1) &&
      (ccd_eq(t + s, T(1)) || t + s < 1)) {
    for (int k = 0; k < 3; k++) wit[k] = x0[k] + s * d1[k] + t * d2[k];
    return dot3(wit, wit);
  }
  T w2[3];
  T dist = d_point_seg_dist2(x0, B, wit);
  T d2s = d_point_seg_dist2(x0, Cc, w2);
  if (d2s < dist) { dist = d2s; wit[0] = w2[0]; wit[1] = w2[1]; wit[2] = w2[2]; }
  d2s = d_point_seg_dist2(B, Cc, w2);
  if (d2s < dist) { dist = d2s; wit[0] = w2[0]; wit[1] = w2[1]; wit[2] = w2[2]; }
  return dist;
}

template <typename T>
__device__ void d_mpr_find_pos(const MprV<T>* P, T* pos) {
  T dir[3], vec[3], b[4];
  d_portal_dir(P, dir);
  cross3(vec, P[1].v, P[2].v); b[0] = dot3(vec, P[3].v);
  cross3(vec, P[3].v, P[2].v); b[1] = dot3(vec, P[0].v);
  cross3(vec, P[0].v, P[1].v); b[2] = dot3(vec, P[3].v);
  cross3(vec, P[2].v, P[1].v); b[3] = dot3(vec, P[0].v);
  T sum = b[0] + b[1] + b[2] + b[3];
  if (ccd_zero(sum) || sum < 0) {
    b[0] = 0;
    cross3(vec, P[2].v, P[3].v); b[1] = dot3(vec, dir);
    cross3(vec, P[3].v, P[1].v); b[2] = dot3(vec, dir);
    cross3(vec, P[1].v, P[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  const T inv = T(1) / sum;
  for (int k = 0; k < 3; k++) {
    T p1 = 0, p2 = 0;
    for (int i = 0; i < 4; i++) { p1 += b[i] * P[i].v1[k]; p2 += b[i] * P[i].v2[k]; }
    pos[k] = T(0.5) * (p1 * inv + p2 * inv);
  }
}

template <typename T, typename G>
__device__ __noinline__ int d_convex_convex(RawCon<T>* out, const G& a, const G& b, T margin) {
  MprV<T> P[4], v4;
  T dir[3], va[3], vb[3], pos[3], nrm[3], depth;
  // portal discovery
  for (int k = 0; k < 3; k++) { P[0].v1[k] = a.pos[k]; P[0].v2[k] = b.pos[k]; P[0].v[k] = a.pos[k] - b.pos[k]; }
  if (ccd_zero(P[0].v[0]) && ccd_zero(P[0].v[1]) && ccd_zero(P[0].v[2])) P[0].v[0] += 10 * CcdEps<T>::v();
  for (int k = 0; k < 3; k++) dir[k] = -P[0].v[k];
  ccd_normalize(dir);
  d_mpr_support(P[1], a, b, dir, margin);
  T dt = dot3(P[1].v, dir);
  if (ccd_zero(dt) || dt < 0) return 0;
  cross3(dir, P[0].v, P[1].v);
  if (ccd_zero(dot3(dir, dir))) {
    if (ccd_zero(P[1].v[0]) && ccd_zero(P[1].v[1]) && ccd_zero(P[1].v[2])) return 0;  // touching: undefined normal
    for (int k = 0; k < 3; k++) { pos[k] = T(0.5) * (P[1].v1[k] + P[1].v2[k]); nrm[k] = P[1].v[k]; }
    depth = norm3(nrm);
    ccd_normalize(nrm);
    set_raw(out[0], margin - depth, pos, nrm);
    return 1;
  }
  ccd_normalize(dir);
  d_mpr_support(P[2], a, b, dir, margin);
  dt = dot3(P[2].v, dir);
  if (ccd_zero(dt) || dt < 0) return 0;
  d_sub3(va, P[1].v, P[0].v);
  d_sub3(vb, P[2].v, P[0].v);
  cross3(dir, va, vb);
  ccd_normalize(dir);
  if (dot3(dir, P[0].v) > 0) {
    const MprV<T> t = P[1]; P[1] = P[2]; P[2] = t;
    for (int k = 0; k < 3; k++) dir[k] = -dir[k];
  }
  for (int guard = 0;; guard++) {
    if (guard > 1000) return 0;
    d_mpr_support(P[3], a, b, dir, margin);
    dt = dot3(P[3].v, dir);
    if (ccd_zero(dt) || dt < 0) return 0;
    bool cont = false;
    cross3(va, P[1].v, P[3].v);
    dt = dot3(va, P[0].v);
    if (dt < 0 && !ccd_zero(dt)) { P[2] = P[3]; cont = true; }
    if (!cont) {
      cross3(va, P[3].v, P[2].v);
      dt = dot3(va, P[0].v);
      if (dt < 0 && !ccd_zero(dt)) { P[1] = P[3]; cont = true; }
    }
    if (!cont) break;
    d_sub3(va, P[1].v, P[0].v);
    d_sub3(vb, P[2].v, P[0].v);
    cross3(dir, va, vb);
    ccd_normalize(dir);
  }
  // refinement
  for (int guard = 0;; guard++) {
    if (guard > 1000) return 0;
    d_portal_dir(P, dir);
    dt = dot3(dir, P[1].v);
    if (ccd_zero(dt) || dt > 0) break;
    d_mpr_support(v4, a, b, dir, margin);
    dt = dot3(v4.v, dir);
    if (!(ccd_zero(dt) || dt > 0) || d_portal_reach_tol(P, v4, dir)) return 0;
    d_expand_portal(P, v4);
  }
  // penetration
  for (int it = 0;; it++) {
    d_portal_dir(P, dir);
    d_mpr_support(v4, a, b, dir, margin);
    if (d_portal_reach_tol(P, v4, dir) || it > 50) {
      T wit[3];
      depth = t_sqrt(d_point_tri_dist2(P[1].v, P[2].v, P[3].v, wit));
      if (ccd_zero(wit[0]) && ccd_zero(wit[1]) && ccd_zero(wit[2])) { nrm[0] = dir[0]; nrm[1] = dir[1]; nrm[2] = dir[2]; }
      else { nrm[0] = wit[0]; nrm[1] = wit[1]; nrm[2] = wit[2]; ccd_normalize(nrm); }
      d_mpr_find_pos(P, pos);
      set_raw(out[0], margin - depth, pos, nrm);
      return 1;
    }
    d_expand_portal(P, v4);
  }
}

// plane against ellipsoid / mesh: support point opposite to the normal (+ up to three more mesh vertices within the margin)
template <typename T>
__device__ __noinline__ int d_plane_convex(RawCon<T>* out, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  T n[3], nd[3], s[3], p[3], dif[3];
  mcol(n, a.mat, 2);
  for (int k = 0; k < 3; k++) nd[k] = -n[k];
  d_support_geom(s, b, nd, T(0));
  d_sub3(dif, s, a.pos);
  const T dist = dot3(dif, n);
  if (dist > margin) return 0;
  for (int k = 0; k < 3; k++) p[k] = s[k] - T(0.5) * dist * n[k];
  set_raw(out[0], dist, p, n);
  int cnt = 1;
  if (b.type == GEOM_MESH) {
    int used[4] = {-1, -1, -1, -1};
    {  // the support vertex itself (same argmax as d_support_geom)
      T l[3];
      for (int k = 0; k < 3; k++) l[k] = b.mat[k] * nd[0] + b.mat[3 + k] * nd[1] + b.mat[6 + k] * nd[2];
      T best = T(-1e30);
      for (int i = 0; i < b.nvert; i++) {
        const T v = b.vert[3 * i] * l[0] + b.vert[3 * i + 1] * l[1] + b.vert[3 * i + 2] * l[2];
        if (v > best) { best = v; used[0] = i; }
      }
    }
    while (cnt < 4) {
      int ib = -1;
      T best = margin, wb[3] = {0, 0, 0};
      for (int i = 0; i < b.nvert; i++) {
        if (i == used[0] || i == used[1] || i == used[2] || i == used[3]) continue;
        T w[3];
        for (int k = 0; k < 3; k++) w[k] = b.mat[3 * k] * b.vert[3 * i] + b.mat[3 * k + 1] * b.vert[3 * i + 1] + b.mat[3 * k + 2] * b.vert[3 * i + 2] + b.pos[k];
        d_sub3(dif, w, a.pos);
        const T di = dot3(dif, n);
        if (di < best || (di == best && ib < 0)) { best = di; ib = i; wb[0] = w[0]; wb[1] = w[1]; wb[2] = w[2]; }
      }
      if (ib < 0) break;
      used[cnt] = ib;
      for (int k = 0; k < 3; k++) p[k] = wb[k] - T(0.5) * best * n[k];
      set_raw(out[cnt], best, p, n);
      cnt++;
    }
  }
  return cnt;
}

template <typename T>
__device__ int narrow_phase(RawCon<T>* out, int t1, int t2, const GeomW<T>& a, const GeomW<T>& b, T margin) {
  T n[3];
  switch (t1 * 8 + t2) {
    case GEOM_PLANE * 8 + GEOM_SPHERE: mcol(n, a.mat, 2); return d_plane_sphere(out, a.pos, n, b.pos, b.size[0], margin);
    case GEOM_PLANE * 8 + GEOM_CAPSULE: return d_plane_capsule(out, a, b, margin);
    case GEOM_PLANE * 8 + GEOM_CYLINDER: return d_plane_cylinder(out, a, b, margin);
    case GEOM_PLANE * 8 + GEOM_BOX: return d_plane_box(out, a, b, margin);
    case GEOM_SPHERE * 8 + GEOM_SPHERE: return d_sphere_sphere(out, a.pos, a.size[0], b.pos, b.size[0], margin);
    case GEOM_SPHERE * 8 + GEOM_CAPSULE: {
      T ax[3], q[3];
      mcol(ax, b.mat, 2);
      const T dif[3] = {a.pos[0] - b.pos[0], a.pos[1] - b.pos[1], a.pos[2] - b.pos[2]};
      const T t = t_min(b.size[1], t_max(-b.size[1], dot3(dif, ax)));
      q[0] = b.pos[0] + t * ax[0]; q[1] = b.pos[1] + t * ax[1]; q[2] = b.pos[2] + t * ax[2];
      return d_sphere_sphere(out, a.pos, a.size[0], q, b.size[0], margin);
    }
    case GEOM_SPHERE * 8 + GEOM_CYLINDER: return d_sphere_cylinder(out, a, b, margin);
    case GEOM_SPHERE * 8 + GEOM_BOX: return d_sphere_box(out, a.pos, a.size[0], b, margin);
    case GEOM_CAPSULE * 8 + GEOM_CAPSULE: return d_capsule_capsule(out, a, b, margin);
    case GEOM_CAPSULE * 8 + GEOM_BOX: return d_capsule_box(out, a, b, margin);
    case GEOM_BOX * 8 + GEOM_BOX: return d_box_box(out, a, b, margin);
    case GEOM_PLANE * 8 + GEOM_ELLIPSOID: case GEOM_PLANE * 8 + GEOM_MESH: return d_plane_convex(out, a, b, margin);
    default:
      if (t1 >= GEOM_SPHERE && t2 <= GEOM_MESH) {
        if constexpr (sizeof(T) == 4) {
          // The general convex routine always runs in fp64.  MPR stops on a comparison against mpr_tolerance = 1e-6, which
          // fp32 rounding (1e-7 on these sizes) moves by whole iterations: at shallow depths the fp32 normal was off by up
          // to 0.2 (12 degrees) against the fp64 oracle on the same pose, while fp64 arithmetic on fp32-rounded poses stays
          // within 7e-3.  Convex pairs are a small share of k_collide and B200 runs fp64 at half the fp32 rate.
          GeomC<double, T> A, B;
          for (int k = 0; k < 3; k++) { A.pos[k] = a.pos[k]; B.pos[k] = b.pos[k]; A.size[k] = a.size[k]; B.size[k] = b.size[k]; }
          for (int k = 0; k < 9; k++) { A.mat[k] = a.mat[k]; B.mat[k] = b.mat[k]; }
          A.type = a.type; A.vert = a.vert; A.nvert = a.nvert; B.type = b.type; B.vert = b.vert; B.nvert = b.nvert;
          RawCon<double> o;
          const int n = d_convex_convex<double>(&o, A, B, (double)margin);
          if (n) {
            out[0].dist = (T)o.dist;
            for (int k = 0; k < 3; k++) { out[0].pos[k] = (T)o.pos[k]; out[0].n[k] = (T)o.n[k]; out[0].tan[k] = 0; }
          }
          return n;
        } else {
          return d_convex_convex<T>(out, a, b, margin);
        }
      }
      return 0;
  }
}

__host__ __device__ inline bool pair_supported(int t1, int t2) {
  switch (t1 * 8 + t2) {
    case GEOM_PLANE * 8 + GEOM_SPHERE: case GEOM_PLANE * 8 + GEOM_CAPSULE: case GEOM_PLANE * 8 + GEOM_CYLINDER:
    case GEOM_PLANE * 8 + GEOM_BOX: case GEOM_SPHERE * 8 + GEOM_SPHERE: case GEOM_SPHERE * 8 + GEOM_CAPSULE:
    case GEOM_SPHERE * 8 + GEOM_CYLINDER: case GEOM_SPHERE * 8 + GEOM_BOX: case GEOM_CAPSULE * 8 + GEOM_CAPSULE:
    case GEOM_CAPSULE * 8 + GEOM_BOX: case GEOM_BOX * 8 + GEOM_BOX:
    case GEOM_PLANE * 8 + GEOM_ELLIPSOID: case GEOM_PLANE * 8 + GEOM_MESH:
      return true;
    default: return t1 >= GEOM_SPHERE && t2 <= GEOM_MESH;  // general convex path (MPR)
  }
}

template <typename T>
__device__ void load_geom(GeomW<T>& g, const MV<T>& m, const KArgs<T>& a, int gi, int env) {
  const DModel& h = *m.h;
  const long long S = a.nenvp;
  for (int k = 0; k < 3; k++) g.pos[k] = a.geom_xpos[(3 * gi + k) * S + env];
  for (int k = 0; k < 9; k++) g.mat[k] = a.geom_xmat[(9 * gi + k) * S + env];
  for (int k = 0; k < 3; k++) g.size[k] = m.f(h.o_geom_size, 3 * gi + k);
  g.type = m.i(h.o_geom_type, gi);
  g.nvert = m.i(h.o_geom_vertnum, gi);
  g.vert = reinterpret_cast<const T*>(a.model + h.o_mesh_vert) + 3 * m.i(h.o_geom_vertadr, gi);
}

// complete contact c of the environment from a raw narrow-phase result: frame, margins, parameter mixing, ids
template <typename T>
__device__ void emit_contact(const MV<T>& m, const KArgs<T>& a, int env, int c, const RawCon<T>& raw, int g1, int g2, int p, T margin) {
  const DModel& h = *m.h;
  const long long S = a.nenvp;
  auto F = [&](int f) -> T& { return a.con[((long long)f * h.nconmax + c) * S + env]; };
  auto I = [&](int f) -> int& { return a.coni[((long long)f * h.nconmax + c) * S + env]; };
  F(CF_DIST) = raw.dist;
  for (int k = 0; k < 3; k++) F(CF_POS + k) = raw.pos[k];
  // complete the frame: normal, tangent hint (Gram-Schmidt) or a fixed pick, then their cross product
  T x[3] = {raw.n[0], raw.n[1], raw.n[2]}, y[3] = {raw.tan[0], raw.tan[1], raw.tan[2]}, z[3];
  normalize3(x);
  if (norm3(y) < T(0.5)) {
    y[0] = 0; y[1] = 0; y[2] = 0;
    if (x[1] < T(0.5) && x[1] > T(-0.5)) y[1] = 1; else y[2] = 1;
  }
  const T dd = dot3(x, y);
  y[0] -= dd * x[0]; y[1] -= dd * x[1]; y[2] -= dd * x[2];
  normalize3(y);
  cross3(z, x, y);
  for (int k = 0; k < 3; k++) { F(CF_FRAME + k) = x[k]; F(CF_FRAME + 3 + k) = y[k]; F(CF_FRAME + 6 + k) = z[k]; }
  const T gap = t_max(m.f(h.o_geom_gap, g1), m.f(h.o_geom_gap, g2));
  F(CF_INCLUDEMARGIN) = margin - gap;
  // parameter mixing
  const int pr1 = m.i(h.o_geom_priority, g1), pr2 = m.i(h.o_geom_priority, g2);
  T fr[3];
  int dim;
  if (pr1 != pr2) {
    const int g = pr1 > pr2 ? g1 : g2;
    dim = m.i(h.o_geom_condim, g);
    for (int k = 0; k < 3; k++) fr[k] = m.f(h.o_geom_friction, 3 * g + k);
    for (int k = 0; k < 2; k++) F(CF_SOLREF + k) = m.f(h.o_geom_solref, 2 * g + k);
    for (int k = 0; k < 5; k++) F(CF_SOLIMP + k) = m.f(h.o_geom_solimp, 5 * g + k);
  } else {
    dim = max(m.i(h.o_geom_condim, g1), m.i(h.o_geom_condim, g2));
    for (int k = 0; k < 3; k++) fr[k] = t_max(m.f(h.o_geom_friction, 3 * g1 + k), m.f(h.o_geom_friction, 3 * g2 + k));
    const T s1 = m.f(h.o_geom_solmix, g1), s2 = m.f(h.o_geom_solmix, g2);
    T mix;
    if (s1 >= Eps<T>::minval() && s2 >= Eps<T>::minval()) mix = s1 / (s1 + s2);
    else if (s1 < Eps<T>::minval() && s2 < Eps<T>::minval()) mix = T(0.5);
    else mix = s1 < Eps<T>::minval() ? T(0) : T(1);
    const T r10 = m.f(h.o_geom_solref, 2 * g1), r20 = m.f(h.o_geom_solref, 2 * g2);
    for (int k = 0; k < 2; k++) {
      const T r1 = m.f(h.o_geom_solref, 2 * g1 + k), r2 = m.f(h.o_geom_solref, 2 * g2 + k);
      F(CF_SOLREF + k) = (r10 > 0 && r20 > 0) ? mix * r1 + (1 - mix) * r2 : t_min(r1, r2);
    }
    for (int k = 0; k < 5; k++) F(CF_SOLIMP + k) = mix * m.f(h.o_geom_solimp, 5 * g1 + k) + (1 - mix) * m.f(h.o_geom_solimp, 5 * g2 + k);
  }
  const T minmu = T(1e-5);
  F(CF_FRICTION) = F(CF_FRICTION + 1) = t_max(minmu, fr[0]);
  F(CF_FRICTION + 2) = t_max(minmu, fr[1]);
  F(CF_FRICTION + 3) = F(CF_FRICTION + 4) = t_max(minmu, fr[2]);
  I(CI_GEOM1) = g1; I(CI_GEOM2) = g2; I(CI_DIM) = dim; I(CI_PAIR) = p; I(CI_EFC) = -1;
}

// K2 + K3: a team of L lanes per environment.  Phase 1: the lanes cull the static pair list side by side (bounding
// spheres; planes by signed centre distance) and compact the survivors, in pair order, into a shared-memory candidate
// list (ballot + popcount).  Phase 2: L candidates at a time go through the narrow phase, one per lane; an exclusive
// scan of the contact counts over the team gives every lane its output slots, so the contact list comes out in
// MuJoCo's pair order whatever the team width (geom ids and pair indices are bit-exact against the sequential oracle).
template <typename T, int BLOCK, int L>
// (measured: a 72-register build, seven CTAs per SM = one wave for 16384 environments: C3 0.141 -> 0.137 ms, PR2 0.87 -> 0.77,
//  but the 20-slot world 0.18 -> 0.33 ms; the 128-register build stays)
#ifdef B2_COLLIDE_MINB
__global__ void __launch_bounds__(BLOCK, B2_COLLIDE_MINB) k_collide(const KArgs<T> a) {
#else
__global__ void __launch_bounds__(BLOCK) k_collide(const KArgs<T> a) {
#endif
  if ((a.flags & B2F_FUSABLE) && a.pending[0] == 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* blob = reinterpret_cast<uint32_t*>(smem_raw + 16);
  const int nwords = reinterpret_cast<const DModel*>(a.model)->nwords;
  stage_model(blob, a.model, nwords, bar);
  MV<T> m{reinterpret_cast<const DModel*>(blob), blob};
  const DModel& h = *m.h;
  const long long S = a.nenvp;
  constexpr int EPB = BLOCK / L;
  const int ntiles = a.ncount / EPB;
  const bool off = h.disableflags & (DSBL_CONSTRAINT | DSBL_CONTACT);
  const int team = threadIdx.x / L, l = threadIdx.x % L;
  const int tshift = (threadIdx.x & 31) & ~(L - 1);
  const unsigned tmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << tshift;
  const int npp = (h.npair + 1) & ~1;
  uint16_t* cand = reinterpret_cast<uint16_t*>(smem_raw + 16 + (((size_t)nwords * 4 + 15) & ~(size_t)15)) + (size_t)team * npp;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * EPB + team;
    int ncand = 0;
    __syncwarp(tmask);
    for (int base = 0; base < h.npair && !off; base += L) {
      const int p = base + l;
      bool pass = false;
      if (p < h.npair) {
        const int g1 = m.i(h.o_pair_geom1, p), g2 = m.i(h.o_pair_geom2, p);
        const int t1 = m.i(h.o_geom_type, g1);
        const T margin = t_max(m.f(h.o_geom_margin, g1), m.f(h.o_geom_margin, g2));
        T x1[3], x2[3];
        for (int k = 0; k < 3; k++) { x1[k] = a.geom_xpos[(3 * g1 + k) * S + env]; x2[k] = a.geom_xpos[(3 * g2 + k) * S + env]; }
        const T dif[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
        if (t1 == GEOM_PLANE) {
          const T n[3] = {a.geom_xmat[(9 * g1 + 2) * S + env], a.geom_xmat[(9 * g1 + 5) * S + env], a.geom_xmat[(9 * g1 + 8) * S + env]};
          pass = !(dot3(dif, n) > m.f(h.o_geom_rbound, g2) + margin);
        } else {
          const T bound = m.f(h.o_geom_rbound, g1) + m.f(h.o_geom_rbound, g2) + margin;
          pass = !(dot3(dif, dif) > bound * bound);
        }
      }
      const unsigned bal = (__ballot_sync(tmask, pass) >> tshift) & (L == 32 ? 0xffffffffu : ((1u << L) - 1u));
      if (pass) cand[ncand + __popc(bal & ((1u << l) - 1u))] = (uint16_t)p;
      ncand += __popc(bal);
    }
    __syncwarp(tmask);
    int ncon = 0;
    for (int base = 0; base < ncand; base += L) {
      const int i = base + l;
      int n = 0, g1 = 0, g2 = 0, p = 0;
      T margin = 0;
      RawCon<T> raw[B2_MAXCONPAIR];
      if (i < ncand) {
        p = cand[i];
        g1 = m.i(h.o_pair_geom1, p); g2 = m.i(h.o_pair_geom2, p);
        margin = t_max(m.f(h.o_geom_margin, g1), m.f(h.o_geom_margin, g2));
        GeomW<T> ga, gb;
        load_geom(ga, m, a, g1, env);
        load_geom(gb, m, a, g2, env);
        n = narrow_phase(raw, ga.type, gb.type, ga, gb, margin);
      }
      int incl = n;
#pragma unroll
      for (int o = 1; o < L; o <<= 1) { const int v = __shfl_up_sync(tmask, incl, o, L); if (l >= o) incl += v; }
      const int total = __shfl_sync(tmask, incl, L - 1, L);
      const int first = ncon + incl - n;
      for (int j = 0; j < n; j++) {
        if (first + j >= h.nconmax) { a.status[env] |= 1; break; }
        emit_contact(m, a, env, first + j, raw[j], g1, g2, p, margin);
      }
      ncon = min(ncon + total, h.nconmax);
    }
    if (l == 0) a.ncon[env] = ncon;
  }
}

}  // namespace b2
