// oracle_collision.cpp — fp64 CPU restatement of broad- and narrow-phase collision: primitive geom pairs and the
// general convex path (MPR) for every other pair of convex geoms.
// TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED: MuJoCo's collision functions are not in
// /root/reference; conventions follow MuJoCo's docs (mjContact: dist < 0 = penetration, pos = midpoint,
// frame[0:3] = normal pointing from geom1 to geom2; geom1 has the lower geom TYPE).  Contact order: the static
// candidate-pair list (mjModel.pair_geom1/2, compiled in MuJoCo's pair-generation order: body pair, then geom pair),
// then emission order inside the pair function.  The reference
// reaches this only through mj_step1 / mj_forward / mj_inverse (src/mj_main.cpp:83, mj_ros.cpp:608,
// mj_hw_interface.cpp:61).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "oracle.h"
#include "oracle_util.h"

using namespace omath;

namespace {

struct G {  // one geom in the world frame
  const mjtNum* pos;
  const mjtNum* mat;  // row-major; column k is local axis k
  const mjtNum* size;
  int type = -1;               // mjtGeom (used by the general convex path only)
  const mjtNum* vert = nullptr;  // mesh vertices in the geom frame, nvert x 3
  int nvert = 0;
};
struct C {  // raw contact before parameter mixing
  mjtNum dist, pos[3], normal[3], tangent[3];
};

inline void col(mjtNum* r, const mjtNum* mat, int k) { r[0] = mat[k]; r[1] = mat[3 + k]; r[2] = mat[6 + k]; }
inline void set_c(C& c, mjtNum dist, const mjtNum* pos, const mjtNum* n) {
  c.dist = dist;
  copy(c.pos, pos, 3);
  copy(c.normal, n, 3);
  zero(c.tangent, 3);
}
inline void axpy3(mjtNum* r, const mjtNum* a, mjtNum s, const mjtNum* b) { for (int k = 0; k < 3; k++) r[k] = a[k] + s * b[k]; }

// sphere (centre c1, radius r1) against sphere; normal from 1 to 2
int sphere_sphere_raw(C* out, const mjtNum* c1, mjtNum r1, const mjtNum* c2, mjtNum r2, mjtNum margin) {
  mjtNum dif[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
  const mjtNum cd = norm3(dif);
  const mjtNum dist = cd - r1 - r2;
  if (dist > margin) return 0;
  mjtNum n[3] = {1, 0, 0};
  if (cd >= mjMINVAL) for (int k = 0; k < 3; k++) n[k] = dif[k] / cd;
  mjtNum p[3];
  axpy3(p, c1, r1 + 0.5 * dist, n);
  set_c(out[0], dist, p, n);
  return 1;
}

int plane_sphere_raw(C* out, const mjtNum* ppos, const mjtNum* n, const mjtNum* c, mjtNum r, mjtNum margin) {
  const mjtNum dif[3] = {c[0] - ppos[0], c[1] - ppos[1], c[2] - ppos[2]};
  const mjtNum dist = dot3(dif, n) - r;
  if (dist > margin) return 0;
  mjtNum p[3];
  axpy3(p, c, -(r + 0.5 * dist), n);
  set_c(out[0], dist, p, n);
  return 1;
}

int plane_sphere(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum n[3];
  col(n, a.mat, 2);
  return plane_sphere_raw(out, a.pos, n, b.pos, b.size[0], margin);
}

int plane_capsule(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum n[3], ax[3], e[3];
  col(n, a.mat, 2);
  col(ax, b.mat, 2);
  int cnt = 0;
  for (int s = 1; s >= -1; s -= 2) {
    axpy3(e, b.pos, s * b.size[1], ax);
    cnt += plane_sphere_raw(out + cnt, a.pos, n, e, b.size[0], margin);
  }
  for (int i = 0; i < cnt; i++) copy(out[i].tangent, ax, 3);  // first tangent along the capsule axis
  return cnt;
}

int plane_cylinder(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum n[3], ax[3];
  col(n, a.mat, 2);
  col(ax, b.mat, 2);
  mjtNum prjaxis = dot3(n, ax);
  if (prjaxis > 0) { for (mjtNum& x : ax) x = -x; prjaxis = -prjaxis; }  // axis points toward the plane
  const mjtNum dif[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
  const mjtNum dist0 = dot3(dif, n);
  // direction in the cap disk that goes deepest into the plane
  mjtNum vec[3];
  for (int k = 0; k < 3; k++) vec[k] = ax[k] * prjaxis - n[k];
  const mjtNum len = norm3(vec);
  if (len >= 1e-12) for (mjtNum& x : vec) x *= b.size[0] / len;
  else { col(vec, b.mat, 0); for (mjtNum& x : vec) x *= b.size[0]; }  // disk parallel to the plane
  const mjtNum prjvec = dot3(vec, n);
  mjtNum axs[3] = {ax[0] * b.size[1], ax[1] * b.size[1], ax[2] * b.size[1]};
  prjaxis *= b.size[1];
  int cnt = 0;
  mjtNum p[3];
  // deepest rim point of the near cap
  mjtNum d1 = dist0 + prjaxis + prjvec;
  if (d1 > margin) return 0;
  for (int k = 0; k < 3; k++) p[k] = b.pos[k] + vec[k] + axs[k] - n[k] * d1 * 0.5;
  set_c(out[cnt++], d1, p, n);
  // same side, far cap
  mjtNum d2 = dist0 - prjaxis + prjvec;
  if (d2 <= margin) {
    for (int k = 0; k < 3; k++) p[k] = b.pos[k] + vec[k] - axs[k] - n[k] * d2 * 0.5;
    set_c(out[cnt++], d2, p, n);
  }
  // two more rim points of the near cap, 120 degrees either side
  const mjtNum prjvec1 = -0.5 * prjvec;
  mjtNum d3 = dist0 + prjaxis + prjvec1;
  if (d3 <= margin) {
    mjtNum side[3];
    cross(side, vec, ax);
    normalize3(side);
    for (mjtNum& x : side) x *= b.size[0] * std::sqrt(3.0) * 0.5;
    for (int s = 1; s >= -1; s -= 2) {
      for (int k = 0; k < 3; k++) p[k] = b.pos[k] + s * side[k] + axs[k] - 0.5 * vec[k] - n[k] * d3 * 0.5;
      set_c(out[cnt++], d3, p, n);
    }
  }
  return cnt;
}

int plane_box(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum n[3];
  col(n, a.mat, 2);
  const mjtNum dif[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
  const mjtNum dist = dot3(dif, n);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    const mjtNum v[3] = {(i & 1 ? 1 : -1) * b.size[0], (i & 2 ? 1 : -1) * b.size[1], (i & 4 ? 1 : -1) * b.size[2]};
    mjtNum corner[3];
    mulMatVec3(corner, b.mat, v);
    const mjtNum ld = dot3(n, corner);
    if (dist + ld > margin || ld > 0) continue;  // too far, or on the upper half of the box
    const mjtNum cd = dist + ld;
    mjtNum p[3];
    for (int k = 0; k < 3; k++) p[k] = b.pos[k] + corner[k] - n[k] * cd * 0.5;
    set_c(out[cnt], cd, p, n);
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

int sphere_sphere(C* out, const G& a, const G& b, mjtNum margin) {
  return sphere_sphere_raw(out, a.pos, a.size[0], b.pos, b.size[0], margin);
}

// nearest point of the segment (centre, axis, half-length) to point p
void nearest_on_segment(mjtNum* r, const mjtNum* centre, const mjtNum* axis, mjtNum half, const mjtNum* p) {
  const mjtNum dif[3] = {p[0] - centre[0], p[1] - centre[1], p[2] - centre[2]};
  const mjtNum t = std::min(half, std::max(-half, dot3(dif, axis)));
  axpy3(r, centre, t, axis);
}

int sphere_capsule(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum ax[3], q[3];
  col(ax, b.mat, 2);
  nearest_on_segment(q, b.pos, ax, b.size[1], a.pos);
  return sphere_sphere_raw(out, a.pos, a.size[0], q, b.size[0], margin);
}

// sphere against a convex solid given the nearest-point result in the solid's local frame.
// c = sphere centre (local), p = nearest point of the solid to c, (inside: c == p) then `nin`/`depth_in` give the
// outward normal and depth of the shallowest exit.
int sphere_solid(C* out, const G& sph, const G& solid, const mjtNum* c, const mjtNum* p, const mjtNum* nin, mjtNum depth_in,
                 mjtNum margin) {
  mjtNum dl[3] = {c[0] - p[0], c[1] - p[1], c[2] - p[2]};
  const mjtNum dn = norm3(dl);
  mjtNum nloc[3], dist;
  const mjtNum r = sph.size[0];
  if (dn > 1e-12) {  // centre outside the solid
    dist = dn - r;
    for (int k = 0; k < 3; k++) nloc[k] = dl[k] / dn;
  } else {
    dist = -depth_in - r;
    copy(nloc, nin, 3);
  }
  if (dist > margin) return 0;
  mjtNum nw[3], pw[3];
  mulMatVec3(nw, solid.mat, nloc);
  for (mjtNum& x : nw) x = -x;  // from the sphere (geom1) toward the solid (geom2)
  axpy3(pw, sph.pos, r + 0.5 * dist, nw);
  set_c(out[0], dist, pw, nw);
  return 1;
}

int sphere_box(C* out, const G& a, const G& b, mjtNum margin) {
  const mjtNum dif[3] = {a.pos[0] - b.pos[0], a.pos[1] - b.pos[1], a.pos[2] - b.pos[2]};
  mjtNum c[3], p[3], nin[3] = {0, 0, 0};
  mulMatTVec3(c, b.mat, dif);
  int kmin = 0;
  mjtNum dmin = 1e300;
  for (int k = 0; k < 3; k++) {
    p[k] = std::min(b.size[k], std::max(-b.size[k], c[k]));
    const mjtNum ex = b.size[k] - std::fabs(c[k]);
    if (ex < dmin) { dmin = ex; kmin = k; }
  }
  nin[kmin] = c[kmin] >= 0 ? 1 : -1;
  return sphere_solid(out, a, b, c, p, nin, dmin, margin);
}

int sphere_cylinder(C* out, const G& a, const G& b, mjtNum margin) {
  const mjtNum dif[3] = {a.pos[0] - b.pos[0], a.pos[1] - b.pos[1], a.pos[2] - b.pos[2]};
  mjtNum c[3], p[3], nin[3] = {0, 0, 0};
  mulMatTVec3(c, b.mat, dif);
  const mjtNum R = b.size[0], H = b.size[1];
  const mjtNum rho = std::sqrt(c[0] * c[0] + c[1] * c[1]);
  const mjtNum s = rho > R ? R / rho : 1;
  p[0] = c[0] * s; p[1] = c[1] * s;
  p[2] = std::min(H, std::max(-H, c[2]));
  const mjtNum ex_r = R - rho, ex_z = H - std::fabs(c[2]);
  mjtNum depth;
  if (ex_r < ex_z) {
    depth = ex_r;
    if (rho > 1e-12) { nin[0] = c[0] / rho; nin[1] = c[1] / rho; } else nin[0] = 1;
  } else {
    depth = ex_z;
    nin[2] = c[2] >= 0 ? 1 : -1;
  }
  return sphere_solid(out, a, b, c, p, nin, depth, margin);
}

// closest points between two segments (centre, unit axis, half-length): returns parameters t1, t2
void segment_segment(const mjtNum* c1, const mjtNum* a1, mjtNum h1, const mjtNum* c2, const mjtNum* a2, mjtNum h2,
                     mjtNum& t1, mjtNum& t2, bool& parallel) {
  const mjtNum dif[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
  const mjtNum b = dot3(a1, a2), u = -dot3(a1, dif), v = dot3(a2, dif);
  const mjtNum det = 1 - b * b;
  parallel = det < 1e-10;
  if (parallel) {
    t1 = 0;
  } else {
    t1 = (u + b * v) / det;
  }
  t1 = std::min(h1, std::max(-h1, t1));
  t2 = std::min(h2, std::max(-h2, v + b * t1));
  t1 = std::min(h1, std::max(-h1, u + b * t2));
}

int capsule_capsule(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum a1[3], a2[3];
  col(a1, a.mat, 2);
  col(a2, b.mat, 2);
  mjtNum t1, t2;
  bool par;
  segment_segment(a.pos, a1, a.size[1], b.pos, a2, b.size[1], t1, t2, par);
  if (!par) {
    mjtNum p1[3], p2[3];
    axpy3(p1, a.pos, t1, a1);
    axpy3(p2, b.pos, t2, a2);
    return sphere_sphere_raw(out, p1, a.size[0], p2, b.size[0], margin);
  }
  // parallel axes: two contacts at the ends of the overlap interval (projected on axis 1)
  const mjtNum dif[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
  const mjtNum s = dot3(a1, a2) >= 0 ? 1 : -1;
  const mjtNum mid = dot3(dif, a1);
  const mjtNum lo = std::max(-a.size[1], mid - b.size[1]), hi = std::min(a.size[1], mid + b.size[1]);
  int cnt = 0;
  if (lo >= hi) {  // no overlap along the axis: nearest end points
    mjtNum p1[3], p2[3];
    axpy3(p1, a.pos, t1, a1);
    axpy3(p2, b.pos, t2, a2);
    return sphere_sphere_raw(out, p1, a.size[0], p2, b.size[0], margin);
  }
  for (int e = 0; e < 2; e++) {
    const mjtNum x = e ? hi : lo;
    mjtNum p1[3], p2[3];
    axpy3(p1, a.pos, x, a1);
    axpy3(p2, b.pos, s * (x - mid), a2);
    cnt += sphere_sphere_raw(out + cnt, p1, a.size[0], p2, b.size[0], margin);
  }
  return cnt;
}

// squared distance from local point c to the box [-s, s], with the nearest point
mjtNum box_dist2(const mjtNum* s, const mjtNum* c, mjtNum* p) {
  mjtNum d2 = 0;
  for (int k = 0; k < 3; k++) {
    p[k] = std::min(s[k], std::max(-s[k], c[k]));
    d2 += (c[k] - p[k]) * (c[k] - p[k]);
  }
  return d2;
}

// capsule (geom1) against box (geom2): spheres of the capsule radius at the segment point nearest to the box and at
// both end points. Nearest segment parameter by golden-section search on the (convex) distance function.
int capsule_box(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum axw[3], ax[3], c0[3];
  col(axw, a.mat, 2);
  const mjtNum dif[3] = {a.pos[0] - b.pos[0], a.pos[1] - b.pos[1], a.pos[2] - b.pos[2]};
  mulMatTVec3(c0, b.mat, dif);  // capsule centre and axis in the box frame
  mulMatTVec3(ax, b.mat, axw);
  const mjtNum h = a.size[1];
  auto f = [&](mjtNum t) {
    mjtNum c[3], p[3];
    axpy3(c, c0, t, ax);
    const mjtNum d2 = box_dist2(b.size, c, p);
    if (d2 > 0) return std::sqrt(d2);
    mjtNum ex = 1e300;  // inside: negative depth so that deeper is "closer"
    for (int k = 0; k < 3; k++) ex = std::min(ex, b.size[k] - std::fabs(c[k]));
    return -ex;
  };
  const mjtNum gr = 0.6180339887498949;
  mjtNum lo = -h, hi = h;
  mjtNum x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo), f1 = f(x1), f2 = f(x2);
  for (int it = 0; it < 40; it++) {
    if (f1 <= f2) { hi = x2; x2 = x1; f2 = f1; x1 = hi - gr * (hi - lo); f1 = f(x1); }
    else { lo = x1; x1 = x2; f1 = f2; x2 = lo + gr * (hi - lo); f2 = f(x2); }
  }
  const mjtNum tbest = 0.5 * (lo + hi);
  const mjtNum ts[3] = {tbest, -h, h};
  int cnt = 0;
  for (int i = 0; i < 3; i++) {
    if (i > 0 && std::fabs(ts[i] - tbest) < 0.1 * h + 1e-9) continue;  // end point coincides with the nearest point
    mjtNum cw[3];
    axpy3(cw, a.pos, ts[i], axw);
    G sph{cw, a.mat, a.size};
    cnt += sphere_box(out + cnt, sph, b, margin);
  }
  return cnt;
}

// ---- box-box: separating-axis test, then face clipping or edge-edge closest points ----
int box_box(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum A[3][3], B[3][3];
  for (int k = 0; k < 3; k++) { col(A[k], a.mat, k); col(B[k], b.mat, k); }
  const mjtNum dif[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
  mjtNum Rm[3][3], Rabs[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { Rm[i][j] = dot3(A[i], B[j]); Rabs[i][j] = std::fabs(Rm[i][j]) + 1e-12; }
  mjtNum best = -1e300;
  int best_code = -1;
  mjtNum best_axis[3] = {0, 0, 0};
  auto consider = [&](const mjtNum* L, mjtNum ra, mjtNum rb, int code, mjtNum bias) {
    const mjtNum t = dot3(dif, L);
    const mjtNum sep = std::fabs(t) - ra - rb;
    if (sep * bias > best + (code >= 1 ? 1e-6 : 0)) {  // a later axis must win by a margin: ties keep the earlier one
      best = sep * bias;
      best_code = code;
      const mjtNum s = t >= 0 ? 1 : -1;
      for (int k = 0; k < 3; k++) best_axis[k] = s * L[k];  // points from a toward b
    }
    return sep;
  };
  for (int i = 0; i < 3; i++) {
    const mjtNum rb = b.size[0] * Rabs[i][0] + b.size[1] * Rabs[i][1] + b.size[2] * Rabs[i][2];
    if (consider(A[i], a.size[i], rb, i, 1) > margin) return 0;
  }
  for (int j = 0; j < 3; j++) {
    const mjtNum ra = a.size[0] * Rabs[0][j] + a.size[1] * Rabs[1][j] + a.size[2] * Rabs[2][j];
    if (consider(B[j], ra, b.size[j], 3 + j, 1) > margin) return 0;
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      mjtNum L[3];
      cross(L, A[i], B[j]);
      const mjtNum ln = norm3(L);
      if (ln < 1e-6) continue;  // parallel edges: covered by the face axes
      for (mjtNum& x : L) x /= ln;
      mjtNum ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) {
        ra += a.size[k] * std::fabs(dot3(A[k], L));
        rb += b.size[k] * std::fabs(dot3(B[k], L));
      }
      if (consider(L, ra, rb, 6 + 3 * i + j, 1) > margin) return 0;
    }
  if (best_code < 0) return 0;

  if (best_code >= 6) {
    // edge-edge: supporting edges of a (direction A[i]) and b (direction B[j]) along best_axis
    const int i = (best_code - 6) / 3, j = (best_code - 6) % 3;
    mjtNum pa[3] = {a.pos[0], a.pos[1], a.pos[2]}, pb[3] = {b.pos[0], b.pos[1], b.pos[2]};
    for (int k = 0; k < 3; k++) {
      if (k != i) { const mjtNum s = dot3(A[k], best_axis) >= 0 ? 1 : -1; for (int c = 0; c < 3; c++) pa[c] += s * a.size[k] * A[k][c]; }
      if (k != j) { const mjtNum s = dot3(B[k], best_axis) >= 0 ? -1 : 1; for (int c = 0; c < 3; c++) pb[c] += s * b.size[k] * B[k][c]; }
    }
    mjtNum t1, t2;
    bool par;
    segment_segment(pa, A[i], a.size[i], pb, B[j], b.size[j], t1, t2, par);
    mjtNum qa[3], qb[3], p[3];
    axpy3(qa, pa, t1, A[i]);
    axpy3(qb, pb, t2, B[j]);
    for (int k = 0; k < 3; k++) p[k] = 0.5 * (qa[k] + qb[k]);
    set_c(out[0], best, p, best_axis);
    return 1;
  }

  // face contact: reference box owns the axis; clip the incident face of the other box against its side planes
  const bool ref_is_a = best_code < 3;
  const G& rf = ref_is_a ? a : b;
  const G& in = ref_is_a ? b : a;
  mjtNum (*RA)[3] = ref_is_a ? A : B;
  mjtNum (*IA)[3] = ref_is_a ? B : A;
  const int ri = ref_is_a ? best_code : best_code - 3;
  mjtNum nref[3];  // outward normal of the reference face, pointing at the incident box
  for (int k = 0; k < 3; k++) nref[k] = ref_is_a ? best_axis[k] : -best_axis[k];
  // incident face: most anti-parallel to nref
  int ii = 0;
  mjtNum mind = 1e300, isgn = 1;
  for (int k = 0; k < 3; k++) {
    const mjtNum dd = dot3(IA[k], nref);
    if (-std::fabs(dd) < mind) { mind = -std::fabs(dd); ii = k; isgn = dd > 0 ? -1 : 1; }
  }
  const int iu = (ii + 1) % 3, iv = (ii + 2) % 3;
  mjtNum poly[16][3], tmp[16][3];
  int np = 4;
  for (int c = 0; c < 4; c++) {
    const mjtNum su = (c == 0 || c == 3) ? -1 : 1, sv = c < 2 ? -1 : 1;
    for (int k = 0; k < 3; k++)
      poly[c][k] = in.pos[k] + isgn * in.size[ii] * IA[ii][k] + su * in.size[iu] * IA[iu][k] + sv * in.size[iv] * IA[iv][k];
  }
  const int ru = (ri + 1) % 3, rv = (ri + 2) % 3;
  const int side_ax[4] = {ru, ru, rv, rv};
  const mjtNum side_sg[4] = {1, -1, 1, -1};
  for (int s = 0; s < 4 && np > 0; s++) {
    // keep the half-space  side_sg * dot(x - rf.pos, RA[ax]) <= size[ax]
    const mjtNum* ax = RA[side_ax[s]];
    const mjtNum lim = rf.size[side_ax[s]];
    int nn = 0;
    for (int c = 0; c < np; c++) {
      const mjtNum* p0 = poly[c];
      const mjtNum* p1 = poly[(c + 1) % np];
      const mjtNum d0v[3] = {p0[0] - rf.pos[0], p0[1] - rf.pos[1], p0[2] - rf.pos[2]};
      const mjtNum d1v[3] = {p1[0] - rf.pos[0], p1[1] - rf.pos[1], p1[2] - rf.pos[2]};
      const mjtNum e0 = side_sg[s] * dot3(d0v, ax) - lim, e1 = side_sg[s] * dot3(d1v, ax) - lim;
      if (e0 <= 0) { copy(tmp[nn], p0, 3); nn++; }
      if ((e0 < 0 && e1 > 0) || (e0 > 0 && e1 < 0)) {
        const mjtNum t = e0 / (e0 - e1);
        for (int k = 0; k < 3; k++) tmp[nn][k] = p0[k] + t * (p1[k] - p0[k]);
        nn++;
      }
    }
    np = nn;
    for (int c = 0; c < np; c++) copy(poly[c], tmp[c], 3);
  }
  int cnt = 0;
  mjtNum nrm[3];  // contact normal from geom1 (a) to geom2 (b)
  copy(nrm, best_axis, 3);
  for (int c = 0; c < np && cnt < mjMAXCONPAIR; c++) {
    const mjtNum dv[3] = {poly[c][0] - rf.pos[0], poly[c][1] - rf.pos[1], poly[c][2] - rf.pos[2]};
    const mjtNum depth = dot3(dv, nref) - rf.size[ri];  // signed distance of the incident vertex to the reference face
    if (depth > margin) continue;
    mjtNum p[3];
    axpy3(p, poly[c], -0.5 * depth, nref);
    set_c(out[cnt++], depth, p, nrm);
  }
  return cnt;
}


// ---------------------------------------------------------------------------------------------------------------
// General convex pairs (s5: "mjc_Convex").  MuJoCo hands every pair without a primitive function to libccd's
// Minkowski Portal Refinement (ccdMPRPenetration; G. Snethen, "XenoCollide", Game Programming Gems 7) with its own
// support functions, mpr_tolerance = 1e-6 and mpr_iterations = 50, and emits ONE contact: dist = margin - depth,
// normal = MPR direction (from geom1 to geom2), pos = midpoint of the two witness points.  libccd is a third-party
// dependency of the closed MuJoCo binary and absent here: the published algorithm is restated (portal discovery,
// refinement until the origin ray crosses the portal, expansion until the support plane is within tolerance of the
// portal, depth = distance from the origin to the portal triangle, position from the barycentric coordinates of the
// origin-ray hit).  Parity unpinned like the rest of this file.
const mjtNum kMprTol = 1e-6;
const int kMprIter = 50;
const mjtNum kCcdEps = 2.220446049250313e-16;

inline bool ccd_zero(mjtNum x) { return std::fabs(x) < kCcdEps; }
inline bool ccd_eq(mjtNum a, mjtNum b) {
  const mjtNum ab = std::fabs(a - b);
  if (ab < kCcdEps) return true;
  const mjtNum fa = std::fabs(a), fb = std::fabs(b);
  return ab < kCcdEps * (fb > fa ? fb : fa);
}
inline mjtNum sgn(mjtNum x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

// support point of one geom (world frame) in world direction `dir` (unit), inflated by margin / 2
void support_geom(mjtNum* res, const G& g, const mjtNum* dir, mjtNum margin) {
  mjtNum l[3], r[3] = {0, 0, 0};
  mulMatTVec3(l, g.mat, dir);
  switch (g.type) {
    case mjGEOM_SPHERE: for (int k = 0; k < 3; k++) r[k] = l[k] * g.size[0]; break;
    case mjGEOM_CAPSULE:
      for (int k = 0; k < 3; k++) r[k] = l[k] * g.size[0];
      r[2] += sgn(l[2]) * g.size[1];
      break;
    case mjGEOM_ELLIPSOID: {
      const mjtNum t[3] = {l[0] * g.size[0], l[1] * g.size[1], l[2] * g.size[2]};
      const mjtNum n = norm3(t);
      if (n >= mjMINVAL) for (int k = 0; k < 3; k++) r[k] = g.size[k] * t[k] / n;
      break;
    }
    case mjGEOM_CYLINDER: {
      const mjtNum t = std::sqrt(l[0] * l[0] + l[1] * l[1]);
      if (t > mjMINVAL) { r[0] = l[0] / t * g.size[0]; r[1] = l[1] / t * g.size[0]; }
      r[2] = sgn(l[2]) * g.size[1];
      break;
    }
    case mjGEOM_BOX: for (int k = 0; k < 3; k++) r[k] = sgn(l[k]) * g.size[k]; break;
    case mjGEOM_MESH: {
      mjtNum best = -1e300;
      int ib = 0;
      for (int i = 0; i < g.nvert; i++) {
        const mjtNum v = dot3(g.vert + 3 * i, l);
        if (v > best) { best = v; ib = i; }
      }
      if (g.nvert) copy(r, g.vert + 3 * ib, 3);
      break;
    }
    default: break;
  }
  mulMatVec3(res, g.mat, r);
  for (int k = 0; k < 3; k++) res[k] += g.pos[k] + 0.5 * margin * dir[k];
}

struct SV { mjtNum v[3], v1[3], v2[3]; };  // point of the Minkowski difference and its two witnesses

void mpr_support(SV& s, const G& a, const G& b, const mjtNum* dir, mjtNum margin) {
  const mjtNum nd[3] = {-dir[0], -dir[1], -dir[2]};
  support_geom(s.v1, a, dir, margin);
  support_geom(s.v2, b, nd, margin);
  for (int k = 0; k < 3; k++) s.v[k] = s.v1[k] - s.v2[k];
}

inline void ccd_normalize(mjtNum* v) { const mjtNum n = norm3(v); for (int k = 0; k < 3; k++) v[k] /= n; }
inline void sub3(mjtNum* r, const mjtNum* a, const mjtNum* b) { for (int k = 0; k < 3; k++) r[k] = a[k] - b[k]; }

void portal_dir(const SV* P, mjtNum* dir) {
  mjtNum e1[3], e2[3];
  sub3(e1, P[2].v, P[1].v);
  sub3(e2, P[3].v, P[1].v);
  cross(dir, e1, e2);
  ccd_normalize(dir);
}

bool portal_reach_tolerance(const SV* P, const SV& v4, const mjtNum* dir) {
  const mjtNum dv4 = dot3(v4.v, dir);
  mjtNum d = dv4 - dot3(P[1].v, dir);
  d = std::min(d, dv4 - dot3(P[2].v, dir));
  d = std::min(d, dv4 - dot3(P[3].v, dir));
  return ccd_eq(d, kMprTol) || d < kMprTol;
}

void expand_portal(SV* P, const SV& v4) {
  mjtNum v4v0[3];
  cross(v4v0, v4.v, P[0].v);
  if (dot3(P[1].v, v4v0) > 0) {
    if (dot3(P[2].v, v4v0) > 0) P[1] = v4; else P[3] = v4;
  } else {
    if (dot3(P[3].v, v4v0) > 0) P[2] = v4; else P[1] = v4;
  }
}

mjtNum point_seg_dist2(const mjtNum* x0, const mjtNum* b, mjtNum* wit) {  // from the origin
  mjtNum d[3];
  sub3(d, b, x0);
  mjtNum t = -dot3(x0, d) / dot3(d, d);
  if (t < 0 || ccd_zero(t)) { copy(wit, x0, 3); }
  else if (t > 1 || ccd_eq(t, 1)) { copy(wit, b, 3); }
  else { for (int k = 0; k < 3; k++) wit[k] = x0[k] + t * d[k]; }
  return dot3(wit, wit);
}

// squared distance from the origin to triangle (x0, B, C) and the nearest point
mjtNum point_tri_dist2(const mjtNum* x0, const mjtNum* B, const mjtNum* Cc, mjtNum* wit) {
  mjtNum d1[3], d2[3];
  sub3(d1, B, x0);
  sub3(d2, Cc, x0);
  const mjtNum v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(x0, d1), q = dot3(x0, d2), r = dot3(d1, d2);
  const mjtNum div = w * v - r * r;
  mjtNum s = -1, t = -1;
  if (!ccd_zero(div)) { s = (q * r - w * p) / div; t = (-s * r - q) / w; }
  if ((ccd_zero(s) || s > 0) && (ccd_eq(s, 1) || s < 1) && (ccd_zero(t) || t > 0) && (ccd_eq(t, 1) || t < 1) &&
      (ccd_eq(t + s, 1) || t + s < 1)) {
    for (int k = 0; k < 3; k++) wit[k] = x0[k] + s * d1[k] + t * d2[k];
    return dot3(wit, wit);
  }
  mjtNum w2[3];
  mjtNum dist = point_seg_dist2(x0, B, wit);
  mjtNum d2s = point_seg_dist2(x0, Cc, w2);
  if (d2s < dist) { dist = d2s; copy(wit, w2, 3); }
  d2s = point_seg_dist2(B, Cc, w2);
  if (d2s < dist) { dist = d2s; copy(wit, w2, 3); }
  return dist;
}

void mpr_find_pos(const SV* P, mjtNum* pos) {
  mjtNum dir[3], vec[3], b[4];
  portal_dir(P, dir);
  cross(vec, P[1].v, P[2].v); b[0] = dot3(vec, P[3].v);
  cross(vec, P[3].v, P[2].v); b[1] = dot3(vec, P[0].v);
  cross(vec, P[0].v, P[1].v); b[2] = dot3(vec, P[3].v);
  cross(vec, P[2].v, P[1].v); b[3] = dot3(vec, P[0].v);
  mjtNum sum = b[0] + b[1] + b[2] + b[3];
  if (ccd_zero(sum) || sum < 0) {
    b[0] = 0;
    cross(vec, P[2].v, P[3].v); b[1] = dot3(vec, dir);
    cross(vec, P[3].v, P[1].v); b[2] = dot3(vec, dir);
    cross(vec, P[1].v, P[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  const mjtNum inv = 1 / sum;
  for (int k = 0; k < 3; k++) {
    mjtNum p1 = 0, p2 = 0;
    for (int i = 0; i < 4; i++) { p1 += b[i] * P[i].v1[k]; p2 += b[i] * P[i].v2[k]; }
    pos[k] = 0.5 * (p1 * inv + p2 * inv);
  }
}

// returns 1 and (depth, dir, pos) when the inflated geoms intersect
int mpr_penetration(const G& a, const G& b, mjtNum margin, mjtNum& depth, mjtNum* dir_out, mjtNum* pos) {
  SV P[4], v4;
  mjtNum dir[3], va[3], vb[3];
  // --- portal discovery ---
  copy(P[0].v1, a.pos, 3);
  copy(P[0].v2, b.pos, 3);
  sub3(P[0].v, a.pos, b.pos);
  if (ccd_zero(P[0].v[0]) && ccd_zero(P[0].v[1]) && ccd_zero(P[0].v[2])) P[0].v[0] += 10 * kCcdEps;
  for (int k = 0; k < 3; k++) dir[k] = -P[0].v[k];
  ccd_normalize(dir);
  mpr_support(P[1], a, b, dir, margin);
  mjtNum dt = dot3(P[1].v, dir);
  if (ccd_zero(dt) || dt < 0) return 0;
  cross(dir, P[0].v, P[1].v);
  if (ccd_zero(dot3(dir, dir))) {
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (P[1].v1[k] + P[1].v2[k]);
    if (ccd_zero(P[1].v[0]) && ccd_zero(P[1].v[1]) && ccd_zero(P[1].v[2])) {  // touching: the normal is undefined, no contact
      return 0;
    }
    depth = norm3(P[1].v);                                                    // origin on the segment v0-v1
    copy(dir_out, P[1].v, 3);
    ccd_normalize(dir_out);
    return 1;
  }
  ccd_normalize(dir);
  mpr_support(P[2], a, b, dir, margin);
  dt = dot3(P[2].v, dir);
  if (ccd_zero(dt) || dt < 0) return 0;
  sub3(va, P[1].v, P[0].v);
  sub3(vb, P[2].v, P[0].v);
  cross(dir, va, vb);
  ccd_normalize(dir);
  if (dot3(dir, P[0].v) > 0) { std::swap(P[1], P[2]); for (int k = 0; k < 3; k++) dir[k] = -dir[k]; }
  for (int guard = 0;; guard++) {
    if (guard > 1000) return 0;
    mpr_support(P[3], a, b, dir, margin);
    dt = dot3(P[3].v, dir);
    if (ccd_zero(dt) || dt < 0) return 0;
    bool cont = false;
    cross(va, P[1].v, P[3].v);
    dt = dot3(va, P[0].v);
    if (dt < 0 && !ccd_zero(dt)) { P[2] = P[3]; cont = true; }
    if (!cont) {
      cross(va, P[3].v, P[2].v);
      dt = dot3(va, P[0].v);
      if (dt < 0 && !ccd_zero(dt)) { P[1] = P[3]; cont = true; }
    }
    if (!cont) break;
    sub3(va, P[1].v, P[0].v);
    sub3(vb, P[2].v, P[0].v);
    cross(dir, va, vb);
    ccd_normalize(dir);
  }
  // --- refinement: until the portal faces the origin from outside ---
  for (int guard = 0;; guard++) {
    if (guard > 1000) return 0;
    portal_dir(P, dir);
    dt = dot3(dir, P[1].v);
    if (ccd_zero(dt) || dt > 0) break;
    mpr_support(v4, a, b, dir, margin);
    dt = dot3(v4.v, dir);
    if (!(ccd_zero(dt) || dt > 0) || portal_reach_tolerance(P, v4, dir)) return 0;
    expand_portal(P, v4);
  }
  // --- penetration: expand until the support plane is within tolerance of the portal ---
  for (int it = 0;; it++) {
    portal_dir(P, dir);
    mpr_support(v4, a, b, dir, margin);
    if (portal_reach_tolerance(P, v4, dir) || it > kMprIter) {
      mjtNum wit[3];
      depth = std::sqrt(point_tri_dist2(P[1].v, P[2].v, P[3].v, wit));
      if (ccd_zero(wit[0]) && ccd_zero(wit[1]) && ccd_zero(wit[2])) copy(dir_out, dir, 3);
      else { copy(dir_out, wit, 3); ccd_normalize(dir_out); }
      mpr_find_pos(P, pos);
      return 1;
    }
    expand_portal(P, v4);
  }
}

int convex_convex(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum depth = 0, dir[3], pos[3];
  if (!mpr_penetration(a, b, margin, depth, dir, pos)) return 0;
  set_c(out[0], margin - depth, pos, dir);
  return 1;
}

// plane against a convex geom without a primitive function (ellipsoid, mesh): the support point opposite to the plane
// normal; a mesh adds up to three more vertices that are within the margin (MuJoCo walks the neighbours of the support
// vertex in the qhull graph — the hull graph is not available here, so the candidates are all vertices, lowest first).
int plane_convex(C* out, const G& a, const G& b, mjtNum margin) {
  mjtNum n[3], nd[3], s[3], p[3];
  col(n, a.mat, 2);
  for (int k = 0; k < 3; k++) nd[k] = -n[k];
  support_geom(s, b, nd, 0);
  mjtNum dif[3];
  sub3(dif, s, a.pos);
  mjtNum dist = dot3(dif, n);
  if (dist > margin) return 0;
  axpy3(p, s, -0.5 * dist, n);
  set_c(out[0], dist, p, n);
  int cnt = 1;
  if (b.type == mjGEOM_MESH) {
    int used[4] = {-1, -1, -1, -1};
    {  // index of the support vertex (same argmax as support_geom)
      mjtNum l[3], best = -1e300;
      mulMatTVec3(l, b.mat, nd);
      for (int i = 0; i < b.nvert; i++) {
        const mjtNum v = dot3(b.vert + 3 * i, l);
        if (v > best) { best = v; used[0] = i; }
      }
    }
    while (cnt < 4) {
      int ib = -1;
      mjtNum best = margin;
      mjtNum wb[3] = {0, 0, 0};
      for (int i = 0; i < b.nvert; i++) {
        if (i == used[0] || i == used[1] || i == used[2] || i == used[3]) continue;
        mjtNum w[3];
        mulMatVec3(w, b.mat, b.vert + 3 * i);
        for (int k = 0; k < 3; k++) w[k] += b.pos[k];
        sub3(dif, w, a.pos);
        const mjtNum di = dot3(dif, n);
        if (di < best || (di == best && ib < 0)) { best = di; ib = i; copy(wb, w, 3); }
      }
      if (ib < 0) break;
      used[cnt] = ib;
      axpy3(p, wb, -0.5 * best, n);
      set_c(out[cnt], best, p, n);
      cnt++;
    }
  }
  return cnt;
}

typedef int (*PairFn)(C*, const G&, const G&, mjtNum);
PairFn pair_fn(int t1, int t2) {
  if (t1 == mjGEOM_PLANE) {
    if (t2 == mjGEOM_SPHERE) return plane_sphere;
    if (t2 == mjGEOM_CAPSULE) return plane_capsule;
    if (t2 == mjGEOM_CYLINDER) return plane_cylinder;
    if (t2 == mjGEOM_BOX) return plane_box;
  } else if (t1 == mjGEOM_SPHERE) {
    if (t2 == mjGEOM_SPHERE) return sphere_sphere;
    if (t2 == mjGEOM_CAPSULE) return sphere_capsule;
    if (t2 == mjGEOM_CYLINDER) return sphere_cylinder;
    if (t2 == mjGEOM_BOX) return sphere_box;
  } else if (t1 == mjGEOM_CAPSULE) {
    if (t2 == mjGEOM_CAPSULE) return capsule_capsule;
    if (t2 == mjGEOM_BOX) return capsule_box;
  } else if (t1 == mjGEOM_BOX && t2 == mjGEOM_BOX) {
    return box_box;
  }
  if (t1 == mjGEOM_PLANE) return (t2 == mjGEOM_ELLIPSOID || t2 == mjGEOM_MESH) ? plane_convex : nullptr;
  if (t1 >= mjGEOM_SPHERE && t1 != mjGEOM_HFIELD && t2 >= mjGEOM_SPHERE && t2 <= mjGEOM_MESH) return convex_convex;
  return nullptr;
}

// complete the contact frame from the normal (and optional first-tangent hint)
void make_frame(mjtNum* frame, const mjtNum* normal, const mjtNum* hint) {
  mjtNum x[3], y[3], z[3];
  copy(x, normal, 3);
  normalize3(x);
  copy(y, hint, 3);
  if (norm3(y) < 0.5) {
    zero(y, 3);
    if (x[1] < 0.5 && x[1] > -0.5) y[1] = 1; else y[2] = 1;
  }
  const mjtNum dd = dot3(x, y);
  for (int k = 0; k < 3; k++) y[k] -= dd * x[k];
  normalize3(y);
  cross(z, x, y);
  copy(frame, x, 3);
  copy(frame + 3, y, 3);
  copy(frame + 6, z, 3);
}

}  // namespace

extern "C" int omj_pair_supported(int t1, int t2) { return pair_fn(t1, t2) != nullptr; }

// test hook: run the general convex path on two explicitly placed primitives; out = {dist, pos[3], normal[3]}
extern "C" int omj_convex_pair(int t1, const mjtNum* pos1, const mjtNum* mat1, const mjtNum* size1, int t2, const mjtNum* pos2,
                               const mjtNum* mat2, const mjtNum* size2, mjtNum margin, mjtNum* out) {
  G a{pos1, mat1, size1, t1}, b{pos2, mat2, size2, t2};
  C c[1];
  const int n = convex_convex(c, a, b, margin);
  if (n) { out[0] = c[0].dist; copy(out + 1, c[0].pos, 3); copy(out + 4, c[0].normal, 3); }
  return n;
}

void omj_collision(const mjModel* m, mjData* d) {
  d->ncon = 0;
  if (m->opt.disableflags & (mjDSBL_CONSTRAINT | mjDSBL_CONTACT)) return;
  C raw[mjMAXCONPAIR + 8];
  for (int p = 0; p < m->npair; p++) {
    const int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
    const int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    PairFn fn = pair_fn(t1, t2);
    if (!fn) continue;
    const mjtNum margin = std::max(m->geom_margin[g1], m->geom_margin[g2]);
    const mjtNum gap = std::max(m->geom_gap[g1], m->geom_gap[g2]);
    // broad phase on bounding spheres (planes: centre distance to the plane)
    const mjtNum* x1 = d->geom_xpos + 3 * g1;
    const mjtNum* x2 = d->geom_xpos + 3 * g2;
    if (t1 == mjGEOM_PLANE) {
      mjtNum n[3];
      col(n, d->geom_xmat + 9 * g1, 2);
      const mjtNum dif[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
      if (dot3(dif, n) > m->geom_rbound[g2] + margin) continue;
    } else {
      const mjtNum dif[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
      const mjtNum bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
      if (dot3(dif, dif) > bound * bound) continue;
    }
    G a{x1, d->geom_xmat + 9 * g1, m->geom_size + 3 * g1, t1}, b{x2, d->geom_xmat + 9 * g2, m->geom_size + 3 * g2, t2};
    if (t1 == mjGEOM_MESH) { const int id = m->geom_dataid[g1]; a.vert = m->mesh_vert + 3 * m->mesh_vertadr[id]; a.nvert = m->mesh_vertnum[id]; }
    if (t2 == mjGEOM_MESH) { const int id = m->geom_dataid[g2]; b.vert = m->mesh_vert + 3 * m->mesh_vertadr[id]; b.nvert = m->mesh_vertnum[id]; }
    const int n = fn(raw, a, b, margin);
    for (int i = 0; i < n; i++) {
      if (d->ncon >= m->nconmax) return;  // cap reached: later pairs are dropped (flagged by the caller via ncon == nconmax)
      mjContact* con = d->contact + d->ncon++;
      std::memset(con, 0, sizeof(*con));
      con->dist = raw[i].dist;
      copy(con->pos, raw[i].pos, 3);
      make_frame(con->frame, raw[i].normal, raw[i].tangent);
      con->includemargin = margin - gap;
      con->geom1 = g1; con->geom2 = g2; con->pair = p;
      con->efc_address = -1;
      // parameter mixing (priority, then max / solmix-weighted average)
      const int pr1 = m->geom_priority[g1], pr2 = m->geom_priority[g2];
      const mjtNum *f1 = m->geom_friction + 3 * g1, *f2 = m->geom_friction + 3 * g2;
      mjtNum fr[3];
      if (pr1 != pr2) {
        const int g = pr1 > pr2 ? g1 : g2;
        con->dim = m->geom_condim[g];
        copy(fr, m->geom_friction + 3 * g, 3);
        copy(con->solref, m->geom_solref + 2 * g, 2);
        copy(con->solimp, m->geom_solimp + 5 * g, 5);
      } else {
        con->dim = std::max(m->geom_condim[g1], m->geom_condim[g2]);
        for (int k = 0; k < 3; k++) fr[k] = std::max(f1[k], f2[k]);
        mjtNum mix;
        const mjtNum s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2];
        if (s1 >= mjMINVAL && s2 >= mjMINVAL) mix = s1 / (s1 + s2);
        else if (s1 < mjMINVAL && s2 < mjMINVAL) mix = 0.5;
        else mix = s1 < mjMINVAL ? 0 : 1;
        const mjtNum *r1 = m->geom_solref + 2 * g1, *r2 = m->geom_solref + 2 * g2;
        if (r1[0] > 0 && r2[0] > 0) for (int k = 0; k < 2; k++) con->solref[k] = mix * r1[k] + (1 - mix) * r2[k];
        else for (int k = 0; k < 2; k++) con->solref[k] = std::min(r1[k], r2[k]);
        for (int k = 0; k < 5; k++) con->solimp[k] = mix * m->geom_solimp[5 * g1 + k] + (1 - mix) * m->geom_solimp[5 * g2 + k];
      }
      con->friction[0] = con->friction[1] = std::max(mjMINMU, fr[0]);
      con->friction[2] = std::max(mjMINMU, fr[1]);
      con->friction[3] = con->friction[4] = std::max(mjMINMU, fr[2]);
      con->mu = con->friction[0];
    }
  }
}
