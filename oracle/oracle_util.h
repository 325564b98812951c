// oracle_util.h — fp64 math helpers for the CPU oracle (test infrastructure only, see oracle.h).
// Spatial-vector conventions as in MuJoCo's docs: 6D motion = [angular; linear], 6D force = [torque; force];
// 10-number inertia = (Ixx,Iyy,Izz,Ixy,Ixz,Iyz, m*cx,m*cy,m*cz, m) about the reference point.
#pragma once
#include <cmath>
#include <cstring>

#include "mujoco/mujoco.h"

namespace omath {

inline void zero(mjtNum* r, int n) { for (int i = 0; i < n; i++) r[i] = 0; }
inline void copy(mjtNum* r, const mjtNum* a, int n) { for (int i = 0; i < n; i++) r[i] = a[i]; }
inline mjtNum dot(const mjtNum* a, const mjtNum* b, int n) { mjtNum s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }
inline mjtNum dot3(const mjtNum* a, const mjtNum* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline mjtNum norm3(const mjtNum* a) { return std::sqrt(dot3(a, a)); }
inline void cross(mjtNum* r, const mjtNum* a, const mjtNum* b) {
  mjtNum x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
inline mjtNum normalize3(mjtNum* a) {
  mjtNum n = norm3(a);
  if (n < mjMINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
inline mjtNum normalize4(mjtNum* q) {
  mjtNum n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < mjMINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (std::fabs(n - 1) > mjMINVAL) { q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n; }
  return n;
}
inline void mulQuat(mjtNum* r, const mjtNum* a, const mjtNum* b) {
  mjtNum w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  mjtNum x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  mjtNum y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  mjtNum z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
inline void quat2mat(mjtNum* m, const mjtNum* q) {
  mjtNum q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  mjtNum q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  mjtNum q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = 2 * (q12 - q03);       m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03);       m[4] = q00 - q11 + q22 - q33; m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02);       m[7] = 2 * (q23 + q01);       m[8] = q00 - q11 - q22 + q33;
}
inline void mulMatVec3(mjtNum* r, const mjtNum* m, const mjtNum* v) {
  mjtNum x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  mjtNum y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  mjtNum z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
inline void mulMatTVec3(mjtNum* r, const mjtNum* m, const mjtNum* v) {
  mjtNum x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  mjtNum y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  mjtNum z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
inline void rotVecQuat(mjtNum* r, const mjtNum* v, const mjtNum* q) {
  mjtNum m[9];
  quat2mat(m, q);
  mulMatVec3(r, m, v);
}
inline void axisAngle2Quat(mjtNum* q, const mjtNum* axis, mjtNum angle) {
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  mjtNum s = std::sin(angle * 0.5);
  q[0] = std::cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
// 3D rotation vector taking quaternion qb to qa, expressed in the frame of qb:  qa = qb * quat(res)
inline void subQuat(mjtNum* res, const mjtNum* qa, const mjtNum* qb) {
  mjtNum qneg[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, qdif[4];
  mulQuat(qdif, qneg, qa);
  mjtNum axis[3] = {qdif[1], qdif[2], qdif[3]};
  mjtNum sin_a_2 = normalize3(axis);
  mjtNum speed = 2 * std::atan2(sin_a_2, qdif[0]);
  if (speed > mjPI) speed -= 2 * mjPI;
  if (sin_a_2 < mjMINVAL) { res[0] = res[1] = res[2] = 0; return; }
  res[0] = axis[0] * speed; res[1] = axis[1] * speed; res[2] = axis[2] * speed;
}
// q <- normalize(q) * quat(vel*scale); vel is expressed in the local frame
inline void quatIntegrate(mjtNum* quat, const mjtNum* vel, mjtNum scale) {
  mjtNum tmp[3] = {vel[0], vel[1], vel[2]}, qrot[4], res[4];
  mjtNum speed = norm3(tmp);
  if (speed < mjMINVAL) { normalize4(quat); return; }
  tmp[0] /= speed; tmp[1] /= speed; tmp[2] /= speed;
  axisAngle2Quat(qrot, tmp, scale * speed);
  normalize4(quat);
  mulQuat(res, quat, qrot);
  copy(quat, res, 4);
}
// inertia about an offset point: res = (R diag(inert) R^T shifted by `dif`, mass*dif, mass)
inline void inertCom(mjtNum* res, const mjtNum* inert, const mjtNum* mat, const mjtNum* dif, mjtNum mass) {
  mjtNum tmp[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      tmp[3 * i + j] = mat[3 * i] * inert[0] * mat[3 * j] + mat[3 * i + 1] * inert[1] * mat[3 * j + 1] + mat[3 * i + 2] * inert[2] * mat[3 * j + 2];
  res[0] = tmp[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] = tmp[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] = tmp[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] = tmp[1] - mass * dif[0] * dif[1];
  res[4] = tmp[2] - mass * dif[0] * dif[2];
  res[5] = tmp[5] - mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2];
  res[9] = mass;
}
// spatial momentum = inertia * motion vector
inline void mulInertVec(mjtNum* res, const mjtNum* i, const mjtNum* v) {
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
// motion axis of a dof expressed about the CoM reference point
inline void dofCom(mjtNum* res, const mjtNum* axis, const mjtNum* offset) {
  if (offset) {
    res[0] = axis[0]; res[1] = axis[1]; res[2] = axis[2];
    cross(res + 3, axis, offset);
  } else {
    res[0] = res[1] = res[2] = 0;
    res[3] = axis[0]; res[4] = axis[1]; res[5] = axis[2];
  }
}
// spatial cross products: motion x motion, motion x* force
inline void crossMotion(mjtNum* res, const mjtNum* vel, const mjtNum* v) {
  mjtNum a[3], b[3], c[3];
  cross(a, vel, v);
  cross(b, vel, v + 3);
  cross(c, vel + 3, v);
  res[0] = a[0]; res[1] = a[1]; res[2] = a[2];
  res[3] = b[0] + c[0]; res[4] = b[1] + c[1]; res[5] = b[2] + c[2];
}
inline void crossForce(mjtNum* res, const mjtNum* vel, const mjtNum* f) {
  mjtNum a[3], b[3], c[3];
  cross(a, vel, f);
  cross(b, vel + 3, f + 3);
  cross(c, vel, f + 3);
  res[0] = a[0] + b[0]; res[1] = a[1] + b[1]; res[2] = a[2] + b[2];
  res[3] = c[0]; res[4] = c[1]; res[5] = c[2];
}

}  // namespace omath
