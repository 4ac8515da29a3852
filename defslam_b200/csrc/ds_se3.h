/*
 * ds_se3.h -- SE(3) arithmetic of the camera vertex, fp64.
 *
 * Semantics follow the reference (paths under the DefSLAM tree):
 *   SE3Quat::exp / ctor / operator* / normalizeRotation
 *       Thirdparty/g2o/g2o/types/se3quat.h:57-59,103-109,223-257,280-285
 *   VertexSE3Expmap::oplusImpl        Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:73-76
 *   Converter::toSE3Quat / toCvMat    Thirdparty/ORBSLAM_2/src/Converter.cc:31-66
 * Quaternions are (x, y, z, w).
 */
#ifndef DS_SE3_H_
#define DS_SE3_H_
#include "ds_common.h"

namespace ds {

struct Pose {
  double q[4];
  double t[3];
};

/* Eigen::Quaterniond(Matrix3d); written without run-time array indexing so that
 * everything stays in registers */
DS_FN void quat_from_R(const double R[9], double q[4]) {
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > (i == 0 ? R[0] : R[4])) i = 2;
    if (i == 0) {        /* j = 1, k = 2 */
      t = sqrt(R[0] - R[4] - R[8] + 1.0);
      q[0] = 0.5 * t; t = 0.5 / t;
      q[3] = (R[7] - R[5]) * t;
      q[1] = (R[3] + R[1]) * t;
      q[2] = (R[6] + R[2]) * t;
    } else if (i == 1) { /* j = 2, k = 0 */
      t = sqrt(R[4] - R[8] - R[0] + 1.0);
      q[1] = 0.5 * t; t = 0.5 / t;
      q[3] = (R[2] - R[6]) * t;
      q[2] = (R[7] + R[5]) * t;
      q[0] = (R[1] + R[3]) * t;
    } else {             /* j = 0, k = 1 */
      t = sqrt(R[8] - R[0] - R[4] + 1.0);
      q[2] = 0.5 * t; t = 0.5 / t;
      q[3] = (R[3] - R[1]) * t;
      q[0] = (R[2] + R[6]) * t;
      q[1] = (R[5] + R[7]) * t;
    }
  }
}

DS_FN void quat_normalize(double q[4]) {
  if (q[3] < 0) {
    q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3];
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

DS_FN void quat_to_R(const double q[4], double R[9]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

DS_FN void quat_mul(const double a[4], const double b[4], double o[4]) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

/* q * v as Eigen evaluates it: v + w*(2 q x v) + q x (2 q x v) */
DS_FN void quat_rot(const double q[4], const double v[3], double o[3]) {
  const double ux = 2 * (q[1] * v[2] - q[2] * v[1]);
  const double uy = 2 * (q[2] * v[0] - q[0] * v[2]);
  const double uz = 2 * (q[0] * v[1] - q[1] * v[0]);
  o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}

DS_FN void pose_map(const Pose &P, const double v[3], double o[3]) {
  quat_rot(P.q, v, o);
  o[0] += P.t[0]; o[1] += P.t[1]; o[2] += P.t[2];
}

DS_FN void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

/* update = (omega, upsilon) */
DS_FN void se3_exp(const double u[6], double q[4], double t[3]) {
  const double w0 = u[0], w1 = u[1], w2 = u[2];
  const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
  const double Om[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
  double Om2[9], R[9], V[9];
  mat3_mul(Om, Om, Om2);
  if (theta < 0.00001) {
#pragma unroll
    for (int i = 0; i < 9; i++) {
      R[i] = ((i % 4) == 0 ? 1.0 : 0.0) + Om[i] + Om2[i];
      V[i] = R[i];
    }
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta),
                 c = (theta - sin(theta)) / pow(theta, 3);
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4) == 0 ? 1.0 : 0.0;
      R[i] = I + a * Om[i] + b * Om2[i];
      V[i] = I + b * Om[i] + c * Om2[i];
    }
  }
  quat_from_R(R, q);
#pragma unroll
  for (int i = 0; i < 3; i++) t[i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
  quat_normalize(q);
}

/* estimate <- exp(update) * estimate */
DS_FN void pose_oplus(Pose &P, const double u[6]) {
  double dq[4], dt[3], rt[3], nq[4];
  se3_exp(u, dq, dt);
  quat_rot(dq, P.t, rt);
  for (int i = 0; i < 3; i++) P.t[i] = dt[i] + rt[i];
  quat_mul(dq, P.q, nq);
  P.q[0] = nq[0]; P.q[1] = nq[1]; P.q[2] = nq[2]; P.q[3] = nq[3];
  quat_normalize(P.q);
}

DS_FN void pose_from_Tcw(const float T[16], Pose &P) {
  const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  quat_from_R(R, P.q);
  quat_normalize(P.q);
  P.t[0] = T[3]; P.t[1] = T[7]; P.t[2] = T[11];
}

DS_FN void pose_to_Tcw(const Pose &P, float T[16]) {
  double R[9];
  quat_to_R(P.q, R);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = (float)R[i * 3 + j];
    T[i * 4 + 3] = (float)P.t[i];
  }
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
}

}  // namespace ds
#endif
