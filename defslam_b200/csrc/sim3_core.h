/*
 * sim3_core.h -- device code of the Sim(3) surface registration and of the min-median scale
 * (last stage of DefLocalMapping::NRSfM; see include/defslam_b200.h for what it replaces).
 *
 * One CTA per keyframe: the points are spread over the threads, the 7x7 normal equations are a
 * fixed-order team reduction of 35 sums, the LM scalar logic runs uniformly on every thread.
 * The Jacobian is the analytic limit of the reference's central differences:
 *   e = p2 - S.map(p1),  y = S.map(p1):   de/d(omega) = [y]x,  de/d(upsilon) = -I,  de/d(sigma) = -y.
 */
#ifndef DS_SIM3_CORE_H_
#define DS_SIM3_CORE_H_
#include "../../include/defslam_b200.h"
#include "ds_common.h"

namespace ds {

struct Sim3 {
  double q[4], t[3], s; /* q = (x,y,z,w) */
};

/* Eigen::Quaterniond(Matrix3d) */
DS_FN void s3_q_from_R(const double R[9], double q[4]) {
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}
DS_FN void s3_q_mul(const double a[4], const double b[4], double o[4]) {
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
DS_FN void s3_q_rot(const double q[4], const double v[3], double o[3]) {
  const double u0 = 2 * (q[1] * v[2] - q[2] * v[1]), u1 = 2 * (q[2] * v[0] - q[0] * v[2]),
               u2 = 2 * (q[0] * v[1] - q[1] * v[0]);
  o[0] = v[0] + q[3] * u0 + (q[1] * u2 - q[2] * u1);
  o[1] = v[1] + q[3] * u1 + (q[2] * u0 - q[0] * u2);
  o[2] = v[2] + q[3] * u2 + (q[0] * u1 - q[1] * u0);
}

/* Sim3(const Vector7d &update)  sim3.h:70-135 */
DS_FN void sim3_exp(const double u[7], Sim3 &S) {
  const double w0 = u[0], w1 = u[1], w2 = u[2], sigma = u[6];
  const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
  const double Om[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Om2[3 * i + j] = Om[3 * i] * Om[j] + Om[3 * i + 1] * Om[3 + j] + Om[3 * i + 2] * Om[6 + j];
  S.s = exp(sigma);
  const double eps = 0.00001;
  double A, B, C, a1 = 1.0, a2 = 1.0;
  if (fabs(sigma) < eps) {
    C = 1;
    if (theta < eps) { A = 1. / 2.; B = 1. / 6.; }
    else {
      const double th2 = theta * theta;
      A = (1 - cos(theta)) / th2;
      B = (theta - sin(theta)) / (th2 * theta);
      a1 = sin(theta) / theta; a2 = (1 - cos(theta)) / (theta * theta);
    }
  } else {
    C = (S.s - 1) / sigma;
    if (theta < eps) {
      const double s2 = sigma * sigma;
      A = ((sigma - 1) * S.s + 1) / s2;
      B = ((0.5 * s2 - sigma + 1) * S.s) / (s2 * sigma);
    } else {
      a1 = sin(theta) / theta; a2 = (1 - cos(theta)) / (theta * theta);
      const double a = S.s * sin(theta), b = S.s * cos(theta), th2 = theta * theta, s2 = sigma * sigma, c = th2 + s2;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / th2;
    }
  }
  double R[9];
  for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a1 * Om[i] + a2 * Om2[i];
  s3_q_from_R(R, S.q);
  for (int i = 0; i < 3; i++) {
    double acc = 0.0;
    for (int j = 0; j < 3; j++) acc += (A * Om[3 * i + j] + B * Om2[3 * i + j] + (i == j ? C : 0.0)) * u[3 + j];
    S.t[i] = acc;
  }
}

DS_FN void sim3_mul(const Sim3 &a, const Sim3 &b, Sim3 &o) {
  Sim3 r;
  s3_q_mul(a.q, b.q, r.q);
  double rt[3];
  s3_q_rot(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.s * rt[i] + a.t[i];
  r.s = a.s * b.s;
  o = r;
}
DS_FN void sim3_map(const Sim3 &S, const double x[3], double o[3]) {
  double r[3];
  s3_q_rot(S.q, x, r);
  for (int i = 0; i < 3; i++) o[i] = S.s * r[i] + S.t[i];
}

struct Sim3Prob {
  int n;
  const float *p1, *p2;
  Sim3 init;
  double chi, huber;
  int max_iterations;
  double *out; /* [16]: q(4) t(3) s chi2 inliers acceptable it0 it1 */
};

/* fixed-order sum of nv (<= 36) values per thread over the team; result in red[0..nv) */
DS_FN void team_sum_vec(const Team &team, const double *v, int nv, double *red) {
#if DS_CUDA
  const int w = team.tid >> 5, nw = (team.nthr + 31) >> 5, lane = team.tid & 31;
  team.sync();
  for (int k = 0; k < nv; k++) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) red[36 + w * 36 + k] = x;
  }
  team.sync();
  if (team.tid < nv) {
    double s = 0.0;
    for (int i = 0; i < nw; i++) s += red[36 + i * 36 + team.tid];
    red[team.tid] = s;
  }
  team.sync();
#else
  (void)team;
  for (int k = 0; k < nv; k++) red[k] = v[k];
#endif
}
/* doubles of shared scratch team_sum_vec needs for nthr threads */
static inline
#if DS_CUDA
__host__ __device__
#endif
int sim3_red_doubles(int nthr) { return 36 + 36 * ((nthr + 31) / 32) + 40; }

DS_FN void edge_error(const Sim3Prob &P, const Sim3 &S, int i, double e[3], double y[3]) {
  const double x1[3] = {(double)P.p1[3 * i], (double)P.p1[3 * i + 1], (double)P.p1[3 * i + 2]};
  sim3_map(S, x1, y);
  for (int c = 0; c < 3; c++) e[c] = (double)P.p2[3 * i + c] - y[c];
}

DS_FN double sim3_robust_chi2(const Team &team, const Sim3Prob &P, const Sim3 &S, double delta, double dsqr, double *red) {
  double chi = 0.0;
  DS_FOR(i, P.n) {
    double e[3], y[3];
    edge_error(P, S, i, e, y);
    const double c2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    chi += c2 <= dsqr ? c2 : 2 * sqrt(c2) * delta - dsqr;
  }
  return team_sum(team, chi, red + 36 + 36 * ((team.nthr + 31) / 32));
}

DS_FN bool ldlt7(const double *H, double lambda, const double *b, double *x) {
  double L[49], v[7];
  for (int j = 0; j < 7; j++) {
    double dj = H[7 * j + j] + lambda;
    for (int k = 0; k < j; k++) { v[k] = L[7 * j + k] * L[7 * k + k]; dj -= L[7 * j + k] * v[k]; }
    L[7 * j + j] = dj;
    if (!(dj > 0.0)) return false;
    for (int i = j + 1; i < 7; i++) {
      double s = H[7 * i + j];
      for (int k = 0; k < j; k++) s -= L[7 * i + k] * v[k];
      L[7 * i + j] = s / dj;
    }
  }
  for (int i = 0; i < 7; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[7 * i + k] * x[k]; x[i] = s; }
  for (int i = 0; i < 7; i++) x[i] /= L[7 * i + i];
  for (int i = 6; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < 7; k++) s -= L[7 * k + i] * x[k]; x[i] = s; }
  return true;
}

/* one optimizer.optimize(max_it) (g2o LM with the ORB-SLAM stop rule); est is updated in place,
 * last = the state of the last evaluation (the errors g2o leaves on the edges) */
DS_FN int sim3_run_lm(const Team &team, const Sim3Prob &P, Sim3 &est, Sim3 &last, double delta, double dsqr, double *red) {
  double lambda = -1., ni = 2.;
  int nBad = 0, it;
  for (it = 0; it < P.max_iterations; it++) {
    double currentChi = sim3_robust_chi2(team, P, est, delta, dsqr, red);
    last = est;
    const double iniChi = currentChi;
    /* 28 entries of the upper triangle of H, 7 of b */
    double acc[35];
    for (int k = 0; k < 35; k++) acc[k] = 0.0;
    DS_FOR(i, P.n) {
      double e[3], y[3];
      edge_error(P, est, i, e, y);
      const double c2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
      const double r1 = c2 <= dsqr ? 1.0 : delta / sqrt(c2);
      /* J (3x7): columns 0-2 [y]x, 3-5 -I, 6 -y */
      const double J[21] = {0, -y[2], y[1], -1, 0, 0, -y[0],
                            y[2], 0, -y[0], 0, -1, 0, -y[1],
                            -y[1], y[0], 0, 0, 0, -1, -y[2]};
      int k = 0;
      for (int a = 0; a < 7; a++)
        for (int b = a; b < 7; b++) acc[k++] += r1 * (J[a] * J[b] + J[7 + a] * J[7 + b] + J[14 + a] * J[14 + b]);
      for (int a = 0; a < 7; a++) acc[28 + a] -= r1 * (J[a] * e[0] + J[7 + a] * e[1] + J[14 + a] * e[2]);
    }
    team_sum_vec(team, acc, 35, red);
    double H[49], b[7];
    {
      int k = 0;
      for (int a = 0; a < 7; a++)
        for (int c = a; c < 7; c++) { H[7 * a + c] = red[k]; H[7 * c + a] = red[k]; k++; }
      for (int a = 0; a < 7; a++) b[a] = red[28 + a];
    }
    if (it == 0) {
      double md = 0.;
      for (int k = 0; k < 7; k++) md = fmax(md, fabs(H[8 * k]));
      lambda = 1e-5 * md; ni = 2; nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      const Sim3 backup = est;
      double dx[7] = {0, 0, 0, 0, 0, 0, 0};
      const bool ok = ldlt7(H, lambda, b, dx);
      Sim3 U;
      sim3_exp(dx, U);
      sim3_mul(U, est, est);
      double tempChi = sim3_robust_chi2(team, P, est, delta, dsqr, red);
      last = est;
      if (!ok) tempChi = DBL_MAX;
      rho = currentChi - tempChi;
      double scale = 0.;
      for (int j = 0; j < 7; j++) scale += dx[j] * (lambda * dx[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && tempChi < DBL_MAX) {
        double alpha = 1. - pow(2 * rho - 1, 3);
        alpha = fmin(alpha, 2. / 3.);
        lambda *= fmax(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni; ni *= 2;
        est = backup;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0) { it++; break; }
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) { it++; break; }
  }
  return it;
}

/* Optimizer::OptimizeHorn for one keyframe.  red: sim3_red_doubles(nthr) doubles of shared memory. */
DS_FN_NOINLINE void sim3_register_one(const Team team, const Sim3Prob &P, double *red) {
  const double delta = (double)(float)sqrt(P.huber);  // const float deltaHuber = sqrt(huber)  DefOptimizer.cc:865
  const double dsqr = (double)(float)(delta * delta);  // RobustKernelHuber::dsqr is a float (robust_kernel_impl.h:84)
  double *rs = red + 36 + 36 * ((team.nthr + 31) / 32);
  Sim3 est = P.init, last = P.init;
  const int it0 = sim3_run_lm(team, P, est, last, delta, dsqr, red);
  const Sim3 first = est;
  int count = 0;
  DS_FOR(i, P.n) {
    double e[3], y[3];
    edge_error(P, last, i, e, y);
    count += !(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] > P.chi);
  }
  count = team_sum_int(team, count, rs);
  const int it1 = sim3_run_lm(team, P, est, last, delta, dsqr, red);
  double chi2 = 0.0;
  DS_FOR(i, P.n) {
    double e[3], y[3];
    edge_error(P, last, i, e, y);
    chi2 += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  }
  chi2 = team_sum(team, chi2, rs);
  if (team.tid == 0) {
    for (int k = 0; k < 4; k++) P.out[k] = first.q[k];
    for (int k = 0; k < 3; k++) P.out[4 + k] = first.t[k];
    P.out[7] = first.s;
    P.out[8] = chi2;
    P.out[9] = count;
    P.out[10] = (chi2 < DBL_MAX && chi2 > -DBL_MAX && (chi2 / count < P.chi)) ? 1.0 : 0.0;
    P.out[11] = it0;
    P.out[12] = it1;
  }
}

/* ------------------------------------------------------------------ min-median scale -- */

DS_FN bool mm_keep(uint64_t seed, uint32_t i, uint32_t j) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((((uint64_t)i) << 32) | (uint64_t)j) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) <= 0.25;
}

DS_FN float mm_residual(const float *mono, const float *stereo, double scale, int j) {
  float r2 = 0.0f;
  for (int k = 0; k < 3; k++) {
    const double r = scale * (double)mono[3 * j + k] - (double)stereo[3 * j + k];
    r2 = (float)((double)r2 + r * r);
  }
#if DS_CUDA
  return __fsqrt_rn(r2);
#else
  return sqrtf(r2);
#endif
}

/* buf: n floats of shared memory; cnt: one shared int; sc: 4 shared doubles */
DS_FN_NOINLINE void scale_min_median_team(const Team team, int n, const float *mono, const float *stereo, uint64_t seed,
                                          float *buf, int *cnt, double *sc, float *scale_out) {
  float min_med = 10000.0f;
  int final_points = 0;
  double best_scale = 0.0;
  bool empty = false;
  for (int i = 0; i < n && !empty; i++) {
    if (!mm_keep(seed, (uint32_t)i, 0xFFFFFFFFu)) continue;
    const double scale = (double)(stereo[3 * i + 2] / mono[3 * i + 2]);
    team.sync();
    if (team.tid == 0) *cnt = 0;
    team.sync();
    DS_FOR(j, n) {
      if (j == i || !mm_keep(seed, (uint32_t)i, (uint32_t)j)) continue;
      buf[atomic_inc_int(cnt)] = mm_residual(mono, stereo, scale, j);
    }
    team.sync();
    const int m = *cnt, size = m - 1;
    final_points++;
    if (size <= 0) { empty = true; break; }
    /* sorted[1 + size/2]: the reference's skip loop also drops the smallest sampled residual */
    const int want = 1 + size / 2;
    DS_FOR(a, m) {
      const float v = buf[a];
      int rank = 0;
      for (int b = 0; b < m; b++) rank += (buf[b] < v) || (buf[b] == v && b < a);
      if (rank == want) sc[0] = (double)v;
    }
    team.sync();
    const float med = (float)sc[0];
    if (med < min_med) { min_med = med; best_scale = scale; }
  }
  if (team.tid == 0) {
    if (empty) { *scale_out = 0.0f; }
    else {
      const float desv = (float)(1.4826 * (1.0 - (5.0 / (final_points - 1.0))) * sqrt((double)min_med));
      float num = 0.0f, den = 0.0f;
      for (int i = 0; i < n; i++) {
        const float residual = mm_residual(mono, stereo, best_scale, i);
#if DS_CUDA
        if ((double)__fdiv_rn(residual, desv) < 2.5) {
          num = __fadd_rn(num, __fmul_rn(stereo[3 * i + 2], mono[3 * i + 2]));
          den = __fadd_rn(den, __fmul_rn(mono[3 * i + 2], mono[3 * i + 2]));
        }
#else
        if ((double)(residual / desv) < 2.5) {
          num += stereo[3 * i + 2] * mono[3 * i + 2];
          den += mono[3 * i + 2] * mono[3 * i + 2];
        }
#endif
      }
      *scale_out = num / den;
    }
  }
}

}  // namespace ds
#endif
