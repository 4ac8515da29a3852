/*
 * sim3_oracle.c -- CPU restatement of the Sim(3) surface registration that ends DefLocalMapping::NRSfM.
 * TEST INFRASTRUCTURE ONLY (see sft_oracle.c header).
 *
 * PINNED TO THE REFERENCE'S OWN CODE for the registration: Optimizer::OptimizeHorn is re-run on the reference's
 * own sim3.h (verbatim), EdgeSim3Simple / VertexSim3ExpmapNoProj (types_seven_dof_expmap.h:96-126,159-188), the
 * numeric Jacobians of base_unary_edge.hpp (verbatim), Huber kernel and Levenberg driver (oracle/g2o_ref_harness.cc
 * -> oracle/_ref/libg2o_sft_ref.so); tests/test_oracle_sft_ref.py compares this file with it live and through
 * tests/golden/sft_ref.npz (estimate 1e-8, first-run iteration count, inliers, verdict).  scaleMinMedian draws
 * from unseeded rand() in the reference (quirk C10) and stays a seeded restatement.
 *
 * Follows (paths under the DefSLAM tree):
 *   Optimizer::OptimizeHorn                Modules/Tracking/DefOptimizer.cc:840-922
 *   EdgeSim3Simple, VertexSim3ExpmapNoProj Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h:96-188
 *   Sim3(update), map, operator*           Thirdparty/g2o/g2o/types/sim3.h:70-160,270-277
 *   BaseUnaryEdge::linearizeOplus          Thirdparty/g2o/g2o/core/base_unary_edge.hpp:82-118
 *   RobustKernelHuber::robustify           Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91
 *   OptimizationAlgorithmLevenberg::solve  Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189
 *   GroundTruthTools::scaleMinMedian       Modules/GroundTruth/GroundTruthCalculator.cc:54-159
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sft_oracle.h"

typedef struct { double q[4], t[3], s; } Sim3; /* q = (x,y,z,w) */

/* Eigen::Quaterniond(Matrix3d) */
static void q_from_R(const double R[9], double q[4]) {
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
static void q_mul(const double a[4], const double b[4], double o[4]) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
static void q_rot(const double q[4], const double v[3], double o[3]) {
  const double uv[3] = {2 * (q[1] * v[2] - q[2] * v[1]), 2 * (q[2] * v[0] - q[0] * v[2]), 2 * (q[0] * v[1] - q[1] * v[0])};
  o[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
  o[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
  o[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

/* Sim3(const Vector7d &update)  sim3.h:70-135 */
static void sim3_exp(const double u[7], Sim3 *S) {
  const double w[3] = {u[0], u[1], u[2]}, ups[3] = {u[3], u[4], u[5]}, sigma = u[6];
  const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double Om[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Om2[3 * i + j] = Om[3 * i] * Om[j] + Om[3 * i + 1] * Om[3 + j] + Om[3 * i + 2] * Om[6 + j];
  S->s = exp(sigma);
  const double eps = 0.00001;
  double A, B, C, R[9], a1 = 1.0, a2 = 1.0;
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
    C = (S->s - 1) / sigma;
    if (theta < eps) {
      const double s2 = sigma * sigma;
      A = ((sigma - 1) * S->s + 1) / s2;
      B = ((0.5 * s2 - sigma + 1) * S->s) / (s2 * sigma);
    } else {
      a1 = sin(theta) / theta; a2 = (1 - cos(theta)) / (theta * theta);
      const double a = S->s * sin(theta), b = S->s * cos(theta), th2 = theta * theta, s2 = sigma * sigma, c = th2 + s2;
      A = (a * sigma + (1 - b) * theta) / (theta * c);
      B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / th2;
    }
  }
  for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a1 * Om[i] + a2 * Om2[i];
  q_from_R(R, S->q);
  for (int i = 0; i < 3; i++) {
    S->t[i] = 0.0;
    for (int j = 0; j < 3; j++) S->t[i] += (A * Om[3 * i + j] + B * Om2[3 * i + j] + (i == j ? C : 0.0)) * ups[j];
  }
}

/* a * b  sim3.h:270-277 */
static void sim3_mul(const Sim3 *a, const Sim3 *b, Sim3 *o) {
  Sim3 r;
  q_mul(a->q, b->q, r.q);
  double rt[3];
  q_rot(a->q, b->t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a->s * rt[i] + a->t[i];
  r.s = a->s * b->s;
  *o = r;
}
static void sim3_map(const Sim3 *S, const double x[3], double o[3]) {
  double r[3];
  q_rot(S->q, x, r);
  for (int i = 0; i < 3; i++) o[i] = S->s * r[i] + S->t[i];
}

typedef struct {
  int n;
  const float *p1, *p2;
  double delta, dsqr;
  Sim3 est;  /* vertex estimate */
  double *err; /* [3n] errors of the LAST evaluation (g2o keeps them on the edges) */
} Reg;

static void compute_errors(Reg *g, const Sim3 *S) {
  for (int i = 0; i < g->n; i++) {
    const double x1[3] = {g->p1[3 * i], g->p1[3 * i + 1], g->p1[3 * i + 2]};
    double m[3];
    sim3_map(S, x1, m);
    for (int c = 0; c < 3; c++) g->err[3 * i + c] = (double)g->p2[3 * i + c] - m[c];
  }
}
static double robust_chi2(const Reg *g) {
  double chi = 0.0;
  for (int i = 0; i < g->n; i++) {
    const double *e = g->err + 3 * i;
    const double c2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    chi += c2 <= g->dsqr ? c2 : 2 * sqrt(c2) * g->delta - g->dsqr;
  }
  return chi;
}

/* numeric Jacobian of edge i at the estimate: central differences, delta 1e-9 */
static void numeric_jacobian(const Reg *g, int i, double J[21]) {
  const double delta = 1e-9, scalar = 1.0 / (2 * delta);
  const double x1[3] = {g->p1[3 * i], g->p1[3 * i + 1], g->p1[3 * i + 2]};
  for (int d = 0; d < 7; d++) {
    double add[7] = {0, 0, 0, 0, 0, 0, 0}, ep[3], em[3];
    Sim3 U, T;
    add[d] = delta;
    sim3_exp(add, &U); sim3_mul(&U, &g->est, &T); sim3_map(&T, x1, ep);
    add[d] = -delta;
    sim3_exp(add, &U); sim3_mul(&U, &g->est, &T); sim3_map(&T, x1, em);
    for (int c = 0; c < 3; c++) {
      const double e1 = (double)g->p2[3 * i + c] - ep[c], e0 = (double)g->p2[3 * i + c] - em[c];
      J[7 * c + d] = scalar * (e1 - e0);
    }
  }
}

int oracle_sim3_jacobian(const defslam_sim3_problem *p, int i, double *J21) {
  Reg g;
  g.n = p->n_points; g.p1 = p->pts1; g.p2 = p->pts2;
  memcpy(g.est.q, p->rot, sizeof(g.est.q)); memcpy(g.est.t, p->trans, sizeof(g.est.t)); g.est.s = p->scale;
  numeric_jacobian(&g, i, J21);
  return 0;
}

static int ldlt7(const double H[49], const double b[7], double x[7]) {
  double L[49], v[7];
  memset(L, 0, sizeof(L));
  for (int j = 0; j < 7; j++) {
    double dj = H[7 * j + j];
    for (int k = 0; k < j; k++) { v[k] = L[7 * j + k] * L[7 * k + k]; dj -= L[7 * j + k] * v[k]; }
    L[7 * j + j] = dj;
    if (!(dj > 0.0)) return 0;
    for (int i = j + 1; i < 7; i++) {
      double s = H[7 * i + j];
      for (int k = 0; k < j; k++) s -= L[7 * i + k] * v[k];
      L[7 * i + j] = s / dj;
    }
  }
  for (int i = 0; i < 7; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[7 * i + k] * x[k]; x[i] = s; }
  for (int i = 0; i < 7; i++) x[i] /= L[7 * i + i];
  for (int i = 6; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < 7; k++) s -= L[7 * k + i] * x[k]; x[i] = s; }
  return 1;
}

/* one optimizer.optimize(max_it); returns the iterations run */
static int run_lm(Reg *g, int max_it) {
  double lambda = -1., ni = 2.;
  int nBad = 0, it;
  const double tau = 1e-5;
  for (it = 0; it < max_it; it++) {
    compute_errors(g, &g->est);
    double currentChi = robust_chi2(g), tempChi;
    const double iniChi = currentChi;
    double H[49], b[7];
    memset(H, 0, sizeof(H)); memset(b, 0, sizeof(b));
    for (int i = 0; i < g->n; i++) {
      double J[21];
      numeric_jacobian(g, i, J);
      const double *e = g->err + 3 * i;
      const double c2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
      const double rho1 = c2 <= g->dsqr ? 1.0 : g->delta / sqrt(c2);
      for (int a = 0; a < 7; a++) {
        for (int c = 0; c < 3; c++) b[a] -= J[7 * c + a] * rho1 * e[c];
        for (int bb = 0; bb < 7; bb++)
          for (int c = 0; c < 3; c++) H[7 * a + bb] += J[7 * c + a] * rho1 * J[7 * c + bb];
      }
    }
    if (it == 0) {
      double md = 0.;
      for (int k = 0; k < 7; k++) md = fmax(md, fabs(H[8 * k]));
      lambda = tau * md; ni = 2; nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      const Sim3 backup = g->est;
      double Hl[49], dx[7] = {0, 0, 0, 0, 0, 0, 0};
      memcpy(Hl, H, sizeof(H));
      for (int k = 0; k < 7; k++) Hl[8 * k] += lambda;
      const int ok = ldlt7(Hl, b, dx);
      Sim3 U;
      sim3_exp(dx, &U);
      sim3_mul(&U, &g->est, &g->est); /* VertexSim3ExpmapNoProj::oplusImpl */
      compute_errors(g, &g->est);
      tempChi = robust_chi2(g);
      if (!ok) tempChi = DBL_MAX;
      rho = currentChi - tempChi;
      double scale = 0.;
      for (int j = 0; j < 7; j++) scale += dx[j] * (lambda * dx[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow(2 * rho - 1, 3);
        alpha = fmin(alpha, 2. / 3.);
        lambda *= fmax(1. / 3., alpha);
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni; ni *= 2;
        g->est = backup; /* pop: the edge errors keep the rejected state */
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0) { it++; break; }
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) { it++; break; }
  }
  return it;
}

int oracle_sim3_register_batched(int32_t nprob, const defslam_sim3_problem *pp, defslam_sim3_result *out, int32_t dev) {
  (void)dev;
  for (int k = 0; k < nprob; k++) {
    const defslam_sim3_problem *p = &pp[k];
    Reg g;
    g.n = p->n_points; g.p1 = p->pts1; g.p2 = p->pts2;
    g.delta = (double)(float)sqrt(p->huber); /* const float deltaHuber = sqrt(huber) */
    g.dsqr = (double)(float)(g.delta * g.delta); /* float dsqr, robust_kernel_impl.h:84 */
    memcpy(g.est.q, p->rot, sizeof(g.est.q)); memcpy(g.est.t, p->trans, sizeof(g.est.t)); g.est.s = p->scale;
    g.err = (double *)malloc(sizeof(double) * 3 * (g.n + 1));
    compute_errors(&g, &g.est);
    out[k].iterations[0] = run_lm(&g, p->max_iterations);
    memcpy(out[k].rot, g.est.q, sizeof(g.est.q)); memcpy(out[k].trans, g.est.t, sizeof(g.est.t)); out[k].scale = g.est.s;
    int count = 0;
    for (int i = 0; i < g.n; i++) {
      const double *e = g.err + 3 * i;
      if (!(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] > p->chi)) count++;
    }
    out[k].iterations[1] = run_lm(&g, p->max_iterations);
    double chi2 = 0.0;
    for (int i = 0; i < g.n; i++) {
      const double *e = g.err + 3 * i;
      chi2 += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    }
    out[k].chi2 = chi2;
    out[k].inliers = count;
    out[k].acceptable = isfinite(chi2) && (chi2 / count < p->chi);
    free(g.err);
  }
  return 0;
}

/* ------------------------------------------------------------------ min-median scale -- */

/* the Bernoulli(0.25) draw that replaces `rand()/RAND_MAX > 0.25 -> skip` */
static int mm_keep(uint64_t seed, uint32_t i, uint32_t j) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((((uint64_t)i) << 32) | (uint64_t)j) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) <= 0.25;
}

static int cmp_f(const void *a, const void *b) {
  const float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

int oracle_scale_min_median(int32_t n, const float *mono, const float *stereo, uint64_t seed, float *scale_out) {
  float min_med = 10000.0f;
  int final_points = 0;
  double best_scale = 0.0;
  float *res = (float *)malloc(sizeof(float) * (n + 1));
  for (int i = 0; i < n; i++) {
    if (!mm_keep(seed, (uint32_t)i, 0xFFFFFFFFu)) continue;
    const double scale = stereo[3 * i + 2] / mono[3 * i + 2];
    int m = 0;
    for (int j = 0; j < n; j++) {
      if (i == j) continue;
      if (!mm_keep(seed, (uint32_t)i, (uint32_t)j)) continue;
      float r2 = 0.0f;
      for (int k = 0; k < 3; k++) {
        const double r = scale * mono[3 * j + k] - stereo[3 * j + k];
        r2 = (float)(r2 + r * r);
      }
      res[m++] = sqrtf(r2);
    }
    qsort(res, m, sizeof(float), cmp_f);
    /* the reference's skip loop also drops the smallest sampled residual (off by one, :103-107) */
    const int size = m - 1;
    final_points++;
    if (size <= 0) { free(res); *scale_out = 0.0f; return 0; }
    const float med = res[1 + size / 2];
    if (med < min_med) { min_med = med; best_scale = scale; }
  }
  const float desv = (float)(1.4826 * (1.0 - (5.0 / (final_points - 1.0))) * sqrt(min_med));
  float num = 0.0f, den = 0.0f;
  for (int i = 0; i < n; i++) {
    float residual = 0.0f;
    for (int k = 0; k < 3; k++) {
      const double r = best_scale * mono[3 * i + k] - stereo[3 * i + k];
      residual = (float)(residual + r * r);
    }
    residual = sqrtf(residual);
    if ((residual / desv) < 2.5) {
      num += stereo[3 * i + 2] * mono[3 * i + 2];
      den += mono[3 * i + 2] * mono[3 * i + 2];
    }
  }
  free(res);
  *scale_out = (float)((double)(num / den));
  return 0;
}
