/*
 * nrsfm_oracle.c -- CPU restatement of the NRSfM mapping stages (Schwarp fit, isometric
 * normals, shape-from-normals).  TEST INFRASTRUCTURE ONLY (see sft_oracle.c header): only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may call it.
 *
 * PARITY UNPINNED BY THE REFERENCE for the iterates: the reference has no tests/fixtures
 * for this path and solves with Ceres (version unpinned, absent here) and Eigen (absent).
 * What is pinned exactly: the cost functions and Jacobians (finite differences in
 * tests/test_oracle_nrsfm.py), the BBS arithmetic underneath (bbs_oracle.c, pinned to the
 * reference's own bbs.cc), the polynomial coefficients (isometric synthetic pair).
 * The trust-region loop below restates Ceres' published Levenberg-Marquardt
 * (TrustRegionMinimizer + LevenbergMarquardtStrategy, default options) and is dense
 * throughout, with the reference's cost structure (dense Jacobians, dense products).
 *
 * Follows (paths under the DefSLAM tree):
 *   Warps::Warp::{Warp,Evaluate,initialize,getEstimates}  Modules/Mapping/Schwarp.cc:38-303
 *   Warps::Schwarzian::{Schwarzian,Evaluate}              Modules/Mapping/Schwarp.cc:305-543
 *   SchwarpDatabase::calculateSchwarps                    Modules/Mapping/SchwarpDatabase.cc:145-349
 *   NormalEstimator::ObtainK1K2                           Modules/Mapping/NormalEstimator.cc:38-229
 *   PolySolver::{getCoefficients,Evaluate}                Modules/Mapping/PolySolver.cc:50-193
 *   ShapeFromNormals::{ctor,obtainM,estimate}             Modules/Mapping/ShapeFromNormals.cc:38-260
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sft_oracle.h"

/* ------------------------------------------------------------------ helpers ---------- */

/* control-point grid sites of Schwarp.cc:321-331 / 399-409.  A site that rounding pushes
 * past the domain end is clamped onto it (the reference would index out of range there). */
static void grid_sites(const defslam_bbs *s, double *X, double *Y) {
  int us = 0;
  for (int i = 0; i < s->nptsu; i++)
    for (int j = 0; j < s->nptsv; j++) {
      double x = (double)((s->umax - s->umin) * i) / (s->nptsu - 1) + s->umin;
      double y = (double)((s->vmax - s->vmin) * j) / (s->nptsv - 1) + s->vmin;
      if (x > s->umax) x = s->umax;
      if (y > s->vmax) y = s->vmax;
      if (x < s->umin) x = s->umin;
      if (y < s->vmin) y = s->vmin;
      X[us] = x;
      Y[us] = y;
      us++;
    }
}

/* dense Cholesky A = L L^T in place (lower); returns 0, or 1 if a pivot is not positive */
static int chol_dense(double *A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !isfinite(d)) return 1;
    const double l = sqrt(d);
    A[(size_t)j * n + j] = l;
    for (int i = j + 1; i < n; i++) {
      double s = A[(size_t)i * n + j];
      const double *ai = A + (size_t)i * n, *aj = A + (size_t)j * n;
      for (int k = 0; k < j; k++) s -= ai[k] * aj[k];
      A[(size_t)i * n + j] = s / l;
    }
  }
  return 0;
}

static void chol_solve(const double *L, int n, double *b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}

/* ------------------------------------------------------------------ Schwarp ---------- */

/* x is [all x; all y] (column-major NC x 2, Schwarp.cc:240-244); BBS wants valdim-interleaved */
static void ctrl_interleave(const double *x, int NC, double *arr) {
  for (int n = 0; n < 2; n++)
    for (int l = 0; l < NC; l++) arr[2 * l + n] = x[n * NC + l];
}

/* residuals (2n + 4NC) and, if J != NULL, the dense Jacobian ((2n+4NC) x 2NC), at p->x */
int oracle_schwarp_evaluate(const defslam_schwarp_problem *p, double *r, double *J) {
  const defslam_bbs *s = &p->bbs;
  const int NC = s->nptsu * s->nptsv, n = p->n_matches, NP = 2 * NC;
  if (s->valdim != 2 || n < 0) return DEFSLAM_EBADARG;
  double *arr = (double *)malloc(sizeof(double) * NP);
  ctrl_interleave(p->x, NC, arr);
  int rc = 0;
  /* ---- data term: Warp::Evaluate Schwarp.cc:235-303 */
  double *u = (double *)malloc(sizeof(double) * (n + 1)), *v = (double *)malloc(sizeof(double) * (n + 1));
  double *val = (double *)malloc(sizeof(double) * 2 * (n + 1));
  for (int i = 0; i < n; i++) { u[i] = p->kp1[2 * i]; v[i] = p->kp1[2 * i + 1]; }
  oracle_bbs_eval(s, arr, n, u, v, 0, 0, val);
  for (int i = 0; i < n; i++) {
    r[i] = p->inv_sigma[i] * ((double)p->kp2[2 * i] - val[2 * i]) * p->fx;
    r[i + n] = p->inv_sigma[i] * ((double)p->kp2[2 * i + 1] - val[2 * i + 1]) * p->fy;
  }
  if (J) {
    memset(J, 0, sizeof(double) * (size_t)(2 * n + 4 * NC) * NP);
    double *C = (double *)malloc(sizeof(double) * (size_t)(n + 1) * NC);
    rc = oracle_bbs_coloc(s, n, u, v, 0, 0, C);
    /* Jdata = -[C fx, 0; 0, C fy] (Schwarp.cc:76-83); Evaluate then hands Ceres row i of Jdata
     * for BOTH residual i and residual i+n (Schwarp.cc:291-299, quirk C6) */
    for (int i = 0; i < n; i++)
      for (int j = 0; j < NC; j++) {
        const double e = -C[(size_t)i * NC + j] * p->fx;
        J[(size_t)i * NP + j] = e;
        J[(size_t)(i + n) * NP + j] = e;
      }
    free(C);
  }
  free(u); free(v); free(val);
  /* ---- Schwarzian term: Schwarp.cc:368-543 */
  double *X = (double *)malloc(sizeof(double) * NC), *Y = (double *)malloc(sizeof(double) * NC);
  grid_sites(s, X, Y);
  double *d10 = (double *)malloc(sizeof(double) * 2 * NC), *d01 = (double *)malloc(sizeof(double) * 2 * NC);
  double *d20 = (double *)malloc(sizeof(double) * 2 * NC), *d02 = (double *)malloc(sizeof(double) * 2 * NC);
  double *d11 = (double *)malloc(sizeof(double) * 2 * NC);
  oracle_bbs_eval(s, arr, NC, X, Y, 1, 0, d10);
  oracle_bbs_eval(s, arr, NC, X, Y, 0, 1, d01);
  oracle_bbs_eval(s, arr, NC, X, Y, 2, 0, d20);
  oracle_bbs_eval(s, arr, NC, X, Y, 0, 2, d02);
  oracle_bbs_eval(s, arr, NC, X, Y, 1, 1, d11);
  double *rs = r + 2 * n;
  const double lam = p->lambda;
  for (int k = 0; k < NC; k++) {
    const double xu = d10[2 * k], yu = d10[2 * k + 1], xv = d01[2 * k], yv = d01[2 * k + 1];
    const double xuu = d20[2 * k], yuu = d20[2 * k + 1], xvv = d02[2 * k], yvv = d02[2 * k + 1];
    const double xuv = d11[2 * k], yuv = d11[2 * k + 1];
    rs[k] = (xuu * yu - yuu * xu) * lam;
    rs[NC + k] = (yvv * xv - xvv * yv) * lam;
    rs[2 * NC + k] = (xuu * yv - yuu * xv + 2 * (xuv * yu - yuv * xu)) * lam;
    rs[3 * NC + k] = (yvv * xu - xvv * yu + 2 * (yuv * xv - xuv * yv)) * lam;
  }
  if (J) {
    /* A = diag(first derivatives) * [Cuu Cuv Cvv], B = diag(second derivatives) * [Cu Cv]
     * and their block differences (Schwarp.cc:461-512), written row by row */
    const size_t sz = (size_t)NC * NC;
    double *Cu = (double *)malloc(sizeof(double) * sz), *Cv = (double *)malloc(sizeof(double) * sz);
    double *Cuu = (double *)malloc(sizeof(double) * sz), *Cvv = (double *)malloc(sizeof(double) * sz);
    double *Cuv = (double *)malloc(sizeof(double) * sz);
    rc |= oracle_bbs_coloc(s, NC, X, Y, 1, 0, Cu);
    rc |= oracle_bbs_coloc(s, NC, X, Y, 0, 1, Cv);
    rc |= oracle_bbs_coloc(s, NC, X, Y, 2, 0, Cuu);
    rc |= oracle_bbs_coloc(s, NC, X, Y, 0, 2, Cvv);
    rc |= oracle_bbs_coloc(s, NC, X, Y, 1, 1, Cuv);
    double *Js = J + (size_t)2 * n * NP;
    for (int k = 0; k < NC; k++) {
      const double xu = d10[2 * k], yu = d10[2 * k + 1], xv = d01[2 * k], yv = d01[2 * k + 1];
      const double xuu = d20[2 * k], yuu = d20[2 * k + 1], xvv = d02[2 * k], yvv = d02[2 * k + 1];
      const double xuv = d11[2 * k], yuv = d11[2 * k + 1];
      for (int c = 0; c < NC; c++) {
        const double cu = Cu[(size_t)k * NC + c], cv = Cv[(size_t)k * NC + c];
        const double cuu = Cuu[(size_t)k * NC + c], cvv = Cvv[(size_t)k * NC + c], cuv = Cuv[(size_t)k * NC + c];
        /* jI */
        Js[(size_t)k * NP + c] = lam * (yu * cuu - yuu * cu);
        Js[(size_t)k * NP + NC + c] = lam * (xuu * cu - xu * cuu);
        /* jJ */
        Js[(size_t)(NC + k) * NP + c] = lam * (yvv * cv - yv * cvv);
        Js[(size_t)(NC + k) * NP + NC + c] = lam * (xv * cvv - xvv * cv);
        /* jM */
        Js[(size_t)(2 * NC + k) * NP + c] = lam * (yv * cuu - yuu * cv + 2 * yu * cuv - 2 * yuv * cu);
        Js[(size_t)(2 * NC + k) * NP + NC + c] = lam * (xuu * cv - xv * cuu + 2 * xuv * cu - 2 * xu * cuv);
        /* jN */
        Js[(size_t)(3 * NC + k) * NP + c] = lam * (yvv * cu - yu * cvv - 2 * yv * cuv + 2 * yuv * cv);
        Js[(size_t)(3 * NC + k) * NP + NC + c] = lam * (xu * cvv - xvv * cu - 2 * xuv * cv + 2 * xv * cuv);
      }
    }
    free(Cu); free(Cv); free(Cuu); free(Cvv); free(Cuv);
  }
  free(X); free(Y); free(d10); free(d01); free(d20); free(d02); free(d11); free(arr);
  return rc ? DEFSLAM_EBADARG : 0;
}

/* Warp::initialize  Schwarp.cc:99-160:  x0 = (C'C + lambda B)^-1 C' q2 */
int oracle_schwarp_init(const defslam_schwarp_problem *p, double *x0) {
  const defslam_bbs *s = &p->bbs;
  const int NC = s->nptsu * s->nptsv, n = p->n_matches;
  double *u = (double *)malloc(sizeof(double) * (n + 1)), *v = (double *)malloc(sizeof(double) * (n + 1));
  for (int i = 0; i < n; i++) { u[i] = p->kp1[2 * i]; v[i] = p->kp1[2 * i + 1]; }
  double *C = (double *)malloc(sizeof(double) * (size_t)(n + 1) * NC);
  int rc = oracle_bbs_coloc(s, n, u, v, 0, 0, C);
  double *A = (double *)malloc(sizeof(double) * (size_t)NC * NC);
  oracle_bbs_bending(s, A);
  for (size_t i = 0; i < (size_t)NC * NC; i++) A[i] *= p->lambda;
  for (int i = 0; i < n; i++) {
    const double *ci = C + (size_t)i * NC;
    for (int a = 0; a < NC; a++) {
      if (ci[a] == 0.0) continue;
      for (int b = 0; b < NC; b++) A[(size_t)a * NC + b] += ci[a] * ci[b];
    }
  }
  for (int d = 0; d < 2; d++) {
    double *rhs = x0 + (size_t)d * NC;
    for (int a = 0; a < NC; a++) rhs[a] = 0.0;
    for (int i = 0; i < n; i++)
      for (int a = 0; a < NC; a++) rhs[a] += C[(size_t)i * NC + a] * (double)p->kp2[2 * i + d];
  }
  if (!rc) {
    if (chol_dense(A, NC)) rc = DEFSLAM_ENUMERIC;
    else { chol_solve(A, NC, x0); chol_solve(A, NC, x0 + NC); }
  }
  free(u); free(v); free(C); free(A);
  return rc;
}

/* Ceres' loss corrector for one residual block with rho'' <= 0 (Huber): residuals and
 * Jacobian rows are scaled by sqrt(rho').  HuberLoss(a): rho(s) = s (s <= a^2), else
 * 2 a sqrt(s) - a^2; rho' = 1 or a / sqrt(s). */
static double huber_rho(double s, double a, double *rho1) {
  const double b = a * a;
  if (s > b) {
    const double r = sqrt(s);
    *rho1 = a / r > DBL_MIN ? a / r : DBL_MIN;
    return 2.0 * a * r - b;
  }
  *rho1 = 1.0;
  return s;
}

#define SCHWARP_HUBER 5.77

static double schwarp_cost(const defslam_schwarp_problem *p, const double *r, double *rho1) {
  const int NC = p->bbs.nptsu * p->bbs.nptsv, n = p->n_matches;
  double sd = 0.0, ss = 0.0;
  for (int i = 0; i < 2 * n; i++) sd += r[i] * r[i];
  for (int i = 0; i < 4 * NC; i++) ss += r[2 * n + i] * r[2 * n + i];
  return 0.5 * (huber_rho(sd, SCHWARP_HUBER, rho1) + ss);
}

/* Ceres trust-region options used by both solves (defaults unless the call site sets them) */
typedef struct {
  int max_iterations;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
} lm_options;

#define LM_INITIAL_RADIUS 1e4
#define LM_MAX_RADIUS 1e16
#define LM_MIN_RADIUS 1e-32
#define LM_MIN_DIAG 1e-6
#define LM_MAX_DIAG 1e32
#define LM_MIN_REL_DECREASE 1e-3
#define LM_MAX_INVALID 5

/* DefORBmatcher::CalculateInitialSchwarp  Modules/Matching/DefORBmatcher.cc:111-187 */
int oracle_schwarp_initial(const defslam_schwarp_problem *p, uint8_t *keep_out, double *err_out) {
  const defslam_bbs *s = &p->bbs;
  const int NC = s->nptsu * s->nptsv, n = p->n_matches, NP = 2 * NC, NR = 2 * n + 4 * NC;
  if (s->valdim != 2 || n <= 0 || !p->x || !keep_out) return DEFSLAM_EBADARG;
  /* x[i] = 0 for the first NCu*NCu*2 entries (:138-142) is overwritten by initialize (:146-147) */
  int rc = oracle_schwarp_init(p, p->x);
  if (rc) return rc;
  int nscrub = s->nptsu * s->nptsu * 2; /* :149 -- NCu*NCu, quirk C9 */
  if (nscrub > NP) nscrub = NP;
  for (int i = 0; i < nscrub; i++)
    if (isnan(p->x[i])) p->x[i] = 0.0;
  double *r = (double *)malloc(sizeof(double) * NR);
  rc = oracle_schwarp_evaluate(p, r, NULL);
  if (rc) { free(r); return rc; }
  /* problem.Evaluate with default options applies HuberLoss(5.77) to the block: Ceres' Corrector scales the
   * residuals by sqrt(rho'(s)), s = squared norm of the block (rho'' <= 0 for Huber: no alpha correction) */
  double sq = 0.0;
  for (int i = 0; i < 2 * n; i++) sq += r[i] * r[i];
  const double delta = 5.77;
  const double rho1 = sq <= delta * delta ? 1.0 : delta / sqrt(sq);
  for (int i = 0; i < n; i++) {
    /* residuals[2*i]^2 + residuals[2*i+1]^2 on the [x block; y block] layout (:171-176) */
    const double e = rho1 * (r[2 * i] * r[2 * i] + r[2 * i + 1] * r[2 * i + 1]);
    if (err_out) err_out[i] = e;
    keep_out[i] = e > 20 ? 0 : 1;
  }
  free(r);
  return 0;
}

int oracle_schwarp_fit(const defslam_schwarp_problem *p, defslam_diffprop *out) {
  const defslam_bbs *s = &p->bbs;
  const int NC = s->nptsu * s->nptsv, n = p->n_matches, NP = 2 * NC, NR = 2 * n + 4 * NC;
  if (s->valdim != 2 || n <= 0 || !p->x) return DEFSLAM_EBADARG;
  int rc = 0;
  if (p->initialize) {
    rc = oracle_schwarp_init(p, p->x);
    if (rc) return rc;
  }
  double *r = (double *)malloc(sizeof(double) * NR), *rc_ = (double *)malloc(sizeof(double) * NR);
  double *J = (double *)malloc(sizeof(double) * (size_t)NR * NP);
  double *H = (double *)malloc(sizeof(double) * (size_t)NP * NP), *A = (double *)malloc(sizeof(double) * (size_t)NP * NP);
  double *g = (double *)malloc(sizeof(double) * NP), *scale = (double *)malloc(sizeof(double) * NP);
  double *step = (double *)malloc(sizeof(double) * NP), *xc = (double *)malloc(sizeof(double) * NP);
  double *x = p->x;
  defslam_schwarp_problem q = *p;
  double rho1, radius = LM_INITIAL_RADIUS, decrease = 2.0;
  int have_scale = 0, need_eval = 1, iters = 0, accepted = 0, invalid = 0;
  double cost = 0.0;
  const lm_options opt = {p->max_iterations, 1e-6, 1e-10, 1e-8};
  for (;;) {
    if (need_eval) {
      q.x = x;
      rc = oracle_schwarp_evaluate(&q, r, J);
      if (rc) break;
      cost = schwarp_cost(p, r, &rho1);
      if (iters == 0 && out) out->cost_initial = cost;
      /* corrector on the data block */
      const double sq = sqrt(rho1);
      for (int i = 0; i < 2 * n; i++) {
        r[i] *= sq;
        for (int j = 0; j < NP; j++) J[(size_t)i * NP + j] *= sq;
      }
      /* gradient (unscaled), Jacobi scaling from the first Jacobian, then scale J */
      double gmax = 0.0;
      for (int j = 0; j < NP; j++) g[j] = 0.0;
      for (int i = 0; i < NR; i++) {
        const double ri = r[i];
        const double *Ji = J + (size_t)i * NP;
        for (int j = 0; j < NP; j++) g[j] += Ji[j] * ri;
      }
      for (int j = 0; j < NP; j++) gmax = fmax(gmax, fabs(g[j]));
      if (!have_scale) {
        for (int j = 0; j < NP; j++) scale[j] = 0.0;
        for (int i = 0; i < NR; i++)
          for (int j = 0; j < NP; j++) scale[j] += J[(size_t)i * NP + j] * J[(size_t)i * NP + j];
        for (int j = 0; j < NP; j++) scale[j] = 1.0 / (1.0 + sqrt(scale[j]));
        have_scale = 1;
      }
      for (int i = 0; i < NR; i++)
        for (int j = 0; j < NP; j++) J[(size_t)i * NP + j] *= scale[j];
      /* H = J'J of the scaled Jacobian */
      memset(H, 0, sizeof(double) * (size_t)NP * NP);
      for (int i = 0; i < NR; i++) {
        const double *Ji = J + (size_t)i * NP;
        for (int a = 0; a < NP; a++) {
          if (Ji[a] == 0.0) continue;
          const double ja = Ji[a];
          double *Ha = H + (size_t)a * NP;
          for (int b = 0; b <= a; b++) Ha[b] += ja * Ji[b];
        }
      }
      need_eval = 0;
      if (gmax <= opt.gradient_tolerance) break;
    }
    if (iters >= opt.max_iterations) break;
    if (radius < LM_MIN_RADIUS) break;
    iters++;
    /* LevenbergMarquardtStrategy::ComputeStep: (J'J + D'D) y = J'r, step = -y */
    for (int a = 0; a < NP; a++)
      for (int b = 0; b <= a; b++) A[(size_t)a * NP + b] = H[(size_t)a * NP + b];
    for (int a = 0; a < NP; a++) {
      double d = H[(size_t)a * NP + a];
      d = fmin(fmax(d, LM_MIN_DIAG), LM_MAX_DIAG);
      A[(size_t)a * NP + a] += d / radius;
    }
    int valid = !chol_dense(A, NP);
    double model_change = 0.0;
    if (valid) {
      for (int j = 0; j < NP; j++) step[j] = g[j] * scale[j];
      chol_solve(A, NP, step);
      for (int j = 0; j < NP; j++) {
        step[j] = -step[j];
        if (!isfinite(step[j])) valid = 0;
      }
    }
    if (valid) {
      /* model_cost_change = -(J step).(r + J step / 2) */
      for (int i = 0; i < NR; i++) {
        const double *Ji = J + (size_t)i * NP;
        double m = 0.0;
        for (int j = 0; j < NP; j++) m += Ji[j] * step[j];
        model_change -= m * (r[i] + 0.5 * m);
      }
      valid = model_change > 0.0;
    }
    if (!valid) {
      if (++invalid >= LM_MAX_INVALID) break;
      radius /= decrease;
      decrease *= 2.0;
      continue;
    }
    invalid = 0;
    double xnorm = 0.0, snorm = 0.0;
    for (int j = 0; j < NP; j++) {
      const double d = step[j] * scale[j];
      xc[j] = x[j] + d;
      xnorm += x[j] * x[j];
      snorm += d * d;
    }
    q.x = xc;
    double rho1c, cost_c = DBL_MAX;
    if (!oracle_schwarp_evaluate(&q, rc_, NULL)) {
      cost_c = schwarp_cost(p, rc_, &rho1c);
      if (!isfinite(cost_c)) cost_c = DBL_MAX;
    }
    if (sqrt(snorm) <= opt.parameter_tolerance * (sqrt(xnorm) + opt.parameter_tolerance)) break;
    if (fabs(cost - cost_c) <= opt.function_tolerance * cost) break;
    const double rel = (cost - cost_c) / model_change;
    if (rel > LM_MIN_REL_DECREASE) {
      memcpy(x, xc, sizeof(double) * NP);
      accepted++;
      const double t = 2.0 * rel - 1.0;
      radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
      radius = fmin(LM_MAX_RADIUS, radius);
      decrease = 2.0;
      need_eval = 1;
    } else {
      radius /= decrease;
      decrease *= 2.0;
    }
  }
  if (out) {
    out->cost_final = cost;
    out->iterations = iters;
    out->accepted = accepted;
  }
  /* ---- DiffProp records: SchwarpDatabase.cc:243-345 */
  if (!rc && out && out->warp_uv) {
    double *arr = (double *)malloc(sizeof(double) * NP);
    ctrl_interleave(x, NC, arr);
    double *u = (double *)malloc(sizeof(double) * n), *v = (double *)malloc(sizeof(double) * n);
    double *val = (double *)malloc(sizeof(double) * 2 * n * 6);
    for (int i = 0; i < n; i++) { u[i] = p->kp1[2 * i]; v[i] = p->kp1[2 * i + 1]; }
    static const int ord[6][2] = {{0, 0}, {1, 0}, {0, 1}, {2, 0}, {1, 1}, {0, 2}};
    for (int k = 0; k < 6; k++) oracle_bbs_eval(s, arr, n, u, v, ord[k][0], ord[k][1], val + (size_t)k * 2 * n);
    for (int i = 0; i < n; i++) {
      /* every quantity passes through a cv::KeyPoint (fp32) */
      const float qx = (float)val[2 * i], qy = (float)val[2 * i + 1];
      const float dux = (float)val[2 * n + 2 * i], duy = (float)val[2 * n + 2 * i + 1];
      const float dvx = (float)val[4 * n + 2 * i], dvy = (float)val[4 * n + 2 * i + 1];
      out->warp_uv[2 * i] = qx;
      out->warp_uv[2 * i + 1] = qy;
      float ex = qx - p->kp2[2 * i], ey = qy - p->kp2[2 * i + 1];
      ex *= (float)p->px_fx;
      ey *= (float)p->px_fy;
      if (out->keep) out->keep[i] = !(sqrt((double)ex * ex + (double)ey * ey) > 10);
      if (out->J12) {
        out->J12[4 * i] = dux; out->J12[4 * i + 1] = duy; out->J12[4 * i + 2] = dvx; out->J12[4 * i + 3] = dvy;
      }
      if (out->J21) {
        volatile float p1 = dux * dvy, p2 = dvx * duy; /* no fused multiply-add across the difference */
        const float det = p1 - p2;
        out->J21[4 * i] = dvy / det;
        out->J21[4 * i + 1] = -dvx / det;
        out->J21[4 * i + 2] = -duy / det;
        out->J21[4 * i + 3] = dux / det;
      }
      if (out->H12) {
        out->H12[6 * i] = (float)val[6 * n + 2 * i];       /* uux */
        out->H12[6 * i + 1] = (float)val[6 * n + 2 * i + 1]; /* uuy */
        out->H12[6 * i + 2] = (float)val[8 * n + 2 * i];   /* uvx */
        out->H12[6 * i + 3] = (float)val[8 * n + 2 * i + 1];
        out->H12[6 * i + 4] = (float)val[10 * n + 2 * i];  /* vvx */
        out->H12[6 * i + 5] = (float)val[10 * n + 2 * i + 1];
      }
    }
    free(arr); free(u); free(v); free(val);
  }
  free(r); free(rc_); free(J); free(H); free(A); free(g); free(scale); free(step); free(xc);
  return rc;
}

/* ------------------------------------------------------------------ normals ---------- */

/* PolySolver::getCoefficients, both polynomials  (PolySolver.cc:50-149).  Coefficient order
 * [x^3, x^2 y, x y^2, y^3, x^2, x y, y^2, x, y, 1]. */
static void poly_coefficients(double a, double b, double c, double d, double t1, double t2, double e1, double e2,
                              double x1, double y1, double x2, double y2, double *q1, double *q2) {
  const double D = a * d - c * b, D2 = D * D;
  const double P = a * x2 + b * y2, Q = c * x2 + d * y2, m = a * c + b * d;
  const double na = a * a + b * b, nc = c * c + d * d;
  const double w = a * x2 * y1 - c * x1 * x2 + b * y1 * y2 - d * x1 * y2;
  const double ee = e1 * e2;
  q1[0] = D * (t1 * ee - D * (e1 * Q - y1 * e2));
  q1[1] = -D * (t2 * ee - D * (e1 * P - x1 * e2));
  q1[2] = 0.0;
  q1[3] = 0.0;
  q1[4] = t2 * (ee * t1 - D * (e1 * Q - 2 * e2 * y1)) - t1 * D * (e1 * P + 2 * e2 * x1) + D2 * (e1 * m - 2 * w);
  q1[5] = e1 * (-e2 * t2 * t2 + 2 * t2 * D * P - na * D2) + e2 * D2;
  q1[6] = 0.0;
  q1[7] = t1 * (e2 * D + 2 * x1 * D * P) - 2 * t2 * (e2 * x1 * t1 + D * w) + e2 * y1 * t2 * t2 +
          D2 * (-2 * x1 * m + y1 * na - Q);
  q1[8] = t2 * D * (e2 - 2 * x1 * P) + x1 * e2 * t2 * t2 + D2 * (x1 * na - P);
  q1[9] = t2 * (e2 * t1 - D * Q) - t1 * D * P + m * D2;

  q2[0] = 0.0;
  q2[1] = 0.0;
  q2[2] = -D * (ee * t1 - D * (e1 * Q - e2 * y1));
  q2[3] = D * (ee * t2 - D * (e1 * P - e2 * x1));
  q2[4] = 0.0;
  q2[5] = e1 * (-e2 * t1 * t1 + D * (2 * t1 * Q - nc * D)) + e2 * D2;
  q2[6] = t2 * (ee * t1 - D * (e1 * Q + 2 * e2 * y1)) - t1 * D * (e1 * P - 2 * e2 * x1) + D2 * (e1 * m + 2 * w);
  q2[7] = t1 * D * (e2 - 2 * y1 * Q) + y1 * (e2 * t1 * t1 + D2 * nc) - D2 * Q;
  q2[8] = t2 * (e2 * D + 2 * y1 * D * Q) + t1 * (-2 * e2 * y1 * t2 + 2 * D * w) + e2 * x1 * t1 * t1 -
          2 * D2 * (m * y1 + 0.5 * P - 0.5 * nc * x1);
  q2[9] = q1[9];
}

/* the fp32 pre-computation of NormalEstimator.cc:88-103 followed by getCoefficients */
static void pair_coefficients(const float *J12, const float *H12, const float *I1, const float *I2, int corrected_t2,
                              double *q1, double *q2) {
  const float a = J12[0], b = J12[1], c = J12[2], d = J12[3];
  const float vvx = H12[4], vvy = H12[5];
  volatile float m1 = -b * vvx / 2, m2 = a * vvy / 2, m3 = -(d * vvx) / 2, m4 = (c * vvy) / 2;
  if (corrected_t2) { m3 = (d * H12[0]) / 2; m4 = -((c * H12[1]) / 2); } /* the transfer's t2, :210 */
  const float t1 = m1 + m2, t2 = m3 + m4;
  volatile float s1 = I1[0] * I1[0], s2 = I1[1] * I1[1], s3 = I2[0] * I2[0], s4 = I2[1] * I2[1];
  const float e1 = 1 + s1 + s2, e2 = 1 + s3 + s4;
  poly_coefficients(a, b, c, d, t1, t2, e1, e2, I1[0], I1[1], I2[0], I2[1], q1, q2);
}

int oracle_polysolver_coefficients(int32_t npairs, const float *J12, const float *H12, const float *I1,
                                   const float *I2, double *eq1, double *eq2) {
  for (int i = 0; i < npairs; i++)
    pair_coefficients(J12 + 4 * i, H12 + 6 * i, I1 + 2 * i, I2 + 2 * i, 0, eq1 + 10 * i, eq2 + 10 * i);
  return 0;
}

/* PolySolver::Evaluate  PolySolver.cc:152-193 */
static void poly_eval(const double *q, double x, double y, double *e, double *jx, double *jy) {
  *e = q[0] * x * x * x + q[1] * x * x * y + q[2] * x * y * y + q[3] * y * y * y + q[4] * x * x + q[5] * x * y +
       q[6] * y * y + q[7] * x + q[8] * y + q[9];
  if (jx) {
    *jx = 3 * q[0] * x * x + 2 * q[1] * x * y + q[2] * y * y + 2 * q[4] * x + q[5] * y + q[7];
    *jy = q[1] * x * x + 2 * q[2] * x * y + 3 * q[3] * y * y + q[5] * x + 2 * q[6] * y + q[8];
  }
}

/* cost, and optionally gradient + J'J of the stacked 2-row blocks */
static double normals_cost(const double *Q, int np, const double *x, double *g, double *H) {
  double c = 0.0;
  if (g) { g[0] = g[1] = 0.0; H[0] = H[1] = H[2] = 0.0; }
  for (int i = 0; i < np; i++)
    for (int k = 0; k < 2; k++) {
      double e, jx, jy;
      poly_eval(Q + 20 * i + 10 * k, x[0], x[1], &e, g ? &jx : NULL, g ? &jy : NULL);
      c += e * e;
      if (g) {
        g[0] += jx * e; g[1] += jy * e;
        H[0] += jx * jx; H[1] += jx * jy; H[2] += jy * jy;
      }
    }
  return 0.5 * c;
}

/* one point: Ceres LM (dense normal Cholesky) with the options of NormalEstimator.cc:137-149.
 * returns the number of trust-region steps */
static int normals_solve_point(const double *Q, int np, double *x, int max_iterations) {
  const lm_options opt = {max_iterations, 1e-10, 1e-8, 1e-8};
  double g[2], H[3], scale[2] = {1, 1}, radius = LM_INITIAL_RADIUS, decrease = 2.0, cost = 0.0;
  int have_scale = 0, need_eval = 1, iters = 0, invalid = 0;
  for (;;) {
    if (need_eval) {
      cost = normals_cost(Q, np, x, g, H);
      if (!have_scale) {
        scale[0] = 1.0 / (1.0 + sqrt(H[0]));
        scale[1] = 1.0 / (1.0 + sqrt(H[2]));
        have_scale = 1;
      }
      need_eval = 0;
      if (fmax(fabs(g[0]), fabs(g[1])) <= opt.gradient_tolerance) break;
    }
    if (iters >= opt.max_iterations) break;
    if (radius < LM_MIN_RADIUS) break;
    iters++;
    /* scaled system */
    const double h00 = H[0] * scale[0] * scale[0], h01 = H[1] * scale[0] * scale[1], h11 = H[2] * scale[1] * scale[1];
    const double g0 = g[0] * scale[0], g1 = g[1] * scale[1];
    const double a00 = h00 + fmin(fmax(h00, LM_MIN_DIAG), LM_MAX_DIAG) / radius;
    const double a11 = h11 + fmin(fmax(h11, LM_MIN_DIAG), LM_MAX_DIAG) / radius;
    int valid = a00 > 0.0;
    double s0 = 0, s1 = 0, model_change = 0.0;
    if (valid) {
      const double l00 = sqrt(a00), l10 = h01 / l00, dd = a11 - l10 * l10;
      valid = dd > 0.0;
      if (valid) {
        const double l11 = sqrt(dd);
        const double y0 = g0 / l00, y1 = (g1 - l10 * y0) / l11;
        s1 = y1 / l11;
        s0 = (y0 - l10 * s1) / l00;
        s0 = -s0; s1 = -s1;
        valid = isfinite(s0) && isfinite(s1);
      }
    }
    if (valid) {
      /* -(J s).(r + J s/2) = -s'g - s'Hs/2 on the scaled quantities */
      model_change = -(s0 * g0 + s1 * g1) - 0.5 * (s0 * (h00 * s0 + h01 * s1) + s1 * (h01 * s0 + h11 * s1));
      valid = model_change > 0.0;
    }
    if (!valid) {
      if (++invalid >= LM_MAX_INVALID) break;
      radius /= decrease; decrease *= 2.0;
      continue;
    }
    invalid = 0;
    const double d0 = s0 * scale[0], d1 = s1 * scale[1];
    const double xc[2] = {x[0] + d0, x[1] + d1};
    double cost_c = normals_cost(Q, np, xc, NULL, NULL);
    if (!isfinite(cost_c)) cost_c = DBL_MAX;
    if (sqrt(d0 * d0 + d1 * d1) <= opt.parameter_tolerance * (sqrt(x[0] * x[0] + x[1] * x[1]) + opt.parameter_tolerance))
      break;
    if (fabs(cost - cost_c) <= opt.function_tolerance * cost) break;
    const double rel = (cost - cost_c) / model_change;
    if (rel > LM_MIN_REL_DECREASE) {
      x[0] = xc[0]; x[1] = xc[1];
      const double t = 2.0 * rel - 1.0;
      radius = fmin(LM_MAX_RADIUS, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
      decrease = 2.0;
      need_eval = 1;
    } else {
      radius /= decrease; decrease *= 2.0;
    }
  }
  return iters;
}

/* ceres::Covariance of the 2-vector: (J'J)^-1 at x; 0 if J'J is rank deficient
 * (reciprocal condition number below Ceres' min_reciprocal_condition_number 1e-14) */
static int normals_covariance(const double *Q, int np, const double *x, double *cov) {
  double g[2], H[3];
  normals_cost(Q, np, x, g, H);
  const double tr = H[0] + H[2], det = H[0] * H[2] - H[1] * H[1];
  const double disc = sqrt(fmax(0.0, 0.25 * tr * tr - det));
  const double lmax = 0.5 * tr + disc, lmin = det / (lmax > 0 ? lmax : 1.0);
  if (!(lmax > 0.0) || !isfinite(lmax) || !(lmin / lmax >= 1e-14)) return 0;
  cov[0] = H[2] / det; cov[1] = -H[1] / det; cov[2] = -H[1] / det; cov[3] = H[0] / det;
  return 1;
}

int oracle_normals_batched(const defslam_normals_problem *p, double *k_out, double *cov_out, float *normal_out,
                           uint8_t *status_out, int32_t *iters_out, float *pair_normal_out,
                           uint8_t *pair_valid_out) {
  for (int i = 0; i < p->n_points; i++) {
    const int j0 = p->pair_ptr[i], j1 = p->pair_ptr[i + 1];
    double *Q = (double *)malloc(sizeof(double) * 20 * (j1 - j0 + 1));
    int np = 0;
    for (int j = j0; j < j1; j++) {
      if (pair_valid_out) pair_valid_out[j] = 0;
      if (p->pair_from_ref && !p->pair_from_ref[j]) continue;
      pair_coefficients(p->J12 + 4 * j, p->H12 + 6 * j, p->I1 + 2 * j, p->I2 + 2 * j, p->corrected_t2, Q + 20 * np,
                        Q + 20 * np + 10);
      np++;
    }
    double x[2] = {p->k_init ? p->k_init[2 * i] : 0.0, p->k_init ? p->k_init[2 * i + 1] : 0.0};
    int status = 0, iters = 0;
    if (np > 0) {
      iters = normals_solve_point(Q, np, x, p->max_iterations);
      double cov[4];
      if (normals_covariance(Q, np, x, cov)) {
        status = 1;
        if (cov_out) memcpy(cov_out + 4 * i, cov, sizeof(cov));
        if (normal_out) {
          const float u = p->ref_uv[2 * i], v = p->ref_uv[2 * i + 1];
          normal_out[3 * i] = (float)x[0];
          normal_out[3 * i + 1] = (float)x[1];
          normal_out[3 * i + 2] = (float)(1 - x[0] * u - x[1] * v);
        }
      } else status = 2;
    }
    free(Q);
    if (k_out) { k_out[2 * i] = x[0]; k_out[2 * i + 1] = x[1]; }
    if (status_out) status_out[i] = (uint8_t)status;
    if (iters_out) iters_out[i] = iters;
    if (status == 2) continue;
    /* transfer to the second keyframe of every pair: NormalEstimator.cc:176-223 */
    for (int j = j0; j < j1; j++) {
      double n0, n1;
      const int from_ref = p->pair_from_ref ? p->pair_from_ref[j] : 1;
      if (from_ref) {
        if (status != 1) continue;
        n0 = x[0]; n1 = x[1];
      } else {
        if (!p->k_first) continue;
        const float f0 = p->k_first[2 * j], f1 = p->k_first[2 * j + 1];
        if (f0 != f0 || f1 != f1) continue;
        n0 = f0; n1 = f1;
      }
      const float *Jf = p->J12 + 4 * j, *Ji = p->J21 + 4 * j, *Hh = p->H12 + 6 * j;
      const float a = Jf[0], b = Jf[1], c = Jf[2], d = Jf[3];
      const float j21_11 = Ji[0], j21_21 = Ji[1], j21_12 = Ji[2], j21_22 = Ji[3];
      volatile float ad = a * d, cb = c * b;
      const float det = ad - cb;
      volatile float m1 = -b * Hh[4] / 2, m2 = a * Hh[5] / 2, m3 = (d * Hh[0]) / 2, m4 = (c * Hh[1]) / 2;
      const float t1 = m1 + m2, t2 = m3 - m4;
      volatile float dt2 = d * t2, bt1 = b * t1, at1 = a * t1, ct2 = c * t2, dd = det * det;
      const float corr1 = (dt2 - bt1) / dd, corr2 = (at1 - ct2) / dd;
      /* float * double products, summed in double (cv::Vec2d norm) */
      const double k1 = (double)j21_11 * n0 + (double)j21_12 * n1 + (double)corr1;
      const double k2 = (double)j21_21 * n0 + (double)j21_22 * n1 + (double)corr2;
      if (pair_normal_out) {
        pair_normal_out[3 * j] = (float)k1;
        pair_normal_out[3 * j + 1] = (float)k2;
        pair_normal_out[3 * j + 2] = (float)(1 - k1 * p->I2[2 * j] - k2 * p->I2[2 * j + 1]);
      }
      if (pair_valid_out) pair_valid_out[j] = 1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ shape from normals */

/* stacked system [M; bending*B; 1'] , rhs [0; 0; NC*mean]   ShapeFromNormals.cc:38-98,178-260 */
int oracle_sfn_system(const defslam_sfn_problem *p, double *A, double *b) {
  const defslam_bbs *s = &p->bbs;
  const int NC = s->nptsu * s->nptsv, n = p->n_normals, rows = 2 * n + NC + 1;
  if (s->valdim != 1) return DEFSLAM_EBADARG;
  double *u = (double *)malloc(sizeof(double) * (n + 1)), *v = (double *)malloc(sizeof(double) * (n + 1));
  for (int i = 0; i < n; i++) { u[i] = p->uv[2 * i]; v[i] = p->uv[2 * i + 1]; }
  const size_t sz = (size_t)(n + 1) * NC;
  double *C = (double *)malloc(sizeof(double) * sz), *Cu = (double *)malloc(sizeof(double) * sz);
  double *Cv = (double *)malloc(sizeof(double) * sz);
  int rc = oracle_bbs_coloc(s, n, u, v, 0, 0, C);
  rc |= oracle_bbs_coloc(s, n, u, v, 1, 0, Cu);
  rc |= oracle_bbs_coloc(s, n, u, v, 0, 1, Cv);
  memset(A, 0, sizeof(double) * (size_t)rows * NC);
  memset(b, 0, sizeof(double) * rows);
  for (int i = 0; i < n; i++) {
    double nx = p->normals[3 * i], ny = p->normals[3 * i + 1], nz = p->normals[3 * i + 2];
    const double nn = sqrt(nx * nx + ny * ny + nz * nz);
    nx /= nn; ny /= nn; nz /= nn;
    const double ne = nx * u[i] + ny * v[i] + nz; /* n . eta */
    for (int c = 0; c < NC; c++) {
      A[(size_t)i * NC + c] = ne * Cu[(size_t)i * NC + c] + nx * C[(size_t)i * NC + c];
      A[(size_t)(i + n) * NC + c] = ne * Cv[(size_t)i * NC + c] + ny * C[(size_t)i * NC + c];
    }
  }
  double *B = (double *)malloc(sizeof(double) * (size_t)NC * NC);
  oracle_bbs_bending(s, B);
  for (int a = 0; a < NC; a++)
    for (int c = 0; c < NC; c++) A[(size_t)(2 * n + a) * NC + c] = p->bending * B[(size_t)a * NC + c];
  for (int c = 0; c < NC; c++) A[(size_t)(2 * n + NC) * NC + c] = 1.0;
  b[2 * n + NC] = NC * p->mean_depth;
  free(u); free(v); free(C); free(Cu); free(Cv); free(B);
  return rc ? DEFSLAM_EBADARG : 0;
}

/* dense Householder QR least squares (Eigen's householderQr().solve): A m x n, m >= n */
static void householder_lstsq(double *A, int m, int n, double *b, double *x) {
  for (int k = 0; k < n; k++) {
    double nrm = 0.0;
    for (int i = k; i < m; i++) nrm += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;
    const double alpha = A[(size_t)k * n + k] > 0 ? -nrm : nrm;
    const double v0 = A[(size_t)k * n + k] - alpha;
    /* v = [v0, A[k+1.., k]]; H = I - 2 v v' / (v'v) */
    double vtv = v0 * v0;
    for (int i = k + 1; i < m; i++) vtv += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    if (vtv == 0.0) continue;
    for (int j = k + 1; j < n; j++) {
      double s = v0 * A[(size_t)k * n + j];
      for (int i = k + 1; i < m; i++) s += A[(size_t)i * n + k] * A[(size_t)i * n + j];
      s = 2.0 * s / vtv;
      A[(size_t)k * n + j] -= s * v0;
      for (int i = k + 1; i < m; i++) A[(size_t)i * n + j] -= s * A[(size_t)i * n + k];
    }
    double s = v0 * b[k];
    for (int i = k + 1; i < m; i++) s += A[(size_t)i * n + k] * b[i];
    s = 2.0 * s / vtv;
    b[k] -= s * v0;
    for (int i = k + 1; i < m; i++) b[i] -= s * A[(size_t)i * n + k];
    A[(size_t)k * n + k] = alpha;
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int j = i + 1; j < n; j++) s -= A[(size_t)i * n + j] * x[j];
    x[i] = s / A[(size_t)i * n + i];
  }
}

static int cmp_float(const void *a, const void *b) {
  const float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

int oracle_sfn_solve(const defslam_sfn_problem *p) {
  const defslam_bbs *s = &p->bbs;
  const int NC = s->nptsu * s->nptsv, n = p->n_normals, rows = 2 * n + NC + 1;
  double *A = (double *)malloc(sizeof(double) * (size_t)rows * NC), *b = (double *)malloc(sizeof(double) * rows);
  double *x = (double *)malloc(sizeof(double) * NC);
  int rc = oracle_sfn_system(p, A, b);
  if (!rc) {
    householder_lstsq(A, rows, NC, b, x);
    for (int i = 0; i < NC; i++)
      if (!isfinite(x[i])) rc = DEFSLAM_ENUMERIC;
  }
  if (!rc) {
    /* median rescale through fp32 (ShapeFromNormals.cc:128-142) */
    float *dv = (float *)malloc(sizeof(float) * NC);
    for (int i = 0; i < NC; i++) dv[i] = (float)x[i];
    qsort(dv, NC, sizeof(float), cmp_float);
    const float corr = 1 / dv[NC / 2];
    free(dv);
    for (int i = 0; i < NC; i++) x[i] = corr * x[i];
    if (p->ctrl_out) memcpy(p->ctrl_out, x, sizeof(double) * NC);
    if (p->n_eval > 0 && p->xyz_out) {
      double *u = (double *)malloc(sizeof(double) * p->n_eval), *v = (double *)malloc(sizeof(double) * p->n_eval);
      double *val = (double *)malloc(sizeof(double) * p->n_eval);
      for (int i = 0; i < p->n_eval; i++) { u[i] = p->eval_uv[2 * i]; v[i] = p->eval_uv[2 * i + 1]; }
      oracle_bbs_eval(s, x, p->n_eval, u, v, 0, 0, val);
      for (int i = 0; i < p->n_eval; i++) {
        p->xyz_out[3 * i] = (float)(u[i] * val[i]);
        p->xyz_out[3 * i + 1] = (float)(v[i] * val[i]);
        p->xyz_out[3 * i + 2] = (float)val[i];
      }
      free(u); free(v); free(val);
    }
  }
  free(A); free(b); free(x);
  return rc;
}
