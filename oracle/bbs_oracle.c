/*
 * bbs_oracle.c -- CPU restatement of the bicubic B-spline routines of the NRSfM path.
 * TEST INFRASTRUCTURE ONLY (see sft_oracle.c header).
 *
 * Pinned: this file is checked in tests/ against the reference's OWN Thirdparty/BBS/bbs.cc,
 * compiled where it lies into oracle/_ref/libbbs_ref.so (oracle/Makefile, target `ref`),
 * and against golden vectors generated from that library (tests/golden/bbs_ref.npz).
 *
 * Follows (paths under the DefSLAM tree):
 *   normalize_with_inter     Thirdparty/BBS/bbs.cc:70-92
 *   eval_basis{,_d,_dd}      Thirdparty/BBS/bbs.cc:95-121
 *   get_deriv_fact           Thirdparty/BBS/bbs.cc:140-145
 *   eval / EvalEigen         Thirdparty/BBS/bbs.cc:155-195, bbs_coloc.cc:610-653
 *   coloc / coloc_deriv      Thirdparty/BBS/bbs.cc:214-340 (dense rows here, as colocEigen
 *                            bbs_coloc.cc:76-207 expands them)
 *   bending_ur / BendingEigen Thirdparty/BBS/bbs.cc:563-640, bbs_coloc.cc:406-507
 * The three 256-entry coefficient tables of bbs.cc:360-554 are NOT copied: they are the
 * integrals over one knot cell of products of basis-function derivatives, evaluated here by
 * exact polynomial integration (the CUDA path uses Gauss-Legendre quadrature instead).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "sft_oracle.h"

static void norm_inter(double xmin, double xmax, int npts, double x, double *nx, int *inter) {
  const int ninter = npts - 3;
  const double width_inter = (xmax - xmin) / ninter;
  if (x == xmax) { *nx = 1.0; *inter = ninter - 1; }
  else if (x < xmin) { *nx = (x - xmin) / width_inter; *inter = -1; }
  else if (x > xmax) { *nx = (x - xmin) / width_inter - ninter; *inter = ninter; }
  else {
    const double scaled = (x - xmin) / width_inter;
    *inter = (int)floor(scaled);
    *nx = scaled - *inter;
  }
}

static void basis(int order, double nx, double *b) {
  if (order == 0) {
    const double nx2 = nx * nx, nx3 = nx2 * nx;
    b[0] = (-nx3 + 3.0 * nx2 - 3.0 * nx + 1.0) / 6.0;
    b[1] = (3.0 * nx3 - 6.0 * nx2 + 4.0) / 6.0;
    b[2] = (-3.0 * nx3 + 3.0 * nx2 + 3.0 * nx + 1.0) / 6.0;
    b[3] = nx3 / 6.0;
  } else if (order == 1) {
    const double nx2 = nx * nx;
    b[0] = (-nx2 + 2 * nx - 1) / 2.0;
    b[1] = (3.0 * nx2 - 4.0 * nx) / 2.0;
    b[2] = (-3 * nx2 + 2 * nx + 1) / 2.0;
    b[3] = nx2 / 2.0;
  } else {
    b[0] = -nx + 1.0;
    b[1] = 3.0 * nx - 2.0;
    b[2] = -3.0 * nx + 1.0;
    b[3] = nx;
  }
}

static double deriv_fact(const defslam_bbs *s, int du, int dv) {
  const double su = (s->umax - s->umin) / (s->nptsu - 3);
  const double sv = (s->vmax - s->vmin) / (s->nptsv - 3);
  return 1.0 / (pow(su, du) * pow(sv, dv));
}

int oracle_bbs_eval(const defslam_bbs *s, const double *ctrl, int32_t nsites, const double *u, const double *v,
                    int32_t du, int32_t dv, double *val) {
  for (int k = 0; k < nsites; k++) {
    double nu, nv, bu[4], bv[4];
    int Iu, Iv;
    norm_inter(s->umin, s->umax, s->nptsu, u[k], &nu, &Iu);
    norm_inter(s->vmin, s->vmax, s->nptsv, v[k], &nv, &Iv);
    if (Iu < 0 || Iu > s->nptsu - 4 || Iv < 0 || Iv > s->nptsv - 4) {
      for (int d = 0; d < s->valdim; d++) val[s->valdim * k + d] = NAN; /* reference: out-of-range read */
      continue;
    }
    basis(du, nu, bu);
    basis(dv, nv, bv);
    const double fact = deriv_fact(s, du, dv);
    for (int d = 0; d < s->valdim; d++) val[s->valdim * k + d] = 0.0;
    for (int iu = 0; iu < 4; iu++)
      for (int iv = 0; iv < 4; iv++) {
        const double bas = bu[iu] * bv[iv];
        int ind = s->valdim * ((iu + Iu) * s->nptsv + iv + Iv);
        for (int d = 0; d < s->valdim; d++) val[s->valdim * k + d] += ctrl[ind++] * bas;
      }
    for (int d = 0; d < s->valdim; d++) val[s->valdim * k + d] *= fact;
  }
  return 0;
}

int oracle_bbs_coloc(const defslam_bbs *s, int32_t nsites, const double *u, const double *v, int32_t du, int32_t dv,
                     double *C) {
  const int NC = s->nptsu * s->nptsv;
  memset(C, 0, sizeof(double) * (size_t)nsites * NC);
  for (int k = 0; k < nsites; k++) {
    double nu, nv;
    int Iu, Iv;
    norm_inter(s->umin, s->umax, s->nptsu, u[k], &nu, &Iu);
    norm_inter(s->vmin, s->vmax, s->nptsv, v[k], &nv, &Iv);
    if (Iu < 0 || Iu > s->nptsu - 4 || Iv < 0 || Iv > s->nptsv - 4) {
      memset(C, 0, sizeof(double) * (size_t)nsites * NC); /* coloc aborts (ret_code 1), matrix left empty */
      return DEFSLAM_EBADARG;
    }
  }
  const int deriv = (du | dv) != 0;
  const double fact = deriv ? deriv_fact(s, du, dv) : 1.0;
  for (int k = 0; k < nsites; k++) {
    double nu, nv, bu[4], bv[4];
    int Iu, Iv;
    norm_inter(s->umin, s->umax, s->nptsu, u[k], &nu, &Iu);
    norm_inter(s->vmin, s->vmax, s->nptsv, v[k], &nv, &Iv);
    basis(du, nu, bu);
    basis(dv, nv, bv);
    for (int iu = 0; iu < 4; iu++)
      for (int iv = 0; iv < 4; iv++) {
        const int col = (iu + Iu) * s->nptsv + iv + Iv;
        C[(size_t)k * NC + col] = deriv ? fact * bu[iu] * bv[iv] : bu[iu] * bv[iv];
      }
  }
  return 0;
}

/* monomial coefficients (constant term first) of the `order`-th derivative of basis a */
static void basis_poly(int order, int a, double c[4]) {
  static const double N[4][4] = {{1. / 6, -3. / 6, 3. / 6, -1. / 6},
                                 {4. / 6, 0., -6. / 6, 3. / 6},
                                 {1. / 6, 3. / 6, 3. / 6, -3. / 6},
                                 {0., 0., 0., 1. / 6}};
  double p[4];
  memcpy(p, N[a], sizeof(p));
  for (int o = 0; o < order; o++) {
    for (int i = 0; i < 3; i++) p[i] = (i + 1) * p[i + 1];
    p[3] = 0.0;
  }
  memcpy(c, p, sizeof(p));
}

/* integral over [0,1] of the product of the order-th derivatives of basis a and b */
static double cell_integral(int order, int a, int b) {
  double p[4], q[4], s = 0.0;
  basis_poly(order, a, p);
  basis_poly(order, b, q);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) s += p[i] * q[j] / (double)(i + j + 1);
  return s;
}

/* dense bending matrix, lambda = 1: accumulation over knot cells as bending_ur does */
int oracle_bbs_bending(const defslam_bbs *s, double *B) {
  const int ny = s->nptsu, nx = s->nptsv, NC = nx * ny;
  const double sy = (s->umax - s->umin) / (s->nptsu - 3);
  const double sx = (s->vmax - s->vmin) / (s->nptsv - 3);
  double coeff[16][16];
  for (int c = 0; c < 16; c++)
    for (int d = 0; d < 16; d++) {
      const int e1 = c / 4, f1 = c % 4, e2 = d / 4, f2 = d % 4;
      const double bxx = cell_integral(0, e1, e2) * cell_integral(2, f1, f2);
      const double byy = cell_integral(2, e1, e2) * cell_integral(0, f1, f2);
      const double bxy = 2.0 * cell_integral(1, e1, e2) * cell_integral(1, f1, f2);
      coeff[c][d] = sy * bxx / pow(sx, 3) + bxy / (sx * sy) + sx * byy / pow(sy, 3);
    }
  memset(B, 0, sizeof(double) * (size_t)NC * NC);
  for (int b = 0; b < ny - 3; b++)
    for (int a = 0; a < nx - 3; a++)
      for (int c = 0; c < 16; c++)
        for (int d = 0; d < 16; d++) {
          const int i = (b + c / 4) * nx + a + c % 4, j = (b + d / 4) * nx + a + d % 4;
          B[(size_t)i * NC + j] += coeff[c][d];
        }
  return 0;
}

/* Surface::getVertex  Modules/Mapping/Surface.cc:125-161 */
int oracle_surface_vertices(const defslam_bbs *s, const double *ctrl, int32_t xs, int32_t ys, float *out) {
  const double t = 0.03;
  int us = 0;
  for (int x = 0; x < xs; x++)
    for (int j = 0; j < ys; j++) {
      const double u = (double)((s->umax - s->umin - 2 * t) * x) / (xs - 1) + (s->umin + t);
      const double v = (double)((s->vmax - s->vmin - 2 * t) * j) / (ys - 1) + (s->vmin + t);
      double d;
      oracle_bbs_eval(s, ctrl, 1, &u, &v, 0, 0, &d);
      out[3 * us] = (float)(u * d);
      out[3 * us + 1] = (float)(v * d);
      out[3 * us + 2] = (float)d;
      us++;
    }
  return 0;
}
