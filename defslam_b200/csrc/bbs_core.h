/*
 * bbs_core.h -- uniform bicubic B-spline arithmetic (device functions).
 *
 * Replaces Thirdparty/BBS for the NRSfM path:
 *   normalize_with_inter   bbs.cc:70-92     -> bbs_normalize
 *   eval_basis{,_d,_dd}    bbs.cc:95-121    -> bbs_basis
 *   get_deriv_fact         bbs.cc:140-145   -> bbs_deriv_fact
 *   eval / EvalEigen       bbs.cc:155-195, bbs_coloc.cc:610-653 -> bbs_eval_site
 *   coloc / coloc_deriv    bbs.cc:214-340, bbs_coloc.cc:76-207  -> bbs_coloc_row
 *   bending_ur/BendingEigen bbs.cc:563-640, bbs_coloc.cc:406-507 -> bbs_bending_entry
 *
 * The reference's three 256-entry bending tables (bbs.cc:360-554) are not
 * reproduced: they are the cell integrals of products of the cubic basis
 * functions' derivatives, which are computed here exactly (4-point
 * Gauss-Legendre is exact for degree <= 7; the integrands have degree <= 6).
 */
#ifndef DS_BBS_CORE_H_
#define DS_BBS_CORE_H_
#include "ds_common.h"

namespace ds {

struct BbsView {
  double umin, umax, vmin, vmax;
  int nptsu, nptsv, valdim;
};

DS_FN void bbs_normalize(double xmin, double xmax, int npts, double x, double &nx, int &inter) {
  const int ninter = npts - 3;
  const double width = (xmax - xmin) / ninter;
  if (x == xmax) { nx = 1.0; inter = ninter - 1; }
  else if (x < xmin) { nx = (x - xmin) / width; inter = -1; }
  else if (x > xmax) { nx = (x - xmin) / width - ninter; inter = ninter; }
  else {
    const double scaled = (x - xmin) / width;
    inter = (int)floor(scaled);
    nx = scaled - inter;
  }
}

DS_FN void bbs_basis(int order, double nx, double b[4]) {
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

DS_FN double bbs_deriv_fact(const BbsView &s, int du, int dv) {
  const double su = (s.umax - s.umin) / (s.nptsu - 3);
  const double sv = (s.vmax - s.vmin) / (s.nptsv - 3);
  /* pow(s, 0|1|2) of the reference is exact-rounded; so is repeated multiplication */
  const double pu = du == 0 ? 1.0 : (du == 1 ? su : su * su);
  const double pv = dv == 0 ? 1.0 : (dv == 1 ? sv : sv * sv);
  return 1.0 / (pu * pv);
}

DS_FN bool bbs_in_domain(const BbsView &s, int Iu, int Iv) {
  return !(Iu < 0 || Iu > s.nptsu - 4 || Iv < 0 || Iv > s.nptsv - 4);
}

/* value (valdim numbers) of the (du,dv) derivative at one site; false and NaNs
 * when the site lies outside the spline domain (the reference indexes out of
 * range there) */
DS_FN bool bbs_eval_site(const BbsView &s, const double *ctrl, double u, double v, int du, int dv, double *val) {
  double nu, nv, bu[4], bv[4];
  int Iu, Iv;
  bbs_normalize(s.umin, s.umax, s.nptsu, u, nu, Iu);
  bbs_normalize(s.vmin, s.vmax, s.nptsv, v, nv, Iv);
  if (!bbs_in_domain(s, Iu, Iv)) {
    for (int d = 0; d < s.valdim; d++) val[d] = NAN;
    return false;
  }
  bbs_basis(du, nu, bu);
  bbs_basis(dv, nv, bv);
  const double fact = bbs_deriv_fact(s, du, dv);
  for (int d = 0; d < s.valdim; d++) val[d] = 0.0;
  for (int iu = 0; iu < 4; iu++)
    for (int iv = 0; iv < 4; iv++) {
      const double bas = bu[iu] * bv[iv];
      int ind = s.valdim * ((iu + Iu) * s.nptsv + iv + Iv);
      for (int d = 0; d < s.valdim; d++) val[d] += ctrl[ind++] * bas;
    }
  for (int d = 0; d < s.valdim; d++) val[d] *= fact;
  return true;
}

/* one dense collocation row (NC entries, 16 non-zero); false outside the domain */
DS_FN bool bbs_coloc_row(const BbsView &s, double u, double v, int du, int dv, double *row) {
  double nu, nv, bu[4], bv[4];
  int Iu, Iv;
  bbs_normalize(s.umin, s.umax, s.nptsu, u, nu, Iu);
  bbs_normalize(s.vmin, s.vmax, s.nptsv, v, nv, Iv);
  if (!bbs_in_domain(s, Iu, Iv)) return false;
  bbs_basis(du, nu, bu);
  bbs_basis(dv, nv, bv);
  const bool deriv = (du | dv) != 0;
  const double fact = deriv ? bbs_deriv_fact(s, du, dv) : 1.0;
  for (int iu = 0; iu < 4; iu++)
    for (int iv = 0; iv < 4; iv++) {
      const int col = (iu + Iu) * s.nptsv + iv + Iv;
      row[col] = deriv ? fact * bu[iu] * bv[iv] : bu[iu] * bv[iv];
    }
  return true;
}

/* integral over [0,1] of N_a^(order)(t) N_b^(order)(t), 4-point Gauss-Legendre */
DS_FN double bbs_cell_integral(int order, int a, int b) {
  const double gx[4] = {0.06943184420297371, 0.33000947820757187, 0.6699905217924281, 0.9305681557970262};
  const double gw[4] = {0.17392742256872692, 0.32607257743127305, 0.32607257743127305, 0.17392742256872692};
  double s = 0.0;
  for (int k = 0; k < 4; k++) {
    double B[4];
    bbs_basis(order, gx[k], B);
    s += gw[k] * B[a] * B[b];
  }
  return s;
}

/* entry (I, J) of the dense bending matrix (lambda = 1), I = iu*nptsv + iv */
DS_FN double bbs_bending_entry(const BbsView &s, int I, int J) {
  const int nx = s.nptsv, ny = s.nptsu;
  const int iu = I / nx, iv = I % nx, ju = J / nx, jv = J % nx;
  const double sy = (s.umax - s.umin) / (s.nptsu - 3);
  const double sx = (s.vmax - s.vmin) / (s.nptsv - 3);
  const int du = iu > ju ? iu - ju : ju - iu, dv = iv > jv ? iv - jv : jv - iv;
  if (du > 3 || dv > 3) return 0.0;
  const double cxx = sy / pow(sx, 3), cxy = 1.0 / (sx * sy), cyy = sx / pow(sy, 3);
  double acc = 0.0;
  const int b0 = (iu > ju ? iu : ju) - 3, b1 = iu < ju ? iu : ju;
  const int a0 = (iv > jv ? iv : jv) - 3, a1 = iv < jv ? iv : jv;
  for (int b = (b0 > 0 ? b0 : 0); b <= b1 && b <= ny - 4; b++)
    for (int a = (a0 > 0 ? a0 : 0); a <= a1 && a <= nx - 4; a++) {
      const int e1 = iu - b, e2 = ju - b, f1 = iv - a, f2 = jv - a;
      /* xx: d2/dv2 ; yy: d2/du2 ; xy: 2 * mixed */
      const double bxx = bbs_cell_integral(0, e1, e2) * bbs_cell_integral(2, f1, f2);
      const double byy = bbs_cell_integral(2, e1, e2) * bbs_cell_integral(0, f1, f2);
      const double bxy = 2.0 * bbs_cell_integral(1, e1, e2) * bbs_cell_integral(1, f1, f2);
      acc += cxx * bxx + cxy * bxy + cyy * byy;
    }
  return acc;
}

}  // namespace ds
#endif
