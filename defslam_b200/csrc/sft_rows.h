/*
 * sft_rows.h -- banded Cholesky + solve of (H + lambda I) dx = b, "row owner" form.
 *
 * Included by sft_core.h (needs its Ctx, chol4, dmma884 ...).  Replaces, like factor_solve(),
 * BlockSolver::solve + LinearSolverDense::solve (block_solver.hpp:354-380, linear_solver_dense.h:65-113)
 * for the arrowhead system of the SfT graph.
 *
 * Left-looking by block rows of 8, one WARP per block row, the row's band (NT = bwp/8 + 1 tiles of 8x8)
 * held in FP64 tensor-core accumulator registers from the moment H is read until the row is final:
 *
 *   for k = I-NBK .. I-1 :   X   = C(I,k) inv(L_kk)^T              (2 DMMA; inv(L_kk) from the chain warp)
 *                            C(I,J) -= X L(J,k)^T,  J = k+1 .. I    (2 DMMA per tile; L(J,k) of the rows above
 *                                                                    from a shared-memory ring, X from registers)
 *   C(I,I) -> chain warp -> L_II, inv(L_II)
 *
 * Compared with the right-looking sliding window (factor_solve): no read-modify-write of the trailing
 * matrix in shared memory (a quarter of the shared-memory traffic: only the B operands are loaded), no
 * CTA-wide barrier per step (rows synchronise through progress counters in shared memory: row I needs
 * tile k of row I-1, which implies every row above it), and the diagonal factorisation -- the latency
 * chain of the whole method -- runs on a warp of its own whose sub-partition does no tensor-core work.
 *
 * Warp roles (256 threads): warp 0 = the dependency chain (last panel tile of each row, last update of its
 * diagonal block, Cholesky of that block, inverse of the factor -- nothing on it crosses a warp boundary);
 * warp 4 (same sub-partition as warp 0) idles so that the chain owns its FP64 pipe; warp 1 = the 8 border
 * rows (camera border x6, right-hand side, pad), i.e. the forward substitution and the Schur complement of
 * the camera block; warps 2,3,5,6,7 = row owners (rows I = o, o+5, ...).
 *
 * Tiles live in the ring in "fragment order" (chunk c, row g, column 4c+q at c*32 + g*4 + q): the layout
 * one m8n8k4 A or B operand load wants, 32 consecutive doubles per chunk, conflict free.  The finished
 * factor goes to global memory row-major per tile (stride LT_STRIDE) for the backward sweep.
 */
#ifndef DS_SFT_ROWS_H_
#define DS_SFT_ROWS_H_
#include "ds_rowchol.h"

namespace ds {

constexpr int ROW_OWNERS = ROWS_OWNERS_;
static_assert(LT_STRIDE == ROWS_LT_STRIDE_, "tile stride of the factor workspace");

static inline
#if DS_CUDA
__host__ __device__
#endif
bool row_mode_nt_supported(int nt) { return nt >= 5 && nt <= 14; } /* 4x4 .. 17x17 regular meshes: one instantiation each */

/* global workspace of the tile-form factor: nblk rows of nt tiles */
static inline
#if DS_CUDA
__host__ __device__
#endif
size_t row_mode_factor_doubles(int nblk, int nt) { return (size_t)nblk * nt * LT_STRIDE; }

/* factor the 8x8 diagonal block held as [A11 0; A21 A22] (packed lower triangles, lambda already on the
 * diagonal): on return the factor with RECIPROCAL diagonal entries; returns true if a pivot is not > 0 */
DS_FN bool diag_factor_regs(double A11[10], double L21[16], double A22[10]) {
  bool bad = chol4(A11);
#pragma unroll
  for (int a = 0; a < 4; a++) {
    double *x = &L21[a * 4];
    x[0] *= A11[0];
    x[1] = (x[1] - x[0] * A11[1]) * A11[2];
    x[2] = (x[2] - x[0] * A11[3] - x[1] * A11[4]) * A11[5];
    x[3] = (x[3] - x[0] * A11[6] - x[1] * A11[7] - x[2] * A11[8]) * A11[9];
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b <= a; b++) {
      double v = A22[a * (a + 1) / 2 + b];
#pragma unroll
      for (int m = 0; m < 4; m++) v -= L21[a * 4 + m] * L21[b * 4 + m];
      A22[a * (a + 1) / 2 + b] = v;
    }
  bad = chol4(A22) || bad;
  return bad;
}

/* column j of Y = inv(L): col[m] = Y[m][j] (zero above the diagonal) */
DS_FN void inv_column_regs(const double *A11, const double *L21, const double *A22, int j, double col[NB]) {
  double sacc[NB];
#pragma unroll
  for (int i = 0; i < NB; i++) sacc[i] = (i == j) ? 1.0 : 0.0;
#pragma unroll
  for (int m = 0; m < NB; m++) {
    const double rd = m < 4 ? A11[m * (m + 1) / 2 + m] : A22[(m - 4) * (m - 3) / 2 + (m - 4)];
    const double xm = sacc[m] * rd;
    col[m] = m >= j ? xm : 0.0;
#pragma unroll
    for (int i = m + 1; i < NB; i++) sacc[i] -= lreg(A11, L21, A22, i, m) * xm;
  }
}


/* Solve (H + lambda I) dx = b in row-owner form.  Same contract as factor_solve(). */
template <int NT>
DS_FN_NOINLINE bool factor_rows(const Team team, double lambda) {
  Ctx &c = ctx_ref();
  constexpr int NBK = NT - 1;
  const int bw = c.pl.bw, bwE = c.pl.bwE, ld = c.pl.ld, Dp = c.pl.Dn_pad, nblk = c.pl.nblk, ES = c.pl.ES;
  const int R = NBK + ROW_OWNERS;
  double *const sm = sm_base();
  double *G = sm + c.sl.G, *Hcc = sm + c.sl.Hcc, *dx = sm + c.sl.dx;
  double *W = sm + c.sl.W;
  double *Eg = c.ws.Eg, *Lt = c.ws.Lb, *Dinv = c.ws.Dinv;
  const double *Hb = c.ws.Hb, *Cg = c.ws.Cg;
  int *flag = (int *)(sm + c.sl.red + 36);
  int *sy = (int *)(sm + c.sl.sy);
  (void)bw; (void)bwE; (void)ld; (void)W;

#if DS_CUDA
  rows_factor<NT>(team.tid, team.nthr, nblk, ROW_OWNERS, SftBandLoader{Hb, ld, bwE, bw, lambda}, W, sm + c.sl.er,
                  sm + c.sl.db, sy, flag, Cg, Eg, ES, G, Hcc, lambda, Lt, Dinv);
#else
  /* emulation: the same block algorithm, one row after the other; finished tiles are read back from the
   * global tile-form factor instead of a ring */
  (void)sy;
  *flag = 0;
  {
    const int lo = bwE - bw;
    double *Et = W; /* border tiles [nblk][64], row-major 8x8 (rows = border rows) */
    double Gs[64];
    for (int i = 0; i < 64; i++) {
      const int a = i >> 3, b = i & 7;
      double v = 0.0;
      if (a < 6 && b < 6) v = Hcc[a * 6 + b] + (a == b ? lambda : 0.0);
      else if (a == 6 && b < 6) v = Hcc[36 + b];
      Gs[i] = v;
    }
    for (int I = 0; I < nblk; I++) {
      double C[NT][64];
      for (int t = 0; t < NT; t++) {
        const int J = I - NBK + t;
        for (int g = 0; g < NB; g++)
          for (int cc = 0; cc < NB; cc++) {
            const int i = NB * I + g, off = NB * J + cc - i + bwE;
            double v = (J >= 0 && off >= lo && off <= bwE) ? Hb[(size_t)i * ld + off] : 0.0;
            if (t == NBK && g == cc) v += lambda;
            C[t][g * 8 + cc] = v;
          }
      }
      double *ltrow = Lt + (size_t)I * NT * LT_STRIDE;
      for (int t = 0; t < NBK; t++) {
        const int k = I - NBK + t;
        if (k < 0) continue;
        const double *Y = Dinv + k * 64; /* row-major inv(L_kk) */
        double X[64];
        for (int g = 0; g < NB; g++)
          for (int cc = 0; cc < NB; cc++) {
            double s = 0.0;
            for (int m = 0; m < NB; m++) s += C[t][g * 8 + m] * Y[cc * 8 + m];
            X[g * 8 + cc] = s;
          }
        for (int i = 0; i < 64; i++) ltrow[t * LT_STRIDE + i] = X[i];
        for (int u = t + 1; u < NT; u++) {
          const int J = I - NBK + u;
          const double *B = u == NBK ? X : Lt + ((size_t)J * NT + (t - u + NBK)) * LT_STRIDE;
          for (int g = 0; g < NB; g++)
            for (int cc = 0; cc < NB; cc++) {
              double s = 0.0;
              for (int m = 0; m < NB; m++) s += X[g * 8 + m] * B[cc * 8 + m];
              C[u][g * 8 + cc] -= s;
            }
        }
      }
      double A11[10], L21[16], A22[10];
      const double *D = C[NBK];
      for (int a = 0; a < 4; a++)
        for (int b = 0; b <= a; b++) { A11[a * (a + 1) / 2 + b] = D[a * 8 + b]; A22[a * (a + 1) / 2 + b] = D[(4 + a) * 8 + 4 + b]; }
      for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) L21[a * 4 + b] = D[(4 + a) * 8 + b];
      if (diag_factor_regs(A11, L21, A22)) *flag = 1;
      for (int j = 0; j < NB; j++) {
        double col[NB];
        inv_column_regs(A11, L21, A22, j, col);
        for (int m = 0; m < NB; m++) Dinv[I * 64 + m * 8 + j] = col[m];
      }
      /* border tile I */
      double E[64];
      for (int g = 0; g < NB; g++)
        for (int cc = 0; cc < NB; cc++) E[g * 8 + cc] = Cg[(size_t)g * ES + NB * I + cc];
      for (int t = 0; t < NBK; t++) {
        const int k = I - NBK + t;
        if (k < 0) continue;
        const double *Ek = Et + k * 64, *B = ltrow + t * LT_STRIDE;
        for (int g = 0; g < NB; g++)
          for (int cc = 0; cc < NB; cc++) {
            double s = 0.0;
            for (int m = 0; m < NB; m++) s += Ek[g * 8 + m] * B[cc * 8 + m];
            E[g * 8 + cc] -= s;
          }
      }
      const double *Y = Dinv + I * 64;
      for (int g = 0; g < NB; g++)
        for (int cc = 0; cc < NB; cc++) {
          double s = 0.0;
          for (int m = 0; m < NB; m++) s += E[g * 8 + m] * Y[cc * 8 + m];
          Et[I * 64 + g * 8 + cc] = s;
          Eg[(size_t)g * ES + NB * I + cc] = s;
        }
      for (int g = 0; g < NB; g++)
        for (int cc = 0; cc < NB; cc++) {
          double s = 0.0;
          for (int m = 0; m < NB; m++) s += Et[I * 64 + g * 8 + m] * Et[I * 64 + cc * 8 + m];
          Gs[g * 8 + cc] -= s;
        }
    }
    for (int i = 0; i < 64; i++) G[i] = Gs[i];
  }
#endif
  prof_mark(team, c, PF_S3);

  /* Schur complement system of the camera: S dc = rhs (6x6 Cholesky in the corner block; row 6 = rhs) */
  if (team.tid == 0) {
    bool ok = true;
    for (int j = 0; j < 6; j++) {
      double d = G[j * 8 + j];
      for (int m = 0; m < j; m++) d -= G[j * 8 + m] * G[j * 8 + m];
      if (!(d > 0.0)) ok = false;
      const double l = sqrt(d);
      G[j * 8 + j] = l;
      for (int i = j + 1; i < 6; i++) {
        double s = G[i * 8 + j];
        for (int m = 0; m < j; m++) s -= G[i * 8 + m] * G[j * 8 + m];
        G[i * 8 + j] = s / l;
      }
    }
    if (!ok) *flag = 1;
    if (*flag == 0) {
      double *y = G + 56;
      for (int i = 0; i < 6; i++) {
        double s = G[48 + i];
        for (int m = 0; m < i; m++) s -= G[i * 8 + m] * y[m];
        y[i] = s / G[i * 8 + i];
      }
      for (int i = 5; i >= 0; i--) {
        double s = y[i];
        for (int m = i + 1; m < 6; m++) s -= G[m * 8 + i] * y[m];
        y[i] = s / G[i * 8 + i];
      }
      for (int i = 0; i < 6; i++) dx[Dp + i] = y[i];
    }
  }
  team.sync();
  prof_mark(team, c, PF_SCHUR);
  if (*flag != 0) { team.sync(); return false; }

  /* v = z - Y^T dc (the border rows are final in global memory: plain loads, written by this CTA before a barrier) */
  DS_FOR(i, Dp) {
    double s = Eg[6 * (size_t)ES + i];
#pragma unroll
    for (int e = 0; e < 6; e++) s -= Eg[e * (size_t)ES + i] * dx[Dp + e];
    dx[i] = s;
  }
#if DS_CUDA
  DS_FOR(i, nblk + 1) sy[i] = 0; /* ready flag of every block row + the sweep's progress */
#endif
  team.sync();

  /* backward sweep L^T dn = v by block rows, bottom up: d = inv(L_kk)^T y, then dx[J] -= L(kb,J)^T d.
   * Row blocks (NT tiles + the inverse of the diagonal block) stream through a ring of NBUF buffers. */
  constexpr int NBUF = ROWS_BWD_BUFS_;
  constexpr int BUFD = NT * LT_STRIDE + 64;
  double *sol = W + NBUF * BUFD;
  prof_mark(team, c, PF_BWD_INIT);
#if DS_CUDA
  rows_backward<NT, NBUF>(team.tid, team.nthr, nblk, W, sy, Lt, Dinv, dx, sol);
  team.sync();
#else
  for (int kb = nblk - 1; kb >= 0; kb--) {
    const int k = kb * NB;
    const int t0 = kb < NBK ? NBK - kb : 0; /* first tile of the row that exists */
    const int nupd = NB * (NBK - t0);
    const double *LR = Lt + (size_t)kb * NT * LT_STRIDE, *Y = Dinv + kb * 64;
    double d[NB], y[NB];
    for (int a = 0; a < NB; a++) y[a] = dx[k + a];
    for (int a = 0; a < NB; a++) {
      /* the device's summation order: two interleaved partial sums over m, zero terms above the diagonal */
      double s0 = 0.0, s1 = 0.0;
      for (int m = 0; m < NB; m += 2) {
        s0 = fma(m >= a ? Y[m * 8 + a] : 0.0, y[m], s0);
        s1 = fma(m + 1 >= a ? Y[(m + 1) * 8 + a] : 0.0, y[m + 1], s1);
      }
      d[a] = s0 + s1;
    }
    for (int a = 0; a < NB; a++) sol[k + a] = d[a];
    for (int jj = 0; jj < nupd; jj++) {
      const int t = t0 + (jj >> 3), cc = jj & 7;
      const int jc = NB * (kb - NBK + t) + cc;
      const double *Lc = LR + t * LT_STRIDE + cc;
      double s = dx[jc];
      for (int a = 0; a < NB; a++) s -= Lc[a * 8] * d[a];
      dx[jc] = s;
    }
  }
#endif
  DS_FOR(i, Dp) dx[i] = sol[i];
  team.sync();
  prof_mark(team, c, PF_BWD);
  return true;
}

}  // namespace ds
#endif
