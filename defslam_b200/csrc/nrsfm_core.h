/*
 * nrsfm_core.h -- device code of the NRSfM mapping stages.
 *
 *   schwarp_fit_one   one CTA fits the Schwarp between two keyframes: collocation
 *                     C'C by cell-sorted gather, Warp::initialize, Ceres-style LM
 *                     on the block-banded normal matrix, DiffProp records.
 *                     replaces Modules/Mapping/Schwarp.cc + SchwarpDatabase.cc:145-349
 *   normals_point     one thread solves the 2-unknown isometric system of a map
 *                     point and transfers the normal along its pairs.
 *                     replaces Modules/Mapping/NormalEstimator.cc:38-229, PolySolver.cc
 *   sfn_solve_one     one CTA solves the shape-from-normals least squares of a
 *                     keyframe (normal equations in packed shared memory +
 *                     corrected semi-normal refinement).
 *                     replaces Modules/Mapping/ShapeFromNormals.cc
 *
 * Not a translation of the reference: no dense Jacobians or Eigen/Ceres objects.
 * The Schwarp normal matrix is assembled directly in block-banded form (one block
 * per control-grid column, couplings reach 3 columns) by threads that *gather*
 * their entry from the sites/cells that touch it -- deterministic, no atomics --
 * and factorised with a sliding shared-memory window.
 *
 * Written against the Team abstraction of ds_common.h, so the same source also
 * compiles as a one-thread team for the CPU-only test tier (tests/emu/).
 */
#ifndef DS_NRSFM_CORE_H_
#define DS_NRSFM_CORE_H_
#include "../../include/defslam_b200.h"
#include "bbs_core.h"
#if defined(DS_NRSFM_PROF) && defined(__CUDACC__)
#include <stdio.h>
/* diagnostic build: the lap counters of ds_rowchol.h summed into a device array (block 0 prints them) */
__device__ long long g_rowprof[40];
enum { PF_X_WARP = 0, PF_X_BWD = 20 };
#define DS_PROF_LOCALS(name, n) long long name[n] = {}
#define DS_PROF_LAP(name, i, t) do { const long long now_ = clock64(); name[i] += now_ - t; t = now_; } while (0)
#define DS_PROF_T0M(var) long long var = clock64()
#define DS_PROF_FLUSH(name, n, idx, cond) do { if (blockIdx.x == 0 && (cond)) for (int i_ = 0; i_ < n; i_++) g_rowprof[idx + i_] += name[i_]; } while (0)
#endif
/* release fence of the progress counters: fence.acq_rel.cta instead of __threadfence_block() (fence.sc.cta).  The
 * initialisation solve is bound by the chain warp, which publishes twice per block row: 250 k -> 197 k cycles. */
#ifndef DS_ROWS_FENCE_LIGHT
#define DS_ROWS_FENCE_LIGHT 1
#endif
#include "ds_rowchol.h"

#if DS_CUDA
#define DS_HD __host__ __device__ __forceinline__
#else
#define DS_HD inline
#endif

namespace ds {

/* Ceres trust-region defaults (TrustRegionMinimizer / LevenbergMarquardtStrategy) */
constexpr double LM_INITIAL_RADIUS = 1e4;
constexpr double LM_MAX_RADIUS = 1e16;
constexpr double LM_MIN_RADIUS = 1e-32;
constexpr double LM_MIN_DIAG = 1e-6;
constexpr double LM_MAX_DIAG = 1e32;
constexpr double LM_MIN_REL_DECREASE = 1e-3;
constexpr int LM_MAX_INVALID = 5;
constexpr double SCHWARP_HUBER = 5.77; /* SchwarpDatabase.cc:205 */

DS_FN double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ===================================================================== *
 *  Block-banded SPD solver
 *  n = nb*bs unknowns; block (I, J) is non-zero for |I-J| <= kb.  Storage
 *  (global memory, L2 resident): block (I, I-d), d = 0..kb, row-major bs x bs at
 *  blk(I, d); the diagonal block is stored full.
 * ===================================================================== */
struct BandDims {
  int bs, nb, kb;
  DS_HD int Wn() const { return (kb + 1) * bs; }
  DS_HD int ldw() const { return Wn() | 1; }
  DS_HD int lbs(int nrhs) const { return Wn() + nrhs; }
  DS_HD size_t blk(int I, int d) const { return ((size_t)I * (kb + 1) + d) * bs * bs; }
  DS_HD size_t doubles() const { return (size_t)nb * (kb + 1) * bs * bs; }
  /* shared doubles needed by bband_solve */
  DS_HD size_t smem_doubles(int nrhs) const {
    return (size_t)(Wn() + nrhs) * ldw() + (size_t)bs * lbs(nrhs) + (size_t)nrhs * bs;
  }
};

/* entry (i, c), i >= c, of a band matrix */
DS_FN double band_at(const double *Hb, const BandDims &bd, int i, int c) {
  const int I = i / bd.bs, J = c / bd.bs;
  return Hb[bd.blk(I, I - J) + (size_t)(i - I * bd.bs) * bd.bs + (c - J * bd.bs)];
}

#if DS_CUDA
#define DS_TX 16
#else
#define DS_TX 1
#endif

/* Solves (diag(sc) H diag(sc) + diag(add)) y = rhs_k, k < nrhs (<= 2), in place in
 * rhs (shared memory, stride ldr).  sc/add may be null.  Lb receives the factor
 * (same layout as Hb).  sh: bd.smem_doubles(nrhs) doubles of shared memory.
 * Right-looking Cholesky on a ring-addressed window of kb+1 block rows; the
 * right-hand sides ride along as extra window rows (forward substitution for
 * free).  One barrier per eliminated column.  Returns false on a non-positive
 * pivot (uniformly over the team). */
DS_FN_NOINLINE bool bband_solve(const Team team, const BandDims bd, const double *Hb, double *Lb, const double *sc,
                                const double *add, double *sh, double *rhs, int nrhs, int ldr) {
  const int bs = bd.bs, nb = bd.nb, kb = bd.kb, Wn = bd.Wn(), ldw = bd.ldw(), LBS = bd.lbs(nrhs);
  double *W = sh, *Lbuf = sh + (size_t)(Wn + nrhs) * ldw, *tv = Lbuf + (size_t)bs * LBS;
  const int TX = DS_TX, tx = team.tid % TX, ty = team.tid / TX, NY = team.nthr / TX > 0 ? team.nthr / TX : 1;
  int off = 0;
#define DS_P(i) (((i) + off) >= Wn ? ((i) + off - Wn) : ((i) + off))
  /* window <- block rows 0..kb */
  {
    const int rows = (nb < kb + 1 ? nb : kb + 1) * bs;
    DS_FOR(idx, rows * Wn) {
      const int i = idx / Wn, c = idx - i * Wn;
      if (c > i) continue;
      double v = band_at(Hb, bd, i, c);
      if (sc) v *= sc[i] * sc[c];
      if (add && i == c) v += add[i];
      W[(size_t)i * ldw + c] = v;
    }
    DS_FOR(idx, nrhs * rows) {
      const int k = idx / rows, c = idx - k * rows;
      W[(size_t)(Wn + k) * ldw + c] = rhs[k * ldr + c];
    }
  }
  team.sync();
  for (int K = 0; K < nb; K++) {
    const int nrow = (kb + 1 < nb - K ? kb + 1 : nb - K) * bs;
    for (int j = 0; j < bs; j++) {
      const int pj = DS_P(j);
      const double piv = W[(size_t)pj * ldw + pj];
      if (!(piv > 0.0) || !(piv < DBL_MAX)) return false;
      const double isq = 1.0 / sqrt(piv), inv = 1.0 / piv;
      /* scaled column j -> panel buffer (rows j.., then the rhs rows) */
      DS_FOR(t, nrow - j + nrhs) {
        const int i = j + t;
        const int row = i < nrow ? DS_P(i) : Wn + (i - nrow);
        Lbuf[(size_t)j * LBS + (i < nrow ? i : Wn + (i - nrow))] = W[(size_t)row * ldw + pj] * isq;
      }
      /* trailing update of the lower triangle and of the rhs rows */
      for (int ii = j + 1 + ty; ii < nrow + nrhs; ii += NY) {
        const int row = ii < nrow ? DS_P(ii) : Wn + (ii - nrow);
        const double a = W[(size_t)row * ldw + pj] * inv;
        const int cmax = ii < nrow ? ii : nrow - 1;
        double *wr = W + (size_t)row * ldw;
        const double *wc = W + pj; /* column pj */
        /* the ring wraps at most once inside [j+1, cmax]: two plain strided segments */
        const int e1 = cmax < Wn - off - 1 ? cmax : Wn - off - 1;
        int c = j + 1 + tx;
#pragma unroll 4
        for (; c <= e1; c += TX) {
          const int pc = c + off;
          wr[pc] -= a * wc[(size_t)pc * ldw];
        }
#pragma unroll 4
        for (; c <= cmax; c += TX) {
          const int pc = c + off - Wn;
          wr[pc] -= a * wc[(size_t)pc * ldw];
        }
      }
      team.sync();
    }
    /* panel -> L (global), y_K -> rhs */
    DS_FOR(idx, bs * nrow) {
      const int j = idx / nrow, i = idx - j * nrow;
      if (i < j) continue;
      const int db = i / bs;
      Lb[bd.blk(K + db, db) + (size_t)(i - db * bs) * bs + j] = Lbuf[(size_t)j * LBS + i];
    }
    DS_FOR(idx, bs * nrhs) {
      const int k = idx / bs, j = idx - k * bs;
      rhs[k * ldr + K * bs + j] = Lbuf[(size_t)j * LBS + Wn + k];
    }
    team.sync();
    off += bs;
    if (off >= Wn) off -= Wn;
    if (K + kb + 1 < nb) {
      /* the freed ring slot becomes block row K+kb+1 (logical rows kb*bs..Wn-1 of step K+1) */
      const int g0 = (K + 1) * bs;
      DS_FOR(idx, bs * Wn) {
        const int i = kb * bs + idx / Wn, c = idx % Wn;
        if (c > i) continue;
        double v = band_at(Hb, bd, g0 + i, g0 + c);
        if (sc) v *= sc[g0 + i] * sc[g0 + c];
        if (add && i == c) v += add[g0 + i];
        W[(size_t)DS_P(i) * ldw + DS_P(c)] = v;
      }
      DS_FOR(idx, nrhs * bs) {
        const int k = idx / bs, c = kb * bs + idx % bs;
        W[(size_t)(Wn + k) * ldw + DS_P(c)] = rhs[k * ldr + g0 + c];
      }
    }
    team.sync();
  }
#undef DS_P
  /* backward sweep L^T x = y, one block column per step, staged through the panel buffer */
#if DS_CUDA
  const int warp = team.tid >> 5, nwarp = team.nthr >> 5 > 0 ? team.nthr >> 5 : 1;
#else
  const int warp = 0, nwarp = 1;
#endif
  for (int K = nb - 1; K >= 0; K--) {
    const int nrow = (kb + 1 < nb - K ? kb + 1 : nb - K) * bs;
    DS_FOR(idx, bs * nrow) {
      const int i = idx / bs, j = idx - i * bs;
      if (i < j) continue;
      const int db = i / bs;
      Lbuf[(size_t)j * LBS + i] = Lb[bd.blk(K + db, db) + (size_t)(i - db * bs) * bs + j];
    }
    team.sync();
    for (int k = warp; k < nrhs; k += nwarp) {
      double *x = rhs + k * ldr + K * bs, *t = tv + k * bs;
      DS_WARP_FOR(j, bs) {
        double s = x[j];
        const double *lc = Lbuf + (size_t)j * LBS;
        for (int i = bs; i < nrow; i++) s -= lc[i] * x[i];
        t[j] = s;
      }
      team.warp_sync();
      for (int i = bs - 1; i >= 0; i--) {
        const double xi = t[i] / Lbuf[(size_t)i * LBS + i];
        team.warp_sync();
        DS_WARP_FOR(j, i) t[j] -= Lbuf[(size_t)j * LBS + i] * xi;
        if (team.lane() == 0) x[i] = xi;
        team.warp_sync();
      }
    }
    team.sync();
  }
  return true;
}

/* ===================================================================== *
 *  The same solve on the FP64 tensor cores: row-owner banded Cholesky of
 *  ds_rowchol.h (the machinery of the SfT step).  The block band is read as a
 *  scalar band of half-bandwidth (kb+1)*bs - 1 in tiles of 8x8; n is padded to
 *  a multiple of 8 with identity rows; the right-hand sides are the border rows.
 * ===================================================================== */
constexpr int ROWS_SWEEP_BUFS = 8;

struct RowPlan {
  int nt;      /* tiles per block row (0: not available, use bband_solve) */
  int owners;  /* row-owner warps the ring has room for */
  int nblk, Dp;
  int bw;      /* scalar half-bandwidth; the row-band copy has bwE = bw rounded up to even, row stride bwE + 1 */
  int smem;    /* doubles of shared memory */
  int ws;      /* doubles of global workspace (tile-form factor, inverses of the diagonal blocks, border rows) */
};

/* instantiated tile counts (a band narrower than NT-1 tiles carries zero tiles on its left) */
static inline
#if DS_CUDA
__host__ __device__
#endif
int rows_nt_for(int nbk) {
  const int need = nbk + 1;
  const int have[] = {6, 8, 10, 12, 14, 16, 18};
  for (int i = 0; i < 7; i++) if (have[i] >= need) return have[i];
  return 0;
}

static inline
#if DS_CUDA
__host__ __device__
#endif
RowPlan rows_plan(int n, int bw, int avail_doubles, int nthreads) {
  RowPlan p{0, 0, 0, 0, 0, 0, 0};
#if DS_CUDA
  const int nt = rows_nt_for((bw + 7) / 8);
  if (nt == 0 || nthreads < 128) return p;
  const int Dp = round_up(n, 8), nblk = Dp / 8;
  const int fixed = nt * 64 + 128 + (3 * nblk + 2 + 32 + 2) / 2 + 2 + 2 * (Dp + 8);
  const int bsz = ROWS_SWEEP_BUFS * (nt * LT_STRIDE + 64);
  for (int own = 5; own >= 2; own--) {
    int w = (nt - 1 + own) * nt * 64;
    if (w < bsz) w = bsz;
    if (fixed + w <= avail_doubles) {
      p.nt = nt; p.owners = own; p.nblk = nblk; p.Dp = Dp; p.bw = bw; p.smem = fixed + w;
      p.ws = nblk * nt * LT_STRIDE + nblk * 64 + 16 * Dp + Dp * ((bw | 1) + 2);
      return p;
    }
  }
#else
  (void)n; (void)bw; (void)avail_doubles; (void)nthreads;
#endif
  return p;
}

#if DS_CUDA
/* The row owners read their block row from a row-band copy of diag(sc) H diag(sc) (row i holds columns i-bw..i at
 * offset c - i + bwE, stride ld = bwE + 1: the layout of the SfT band, 16-byte pairs), padded with identity rows.
 * It is written once per build of H -- not per solve: reading the block band directly cost the owners a table
 * look-up, a scaling product and an 8-byte load per entry, 9 k cycles per block row, a third of the factorisation. */
DS_FN double *rows_hr(const RowPlan &plan, double *wsg) {
  return wsg + (size_t)plan.nblk * plan.nt * LT_STRIDE + (size_t)plan.nblk * 64 + 16 * (size_t)plan.Dp;
}
DS_FN_NOINLINE void band_to_rows(const Team team, const BandDims bd, const RowPlan plan, const double *Hb, const double *sc,
                                 double *wsg) {
  const int n = bd.nb * bd.bs, bs = bd.bs, bw = plan.bw, bwE = (bw + 1) & ~1, ld = bwE + 1, W1 = bw + 1;
  double *Hr = rows_hr(plan, wsg);
  DS_FOR(idx, plan.Dp * W1) {
    const int i = idx / W1, c = i - bw + (idx - i * W1);
    double v = 0.0;
    if (i >= n) v = c == i ? 1.0 : 0.0;
    else if (c >= 0) {
      const int Ib = i / bs, Jb = c / bs;
      if (Ib - Jb <= bd.kb) {
        v = Hb[bd.blk(Ib, Ib - Jb) + (size_t)(i - Ib * bs) * bs + (c - Jb * bs)];
        if (sc) v *= sc[i] * sc[c];
      }
    }
    Hr[(size_t)i * ld + (c - i + bwE)] = v;
  }
  team.sync();
}

struct RowBandLoader {
  const double *Hr;
  int ld, bwE, bw;
  const double *add; /* added to the diagonal of rows < n (may be null) */
  int n;
  template <int NT>
  DS_FN void load(int I, int g, int q, double *a0, double *a1) const {
    constexpr int NBK = NT - 1;
    const int lo = bwE - bw;
    const int i = NB * I + g;
    const double *rowp = Hr + (size_t)i * ld;
    /* Straight-line for the tiles left of the diagonal: every pair is requested (from a clamped, valid address where
     * it lies outside the band) before any is used.  With a branch per tile the loads went out one memory round trip
     * after the other: 4 k cycles per block row of an owner, a quarter of its time, with the chain warp waiting. */
#pragma unroll
    for (int t = 0; t < NBK; t++) {
      const int off = NB * (I - NBK + t) + 2 * q - i + bwE; /* even, <= bwE - 2 */
      /* a row is 16-byte aligned at offsets of its own parity; off = -1 (odd row, its first column the second of the
       * pair) reads the last slot of the row above with it */
      const dbl2 w = *(const dbl2 *)(rowp + (off >= -1 ? off : (i & 1)));
      a0[t] = w.x; a1[t] = w.y;
    }
    {
      /* the diagonal tile: columns up to the diagonal (offset bwE) only */
      const int off = NB * I + 2 * q - i + bwE;
      double x0 = 0.0, x1 = 0.0;
      if (off + 1 <= bwE) { const dbl2 w = *(const dbl2 *)(rowp + off); x0 = w.x; x1 = w.y; }
      else if (off == bwE) x0 = rowp[off];
      a0[NBK] = x0; a1[NBK] = x1;
    }
#pragma unroll
    for (int t = 0; t < NBK; t++) {
      const int J = I - NBK + t;
      const int off = NB * J + 2 * q - i + bwE;
      a0[t] = (J >= 0 && off >= lo) ? a0[t] : 0.0;
      a1[t] = (J >= 0 && off + 1 >= lo) ? a1[t] : 0.0;
    }
    const double dadd = (add && i < n) ? add[i] : 0.0;
    if (2 * q == g) a0[NT - 1] += dadd;
    if (2 * q + 1 == g) a1[NT - 1] += dadd;
  }
  DS_FN void prefetch(int I, int lane) const {
    const char *seg = (const char *)(Hr + (size_t)(NB * I + (lane >> 2)) * ld + (bwE - bw));
    const int nline = ((bw + 1) * 8 + 127) / 128 + 1;
    for (int l = (lane & 3); l < nline; l += 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(seg + 128 * l));
  }
};

/* Same contract as bband_solve() with the matrix taken from the row-band copy (band_to_rows); wsg: plan.ws doubles
 * of global workspace; sh: plan.smem doubles. */
template <int NT>
DS_FN_NOINLINE bool bband_solve_rows(const Team team, const BandDims bd, const RowPlan plan, double *wsg, const double *add,
                                     double *sh, double *rhs, int nrhs, int ldr) {
  const int n = bd.nb * bd.bs, Dp = plan.Dp, nblk = plan.nblk;
  const int R = NT - 1 + plan.owners;
  int wsz = R * NT * 64;
  { const int bsz = ROWS_SWEEP_BUFS * (NT * LT_STRIDE + 64); if (wsz < bsz) wsz = bsz; }
  double *W = sh, *er = W + wsz, *db = er + NT * 64, *dx = db + 128, *sol = dx + Dp + 8;
  int *sy = (int *)(sol + Dp + 8), *flag = sy + 3 * nblk + 2 + 32 + 1;
  double *Lt = wsg, *Dinv = Lt + (size_t)nblk * NT * LT_STRIDE, *Cg = Dinv + (size_t)nblk * 64, *Eg = Cg + 8 * (size_t)Dp;
  const int bwE = (plan.bw + 1) & ~1;
#if defined(DS_NRSFM_PROF)
  const long long q0 = clock64();
#endif
  /* the right-hand sides as border rows (the other rows of the border tile are zero) */
  DS_FOR(idx, 8 * Dp) {
    const int r = idx / Dp, i = idx - r * Dp;
    Cg[idx] = (r < nrhs && i < n) ? rhs[r * ldr + i] : 0.0;
  }
  team.sync();
#if defined(DS_NRSFM_PROF)
  const long long q1 = clock64();
#endif
  rows_factor<NT>(team.tid, team.nthr, nblk, plan.owners, RowBandLoader{rows_hr(plan, wsg), bwE + 1, bwE, plan.bw, add, n}, W, er,
                  db, sy, flag, Cg, Eg, Dp, (double *)nullptr, (const double *)nullptr, 0.0, Lt, Dinv);
#if defined(DS_NRSFM_PROF)
  const long long q2 = clock64();
#endif
  if (*flag != 0) { team.sync(); return false; }
  for (int r = 0; r < nrhs; r++) {
    DS_FOR(i, Dp) dx[i] = Eg[(size_t)r * Dp + i];
    DS_FOR(i, nblk + 1) sy[i] = 0;
    team.sync();
    rows_backward<NT, ROWS_SWEEP_BUFS>(team.tid, team.nthr, nblk, W, sy, Lt, Dinv, dx, sol);
    team.sync();
    DS_FOR(i, n) rhs[r * ldr + i] = sol[i];
    team.sync();
  }
#if defined(DS_NRSFM_PROF)
  if (team.tid == 0 && blockIdx.x == 0) {
    printf("[nrsfm prof] NT %d nblk %d owners %d: staging %lld, factorisation %lld, backward sweeps %lld cycles; chain wait %lld busy %lld; owners (load / steps):"
           " %lld/%lld %lld/%lld %lld/%lld %lld/%lld %lld/%lld\n", NT, nblk,
           plan.owners, q1 - q0, q2 - q1, clock64() - q2, g_rowprof[0], g_rowprof[1], g_rowprof[2], g_rowprof[3], g_rowprof[4],
           g_rowprof[5], g_rowprof[6], g_rowprof[7], g_rowprof[8], g_rowprof[9], g_rowprof[10], g_rowprof[11]);
    for (int i = 0; i < 40; i++) g_rowprof[i] = 0;
  }
#endif
  return true;
}
#endif

/* dispatch: tensor-core row form where the plan has one (the caller has run band_to_rows since H or sc last
 * changed), the scalar window otherwise */
DS_FN bool bband_solve_any(const Team team, const BandDims bd, const RowPlan plan, const double *Hb, double *Lb,
                           const double *sc, const double *add, double *sh, double *rhs, int nrhs, int ldr) {
#if DS_CUDA
  switch (plan.nt) {
    case 6: return bband_solve_rows<6>(team, bd, plan, Lb, add, sh, rhs, nrhs, ldr);
    case 8: return bband_solve_rows<8>(team, bd, plan, Lb, add, sh, rhs, nrhs, ldr);
    case 10: return bband_solve_rows<10>(team, bd, plan, Lb, add, sh, rhs, nrhs, ldr);
    case 12: return bband_solve_rows<12>(team, bd, plan, Lb, add, sh, rhs, nrhs, ldr);
    case 14: return bband_solve_rows<14>(team, bd, plan, Lb, add, sh, rhs, nrhs, ldr);
    case 16: return bband_solve_rows<16>(team, bd, plan, Lb, add, sh, rhs, nrhs, ldr);
    case 18: return bband_solve_rows<18>(team, bd, plan, Lb, add, sh, rhs, nrhs, ldr);
    default: break;
  }
#else
  (void)plan;
#endif
  return bband_solve(team, bd, Hb, Lb, sc, add, sh, rhs, nrhs, ldr);
}

/* ===================================================================== *
 *  Sites sorted by knot cell (shared by the Schwarp data term and SfN)
 * ===================================================================== */
struct CellSort {
  int n, ncu, ncv; /* ncu x ncv knot cells */
  int *cell;       /* [n]     cell of each site, -1 outside the domain  */
  int *start;      /* [ncu*ncv+1]                                        */
  int *perm;       /* [n]     sites grouped by cell, original order kept */
  double *taps;    /* [n*8]   bu[4], bv[4] of each site                  */
};

/* fills cs from fp32 sites (x, y interleaved).  Returns the number of sites outside
 * the spline domain (the reference's colocEigen aborts on any, bbs_coloc.cc:106-109). */
DS_FN_NOINLINE int cell_sort(const Team team, const BbsView &s, const float *xy, CellSort &cs, double *red) {
  int bad = 0;
  DS_FOR(m, cs.n) {
    double nu, nv;
    int Iu, Iv;
    bbs_normalize(s.umin, s.umax, s.nptsu, (double)xy[2 * m], nu, Iu);
    bbs_normalize(s.vmin, s.vmax, s.nptsv, (double)xy[2 * m + 1], nv, Iv);
    if (!bbs_in_domain(s, Iu, Iv)) { cs.cell[m] = -1; bad++; continue; }
    cs.cell[m] = Iu * cs.ncv + Iv;
    bbs_basis(0, nu, cs.taps + 8 * (size_t)m);
    bbs_basis(0, nv, cs.taps + 8 * (size_t)m + 4);
  }
  bad = team_sum_int(team, bad, red);
  const int nc = cs.ncu * cs.ncv;
#if DS_CUDA
  /* one warp per cell, 32 sites per trip: count by ballot, place by the prefix of the ballot (original order kept) */
  const int warp = team.tid >> 5, lane = team.tid & 31, nwarp = team.nthr >> 5;
  for (int c = warp; c < nc; c += nwarp) {
    int cnt = 0;
    for (int m0 = 0; m0 < cs.n; m0 += 32) {
      const int m = m0 + lane;
      cnt += __popc(__ballot_sync(0xffffffffu, m < cs.n && cs.cell[m] == c));
    }
    if (lane == 0) cs.start[c + 1] = cnt;
  }
#else
  DS_FOR(c, nc) {
    int cnt = 0;
    for (int m = 0; m < cs.n; m++) cnt += cs.cell[m] == c;
    cs.start[c + 1] = cnt;
  }
#endif
  team.sync();
  if (team.tid == 0) {
    cs.start[0] = 0;
    for (int c = 0; c < nc; c++) cs.start[c + 1] += cs.start[c];
  }
  team.sync();
#if DS_CUDA
  for (int c = warp; c < nc; c += nwarp) {
    int o = cs.start[c];
    for (int m0 = 0; m0 < cs.n; m0 += 32) {
      const int m = m0 + lane;
      const bool mine = m < cs.n && cs.cell[m] == c;
      const unsigned bal = __ballot_sync(0xffffffffu, mine);
      if (mine) cs.perm[o + __popc(bal & ((1u << lane) - 1u))] = m;
      o += __popc(bal);
    }
  }
#else
  DS_FOR(c, nc) {
    int o = cs.start[c];
    for (int m = 0; m < cs.n; m++)
      if (cs.cell[m] == c) cs.perm[o++] = m;
  }
#endif
  team.sync();
  return bad;
}

/* cell integrals of basis-derivative products, [order][a][b] (48 doubles) */
DS_FN void fill_cell_integrals(const Team team, double *ci) {
  DS_FOR(t, 48) ci[t] = bbs_cell_integral(t / 16, (t / 4) & 3, t & 3);
  team.sync();
}

/* bending matrix entry from the cell-integral table (same sum as bbs_bending_entry); the three scale factors of the
 * grid are formed once per problem (five divisions), not once per entry */
struct BendCoef { double cxx, cxy, cyy; };
DS_FN BendCoef bending_coef(const BbsView &s) {
  const double sy = (s.umax - s.umin) / (s.nptsu - 3);
  const double sx = (s.vmax - s.vmin) / (s.nptsv - 3);
  return BendCoef{sy / (sx * sx * sx), 1.0 / (sx * sy), sx / (sy * sy * sy)};
}
DS_FN double bending_entry_tab(const BbsView &s, const BendCoef &bc, const double *ci, int iu, int iv, int ju, int jv) {
  const int du = iu > ju ? iu - ju : ju - iu, dv = iv > jv ? iv - jv : jv - iv;
  if (du > 3 || dv > 3) return 0.0;
  const int nx = s.nptsv, ny = s.nptsu;
  const double cxx = bc.cxx, cxy = bc.cxy, cyy = bc.cyy;
  double acc = 0.0;
  const int b0 = (iu > ju ? iu : ju) - 3, b1 = iu < ju ? iu : ju;
  const int a0 = (iv > jv ? iv : jv) - 3, a1 = iv < jv ? iv : jv;
  for (int b = (b0 > 0 ? b0 : 0); b <= b1 && b <= ny - 4; b++)
    for (int a = (a0 > 0 ? a0 : 0); a <= a1 && a <= nx - 4; a++) {
      const int e1 = iu - b, e2 = ju - b, f1 = iv - a, f2 = jv - a;
      const double u0 = ci[e1 * 4 + e2], u1 = ci[16 + e1 * 4 + e2], u2 = ci[32 + e1 * 4 + e2];
      const double v0 = ci[f1 * 4 + f2], v1 = ci[16 + f1 * 4 + f2], v2 = ci[32 + f1 * 4 + f2];
      acc += cxx * (u0 * v2) + cxy * (2.0 * u1 * v1) + cyy * (u2 * v0);
    }
  return acc;
}
DS_FN double bending_entry_tab(const BbsView &s, const double *ci, int iu, int iv, int ju, int jv) {
  return bending_entry_tab(s, bending_coef(s), ci, iu, iv, ju, jv);
}

/* ===================================================================== *
 *  Schwarp fit
 * ===================================================================== */
struct SchwarpProb {
  BbsView bbs; /* valdim 2 */
  int n;
  const float *kp1, *kp2, *isig;
  double lambda, fx, fy, px_fx, px_fy;
  int max_iterations, initialize;
  double *x; /* [2*NC] in/out, [all x; all y] */
  /* outputs (device) */
  float *warp_uv, *J12, *J21, *H12;
  uint8_t *keep;
  double *scalars; /* cost_initial, cost_final, iterations, accepted, status */
};

/* per-CTA scratch in global memory */
struct SchwarpWs {
  int *cell, *cstart, *perm;
  double *taps;   /* [n*8]                      */
  double *CtC;    /* band, bs = nptsv           */
  double *Hb;     /* band, bs = 2*nptsv         */
  double *Lb;     /* band (factor), bs = 2*nptsv */
  double *Js;     /* [NC][4][16][2] Schwarzian Jacobian taps */
  double *rdata;  /* [2n]                       */
  double *sdv;    /* [10*NC] + [4*NC]: site derivatives and Schwarzian residuals when they do
                     not fit in shared memory next to the solver window (large grids) */
};

struct SchwarpSizes {
  size_t cell, cstart, perm, taps, CtC, Hb, Lb, Js, rdata, sdv, total;
};

/* shared-memory carve-up (doubles); sdv < 0: site arrays live in the global workspace */
struct SchwarpSmem {
  int x, xc, g, scale, add, step, sdv, rs, ci, red, su, ints, solver, total;
  RowPlan p1, p2; /* tensor-core row form of the initialisation solve (bs = nptsv) and of the LM solve (bs = 2 nptsv) */
};
constexpr int SCHWARP_SMEM_LIMIT_DOUBLES = 227 * 1024 / 8 - 64;
#ifndef NRSFM_THREADS_
#define NRSFM_THREADS_ 512
#endif
static inline
#if DS_CUDA
__host__ __device__
#endif
SchwarpSmem schwarp_smem(int nptsu, int nptsv) {
  SchwarpSmem m;
  const int NC = nptsu * nptsv, NP = 2 * NC;
  BandDims b1{nptsv, nptsu, 3}, b2{2 * nptsv, nptsu, 3};
  const int nints = (3 * (nptsu + nptsv) + 1) / 2 + 2;
  const int base = 6 * NP + 48 + 40 + nints + 1;
  /* the solver region: the row form of the LM solve if its ring fits (with the site arrays in shared memory if
   * that fits too, else with them in the global workspace), else the scalar window */
  bool sites_in_smem = false;
  m.p1 = m.p2 = RowPlan{0, 0, 0, 0, 0, 0, 0};
  int solver = 0;
  for (int pass = 0; pass < 2 && m.p2.nt == 0; pass++) {
    const int avail = SCHWARP_SMEM_LIMIT_DOUBLES - base - (pass == 0 ? 14 * NC : 0);
    const RowPlan p2 = rows_plan(NP, 4 * b2.bs - 1, avail, NRSFM_THREADS_);
    if (p2.nt == 0) continue;
    m.p2 = p2;
    sites_in_smem = pass == 0;
    solver = p2.smem;
    const RowPlan p1 = rows_plan(NC, 4 * b1.bs - 1, avail, NRSFM_THREADS_);
    if (p1.nt > 0) { m.p1 = p1; if (p1.smem > solver) solver = p1.smem; }
    else if ((int)b1.smem_doubles(2) <= avail) { if ((int)b1.smem_doubles(2) > solver) solver = (int)b1.smem_doubles(2); }
    else { m.p2 = RowPlan{0, 0, 0, 0, 0, 0, 0}; } /* (cannot happen for grids the scalar path accepts) */
  }
  if (m.p2.nt == 0) {
    solver = (int)b2.smem_doubles(2);
    sites_in_smem = base + 14 * NC + solver <= SCHWARP_SMEM_LIMIT_DOUBLES;
  }
  int o = 0;
  m.x = o; o += NP;
  m.xc = o; o += NP;
  m.g = o; o += NP;      /* gradient, block order */
  m.scale = o; o += NP;  /* Jacobi scaling, block order */
  m.add = o; o += NP;
  m.step = o; o += NP;   /* rhs / solution (2 columns for the initialisation: NP >= 2*NC) */
  m.sdv = sites_in_smem ? o : -1; /* per site: xu yu xv yv xuu yuu xvv yvv xuv yuv */
  if (sites_in_smem) o += 10 * NC;
  m.rs = sites_in_smem ? o : -1;
  if (sites_in_smem) o += 4 * NC;
  m.ci = o; o += 48;
  m.red = o; o += 40;
  m.su = o; o += 0;
  m.ints = o; o += nints; /* site intervals + lo/hi site ranges (ints) */
  o = (o + 1) & ~1;       /* the solver region holds 16-byte tiles */
  m.solver = o; o += solver;
  m.total = o;
  return m;
}

static inline
#if DS_CUDA
__host__ __device__
#endif
SchwarpSizes schwarp_ws_sizes(int nptsu, int nptsv, int nmax) {
  SchwarpSizes z;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t NC = (size_t)nptsu * nptsv;
  size_t o = 0;
  z.cell = o; o += al(sizeof(int) * nmax);
  z.cstart = o; o += al(sizeof(int) * ((size_t)(nptsu - 3) * (nptsv - 3) + 1));
  z.perm = o; o += al(sizeof(int) * nmax);
  z.taps = o; o += al(sizeof(double) * 8 * nmax);
  z.CtC = o; o += al(sizeof(double) * nptsu * 4 * nptsv * nptsv);
  z.Hb = o; o += al(sizeof(double) * nptsu * 4 * 4 * nptsv * nptsv);
  {
    const SchwarpSmem m = schwarp_smem(nptsu, nptsv);
    size_t lb = (size_t)nptsu * 4 * 4 * nptsv * nptsv;
    if ((size_t)m.p1.ws > lb) lb = (size_t)m.p1.ws;
    if ((size_t)m.p2.ws > lb) lb = (size_t)m.p2.ws;
    const size_t gram = (size_t)(nptsu - 3) * (nptsv - 3) * 256; /* per-cell Gram matrices of the C'C build */
    if (gram > lb) lb = gram;
    z.Lb = o; o += al(sizeof(double) * lb);
  }
  z.Js = o; o += al(sizeof(double) * NC * 128);
  z.rdata = o; o += al(sizeof(double) * 2 * nmax);
  z.sdv = o; o += al(sizeof(double) * 14 * NC);
  z.total = o;
  return z;
}

DS_FN double *site_derivs(const SchwarpWs &ws, double *sh, const SchwarpSmem &m) { return m.sdv >= 0 ? sh + m.sdv : ws.sdv; }
DS_FN double *site_resid(const SchwarpWs &ws, double *sh, const SchwarpSmem &m, int NC) {
  return m.rs >= 0 ? sh + m.rs : ws.sdv + 10 * (size_t)NC;
}

struct SiteMap {
  int *Su, *Sv;   /* knot interval of grid site i / j                 */
  int *lou, *hiu; /* site range whose 4-tap window covers control pu  */
  int *lov, *hiv;
};

DS_FN SiteMap site_map(double *base, int nptsu, int nptsv) {
  SiteMap sm;
  int *p = (int *)base;
  sm.Su = p; p += nptsu;
  sm.lou = p; p += nptsu;
  sm.hiu = p; p += nptsu;
  sm.Sv = p; p += nptsv;
  sm.lov = p; p += nptsv;
  sm.hiv = p;
  return sm;
}

/* grid site coordinate (Schwarp.cc:321-331), clamped onto the domain where rounding
 * would push it past the end (the reference would index out of range there) */
DS_FN double grid_coord(double lo, double hi, int i, int npts) {
  double x = (double)((hi - lo) * i) / (npts - 1) + lo;
  if (x > hi) x = hi;
  if (x < lo) x = lo;
  return x;
}

/* derivative taps of one grid coordinate: b0 (value unused), b1 = d/dx, b2 = d2/dx2 weights */
DS_FN void site_axis(double lo, double hi, int npts, int i, int &I, double b[3][4]) {
  double nx;
  bbs_normalize(lo, hi, npts, grid_coord(lo, hi, i, npts), nx, I);
  bbs_basis(0, nx, b[0]);
  bbs_basis(1, nx, b[1]);
  bbs_basis(2, nx, b[2]);
}

/* control value c(pu, pv, coord) from x laid out [all x; all y] */
DS_FN double ctrl_at(const double *x, int NC, int nptsv, int pu, int pv, int coord) {
  return x[coord * NC + pu * nptsv + pv];
}

/* Residuals at xs: data residuals -> ws.rdata (if store), Schwarzian residuals -> sm.rs and
 * site derivatives -> sm.sdv (if store).  Returns 0.5*(rho(sd) + ss); *rho1 = rho'. */
DS_FN_NOINLINE double schwarp_eval(const Team team, const SchwarpProb &P, const SchwarpWs &ws, double *sh,
                                   const SchwarpSmem &m, const double *xs, bool store, double *rho1) {
  const BbsView &s = P.bbs;
  const int NC = s.nptsu * s.nptsv;
  double sd = 0.0;
  DS_FOR(i, P.n) {
    const int cell = ws.cell[i];
    const int Iu = cell / (s.nptsv - 3), Iv = cell - Iu * (s.nptsv - 3);
    const double *tp = ws.taps + 8 * (size_t)i;
    double wx = 0.0, wy = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const double bas = tp[a] * tp[4 + b];
        const int l = (Iu + a) * s.nptsv + Iv + b;
        wx += xs[l] * bas;
        wy += xs[NC + l] * bas;
      }
    const double is = (double)P.isig[i];
    const double rx = is * ((double)P.kp2[2 * i] - wx) * P.fx;
    const double ry = is * ((double)P.kp2[2 * i + 1] - wy) * P.fy;
    if (store) { ws.rdata[i] = rx; ws.rdata[P.n + i] = ry; }
    sd += rx * rx + ry * ry;
  }
  sd = team_sum(team, sd, sh + m.red);
  double ss = 0.0;
  const double fu = bbs_deriv_fact(s, 1, 0), fv = bbs_deriv_fact(s, 0, 1);
  DS_FOR(k, NC) {
    const int i = k / s.nptsv, j = k - i * s.nptsv;
    int Iu, Iv;
    double bu[3][4], bv[3][4];
    site_axis(s.umin, s.umax, s.nptsu, i, Iu, bu);
    site_axis(s.vmin, s.vmax, s.nptsv, j, Iv, bv);
    double d[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++) {
        const int l = (Iu + a) * s.nptsv + Iv + b;
        const double cx = xs[l], cy = xs[NC + l];
        const double w10 = bu[1][a] * bv[0][b], w01 = bu[0][a] * bv[1][b];
        const double w20 = bu[2][a] * bv[0][b], w02 = bu[0][a] * bv[2][b], w11 = bu[1][a] * bv[1][b];
        d[0] += cx * w10; d[1] += cy * w10;
        d[2] += cx * w01; d[3] += cy * w01;
        d[4] += cx * w20; d[5] += cy * w20;
        d[6] += cx * w02; d[7] += cy * w02;
        d[8] += cx * w11; d[9] += cy * w11;
      }
    d[0] *= fu; d[1] *= fu; d[2] *= fv; d[3] *= fv;
    d[4] *= fu * fu; d[5] *= fu * fu; d[6] *= fv * fv; d[7] *= fv * fv; d[8] *= fu * fv; d[9] *= fu * fv;
    const double xu = d[0], yu = d[1], xv = d[2], yv = d[3], xuu = d[4], yuu = d[5], xvv = d[6], yvv = d[7],
                 xuv = d[8], yuv = d[9];
    const double lam = P.lambda;
    const double r0 = (xuu * yu - yuu * xu) * lam;
    const double r1 = (yvv * xv - xvv * yv) * lam;
    const double r2 = (xuu * yv - yuu * xv + 2 * (xuv * yu - yuv * xu)) * lam;
    const double r3 = (yvv * xu - xvv * yu + 2 * (yuv * xv - xuv * yv)) * lam;
    if (store) {
      double *sv = site_derivs(ws, sh, m) + 10 * k;
      for (int q = 0; q < 10; q++) sv[q] = d[q];
      double *rs = site_resid(ws, sh, m, NC);
      rs[k] = r0; rs[NC + k] = r1; rs[2 * NC + k] = r2; rs[3 * NC + k] = r3;
    }
    ss += r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
  }
  ss = team_sum(team, ss, sh + m.red);
  double rho;
  const double b = SCHWARP_HUBER * SCHWARP_HUBER;
  if (sd > b) {
    const double r = sqrt(sd);
    const double r1 = SCHWARP_HUBER / r;
    *rho1 = r1 > DBL_MIN ? r1 : DBL_MIN;
    rho = 2.0 * SCHWARP_HUBER * r - b;
  } else {
    *rho1 = 1.0;
    rho = sd;
  }
  return 0.5 * (rho + ss);
}

/* Schwarzian Jacobian taps of every grid site: Js[site][row 0..3][tap a*4+b][coord] (Schwarp.cc:461-512) */
DS_FN_NOINLINE void schwarp_site_jacobians(const Team team, const SchwarpProb &P, const SchwarpWs &ws, double *sh,
                                           const SchwarpSmem &m) {
  const BbsView &s = P.bbs;
  const int NC = s.nptsu * s.nptsv;
  const double fu = bbs_deriv_fact(s, 1, 0), fv = bbs_deriv_fact(s, 0, 1), lam = P.lambda;
  DS_FOR(idx, NC * 16) {
    const int k = idx >> 4, t = idx & 15, a = t >> 2, b = t & 3;
    const int i = k / s.nptsv, j = k - i * s.nptsv;
    int Iu, Iv;
    double bu[3][4], bv[3][4];
    site_axis(s.umin, s.umax, s.nptsu, i, Iu, bu);
    site_axis(s.vmin, s.vmax, s.nptsv, j, Iv, bv);
    const double cu = fu * bu[1][a] * bv[0][b], cv = fv * bu[0][a] * bv[1][b];
    const double cuu = fu * fu * bu[2][a] * bv[0][b], cvv = fv * fv * bu[0][a] * bv[2][b],
                 cuv = fu * fv * bu[1][a] * bv[1][b];
    const double *d = site_derivs(ws, sh, m) + 10 * k;
    const double xu = d[0], yu = d[1], xv = d[2], yv = d[3], xuu = d[4], yuu = d[5], xvv = d[6], yvv = d[7],
                 xuv = d[8], yuv = d[9];
    double *J = ws.Js + (size_t)k * 128 + 2 * t;
    J[0] = lam * (yu * cuu - yuu * cu);
    J[1] = lam * (xuu * cu - xu * cuu);
    J[32] = lam * (yvv * cv - yv * cvv);
    J[33] = lam * (xv * cvv - xvv * cv);
    J[64] = lam * (yv * cuu - yuu * cv + 2 * yu * cuv - 2 * yuv * cu);
    J[65] = lam * (xuu * cv - xv * cuu + 2 * xuv * cu - 2 * xu * cuv);
    J[96] = lam * (yvv * cu - yu * cvv - 2 * yv * cuv + 2 * yuv * cv);
    J[97] = lam * (xu * cvv - xvv * cu - 2 * xuv * cv + 2 * xv * cuv);
  }
  team.sync();
}

/* gradient (unscaled, block order) and the band of J'J at the stored state.
 * Data block: rows i and i+n both carry -fx*C_i on the x columns (quirk C6), times the
 * corrector sqrt(rho'). */
DS_FN_NOINLINE void schwarp_build(const Team team, const SchwarpProb &P, const SchwarpWs &ws, double *sh,
                                  const SchwarpSmem &m, double rho1) {
  const BbsView &s = P.bbs;
  const int nu = s.nptsu, nv = s.nptsv, NC = nu * nv, ncv = nv - 3;
  const SiteMap sm = site_map(sh + m.ints, nu, nv);
  const double *rs = site_resid(ws, sh, m, NC);
  double *g = sh + m.g;
  /* gradient */
  DS_FOR(Pi, 2 * NC) {
    const int pu = Pi / (2 * nv), rem = Pi - pu * 2 * nv, pv = rem >> 1, co = rem & 1;
    double acc = 0.0;
    for (int i = sm.lou[pu]; i <= sm.hiu[pu]; i++)
      for (int j = sm.lov[pv]; j <= sm.hiv[pv]; j++) {
        const int k = i * nv + j, t = (pu - sm.Su[i]) * 4 + (pv - sm.Sv[j]);
        const double *J = ws.Js + (size_t)k * 128 + 2 * t + co;
        acc += J[0] * rs[k] + J[32] * rs[NC + k] + J[64] * rs[2 * NC + k] + J[96] * rs[3 * NC + k];
      }
    if (co == 0) {
      double dsum = 0.0;
      for (int Iu = (pu - 3 > 0 ? pu - 3 : 0); Iu <= pu && Iu <= nu - 4; Iu++)
        for (int Iv = (pv - 3 > 0 ? pv - 3 : 0); Iv <= pv && Iv <= nv - 4; Iv++) {
          const int c = Iu * ncv + Iv;
          for (int q = ws.cstart[c]; q < ws.cstart[c + 1]; q++) {
            const int mi = ws.perm[q];
            const double *tp = ws.taps + 8 * (size_t)mi;
            dsum += tp[pu - Iu] * tp[4 + pv - Iv] * (ws.rdata[mi] + ws.rdata[P.n + mi]);
          }
        }
      acc += -P.fx * rho1 * dsum;
    }
    g[Pi] = acc;
  }
  /* band of H.  One work item per coupled pair of control points ((pu, pv), (qu, qv)), qu = pu - d, |pv - qv| <= 3:
   * its 2 x 2 entries (x/y of either point) share the sites and the taps, so the site Jacobians are read as 16-byte
   * pairs once for four sums; every other entry of the band is a structural zero (filled first). */
  const BandDims b2{2 * nv, nu, 3}, b1{nv, nu, 3};
  const int bs = b2.bs;
  const double wdata = 2.0 * P.fx * P.fx * rho1;
  DS_FOR(idx, nu * 4 * bs * bs) ws.Hb[idx] = 0.0;
  team.sync();
  DS_FOR(item, nu * 4 * nv * 7) {
    const int e = item % (nv * 7), Id = item / (nv * 7), d = Id & 3, I = Id >> 2;
    const int pv = e / 7, qv = pv - 3 + (e - pv * 7);
    const int pu = I, qu = I - d;
    if (qu < 0 || qv < 0 || qv >= nv) continue;
    const int i0 = sm.lou[pu] > sm.lou[qu] ? sm.lou[pu] : sm.lou[qu];
    const int i1 = sm.hiu[pu] < sm.hiu[qu] ? sm.hiu[pu] : sm.hiu[qu];
    const int j0 = sm.lov[pv] > sm.lov[qv] ? sm.lov[pv] : sm.lov[qv];
    const int j1 = sm.hiv[pv] < sm.hiv[qv] ? sm.hiv[pv] : sm.hiv[qv];
    double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0; /* a[cp][cq] */
    for (int i = i0; i <= i1; i++)
      for (int j = j0; j <= j1; j++) {
        const int k = i * nv + j;
        const int tp = (pu - sm.Su[i]) * 4 + (pv - sm.Sv[j]), tq = (qu - sm.Su[i]) * 4 + (qv - sm.Sv[j]);
        const dbl2 *Jp = (const dbl2 *)(ws.Js + (size_t)k * 128 + 2 * tp), *Jq = (const dbl2 *)(ws.Js + (size_t)k * 128 + 2 * tq);
        const dbl2 p0 = Jp[0], p1 = Jp[16], p2 = Jp[32], p3 = Jp[48];
        const dbl2 q0 = Jq[0], q1 = Jq[16], q2 = Jq[32], q3 = Jq[48];
        a00 += p0.x * q0.x + p1.x * q1.x + p2.x * q2.x + p3.x * q3.x;
        a01 += p0.x * q0.y + p1.x * q1.y + p2.x * q2.y + p3.x * q3.y;
        a10 += p0.y * q0.x + p1.y * q1.x + p2.y * q2.x + p3.y * q3.x;
        a11 += p0.y * q0.y + p1.y * q1.y + p2.y * q2.y + p3.y * q3.y;
      }
    a00 += wdata * ws.CtC[b1.blk(I, d) + (size_t)pv * nv + qv];
    double *out = ws.Hb + b2.blk(I, d) + (size_t)(2 * pv) * bs + 2 * qv;
    out[0] = a00; out[1] = a01;
    out[bs] = a10; out[bs + 1] = a11;
  }
  team.sync();
}

/* DiffProp records of every match (SchwarpDatabase.cc:243-345): fp32 like cv::KeyPoint */
DS_FN_NOINLINE void schwarp_diffprop(const Team team, const SchwarpProb &P, const SchwarpWs &ws, const double *xs) {
  const BbsView &s = P.bbs;
  const int NC = s.nptsu * s.nptsv;
  const double fu = bbs_deriv_fact(s, 1, 0), fv = bbs_deriv_fact(s, 0, 1);
  DS_FOR(i, P.n) {
    double nu, nv, bu[3][4], bv[3][4];
    int Iu, Iv;
    bbs_normalize(s.umin, s.umax, s.nptsu, (double)P.kp1[2 * i], nu, Iu);
    bbs_normalize(s.vmin, s.vmax, s.nptsv, (double)P.kp1[2 * i + 1], nv, Iv);
    for (int o = 0; o < 3; o++) { bbs_basis(o, nu, bu[o]); bbs_basis(o, nv, bv[o]); }
    double d[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++) {
        const int l = (Iu + a) * s.nptsv + Iv + b;
        const double cx = xs[l], cy = xs[NC + l];
        const double w00 = bu[0][a] * bv[0][b], w10 = bu[1][a] * bv[0][b], w01 = bu[0][a] * bv[1][b];
        const double w20 = bu[2][a] * bv[0][b], w11 = bu[1][a] * bv[1][b], w02 = bu[0][a] * bv[2][b];
        d[0] += cx * w00; d[1] += cy * w00;
        d[2] += cx * w10; d[3] += cy * w10;
        d[4] += cx * w01; d[5] += cy * w01;
        d[6] += cx * w20; d[7] += cy * w20;
        d[8] += cx * w11; d[9] += cy * w11;
        d[10] += cx * w02; d[11] += cy * w02;
      }
    const float qx = (float)d[0], qy = (float)d[1];
    const float dux = (float)(d[2] * fu), duy = (float)(d[3] * fu), dvx = (float)(d[4] * fv), dvy = (float)(d[5] * fv);
    if (P.warp_uv) { P.warp_uv[2 * i] = qx; P.warp_uv[2 * i + 1] = qy; }
    if (P.keep) {
#if DS_CUDA
      const float ex = __fmul_rn(__fsub_rn(qx, P.kp2[2 * i]), (float)P.px_fx);
      const float ey = __fmul_rn(__fsub_rn(qy, P.kp2[2 * i + 1]), (float)P.px_fy);
#else
      const float ex = (qx - P.kp2[2 * i]) * (float)P.px_fx, ey = (qy - P.kp2[2 * i + 1]) * (float)P.px_fy;
#endif
      P.keep[i] = !(sqrt((double)ex * ex + (double)ey * ey) > 10);
    }
    if (P.J12) { P.J12[4 * i] = dux; P.J12[4 * i + 1] = duy; P.J12[4 * i + 2] = dvx; P.J12[4 * i + 3] = dvy; }
    if (P.J21) {
#if DS_CUDA
      const float det = __fsub_rn(__fmul_rn(dux, dvy), __fmul_rn(dvx, duy));
      P.J21[4 * i] = __fdiv_rn(dvy, det);
      P.J21[4 * i + 1] = __fdiv_rn(-dvx, det);
      P.J21[4 * i + 2] = __fdiv_rn(-duy, det);
      P.J21[4 * i + 3] = __fdiv_rn(dux, det);
#else
      const float p1 = dux * dvy, p2 = dvx * duy, det = p1 - p2;
      P.J21[4 * i] = dvy / det; P.J21[4 * i + 1] = -dvx / det; P.J21[4 * i + 2] = -duy / det; P.J21[4 * i + 3] = dux / det;
#endif
    }
    if (P.H12) {
      P.H12[6 * i] = (float)(d[6] * (fu * fu));
      P.H12[6 * i + 1] = (float)(d[7] * (fu * fu));
      P.H12[6 * i + 2] = (float)(d[8] * (fu * fv));
      P.H12[6 * i + 3] = (float)(d[9] * (fu * fv));
      P.H12[6 * i + 4] = (float)(d[10] * (fv * fv));
      P.H12[6 * i + 5] = (float)(d[11] * (fv * fv));
    }
  }
}

/* status written to scalars[4] */
enum { SCHWARP_OK = 0, SCHWARP_OUT_OF_DOMAIN = 1, SCHWARP_INIT_FAILED = 2 };

/* One keyframe pair, start to finish.  sh: schwarp_smem(...).total doubles of shared memory. */
DS_FN_NOINLINE void schwarp_fit_one(const Team team, const SchwarpProb &P, const SchwarpWs &ws, double *sh) {
  const BbsView &s = P.bbs;
  const int nu = s.nptsu, nv = s.nptsv, NC = nu * nv, NP = 2 * NC;
  const SchwarpSmem m = schwarp_smem(nu, nv);
  double *x = sh + m.x, *xc = sh + m.xc, *g = sh + m.g, *scale = sh + m.scale, *add = sh + m.add,
         *step = sh + m.step;
  const SiteMap sm = site_map(sh + m.ints, nu, nv);
  const BandDims b1{nv, nu, 3}, b2{2 * nv, nu, 3};
  fill_cell_integrals(team, sh + m.ci);
  /* grid-site intervals and, per control index, the range of sites whose window covers it */
  if (team.tid == 0) {
    double b[3][4];
    for (int i = 0; i < nu; i++) site_axis(s.umin, s.umax, nu, i, sm.Su[i], b);
    for (int j = 0; j < nv; j++) site_axis(s.vmin, s.vmax, nv, j, sm.Sv[j], b);
    for (int p = 0; p < nu; p++) {
      int lo = nu, hi = -1;
      for (int i = 0; i < nu; i++)
        if (sm.Su[i] <= p && p <= sm.Su[i] + 3) { if (i < lo) lo = i; hi = i; }
      sm.lou[p] = lo; sm.hiu[p] = hi;
    }
    for (int p = 0; p < nv; p++) {
      int lo = nv, hi = -1;
      for (int j = 0; j < nv; j++)
        if (sm.Sv[j] <= p && p <= sm.Sv[j] + 3) { if (j < lo) lo = j; hi = j; }
      sm.lov[p] = lo; sm.hiv[p] = hi;
    }
  }
  DS_FOR(i, NP) x[i] = P.x[i];
  team.sync();
#if defined(DS_NRSFM_PROF) && DS_CUDA
  long long pfp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pft = clock64();
#define PF_LAP(i) do { const long long n_ = clock64(); pfp[i] += n_ - pft; pft = n_; } while (0)
#else
#define PF_LAP(i) do {} while (0)
#endif
  CellSort cs{P.n, nu - 3, nv - 3, ws.cell, ws.cstart, ws.perm, ws.taps};
  const int bad = cell_sort(team, s, P.kp1, cs, sh + m.red);
  if (bad) {
    if (team.tid == 0) { P.scalars[0] = P.scalars[1] = 0.0; P.scalars[2] = P.scalars[3] = 0.0; P.scalars[4] = SCHWARP_OUT_OF_DOMAIN; }
    return;
  }
  PF_LAP(0);
  /* C'C in band form (bs = nptsv).  Per knot cell the 16 x 16 Gram matrix of its matches' tap products first
   * (fixed order over the cell's matches; scratch: the factor workspace, not in use yet), then every band entry
   * sums the <= 16 cells that cover both control points. */
  {
    const int ncv = nv - 3, ncells = (nu - 3) * ncv;
    double *G = ws.Lb;
    DS_FOR(item, ncells * 256) {
      const int c = item >> 8, ab = item & 255, ta = ab >> 4, tb = ab & 15;
      double acc = 0.0;
      for (int q = ws.cstart[c]; q < ws.cstart[c + 1]; q++) {
        const double *tp = ws.taps + 8 * (size_t)ws.perm[q];
        acc += tp[ta >> 2] * tp[4 + (ta & 3)] * tp[tb >> 2] * tp[4 + (tb & 3)];
      }
      G[item] = acc;
    }
    team.sync();
    DS_FOR(idx, nu * 4 * nv * nv) {
      const int qv = idx % nv, pv = (idx / nv) % nv, d = (idx / (nv * nv)) & 3, pu = idx / (4 * nv * nv), qu = pu - d;
      double acc = 0.0;
      const int dv = pv > qv ? pv - qv : qv - pv;
      if (qu >= 0 && dv <= 3) {
        const int u0 = pu - 3 > 0 ? pu - 3 : 0, u1 = qu < nu - 4 ? qu : nu - 4;
        const int vhi = pv < qv ? pv : qv, vlo = (pv > qv ? pv : qv) - 3;
        for (int Iu = u0; Iu <= u1; Iu++)
          for (int Iv = (vlo > 0 ? vlo : 0); Iv <= vhi && Iv <= nv - 4; Iv++)
            acc += G[(size_t)(Iu * ncv + Iv) * 256 + ((pu - Iu) * 4 + (pv - Iv)) * 16 + (qu - Iu) * 4 + (qv - Iv)];
      }
      ws.CtC[idx] = acc;
    }
  }
  team.sync();
  PF_LAP(1);
  int status = SCHWARP_OK;
  if (P.initialize) {
    /* Warp::initialize (Schwarp.cc:99-160): (C'C + lambda B) x0 = C' q2, two right-hand sides.
     * The system is assembled into Hb (reused later) with bs = nptsv. */
    const BendCoef bc = bending_coef(s);
    DS_FOR(idx, nu * 4 * nv * nv) {
      const int qv = idx % nv, pv = (idx / nv) % nv, d = (idx / (nv * nv)) & 3, pu = idx / (4 * nv * nv), qu = pu - d;
      ws.Hb[idx] = qu >= 0 ? ws.CtC[idx] + P.lambda * bending_entry_tab(s, bc, sh + m.ci, pu, pv, qu, qv) : 0.0;
    }
    DS_FOR(p, NC) {
      const int pu = p / nv, pv = p - pu * nv;
      double ax = 0.0, ay = 0.0;
      for (int Iu = (pu - 3 > 0 ? pu - 3 : 0); Iu <= pu && Iu <= nu - 4; Iu++)
        for (int Iv = (pv - 3 > 0 ? pv - 3 : 0); Iv <= pv && Iv <= nv - 4; Iv++) {
          const int c = Iu * (nv - 3) + Iv;
          for (int q = ws.cstart[c]; q < ws.cstart[c + 1]; q++) {
            const int mi = ws.perm[q];
            const double *tp = ws.taps + 8 * (size_t)mi;
            const double w = tp[pu - Iu] * tp[4 + pv - Iv];
            ax += w * (double)P.kp2[2 * mi];
            ay += w * (double)P.kp2[2 * mi + 1];
          }
        }
      step[p] = ax;
      step[NC + p] = ay;
    }
    team.sync();
#if DS_CUDA
    if (m.p1.nt > 0) band_to_rows(team, b1, m.p1, ws.Hb, nullptr, ws.Lb);
#endif
    const bool ok = bband_solve_any(team, b1, m.p1, ws.Hb, ws.Lb, nullptr, nullptr, sh + m.solver, step, 2, NC);
    team.sync();
    if (ok) {
      DS_FOR(i, NP) x[i] = step[i];
    } else {
      status = SCHWARP_INIT_FAILED;
    }
    team.sync();
  }
  PF_LAP(2);
#if defined(DS_NRSFM_PROF) && DS_CUDA
  long long pf_solve = 0;
  const long long pf_lm0 = clock64();
#endif
  /* ---- Ceres-style Levenberg-Marquardt */
  double radius = LM_INITIAL_RADIUS, decrease = 2.0, cost = 0.0, rho1 = 1.0, cost_initial = 0.0;
  int iters = 0, accepted = 0, invalid = 0;
  bool have_scale = false, need_eval = true;
  while (status == SCHWARP_OK) {
    if (need_eval) {
      PF_LAP(7);
      cost = schwarp_eval(team, P, ws, sh, m, x, true, &rho1);
      if (iters == 0) cost_initial = cost;
      team.sync();
      PF_LAP(3);
      schwarp_site_jacobians(team, P, ws, sh, m);
      PF_LAP(4);
      schwarp_build(team, P, ws, sh, m, rho1);
      PF_LAP(5);
      double gmax = 0.0;
      DS_FOR(i, NP) gmax = fmax(gmax, fabs(g[i]));
      gmax = team_max(team, gmax, sh + m.red);
      if (!have_scale) {
        DS_FOR(i, NP) scale[i] = 1.0 / (1.0 + sqrt(band_at(ws.Hb, b2, i, i)));
        have_scale = true;
        team.sync();
      }
#if DS_CUDA
      if (m.p2.nt > 0) band_to_rows(team, b2, m.p2, ws.Hb, scale, ws.Lb);
#endif
      need_eval = false;
      if (gmax <= 1e-10) break;
    }
    if (iters >= P.max_iterations) break;
    if (radius < LM_MIN_RADIUS) break;
    iters++;
    /* (S H S + D'D) y = S g ; step = -y */
    DS_FOR(i, NP) {
      const double h = band_at(ws.Hb, b2, i, i) * scale[i] * scale[i];
      add[i] = clampd(h, LM_MIN_DIAG, LM_MAX_DIAG) / radius;
      step[i] = g[i] * scale[i];
    }
    team.sync();
#if defined(DS_NRSFM_PROF) && DS_CUDA
    const long long pf0 = clock64();
#endif
    bool valid = bband_solve_any(team, b2, m.p2, ws.Hb, ws.Lb, scale, add, sh + m.solver, step, 1, NP);
    team.sync();
#if defined(DS_NRSFM_PROF) && DS_CUDA
    pf_solve += clock64() - pf0;
#endif
    double model_change = 0.0;
    if (valid) {
      /* -(J s).(r + J s/2) with s = -y: (y'Sg + y'D y)/2 from the solved system */
      double a = 0.0, fin = 0.0;
      DS_FOR(i, NP) {
        const double y = step[i];
        a += y * (g[i] * scale[i]) + y * y * add[i];
        if (!(fabs(y) < DBL_MAX)) fin += 1.0;
      }
      a = team_sum(team, a, sh + m.red);
      fin = team_sum(team, fin, sh + m.red);
      model_change = 0.5 * a;
      valid = fin == 0.0 && model_change > 0.0;
    }
    if (!valid) {
      if (++invalid >= LM_MAX_INVALID) break;
      radius /= decrease;
      decrease *= 2.0;
      continue;
    }
    invalid = 0;
    /* candidate: x + S * step (block order -> [all x; all y]) */
    double xn = 0.0, sn = 0.0;
    DS_FOR(Pi, NP) {
      const int pu = Pi / (2 * nv), rem = Pi - pu * 2 * nv, pv = rem >> 1, co = rem & 1;
      const int l = co * NC + pu * nv + pv;
      const double dlt = -step[Pi] * scale[Pi];
      xc[l] = x[l] + dlt;
      xn += x[l] * x[l];
      sn += dlt * dlt;
    }
    xn = team_sum(team, xn, sh + m.red);
    sn = team_sum(team, sn, sh + m.red);
    double rho1c;
    double cost_c = schwarp_eval(team, P, ws, sh, m, xc, false, &rho1c);
    if (!(cost_c < DBL_MAX)) cost_c = DBL_MAX;
    if (sqrt(sn) <= 1e-8 * (sqrt(xn) + 1e-8)) break;
    if (fabs(cost - cost_c) <= 1e-6 * cost) break;
    const double rel = (cost - cost_c) / model_change;
    if (rel > LM_MIN_REL_DECREASE) {
      team.sync();
      DS_FOR(i, NP) x[i] = xc[i];
      accepted++;
      const double t = 2.0 * rel - 1.0;
      radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
      radius = fmin(LM_MAX_RADIUS, radius);
      decrease = 2.0;
      need_eval = true;
      team.sync();
    } else {
      radius /= decrease;
      decrease *= 2.0;
    }
  }
  team.sync();
  PF_LAP(7);
  if (status == SCHWARP_OK) {
    DS_FOR(i, NP) P.x[i] = x[i];
    schwarp_diffprop(team, P, ws, x);
  }
  PF_LAP(6);
#if defined(DS_NRSFM_PROF) && DS_CUDA
  if (team.tid == 0 && blockIdx.x == 0)
    printf("[nrsfm prof] fit: cell sort %lld, C'C %lld, initialisation %lld, eval %lld, site Jacobians %lld, build %lld, DiffProp %lld, LM loop rest (solves, candidate evals) %lld\n",
           pfp[0], pfp[1], pfp[2], pfp[3], pfp[4], pfp[5], pfp[6], pfp[7]);
#endif
  if (team.tid == 0) {
    P.scalars[0] = cost_initial;
    P.scalars[1] = cost;
    P.scalars[2] = iters;
    P.scalars[3] = accepted;
    P.scalars[4] = status;
#if defined(DS_NRSFM_PROF) && DS_CUDA
    P.scalars[0] = (double)pf_solve;              /* cycles inside the LM solves */
    P.scalars[1] = (double)(clock64() - pf_lm0);  /* cycles of the LM loop + DiffProp */
#endif
  }
}

/* Residuals and dense Jacobian at P.x (parity hook; thread-per-row, no shared state).
 * r: [2n+4NC], J: [(2n+4NC) x 2NC] row-major or null.  Uses the same device arithmetic
 * as the fit (bbs taps, site derivatives). */
DS_FN void schwarp_row(const SchwarpProb &P, int row, double *r, double *J) {
  const BbsView &s = P.bbs;
  const int NC = s.nptsu * s.nptsv, NP = 2 * NC, n = P.n;
  const double *xs = P.x;
  if (row < 2 * n) {
    const int i = row < n ? row : row - n, co = row < n ? 0 : 1;
    double nu, nv, bu[4], bv[4];
    int Iu, Iv;
    bbs_normalize(s.umin, s.umax, s.nptsu, (double)P.kp1[2 * i], nu, Iu);
    bbs_normalize(s.vmin, s.vmax, s.nptsv, (double)P.kp1[2 * i + 1], nv, Iv);
    if (!bbs_in_domain(s, Iu, Iv)) { r[row] = NAN; return; }
    bbs_basis(0, nu, bu);
    bbs_basis(0, nv, bv);
    double w = 0.0;
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++) {
        const int l = (Iu + a) * s.nptsv + Iv + b;
        w += xs[co * NC + l] * (bu[a] * bv[b]);
        if (J) J[(size_t)row * NP + l] = -(bu[a] * bv[b]) * P.fx; /* both rows: x columns, fx (quirk C6) */
      }
    r[row] = (double)P.isig[i] * ((double)P.kp2[2 * i + co] - w) * (co ? P.fy : P.fx);
    return;
  }
  const int q = (row - 2 * n) / NC, k = (row - 2 * n) - q * NC;
  const int i = k / s.nptsv, j = k - i * s.nptsv;
  int Iu, Iv;
  double bu[3][4], bv[3][4];
  site_axis(s.umin, s.umax, s.nptsu, i, Iu, bu);
  site_axis(s.vmin, s.vmax, s.nptsv, j, Iv, bv);
  const double fu = bbs_deriv_fact(s, 1, 0), fv = bbs_deriv_fact(s, 0, 1), lam = P.lambda;
  double d[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      const int l = (Iu + a) * s.nptsv + Iv + b;
      const double cx = xs[l], cy = xs[NC + l];
      const double w10 = fu * bu[1][a] * bv[0][b], w01 = fv * bu[0][a] * bv[1][b];
      const double w20 = fu * fu * bu[2][a] * bv[0][b], w02 = fv * fv * bu[0][a] * bv[2][b],
                   w11 = fu * fv * bu[1][a] * bv[1][b];
      d[0] += cx * w10; d[1] += cy * w10; d[2] += cx * w01; d[3] += cy * w01;
      d[4] += cx * w20; d[5] += cy * w20; d[6] += cx * w02; d[7] += cy * w02; d[8] += cx * w11; d[9] += cy * w11;
    }
  const double xu = d[0], yu = d[1], xv = d[2], yv = d[3], xuu = d[4], yuu = d[5], xvv = d[6], yvv = d[7], xuv = d[8],
               yuv = d[9];
  double rv;
  if (q == 0) rv = xuu * yu - yuu * xu;
  else if (q == 1) rv = yvv * xv - xvv * yv;
  else if (q == 2) rv = xuu * yv - yuu * xv + 2 * (xuv * yu - yuv * xu);
  else rv = yvv * xu - xvv * yu + 2 * (yuv * xv - xuv * yv);
  r[row] = rv * lam;
  if (!J) return;
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      const int l = (Iu + a) * s.nptsv + Iv + b;
      const double cu = fu * bu[1][a] * bv[0][b], cv = fv * bu[0][a] * bv[1][b];
      const double cuu = fu * fu * bu[2][a] * bv[0][b], cvv = fv * fv * bu[0][a] * bv[2][b],
                   cuv = fu * fv * bu[1][a] * bv[1][b];
      double jx, jy;
      if (q == 0) { jx = yu * cuu - yuu * cu; jy = xuu * cu - xu * cuu; }
      else if (q == 1) { jx = yvv * cv - yv * cvv; jy = xv * cvv - xvv * cv; }
      else if (q == 2) { jx = yv * cuu - yuu * cv + 2 * yu * cuv - 2 * yuv * cu; jy = xuu * cv - xv * cuu + 2 * xuv * cu - 2 * xu * cuv; }
      else { jx = yvv * cu - yu * cvv - 2 * yv * cuv + 2 * yuv * cv; jy = xu * cvv - xvv * cu - 2 * xuv * cv + 2 * xv * cuv; }
      J[(size_t)row * NP + l] = lam * jx;
      J[(size_t)row * NP + NC + l] = lam * jy;
    }
}

/* ===================================================================== *
 *  Isometric normals
 * ===================================================================== */
struct NormalsProb {
  int n_points, npairs;
  const int *pair_ptr;
  const float *J12, *J21, *H12, *I1, *I2, *k_first, *ref_uv;
  const uint8_t *from_ref;
  const double *k_init;
  int max_iterations, corrected_t2;
  double *Q; /* scratch [20][npairs] (coefficient-major) */
  double *k_out, *cov_out;
  float *normal_out, *pair_normal_out;
  uint8_t *status_out, *pair_valid_out;
  int *iters_out;
};

#if DS_CUDA
#define DS_F32(expr_cuda, expr_host) (expr_cuda)
#else
#define DS_F32(expr_cuda, expr_host) (expr_host)
#endif
DS_FN float f32_mul(float a, float b) { return DS_F32(__fmul_rn(a, b), a * b); }
DS_FN float f32_add(float a, float b) { return DS_F32(__fadd_rn(a, b), a + b); }
DS_FN float f32_sub(float a, float b) { return DS_F32(__fsub_rn(a, b), a - b); }
DS_FN float f32_div(float a, float b) { return DS_F32(__fdiv_rn(a, b), a / b); }

/* both cubic polynomials of one pair (PolySolver.cc:50-149), coefficient order
 * [x^3, x^2 y, x y^2, y^3, x^2, x y, y^2, x, y, 1]; the inputs t1,t2,e1,e2 are formed in
 * fp32 like NormalEstimator.cc:88-103 */
DS_FN void pair_polynomials(const float *J12, const float *H12, const float *I1, const float *I2, int corrected_t2,
                            double *q1, double *q2, int stride) {
  const float fa = J12[0], fb = J12[1], fc = J12[2], fd = J12[3];
  const float t1f = f32_add(f32_div(f32_mul(-fb, H12[4]), 2.f), f32_div(f32_mul(fa, H12[5]), 2.f));
  const float t2f = corrected_t2 ? f32_sub(f32_div(f32_mul(fd, H12[0]), 2.f), f32_div(f32_mul(fc, H12[1]), 2.f))
                                 : f32_add(f32_div(-f32_mul(fd, H12[4]), 2.f), f32_div(f32_mul(fc, H12[5]), 2.f));
  const float e1f = f32_add(f32_add(1.f, f32_mul(I1[0], I1[0])), f32_mul(I1[1], I1[1]));
  const float e2f = f32_add(f32_add(1.f, f32_mul(I2[0], I2[0])), f32_mul(I2[1], I2[1]));
  const double a = fa, b = fb, c = fc, d = fd, t1 = t1f, t2 = t2f, e1 = e1f, e2 = e2f;
  const double x1 = I1[0], y1 = I1[1], x2 = I2[0], y2 = I2[1];
  const double D = a * d - c * b, D2 = D * D;
  const double Pq = a * x2 + b * y2, Qq = c * x2 + d * y2, mm = a * c + b * d;
  const double na = a * a + b * b, nc = c * c + d * d;
  const double w = a * x2 * y1 - c * x1 * x2 + b * y1 * y2 - d * x1 * y2;
  const double ee = e1 * e2;
  q1[0 * stride] = D * (t1 * ee - D * (e1 * Qq - y1 * e2));
  q1[1 * stride] = -D * (t2 * ee - D * (e1 * Pq - x1 * e2));
  q1[2 * stride] = 0.0;
  q1[3 * stride] = 0.0;
  q1[4 * stride] = t2 * (ee * t1 - D * (e1 * Qq - 2 * e2 * y1)) - t1 * D * (e1 * Pq + 2 * e2 * x1) + D2 * (e1 * mm - 2 * w);
  q1[5 * stride] = e1 * (-e2 * t2 * t2 + 2 * t2 * D * Pq - na * D2) + e2 * D2;
  q1[6 * stride] = 0.0;
  q1[7 * stride] = t1 * (e2 * D + 2 * x1 * D * Pq) - 2 * t2 * (e2 * x1 * t1 + D * w) + e2 * y1 * t2 * t2 +
                   D2 * (-2 * x1 * mm + y1 * na - Qq);
  q1[8 * stride] = t2 * D * (e2 - 2 * x1 * Pq) + x1 * e2 * t2 * t2 + D2 * (x1 * na - Pq);
  const double c00 = t2 * (e2 * t1 - D * Qq) - t1 * D * Pq + mm * D2;
  q1[9 * stride] = c00;
  q2[0 * stride] = 0.0;
  q2[1 * stride] = 0.0;
  q2[2 * stride] = -D * (ee * t1 - D * (e1 * Qq - e2 * y1));
  q2[3 * stride] = D * (ee * t2 - D * (e1 * Pq - e2 * x1));
  q2[4 * stride] = 0.0;
  q2[5 * stride] = e1 * (-e2 * t1 * t1 + D * (2 * t1 * Qq - nc * D)) + e2 * D2;
  q2[6 * stride] = t2 * (ee * t1 - D * (e1 * Qq + 2 * e2 * y1)) - t1 * D * (e1 * Pq - 2 * e2 * x1) + D2 * (e1 * mm + 2 * w);
  q2[7 * stride] = t1 * D * (e2 - 2 * y1 * Qq) + y1 * (e2 * t1 * t1 + D2 * nc) - D2 * Qq;
  q2[8 * stride] = t2 * (e2 * D + 2 * y1 * D * Qq) + t1 * (-2 * e2 * y1 * t2 + 2 * D * w) + e2 * x1 * t1 * t1 -
                   2 * D2 * (mm * y1 + 0.5 * Pq - 0.5 * nc * x1);
  q2[9 * stride] = c00;
}

/* cost 0.5*sum e^2 over the pairs [j0, j1) that enter the system; optionally g = J'e, H = J'J */
DS_FN double normals_cost(const NormalsProb &P, int j0, int j1, double x, double y, double *g, double *H) {
  double c = 0.0;
  if (g) { g[0] = g[1] = 0.0; H[0] = H[1] = H[2] = 0.0; }
  const int st = P.npairs;
  const double xx = x * x, xy = x * y, yy = y * y;
  for (int j = j0; j < j1; j++) {
    if (P.from_ref && !P.from_ref[j]) continue;
    for (int k = 0; k < 2; k++) {
      const double *q = P.Q + (size_t)(10 * k) * st + j;
      const double q0 = q[0], q1 = q[st], q2 = q[2 * (size_t)st], q3 = q[3 * (size_t)st], q4 = q[4 * (size_t)st],
                   q5 = q[5 * (size_t)st], q6 = q[6 * (size_t)st], q7 = q[7 * (size_t)st], q8 = q[8 * (size_t)st],
                   q9 = q[9 * (size_t)st];
      const double e = q0 * xx * x + q1 * xx * y + q2 * x * yy + q3 * yy * y + q4 * xx + q5 * xy + q6 * yy + q7 * x +
                       q8 * y + q9;
      c += e * e;
      if (g) {
        const double jx = 3 * q0 * xx + 2 * q1 * xy + q2 * yy + 2 * q4 * x + q5 * y + q7;
        const double jy = q1 * xx + 2 * q2 * xy + 3 * q3 * yy + q5 * x + 2 * q6 * y + q8;
        g[0] += jx * e; g[1] += jy * e;
        H[0] += jx * jx; H[1] += jx * jy; H[2] += jy * jy;
      }
    }
  }
  return 0.5 * c;
}

/* One map point: polynomial system, Ceres-style LM with the options of
 * NormalEstimator.cc:137-149, covariance, normal, transfer along the pairs (:176-223). */
DS_FN void normals_point(const NormalsProb &P, int i) {
  const int j0 = P.pair_ptr[i], j1 = P.pair_ptr[i + 1];
  int np = 0;
  for (int j = j0; j < j1; j++) {
    if (P.pair_valid_out) P.pair_valid_out[j] = 0;
    if (P.from_ref && !P.from_ref[j]) continue;
    pair_polynomials(P.J12 + 4 * j, P.H12 + 6 * j, P.I1 + 2 * j, P.I2 + 2 * j, P.corrected_t2, P.Q + j,
                     P.Q + (size_t)10 * P.npairs + j, P.npairs);
    np++;
  }
  double x[2] = {P.k_init ? P.k_init[2 * i] : 0.0, P.k_init ? P.k_init[2 * i + 1] : 0.0};
  int status = 0, iters = 0;
  if (np > 0) {
    double g[2], H[3], scale[2] = {1, 1}, radius = LM_INITIAL_RADIUS, decrease = 2.0, cost = 0.0;
    bool have_scale = false, need_eval = true;
    int invalid = 0;
    for (;;) {
      if (need_eval) {
        cost = normals_cost(P, j0, j1, x[0], x[1], g, H);
        if (!have_scale) {
          scale[0] = 1.0 / (1.0 + sqrt(H[0]));
          scale[1] = 1.0 / (1.0 + sqrt(H[2]));
          have_scale = true;
        }
        need_eval = false;
        if (fmax(fabs(g[0]), fabs(g[1])) <= 1e-8) break;
      }
      if (iters >= P.max_iterations) break;
      if (radius < LM_MIN_RADIUS) break;
      iters++;
      const double h00 = H[0] * scale[0] * scale[0], h01 = H[1] * scale[0] * scale[1], h11 = H[2] * scale[1] * scale[1];
      const double g0 = g[0] * scale[0], g1 = g[1] * scale[1];
      const double a00 = h00 + clampd(h00, LM_MIN_DIAG, LM_MAX_DIAG) / radius;
      const double a11 = h11 + clampd(h11, LM_MIN_DIAG, LM_MAX_DIAG) / radius;
      bool valid = a00 > 0.0;
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
          valid = fabs(s0) < DBL_MAX && fabs(s1) < DBL_MAX;
        }
      }
      if (valid) {
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
      const double xc0 = x[0] + d0, xc1 = x[1] + d1;
      double cost_c = normals_cost(P, j0, j1, xc0, xc1, nullptr, nullptr);
      if (!(cost_c < DBL_MAX)) cost_c = DBL_MAX;
      if (sqrt(d0 * d0 + d1 * d1) <= 1e-8 * (sqrt(x[0] * x[0] + x[1] * x[1]) + 1e-8)) break;
      if (fabs(cost - cost_c) <= 1e-10 * cost) break;
      const double rel = (cost - cost_c) / model_change;
      if (rel > LM_MIN_REL_DECREASE) {
        x[0] = xc0; x[1] = xc1;
        const double t = 2.0 * rel - 1.0;
        radius = fmin(LM_MAX_RADIUS, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
        decrease = 2.0;
        need_eval = true;
      } else {
        radius /= decrease; decrease *= 2.0;
      }
    }
    /* ceres::Covariance: (J'J)^-1, rank deficient below a reciprocal condition number of 1e-14 */
    normals_cost(P, j0, j1, x[0], x[1], g, H);
    const double tr = H[0] + H[2], det = H[0] * H[2] - H[1] * H[1];
    const double disc = sqrt(fmax(0.0, 0.25 * tr * tr - det));
    const double lmax = 0.5 * tr + disc, lmin = det / (lmax > 0 ? lmax : 1.0);
    if (!(lmax > 0.0) || !(lmax < DBL_MAX) || !(lmin / lmax >= 1e-14)) {
      status = 2;
    } else {
      status = 1;
      if (P.cov_out) {
        P.cov_out[4 * i] = H[2] / det; P.cov_out[4 * i + 1] = -H[1] / det;
        P.cov_out[4 * i + 2] = -H[1] / det; P.cov_out[4 * i + 3] = H[0] / det;
      }
      if (P.normal_out) {
        const float u = P.ref_uv[2 * i], v = P.ref_uv[2 * i + 1];
        P.normal_out[3 * i] = (float)x[0];
        P.normal_out[3 * i + 1] = (float)x[1];
        P.normal_out[3 * i + 2] = (float)(1 - x[0] * u - x[1] * v);
      }
    }
  }
  if (P.k_out) { P.k_out[2 * i] = x[0]; P.k_out[2 * i + 1] = x[1]; }
  if (P.status_out) P.status_out[i] = (uint8_t)status;
  if (P.iters_out) P.iters_out[i] = iters;
  if (status == 2) return;
  for (int j = j0; j < j1; j++) {
    double n0, n1;
    const int from_ref = P.from_ref ? P.from_ref[j] : 1;
    if (from_ref) {
      if (status != 1) continue;
      n0 = x[0]; n1 = x[1];
    } else {
      if (!P.k_first) continue;
      const float f0 = P.k_first[2 * j], f1 = P.k_first[2 * j + 1];
      if (f0 != f0 || f1 != f1) continue;
      n0 = f0; n1 = f1;
    }
    const float *Jf = P.J12 + 4 * j, *Ji = P.J21 + 4 * j, *Hh = P.H12 + 6 * j;
    const float a = Jf[0], b = Jf[1], c = Jf[2], d = Jf[3];
    const float det = f32_sub(f32_mul(a, d), f32_mul(c, b));
    const float t1 = f32_add(f32_div(f32_mul(-b, Hh[4]), 2.f), f32_div(f32_mul(a, Hh[5]), 2.f));
    const float t2 = f32_sub(f32_div(f32_mul(d, Hh[0]), 2.f), f32_div(f32_mul(c, Hh[1]), 2.f));
    const float dd = f32_mul(det, det);
    const float corr1 = f32_div(f32_sub(f32_mul(d, t2), f32_mul(b, t1)), dd);
    const float corr2 = f32_div(f32_sub(f32_mul(a, t1), f32_mul(c, t2)), dd);
    const double k1 = (double)Ji[0] * n0 + (double)Ji[2] * n1 + (double)corr1;
    const double k2 = (double)Ji[1] * n0 + (double)Ji[3] * n1 + (double)corr2;
    if (P.pair_normal_out) {
      P.pair_normal_out[3 * j] = (float)k1;
      P.pair_normal_out[3 * j + 1] = (float)k2;
      P.pair_normal_out[3 * j + 2] = (float)(1 - k1 * (double)P.I2[2 * j] - k2 * (double)P.I2[2 * j + 1]);
    }
    if (P.pair_valid_out) P.pair_valid_out[j] = 1;
  }
}

/* ===================================================================== *
 *  Shape from normals
 * ===================================================================== */
struct SfnProb {
  BbsView bbs; /* valdim 1 */
  int n, n_eval;
  const float *uv, *normals, *eval_uv;
  double bending, mean_depth;
  double *ctrl_out;
  float *xyz_out;
  int *rc_out;
};

struct SfnWs {
  int *cell, *cstart, *perm;
  double *taps;  /* [n*8]   (cell_sort scratch)              */
  double *mrow;  /* [n][2][16] the two M rows of each normal */
  double *B;     /* [NC*NC] bending * B                      */
  double *N;     /* packed lower N = A'A when it does not fit in shared memory (else unused) */
  double *res;   /* [2n + NC + 1] residual of the stacked system */
  double *G;     /* [cells][16][16] per-cell Gram matrices of the M rows */
};

struct SfnSizes {
  size_t cell, cstart, perm, taps, mrow, B, N, res, G, total;
};
static inline
#if DS_CUDA
__host__ __device__
#endif
SfnSizes sfn_ws_sizes(int nptsu, int nptsv, int nmax) {
  SfnSizes z;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t NC = (size_t)nptsu * nptsv;
  size_t o = 0;
  z.cell = o; o += al(sizeof(int) * (nmax + 1));
  z.cstart = o; o += al(sizeof(int) * ((size_t)(nptsu - 3) * (nptsv - 3) + 1));
  z.perm = o; o += al(sizeof(int) * (nmax + 1));
  z.taps = o; o += al(sizeof(double) * 8 * (nmax + 1));
  z.mrow = o; o += al(sizeof(double) * 32 * (nmax + 1));
  z.B = o; o += al(sizeof(double) * NC * NC);
  z.N = o; o += al(sizeof(double) * NC * (NC + 1) / 2);
  z.res = o; o += al(sizeof(double) * (2 * (size_t)nmax + NC + 1));
  z.G = o; o += al(sizeof(double) * (size_t)(nptsu - 3) * (nptsv - 3) * 256);
  z.total = o;
  return z;
}

/* shared doubles: x, rhs, colbuf, ci, red, floats for the median (+ packed N if it fits) */
static DS_HD int sfn_smem_fixed(int NC) { return 4 * NC + 48 + 40 + (NC + 1) / 2 + 2; }

DS_FN size_t pk(int i, int c) { return (size_t)i * (i + 1) / 2 + c; }

/* in-place Cholesky of a packed lower-triangular SPD matrix; col: n doubles of shared memory */
DS_FN_NOINLINE bool packed_chol(const Team team, double *A, int n, double *col) {
  const int TX = DS_TX, tx = team.tid % TX, ty = team.tid / TX, NY = team.nthr / TX > 0 ? team.nthr / TX : 1;
  for (int j = 0; j < n; j++) {
    const double piv = A[pk(j, j)];
    if (!(piv > 0.0) || !(piv < DBL_MAX)) return false;
    const double isq = 1.0 / sqrt(piv);
    DS_FOR(t, n - j - 1) {
      const int i = j + 1 + t;
      const double v = A[pk(i, j)] * isq;
      col[i] = v;
      A[pk(i, j)] = v;
    }
    team.sync();
    if (team.tid == 0) A[pk(j, j)] = piv * isq; /* nobody reads the pivot during the update */
    for (int i = j + 1 + ty; i < n; i += NY) {
      const double a = col[i];
      double *ar = A + pk(i, 0);
      for (int c = j + 1 + tx; c <= i; c += TX) ar[c] -= a * col[c];
    }
    team.sync();
  }
  return true;
}

/* x <- (L L')^-1 x, x in shared memory */
DS_FN_NOINLINE void packed_solve(const Team team, const double *L, int n, double *x) {
  for (int j = 0; j < n; j++) {
    const double xj = x[j] / L[pk(j, j)];
    team.sync();
    if (team.tid == 0) x[j] = xj;
    DS_FOR(t, n - j - 1) {
      const int i = j + 1 + t;
      x[i] -= L[pk(i, j)] * xj;
    }
    team.sync();
  }
  for (int j = n - 1; j >= 0; j--) {
    const double xj = x[j] / L[pk(j, j)];
    team.sync();
    if (team.tid == 0) x[j] = xj;
    DS_FOR(i, j) x[i] -= L[pk(j, i)] * xj;
    team.sync();
  }
}

/* ---- the same factorisation on the FP64 tensor cores: N as 8x8 tiles of its lower triangle in shared memory, every
 * tile in operand-fragment order (element (r, c) at (c/4)*32 + r*4 + c%4: one m8n8k4 operand load is 32 consecutive
 * doubles, an accumulator pair one 16-byte access), right-looking by block columns: diagonal block by one warp (the
 * 8x8 Cholesky + inverse of ds_rowchol.h), panel X = C inv(L_kk)^T and trailing update C(i,j) -= X_i X_j^T by all
 * warps with two DMMAs per tile; three barriers per block column instead of two per column.  Yinv: the inverses of the
 * diagonal blocks (what the substitutions need).  n padded to 8 nt with identity rows. */
static DS_HD int sfn_tile_doubles(int NC) { const int nt = (NC + 7) / 8; return nt * (nt + 1) / 2 * 64 + nt * 64; }
DS_FN int tile_id(int i, int j) { return i * (i + 1) / 2 + j; }
DS_FN int tile_elem(int r, int c) { return (c >> 2) * 32 + r * 4 + (c & 3); }
#if DS_CUDA
DS_FN_NOINLINE bool sfn_tile_chol(const Team team, double *T, double *Yinv, int nt, int *flag) {
  const int warp = team.tid >> 5, lane = team.tid & 31, nwarp = team.nthr >> 5;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t T32 = smem_u32(T), Y32 = smem_u32(Yinv);
  const uint32_t pair_off = frag_pair_off(g, q), lane_off = 8u * (uint32_t)lane;
  if (team.tid == 0) *flag = 0;
  team.sync();
  for (int k = 0; k < nt; k++) {
    if (warp == 0) {
      const double *D = T + (size_t)tile_id(k, k) * 64;
      double a[36];
#pragma unroll
      for (int i = 0; i < NB; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) a[i * (i + 1) / 2 + j] = D[tile_elem(i, j)];
      const bool bad = chol8_regs(a);
      double col[NB];
      const int j = lane & 7;
      invcol8_regs(a, j, col);
      if (lane < NB) {
        double *Y = Yinv + (size_t)k * 64;
#pragma unroll
        for (int m = 0; m < NB; m++) Y[tile_elem(m, j)] = col[m];
      }
      if (bad && lane == 0) *flag = 1;
    }
    team.sync();
    if (*flag != 0) return false;
    const uint32_t yk = Y32 + 512u * (uint32_t)k;
    for (int i = k + 1 + warp; i < nt; i += nwarp) {
      const uint32_t tile = T32 + 512u * (uint32_t)tile_id(i, k);
      const double y0 = lds_f64(yk + lane_off), y1 = lds_f64(yk + 256u + lane_off);
      const double c0 = lds_f64(tile + lane_off), c1 = lds_f64(tile + 256u + lane_off);
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, c0, y0);
      dmma884(x0, x1, c1, y1);
      __syncwarp(); /* every lane has read C before X overwrites it */
      sts_v2f64(tile + pair_off, x0, x1);
    }
    team.sync();
    const int ntr = nt - 1 - k, total = ntr * (ntr + 1) / 2;
    for (int e = warp; e < total; e += nwarp) {
      int ii = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
      while ((ii + 1) * (ii + 2) / 2 <= e) ii++;
      while (ii * (ii + 1) / 2 > e) ii--;
      const int jj = e - ii * (ii + 1) / 2;
      const int i = k + 1 + ii, j = k + 1 + jj;
      const uint32_t tc = T32 + 512u * (uint32_t)tile_id(i, j) + pair_off;
      const uint32_t ta = T32 + 512u * (uint32_t)tile_id(i, k), tb = T32 + 512u * (uint32_t)tile_id(j, k);
      dbl2 acc = lds_v2f64(tc);
      const double a0 = -lds_f64(ta + lane_off), a1 = -lds_f64(ta + 256u + lane_off);
      const double b0 = lds_f64(tb + lane_off), b1 = lds_f64(tb + 256u + lane_off);
      dmma884(acc.x, acc.y, a0, b0);
      dmma884(acc.x, acc.y, a1, b1);
      sts_v2f64(tc, acc.x, acc.y);
    }
    team.sync();
  }
  return true;
}

/* x <- (L L')^-1 x on the tile-form factor; x: 8 nt doubles of shared memory.  One warp: the block recurrences have no
 * parallelism across block rows and a CTA-wide barrier per column is what the packed substitution spends its time on. */
DS_FN_NOINLINE void sfn_tile_solve(const Team team, const double *T, const double *Yinv, int nt, double *x) {
  const int warp = team.tid >> 5, lane = team.tid & 31;
  if (warp == 0) {
    const uint32_t T32 = smem_u32(T), Y32 = smem_u32(Yinv), X32 = smem_u32(x);
    const int m8 = lane & 7;
    for (int pass = 0; pass < 2; pass++) { /* 0: L y = r, top down; 1: L' x = y, bottom up */
      for (int kk = 0; kk < nt; kk++) {
        const int k = pass ? nt - 1 - kk : kk;
        /* the block's solution: lane m (mod 8) forms entry m of inv(L_kk) r_k (pass 0) or inv(L_kk)' r_k (pass 1) from
         * the stored inverse (zeros above its diagonal), two interleaved partial sums */
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < NB; j += 2) {
          const double ya = lds_f64(Y32 + 8u * (uint32_t)(k * 64 + (pass ? tile_elem(j, m8) : tile_elem(m8, j))));
          const double yb = lds_f64(Y32 + 8u * (uint32_t)(k * 64 + (pass ? tile_elem(j + 1, m8) : tile_elem(m8, j + 1))));
          s0 = fma(ya, lds_f64(X32 + 8u * (uint32_t)(NB * k + j)), s0);
          s1 = fma(yb, lds_f64(X32 + 8u * (uint32_t)(NB * k + j + 1)), s1);
        }
        const double mine = s0 + s1;
        double v[NB];
#pragma unroll
        for (int c = 0; c < NB; c++) v[c] = __shfl_sync(0xffffffffu, mine, c);
        __syncwarp();
        if (lane < NB) x[NB * k + lane] = mine;
        /* the other blocks: pass 0 rows below (tiles (I, k), entry r of block I), pass 1 columns above (tiles (k, J),
         * entry c of block J); eight entries per lane in flight */
        const int cnt = (pass ? k : nt - 1 - k) * NB;
        for (int base = 0; base < cnt; base += 256) {
          double acc[8];
          uint32_t xa[8], la[8];
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const int idx = base + lane + 32 * r;
            const int ok = idx < cnt, id = ok ? idx : 0;
            const int blk = pass ? (id >> 3) : k + 1 + (id >> 3), e = id & 7;
            xa[r] = ok ? X32 + 8u * (uint32_t)(NB * blk + e) : 0u;
            /* pass 0: L(blk, k)[e][c], c = 0..7; pass 1: L(k, blk)[r'][e], r' = 0..7 */
            la[r] = T32 + 512u * (uint32_t)(pass ? tile_id(k, blk) : tile_id(blk, k)) + 8u * (uint32_t)(pass ? tile_elem(0, e) : tile_elem(e, 0));
            acc[r] = ok ? lds_f64(xa[r]) : 0.0;
          }
#pragma unroll
          for (int c = 0; c < NB; c++) {
            /* element (e, c) sits at (c/4)*32 + e*4 + c%4, element (c, e) at (e/4)*32 + c*4 + e%4 */
            const uint32_t step = pass ? 8u * (uint32_t)(c * 4) : 8u * (uint32_t)((c >> 2) * 32 + (c & 3));
#pragma unroll
            for (int r = 0; r < 8; r++) acc[r] = fma(-lds_f64(la[r] + step), v[c], acc[r]);
          }
#pragma unroll
          for (int r = 0; r < 8; r++)
            if (xa[r]) asm volatile("st.shared.f64 [%0], %1;" ::"r"(xa[r]), "d"(acc[r]) : "memory");
        }
        __syncwarp();
      }
    }
  }
  team.sync();
}
#endif

/* the two M rows of normal i (ShapeFromNormals.cc:234-258) as 16 taps each; false outside the domain */
DS_FN bool sfn_rows(const BbsView &s, const float *uv, const float *nrm, int i, int &Iu, int &Iv, double *m1,
                    double *m2) {
  const double u = uv[2 * i], v = uv[2 * i + 1];
  double nu, nv, bu0[4], bv0[4], bu1[4], bv1[4];
  bbs_normalize(s.umin, s.umax, s.nptsu, u, nu, Iu);
  bbs_normalize(s.vmin, s.vmax, s.nptsv, v, nv, Iv);
  if (!bbs_in_domain(s, Iu, Iv)) return false;
  bbs_basis(0, nu, bu0); bbs_basis(0, nv, bv0); bbs_basis(1, nu, bu1); bbs_basis(1, nv, bv1);
  const double fu = bbs_deriv_fact(s, 1, 0), fv = bbs_deriv_fact(s, 0, 1);
  double nx = nrm[3 * i], ny = nrm[3 * i + 1], nz = nrm[3 * i + 2];
  const double nn = sqrt(nx * nx + ny * ny + nz * nz);
  nx /= nn; ny /= nn; nz /= nn;
  const double ne = nx * u + ny * v + nz;
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      const double c0 = bu0[a] * bv0[b], cu = fu * bu1[a] * bv0[b], cv = fv * bu0[a] * bv1[b];
      m1[a * 4 + b] = ne * cu + nx * c0;
      m2[a * 4 + b] = ne * cv + ny * c0;
    }
  return true;
}

/* One keyframe.  n_mode 0: packed N in the global workspace, 1: packed N in shared memory, 2 (device only): N as
 * 8x8 tiles in shared memory, factorised on the tensor cores.  sh: sfn_smem_fixed(NC) (+ NC(NC+1)/2 in mode 1,
 * + sfn_tile_doubles(NC) in mode 2) doubles.
 * rc: 0 ok, DEFSLAM_EBADARG site outside the domain, DEFSLAM_ENUMERIC not finite / not SPD. */
DS_FN_NOINLINE void sfn_solve_one(const Team team, const SfnProb &P, const SfnWs &ws, double *sh, int n_mode) {
  const BbsView &s = P.bbs;
  const int nu = s.nptsu, nv = s.nptsv, NC = nu * nv, n = P.n, ncv = nv - 3;
  double *x = sh, *rhs = sh + NC, *col = sh + 2 * NC, *acc = sh + 3 * NC, *ci = sh + 4 * NC, *red = ci + 48;
  float *fl = (float *)(red + 40);
  /* (the tiles are accessed 16 bytes at a time: their offset is rounded up to even, the host adds the slack) */
  double *N = n_mode == 2 ? (red + 40 + (((NC + 1) / 2 + 2 + 1) & ~1)) : (n_mode ? (red + 40 + (NC + 1) / 2 + 2) : ws.N);
  const bool tiles = n_mode == 2;
  const int nt = (NC + 7) / 8, Dp = NB * nt;
  double *Yinv = N + (size_t)nt * (nt + 1) / 2 * 64; /* (mode 2) */
  (void)Dp; (void)Yinv;
  fill_cell_integrals(team, ci);
  CellSort cs{n, nu - 3, nv - 3, ws.cell, ws.cstart, ws.perm, ws.taps};
  int bad = cell_sort(team, s, P.uv, cs, red);
  DS_FOR(i, n) {
    int Iu, Iv;
    sfn_rows(s, P.uv, P.normals, i, Iu, Iv, ws.mrow + 32 * (size_t)i, ws.mrow + 32 * (size_t)i + 16);
  }
  /* bending * B, dense (rows of the stacked system) */
  const BendCoef bcoef = bending_coef(s);
  DS_FOR(idx, NC * NC) {
    const int p = idx / NC, q = idx - p * NC;
    ws.B[idx] = P.bending * bending_entry_tab(s, bcoef, ci, p / nv, p % nv, q / nv, q % nv);
  }
  team.sync();
  if (bad) {
    if (team.tid == 0 && P.rc_out) *P.rc_out = DEFSLAM_EBADARG;
    return;
  }
  if (tiles) { /* padding rows: identity; the upper halves of the diagonal tiles mirror the lower ones */
    DS_FOR(idx, nt * (nt + 1) / 2 * 64) N[idx] = 0.0;
    team.sync();
    DS_FOR(i, Dp - NC) N[(size_t)tile_id(nt - 1, nt - 1) * 64 + tile_elem((NC + i) & 7, (NC + i) & 7)] = 1.0;
  }
  /* per knot cell the 16 x 16 Gram matrix of the M rows of its normals (fixed order over the cell's normals): an
   * entry of M'M then sums <= 16 cells instead of walking their normals */
  DS_FOR(item, (nu - 3) * ncv * 256) {
    const int c = item >> 8, tp = (item >> 4) & 15, tq = item & 15;
    double a = 0.0;
    for (int k = ws.cstart[c]; k < ws.cstart[c + 1]; k++) {
      const double *mr = ws.mrow + 32 * (size_t)ws.perm[k];
      a += mr[tp] * mr[tq] + mr[16 + tp] * mr[16 + tq];
    }
    ws.G[item] = a;
  }
  team.sync();
  /* N = M'M + B'B + 1 1' (packed lower) */
  DS_FOR(idx, NC * (NC + 1) / 2) {
    /* row from the packed index */
    int p = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((size_t)(p + 1) * (p + 2) / 2 <= (size_t)idx) p++;
    while ((size_t)p * (p + 1) / 2 > (size_t)idx) p--;
    const int q = idx - p * (p + 1) / 2;
    const int pu = p / nv, pv = p - pu * nv, qu = q / nv, qv = q - qu * nv;
    double a = 1.0;
    const int du = pu - qu, dv = pv > qv ? pv - qv : qv - pv; /* pu >= qu */
    if (du <= 3 && dv <= 3) {
      const int vhi = pv < qv ? pv : qv, vlo = (pv > qv ? pv : qv) - 3;
      for (int Iu = (pu - 3 > 0 ? pu - 3 : 0); Iu <= qu && Iu <= nu - 4; Iu++)
        for (int Iv = (vlo > 0 ? vlo : 0); Iv <= vhi && Iv <= nv - 4; Iv++) {
          const int c = Iu * ncv + Iv;
          const int tp = (pu - Iu) * 4 + pv - Iv, tq = (qu - Iu) * 4 + qv - Iv;
          a += ws.G[(size_t)c * 256 + tp * 16 + tq];
        }
    }
    if (du <= 6 && dv <= 6) {
      /* (B'B)[p,q] = sum_k B[k,p] B[k,q], k within 3 grid steps of both */
      const int ku0 = (pu - 3 > 0 ? pu - 3 : 0), ku1 = qu + 3 < nu - 1 ? qu + 3 : nu - 1;
      const int kv0 = ((pv > qv ? pv : qv) - 3 > 0 ? (pv > qv ? pv : qv) - 3 : 0);
      const int kv1 = (pv < qv ? pv : qv) + 3 < nv - 1 ? (pv < qv ? pv : qv) + 3 : nv - 1;
      double bb = 0.0;
      for (int ku = ku0; ku <= ku1; ku++)
        for (int kv = kv0; kv <= kv1; kv++) {
          const int k = ku * nv + kv;
          bb += ws.B[(size_t)p * NC + k] * ws.B[(size_t)q * NC + k]; /* B is symmetric: rows instead of columns (contiguous in k) */
        }
      a += bb;
    }
    if (tiles) {
      N[(size_t)tile_id(p >> 3, q >> 3) * 64 + tile_elem(p & 7, q & 7)] = a;
      if ((p >> 3) == (q >> 3) && p != q) N[(size_t)tile_id(p >> 3, q >> 3) * 64 + tile_elem(q & 7, p & 7)] = a;
    } else {
      N[idx] = a;
    }
  }
  DS_FOR(p, NC) x[p] = NC * P.mean_depth;
  team.sync();
  /* (mode 2) the substitutions work on a vector padded to 8 nt entries: col and acc together */
  double *work = col;
  int *tflag = (int *)(red + 36);
  (void)work; (void)tflag;
#if DS_CUDA
#define SFN_SOLVE(vec)                                            \
  do {                                                            \
    if (tiles) {                                                  \
      DS_FOR(i_, Dp) work[i_] = i_ < NC ? (vec)[i_] : 0.0;        \
      team.sync();                                                \
      sfn_tile_solve(team, N, Yinv, nt, work);                    \
      DS_FOR(i_, NC)(vec)[i_] = work[i_];                         \
      team.sync();                                                \
    } else {                                                      \
      packed_solve(team, N, NC, (vec));                           \
    }                                                             \
  } while (0)
  bool ok = tiles ? sfn_tile_chol(team, N, Yinv, nt, tflag) : packed_chol(team, N, NC, col);
#else
#define SFN_SOLVE(vec) packed_solve(team, N, NC, (vec))
  bool ok = packed_chol(team, N, NC, col);
#endif
  if (ok) {
    SFN_SOLVE(x);
    /* two sweeps of corrected semi-normal equations: r = b - A x, x += N^-1 A'r */
    for (int sweep = 0; sweep < 2; sweep++) {
      DS_FOR(i, n) {
        const int cell = ws.cell[i], Iu = cell / ncv, Iv = cell - Iu * ncv;
        const double *mr = ws.mrow + 32 * (size_t)i;
        double r1 = 0.0, r2 = 0.0;
        for (int a = 0; a < 4; a++)
          for (int b = 0; b < 4; b++) {
            const double xv = x[(Iu + a) * nv + Iv + b];
            r1 -= mr[a * 4 + b] * xv;
            r2 -= mr[16 + a * 4 + b] * xv;
          }
        ws.res[i] = r1;
        ws.res[n + i] = r2;
      }
      DS_FOR(k, NC) {
        double r = 0.0;
        const double *br = ws.B + (size_t)k * NC;
        const int ku = k / nv, kv = k - ku * nv;
        for (int pu = (ku - 3 > 0 ? ku - 3 : 0); pu <= ku + 3 && pu < nu; pu++)
          for (int pv = (kv - 3 > 0 ? kv - 3 : 0); pv <= kv + 3 && pv < nv; pv++) r -= br[pu * nv + pv] * x[pu * nv + pv];
        ws.res[2 * n + k] = r;
      }
      double sx = 0.0;
      DS_FOR(p, NC) sx += x[p];
      sx = team_sum(team, sx, red);
      const double rl = NC * P.mean_depth - sx;
      team.sync();
      DS_FOR(p, NC) {
        const int pu = p / nv, pv = p - pu * nv;
        double a = rl;
        for (int Iu = (pu - 3 > 0 ? pu - 3 : 0); Iu <= pu && Iu <= nu - 4; Iu++)
          for (int Iv = (pv - 3 > 0 ? pv - 3 : 0); Iv <= pv && Iv <= nv - 4; Iv++) {
            const int c = Iu * ncv + Iv, tp = (pu - Iu) * 4 + pv - Iv;
            for (int k = ws.cstart[c]; k < ws.cstart[c + 1]; k++) {
              const int mi = ws.perm[k];
              const double *mr = ws.mrow + 32 * (size_t)mi;
              a += mr[tp] * ws.res[mi] + mr[16 + tp] * ws.res[n + mi];
            }
          }
        for (int ku = (pu - 3 > 0 ? pu - 3 : 0); ku <= pu + 3 && ku < nu; ku++)
          for (int kv = (pv - 3 > 0 ? pv - 3 : 0); kv <= pv + 3 && kv < nv; kv++) {
            const int k = ku * nv + kv;
            a += ws.B[(size_t)k * NC + p] * ws.res[2 * n + k];
          }
        rhs[p] = a;
      }
      team.sync();
      SFN_SOLVE(rhs);
      DS_FOR(p, NC) x[p] += rhs[p];
      team.sync();
    }
  }
  int nonfinite = 0;
  DS_FOR(p, NC) nonfinite += !(fabs(x[p]) < DBL_MAX);
  nonfinite = team_sum_int(team, nonfinite, red);
  if (!ok || nonfinite) {
    if (team.tid == 0 && P.rc_out) *P.rc_out = DEFSLAM_ENUMERIC;
    return;
  }
  /* median rescale through fp32 (ShapeFromNormals.cc:128-142): element of rank NC/2 */
  DS_FOR(p, NC) fl[p] = (float)x[p];
  team.sync();
  DS_FOR(p, NC) {
    const float v = fl[p];
    int rank = 0;
    for (int q = 0; q < NC; q++) rank += (fl[q] < v) || (fl[q] == v && q < p);
    if (rank == NC / 2) acc[0] = (double)(1.f / v);
  }
  team.sync();
  const float corr = (float)acc[0];
  team.sync();
  DS_FOR(p, NC) {
    x[p] = corr * x[p];
    if (P.ctrl_out) P.ctrl_out[p] = x[p];
  }
  team.sync();
  DS_FOR(i, P.n_eval) {
    const double u = P.eval_uv[2 * i], v = P.eval_uv[2 * i + 1];
    double d;
    bbs_eval_site(s, x, u, v, 0, 0, &d);
    P.xyz_out[3 * i] = (float)(u * d);
    P.xyz_out[3 * i + 1] = (float)(v * d);
    P.xyz_out[3 * i + 2] = (float)d;
  }
  if (team.tid == 0 && P.rc_out) *P.rc_out = 0;
}

/* one row of the stacked system (parity hook) */
DS_FN void sfn_system_row(const SfnProb &P, const double *ci, int row, double *A, double *b) {
  const BbsView &s = P.bbs;
  const int NC = s.nptsu * s.nptsv, n = P.n;
  double *ar = A + (size_t)row * NC;
  for (int c = 0; c < NC; c++) ar[c] = 0.0;
  b[row] = 0.0;
  if (row < 2 * n) {
    const int i = row < n ? row : row - n;
    int Iu, Iv;
    double m1[16], m2[16];
    if (!sfn_rows(s, P.uv, P.normals, i, Iu, Iv, m1, m2)) return;
    for (int a = 0; a < 4; a++)
      for (int bb = 0; bb < 4; bb++) ar[(Iu + a) * s.nptsv + Iv + bb] = row < n ? m1[a * 4 + bb] : m2[a * 4 + bb];
  } else if (row < 2 * n + NC) {
    const int p = row - 2 * n;
    for (int q = 0; q < NC; q++)
      ar[q] = P.bending * bending_entry_tab(s, ci, p / s.nptsv, p % s.nptsv, q / s.nptsv, q % s.nptsv);
  } else {
    for (int c = 0; c < NC; c++) ar[c] = 1.0;
    b[row] = NC * P.mean_depth;
  }
}

}  // namespace ds
#endif
