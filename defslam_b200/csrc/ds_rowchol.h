/*
 * ds_rowchol.h -- banded Cholesky in "row owner" form on the FP64 tensor cores, shared by the SfT solve
 * (sft_rows.h: arrowhead system of the LM step) and the NRSfM stages (nrsfm_core.h: block-banded normal
 * matrix of the Schwarp fit).
 *
 * Left-looking by block rows of 8, one WARP per block row, the row's band (NT tiles of 8x8) held in DMMA
 * accumulator registers from the moment the matrix is read until the row is final; finished rows live in a
 * shared-memory ring in operand-fragment order; rows synchronise through progress counters; the dependency
 * chain (last panel tile, last update of the diagonal block, its Cholesky, the inverse of its factor) runs
 * inside one warp.  Up to 8 border rows (right-hand sides; for SfT also the camera border) ride along as a
 * further "row".  See sft_rows.h for the description of the method and DESIGN.md section 4 for the numbers.
 *
 * Device only (the CPU emulation tier has its own serial restatements next to each caller).
 */
#ifndef DS_ROWCHOL_H_
#define DS_ROWCHOL_H_
#include "ds_common.h"
#include "ds_tma.h"

namespace ds {

constexpr int LT_STRIDE = 64; /* doubles per 8x8 tile of the factor in global memory: tiles are contiguous (a padded
                                 stride of 72 kept the backward sweep's column reads of neighbouring tiles on different
                                 banks; the 11 % of factor traffic it cost weighs more) */

#if DS_CUDA
#ifndef DS_PROF_LOCALS /* per-phase cycle counters exist in the SfT profile build only */
#define DS_PROF_LOCALS(name, n) do {} while (0)
#define DS_PROF_LAP(name, i, t) do {} while (0)
#define DS_PROF_T0M(var) do {} while (0)
#define DS_PROF_FLUSH(name, n, idx, cond) do {} while (0)
#endif
#ifndef DS_DMMA884_DEFINED
#define DS_DMMA884_DEFINED
/* D(8x8) += A(8x4) B(4x8) on the FP64 tensor cores.  Lane T holds a = A[T/4][T%4],
 * b = B[T%4][T/4], and d0,d1 = D[T/4][2*(T%4)], D[T/4][2*(T%4)+1]. */
DS_FN void dmma884(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
#endif
DS_FN int ld_vol_s32(const int *p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
DS_FN dbl2 ldcg_dbl2(const dbl2 *p) {
  dbl2 v;
  asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
DS_FN void st_vol_s32(int *p, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
/* every lane polls (broadcast read); the data guarded by the counter is read after it in program order */
DS_FN void spin_ge(const int *p, int need) {
  while (ld_vol_s32(p) < need) {
  }
}
/* waits that are not on the critical path (border rows, producers of the backward sweep) back off between polls:
 * a spinning warp takes issue slots and shared-memory cycles from the SM's other CTA */
#ifndef DS_SPIN_SLEEP_NS
#define DS_SPIN_SLEEP_NS 0
#endif
DS_FN void spin_ge_relaxed(const int *p, int need) {
  while (ld_vol_s32(p) < need) {
#if DS_SPIN_SLEEP_NS > 0
    __nanosleep(DS_SPIN_SLEEP_NS);
#endif
  }
}
/* publish: the warp's shared-memory stores first, then the counter */
#ifndef DS_ROWS_FENCE_LIGHT
#define DS_ROWS_FENCE_LIGHT 0
#endif
#ifndef DS_ROWS_PREFETCH
#define DS_ROWS_PREFETCH 0
#endif
#ifndef DS_ROWS_BORDER_ON_CHAIN_SP
#define DS_ROWS_BORDER_ON_CHAIN_SP 0
#endif
#ifndef DS_ROWS_LOAD_CG
#define DS_ROWS_LOAD_CG 0
#endif
#ifndef DS_ROWS_SPLIT_CHAINS
#define DS_ROWS_SPLIT_CHAINS 0
#endif
DS_FN void publish(int *p, int v, int lane) {
  __syncwarp();
#if DS_ROWS_FENCE_LIGHT
  asm volatile("fence.acq_rel.cta;" ::: "memory");
#else
  __threadfence_block();
#endif
  if (lane == 0) st_vol_s32(p, v);
}
/* byte offset of this lane's accumulator pair inside a tile in fragment order */
DS_FN uint32_t frag_pair_off(int g, int q) { return 8u * (uint32_t)((q >> 1) * 32 + g * 4 + 2 * (q & 1)); }

struct RowShared {
  uint32_t ring;   /* shared address of the ring: slot (I % R), tile t at ((I % R) * NT + t) * 512 bytes */
  uint32_t ering;  /* border tiles, slot k % NT */
  double *dbuf;    /* [2][64] diagonal blocks on their way to the chain warp, row-major (= accumulator order) */
  int *prog, *ddone, *pre, *edone;
};

/* ------------------------------------------------------------------ chain warp (warp 0) */
/* branch-free reciprocal square root: hardware seed (MUFU.RSQ64H, ~2^-20) and one third-order step
 * y1 = y0 (1 + e/2 + 3 e^2/8), e = 1 - d y0^2 (relative error ~ e^3, below 2^-58).  The library rsqrt() carries a
 * slow-path call per use, which ends the basic block and keeps the scheduler from overlapping the latency of the
 * pivot chain with the independent updates of the block (tools/chainbench.cu: 1294 -> 1133 cycles per block).
 * Pivots are > 0 and far from the denormal range (lambda sits on every diagonal entry); a non-positive pivot
 * gives NaN/Inf, is flagged, and the solve is reported as failed like LinearSolverDense::solve does. */
DS_FN double rsqrt_fast(double d) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double h = y0 * y0;
  const double e = fma(-d, h, 1.0);
  const double p = fma(0.375, e, 0.5);
  const double t = y0 * e;
  return fma(t, p, y0);
}
/* right-looking Cholesky of the packed lower triangle a[i(i+1)/2 + j] in registers, reciprocal diagonals */
DS_FN bool chol8_regs(double *a) {
  bool bad = false;
#pragma unroll
  for (int k = 0; k < NB; k++) {
    bad = bad || !(a[k * (k + 1) / 2 + k] > 0.0);
    const double inv = rsqrt_fast(a[k * (k + 1) / 2 + k]);
    a[k * (k + 1) / 2 + k] = inv;
#pragma unroll
    for (int i = k + 1; i < NB; i++) a[i * (i + 1) / 2 + k] *= inv;
#pragma unroll
    for (int i = k + 1; i < NB; i++)
#pragma unroll
      for (int j = k + 1; j <= i; j++) a[i * (i + 1) / 2 + j] -= a[i * (i + 1) / 2 + k] * a[j * (j + 1) / 2 + k];
  }
  return bad;
}
DS_FN void invcol8_regs(const double *a, int j, double *col) {
  double s[NB];
#pragma unroll
  for (int i = 0; i < NB; i++) s[i] = i == j ? 1.0 : 0.0;
#pragma unroll
  for (int m = 0; m < NB; m++) {
    const double xm = s[m] * a[m * (m + 1) / 2 + m];
    col[m] = m >= j ? xm : 0.0;
#pragma unroll
    for (int i = m + 1; i < NB; i++) s[i] -= a[i * (i + 1) / 2 + m] * xm;
  }
}

/* The whole dependency chain of the factorisation runs here without leaving the warp: for row I the last panel tile
 * X = C(I,I-1) inv(L_{I-1,I-1})^T, the last update of the diagonal block C(I,I) -= X X^T, the Cholesky of that block
 * and the inverse of its factor.  The row's owner has brought both tiles up to column I-2 and handed them over
 * (pre[I]) long before they are needed. */
template <int NT>
DS_FN void rows_chain_warp(const RowShared &S, int nblk, int R, double *Lt, double *Dinv, int *flag, int lane) {
  constexpr int NBK = NT - 1;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t pair_off = frag_pair_off(g, q), lane_off = 8u * (uint32_t)lane;
  bool bad_any = false;
  DS_PROF_LOCALS(pacc, 2);
  DS_PROF_T0M(pt);
  for (int I = 0; I < nblk; I++) {
    spin_ge(&S.pre[I], 1);
    DS_PROF_LAP(pacc, 0, pt); /* waiting for the owner's hand-over */
    double *D = S.dbuf + (I & 1) * 64; /* row-major 8x8 = accumulator order: lane's pair at 2*lane */
    dbl2 dd = *(const dbl2 *)(D + 2 * lane);
    double x0 = 0.0, x1 = 0.0;
    const uint32_t tile = S.ring + 512u * (uint32_t)((I % R) * NT + (NBK - 1));
    if (I > 0) {
      const uint32_t inv = S.ring + 512u * (uint32_t)(((I - 1) % R) * NT + NBK);
      const double y0 = lds_f64(inv + lane_off), y1 = lds_f64(inv + 256u + lane_off);
      const double c0 = lds_f64(tile + lane_off), c1 = lds_f64(tile + 256u + lane_off);
      dmma884(x0, x1, c0, y0);
      dmma884(x0, x1, c1, y1);
      __syncwarp(); /* every lane has read C before X overwrites it */
      sts_v2f64(tile + pair_off, x0, x1);
      __syncwarp();
      const double xa0 = lds_f64(tile + lane_off), xa1 = lds_f64(tile + 256u + lane_off);
      publish(&S.prog[I], NBK, lane); /* tile (I, I-1) is final */
      dmma884(dd.x, dd.y, -xa0, xa0);
      dmma884(dd.x, dd.y, -xa1, xa1);
      *(dbl2 *)(D + 2 * lane) = dd;
      __syncwarp();
    }
    double a[36];
#pragma unroll
    for (int i = 0; i < NB; i++)
#pragma unroll
      for (int jj = 0; jj <= i; jj++) a[i * (i + 1) / 2 + jj] = D[i * 8 + jj];
    bad_any = chol8_regs(a) || bad_any;
    double col[NB];
    const int j = lane & 7;
    invcol8_regs(a, j, col);
    if (lane < NB) {
      /* inv(L_II) takes the diagonal slot of row I in the ring, in fragment order: Y[m][j] at (j/4)*32 + m*4 + j%4 */
      const uint32_t it = S.ring + 512u * (uint32_t)((I % R) * NT + NBK);
#pragma unroll
      for (int m = 0; m < NB; m++)
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(it + 8u * (uint32_t)((j >> 2) * 32 + m * 4 + (j & 3))), "d"(col[m]) : "memory");
    }
    publish(&S.ddone[I], 1, lane);
    /* global copies for the backward sweep, off the chain */
    if (lane < NB) {
      double *dg = Dinv + I * 64;
#pragma unroll
      for (int m = 0; m < NB; m++) dg[m * 8 + j] = col[m];
    }
    if (I > 0) *(dbl2 *)(Lt + ((size_t)I * NT + (NBK - 1)) * LT_STRIDE + g * 8 + 2 * q) = dbl2{x0, x1};
    DS_PROF_LAP(pacc, 1, pt); /* last tile + factor + inverse + publish */
  }
  DS_PROF_FLUSH(pacc, 2, PF_X_WARP, lane == 0);
  if (bad_any && lane == 0) *flag = 1;
}


/* the SfT band: row i holds H[i][j] at column j - i + bwE, stride ld (sft_core.h); lambda on the diagonal */
struct SftBandLoader {
  const double *Hb;
  int ld, bwE, bw;
  double lambda;
  template <int NT>
  DS_FN void load(int I, int g, int q, double *a0, double *a1) const {
    constexpr int NBK = NT - 1;
    const int lo = bwE - bw;
    const int i = NB * I + g;
    const double *rowp = Hb + (size_t)i * ld;
    /* Straight-line for the tiles left of the diagonal: every pair is requested (from a clamped, valid address where
     * it lies outside the band) before any is used.  With a branch per tile the loads went out one memory round trip
     * after the other: 4 k cycles per block row of an owner, a quarter of its time, with the chain warp waiting. */
#pragma unroll
    for (int t = 0; t < NBK; t++) {
      const int off = NB * (I - NBK + t) + 2 * q - i + bwE; /* even, <= bwE - 2 */
      /* a row is 16-byte aligned at offsets of its own parity; off = -1 (odd row, its first column the second of the
       * pair) reads the last slot of the row above with it */
#if DS_ROWS_LOAD_CG
      const dbl2 w = ldcg_dbl2((const dbl2 *)(rowp + (off >= -1 ? off : (i & 1))));
#else
      const dbl2 w = *(const dbl2 *)(rowp + (off >= -1 ? off : (i & 1)));
#endif
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
    if (2 * q == g) a0[NT - 1] += lambda;
    if (2 * q + 1 == g) a1[NT - 1] += lambda;
  }
  /* L2 prefetch of block row I: 4 lanes per row, one 128-byte line each per trip */
  DS_FN void prefetch(int I, int lane) const {
    const char *seg = (const char *)(Hb + (size_t)(NB * I + (lane >> 2)) * ld + (bwE - bw));
    const int nline = ((bw + 1) * 8 + 127) / 128 + 1;
    for (int l = (lane & 3); l < nline; l += 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(seg + 128 * l));
  }
};

/* ------------------------------------------------------------------ row owners */
template <int NT, class Loader>
DS_FN void rows_owner_warp(const RowShared &S, int widx, int nown, int nblk, int R, const Loader &ldr, double *Lt, int lane) {
  constexpr int NBK = NT - 1;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t pair_off = frag_pair_off(g, q), lane_off = 8u * (uint32_t)lane;
  DS_PROF_LOCALS(oacc, 2);
  DS_PROF_T0M(ot);
  for (int I = widx; I < nblk; I += nown) {
    /* ---- the row's band of H into accumulator registers (lane (g,q): row g, columns 2q, 2q+1 of each tile) */
    double a0[NT], a1[NT];
    ldr.template load<NT>(I, g, q, a0, a1);
    DS_PROF_LAP(oacc, 0, ot); /* the row's band read */
#if DS_ROWS_PREFETCH
    /* experiment (not kept: C2 batch 3 % slower): L2 prefetch of the band of the row this warp takes next */
    if (I + nown < nblk) ldr.prefetch(I + nown, lane);
#endif
    /* the slot's previous tenant (row I-R) was last read by the border warp */
    if (I >= R) spin_ge(S.edone, I - R + 1);
    const uint32_t slot = S.ring + 512u * (uint32_t)((I % R) * NT);
    double *ltrow = Lt + (size_t)I * NT * LT_STRIDE;
#pragma unroll
    for (int t = 0; t < NBK - 1; t++) {
      const int k = I - NBK + t;
      if (k < 0) continue; /* rows at the top: the tile does not exist */
      /* row I-1 has finished its tile of column k (=> so has every row above it, and inv(L_kk) exists) */
#if defined(DS_SPIN_SLEEP_OWNERS)
      spin_ge_relaxed(&S.prog[I - 1], t + 2);
#else
      spin_ge(&S.prog[I - 1], t + 2);
#endif
      const uint32_t tile = slot + 512u * (uint32_t)t;
      const uint32_t inv = S.ring + 512u * (uint32_t)((k % R) * NT + NBK);
      /* X = C inv(L_kk)^T: C goes through the tile's own ring location to change from accumulator to operand order */
      const double y0 = lds_f64(inv + lane_off), y1 = lds_f64(inv + 256u + lane_off);
      sts_v2f64(tile + pair_off, a0[t], a1[t]);
      __syncwarp();
      const double c0 = lds_f64(tile + lane_off), c1 = lds_f64(tile + 256u + lane_off);
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, c0, y0);
      dmma884(x0, x1, c1, y1);
      __syncwarp(); /* every lane has read C before X overwrites it */
      sts_v2f64(tile + pair_off, x0, x1);
      __syncwarp();
      const double xa0 = lds_f64(tile + lane_off), xa1 = lds_f64(tile + 256u + lane_off);
      publish(&S.prog[I], t + 1, lane);
      *(dbl2 *)(ltrow + t * LT_STRIDE + g * 8 + 2 * q) = dbl2{x0, x1}; /* for the backward sweep */
      const double n0 = -xa0, n1 = -xa1;
      /* C(I,J) -= X L(J,k)^T for the tiles to the right; L(J,k) is tile t - u + NBK of row J = I - NBK + u */
#pragma unroll
      for (int u = t + 1; u < NBK; u++) {
        const int J = I - NBK + u;
        const uint32_t bt = S.ring + 512u * (uint32_t)((J % R) * NT + (t - u + NBK));
        const double b0 = lds_f64(bt + lane_off), b1 = lds_f64(bt + 256u + lane_off);
        dmma884(a0[u], a1[u], n0, b0);
        dmma884(a0[u], a1[u], n1, b1);
      }
      dmma884(a0[NBK], a1[NBK], n0, xa0); /* the diagonal block: L(I,k) is X itself */
      dmma884(a0[NBK], a1[NBK], n1, xa1);
    }
    /* hand the last panel tile (operand order, in its ring location) and the diagonal block (row-major = accumulator
     * order) over to the chain warp: both are complete up to column I-2 */
    sts_v2f64(slot + 512u * (uint32_t)(NBK - 1) + pair_off, a0[NBK - 1], a1[NBK - 1]);
    *(dbl2 *)(S.dbuf + (I & 1) * 64 + 2 * lane) = dbl2{a0[NBK], a1[NBK]};
    publish(&S.pre[I], 1, lane);
    DS_PROF_LAP(oacc, 1, ot);
  }
  DS_PROF_FLUSH(oacc, 2, PF_X_WARP + 2 + 2 * widx, lane == 0);
}

/* ------------------------------------------------------------------ border rows (warp 4) */
template <int NT>
DS_FN void rows_border_warp(const RowShared &S, int nblk, int R, const double *Cg, double *Eg, int ES, double *G,
                            const double *Hcc, double lambda, int lane) {
  constexpr int NBK = NT - 1;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t pair_off = frag_pair_off(g, q), lane_off = 8u * (uint32_t)lane;
  double ga0 = 0.0, ga1 = 0.0, gb0 = 0.0, gb1 = 0.0; /* sum over J of E_J E_J^T, two independent chains */
  for (int J = 0; J < nblk; J++) {
    /* border block of H (rows: camera border 0-5, rhs 6, pad 7; columns 8J..8J+7) */
    const dbl2 h = *(const dbl2 *)(Cg + (size_t)g * ES + NB * J + 2 * q);
    double e0 = h.x, e1 = h.y, f0 = 0.0, f1 = 0.0;
    spin_ge_relaxed(&S.ddone[J], 1); /* row J is final (its tiles and inv(L_JJ) are in the ring) */
    const uint32_t slot = S.ring + 512u * (uint32_t)((J % R) * NT);
#pragma unroll
    for (int t = 0; t < NBK; t++) {
      const int k = J - NBK + t;
      if (k < 0) continue;
      const uint32_t et = S.ering + 512u * (uint32_t)(k % NT), bt = slot + 512u * (uint32_t)t;
      const double x0 = -lds_f64(et + lane_off), x1 = -lds_f64(et + 256u + lane_off);
      const double b0 = lds_f64(bt + lane_off), b1 = lds_f64(bt + 256u + lane_off);
      if (t & 1) { dmma884(f0, f1, x0, b0); dmma884(f0, f1, x1, b1); }
      else { dmma884(e0, e1, x0, b0); dmma884(e0, e1, x1, b1); }
    }
    e0 += f0; e1 += f1;
    /* E_J = (...) inv(L_JJ)^T */
    const uint32_t tile = S.ering + 512u * (uint32_t)(J % NT), inv = slot + 512u * (uint32_t)NBK;
    const double y0 = lds_f64(inv + lane_off), y1 = lds_f64(inv + 256u + lane_off);
    __syncwarp(); /* the slot's previous tenant (E tile J-NT) is no longer needed by any lane */
    sts_v2f64(tile + pair_off, e0, e1);
    __syncwarp();
    const double c0 = lds_f64(tile + lane_off), c1 = lds_f64(tile + 256u + lane_off);
    double x0 = 0.0, x1 = 0.0;
    dmma884(x0, x1, c0, y0);
    dmma884(x0, x1, c1, y1);
    __syncwarp();
    sts_v2f64(tile + pair_off, x0, x1);
    *(dbl2 *)(Eg + (size_t)g * ES + NB * J + 2 * q) = dbl2{x0, x1};
    __syncwarp();
    const double xa0 = lds_f64(tile + lane_off), xa1 = lds_f64(tile + 256u + lane_off);
    dmma884(ga0, ga1, xa0, xa0);
    dmma884(gb0, gb1, xa1, xa1);
    publish(S.edone, J + 1, lane);
  }
  /* corner: camera block + lambda (rows 0-5), right-hand side (row 6), minus the accumulated E E^T
   * (G == nullptr: plain right-hand sides, no corner system) */
  if (G == nullptr) return;
  ga0 += gb0; ga1 += gb1;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int b = 2 * q + s;
    double v = 0.0;
    if (g < 6 && b < 6) v = Hcc[g * 6 + b] + (g == b ? lambda : 0.0);
    else if (g == 6 && b < 6) v = Hcc[36 + b];
    G[g * 8 + b] = v - (s ? ga1 : ga0);
  }
}

/* The factorisation proper.  sy: 3*nblk + 2 progress counters followed by one int per warp; W: ring of
 * R = NT-1 + max_owners block rows of NT tiles; er: ring of NT border tiles; db: 128 doubles.  Cg / Eg: the 8 border
 * rows (stride ES) before / after the forward substitution.  G != nullptr: the corner (Schur) block of the SfT
 * arrowhead system is formed from Hcc and lambda.  All warps of the CTA call; ends with a CTA-wide barrier.
 * *flag != 0 afterwards: a pivot was not positive. */
template <int NT, class Loader>
DS_FN void rows_factor(int tid, int nthr, int nblk, int max_owners, const Loader &ldr, double *W, double *er, double *db,
                       int *sy, int *flag, const double *Cg, double *Eg, int ES, double *G, const double *Hcc,
                       double lambda, double *Lt, double *Dinv) {
  constexpr int NBK = NT - 1;
  const int R = NBK + max_owners;
  const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
  for (int i = tid; i < 3 * nblk + 2; i += nthr) sy[i] = 0;
  if (tid == 0) *flag = 0;
  /* Roles follow the HARDWARE warp slot (%warpid; scheduler / sub-partition = slot % 4), not the logical warp
   * index: the second CTA of an SM gets its slots rotated (tools/warpmap.cu: logical warp 0 -> slot 9), and the
   * dependency chain must sit on a sub-partition where neither CTA issues tensor-core work.  Every CTA puts its
   * chain warp on sub-partition 0 and idles its other warps there; results do not depend on who does what. */
  int *wsp = sy + 3 * nblk + 2;
  if (lane == 0) {
    unsigned wid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
#if DS_ROWS_SPLIT_CHAINS
    wsp[warp] = (int)wid;
#else
    wsp[warp] = (int)(wid & 3u);
#endif
  }
  __syncthreads();
  {
#if DS_ROWS_SPLIT_CHAINS
    /* experiment: the two CTAs of an SM put their chains on DIFFERENT sub-partitions (0 for the CTA in warp slots
     * 0-7, 1 for the one in slots 8-15) and both keep their tensor-core work off sub-partitions 0 and 1: two chains
     * sharing a sub-partition run 1.5x slower each (tools/chainbench.cu: 904 -> 1375 cycles per block) */
    int minslot = 1 << 30;
    for (int w = 0; w < nwarp; w++) minslot = wsp[w] < minslot ? wsp[w] : minslot;
    unsigned nsm;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    const bool paired = gridDim.x > nsm && nwarp == 8;
    const int want_sp = paired ? ((minslot >> 3) & 1) : 0;
    int chain_w = 0;
    for (int w = nwarp - 1; w >= 0; w--) if ((wsp[w] & 3) == want_sp) chain_w = w;
    const int chain_sp = wsp[chain_w] & 3;
    __syncthreads();
    if (lane == 0) { const int sp = wsp[warp] & 3; wsp[warp] = (paired && sp < 2) ? chain_sp : sp; } /* sub-partitions 0, 1: no workers */
    __syncthreads();
#else
    int chain_w = 0;
    for (int w = nwarp - 1; w >= 0; w--) if (wsp[w] == 0) chain_w = w;
    const int chain_sp = wsp[chain_w];
#endif
    /* the border warp (24 DMMAs per block row, a sixth of an owner's load) shares the chain's sub-partition when a
     * second warp sits there only with DS_ROWS_BORDER_ON_CHAIN_SP (measured: C2 1 % faster, C1/C3/C4 5-9 % slower --
     * the chain loses more than the sixth owner gains; default off); every other warp on that sub-partition idles */
    int border_w = -1;
#if DS_ROWS_BORDER_ON_CHAIN_SP
    for (int w = nwarp - 1; w >= 0; w--) if (w != chain_w && wsp[w] == chain_sp) border_w = w;
#endif
    /* workers: every warp off the chain's sub-partition; without a border warp yet the first one takes that role */
    int my = -1, nworkers = 0;
    for (int w = 0; w < nwarp; w++) {
      if (w == chain_w || wsp[w] == chain_sp) continue;
      if (w == warp) my = nworkers;
      nworkers++;
    }
    if (nworkers < 2) { /* degenerate slot assignment: fall back to logical roles */
      my = warp == chain_w ? -1 : (warp > chain_w ? warp - 1 : warp);
      nworkers = nwarp - 1;
      border_w = -1;
    }
    if (border_w >= 0) { if (warp == border_w) my = 0; else if (my >= 0) my += 1; nworkers += 1; }
    const int nown = nworkers - 1 < max_owners ? nworkers - 1 : max_owners;
    RowShared S;
    S.ring = smem_u32(W);
    S.ering = smem_u32(er);
    S.dbuf = db;
    S.prog = sy; S.ddone = sy + nblk; S.pre = sy + 2 * nblk; S.edone = sy + 3 * nblk;
    if (warp == chain_w) rows_chain_warp<NT>(S, nblk, R, Lt, Dinv, flag, lane);
    else if (my == 0) rows_border_warp<NT>(S, nblk, R, Cg, Eg, ES, G, Hcc, lambda, lane);
    else if (my > 0 && my <= nown) rows_owner_warp<NT>(S, my - 1, nown, nblk, R, ldr, Lt, lane);
  }
  __syncthreads();
}

/* Backward sweep L^T x = y by block rows, bottom up: d = inv(L_kk)^T y_k, then y_J -= L(k,J)^T d.  dx (shared
 * memory): y on entry; sol (shared memory): x on return.  W: NBUF buffers of BUFD = NT*LT_STRIDE + 64 doubles (the
 * idle tile ring); sy: nblk + 1 ints, zeroed before a CTA-wide barrier that precedes the call.  All warps of the
 * CTA call; the caller synchronises afterwards.
 *
 * ONE warp runs the sweep: the chain d_k -> y_{k-1} -> d_{k-1} has no parallelism across block rows, and a
 * CTA-wide barrier per block row cost more than the 8 x NBK columns of a row give back when they are spread over
 * eight warps (1.2 k cycles per block row).  As a single warp a row is: its columns of the row block, its entries
 * of inv(L_kk) and y_k requested from shared memory up front, d by lanes + broadcast, <= 4 independent 8-term
 * chains per lane, one warp barrier.  The other warps are the producers: warp w copies block rows w-1, w-1+P, ...
 * of the factor from the workspace (L2) into the ring with plain 16-byte loads and raises the row's flag; they
 * follow the sweep's progress counter NBUF rows ahead.  (Bulk copies issued by the sweeping warp itself cost it
 * 320 cycles per row for the issue and 90 for each mbarrier poll -- measured -- on a row that needs about 300.) */
template <int NT, int NBUF>
DS_FN void rows_backward(int tid, int nthr, int nblk, double *W, int *sy, const double *Lt, const double *Dinv, double *dx,
                         double *sol) {
  constexpr int NBK = NT - 1;
  constexpr int BUFD = NT * LT_STRIDE + 64;
  const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
  int *ready = sy, *done = sy + nblk;
  const int P = nwarp - 1;
  if (warp > 0) {
    constexpr int NV = (BUFD / 2 + 31) / 32; /* 16-byte pieces of a row block per lane */
    for (int j = warp - 1; j < nblk; j += P) {
      const int kbj = nblk - 1 - j;
      const dbl2 *srcL = (const dbl2 *)(Lt + (size_t)kbj * NT * LT_STRIDE), *srcY = (const dbl2 *)(Dinv + kbj * 64);
      if (j >= NBUF) spin_ge_relaxed(done, j - NBUF + 1); /* the buffer's previous tenant has been consumed */
      dbl2 *dst = (dbl2 *)(W + (j % NBUF) * BUFD);
#pragma unroll
      for (int i0 = 0; i0 < NV; i0 += 8) { /* eight loads in flight per lane */
        dbl2 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int e = lane + 32 * (i0 + i);
          if (i0 + i < NV) {
            if (e < NT * LT_STRIDE / 2) v[i] = ldcg_dbl2(srcL + e);
            else if (e < BUFD / 2) v[i] = ldcg_dbl2(srcY + (e - NT * LT_STRIDE / 2));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int e = lane + 32 * (i0 + i);
          if (i0 + i < NV && e < BUFD / 2) dst[e] = v[i];
        }
      }
      publish(&ready[j], 1, lane);
    }
  } else {
    constexpr int RN = (NB * NBK + 31) / 32; /* columns of a block row per lane */
    DS_PROF_LOCALS(bwacc, 4);
    DS_PROF_T0M(bwt);
    spin_ge(&ready[0], 1);
    DS_PROF_LAP(bwacc, 2, bwt);
    for (int kb = nblk - 1; kb >= 0; kb--) {
      const int k = kb * NB;
      const int j = nblk - 1 - kb, buf = j % NBUF;
      const double *LR = W + buf * BUFD, *Y = LR + NT * LT_STRIDE;
      const int t0 = kb < NBK ? NBK - kb : 0; /* first tile of the row that exists */
      const int nupd = NB * (NBK - t0);
      /* the block next to the diagonal first: the next row's d depends on it alone */
      double lv[RN][NB], yv[RN];
      int jcs[RN];
      bool valid[RN];
#pragma unroll
      for (int r = 0; r < RN; r++) {
        /* lanes past the row's last column load from a valid address and do not store */
        const int jj = lane + 32 * r;
        const int tt = NBK - 1 - (jj >> 3), t = tt > 0 ? tt : 0, cc = jj & 7;
        valid[r] = jj < nupd;
        jcs[r] = valid[r] ? NB * (kb - NBK + t) + cc : 0;
        const double *Lc = LR + t * LT_STRIDE + cc; /* Lc[a*8] = L[k+a][jc] */
#pragma unroll
        for (int a = 0; a < NB; a++) lv[r][a] = Lc[a * 8];
        yv[r] = dx[jcs[r]];
      }
      const int la = lane & 7;
      double ya[NB], xk[NB];
      /* (the entries of inv(L_kk) above the diagonal are stored zeros) */
#pragma unroll
      for (int m = 0; m < NB; m++) { ya[m] = Y[m * 8 + la]; xk[m] = dx[k + m]; }
      /* the next row's flag is read now and looked at after this row's arithmetic */
      const int next_ready = kb > 0 ? ld_vol_s32(&ready[j + 1]) : 1;
      /* d[a] = sum_{m >= a} Y[m][a] y[m] (Y = inv(L_kk), row-major): lane a (mod 8) forms d[a], then broadcast */
      double d[NB];
      {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int m = 0; m < NB; m += 2) {
          s0 = fma(ya[m], xk[m], s0);
          s1 = fma(ya[m + 1], xk[m + 1], s1);
        }
        const double mine = s0 + s1;
        if (lane < NB) sol[k + lane] = mine;
#pragma unroll
        for (int a = 0; a < NB; a++) d[a] = __shfl_sync(0xffffffffu, mine, a);
      }
#pragma unroll
      for (int a = 0; a < NB; a++)
#pragma unroll
        for (int r = 0; r < RN; r++) yv[r] -= lv[r][a] * d[a];
#pragma unroll
      for (int r = 0; r < RN; r++)
        if (valid[r]) dx[jcs[r]] = yv[r];
      __syncwarp();
      if (lane == 0) st_vol_s32(done, j + 1);
      DS_PROF_LAP(bwacc, 1, bwt); /* loads, d, update, stores */
      if (next_ready < 1) {
        spin_ge(&ready[j + 1], 1);
        DS_PROF_LAP(bwacc, 2, bwt); /* the copy had not landed */
      }
    }
    DS_PROF_FLUSH(bwacc, 4, PF_X_BWD, lane == 0);
  }
}

#endif /* DS_CUDA */

}  // namespace ds
#endif
