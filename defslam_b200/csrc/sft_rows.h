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

namespace ds {

constexpr int ROW_OWNERS = ROWS_OWNERS_;
constexpr int LT_STRIDE = ROWS_LT_STRIDE_; /* doubles per 8x8 tile of the factor in global memory (64 + pad: the backward
                                 sweep's column reads of two neighbouring tiles fall on different banks) */

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

#if DS_CUDA
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

/* ------------------------------------------------------------------ row owners */
template <int NT>
DS_FN void rows_owner_warp(const RowShared &S, int widx, int nown, int nblk, int R, const double *Hb, double *Lt, int ld,
                           int bwE, int bw, double lambda, int lane) {
  constexpr int NBK = NT - 1;
  const int g = lane >> 2, q = lane & 3;
  const int lo = bwE - bw;
  const uint32_t pair_off = frag_pair_off(g, q), lane_off = 8u * (uint32_t)lane;
  DS_PROF_LOCALS(oacc, 2);
  DS_PROF_T0M(ot);
  for (int I = widx; I < nblk; I += nown) {
    /* ---- the row's band of H into accumulator registers (lane (g,q): row g, columns 2q, 2q+1 of each tile) */
    double a0[NT], a1[NT];
    {
      const int i = NB * I + g;
      const double *rowp = Hb + (size_t)i * ld;
#pragma unroll
      for (int t = 0; t < NT; t++) {
        const int J = I - NBK + t;
        const int off = NB * J + 2 * q - i + bwE;
        const bool v0 = J >= 0 && off >= lo && off <= bwE, v1 = J >= 0 && off + 1 >= lo && off + 1 <= bwE;
        double x0 = 0.0, x1 = 0.0;
        if (v0 && v1) { const dbl2 v = *(const dbl2 *)(rowp + off); x0 = v.x; x1 = v.y; }
        else { if (v0) x0 = rowp[off]; if (v1) x1 = rowp[off + 1]; }
        a0[t] = x0; a1[t] = x1;
      }
      if (2 * q == g) a0[NT - 1] += lambda;
      if (2 * q + 1 == g) a1[NT - 1] += lambda;
    }
#if DS_ROWS_PREFETCH
    /* the band of the row this warp takes next: HBM -> L2 while this row is being worked on */
    if (I + nown < nblk) {
      const int i2 = NB * (I + nown) + (lane >> 2);
      const char *seg = (const char *)(Hb + (size_t)i2 * ld + lo);
      const int nline = ((bw + 1) * 8 + 127) / 128 + 1;
      for (int l = (lane & 3); l < nline; l += 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(seg + 128 * l));
    }
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
  /* corner: camera block + lambda (rows 0-5), right-hand side (row 6), minus the accumulated E E^T */
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
#endif /* DS_CUDA */

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
  const int warp = team.tid >> 5, lane = team.tid & 31, nwarp = team.nthr >> 5;
  DS_FOR(i, 3 * nblk + 2) sy[i] = 0;
  if (team.tid == 0) *flag = 0;
  /* Roles follow the HARDWARE warp slot (%warpid; scheduler / sub-partition = slot % 4), not the logical warp
   * index: the second CTA of an SM gets its slots rotated (tools/warpmap.cu: logical warp 0 -> slot 9), and the
   * dependency chain must sit on a sub-partition where neither CTA issues tensor-core work.  Every CTA puts its
   * chain warp on sub-partition 0 and idles its other warps there; results do not depend on who does what. */
  int *wsp = sy + 3 * nblk + 2;
  if (lane == 0) {
    unsigned wid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    wsp[warp] = (int)(wid & 3u);
  }
  team.sync();
  {
    int chain_w = 0;
    for (int w = nwarp - 1; w >= 0; w--) if (wsp[w] == 0) chain_w = w;
    const int chain_sp = wsp[chain_w];
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
    const int nown = nworkers - 1 < ROW_OWNERS ? nworkers - 1 : ROW_OWNERS;
    RowShared S;
    S.ring = smem_u32(W);
    S.ering = smem_u32(sm + c.sl.er);
    S.dbuf = sm + c.sl.db;
    S.prog = sy; S.ddone = sy + nblk; S.pre = sy + 2 * nblk; S.edone = sy + 3 * nblk;
    if (warp == chain_w) rows_chain_warp<NT>(S, nblk, R, Lt, Dinv, flag, lane);
    else if (my == 0) rows_border_warp<NT>(S, nblk, R, Cg, Eg, ES, G, Hcc, lambda, lane);
    else if (my > 0 && my <= nown) rows_owner_warp<NT>(S, my - 1, nown, nblk, R, Hb, Lt, ld, bwE, bw, lambda, lane);
  }
  team.sync();
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
  /* ONE warp runs the sweep: the chain d_k -> y_{k-1} -> d_{k-1} has no parallelism across block rows, and a
   * CTA-wide barrier per block row cost more than the 8 x NBK columns of a row give back when they are spread over
   * eight warps (1.2 k cycles per block row).  As a single warp a row is: its columns of the row block, its entries
   * of inv(L_kk) and y_k requested from shared memory up front, d by lanes + broadcast, <= 4 independent 8-term
   * chains per lane, one warp barrier.  The other warps are the producers: warp w copies block rows w-1, w-1+P, ...
   * of the factor from the workspace (L2) into the ring with plain 16-byte loads and raises the row's flag; they
   * follow the sweep's progress counter NBUF rows ahead.  (Bulk copies issued by the sweeping warp itself cost it
   * 320 cycles per row for the issue and 90 for each mbarrier poll -- measured -- on a row that needs about 300.)
   * Same operations in the same order per entry as before: bit-identical. */
  {
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
