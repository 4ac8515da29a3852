/*
 * sft_core.h -- the Shape-from-Template solve as ONE persistent CTA per frame.
 *
 * Replaces, for the graph Modules/Tracking/DefOptimizer.cc:251-578 builds:
 *   sft_types.h edges (computeError / linearizeOplus)      -> eval_state(), build_system()
 *   BaseMultiEdge/BinaryEdge/UnaryEdge::constructQuadraticForm
 *   + RobustKernelHuber::robustify                          -> build_system()
 *   BlockSolver::buildSystem/setLambda/solve + LinearSolverDense::solve
 *                                                           -> factor_solve()
 *   OptimizationAlgorithmLevenberg::solve + SparseOptimizer::optimize/update
 *                                                           -> sft_solve_one()
 *   outlier classing / repError / write-back (DefOptimizer.cc:515-577)
 *                                                           -> finalize()
 *
 * Not a translation: the graph is never materialised.  Because the reference's
 * node Jacobian of a reprojection edge is  b_k * A(node k)  with A depending
 * only on the node and the pose (sft_types.h:176-205, quirk C1), the whole
 * reprojection part of J^T W J collapses to per-facet sums of 48 scalars plus a
 * 6x6 camera block; curvature and stretch terms are rank-1 per centre / edge.
 * H is assembled by a deterministic gather (one thread per 3x3 block) into an
 * arrowhead layout -- node band (half bandwidth bw) + 6 camera rows -- and
 * (H + lambda I) dx = b is solved by a blocked banded Cholesky that slides an
 * (NB+bw) x (bw+1) window through shared memory, with the camera border and the
 * right-hand side carried as 8 extra rows of the same trailing update.
 *
 * Variable order inside the kernel: node 0 xyz, node 1 xyz, ..., camera
 * (omega, upsilon) last.  Fixed nodes (outside OptLap) keep identity rows.
 */
#ifndef DS_SFT_CORE_H_
#define DS_SFT_CORE_H_

#include "ds_common.h"
#include "ds_plan.h"
#include "ds_se3.h"
#include "ds_tma.h"

namespace ds {

struct ResultScalars {
  float Tcw[16];
  float rep_error;
  int n_inliers;
  int lm_iterations;
  int lm_trials;
  double chi2_initial;
  double chi2_final;
  double lambda_final;
  int status;
  int n_viewed;
  int n_optlap;
  int pad;
};

enum { MODE_SOLVE = 0, MODE_NORMAL_EQ = 1 };

struct ProbView {
  const PlanView *plan;    /* resident with the template */
  int mode;
  int n_matches, n_kp, max_it, layers;
  int trace_cap;
  int e_in_smem;
  int x_in_smem;           /* node positions / LM step in shared memory (else in the global workspace) */
  int variant;             /* 1: the matches-given overload (DefOptimizer.cc:582-837) */
  double curv_len;         /* variant 1: EdgeMeanCurvature::lenghtEdge_ (quirk C8) */
  int row_nt;              /* > 0: row-owner factorisation (sft_rows.h) with this many tiles per block row; 0: sliding window */
  double fx, fy, cx, cy;
  double reg_lap, reg_inex, reg_temp;
  float Tcw[16];
  const double *node_xyz;  /* [3n] */
  const int *match_nodes;  /* [3M] */
  const double *match_bary;/* [3M] */
  const float *match_uv;   /* [2M] */
  const float *match_isig; /* [M]  */
  double *out_nodes;       /* [3n] */
  uint8_t *out_outlier;    /* [M]  */
  uint8_t *out_role;       /* [n]  */
  double *out_trace;       /* [4*trace_cap] */
  ResultScalars *out_res;
  double *out_H;           /* MODE_NORMAL_EQ: [D*D] */
  double *out_b;           /* MODE_NORMAL_EQ: [D]   */
};

/* per-CTA scratch in global memory (L2 resident) */
struct Workspace {
  double *Hb;    /* [Dn_pad*ld]   band of H (lower, row i holds cols i-bw..i) */
  double *Lb;    /* [Dn_pad*ld]   band of the Cholesky factor                 */
  double *Dinv;  /* [nblk*64]     inverses of the diagonal blocks of L        */
  double *Cg;    /* [8*Dn_pad]    rows 0-5 camera border, row 6 b_n, row 7 0  */
  double *Eg;    /* [8*Dn_pad]    working copy when it does not fit in smem   */
  double *F;     /* [NFACC*nf]    per-facet accumulators                      */
  double *S;     /* [NMSCR*M]     per-match scratch                           */
  double *xb;    /* [Dn_pad]      LM backup of the node positions (push/pop)  */
  double *xg;    /* [Dn_pad]      node positions when they do not fit in smem */
  double *dxg;   /* [Dn_pad+8]    LM step when it does not fit in smem        */
  int *mfac;     /* [M]  facet<<6 | slot0 | slot1<<2 | slot2<<4               */
  int *mperm;    /* [M]  matches grouped by facet                             */
  double *mrec;  /* [8M] per-match records in facet-grouped order (see match_record) */
  int *fptr;     /* [nf+1] */
  int *fcnt;     /* [nf]   */
};

struct WorkspaceSizes {
  size_t band, dinv, cg, F, S, M, nf, dp; /* element counts (max over the batch) */
};

static inline
#if DS_CUDA
__host__ __device__
#endif
size_t workspace_bytes(const WorkspaceSizes &z) {
  size_t b = sizeof(double) * (2 * z.band + z.dinv + 2 * z.cg + z.F + z.S + 8 * z.M + 2 + 3 * z.dp + 8) + sizeof(int) * (2 * z.M + 2 * z.nf + 2);
  return (b + 255) & ~(size_t)255;
}

static inline
#if DS_CUDA
__host__ __device__
#endif
Workspace carve_workspace(uint8_t *base, const WorkspaceSizes &z) {
  Workspace w;
  double *d = (double *)base;
  w.Hb = d; d += z.band;
  w.Lb = d; d += z.band;
  w.Dinv = d; d += z.dinv;
  w.Cg = d; d += z.cg;
  w.Eg = d; d += z.cg;
  w.F = d; d += z.F;
  w.S = d; d += z.S;
  if ((d - (double *)base) & 1) d++; /* 16-byte records */
  w.mrec = d; d += 8 * z.M;
  w.xb = d; d += z.dp;
  w.xg = d; d += z.dp;
  w.dxg = d; d += z.dp + 8;
  int *i = (int *)d;
  w.mfac = i; i += z.M;
  w.mperm = i; i += z.M;
  w.fptr = i; i += z.nf + 1;
  w.fcnt = i;
  return w;
}

/* shared-memory carve-up (offsets in doubles) -- same function on host (size)
 * and device (pointers) */
struct SmemLayout {
  int W, E, P, x, xb, dx, Lkk, invL, G, Hcc, red, pose, tiles, flags, total;
  int er, db, sy; /* row-owner factorisation: border-tile ring, diagonal hand-off buffers, progress counters */
};

/* row-owner factorisation (sft_rows.h): ring slots and sizes */
constexpr int ROWS_OWNERS_ = 5;
constexpr int ROWS_LT_STRIDE_ = 64;
#ifndef DS_BWD_BUFS
#define DS_BWD_BUFS 8
#endif
constexpr int ROWS_BWD_BUFS_ = DS_BWD_BUFS; /* row blocks of the factor in flight during the backward sweep */

DS_FN int asm_scratch_doubles(int n, int ne) { return 11 * n + 5 * ne; }

/* panel buffer P[2][HS]: two halves (columns 0-3 / 4-7 of the panel) of PR = bwp + 8 rows x 4; the
 * half stride HS = 4 PR + 8 is 8 (mod 16) doubles, so the 16-byte stores of a quarter-warp that
 * straddle both halves fall on different banks */
static inline
#if DS_CUDA
__host__ __device__
#endif
int panel_half_stride(int bwp) { return 4 * (bwp + 8) + 8; }
static inline
#if DS_CUDA
__host__ __device__
#endif
int panel_doubles(int bwp) { const int d = 2 * panel_half_stride(bwp); return d > 216 ? d : 216; }

static inline
#if DS_CUDA
__host__ __device__
#endif
SmemLayout smem_layout(int n_nodes, int n_edges, int Dn_pad, int bwp, int ld, int Wr, int ES, bool e_in_smem,
                       bool x_in_smem = true, int row_nt = 0) {
  SmemLayout L;
  int o = 0;
  const int asz = 11 * n_nodes + 5 * n_edges; /* assembly scratch */
  L.er = L.db = L.sy = 0;
  if (row_nt > 0) {
    /* ring of (NT-1 + owners) block rows of NT tiles; backward sweep: 4 x (row block + inverse) + the solution */
    const int nblk = Dn_pad / NB;
    int wsz = (row_nt - 1 + ROWS_OWNERS_) * row_nt * 64;
    const int bsz = ROWS_BWD_BUFS_ * (row_nt * ROWS_LT_STRIDE_ + 64) + Dn_pad, esz = nblk * 64;
    if (asz > wsz) wsz = asz;
    if (bsz > wsz) wsz = bsz;
    if (esz > wsz) wsz = esz; /* (emulation keeps the border tiles there) */
    L.W = o;  o += wsz; o = (o + 1) & ~1;
    L.E = o;
    L.P = o;
    L.er = o; o += row_nt * 64;
    L.db = o; o += 128;
    L.sy = o; o += (3 * nblk + 2 + 16 + 1) / 2 + 1; o = (o + 1) & ~1; /* counters + the warps' sub-partitions */
  } else {
  int wsz = Wr * ld;
  /* backward sweep: ring of 4 row blocks + the solution vector */
  const int bsz = 4 * (NB * ld) + Dn_pad;
  if (asz > wsz) wsz = asz;
  if (bsz > wsz) wsz = bsz;
  L.W = o;    o += wsz; o = (o + 1) & ~1;
  L.E = o;    o += e_in_smem ? 8 * ES : 0;
  L.P = o;    o += panel_doubles(bwp); o = (o + 1) & ~1;
  }
  L.x = o;    o += x_in_smem ? Dn_pad : 0;
  L.xb = o;   /* (backup lives in global memory) */
  L.dx = o;   o += x_in_smem ? Dn_pad + 8 : 0;
  L.Lkk = o;  /* (unused) */
  L.invL = o; o += 192;  /* inv(L_kk), row stride 12; two buffers (step parity) */
  L.G = o;    o += 64;
  L.Hcc = o;  o += 48;   /* 36 Hcc + 6 bc + 6 dc(stale) */
  L.red = o;  o += 40;
  L.pose = o; o += 16;   /* pose (7) + backup (7) */
  {
    const int nrow = bwp / NB + 1;
    L.tiles = o; /* one int4 per trailing-update tile + per-warp ranges (sliding-window path only) */
    if (row_nt == 0) o += 2 * (nrow * (nrow + 1) / 2) + 10;
  }
  L.flags = o; o += (2 * n_nodes + 7) / 8 + 1;
  L.total = o;
  return L;
}

struct Ctx {
  PlanView pl;
  ProbView pb;
  Workspace ws;
  SmemLayout sl;
  int n_optlap, n_viewed, n_str;
  double info_ref, info_curv, info_str;
  double hub_delta, hub_dsqr;
  double inv_n;
  double info_uniform;     /* > 0: every reprojection edge has this information (matches-given overload) */
  long long *prof;  /* optional per-phase cycle counters (global), CTA 0 only */
  long long prof_last;
  uint32_t ph[16];  /* phase parity of each mbarrier */
  uint64_t mbar[16]; /* 0: forward window, 1..: backward ring (fixed address for the whole launch) */
};

/* doubles reserved at the head of shared memory for the CTA-wide context */
constexpr int CTX_DOUBLES = (int)((sizeof(Ctx) + 15) / 16) * 2;

/* All shared-memory pointers are derived from the kernel's dynamic shared array
 * inside each function, never carried through the context: that way the
 * compiler knows the address space and emits LDS/STS with 32-bit addressing
 * instead of generic 64-bit loads. */
#if DS_CUDA
extern __shared__ __align__(16) double ds_smem_raw[];
#define DS_SMEM ds_smem_raw
#else
static double *ds_smem_emu = nullptr;
#define DS_SMEM ds_smem_emu
#endif
DS_FN Ctx &ctx_ref() { return *(Ctx *)DS_SMEM; }
DS_FN double *sm_base() { return DS_SMEM + CTX_DOUBLES; }
/* node positions x and LM step dx: shared memory (XS) or, for meshes whose window leaves no room
 * (25 x 25 and up), the CTA's global workspace.  Compile-time so the common case keeps LDS/STS. */
template <bool XS> DS_FN double *xvec(const Ctx &c) { return XS ? sm_base() + c.sl.x : c.ws.xg; }
template <bool XS> DS_FN double *dxvec(const Ctx &c) { return XS ? sm_base() + c.sl.dx : c.ws.dxg; }
DS_FN uint8_t *viewed_ptr(const Ctx &c) { return (uint8_t *)(sm_base() + c.sl.flags); }
DS_FN uint8_t *freev_ptr(const Ctx &c) { return (uint8_t *)(sm_base() + c.sl.flags) + c.pl.n_nodes; }

/* phase-cycle accounting (diagnostics; enabled by DEFSLAM_PROFILE=1 on the host side) */
enum { PF_PROLOGUE = 0, PF_EVAL_STORE, PF_BUILD, PF_FS_INIT, PF_S1, PF_S1_WAIT, PF_S2, PF_S3, PF_SCHUR, PF_BWD_INIT,
       PF_BWD, PF_UPDATE, PF_EVAL_TRIAL, PF_LM_SCALAR, PF_FINALIZE, PF_COUNT,
       /* extras (not part of the total): busy cycles of each warp inside S3, look-ahead factor, steps */
       PF_X_WARP = PF_COUNT, PF_X_DIAG = PF_COUNT + 16, PF_X_STEPS, PF_X_S2W, PF_X_BUILD = PF_X_S2W + 16,
       PF_X_BUILDS = PF_X_BUILD + 8, PF_X_BWD, PF_X_FS = PF_X_BWD + 4, PF_TOTAL = PF_X_FS + 4 };
#if DS_CUDA && defined(DS_PROFILE)
#define DS_PROF_T0(var) const long long var = clock64()
#define DS_PROF_ADD(idx, t0, cond) do { if (ctx_ref().prof != nullptr && (cond)) ctx_ref().prof[idx] += clock64() - (t0); } while (0)
#define DS_PROF_INC(idx, cond) do { if (ctx_ref().prof != nullptr && (cond)) ctx_ref().prof[idx] += 1; } while (0)
/* register accumulators for fine-grained sections (a global += per section would stall the warp) */
#define DS_PROF_LOCALS(name, n) long long name[n] = {}
#define DS_PROF_LAP(name, i, t) do { const long long now_ = clock64(); name[i] += now_ - t; t = now_; } while (0)
#define DS_PROF_T0M(var) long long var = clock64()
#define DS_PROF_FLUSH(name, n, idx, cond) do { if (ctx_ref().prof != nullptr && (cond)) for (int i_ = 0; i_ < n; i_++) ctx_ref().prof[idx + i_] += name[i_]; } while (0)
#else
#define DS_PROF_LOCALS(name, n) do {} while (0)
#define DS_PROF_LAP(name, i, t) do {} while (0)
#define DS_PROF_T0M(var) do {} while (0)
#define DS_PROF_FLUSH(name, n, idx, cond) do {} while (0)
#define DS_PROF_T0(var) do {} while (0)
#define DS_PROF_ADD(idx, t0, cond) do {} while (0)
#define DS_PROF_INC(idx, cond) do {} while (0)
#endif
DS_FN void prof_mark(const Team team, Ctx &cx, int idx) {
  Ctx &c = ctx_ref();
  (void)cx;
#if DS_CUDA && defined(DS_PROFILE)
  if (c.prof != nullptr && team.tid == 0) {
    const long long now = clock64();
    c.prof[idx] += now - c.prof_last;
    c.prof_last = now;
  }
#else
  (void)team; (void)c; (void)idx;
#endif
}

/* Layouts (chosen for the FP64 tensor-core tiles of the factorisation):
 *  - band row i holds H[i][j] at column j - i + bwE (bwE even), row stride ld
 *    odd with ld == 9 (mod 16): the two adjacent columns a DMMA accumulator
 *    lane owns are 16-byte aligned in every row, and the 8 rows of a tile fall
 *    on distinct bank groups;
 *  - the 8 border rows (camera border 0-5, rhs 6, unused 7) are E[e*ES + col];
 *  - the panel buffer is P[h][r][4]: column cc = 4h + c4 of panel row r, so that
 *    one DMMA operand fragment (row T/4, k = T%4) is 32 consecutive doubles. */
DS_FN int pidx(int cc, int r, int HS) { return (cc >> 2) * HS + r * 4 + (cc & 3); }

/* ------------------------------------------------------------ prologue -- */

/* returns 0 or an error code (uniform over the team) */
template <bool XS>
DS_FN_NOINLINE int prologue(const Team team, Ctx &cx) {
  Ctx &c = ctx_ref();
  (void)cx;
  const PlanView &pl = c.pl;
  const ProbView &pb = c.pb;
  const int n = pl.n_nodes, nf = pl.n_facets, M = pb.n_matches;
  double *x = xvec<XS>(c);
  double *red = sm_base() + c.sl.red;

  DS_FOR(i, pl.Dn_pad) x[i] = i < pl.Dn ? pb.node_xyz[i] : 0.0;
  DS_FOR(i, pl.Dn_pad + 8) dxvec<XS>(c)[i] = 0.0;
  DS_FOR(i, 192) sm_base()[c.sl.invL + i] = 0.0;
  DS_FOR(i, 2 * n) viewed_ptr(c)[i] = 0; /* viewed + freev are contiguous */
  DS_FOR(f, nf) c.ws.fcnt[f] = 0;
  if (team.tid == 0) {
    Pose P;
    pose_from_Tcw(pb.Tcw, P);
    double *ps = sm_base() + c.sl.pose;
    for (int k = 0; k < 4; k++) ps[k] = P.q[k];
    for (int k = 0; k < 3; k++) ps[4 + k] = P.t[k];
  }
  /* zero the band workspace; padding rows are identity */
  {
    const int tot = pl.Dn_pad * pl.ld;
    DS_FOR(i, tot) c.ws.Hb[i] = 0.0;
    DS_FOR(i, 8 * pl.ES) c.ws.Cg[i] = 0.0;
  }
  team.sync();
  DS_FOR(i, pl.Dn_pad - pl.Dn) c.ws.Hb[(pl.Dn + i) * pl.ld + pl.bwE] = 1.0;

  /* facet of every match (DefMapPoint::getFacet) + viewed nodes
   * (DefOptimizer.cc:326-335) */
  int bad = 0;
  DS_FOR(m, M) {
    int v[3] = {pb.match_nodes[3 * m], pb.match_nodes[3 * m + 1], pb.match_nodes[3 * m + 2]};
    int code = -1;
    if (v[0] >= 0 && v[0] < n && v[1] >= 0 && v[1] < n && v[2] >= 0 && v[2] < n) {
      for (int k = pl.nf_ptr[v[0]]; k < pl.nf_ptr[v[0] + 1] && code < 0; k++) {
        const int f = pl.nf_ent[k] >> 2;
        const int *fv = &pl.facets[3 * f];
        int s[3];
        bool ok = true;
        for (int a = 0; a < 3; a++) {
          s[a] = (v[a] == fv[0]) ? 0 : (v[a] == fv[1]) ? 1 : (v[a] == fv[2]) ? 2 : -1;
          ok = ok && s[a] >= 0;
        }
        if (ok && s[0] != s[1] && s[1] != s[2] && s[0] != s[2]) code = (f << 6) | s[0] | (s[1] << 2) | (s[2] << 4);
      }
    }
    c.ws.mfac[m] = code;
    if (code < 0) bad++;
    else {
      atomic_inc_int(&c.ws.fcnt[code >> 6]);
      viewed_ptr(c)[v[0]] = 1; viewed_ptr(c)[v[1]] = 1; viewed_ptr(c)[v[2]] = 1;
    }
  }
  bad = team_sum_int(team, bad, red);
  if (bad > 0) return DEFSLAM_EBADARG;

  /* OptLap = Viewed U ring1(Viewed)  (DefOptimizer.cc:384-406, quirk C3) */
  DS_FOR(v, n) {
    int fr = viewed_ptr(c)[v];
    if (pb.variant == 1) fr = 1; /* every node is free (DefOptimizer.cc:615-620) */
    else if (!fr && pb.layers >= 1)
      for (int k = pl.nbr_ptr[v]; k < pl.nbr_ptr[v + 1]; k++) fr |= viewed_ptr(c)[pl.nbr_idx[k]];
    freev_ptr(c)[v] = (uint8_t)fr;
  }
  if (team.tid == 0) { /* exclusive scan of the facet histogram */
    int s = 0;
    for (int f = 0; f < nf; f++) { c.ws.fptr[f] = s; s += c.ws.fcnt[f]; }
    c.ws.fptr[nf] = s;
  }
  team.sync();
  int cnt_v = 0, cnt_o = 0, cnt_s = 0;
  DS_FOR(v, n) { cnt_v += viewed_ptr(c)[v]; cnt_o += freev_ptr(c)[v]; }
  DS_FOR(e, pl.n_edges) cnt_s += (freev_ptr(c)[pl.edge_ab[2 * e]] | freev_ptr(c)[pl.edge_ab[2 * e + 1]]);
  DS_FOR(f, nf) c.ws.fcnt[f] = 0;
  const int n_viewed = team_sum_int(team, cnt_v, red);
  const int n_optlap = team_sum_int(team, cnt_o, red);
  const int n_str = team_sum_int(team, cnt_s, red);
  DS_FOR(m, M) {
    const int f = c.ws.mfac[m] >> 6;
    c.ws.mperm[c.ws.fptr[f] + atomic_inc_int(&c.ws.fcnt[f])] = m;
  }
  team.sync();
  DS_FOR(f, nf) { /* fixed order inside a facet: ascending match index */
    const int b = c.ws.fptr[f], e = c.ws.fptr[f + 1];
    for (int i = b + 1; i < e; i++) {
      const int key = c.ws.mperm[i];
      int j = i - 1;
      while (j >= b && c.ws.mperm[j] > key) { c.ws.mperm[j + 1] = c.ws.mperm[j]; j--; }
      c.ws.mperm[j + 1] = key;
    }
  }
  team.sync();
  /* per-match records in facet-grouped order: the error passes (once per LM iteration and once per trial) then read
   * one contiguous 64-byte record per match instead of the permutation and, through it, four of the caller's arrays
   * (two dependent round trips of ~3 k cycles each under this kernel's load) */
  DS_FOR(sidx, M) {
    const int m = c.ws.mperm[sidx];
    double *r = c.ws.mrec + 8 * (size_t)sidx;
    r[0] = pb.match_bary[3 * m]; r[1] = pb.match_bary[3 * m + 1]; r[2] = pb.match_bary[3 * m + 2];
    int *ri = (int *)(r + 3);
    float *rf = (float *)(r + 3);
    ri[0] = pb.match_nodes[3 * m]; ri[1] = pb.match_nodes[3 * m + 1]; ri[2] = pb.match_nodes[3 * m + 2];
    rf[3] = pb.match_isig[m]; rf[4] = pb.match_uv[2 * m]; rf[5] = pb.match_uv[2 * m + 1];
    ri[6] = c.ws.mfac[m]; ri[7] = m;
  }
  team.sync();

  /* information matrices (DefOptimizer.cc:339-340,376-378,458,499); the context
   * is shared by the CTA: one writer */
  if (team.tid == 0) {
    c.n_viewed = n_viewed; c.n_optlap = n_optlap; c.n_str = n_str;
    c.inv_n = 1.0; /* division by N is done per match as float / int like the reference */
    c.info_ref = pb.reg_temp / pow(pl.median_len, 2);
    c.info_curv = n_optlap > 0 ? pb.reg_lap / (double)n_optlap : 0.0;
    c.info_str = n_str > 0 ? pb.reg_inex / (double)n_str : 0.0;
    c.info_uniform = 0.0;
    float deltaMono = (float)sqrt(5.991);
    if (pb.variant == 1) { /* DefOptimizer.cc:582-837 */
      c.info_ref = 0.0;                                              /* temporal edges never added (:688) */
      c.info_curv = n_viewed > 0 ? pb.reg_lap / (double)n_viewed : 0.0; /* OptLap = ViewedNodes (:693,:755) */
      c.info_uniform = 1.0 / (double)(M > 0 ? M : 1);                              /* Identity / double(matches.size()) (:655) */
      deltaMono = 0.5f;                                              /* :625 */
    }
    c.hub_delta = (double)deltaMono;
    c.hub_dsqr = (double)(float)(c.hub_delta * c.hub_delta);  // RobustKernelHuber::dsqr is a float (robust_kernel_impl.h:84)
  }
  team.sync();
  return 0;
}

/* -------------------------------------------------- errors and chi2 ---- */

DS_FN void load_pose(const double *ps, Pose &P) {
  for (int k = 0; k < 4; k++) P.q[k] = ps[k];
  for (int k = 0; k < 3; k++) P.t[k] = ps[4 + k];
}

/* reprojection error of one match at (x, P)  -- sft_types.h:102-133 */
DS_FN void reproj_error(const Ctx &c, const double *x, const Pose &P, int m, double e[2], double Pc[3]) {
  const ProbView &pb = c.pb;
  const int v0 = pb.match_nodes[3 * m], v1 = pb.match_nodes[3 * m + 1], v2 = pb.match_nodes[3 * m + 2];
  const double b0 = pb.match_bary[3 * m], b1 = pb.match_bary[3 * m + 1], b2 = pb.match_bary[3 * m + 2];
  double Pw[3];
  for (int k = 0; k < 3; k++) Pw[k] = b0 * x[3 * v0 + k] + b1 * x[3 * v1 + k] + b2 * x[3 * v2 + k];
  pose_map(P, Pw, Pc);
  const double u = Pc[0] / Pc[2] * pb.fx + pb.cx;
  const double v = Pc[1] / Pc[2] * pb.fy + pb.cy;
  e[0] = (double)pb.match_uv[2 * m] - u;
  e[1] = (double)pb.match_uv[2 * m + 1] - v;
}

DS_FN double match_info(const Ctx &c, int m) {
  return c.info_uniform > 0.0 ? c.info_uniform : (double)c.pb.match_isig[m] / (double)c.pb.n_kp;
}

/* the same from the facet-grouped record of match slot s (prologue): b[3], v[3], 1/sigma^2, observation, the facet
 * code of the match (mfac) and its index in the caller's order */
struct MatchRec {
  double b[3];
  int v[3];
  float isig, u, w;
  int code, m;
};
DS_FN MatchRec match_record(const Ctx &c, int s) {
  MatchRec r;
  const double *p = c.ws.mrec + 8 * (size_t)s;
#if DS_CUDA
  const dbl2 q0 = *(const dbl2 *)p, q1 = *(const dbl2 *)(p + 2), q2 = *(const dbl2 *)(p + 4), q3 = *(const dbl2 *)(p + 6);
  r.b[0] = q0.x; r.b[1] = q0.y; r.b[2] = q1.x;
  r.v[0] = (int)((unsigned long long)__double_as_longlong(q1.y) & 0xffffffffu);
  r.v[1] = (int)((unsigned long long)__double_as_longlong(q1.y) >> 32);
  r.v[2] = (int)((unsigned long long)__double_as_longlong(q2.x) & 0xffffffffu);
  r.isig = __int_as_float((int)((unsigned long long)__double_as_longlong(q2.x) >> 32));
  r.u = __int_as_float((int)((unsigned long long)__double_as_longlong(q2.y) & 0xffffffffu));
  r.w = __int_as_float((int)((unsigned long long)__double_as_longlong(q2.y) >> 32));
  r.code = (int)((unsigned long long)__double_as_longlong(q3.x) & 0xffffffffu);
  r.m = (int)((unsigned long long)__double_as_longlong(q3.x) >> 32);
#else
  const int *pi = (const int *)(p + 3);
  const float *pf = (const float *)(p + 3);
  r.b[0] = p[0]; r.b[1] = p[1]; r.b[2] = p[2];
  r.v[0] = pi[0]; r.v[1] = pi[1]; r.v[2] = pi[2];
  r.isig = pf[3]; r.u = pf[4]; r.w = pf[5];
  r.code = pi[6]; r.m = pi[7];
#endif
  return r;
}
DS_FN void reproj_error_rec(const Ctx &c, const double *x, const Pose &P, const MatchRec &r, double e[2], double Pc[3]) {
  const ProbView &pb = c.pb;
  double Pw[3];
  for (int k = 0; k < 3; k++) Pw[k] = r.b[0] * x[3 * r.v[0] + k] + r.b[1] * x[3 * r.v[1] + k] + r.b[2] * x[3 * r.v[2] + k];
  pose_map(P, Pw, Pc);
  const double u = Pc[0] / Pc[2] * pb.fx + pb.cx;
  const double v = Pc[1] / Pc[2] * pb.fy + pb.cy;
  e[0] = (double)r.u - u;
  e[1] = (double)r.w - v;
}
DS_FN double match_info_rec(const Ctx &c, const MatchRec &r) {
  return c.info_uniform > 0.0 ? c.info_uniform : (double)r.isig / (double)c.pb.n_kp;
}

/* curvature residual of centre i: delta = x_i - sum(w x_j)/W  (sft_types.h:257-291) */
DS_FN double curv_residual(const Ctx &c, const double *x, int i, double d[3], double &nrm) {
  const PlanView &pl = c.pl;
  double a0 = 0, a1 = 0, a2 = 0;
  for (int k = pl.nbr_ptr[i]; k < pl.nbr_ptr[i + 1]; k++) {
    const int j = pl.nbr_idx[k];
    const double w = pl.nbr_w[k];
    a0 = a0 + w * x[3 * j]; a1 = a1 + w * x[3 * j + 1]; a2 = a2 + w * x[3 * j + 2];
  }
  const double W = pl.sum_w[i];
  d[0] = x[3 * i] - a0 / W; d[1] = x[3 * i + 1] - a1 / W; d[2] = x[3 * i + 2] - a2 / W;
  nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  return nrm - pl.kappa0[i];
}

/* SparseOptimizer::computeActiveErrors + activeRobustChi2 at state (x, pose).
 * store=true additionally fills everything build_system() gathers from:
 * per-match scratch S, per-node A / centre / edge quantities (overlaid on the
 * window region of shared memory). */
template <bool XS, bool STORE>
DS_FN_NOINLINE double eval_state(const Team team, Ctx &cx) {
  constexpr bool store = STORE;
  Ctx &c = ctx_ref();
  (void)cx;
  const double *x = xvec<XS>(c), *ps = sm_base() + c.sl.pose;
  const PlanView &pl = c.pl;
  const ProbView &pb = c.pb;
  const int n = pl.n_nodes, ne = pl.n_edges, M = pb.n_matches;
  Pose P;
  load_pose(ps, P);
  double *A = sm_base() + c.sl.W;      /* [6n] */
  double *cd = A + 6 * n;         /* [3n] unit Laplacian direction */
  double *cg = cd + 3 * n;        /* [n]  info_curv * S_i (0 if inactive) */
  double *cr = cg + n;            /* [n]  |delta| - kappa0 */
  double *eu = cr + n;            /* [3ne] J_a of the stretch edge */
  double *er = eu + 3 * ne;       /* [ne] stretch residual */
  double *es = er + ne;           /* [ne] info_str if active else 0 */
  double chi = 0.0;
  double R[9];
  if (store) quat_to_R(P.q, R);
  /* camera block of J^T W J (upper triangle, 21) and J_c^T W e (6): no per-facet resolution is
   * needed, so every thread sums its own matches here, where the Jacobian is in registers */
  double cam[STORE ? 27 : 1];
  if (store) {
#pragma unroll
    for (int k = 0; k < 27; k++) cam[k] = 0.0;
  }

  /* reprojection edges, in facet-grouped order */
  DS_FOR(s, M) {
    const MatchRec mr = match_record(c, s);
    double e[2], Pc[3];
    reproj_error_rec(c, x, P, mr, e, Pc);
    const double info = match_info_rec(c, mr);
    const double c2 = e[0] * info * e[0] + e[1] * info * e[1];
    double rho0, rho1; /* RobustKernelHuber::robustify robust_kernel_impl.cpp:78-91 */
    if (c2 <= c.hub_dsqr) { rho0 = c2; rho1 = 1.0; }
    else { const double sq = sqrt(c2); rho0 = 2 * sq * c.hub_delta - c.hub_dsqr; rho1 = c.hub_delta / sq; }
    chi += rho0;
    if (store) {
      double *S = c.ws.S;
      /* camera Jacobian at the interpolated point, from the node images like
       * the reference (xyz = sum b_k (R x_k + t))  sft_types.h:151-174 */
      const int v[3] = {mr.v[0], mr.v[1], mr.v[2]};
      const double b[3] = {mr.b[0], mr.b[1], mr.b[2]};
      double xk[3][3];
      for (int k = 0; k < 3; k++) pose_map(P, &x[3 * v[k]], xk[k]);
      const double X = xk[0][0] * b[0] + xk[1][0] * b[1] + xk[2][0] * b[2];
      const double Y = xk[0][1] * b[0] + xk[1][1] * b[1] + xk[2][1] * b[2];
      const double Z = xk[0][2] * b[0] + xk[1][2] * b[1] + xk[2][2] * b[2];
      const double Z2 = Z * Z, fx = pb.fx, fy = pb.fy;
      S[0 * M + s] = e[0];
      S[1 * M + s] = e[1];
      S[2 * M + s] = rho1 * info;
      S[3 * M + s] = X * Y / Z2 * fx;
      S[4 * M + s] = -(1 + (X * X / Z2)) * fx;
      S[5 * M + s] = Y / Z * fx;
      S[6 * M + s] = -1. / Z * fx;
      S[7 * M + s] = 0;
      S[8 * M + s] = X / Z2 * fx;
      S[9 * M + s] = (1 + Y * Y / Z2) * fy;
      S[10 * M + s] = -X * Y / Z2 * fy;
      S[11 * M + s] = -X / Z * fy;
      S[12 * M + s] = 0;
      S[13 * M + s] = -1. / Z * fy;
      S[14 * M + s] = Y / Z2 * fy;
      const int code = mr.code;
      S[(15 + (code & 3)) * M + s] = b[0];
      S[(15 + ((code >> 2) & 3)) * M + s] = b[1];
      S[(15 + ((code >> 4) & 3)) * M + s] = b[2];
      {
        const double w = rho1 * info;
        const double J[12] = {X * Y / Z2 * fx, -(1 + (X * X / Z2)) * fx, Y / Z * fx, -1. / Z * fx, 0, X / Z2 * fx,
                              (1 + Y * Y / Z2) * fy, -X * Y / Z2 * fy, -X / Z * fy, 0, -1. / Z * fy, Y / Z2 * fy};
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int q = r; q < 6; q++) cam[STORE ? k++ : 0] += w * (J[r] * J[q] + J[6 + r] * J[6 + q]);
#pragma unroll
        for (int r = 0; r < 6; r++) cam[STORE ? 21 + r : 0] += w * (J[r] * e[0] + J[6 + r] * e[1]);
      }
    }
  }
  if (store) {
    /* fixed-shape reduction: lanes by shuffle, then one partial per warp in the panel buffer
     * (free outside factor_solve); build_system adds the warps in order */
    double *part = sm_base() + c.sl.P;
#if DS_CUDA
    const int warp = team.tid >> 5, nwarp = team.nthr >> 5;
#pragma unroll
    for (int k = 0; k < 27; k++) {
      double v = cam[STORE ? k : 0];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if ((team.tid & 31) == 0) part[k * nwarp + warp] = v;
    }
#else
    for (int k = 0; k < 27; k++) part[k] = cam[STORE ? k : 0];
#endif
  }
  /* temporal (EdgesReference sft_types.h:403-408), curvature, per-node A */
  DS_FOR(v, n) {
    if (viewed_ptr(c)[v]) {
      const double e0 = x[3 * v] - pl.rest[3 * v], e1 = x[3 * v + 1] - pl.rest[3 * v + 1],
                   e2 = x[3 * v + 2] - pl.rest[3 * v + 2];
      chi += (e0 * e0 + e1 * e1 + e2 * e2) * c.info_ref;
    }
    /* curvature centres: OptLap, which is the free set -- or the viewed nodes in the matches-given overload */
    const bool in_lap = pb.variant == 1 ? viewed_ptr(c)[v] != 0 : freev_ptr(c)[v] != 0;
    const bool active = in_lap && !pl.boundary[v] && (pl.nbr_ptr[v + 1] > pl.nbr_ptr[v]);
    if (active) {
      double d[3], nrm;
      const double r = curv_residual(c, x, v, d, nrm);
      /* deg(v) copies of the residual, each divided by its edge length -- or all by the caller's lenghtEdge_ (C8) */
      const double g = c.info_curv * (pb.variant == 1 ? (double)(pl.nbr_ptr[v + 1] - pl.nbr_ptr[v]) / (pb.curv_len * pb.curv_len)
                                                       : pl.inv_len2[v]);
      chi += r * r * g;
      if (store) {
        const double inv = nrm < 1E-15 ? 0.0 : 1.0 / nrm;
        cd[3 * v] = d[0] * inv; cd[3 * v + 1] = d[1] * inv; cd[3 * v + 2] = d[2] * inv;
        cg[v] = g; cr[v] = r;
      }
    } else if (store) {
      cd[3 * v] = cd[3 * v + 1] = cd[3 * v + 2] = 0.0; cg[v] = 0.0; cr[v] = 0.0;
    }
    if (store) { /* A_v = -(1/z) [fx 0 -x/z fx; 0 fy -y/z fy] R   sft_types.h:176-190 */
      double p[3];
      pose_map(P, &x[3 * v], p);
      const double t02 = -p[0] / p[2] * pb.fx, t12 = -p[1] / p[2] * pb.fy, mz = -1. / p[2];
      for (int k = 0; k < 3; k++) {
        A[6 * v + k] = mz * (pb.fx * R[k] + t02 * R[6 + k]);
        A[6 * v + 3 + k] = mz * (pb.fy * R[3 + k] + t12 * R[6 + k]);
      }
    }
  }
  /* stretch (EdgesStreching sft_types.h:353-378) */
  DS_FOR(e, ne) {
    const int a = pl.edge_ab[2 * e], b = pl.edge_ab[2 * e + 1];
    const bool active = freev_ptr(c)[a] || freev_ptr(c)[b];
    if (active) {
      const double d0 = x[3 * a] - x[3 * b], d1 = x[3 * a + 1] - x[3 * b + 1], d2 = x[3 * a + 2] - x[3 * b + 2];
      const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
      const double l0 = pl.edge_len0[e];
      const double r = nrm * (1.0 / l0) - 1.0;
      chi += r * c.info_str * r;
      if (store) {
        const double ddo = 1.0 / (nrm * l0);
        eu[3 * e] = d0 * ddo; eu[3 * e + 1] = d1 * ddo; eu[3 * e + 2] = d2 * ddo;
        er[e] = r; es[e] = c.info_str;
      }
    } else if (store) {
      eu[3 * e] = eu[3 * e + 1] = eu[3 * e + 2] = 0.0; er[e] = 0.0; es[e] = 0.0;
    }
  }
  return team_sum(team, chi, sm_base() + c.sl.red);
}

/* ------------------------------------------------ normal equations ----- */

/* BlockSolver::buildSystem for this graph.  Needs eval_state(store=true) at the
 * same state.  Produces Hb (band), Cg rows 0-5 (camera border), Cg row 6 (b_n),
 * Hcc/bc (shared).  Returns max |diag| over the free variables
 * (computeLambdaInit, optimization_algorithm_levenberg.cpp:166-180). */
template <bool XS>
DS_FN_NOINLINE double build_system(const Team team, Ctx &cx) {
  Ctx &c = ctx_ref();
  (void)cx;
  const PlanView &pl = c.pl;
  const int n = pl.n_nodes, ne = pl.n_edges, nf = pl.n_facets, M = c.pb.n_matches;
  const int bwE = pl.bwE, ld = pl.ld, ES = pl.ES;
  const double *A = sm_base() + c.sl.W;
  const double *cd = A + 6 * n, *cg = cd + 3 * n, *cr = cg + n, *eu = cr + n, *er = eu + 3 * ne, *es = er + ne;
  double *F = c.ws.F;
  const double *S = c.ws.S;
  double *Hcc = sm_base() + c.sl.Hcc;
  team.sync();
  DS_PROF_T0(bt0);
  DS_PROF_INC(PF_X_BUILDS, team.tid == 0);

  /* (0) camera-camera block and b_c: the per-warp partial sums eval_state left in the panel
   * buffer, added in warp order -- first, because the staging area of (1) covers that buffer */
  {
    const double *part = sm_base() + c.sl.P;
#if DS_CUDA
    const int nwarp = team.nthr >> 5;
#else
    const int nwarp = 1;
#endif
    DS_FOR(k, 27) {
      double s = 0.0;
      for (int w = 0; w < nwarp; w++) s += part[k * nwarp + w];
      if (k < 21) {
        int r = 0, rem = k;
        while (rem >= 6 - r) { rem -= 6 - r; r++; }
        const int q = r + rem;
        Hcc[r * 6 + q] = s; Hcc[q * 6 + r] = s;
      } else {
        Hcc[36 + (k - 21)] = -s;
      }
    }
  }

  /* (1) per-facet sums.  item = (facet, group).  The per-match scratch of a run of facets is
   * first staged into the shared memory that is idle during assembly (rest of the window, the
   * border rows, the panel buffer) by all threads -- coalesced, every load independent -- so the
   * sums read shared memory instead of paying one L2/HBM round trip per match. */
  {
    const int asz = asm_scratch_doubles(n, ne);
    double *stage = sm_base() + c.sl.W + asz;
    const int psz = panel_doubles(pl.bwp);
    /* the tail of the area holds the chunk's facet offsets (nf + 1 ints) */
    const int avail = (c.sl.P + psz) - (c.sl.W + asz) - (nf + 2) / 2 - 1;
    const int cap = avail > 0 ? avail / NMSCR : 0; /* matches per chunk */
    int *fps = (int *)(stage + (avail > 0 ? avail : 0));
    int fb = 0;
    DS_PROF_LOCALS(fsacc, 4);
    DS_PROF_T0M(fst);
    while (fb < nf) {
      const int sb0 = c.ws.fptr[fb];
      int fe = nf, s1 = M;
      if (M - sb0 > cap) {
        fe = cap > 0 ? c.ws.mfac[c.ws.mperm[sb0 + cap]] >> 6 : fb; /* first facet that does not fit entirely */
        s1 = c.ws.fptr[fe];
      }
      const double *SS;
      int sst;
      if (fe > fb) { /* stage matches sb0..s1 */
        const int cnt = s1 - sb0;
        team.sync(); /* the previous chunk has been consumed */
        DS_PROF_LAP(fsacc, 0, fst);
        /* plane by plane, all 18 loads of a thread in flight together */
        DS_FOR(j, cnt) {
          double v[NMSCR];
#pragma unroll
          for (int pln = 0; pln < NMSCR; pln++) v[pln] = S[pln * M + sb0 + j];
#pragma unroll
          for (int pln = 0; pln < NMSCR; pln++) stage[pln * cap + j] = v[pln];
        }
        /* facet offsets of the chunk, relative to it (int view of the tail of the staging area) */
        DS_FOR(i, fe - fb + 1) fps[i] = c.ws.fptr[fb + i] - sb0;
        DS_PROF_LAP(fsacc, 1, fst);
        team.sync();
        DS_PROF_LAP(fsacc, 2, fst);
        SS = stage; sst = cap;
      } else {      /* a single facet with more matches than the staging area holds: from global */
        fe = fb + 1;
        SS = S; sst = M;
      }
      const int nfc = fe - fb;
      /* group-major: the lanes of a warp run the same group on consecutive facets (no divergence
       * between the four code paths, neighbouring rows of the scratch) */
      DS_FOR(it, nfc * 4) {
        const int g = it / nfc, f = fb + (it - g * nfc);
        int sb, se;
        if (SS == stage) { sb = fps[f - fb]; se = fps[f - fb + 1]; }
        else { sb = c.ws.fptr[f]; se = c.ws.fptr[f + 1]; }
    /* Two matches per trip: the loads of both are independent, so a facet with n matches costs
     * ceil(n/2) memory round trips.  The second slot of an odd tail re-reads the last match with
     * weight 0 (adds +0: the sums are the same as match-by-match). */
    if (g == 0) {
      double a[12];
#pragma unroll
      for (int k = 0; k < 12; k++) a[k] = 0.0;
      for (int s0 = sb; s0 < se; s0 += 2) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const bool ok = s0 + u < se;
          const int s = ok ? s0 + u : s0;
          const double w = ok ? SS[2 * sst + s] : 0.0, e0 = SS[s], e1 = SS[sst + s];
          const double b0 = SS[15 * sst + s], b1 = SS[16 * sst + s], b2 = SS[17 * sst + s];
          a[0] += w * b0 * b0; a[1] += w * b0 * b1; a[2] += w * b0 * b2;
          a[3] += w * b1 * b1; a[4] += w * b1 * b2; a[5] += w * b2 * b2;
          a[6] += w * b0 * e0; a[7] += w * b0 * e1;
          a[8] += w * b1 * e0; a[9] += w * b1 * e1;
          a[10] += w * b2 * e0; a[11] += w * b2 * e1;
        }
      }
#pragma unroll
      for (int k = 0; k < 12; k++) F[k * nf + f] = a[k];
    } else {
      double a[12];
#pragma unroll
      for (int k = 0; k < 12; k++) a[k] = 0.0;
      for (int s0 = sb; s0 < se; s0 += 2) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const bool ok = s0 + u < se;
          const int s = ok ? s0 + u : s0;
          const double wb = (ok ? SS[2 * sst + s] : 0.0) * SS[(14 + g) * sst + s];
#pragma unroll
          for (int k = 0; k < 12; k++) a[k] += wb * SS[(3 + k) * sst + s];
        }
      }
#pragma unroll
      for (int k = 0; k < 12; k++) F[(12 * g + k) * nf + f] = a[k];
    }
      }
      DS_PROF_LAP(fsacc, 3, fst);
      fb = fe;
    }
    DS_PROF_FLUSH(fsacc, 4, PF_X_FS, team.tid == 0);
  }
  team.sync();
  DS_PROF_ADD(PF_X_BUILD + 0, bt0, team.tid == 0);
  DS_PROF_T0(bt1);

  DS_PROF_ADD(PF_X_BUILD + 1, bt1, team.tid == 0);
  DS_PROF_T0(bt2);
  /* (3) node-node blocks: gather */
  double maxd = 0.0;
  DS_FOR(bi, pl.n_blk) {
    /* packed plan record: everything the block needs from the plan in one 32-byte read */
    const int *hd = &pl.blk_hdr[8 * bi];
#if DS_CUDA
    const int4 h0 = *(const int4 *)hd, h1 = *(const int4 *)(hd + 4);
    const int p = h0.x, q = h0.y, edge = h0.z, c0 = h0.w, nctr = h1.x, nfac = h1.y, f0 = h1.z, f1 = h1.w;
#else
    const int p = hd[0], q = hd[1], edge = hd[2], c0 = hd[3], nctr = hd[4], nfac = hd[5], f0 = hd[6], f1 = hd[7];
#endif
    double h[9];
    for (int k = 0; k < 9; k++) h[k] = 0.0;
    const bool fp = freev_ptr(c)[p], fq = freev_ptr(c)[q];
    if (fp && fq) {
      /* reprojection: (sum over shared facets of sum_m w b_p b_q) * A_p^T A_q */
      double beta = 0.0;
      if (nfac <= 2) {
        if (nfac > 0) beta += F[f0];
        if (nfac > 1) beta += F[f1];
      } else {
        for (int k = 0; k < nfac; k++) beta += F[pl.blk_fidx[f0 + k]];
      }
      const double *Ap = &A[6 * p], *Aq = &A[6 * q];
      for (int r = 0; r < 3; r++)
        for (int s = 0; s < 3; s++) h[3 * r + s] = beta * (Ap[r] * Aq[s] + Ap[3 + r] * Aq[3 + s]);
      /* curvature: g_i c_ip c_iq d_i d_i^T.  Four centres at a time: the centre indices and the
       * two coefficients of a batch are independent loads; accumulation order unchanged. */
      {
        const int c1 = c0 + nctr;
        for (int kb4 = c0; kb4 < c1; kb4 += 4) {
          int ci[4];
          double cpv[4], cqv[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int k = kb4 + u < c1 ? kb4 + u : c1 - 1;
            ci[u] = pl.blk_ctr_i[k];
            cpv[u] = pl.blk_ctr_pq[2 * k];
            cqv[u] = pl.blk_ctr_pq[2 * k + 1];
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (kb4 + u >= c1) break;
            const int i = ci[u];
            const double g = cg[i];
            if (g == 0.0) continue;
            const double gc = g * cpv[u] * cqv[u];
            const double *d = &cd[3 * i];
            for (int r = 0; r < 3; r++)
              for (int s = 0; s < 3; s++) h[3 * r + s] += gc * d[r] * d[s];
          }
        }
      }
      /* stretch */
      if (p == q) {
        for (int k = pl.ne_ptr[p]; k < pl.ne_ptr[p + 1]; k++) {
          const int e = pl.ne_ent[k] >> 1;
          const double w = es[e];
          const double *u = &eu[3 * e];
          for (int r = 0; r < 3; r++)
            for (int s = 0; s < 3; s++) h[3 * r + s] += w * u[r] * u[s];
        }
        if (viewed_ptr(c)[p]) { h[0] += c.info_ref; h[4] += c.info_ref; h[8] += c.info_ref; }
      } else if (edge >= 0) {
        const int e = edge;
        const double w = es[e];
        const double *u = &eu[3 * e];
        for (int r = 0; r < 3; r++)
          for (int s = 0; s < 3; s++) h[3 * r + s] -= w * u[r] * u[s];
      }
    } else if (p == q) {
      h[0] = h[4] = h[8] = 1.0; /* fixed vertex: identity rows */
    }
    /* rows of p (i), cols of q (j <= i) */
    for (int r = 0; r < 3; r++) {
      const int i = 3 * p + r;
      for (int s = 0; s < 3; s++) {
        const int j = 3 * q + s;
        if (j > i) continue;
        st_stream(&c.ws.Hb[i * ld + (j - i + bwE)], h[3 * r + s]);
      }
      if (p == q && fp) maxd = fmax(maxd, fabs(h[4 * r]));
    }
  }

  DS_PROF_ADD(PF_X_BUILD + 2, bt2, team.tid == 0);
  DS_PROF_T0(bt3);
  /* (4) per-node gradient b_n and camera border C */
  DS_FOR(p, n) {
    double b[3] = {0, 0, 0}, C[18];
    for (int k = 0; k < 18; k++) C[k] = 0.0;
    if (freev_ptr(c)[p]) {
      const double *Ap = &A[6 * p];
      double be0 = 0, be1 = 0, bj[12];
      for (int k = 0; k < 12; k++) bj[k] = 0.0;
      for (int k = pl.nf_ptr[p]; k < pl.nf_ptr[p + 1]; k++) {
        const int f = pl.nf_ent[k] >> 2, sl = pl.nf_ent[k] & 3;
        be0 += F[(6 + 2 * sl) * nf + f];
        be1 += F[(7 + 2 * sl) * nf + f];
        for (int t = 0; t < 12; t++) bj[t] += F[(12 * (sl + 1) + t) * nf + f];
      }
      for (int r = 0; r < 3; r++) b[r] = -(Ap[r] * be0 + Ap[3 + r] * be1);
      /* C[a][r] = sum_rows bj[row][a] * Ap[row][r] */
      for (int a = 0; a < 6; a++)
        for (int r = 0; r < 3; r++) C[3 * a + r] = bj[a] * Ap[r] + bj[6 + a] * Ap[3 + r];
      if (viewed_ptr(c)[p])
        for (int r = 0; r < 3; r++) b[r] -= c.info_ref * (xvec<XS>(c)[3 * p + r] - pl.rest[3 * p + r]);
      for (int k = pl.nc_ptr[p]; k < pl.nc_ptr[p + 1]; k++) {
        const int i = pl.nc_ent[2 * k], ip = pl.nc_ent[2 * k + 1];
        const double g = cg[i];
        if (g == 0.0) continue;
        const double cp = ip < 0 ? 1.0 : -pl.nbr_c[ip];
        const double s = g * cp * cr[i];
        for (int r = 0; r < 3; r++) b[r] -= s * cd[3 * i + r];
      }
      for (int k = pl.ne_ptr[p]; k < pl.ne_ptr[p + 1]; k++) {
        const int ent = pl.ne_ent[k], e = ent >> 1;
        const double s = (ent & 1) ? es[e] * er[e] : -es[e] * er[e];
        for (int r = 0; r < 3; r++) b[r] += s * eu[3 * e + r];
      }
    }
    for (int r = 0; r < 3; r++) {
      c.ws.Cg[6 * ES + 3 * p + r] = b[r];
      for (int a = 0; a < 6; a++) c.ws.Cg[a * ES + 3 * p + r] = C[3 * a + r];
    }
  }
  DS_PROF_ADD(PF_X_BUILD + 3, bt3, team.tid == 0);
  DS_PROF_T0(bt4);
  maxd = team_max(team, maxd, sm_base() + c.sl.red); /* also a barrier: Hcc complete */
  DS_PROF_ADD(PF_X_BUILD + 4, bt4, team.tid == 0);
  for (int k = 0; k < 6; k++) maxd = fmax(maxd, fabs(Hcc[7 * k]));
  return maxd;
}

/* ------------------------------------------- banded Cholesky + solve --- */

/* Cholesky of one NB x NB diagonal block, by warp 0.
 * in : the block's lower triangle in the window (lambda is added to the diagonal)
 * out: L_kk written back to the window with the RECIPROCAL of each diagonal
 *      entry on the diagonal (only 1/L_ii is ever needed afterwards), and a
 *      dense copy in Lkk[8][8] (shared) for the panel solve; *flag set when a
 *      pivot is not > 0.
 * Every lane holds the whole block in registers and runs the same elimination:
 * the critical path is 8 x (rsqrt + mul + fma) with no shuffle or shared-memory
 * round trip on it; lanes only split the stores. */
DS_FN double ds_rsqrt(double d) {
#if DS_CUDA
#if DS_FAST_RSQRT
  /* branch-free: hardware seed (MUFU.RSQ64H, ~2^-22) and one third-order step
   * y1 = y0 (1 + e/2 + 3 e^2/8), e = 1 - d y0^2  (relative error ~ e^3 -> below 2^-60).
   * The library rsqrt() carries a slow-path call per use, which ends the basic block and stops the
   * scheduler from overlapping the 62-cycle latency with the independent updates of the block.
   * Pivots here are > 0 and far from the denormal range (lambda is added to every diagonal entry);
   * a non-positive pivot is caught by the caller before the value is used. */
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double h = y0 * y0;
  const double e = fma(-d, h, 1.0);
  const double p = fma(0.375, e, 0.5);
  const double t = y0 * e;
  return fma(t, p, y0);
#else
  return rsqrt(d);
#endif
#else
  return 1.0 / sqrt(d);
#endif
}

/* in-register Cholesky of a 4x4 lower triangle (10 entries, row-major packed:
 * 00 10 11 20 21 22 30 31 32 33); diagonal entries are replaced by 1/L_ii */
DS_FN bool chol4(double t[10]) {
  bool bad = false;
  double inv;
  bad = bad || !(t[0] > 0.0); inv = ds_rsqrt(t[0]); t[0] = inv;
  t[1] *= inv; t[3] *= inv; t[6] *= inv;
  t[2] -= t[1] * t[1]; t[4] -= t[3] * t[1]; t[7] -= t[6] * t[1];
  t[5] -= t[3] * t[3]; t[8] -= t[6] * t[3]; t[9] -= t[6] * t[6];
  bad = bad || !(t[2] > 0.0); inv = ds_rsqrt(t[2]); t[2] = inv;
  t[4] *= inv; t[7] *= inv;
  t[5] -= t[4] * t[4]; t[8] -= t[7] * t[4]; t[9] -= t[7] * t[7];
  bad = bad || !(t[5] > 0.0); inv = ds_rsqrt(t[5]); t[5] = inv;
  t[8] *= inv;
  t[9] -= t[8] * t[8];
  bad = bad || !(t[9] > 0.0); inv = ds_rsqrt(t[9]); t[9] = inv;
  return bad;
}

constexpr int ILS = 12; /* row stride of inv(L_kk) in shared memory */

/* entry (i, m), i > m, of the factor held as [A11 0; L21 A22] (packed triangles, reciprocal
 * diagonals); indices are compile-time constants after unrolling */
DS_FN double lreg(const double *A11, const double *L21, const double *A22, int i, int m) {
  return i < 4 ? A11[i * (i + 1) / 2 + m] : (m < 4 ? L21[(i - 4) * 4 + m] : A22[(i - 4) * (i - 3) / 2 + (m - 4)]);
}

/* column j of X = inv(L): s_i = delta_ij - sum_{m<i} L[i][m] X[m];  X[i] = s_i / L_ii.
 * The strict upper part of invL stays zero (prologue). */
DS_FN void inv_column(const double *A11, const double *L21, const double *A22, int j, bool store, double *invL) {
  double sacc[NB];
#pragma unroll
  for (int i = 0; i < NB; i++) sacc[i] = (i == j) ? 1.0 : 0.0;
#pragma unroll
  for (int m = 0; m < NB; m++) {
    const double rd = m < 4 ? A11[m * (m + 1) / 2 + m] : A22[(m - 4) * (m - 3) / 2 + (m - 4)];
    const double xm = sacc[m] * rd;
    if (store && m >= j) invL[m * ILS + j] = xm;
#pragma unroll
    for (int i = m + 1; i < NB; i++) sacc[i] -= lreg(A11, L21, A22, i, m) * xm;
  }
}

/* Cholesky of one NB x NB diagonal block + inverse of its factor, by warp 0.
 * in : the block's lower triangle in the window (lambda is added to the diagonal)
 * out: L_kk written back to the window and to Lkk[8][8] (shared), with the
 *      RECIPROCAL of each diagonal entry on the diagonal (only 1/L_ii is ever
 *      needed afterwards); inv(L_kk) -> invL (shared, row stride ILS);
 *      *flag set when a pivot is not > 0.
 * Phase 1: two-level [A11 0; A21 A22] with 4x4 blocks; every lane runs the same
 * arithmetic on registers, so the dependent chain (8 x rsqrt-mul-fma) has no
 * shuffle or shared-memory hop on it.  Phase 2: lane j builds column j of the
 * inverse by forward substitution from the shared copy. */
DS_FN void diag_factor(const Team team, double *W, int kslot, int Wr, int ld, int bwE, double lambda, double *invL,
                       int *flag) {
  const int lane = team.lane();
  double *rowp[NB]; /* rowp[a][t] = entry (a, a - (bwE - t)); diagonal at rowp[a][bwE] */
#pragma unroll
  for (int a = 0; a < NB; a++) {
    int slot = kslot + a;
    if (slot >= Wr) slot -= Wr;
    rowp[a] = W + slot * ld + bwE - a; /* rowp[a][b] = entry (row a, col b) of the block */
  }
  double A11[10], L21[16], A22[10];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b <= a; b++) A11[a * (a + 1) / 2 + b] = rowp[a][b] + (a == b ? lambda : 0.0);
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) L21[a * 4 + b] = rowp[4 + a][b];
  bool bad = chol4(A11);
  /* L21 = A21 L11^-T, the four rows are independent chains */
#pragma unroll
  for (int a = 0; a < 4; a++) {
    double *x = &L21[a * 4];
    x[0] *= A11[0];
    x[1] = (x[1] - x[0] * A11[1]) * A11[2];
    x[2] = (x[2] - x[0] * A11[3] - x[1] * A11[4]) * A11[5];
    x[3] = (x[3] - x[0] * A11[6] - x[1] * A11[7] - x[2] * A11[8]) * A11[9];
  }
  /* every lane has read A11 / A21 before any lane overwrites them */
  team.warp_sync();
  /* rows 0-3: every lane stores the same values (benign, one wavefront each) */
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < NB; b++) {
      if (b <= a) rowp[a][b] = A11[a * (a + 1) / 2 + b];
    }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b <= a; b++) {
      double v = rowp[4 + a][4 + b] + (a == b ? lambda : 0.0);
#pragma unroll
      for (int m = 0; m < 4; m++) v -= L21[a * 4 + m] * L21[b * 4 + m];
      A22[a * (a + 1) / 2 + b] = v;
    }
  bad = chol4(A22) || bad;
  team.warp_sync(); /* every lane has read A22 */
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < NB; b++) {
      if (b <= 4 + a) rowp[4 + a][b] = b < 4 ? L21[a * 4 + b] : A22[a * (a + 1) / 2 + (b - 4)];
    }
  if (bad && lane == 0) *flag = 1;
  /* Phase 2 from the registers of phase 1 (no shared-memory round trip, no warp barrier): the
   * substitution for column lane%8 is straight-line code the scheduler interleaves with the
   * elimination above, one step behind it. */
#if DS_CUDA
  inv_column(A11, L21, A22, lane & 7, lane < NB, invL);
#else
  for (int j = 0; j < NB; j++) inv_column(A11, L21, A22, j, true, invL);
#endif
}

#if DS_CUDA && !defined(DS_DMMA884_DEFINED)
#define DS_DMMA884_DEFINED
/* D(8x8) += A(8x4) B(4x8) on the FP64 tensor cores.  Lane T holds a = A[T/4][T%4],
 * b = B[T%4][T/4], and d0,d1 = D[T/4][2*(T%4)], D[T/4][2*(T%4)+1]. */
DS_FN void dmma884(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
#endif

/* X (8x8) = R (8x8) * C^T (8x8), R[g][m] = ra[g*sa + m], C[c][m] = cb[c*sb + m]; lane
 * (g = T/4, q = T%4) receives X[g][2q], X[g][2q+1].  Used for the panel solve
 * (C = inv(L_kk)) and, through P, for the trailing update. */
DS_FN void tile_mul_nt(int lane, const double *ra, int sa, const double *cb, int sb, double &d0, double &d1) {
  const int g = lane >> 2, q = lane & 3;
#if DS_CUDA
  d0 = 0.0; d1 = 0.0;
  dmma884(d0, d1, ra[g * sa + q], cb[g * sb + q]);
  dmma884(d0, d1, ra[g * sa + 4 + q], cb[g * sb + 4 + q]);
#else
  double s0 = 0.0, s1 = 0.0;
  for (int m = 0; m < NB; m++) {
    s0 += ra[g * sa + m] * cb[(2 * q) * sb + m];
    s1 += ra[g * sa + m] * cb[(2 * q + 1) * sb + m];
  }
  d0 = s0; d1 = s1;
#endif
}

/* same product with both operands taken from the panel buffer P[h][r][4]:
 * X[g][c] = sum_cc P(rA+g, cc) P(rB+c, cc) */
DS_FN void tile_mul_pp(int lane, const double *P, int HS, int rA, int rB, double &d0, double &d1) {
  const int g = lane >> 2, q = lane & 3;
#if DS_CUDA
  d0 = 0.0; d1 = 0.0;
  dmma884(d0, d1, P[(rA + g) * 4 + q], P[(rB + g) * 4 + q]);
  dmma884(d0, d1, P[HS + (rA + g) * 4 + q], P[HS + (rB + g) * 4 + q]);
#else
  double s0[2] = {0.0, 0.0}, s1[2] = {0.0, 0.0}; /* k = 0..3 and k = 4..7 summed separately, like the kernel */
  for (int cc = 0; cc < NB; cc++) {
    const double a = P[pidx(cc, rA + g, HS)];
    s0[cc >> 2] += a * P[pidx(cc, rB + 2 * q, HS)];
    s1[cc >> 2] += a * P[pidx(cc, rB + 2 * q + 1, HS)];
  }
  d0 = s0[0] + s0[1]; d1 = s1[0] + s1[1];
#endif
}

/* One 8-row tile of the panel solve: X = A inv(L_kk)^T, in place, and into the
 * panel buffer.  base[g*rs + cc] = A[g][cc]; the entry is inside the band iff
 * o0 + og*g + cc >= lo (always, for the border rows). */
DS_FN void panel_tile(const Team team, double *base, int rs, int o0, int og, int lo, const double *invL, double *P,
                      int HS, int r0) {
#if DS_CUDA
  const int lane = team.tid & 31, g = lane >> 2, q = lane & 3;
  const int o = o0 + og * g;
  const double a0 = (o + q >= lo) ? base[g * rs + q] : 0.0;
  const double a1 = (o + 4 + q >= lo) ? base[g * rs + 4 + q] : 0.0;
  double d0 = 0.0, d1 = 0.0;
  dmma884(d0, d1, a0, invL[g * ILS + q]);
  dmma884(d0, d1, a1, invL[g * ILS + 4 + q]);
  /* mma.sync is warp-synchronous: every lane has read its operands (the barrier states it) */
  __syncwarp();
  if (o + 2 * q >= lo) base[g * rs + 2 * q] = d0;
  if (o + 2 * q + 1 >= lo) base[g * rs + 2 * q + 1] = d1;
  *(dbl2 *)&P[pidx(2 * q, r0 + g, HS)] = dbl2{d0, d1};
#else
  (void)team;
  double X[NB][NB];
  for (int g = 0; g < NB; g++) {
    const int o = o0 + og * g;
    for (int cc = 0; cc < NB; cc++) {
      double sx = 0.0;
      for (int m = 0; m < NB; m++) {
        const double a = (o + m >= lo) ? base[g * rs + m] : 0.0;
        sx += a * invL[cc * ILS + m];
      }
      X[g][cc] = sx;
    }
  }
  for (int g = 0; g < NB; g++) {
    const int o = o0 + og * g;
    for (int cc = 0; cc < NB; cc++) {
      if (o + cc >= lo) base[g * rs + cc] = X[g][cc];
      P[pidx(cc, r0 + g, HS)] = X[g][cc];
    }
  }
#endif
}

/* C(pair) -= d for the 16-byte pair a DMMA accumulator lane owns */
DS_FN void sub_pair(double *dst, double d0, double d1) {
  dbl2 *p2 = (dbl2 *)dst;
  dbl2 cv = *p2;
  cv.x -= d0; cv.y -= d1;
  *p2 = cv;
}

/* One trailing-update tile, in the form the inner loop consumes it. */
struct alignas(16) TileDesc {
  int pa, pb;   /* 32 * first panel row of the A / B operand = its byte offset into P (rows are
                   multiples of 8, the border rows start at bwp) */
  int kind;     /* 1 window tile, all 64 entries inside the band and below the diagonal;
                   2 window tile that needs per-entry checks; 3 border-row tile; 4 corner */
  int doff;     /* 8 * (kinds 1-2: rB - rA + bwE, the band offset of the tile's first column in its first
                   row; kind 3: NB + rB, the column of the border rows relative to k) = byte offset */
};

/* cost of the look-ahead diagonal factorisation in units of one trailing-update tile (measured:
 * ~2.8 k cycles against ~320 per tile with two resident CTAs) */
constexpr int LOOKAHEAD_TILES = 9;

/* Tiles in row-major order over the lower triangle of tile rows, the border tile row last; built
 * once per problem.  Warp w > 0 owns the contiguous range [wstart[w], wstart[w+1]) of tiles
 * 1..nshared-1 (consecutive tiles of a tile row share their A operand); warp 0 owns tile 0 -- the
 * next diagonal block, which it then factors (look-ahead) -- and the n0 tiles
 * [nshared, ntiles) so that all warps finish together.  wstart[nwarp] = nshared. */
DS_FN void build_tile_table(const Team team, TileDesc *tt, int *wstart, int nwarp, int nt8, int bwp, int bw, int bwE) {
  const int nrow = nt8 + 1, ntiles = nrow * (nrow + 1) / 2;
  DS_FOR(t, ntiles) {
    int ti = 0, tj = t;
    while (tj > ti) { tj -= ti + 1; ti++; }
    const bool erow = ti == nt8, ecol = tj == nt8;
    const int rA = erow ? bwp : ti * NB, rB = ecol ? bwp : tj * NB;
    TileDesc d;
    d.pa = 32 * rA;
    d.pb = 32 * rB;
    if (erow) { d.kind = ecol ? 4 : 3; d.doff = 8 * (NB + rB); }
    else { d.kind = (ti > tj && NB * (ti - tj) + (NB - 1) <= bw) ? 1 : 2; d.doff = 8 * (rB - rA + bwE); }
    tt[t] = d;
  }
  /* warp 0's fixed load per step: its panel tile, tile 0 and the look-ahead factorisation; the other
   * warps share nt8 panel tiles and the remaining trailing tiles */
  /* DS_ISO: warp 4 shares its scheduler (SM sub-partition = warp id mod 4) and FP64 pipe with warp 0, whose
   * diagonal factorisation is the latency chain of the step; it takes the TMA duty and no tensor-core work, so the
   * chain's DFMAs do not queue behind 16-cycle DMMAs */
  const int idle_w = (DS_ISO && nwarp == 8) ? 4 : -1;
  const int nwork = nwarp - 1 - (idle_w >= 0 ? 1 : 0); /* warps that run S2/S3 next to warp 0 */
  int n0 = ((ntiles - 1) + nt8 - (LOOKAHEAD_TILES + 2) * nwork) / (nwork + 1);
  if (n0 < 0) n0 = 0;
  const int nshared = ntiles - n0;
  team.sync();
  /* Contiguous ranges of equal COST for the worker warps: a panel tile the warp solves before the update
   * (S2, row tiles 1 + wi, 1 + wi + nwork, ...) counts 1.4 tiles, a tile that needs per-entry checks
   * (diagonal, corner) 1.5 -- ratios from the per-warp cycle counters of the profile build. */
  if (team.tid == 0) {
    wstart[0] = nshared;
    if (nwarp > 1) {
      int total = 0;
      for (int t = 1; t < nshared; t++) total += (tt[t].kind == 2 || tt[t].kind == 4) ? 15 : 10;
      for (int rt = 1; rt <= nt8; rt++) total += 14;
      int cur = 1, wi = 0;
      for (int w = 1; w < nwarp; w++) {
        wstart[w] = cur;
        if (w == idle_w) continue; /* empty range: wstart[w+1] = cur as well */
        int acc = 0;
        for (int rt = 1 + wi; rt <= nt8; rt += nwork) acc += 14;
        const int budget = total / (nwork - wi);
        if (wi == nwork - 1) cur = nshared;
        while (cur < nshared) {
          const int cst = (tt[cur].kind == 2 || tt[cur].kind == 4) ? 15 : 10;
          if (acc + cst / 2 >= budget) break;
          acc += cst;
          cur++;
        }
        total -= acc;
        wi++;
      }
      wstart[nwarp] = nshared;
    } else {
      wstart[1] = nshared;
    }
  }
}

#if DS_CUDA
/* C -= X for the 8x8 product a trailing-update tile produced (d0, d1 = this lane's pair) */
DS_FN void s3_store(const TileDesc td, double d0, double d1, double *W, double *Eb, double *G, int ks8, int k, int Wr,
                    int ld, int ES, int bwE, int lo, int g, int q) {
  if (td.kind <= 2) {
    /* entry (i, j): i = k+8+rA+g, j = k+8+rB+2q(+1); band offset j-i+bwE */
    int s0 = ks8 + (td.pa >> 5);
    if (s0 >= Wr) s0 -= Wr;
    const int off = (td.doff >> 3) + 2 * q - g;
    double *dst = W + (s0 + g) * ld + off;
    if (td.kind == 1) {
      sub_pair(dst, d0, d1);
    } else {
      const bool v0 = off <= bwE && off >= lo, v1 = off + 1 <= bwE && off + 1 >= lo;
      if (v0 && v1) sub_pair(dst, d0, d1);
      else {
        if (v0) dst[0] -= d0;
        if (v1) dst[1] -= d1;
      }
    }
  } else if (td.kind == 3) {
    sub_pair(Eb + g * ES + k + (td.doff >> 3) + 2 * q, d0, d1);
  } else {
    if (2 * q <= g) G[g * 8 + 2 * q] -= d0;
    if (2 * q + 1 <= g) G[g * 8 + 2 * q + 1] -= d1;
  }
}

/* tiles [beg, end) of the table, two at a time: the operand loads, the loads of the two
 * destination pairs and the four DMMAs of a pair of tiles are independent, and consecutive tiles
 * of a tile row reuse the A fragments.  All shared-memory traffic of the common case (kind 1)
 * goes through 32-bit shared addresses.  CHECK: steps near the end of the matrix, where the
 * tiles of rows beyond it are skipped. */
template <bool CHECK>
DS_FN void s3_range(const TileDesc *tt, int beg, int end, const double *P, int HS, double *W, double *Eb, double *G,
                    int ks8, int k, int Wr, int ld, int ES, int bwE, int lo, int n_trail, int lane, bool e_smem) {
  const int g = lane >> 2, q = lane & 3;
  const uint32_t tt_s = smem_u32(tt);
  /* lane's pair in border row g at column k (shared-memory border rows only) */
  const uint32_t El = e_smem ? smem_u32(Eb) + 8u * (uint32_t)(g * ES + 2 * q + k) : 0u;
  const uint32_t Plo = smem_u32(P) + 8u * (uint32_t)(g * 4 + q), Phi = Plo + 8u * (uint32_t)HS;
  const uint32_t Wl = smem_u32(W) + 8u * (uint32_t)(g * ld + 2 * q - g); /* lane's pair in row g of a tile */
  const uint32_t ldb = 8u * (uint32_t)ld;
  const int nt4 = 32 * n_trail;
  for (int i = beg; i < end; i += 2) {
    const bool two = i + 1 < end;
    const int4 u0 = lds_v4s32(tt_s + 16u * (uint32_t)i), u1 = lds_v4s32(tt_s + 16u * (uint32_t)(two ? i + 1 : i));
    bool do0 = true, do1 = two;
    if (CHECK) {
      do0 = u0.z == 4 || (u0.y < nt4 && (u0.z == 3 || u0.x < nt4));
      do1 = two && (u1.z == 4 || (u1.y < nt4 && (u1.z == 3 || u1.x < nt4)));
      if (!do0 && !do1) continue;
    }
    const double a0l = lds_f64(Plo + u0.x), a0h = lds_f64(Phi + u0.x);
    const double b0l = lds_f64(Plo + u0.y), b0h = lds_f64(Phi + u0.y);
    double a1l = a0l, a1h = a0h;
    if (u1.x != u0.x) { a1l = lds_f64(Plo + u1.x); a1h = lds_f64(Phi + u1.x); }
    const double b1l = lds_f64(Plo + u1.y), b1h = lds_f64(Phi + u1.y);
    /* destination of a window tile: row slot of its first row in the ring, band offset doff */
    int s0 = ks8 + (u0.x >> 5), s1 = ks8 + (u1.x >> 5);
    if (s0 >= Wr) s0 -= Wr;
    if (s1 >= Wr) s1 -= Wr;
    const bool e0 = e_smem && u0.z == 3, e1 = e_smem && u1.z == 3;
    const uint32_t d0 = e0 ? El + (uint32_t)u0.w : Wl + ldb * (uint32_t)s0 + (uint32_t)u0.w;
    const uint32_t d1 = e1 ? El + (uint32_t)u1.w : Wl + ldb * (uint32_t)s1 + (uint32_t)u1.w;
    const bool f0 = do0 && (u0.z == 1 || e0), f1 = do1 && (u1.z == 1 || e1);
    dbl2 v0 = {0.0, 0.0}, v1 = {0.0, 0.0};
    if (f0) v0 = lds_v2f64(d0);
    if (f1) v1 = lds_v2f64(d1);
    /* four independent DMMAs (k = 0..3 and k = 4..7 of each tile), summed afterwards: no
     * accumulator dependency between tensor-core instructions */
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0, e00 = 0.0, e01 = 0.0, e10 = 0.0, e11 = 0.0;
    dmma884(c00, c01, a0l, b0l);
    dmma884(c10, c11, a1l, b1l);
    dmma884(e00, e01, a0h, b0h);
    dmma884(e10, e11, a1h, b1h);
    c00 += e00; c01 += e01; c10 += e10; c11 += e11;
    if (f0) sts_v2f64(d0, v0.x - c00, v0.y - c01);
    else if (do0) {
      TileDesc td; td.pa = u0.x; td.pb = u0.y; td.kind = u0.z; td.doff = u0.w;
      s3_store(td, c00, c01, W, Eb, G, ks8, k, Wr, ld, ES, bwE, lo, g, q);
    }
    if (f1) sts_v2f64(d1, v1.x - c10, v1.y - c11);
    else if (do1) {
      TileDesc td; td.pa = u1.x; td.pb = u1.y; td.kind = u1.z; td.doff = u1.w;
      s3_store(td, c10, c11, W, Eb, G, ks8, k, Wr, ld, ES, bwE, lo, g, q);
    }
  }
}
#endif

#if DS_CUDA
/* named barriers: producers arrive, consumers sync (count = all threads of the CTA) */
DS_FN void named_arrive(int id, int count) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
DS_FN void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
#endif

}  // namespace ds
#include "sft_rows.h"
namespace ds {

/* Solve (H + lambda I) dx = b.  dx -> sm[sl.dx] (nodes, then camera at Dn_pad).
 * Returns false if a pivot is not positive (LinearSolverDense::solve returning
 * false, linear_solver_dense.h:107-112); dx is then left untouched (stale), as
 * in the reference. */
template <bool XS>
DS_FN_NOINLINE bool factor_solve(const Team team, Ctx &cx, double lambda) {
  Ctx &c = ctx_ref();
  (void)cx;
  const int bw = c.pl.bw, bwE = c.pl.bwE, ld = c.pl.ld, Dp = c.pl.Dn_pad, Wr = c.pl.Wr, bwp = c.pl.bwp,
            nblk = c.pl.nblk, ES = c.pl.ES;
  const int HS = panel_half_stride(bwp); /* panel rows: trailing rows, then the 8 border rows */
  double *const sm = sm_base();
  double *W = sm + c.sl.W, *P = sm + c.sl.P, *invL = sm + c.sl.invL;
  double *G = sm + c.sl.G, *Hcc = sm + c.sl.Hcc, *dx = dxvec<XS>(c);
  const bool e_smem = c.pb.e_in_smem != 0;
  double *Es = sm_base() + c.sl.E, *Eg = c.ws.Eg;
  const double *Hb = c.ws.Hb;
  double *Lb = c.ws.Lb;
  int *flag = (int *)(sm + c.sl.red + 36);
  const int nt8 = bwp / NB;
  const TileDesc *tiles = (const TileDesc *)(sm + c.sl.tiles);
  const int *wstart = (const int *)(tiles + (nt8 + 1) * (nt8 + 2) / 2);
#if DS_CUDA
  const int warp = team.tid >> 5, nwarp = team.nthr >> 5, lane = team.tid & 31;
#else
  const int warp = 0, nwarp = 1;
#endif

  /* Everything H/border related in global memory was written with plain stores
   * (build_system); order them before the TMA reads. */
  fence_proxy_async();
  team.sync();
  uint64_t *bar0 = &c.mbar[0];
  uint32_t ph0 = c.ph[0];
  const uint64_t pol = l2_policy_evict_first();
  /* the factor is written top-down and read back bottom-up within the same trial (LIFO) */
#ifndef DS_LB_POLICY
#define DS_LB_POLICY 0
#endif
  const uint64_t polL = DS_LB_POLICY == 0 ? pol : (DS_LB_POLICY == 1 ? l2_policy_evict_normal() : l2_policy_evict_last());
  /* window <- first Wr rows of H; border/rhs working copy (bulk async); corner */
  {
    const int rows = Wr < Dp ? Wr : Dp;
    if (team.tid == 0) {
      const uint32_t bw_bytes = (uint32_t)(rows * ld * sizeof(double));
      const uint32_t e_bytes = e_smem ? (uint32_t)(8 * ES * sizeof(double)) : 0u;
      mbar_expect_tx(bar0, bw_bytes + e_bytes);
      tma_load_1d_stream(W, Hb, bw_bytes, bar0, pol);
      if (e_smem) tma_load_1d(Es, c.ws.Cg, e_bytes, bar0);
    }
    if (!e_smem) {
      const double *Cg = c.ws.Cg;
      DS_FOR(i, 8 * ES) Eg[i] = Cg[i];
    }
    DS_FOR(i, 64) {
      const int a = i >> 3, b = i & 7;
      double v = 0.0;
      if (a < 6 && b < 6) v = Hcc[a * 6 + b] + (a == b ? lambda : 0.0);
      else if (a == 6 && b < 6) v = Hcc[36 + b];
      G[i] = v;
    }
    if (team.tid == 0) *flag = 0;
    mbar_wait(bar0, ph0);
    ph0 ^= 1u;
  }
  team.sync();
  prof_mark(team, c, PF_FS_INIT);

  int kslot = 0; /* k mod Wr */
  /* S1 of the first block; every later diagonal block is factored by warp 0 during the
   * trailing update of the step before it (look-ahead), off the critical path */
  if (team.warp0()) {
    diag_factor(team, W, 0, Wr, ld, bwE, lambda, invL, flag);
    fence_proxy_async_smem(); /* rows 0..7 are final: they leave through the async proxy */
  }
  team.sync();
  /* One step per 8-row block.  Warp 0 is the critical path: the panel tile right under the
   * diagonal block, the update of the next diagonal block (tile 0) and its factorisation (S1 of
   * step kb+1, look-ahead).  The other warps solve the rest of the panel (S2), meet on a named
   * barrier that warp 0 only arrives at, and run the trailing update (S3).  One CTA-wide
   * barrier per step.  inv(L_kk) is double-buffered: warp 0 writes the next one while the
   * others may still read the current one. */
#if DS_CUDA
  /* TMA duty (bulk store of finished rows, refill of the freed slot): warp 1, or the warp kept free of
   * tensor-core work (DS_ISO) */
  const int iso_w = (DS_ISO && nwarp == 8) ? 4 : -1;
  const int tma_tid = iso_w >= 0 ? 32 * iso_w : 32;
#else
  const int tma_tid = 0;
#endif
  for (int kb = 0; kb < nblk; kb++) {
    const int k = kb * NB;
    const double *invLk = invL + 96 * (kb & 1);
    double *invLn = invL + 96 * ((kb + 1) & 1);
    /* finished rows k..k+7 -> L band in global memory (TMA bulk store) */
    if (team.tid == tma_tid)
      tma_store_1d_stream(Lb + k * ld, W + kslot * ld, (uint32_t)(NB * ld * sizeof(double)), polL);
    prof_mark(team, c, PF_S1);
    /* rows requested during the previous step (they enter this step's panel) */
    if (kb > 0 && (k - NB) + Wr < Dp) { mbar_wait(bar0, ph0); ph0 ^= 1u; }
    prof_mark(team, c, PF_S1_WAIT);

    /* S2: panel = (rows below, then the 8 border rows) times L_kk^-T, one 8-row
     * tile per warp: X = A inv(L_kk)^T on the tensor cores */
    const int n_trail = (Dp - (k + NB)) < bwp ? (Dp - (k + NB)) : bwp;
    const int nrt = n_trail / NB; /* trailing row tiles that exist */
    const int nrow = nt8 + 1;
    const int ntiles = nrow * (nrow + 1) / 2;
    const int lo = bwE - bw;
    DS_PROF_T0(s2t0);
#if DS_CUDA
    /* warp 0: the tile under the diagonal block; worker wi: row tiles 1 + wi, 1 + wi + nwork, ... */
    const int nwork = nwarp - 1 - (iso_w >= 0 ? 1 : 0);
    const int wi = warp - 1 - (iso_w >= 0 && warp > iso_w ? 1 : 0);
    const int rt_first = warp == 0 ? 0 : (warp == iso_w ? nt8 + 1 : 1 + wi), rt_step = warp == 0 ? nt8 + 1 : nwork;
#else
    const int rt_first = 0, rt_step = 1;
#endif
    for (int rt = rt_first; rt <= nt8; rt += rt_step) {
      const bool erow = rt == nt8;
      if (!erow && rt >= nrt) continue; /* rows beyond the matrix: never read by S3 */
      const int r0 = erow ? bwp : rt * NB; /* first panel row of this tile */
      if (!erow) {
        /* trailing rows i = k+8+r0+g: column k+cc sits at band offset o0+cc with
         * o0 = bwE-8-r0-g; offsets below lo = bwE-bw are outside the band (zero).
         * The 8 rows of a tile never wrap inside the ring (Wr, r0 multiples of 8). */
        int s0 = kslot + NB + r0;
        if (s0 >= Wr) s0 -= Wr;
        panel_tile(team, W + s0 * ld + (bwE - NB - r0), ld - 1, bwE - NB - r0, -1, bwE - bw, invLk, P, HS, r0);
      } else if (e_smem) {
        panel_tile(team, Es + k, ES, bwE, 0, bwE - bw, invLk, P, HS, r0);
      } else {
        panel_tile(team, Eg + k, ES, bwE, 0, bwE - bw, invLk, P, HS, r0);
      }
    }
    DS_PROF_ADD(PF_X_S2W + warp, s2t0, (team.tid & 31) == 0);
#if DS_CUDA
    {
      double *Eb = e_smem ? Es : Eg;
      const int ks8 = kslot + NB;
      const bool full = n_trail == bwp;
      const bool w0_tail = wstart[0] < ntiles; /* warp 0 also owns trailing tiles: it needs the whole panel */
      DS_PROF_INC(PF_X_STEPS, team.tid == 0);
      if (warp == 0) {
        named_arrive(1, team.nthr); /* panel rows 0..7 are in P */
        DS_PROF_T0(s3t0);
        team.warp_sync();
        /* S3, tile 0: the next diagonal block, the only tile that touches rows k+8..k+15 */
        if (n_trail > 0) s3_range<false>(tiles, 0, 1, P, HS, W, Eb, G, ks8, k, Wr, ld, ES, bwE, lo, n_trail, lane, e_smem);
        if (kb + 1 < nblk) {
          team.warp_sync();
          int ns = kslot + NB;
          if (ns >= Wr) ns -= Wr;
          DS_PROF_T0(dgt0);
          diag_factor(team, W, ns, Wr, ld, bwE, lambda, invLn, flag);
          DS_PROF_ADD(PF_X_DIAG, dgt0, lane == 0);
        }
        if (w0_tail) {
          named_sync(2, team.nthr);
          if (full) s3_range<false>(tiles, wstart[0], ntiles, P, HS, W, Eb, G, ks8, k, Wr, ld, ES, bwE, lo, n_trail, lane, e_smem);
          else s3_range<true>(tiles, wstart[0], ntiles, P, HS, W, Eb, G, ks8, k, Wr, ld, ES, bwE, lo, n_trail, lane, e_smem);
        }
        DS_PROF_ADD(PF_X_WARP + warp, s3t0, lane == 0);
      } else {
        if (w0_tail) named_arrive(2, team.nthr);
        named_sync(1, team.nthr); /* the whole panel is in P; every thread is past the mbarrier wait */
        prof_mark(team, c, PF_S2);
        /* the bulk store of rows k..k+7 has read its source by now: refill the freed slot with
         * rows k+Wr.. (needed by the next step's panel) */
        if (team.tid == tma_tid) {
          tma_store_wait_read();
          if (k + Wr < Dp) {
            const uint32_t bytes = (uint32_t)(NB * ld * sizeof(double));
            mbar_expect_tx(bar0, bytes);
            tma_load_1d_stream(W + kslot * ld, Hb + (k + Wr) * ld, bytes, bar0, pol);
          }
        }
        DS_PROF_T0(s3t0);
        /* S3: trailing update C -= P_I P_J^T on 8x8 tiles (FP64 tensor cores) */
        const int beg = wstart[warp], end = wstart[warp + 1];
        if (full) s3_range<false>(tiles, beg, end, P, HS, W, Eb, G, ks8, k, Wr, ld, ES, bwE, lo, n_trail, lane, e_smem);
        else s3_range<true>(tiles, beg, end, P, HS, W, Eb, G, ks8, k, Wr, ld, ES, bwE, lo, n_trail, lane, e_smem);
        DS_PROF_ADD(PF_X_WARP + warp, s3t0, lane == 0);
      }
    }
#else
    /* emulation: one thread, the phases in sequence */
    (void)wstart;
    if (k + Wr < Dp) tma_load_1d_stream(W + kslot * ld, Hb + (k + Wr) * ld, (uint32_t)(NB * ld * sizeof(double)), bar0, pol);
    for (int t = 0; t < ntiles; t++) {
      const TileDesc td = tiles[t];
      const int rA = td.pa >> 5, rB = td.pb >> 5;
      bool skip = td.kind <= 2 ? (rA >= n_trail || rB >= n_trail) : (td.kind == 3 && rB >= n_trail);
      if (!skip) {
        for (int T = 0; T < 32; T++) {
          const int g = T >> 2, q = T & 3;
          double d0, d1;
          tile_mul_pp(T, P, HS, rA, rB, d0, d1);
          if (td.kind <= 2) {
            int s0 = kslot + NB + rA;
            if (s0 >= Wr) s0 -= Wr;
            const int off = (td.doff >> 3) + 2 * q - g;
            double *dst = W + (s0 + g) * ld + off;
            if (off <= bwE && off >= lo) dst[0] -= d0;
            if (off + 1 <= bwE && off + 1 >= lo) dst[1] -= d1;
          } else if (td.kind == 3) {
            const int eo = g * ES + k + (td.doff >> 3) + 2 * q;
            if (e_smem) sub_pair(Es + eo, d0, d1);
            else sub_pair(Eg + eo, d0, d1);
          } else {
            if (2 * q <= g) G[g * 8 + 2 * q] -= d0;
            if (2 * q + 1 <= g) G[g * 8 + 2 * q + 1] -= d1;
          }
        }
      }
      if (t == 0 && kb + 1 < nblk) {
        int ns = kslot + NB;
        if (ns >= Wr) ns -= Wr;
        diag_factor(team, W, ns, Wr, ld, bwE, lambda, invLn, flag);
      }
    }
#endif
    /* rows written here are bulk-stored (async proxy) after a later barrier */
    fence_proxy_async_smem();
    team.sync();
    prof_mark(team, c, PF_S3);
    kslot += NB;
    if (kslot >= Wr) kslot -= Wr;
  }

  /* Schur complement system of the camera: S dc = rhs  (6x6 Cholesky, in place
   * in the shared corner block; row 6 of G is the right-hand side) */
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
      double *y = G + 56; /* row 7 of the corner is unused */
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
  /* all bulk stores of L must have landed before the TMA reads of the backward sweep */
  if (team.tid == tma_tid) tma_store_wait_all();
  if (team.tid == 0) c.ph[0] = ph0;
  team.sync();
  prof_mark(team, c, PF_SCHUR);
  if (*flag != 0) { team.sync(); return false; }

  /* v = z - Y^T dc */
  if (e_smem) {
    DS_FOR(i, Dp) {
      double s = Es[6 * ES + i];
#pragma unroll
      for (int e = 0; e < 6; e++) s -= Es[e * ES + i] * dx[Dp + e];
      dx[i] = s;
    }
  } else {
    DS_FOR(i, Dp) {
      double s = Eg[6 * ES + i];
#pragma unroll
      for (int e = 0; e < 6; e++) s -= Eg[e * ES + i] * dx[Dp + e];
      dx[i] = s;
    }
  }
  /* backward sweep  L^T dn = v, one block of NB rows per step.  Rows of L stream
   * back from global memory through a ring of NBUF buffers in the (now free)
   * window, NBUF-1 steps ahead. */
  constexpr int NBUF = 4;
  const int bufsz = NB * ld;
  double *sol = W + NBUF * bufsz;
  const uint32_t row_bytes = (uint32_t)(NB * ld * sizeof(double));
  uint32_t phb[NBUF];
#pragma unroll
  for (int b = 0; b < NBUF; b++) phb[b] = c.ph[1 + b];
  team.sync(); /* all reads of the window / E done before the ring overwrites it */
  if (team.tid == 0) {
    for (int j = 0; j < NBUF - 1 && j < nblk; j++) {
      const int kbj = nblk - 1 - j;
      mbar_expect_tx(&c.mbar[1 + (j % NBUF)], row_bytes);
      tma_load_1d_stream(W + (j % NBUF) * bufsz, Lb + kbj * NB * ld, row_bytes, &c.mbar[1 + (j % NBUF)], polL);
    }
  }
  prof_mark(team, c, PF_BWD_INIT);
  DS_PROF_LOCALS(bwacc, 4);
  for (int kb = nblk - 1; kb >= 0; kb--) {
    const int k = kb * NB;
    const int j = nblk - 1 - kb, buf = j % NBUF;
    /* request block j+NBUF-1: its buffer was last read at step j-1, before the
     * barrier that ended that step */
    if (team.tid == 0) {
      const int jn = j + NBUF - 1;
      if (jn < nblk) {
        const int kbn = nblk - 1 - jn, bn = jn % NBUF;
        mbar_expect_tx(&c.mbar[1 + bn], row_bytes);
        tma_load_1d_stream(W + bn * bufsz, Lb + kbn * NB * ld, row_bytes, &c.mbar[1 + bn], polL);
      }
    }
    const double *LR = W + buf * bufsz;
    DS_PROF_T0M(bwt);
#pragma unroll
    for (int b = 0; b < NBUF; b++)
      if (b == buf) { mbar_wait(&c.mbar[1 + b], phb[b]); phb[b] ^= 1u; }
    DS_PROF_LAP(bwacc, 0, bwt);
    /* d = L_kk^-T y by backward substitution (1/L_ii on the diagonal); every
     * thread that needs d (the updaters of dx[k-bw .. k+7]) computes it
     * redundantly from shared memory: no barrier between solve and update */
    const int j0 = k - bw > 0 ? k - bw : 0;
    const int nupd = k - j0; /* entries left of the block that change */
    const bool active = team.tid < (nupd > NB ? nupd : NB);
    double d[NB];
    if (active) {
#pragma unroll
      for (int a = 0; a < NB; a++) d[a] = dx[k + a];
      /* the 36 entries of L_kk first (independent loads), then the dependent chain */
      double Lk[NB * (NB + 1) / 2];
#pragma unroll
      for (int a = 0; a < NB; a++)
#pragma unroll
        for (int m = 0; m <= a; m++) Lk[a * (a + 1) / 2 + m] = LR[a * ld + bwE - a + m];
#pragma unroll
      for (int a = NB - 1; a >= 0; a--) {
        d[a] *= Lk[a * (a + 1) / 2 + a];
#pragma unroll
        for (int m = 0; m < a; m++) d[m] -= d[a] * Lk[a * (a + 1) / 2 + m];
      }
    }
    DS_PROF_LAP(bwacc, 1, bwt);
    /* the solution goes to its own vector: dx[k..k+7] is still being read as the right-hand
     * side by slower threads, and one barrier per step is enough */
    if (active) {
      if (team.tid == 0) {
#pragma unroll
        for (int a = 0; a < NB; a++) sol[k + a] = d[a];
      }
      DS_FOR(jj, nupd) {
        const int jc = j0 + jj;
        double s = dx[jc];
        const double *Lc = LR + (jc - k + bwE); /* Lc[a*(ld-1)] = L[k+a][jc] */
        const int lo = bwE - bw;
#pragma unroll
        for (int a = 0; a < NB; a++) {
          if (jc - k - a + bwE >= lo) s -= Lc[a * (ld - 1)] * d[a];
        }
        dx[jc] = s;
      }
    }
    DS_PROF_LAP(bwacc, 2, bwt);
    team.sync();
    DS_PROF_LAP(bwacc, 3, bwt);
  }
  DS_PROF_FLUSH(bwacc, 4, PF_X_BWD, team.tid == 0);
  if (team.tid == 0) {
#pragma unroll
    for (int b = 0; b < NBUF; b++) c.ph[1 + b] = phb[b];
  }
  DS_FOR(i, Dp) dx[i] = sol[i];
  team.sync();
  prof_mark(team, c, PF_BWD);
  return true;
}

/* ------------------------------------------------------------- LM ------ */

template <bool XS>
DS_FN void apply_update(const Team team, Ctx &cx) {
  Ctx &c = ctx_ref();
  (void)cx;
  const PlanView &pl = c.pl;
  double *x = xvec<XS>(c), *dx = dxvec<XS>(c);
  /* VertexSBAPointXYZ::oplusImpl (types_sba.h:52-56); fixed nodes have dx = 0 */
  DS_FOR(i, pl.Dn) x[i] += dx[i];
  if (team.tid == 0) { /* VertexSE3Expmap::oplusImpl */
    Pose P;
    load_pose(sm_base() + c.sl.pose, P);
    pose_oplus(P, &dx[pl.Dn_pad]);
    double *ps = sm_base() + c.sl.pose;
    for (int k = 0; k < 4; k++) ps[k] = P.q[k];
    for (int k = 0; k < 3; k++) ps[4 + k] = P.t[k];
  }
  team.sync();
}

DS_FN void expand_dense(const Team team, Ctx &cx, double chi) {
  Ctx &c = ctx_ref();
  (void)cx;
  const PlanView &pl = c.pl;
  const int Dn = pl.Dn, D = Dn + 6, bw = pl.bw, bwE = pl.bwE, ld = pl.ld, ES = pl.ES;
  const double *Hcc = sm_base() + c.sl.Hcc;
  team.sync();
  if (c.pb.out_H) {
    DS_FOR(idx, D * D) {
      const int i = idx / D, j = idx - i * D;
      const int hi = i > j ? i : j, lo = i > j ? j : i;
      double v = 0.0;
      if (hi < Dn) { if (hi - lo <= bw) v = c.ws.Hb[hi * ld + (lo - hi + bwE)]; }
      else if (lo < Dn) v = c.ws.Cg[(hi - Dn) * ES + lo];
      else v = Hcc[(hi - Dn) * 6 + (lo - Dn)];
      c.pb.out_H[idx] = v;
    }
  }
  if (c.pb.out_b) {
    DS_FOR(i, D) {
      double v = i < Dn ? c.ws.Cg[6 * ES + i] : Hcc[36 + (i - Dn)];
      if (i < Dn && !freev_ptr(c)[i / 3]) v = 0.0;
      c.pb.out_b[i] = v;
    }
  }
  if (team.tid == 0) {
    c.pb.out_res->chi2_initial = chi;
    c.pb.out_res->chi2_final = chi;
    c.pb.out_res->status = 0;
    c.pb.out_res->n_viewed = c.n_viewed;
    c.pb.out_res->n_optlap = c.n_optlap;
  }
}

/* DefOptimizer.cc:515-577 */
template <bool XS>
DS_FN_NOINLINE void finalize(const Team team, Ctx &cx, bool last_rejected, int iterations, int trials, double chi_ini,
                             double chi_fin, double lambda) {
  Ctx &c = ctx_ref();
  (void)cx;
  const PlanView &pl = c.pl;
  const ProbView &pb = c.pb;
  const int M = pb.n_matches, n = pl.n_nodes;
  const double *x = xvec<XS>(c), *xb = c.ws.xb;
  Pose P, Pb;
  load_pose(sm_base() + c.sl.pose, P);
  load_pose(sm_base() + c.sl.pose + 8, Pb);
  /* e->chi2() is the error of the LAST computeActiveErrors, i.e. of the last LM
   * trial whether it was accepted or not (the state was popped, the edges were
   * not re-evaluated). */
  const double *xl = last_rejected ? xb : x;
  const Pose &Pl = last_rejected ? Pb : P;
  int nbad = 0, cnt = 0;
  double sum = 0.0;
  DS_FOR(m, M) {
    double e[2], Pc[3];
    reproj_error(c, xl, Pl, m, e, Pc);
    const double info = match_info(c, m);
    const float chi2 = (float)(e[0] * info * e[0] + e[1] * info * e[1]);
    /* matches-given overload: outlier <=> deltaMono < |e| (DefOptimizer.cc:806-818) */
    const bool out = pb.variant == 1 ? c.hub_delta < sqrt(pow(e[0], 2) + pow(e[1], 2)) : chi2 > 5.991;
    if (pb.out_outlier) pb.out_outlier[m] = out ? 1 : 0;
    if (out) nbad++;
    else {
      reproj_error(c, x, P, m, e, Pc);
      sum += sqrt(pow(e[0], 2) + pow(e[1], 2));
      cnt++;
    }
  }
  nbad = team_sum_int(team, nbad, sm_base() + c.sl.red);
  cnt = team_sum_int(team, cnt, sm_base() + c.sl.red);
  sum = team_sum(team, sum, sm_base() + c.sl.red);
  if (pb.out_nodes) DS_FOR(i, pl.Dn) pb.out_nodes[i] = x[i];
  if (pb.out_role) DS_FOR(v, n) pb.out_role[v] = (uint8_t)(viewed_ptr(c)[v] | (freev_ptr(c)[v] << 1));
  if (team.tid == 0) {
    ResultScalars *r = pb.out_res;
    pose_to_Tcw(P, r->Tcw);
    r->rep_error = (float)(sum / (double)(unsigned)cnt);
    r->n_inliers = M - nbad;
    r->lm_iterations = iterations;
    r->lm_trials = trials;
    r->chi2_initial = chi_ini;
    r->chi2_final = chi_fin;
    r->lambda_final = lambda;
    r->status = 0;
    r->n_viewed = c.n_viewed;
    r->n_optlap = c.n_optlap;
  }
}

/* One frame, start to finish.  All threads of the team call this. */
template <bool XS>
DS_FN_NOINLINE void sft_solve_one(const Team team, Ctx &cx) {
  Ctx &c = ctx_ref();
  (void)cx;
  const PlanView &pl = c.pl;
  const ProbView &pb = c.pb;
  double *x = xvec<XS>(c), *xb = c.ws.xb, *dx = dxvec<XS>(c);
  double *ps = sm_base() + c.sl.pose, *psb = ps + 8;
  double *red = sm_base() + c.sl.red;
  const int Dp = pl.Dn_pad;

  const int rc = prologue<XS>(team, c);
  if (pb.row_nt == 0) {
    TileDesc *tt = (TileDesc *)(sm_base() + c.sl.tiles);
    const int nt8 = pl.bwp / NB;
#if DS_CUDA
    const int nwarp = team.nthr >> 5;
#else
    const int nwarp = 1;
#endif
    build_tile_table(team, tt, (int *)(tt + (nt8 + 1) * (nt8 + 2) / 2), nwarp, nt8, pl.bwp, pl.bw, pl.bwE);
  }
  team.sync();
  prof_mark(team, c, PF_PROLOGUE);
  if (rc != 0) {
    if (team.tid == 0) pb.out_res->status = rc;
    return;
  }
  if (pb.mode == MODE_NORMAL_EQ) {
    const double chi = eval_state<XS, true>(team, c);
    build_system<XS>(team, c);
    expand_dense(team, c, chi);
    return;
  }

  /* OptimizationAlgorithmLevenberg::solve, optimization_algorithm_levenberg.cpp:61-164,
   * driven by SparseOptimizer::optimize, sparse_optimizer.cpp:403-475 */
  const int max_it = pb.max_it > 0 ? pb.max_it : 50;
  const double tau = 1e-5, goodUpper = 2. / 3., goodLower = 1. / 3.;
  const int maxTrials = 10;
  double lambda = -1., ni = 2., chi_ini0 = 0., chi_fin = 0.;
  int nBad = 0, iterations = 0, trials = 0;
  bool last_rejected = false;
  for (int it = 0; it < max_it; it++) {
    double currentChi = eval_state<XS, true>(team, c);
    prof_mark(team, c, PF_EVAL_STORE);
    double tempChi = currentChi;
    const double iniChi = currentChi;
    if (it == 0) chi_ini0 = currentChi;
    const double maxDiag = build_system<XS>(team, c);
    prof_mark(team, c, PF_BUILD);
    if (it == 0) { lambda = tau * maxDiag; ni = 2; nBad = 0; }
    const double lambda_start = lambda;
    double rho = 0;
    int qmax = 0;
    do {
      /* push */
      team.sync();
      DS_FOR(i, pl.Dn) xb[i] = x[i];
      if (team.tid == 0) for (int k = 0; k < 7; k++) psb[k] = ps[k];
      prof_mark(team, c, PF_LM_SCALAR);
      bool ok2;
      if (XS && pb.row_nt > 0) {
        switch (pb.row_nt) { /* tiles per block row = bwp/8 + 1: 5 (4x4 mesh) .. 14 (17x17) */
          case 5: ok2 = factor_rows<5>(team, lambda); break;
          case 6: ok2 = factor_rows<6>(team, lambda); break;
          case 7: ok2 = factor_rows<7>(team, lambda); break;
          case 8: ok2 = factor_rows<8>(team, lambda); break;
          case 9: ok2 = factor_rows<9>(team, lambda); break;
          case 10: ok2 = factor_rows<10>(team, lambda); break;
          case 11: ok2 = factor_rows<11>(team, lambda); break;
          case 12: ok2 = factor_rows<12>(team, lambda); break;
          case 13: ok2 = factor_rows<13>(team, lambda); break;
          default: ok2 = factor_rows<14>(team, lambda); break;
        }
      } else {
        ok2 = factor_solve<XS>(team, c, lambda);
      }
      apply_update<XS>(team, c);
      prof_mark(team, c, PF_UPDATE);
      tempChi = eval_state<XS, false>(team, c);
      prof_mark(team, c, PF_EVAL_TRIAL);
      if (!ok2) tempChi = DBL_MAX;
      rho = currentChi - tempChi;
      double scale = 0.; /* computeScale :182-189 */
      DS_FOR(j, pl.Dn) scale += dx[j] * (lambda * dx[j] + c.ws.Cg[6 * pl.ES + j]);
      scale = team_sum(team, scale, red);
      for (int j = 0; j < 6; j++) scale += dx[Dp + j] * (lambda * dx[Dp + j] + (sm_base() + c.sl.Hcc)[36 + j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = fmin(alpha, goodUpper);
        const double scaleFactor = fmax(goodLower, alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        last_rejected = false;
      } else {
        lambda *= ni;
        ni *= 2;
        /* pop: restore; the backup buffers keep the rejected trial state */
        team.sync();
        DS_FOR(i, pl.Dn) { const double t = x[i]; x[i] = xb[i]; xb[i] = t; }
        if (team.tid == 0) for (int k = 0; k < 7; k++) { const double t = ps[k]; ps[k] = psb[k]; psb[k] = t; }
        team.sync();
        last_rejected = true;
      }
      qmax++;
      trials++;
    } while (rho < 0 && qmax < maxTrials);
    if (pb.out_trace && it < pb.trace_cap && team.tid == 0) {
      pb.out_trace[4 * it + 0] = iniChi; pb.out_trace[4 * it + 1] = lambda_start;
      pb.out_trace[4 * it + 2] = (double)qmax; pb.out_trace[4 * it + 3] = currentChi;
    }
    chi_fin = currentChi;
    iterations = it + 1;
    if (qmax == maxTrials || rho == 0) break;
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) break;
  }
  team.sync();
  prof_mark(team, c, PF_LM_SCALAR);
  finalize<XS>(team, c, last_rejected, iterations, trials, chi_ini0, chi_fin, lambda);
  prof_mark(team, c, PF_FINALIZE);
}

/* Entry shared by the CUDA kernel and the emulation: bind a problem to a team,
 * its shared memory and its global workspace, then solve it.  The context is
 * kept in shared memory (one copy per CTA): with the shared-memory carve-out
 * this kernel uses there is almost no L1 left for a per-thread stack copy. */
template <bool XS>
DS_FN void sft_run_problem(const Team &team, const ProbView &pv, double *smem, uint8_t *ws_base,
                           const WorkspaceSizes &z, bool first_of_launch, long long *prof) {
#if !DS_CUDA
  ds_smem_emu = smem;
#else
  (void)smem;
#endif
  Ctx &c = ctx_ref();
  double *sm = sm_base();
  team.sync();
  if (team.tid == 0) {
    if (first_of_launch) {
      for (int i = 0; i < 16; i++) { mbar_init(&c.mbar[i], 1); c.ph[i] = 0; }
      fence_mbar_init();
    }
    c.prof = prof;
#if DS_CUDA
    c.prof_last = clock64();
#endif
    c.pb = pv;
    c.pl = *pv.plan;
    c.ws = carve_workspace(ws_base, z);
    c.sl = smem_layout(c.pl.n_nodes, c.pl.n_edges, c.pl.Dn_pad, c.pl.bwp, c.pl.ld, c.pl.Wr, c.pl.ES, pv.e_in_smem != 0,
                       pv.x_in_smem != 0, pv.row_nt);
  }
  team.sync();
  sft_solve_one<XS>(team, c);
}

}  // namespace ds
#endif
