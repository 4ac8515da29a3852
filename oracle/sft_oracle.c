/*
 * sft_oracle.c -- CPU restatement of DefSLAM's Shape-from-Template solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build or call it, and there only as the checker
 * or the timed CPU baseline.
 *
 * PARITY PINNED TO THE REFERENCE'S OWN CODE: the reference ships no tests or
 * vectors for this path and cannot be built as a whole here (Eigen, OpenCV,
 * Ceres are absent), but the files that hold the path's arithmetic compile
 * against a minimal stand-in for Eigen (oracle/ref_shim/): sft_types.h,
 * se3quat.h and core/base_{unary,binary,multi}_edge.hpp verbatim where they lie,
 * the Levenberg driver (optimization_algorithm_levenberg.cpp:43-189), the Huber
 * kernel (robust_kernel_impl.cpp:65-91), activeRobustChi2/update
 * (sparse_optimizer.cpp:104-120,477-491) and the vertex oplus bodies by line
 * range (oracle/Makefile target g2oref -> oracle/_ref/libg2o_sft_ref.so,
 * harness g2o_ref_harness.cc).  tests/test_oracle_sft_ref.py checks this file
 * against that library (per-edge errors and Jacobians to 1e-15, H/b/chi2 to
 * 1e-13, identical LM iteration/trial pattern, lambda and chi2 traces, final
 * nodes to 1e-9) live and through tests/golden/sft_ref.npz.  What remains a
 * restatement: the graph construction of DefOptimizer.cc:251-513 (needs the
 * Frame/Map/OpenCV types), BlockSolver bookkeeping, and Eigen's dense LDLT.
 *
 * What it follows (paths under the DefSLAM tree):
 *   graph construction      Modules/Tracking/DefOptimizer.cc:251-578
 *   residuals + Jacobians   Thirdparty/g2o/g2o/types/sft_types.h:75-411
 *   quadratic forms         Thirdparty/g2o/g2o/core/base_multi_edge.hpp:36-48,171-222
 *                           base_binary_edge.hpp:57-130, base_unary_edge.hpp:43-72
 *   Huber                   Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91
 *   LM                      Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189
 *   dense solve             Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:65-113
 *   SE3 exp / update        Thirdparty/g2o/g2o/types/se3quat.h:223-257,
 *                           types_six_dof_expmap.h:73-76
 *
 * It deliberately keeps the reference's cost structure: the graph is rebuilt
 * per call, every edge (including the deg(i) duplicated curvature edges, quirk
 * C2) is linearised separately, the normal matrix is dense over the free
 * variables and is re-factorised by a dense LDL^T on every LM trial, single
 * threaded (g2o OpenMP is off in the reference, Thirdparty/g2o/config.h:4).
 *
 * Deviation that cannot be avoided: vertex/edge order.  The reference orders
 * nodes and edges by heap address (std::set<Node*>, quirk C13); here nodes are
 * in index order, neighbours ascending, edges in edge-list order.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/defslam_b200.h"
#include "sft_oracle.h"

/* ------------------------------------------------------------------ SE3 -- */

/* Eigen::Quaterniond(Matrix3d) -- the conversion SE3Quat(R,t) relies on
 * (se3quat.h:57-59).  q = (x,y,z,w). */
static void quat_from_R(const double R[9], double q[4]) {
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}

/* SE3Quat::normalizeRotation  se3quat.h:280-285 */
static void quat_normalize(double q[4]) {
  if (q[3] < 0) {
    q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3];
  }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

/* Eigen::Quaternion::toRotationMatrix */
static void quat_to_R(const double q[4], double R[9]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* a*b, Eigen quaternion product */
static void quat_mul(const double a[4], const double b[4], double o[4]) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

/* q * v  (Eigen: v + w*(2 q x v) + q x (2 q x v)) */
static void quat_rot(const double q[4], const double v[3], double o[3]) {
  double uv[3] = {2 * (q[1] * v[2] - q[2] * v[1]), 2 * (q[2] * v[0] - q[0] * v[2]),
                  2 * (q[0] * v[1] - q[1] * v[0])};
  o[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
  o[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
  o[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

static void mat3_mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

/* SE3Quat::exp  se3quat.h:223-257 ; update = (omega, upsilon) */
static void se3_exp(const double u[6], double q[4], double t[3]) {
  const double w[3] = {u[0], u[1], u[2]}, ups[3] = {u[3], u[4], u[5]};
  const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double Om[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double Om2[9], R[9], V[9];
  mat3_mul(Om, Om, Om2);
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) R[i] = ((i % 4) == 0 ? 1.0 : 0.0) + Om[i] + Om2[i];
    memcpy(V, R, sizeof(R));
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta),
                 c = (theta - sin(theta)) / pow(theta, 3);
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4) == 0 ? 1.0 : 0.0;
      R[i] = I + a * Om[i] + b * Om2[i];
      V[i] = I + b * Om[i] + c * Om2[i];
    }
  }
  quat_from_R(R, q);
  for (int i = 0; i < 3; i++) t[i] = V[i * 3] * ups[0] + V[i * 3 + 1] * ups[1] + V[i * 3 + 2] * ups[2];
  quat_normalize(q); /* SE3Quat(q,t) ctor normalises */
}

/* VertexSE3Expmap::oplusImpl: estimate = exp(update) * estimate
 * (types_six_dof_expmap.h:73-76, SE3Quat::operator* se3quat.h:103-109) */
static void pose_oplus(double q[4], double t[3], const double u[6]) {
  double dq[4], dt[3], rt[3], nq[4];
  se3_exp(u, dq, dt);
  quat_rot(dq, t, rt);
  for (int i = 0; i < 3; i++) t[i] = dt[i] + rt[i];
  quat_mul(dq, q, nq);
  memcpy(q, nq, sizeof(nq));
  quat_normalize(q);
}

/* Converter::toSE3Quat  Thirdparty/ORBSLAM_2/src/Converter.cc:31-43 */
static void pose_from_Tcw(const float T[16], double q[4], double t[3]) {
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  quat_from_R(R, q);
  quat_normalize(q);
  t[0] = T[3]; t[1] = T[7]; t[2] = T[11];
}

/* Converter::toCvMat(SE3Quat)  Converter.cc:45-48,57-64 */
static void pose_to_Tcw(const double q[4], const double t[3], float T[16]) {
  double R[9];
  quat_to_R(q, R);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = (float)R[i * 3 + j];
    T[i * 4 + 3] = (float)t[i];
  }
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
}

/* ---------------------------------------------------------------- graph -- */

#include "sft_oracle_graph.h"

void oracle_graph_free(Graph *g) {
  if (!g) return;
  for (int i = 0; i < g->n_curv; i++) {
    free(g->curv[i].v); free(g->curv[i].w); free(g->curv[i].J);
  }
  free(g->x); free(g->idx); free(g->rep); free(g->ref); free(g->curv); free(g->str);
  free(g->viewed); free(g->optlap);
  free(g->H); free(g->b); free(g->dx); free(g->Hwork); free(g->diag_backup);
}

/* DefOptimizer.cc:251-507: build the graph */
int oracle_graph_build(Graph *g, const defslam_sft_problem *p) {
  const defslam_template_desc *td = p->tmpl_desc;
  if (!td) return DEFSLAM_EBADARG;
  const int n = td->n_nodes;
  memset(g, 0, sizeof(*g));
  g->n_nodes = n;
  g->n_matches = p->n_matches;
  g->fx = p->fx; g->fy = p->fy; g->cx = p->cx; g->cy = p->cy;
  pose_from_Tcw(p->T_cw, g->q, g->t); /* :269 */
  g->x = (double *)malloc(sizeof(double) * 3 * n);
  memcpy(g->x, p->node_xyz, sizeof(double) * 3 * n); /* setMeshNodes :926-952 */
  g->viewed = (uint8_t *)calloc(n, 1);
  g->optlap = (uint8_t *)calloc(n, 1);
  g->idx = (int *)malloc(sizeof(int) * n);

  g->variant = p->matches_given ? 1 : 0;
  if (g->variant && !(p->curv_edge_len > 0.0)) return DEFSLAM_EBADARG;
  /* const float deltaMono = sqrt(5.991);  :286 (float!) ; const float deltaMono = 0.5 in the overload :625 */
  const float deltaMono = g->variant ? 0.5f : (float)sqrt(5.991);
  g->huber_delta = (double)deltaMono;
  g->huber_dsqr = (double)(float)(g->huber_delta * g->huber_delta); /* setDelta robust_kernel_impl.cpp:65-69 into "float dsqr" (robust_kernel_impl.h:84) */

  /* ---- reprojection edges :293-361 */
  g->rep = (EdgeReproj *)calloc(p->n_matches > 0 ? p->n_matches : 1, sizeof(EdgeReproj));
  g->n_rep = p->n_matches;
  const int N = p->n_frame_keypoints;
  for (int m = 0; m < p->n_matches; m++) {
    EdgeReproj *e = &g->rep[m];
    e->m = m;
    for (int k = 0; k < 3; k++) {
      e->v[k] = p->match_nodes[3 * m + k];
      if (e->v[k] < 0 || e->v[k] >= n) return DEFSLAM_EBADARG;
      e->bary[k] = p->match_bary[3 * m + k];
      g->viewed[e->v[k]] = 1; /* ViewedNodes.insert :331 */
    }
    e->obs[0] = (double)p->match_uv[2 * m];
    e->obs[1] = (double)p->match_uv[2 * m + 1];
    /* Identity * invSigma2 / N with float invSigma2, int N  :339-340 ; Identity / double(matches.size()) :655 */
    e->info = g->variant ? 1.0 / (double)p->n_matches : (double)p->match_inv_sigma2[m] / (double)N;
  }

  /* ---- temporal edges for viewed nodes :363-382 */
  const double mlen = td->edge_median_len;
  g->info_ref = p->reg_temp / pow(mlen, 2);
  g->ref = (EdgeRef *)calloc(n, sizeof(EdgeRef));
  for (int v = 0; v < n; v++)
    if (g->viewed[v]) {
      g->n_viewed += g->variant; /* (counted below for the first overload) */
      if (g->variant) continue;  /* the overload builds these edges but never adds them (:676-689) */
      EdgeRef *e = &g->ref[g->n_ref++];
      e->v = v;
      for (int c = 0; c < 3; c++) e->meas[c] = td->node_rest_xyz[3 * v + c];
      g->n_viewed++;
    }

  /* ---- OptLap = Viewed U ring1(Viewed) :384-406 (layers>1 == 1, quirk C3) */
  memcpy(g->optlap, g->viewed, n);
  if (g->variant) memset(g->optlap, 1, n); /* every node is free (:615-620); g->optlap = the free set */
  else if (p->neighbour_layers >= 1)
    for (int v = 0; v < n; v++)
      if (g->viewed[v])
        for (int k = td->nbr_ptr[v]; k < td->nbr_ptr[v + 1]; k++) g->optlap[td->nbr_idx[k]] = 1;
  g->n_optlap = 0;
  for (int v = 0; v < n; v++) g->n_optlap += g->optlap[v];

  /* free variables: camera rows 0..5, then OptLap nodes ascending (:414-419) */
  int row = 6;
  for (int v = 0; v < n; v++) {
    if (g->optlap[v]) { g->idx[v] = row; row += 3; } else g->idx[v] = -1;
  }
  g->D = row;

  /* node -> incident edges (Node::getEdges) */
  int *deg = (int *)calloc(n + 1, sizeof(int));
  for (int e = 0; e < td->n_edges; e++) { deg[td->edge_ab[2 * e] + 1]++; deg[td->edge_ab[2 * e + 1] + 1]++; }
  for (int v = 0; v < n; v++) deg[v + 1] += deg[v];
  int *inc = (int *)malloc(sizeof(int) * (2 * td->n_edges + 1));
  int *fill = (int *)calloc(n, sizeof(int));
  for (int e = 0; e < td->n_edges; e++)
    for (int s = 0; s < 2; s++) {
      int v = td->edge_ab[2 * e + s];
      inc[deg[v] + fill[v]++] = e;
    }

  /* ---- curvature edges :411-463, one per incident edge of every
   *      non-boundary OptLap node (quirk C2) */
  int ncurv = 0;
  /* curvature centres: OptLap -- in the overload OptLap = ViewedNodes (:693) although every node is free */
  const uint8_t *lap = g->variant ? g->viewed : g->optlap;
  g->n_curv_den = g->variant ? g->n_viewed : g->n_optlap;
  for (int v = 0; v < n; v++)
    if (lap[v] && !td->node_boundary[v]) ncurv += deg[v + 1] - deg[v];
  g->curv = (EdgeCurv *)calloc(ncurv > 0 ? ncurv : 1, sizeof(EdgeCurv));
  g->info_curv = g->n_curv_den > 0 ? p->reg_lap / (double)g->n_curv_den : 0.0; /* :458, :755 */
  for (int v = 0; v < n; v++) {
    if (!(lap[v] && !td->node_boundary[v])) continue;
    const int nn = td->nbr_ptr[v + 1] - td->nbr_ptr[v];
    for (int ie = deg[v]; ie < deg[v + 1]; ie++) {
      EdgeCurv *e = &g->curv[g->n_curv++];
      e->nv = nn + 1;
      e->v = (int *)malloc(sizeof(int) * e->nv);
      e->w = (double *)malloc(sizeof(double) * (nn > 0 ? nn : 1));
      e->J = (double *)calloc(3 * e->nv, sizeof(double));
      e->v[0] = v;
      for (int k = 0; k < nn; k++) {
        e->v[k + 1] = td->nbr_idx[td->nbr_ptr[v] + k];
        e->w[k] = td->nbr_w[td->nbr_ptr[v] + k];
      }
      /* setDistanceEdges((*ite)->getDist()) :446 ; never set in the overload (quirk C8): the caller's value */
      e->len = g->variant ? p->curv_edge_len : td->edge_len0[inc[ie]];
      e->kappa0 = td->node_kappa0[v];  /* GetMeanCurvatureInitial :449-453 */
    }
  }

  /* ---- stretch edges: every mesh edge incident to an OptLap node :466-507 */
  g->str = (EdgeStretch *)calloc(td->n_edges > 0 ? td->n_edges : 1, sizeof(EdgeStretch));
  for (int e = 0; e < td->n_edges; e++) {
    const int a = td->edge_ab[2 * e], b = td->edge_ab[2 * e + 1];
    if (g->optlap[a] || g->optlap[b]) {
      EdgeStretch *s = &g->str[g->n_str++];
      s->a = a; s->b = b; s->len0 = td->edge_len0[e];
    }
  }
  g->info_str = g->n_str > 0 ? p->reg_inex / (double)g->n_str : 0.0; /* :499 */

  free(deg); free(inc); free(fill);

  g->H = (double *)calloc((size_t)g->D * g->D, sizeof(double));
  g->Hwork = (double *)calloc((size_t)g->D * g->D, sizeof(double));
  g->b = (double *)calloc(g->D, sizeof(double));
  g->dx = (double *)calloc(g->D, sizeof(double));
  g->diag_backup = (double *)calloc(g->D, sizeof(double));
  return 0;
}

/* ------------------------------------------------- errors and Jacobians -- */

/* EdgeNodesCamera::computeError  sft_types.h:102-133 */
static void reproj_error(const Graph *g, EdgeReproj *e) {
  double Pw[3], Pc[3];
  for (int c = 0; c < 3; c++)
    Pw[c] = e->bary[0] * g->x[3 * e->v[0] + c] + e->bary[1] * g->x[3 * e->v[1] + c] +
            e->bary[2] * g->x[3 * e->v[2] + c];
  quat_rot(g->q, Pw, Pc);
  for (int c = 0; c < 3; c++) Pc[c] += g->t[c];
  const double u = Pc[0] / Pc[2] * g->fx + g->cx;
  const double v = Pc[1] / Pc[2] * g->fy + g->cy;
  e->err[0] = e->obs[0] - u;
  e->err[1] = e->obs[1] - v;
}

/* EdgeNodesCamera::linearizeOplus  sft_types.h:137-206 (quirk C1 kept) */
static void reproj_linearize(const Graph *g, EdgeReproj *e) {
  double R[9], xk[3][3], xyz[3] = {0, 0, 0};
  quat_to_R(g->q, R);
  for (int k = 0; k < 3; k++) {
    quat_rot(g->q, &g->x[3 * e->v[k]], xk[k]);
    for (int c = 0; c < 3; c++) xk[k][c] += g->t[c];
  }
  for (int c = 0; c < 3; c++) xyz[c] = xk[0][c] * e->bary[0] + xk[1][c] * e->bary[1] + xk[2][c] * e->bary[2];
  const double fx = g->fx, fy = g->fy;
  double x = xyz[0], y = xyz[1], z = xyz[2], z_2 = z * z;
  double *J = e->Jc;
  J[0] = x * y / z_2 * fx;
  J[1] = -(1 + (x * x / z_2)) * fx;
  J[2] = y / z * fx;
  J[3] = -1. / z * fx;
  J[4] = 0;
  J[5] = x / z_2 * fx;
  J[6] = (1 + y * y / z_2) * fy;
  J[7] = -x * y / z_2 * fy;
  J[8] = -x / z * fy;
  J[9] = 0;
  J[10] = -1. / z * fy;
  J[11] = y / z_2 * fy;
  for (int k = 0; k < 3; k++) {
    x = xk[k][0]; y = xk[k][1]; z = xk[k][2];
    const double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 3; c++) {
        const double tr = tmp[r * 3] * R[c] + tmp[r * 3 + 1] * R[3 + c] + tmp[r * 3 + 2] * R[6 + c];
        e->Jn[k][r * 3 + c] = -1. / z * tr * e->bary[k];
      }
  }
}

/* EdgeMeanCurvature::computeError  sft_types.h:257-291 */
static void curv_error(const Graph *g, EdgeCurv *e) {
  double acc[3] = {0, 0, 0}, sw = 0;
  for (int k = 1; k < e->nv; k++) {
    for (int c = 0; c < 3; c++) acc[c] = acc[c] + e->w[k - 1] * g->x[3 * e->v[k] + c];
    sw = sw + e->w[k - 1];
  }
  e->sumw = sw;
  for (int c = 0; c < 3; c++) e->mc[c] = g->x[3 * e->v[0] + c] - acc[c] / sw;
  e->mcn = sqrt(e->mc[0] * e->mc[0] + e->mc[1] * e->mc[1] + e->mc[2] * e->mc[2]);
  e->err = (e->mcn - e->kappa0) / e->len;
}

/* EdgeMeanCurvature::linearizeOplus  sft_types.h:293-311 */
static void curv_linearize(EdgeCurv *e) {
  for (int i = 0; i < e->nv; i++) {
    double *J = &e->J[3 * i];
    if (e->mcn < 1E-15) { J[0] = J[1] = J[2] = 0.0; continue; }
    const double wa = (i == 0) ? 1.0 : -(e->w[i - 1] / e->sumw);
    for (int c = 0; c < 3; c++) J[c] = wa * e->mc[c] / (e->mcn * e->len);
  }
}

/* EdgesStreching  sft_types.h:353-378 */
static void stretch_error(const Graph *g, EdgeStretch *e) {
  double d[3];
  for (int c = 0; c < 3; c++) d[c] = g->x[3 * e->a + c] - g->x[3 * e->b + c];
  const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  e->err = nrm * (1.0 / e->len0) - 1.0;
}
static void stretch_linearize(const Graph *g, EdgeStretch *e) {
  double d[3];
  for (int c = 0; c < 3; c++) d[c] = g->x[3 * e->a + c] - g->x[3 * e->b + c];
  const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const double ddo = 1.0 / (nrm * e->len0);
  for (int c = 0; c < 3; c++) e->Ja[c] = d[c] * ddo;
}

/* EdgesReference::computeError  sft_types.h:403-408 */
static void ref_error(const Graph *g, EdgeRef *e) {
  for (int c = 0; c < 3; c++) e->err[c] = g->x[3 * e->v + c] - e->meas[c];
}

/* SparseOptimizer::computeActiveErrors */
static void compute_active_errors(Graph *g) {
  for (int i = 0; i < g->n_rep; i++) reproj_error(g, &g->rep[i]);
  for (int i = 0; i < g->n_ref; i++) ref_error(g, &g->ref[i]);
  for (int i = 0; i < g->n_curv; i++) curv_error(g, &g->curv[i]);
  for (int i = 0; i < g->n_str; i++) stretch_error(g, &g->str[i]);
}

/* RobustKernelHuber::robustify  robust_kernel_impl.cpp:78-91 */
static void huber(const Graph *g, double e2, double rho[2]) {
  if (e2 <= g->huber_dsqr) { rho[0] = e2; rho[1] = 1.; }
  else {
    const double sq = sqrt(e2);
    rho[0] = 2 * sq * g->huber_delta - g->huber_dsqr;
    rho[1] = g->huber_delta / sq;
  }
}

/* SparseOptimizer::activeRobustChi2  sparse_optimizer.cpp:104-120
 * (edges in insertion order: reprojection, temporal, curvature, stretch) */
static double active_robust_chi2(const Graph *g) {
  double chi = 0.0, rho[2];
  for (int i = 0; i < g->n_rep; i++) {
    const EdgeReproj *e = &g->rep[i];
    const double c2 = e->err[0] * e->info * e->err[0] + e->err[1] * e->info * e->err[1];
    huber(g, c2, rho);
    chi += rho[0];
  }
  for (int i = 0; i < g->n_ref; i++) {
    const EdgeRef *e = &g->ref[i];
    chi += (e->err[0] * e->err[0] + e->err[1] * e->err[1] + e->err[2] * e->err[2]) * g->info_ref;
  }
  for (int i = 0; i < g->n_curv; i++) chi += g->curv[i].err * g->info_curv * g->curv[i].err;
  for (int i = 0; i < g->n_str; i++) chi += g->str[i].err * g->info_str * g->str[i].err;
  return chi;
}

/* accumulate  H(ri.., rj..) += Ji^T * w * Jj  for an r-row residual block.
 * Ji: r x ci row-major, Jj: r x cj.  Only called with ri <= rj (upper). */
static void add_JtWJ(Graph *g, int ri, int ci, const double *Ji, int rj, int cj, const double *Jj,
                     int r, double w) {
  const int D = g->D;
  for (int a = 0; a < ci; a++)
    for (int b = 0; b < cj; b++) {
      double s = 0;
      for (int k = 0; k < r; k++) s += Ji[k * ci + a] * w * Jj[k * cj + b];
      g->H[(size_t)(ri + a) * D + rj + b] += s;
    }
}
static void add_Jtr(Graph *g, int ri, int ci, const double *Ji, int r, const double *wr) {
  for (int a = 0; a < ci; a++) {
    double s = 0;
    for (int k = 0; k < r; k++) s += Ji[k * ci + a] * wr[k];
    g->b[ri + a] += s;
  }
}

/* BlockSolver::buildSystem  block_solver.hpp:502-560: zero, linearise every
 * edge, constructQuadraticForm (upper blocks only), then mirror. */
static void build_system(Graph *g) {
  const int D = g->D;
  memset(g->H, 0, sizeof(double) * (size_t)D * D);
  memset(g->b, 0, sizeof(double) * D);
  double rho[2];
  /* reprojection: BaseMultiEdge::constructQuadraticForm base_multi_edge.hpp:36-48 */
  for (int i = 0; i < g->n_rep; i++) {
    EdgeReproj *e = &g->rep[i];
    reproj_linearize(g, e);
    const double c2 = e->err[0] * e->info * e->err[0] + e->err[1] * e->info * e->err[1];
    huber(g, c2, rho);
    const double w = rho[1] * e->info; /* robustInformation, base_edge.h:96-102 */
    const double wr[2] = {-e->info * e->err[0] * rho[1], -e->info * e->err[1] * rho[1]};
    /* vertex 0 = camera (free) */
    add_JtWJ(g, 0, 6, e->Jc, 0, 6, e->Jc, 2, w);
    add_Jtr(g, 0, 6, e->Jc, 2, wr);
    for (int k = 0; k < 3; k++) {
      const int rk = g->idx[e->v[k]];
      if (rk < 0) continue;
      add_JtWJ(g, 0, 6, e->Jc, rk, 3, e->Jn[k], 2, w);
    }
    for (int k = 0; k < 3; k++) {
      const int rk = g->idx[e->v[k]];
      if (rk < 0) continue;
      add_JtWJ(g, rk, 3, e->Jn[k], rk, 3, e->Jn[k], 2, w);
      add_Jtr(g, rk, 3, e->Jn[k], 2, wr);
      for (int l = k + 1; l < 3; l++) {
        const int rl = g->idx[e->v[l]];
        if (rl < 0) continue;
        if (rk <= rl) add_JtWJ(g, rk, 3, e->Jn[k], rl, 3, e->Jn[l], 2, w);
        else add_JtWJ(g, rl, 3, e->Jn[l], rk, 3, e->Jn[k], 2, w);
      }
    }
  }
  /* temporal: BaseUnaryEdge::constructQuadraticForm base_unary_edge.hpp:43-72 */
  static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < g->n_ref; i++) {
    EdgeRef *e = &g->ref[i];
    const int rv = g->idx[e->v];
    if (rv < 0) continue;
    add_JtWJ(g, rv, 3, I3, rv, 3, I3, 3, g->info_ref);
    const double wr[3] = {-g->info_ref * e->err[0], -g->info_ref * e->err[1], -g->info_ref * e->err[2]};
    add_Jtr(g, rv, 3, I3, 3, wr);
  }
  /* curvature: multi-edge, fixed vertices skipped (base_multi_edge.hpp:173-177,199) */
  for (int i = 0; i < g->n_curv; i++) {
    EdgeCurv *e = &g->curv[i];
    curv_linearize(e);
    const double wr = -g->info_curv * e->err;
    for (int a = 0; a < e->nv; a++) {
      const int ra = g->idx[e->v[a]];
      if (ra < 0) continue;
      add_JtWJ(g, ra, 3, &e->J[3 * a], ra, 3, &e->J[3 * a], 1, g->info_curv);
      add_Jtr(g, ra, 3, &e->J[3 * a], 1, &wr);
      for (int c = a + 1; c < e->nv; c++) {
        const int rc = g->idx[e->v[c]];
        if (rc < 0) continue;
        if (ra <= rc) add_JtWJ(g, ra, 3, &e->J[3 * a], rc, 3, &e->J[3 * c], 1, g->info_curv);
        else add_JtWJ(g, rc, 3, &e->J[3 * c], ra, 3, &e->J[3 * a], 1, g->info_curv);
      }
    }
  }
  /* stretch: BaseBinaryEdge::constructQuadraticForm base_binary_edge.hpp:57-130 */
  for (int i = 0; i < g->n_str; i++) {
    EdgeStretch *e = &g->str[i];
    stretch_linearize(g, e);
    const double Jb[3] = {-e->Ja[0], -e->Ja[1], -e->Ja[2]};
    const double wr = -g->info_str * e->err;
    const int ra = g->idx[e->a], rb = g->idx[e->b];
    if (ra >= 0) {
      add_JtWJ(g, ra, 3, e->Ja, ra, 3, e->Ja, 1, g->info_str);
      add_Jtr(g, ra, 3, e->Ja, 1, &wr);
    }
    if (rb >= 0) {
      add_JtWJ(g, rb, 3, Jb, rb, 3, Jb, 1, g->info_str);
      add_Jtr(g, rb, 3, Jb, 1, &wr);
    }
    if (ra >= 0 && rb >= 0) {
      if (ra <= rb) add_JtWJ(g, ra, 3, e->Ja, rb, 3, Jb, 1, g->info_str);
      else add_JtWJ(g, rb, 3, Jb, ra, 3, e->Ja, 1, g->info_str);
    }
  }
  /* mirror upper -> lower (linear_solver_dense.h:93-98 does it at copy time) */
  for (int i = 0; i < D; i++)
    for (int j = i + 1; j < D; j++) g->H[(size_t)j * D + i] = g->H[(size_t)i * D + j];
}

/* LinearSolverDense::solve  linear_solver_dense.h:65-113: dense LDL^T of the
 * full matrix, fail unless positive.  Eigen's LDLT pivots on the diagonal;
 * without pivoting the factors differ but the solution agrees to rounding.
 * Left-looking, row-major, contiguous dot products. */
int oracle_dense_ldlt_solve(int D, const double *H, double *L, const double *b, double *x) {
  /* L is D*D scratch: strictly-lower holds L, diagonal holds d */
  double *v = (double *)malloc(sizeof(double) * D);
  int ok = 1;
  for (int j = 0; j < D; j++) {
    double *Lj = &L[(size_t)j * D];
    double dj = H[(size_t)j * D + j];
    for (int k = 0; k < j; k++) { v[k] = Lj[k] * L[(size_t)k * D + k]; dj -= Lj[k] * v[k]; }
    Lj[j] = dj;
    if (!(dj > 0.0)) { ok = 0; break; }
    const double inv = 1.0 / dj;
    for (int i = j + 1; i < D; i++) {
      double *Li = &L[(size_t)i * D];
      double s = H[(size_t)i * D + j];
      for (int k = 0; k < j; k++) s -= Li[k] * v[k];
      Li[j] = s * inv;
    }
  }
  if (ok) {
    for (int i = 0; i < D; i++) {
      double s = b[i];
      const double *Li = &L[(size_t)i * D];
      for (int k = 0; k < i; k++) s -= Li[k] * x[k];
      x[i] = s;
    }
    for (int i = 0; i < D; i++) x[i] /= L[(size_t)i * D + i];
    for (int i = D - 1; i >= 0; i--) {
      double s = x[i];
      for (int k = i + 1; k < D; k++) s -= L[(size_t)k * D + i] * x[k];
      x[i] = s;
    }
  }
  free(v);
  return ok;
}

/* SparseOptimizer::update  sparse_optimizer.cpp:477-491 */
static void apply_update(Graph *g, const double *dx) {
  pose_oplus(g->q, g->t, dx); /* camera rows 0..5 */
  for (int v = 0; v < g->n_nodes; v++) {
    const int r = g->idx[v];
    if (r < 0) continue;
    for (int c = 0; c < 3; c++) g->x[3 * v + c] += dx[r + c]; /* VertexSBAPointXYZ::oplusImpl types_sba.h:52-56 */
  }
}

/* ------------------------------------------------------------------- LM -- */

typedef struct {
  int iterations, trials;
  double chi2_initial, chi2_final, lambda;
} LMStats;

/* SparseOptimizer::optimize (sparse_optimizer.cpp:403-475) driving
 * OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-164) */
static void run_lm(Graph *g, int max_iterations, LMStats *st, double *trace, int trace_cap) {
  const int D = g->D;
  double lambda = -1., ni = 2.;
  int nBad = 0;
  const double tau = 1e-5, goodUpper = 2. / 3., goodLower = 1. / 3.;
  const int maxTrials = 10;
  double *xb = (double *)malloc(sizeof(double) * 3 * g->n_nodes);
  double qb[4], tb[3];
  memset(st, 0, sizeof(*st));
  int it;
  for (it = 0; it < max_iterations; it++) {
    compute_active_errors(g);
    double currentChi = active_robust_chi2(g);
    double tempChi = currentChi;
    const double iniChi = currentChi;
    if (it == 0) st->chi2_initial = currentChi;
    build_system(g);
    if (it == 0) { /* computeLambdaInit :166-180 */
      double maxDiag = 0.;
      for (int k = 0; k < D; k++) maxDiag = fmax(fabs(g->H[(size_t)k * D + k]), maxDiag);
      lambda = tau * maxDiag;
      ni = 2;
      nBad = 0;
    }
    const double lambda_start = lambda;
    double rho = 0;
    int qmax = 0;
    do {
      /* push */
      memcpy(xb, g->x, sizeof(double) * 3 * g->n_nodes);
      memcpy(qb, g->q, sizeof(qb)); memcpy(tb, g->t, sizeof(tb));
      /* setLambda(lambda, backup=true)  block_solver.hpp:564-589 */
      for (int k = 0; k < D; k++) { g->diag_backup[k] = g->H[(size_t)k * D + k]; g->H[(size_t)k * D + k] += lambda; }
      const int ok2 = oracle_dense_ldlt_solve(D, g->H, g->Hwork, g->b, g->dx);
      apply_update(g, g->dx);
      /* restoreDiagonal */
      for (int k = 0; k < D; k++) g->H[(size_t)k * D + k] = g->diag_backup[k];
      compute_active_errors(g);
      tempChi = active_robust_chi2(g);
      if (!ok2) tempChi = DBL_MAX;
      rho = (currentChi - tempChi);
      double scale = 0.; /* computeScale :182-189 */
      for (int j = 0; j < D; j++) scale += g->dx[j] * (lambda * g->dx[j] + g->b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = fmin(alpha, goodUpper);
        const double scaleFactor = fmax(goodLower, alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        /* pop */
        memcpy(g->x, xb, sizeof(double) * 3 * g->n_nodes);
        memcpy(g->q, qb, sizeof(qb)); memcpy(g->t, tb, sizeof(tb));
      }
      qmax++;
      st->trials++;
    } while (rho < 0 && qmax < maxTrials);
    if (trace && it < trace_cap) {
      trace[4 * it + 0] = iniChi; trace[4 * it + 1] = lambda_start;
      trace[4 * it + 2] = (double)qmax; trace[4 * it + 3] = currentChi;
    }
    st->chi2_final = currentChi;
    st->lambda = lambda;
    st->iterations = it + 1;
    if (qmax == maxTrials || rho == 0) break; /* Terminate */
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) break;
  }
  free(xb);
}

/* ----------------------------------------------------------- public API -- */

int oracle_sft_solve(const defslam_sft_problem *p, defslam_sft_result *r) {
  if (!p || !r || !p->tmpl_desc || !p->node_xyz) return DEFSLAM_EBADARG;
  Graph g;
  int rc = oracle_graph_build(&g, p);
  if (rc) { oracle_graph_free(&g); return rc; }
  const int n = g.n_nodes;
  /* e->computeError() at edge creation (:350) */
  compute_active_errors(&g);
  LMStats st;
  const int maxit = p->max_iterations > 0 ? p->max_iterations : 50;
  run_lm(&g, maxit, &st, r->trace, r->trace_capacity);

  /* outlier classing :515-537.  e->chi2() uses the error of the LAST
   * computeActiveErrors, i.e. of the last LM trial (rejected or not): the
   * edges are not re-evaluated after a pop() because mvbOutlier[idx] is false. */
  int nBad = 0;
  uint8_t *outl = (uint8_t *)calloc(g.n_rep > 0 ? g.n_rep : 1, 1);
  for (int i = 0; i < g.n_rep; i++) {
    const EdgeReproj *e = &g.rep[i];
    const float chi2 = (float)(e->err[0] * e->info * e->err[0] + e->err[1] * e->info * e->err[1]);
    /* overload: deltaMono < sqrt(a0^2 + a1^2)  :806-818 */
    const int out = g.variant ? (g.huber_delta < sqrt(pow(e->err[0], 2) + pow(e->err[1], 2))) : (chi2 > 5.991);
    if (out) { outl[i] = 1; nBad++; }
  }
  /* mean reprojection error over inliers, errors recomputed at the final
   * estimate :538-559 */
  double sumError = 0.0;
  unsigned cnt = 0;
  for (int i = 0; i < g.n_rep; i++)
    if (!outl[i]) {
      reproj_error(&g, &g.rep[i]);
      sumError += sqrt(pow(g.rep[i].err[0], 2) + pow(g.rep[i].err[1], 2));
      cnt++;
    }
  r->rep_error = (float)(sumError / cnt);
  pose_to_Tcw(g.q, g.t, r->T_cw_out);
  if (r->node_xyz_out) memcpy(r->node_xyz_out, g.x, sizeof(double) * 3 * n); /* updateNodes :955-968 */
  if (r->outlier_out) memcpy(r->outlier_out, outl, g.n_rep);
  if (r->node_role_out)
    for (int v = 0; v < n; v++) r->node_role_out[v] = (uint8_t)(g.viewed[v] | (g.optlap[v] << 1));
  r->n_inliers = g.n_rep - nBad;
  r->lm_iterations = st.iterations;
  r->lm_trials = st.trials;
  r->chi2_initial = st.chi2_initial;
  r->chi2_final = st.chi2_final;
  r->lambda_final = st.lambda;
  r->status = 0;
  free(outl);
  oracle_graph_free(&g);
  return 0;
}

/* H, b, chi2 at the current state, in the ABI variable order of
 * defslam_sft_normal_equations (nodes first, camera last, fixed nodes =
 * identity rows). */
int oracle_sft_normal_equations(const defslam_sft_problem *p, double *H_dense, double *b, double *chi2) {
  if (!p || !p->tmpl_desc) return DEFSLAM_EBADARG;
  Graph g;
  int rc = oracle_graph_build(&g, p);
  if (rc) { oracle_graph_free(&g); return rc; }
  compute_active_errors(&g);
  if (chi2) *chi2 = active_robust_chi2(&g);
  build_system(&g);
  const int n = g.n_nodes, Dabi = 3 * n + 6;
  int *map = (int *)malloc(sizeof(int) * Dabi); /* abi row -> oracle row or -1 */
  for (int v = 0; v < n; v++)
    for (int c = 0; c < 3; c++) map[3 * v + c] = g.idx[v] < 0 ? -1 : g.idx[v] + c;
  for (int c = 0; c < 6; c++) map[3 * n + c] = c;
  for (int i = 0; i < Dabi; i++) {
    if (b) b[i] = map[i] < 0 ? 0.0 : g.b[map[i]];
    if (H_dense)
      for (int j = 0; j < Dabi; j++) {
        double v = 0.0;
        if (map[i] >= 0 && map[j] >= 0) v = g.H[(size_t)map[i] * g.D + map[j]];
        else if (i == j) v = 1.0;
        H_dense[(size_t)i * Dabi + j] = v;
      }
  }
  free(map);
  oracle_graph_free(&g);
  return 0;
}

/* Per-family residual vector + dense Jacobian w.r.t. the ABI variable order,
 * for finite-difference tests.  rows: 2*n_rep, then 3*n_ref, n_curv, n_str.
 * J may be NULL.  Returns number of residual rows (or <0). */
int oracle_sft_residuals(const defslam_sft_problem *p, double *res, double *J, int max_rows) {
  if (!p || !p->tmpl_desc) return DEFSLAM_EBADARG;
  Graph g;
  int rc = oracle_graph_build(&g, p);
  if (rc) { oracle_graph_free(&g); return rc; }
  compute_active_errors(&g);
  const int n = g.n_nodes, Dabi = 3 * n + 6;
  const int rows = 2 * g.n_rep + 3 * g.n_ref + g.n_curv + g.n_str;
  if (rows > max_rows) { oracle_graph_free(&g); return rows; }
  if (J) memset(J, 0, sizeof(double) * (size_t)rows * Dabi);
  int r0 = 0;
  for (int i = 0; i < g.n_rep; i++, r0 += 2) {
    EdgeReproj *e = &g.rep[i];
    res[r0] = e->err[0]; res[r0 + 1] = e->err[1];
    if (J) {
      reproj_linearize(&g, e);
      for (int r = 0; r < 2; r++) {
        for (int c = 0; c < 6; c++) J[(size_t)(r0 + r) * Dabi + 3 * n + c] = e->Jc[r * 6 + c];
        for (int k = 0; k < 3; k++)
          for (int c = 0; c < 3; c++) J[(size_t)(r0 + r) * Dabi + 3 * e->v[k] + c] += e->Jn[k][r * 3 + c];
      }
    }
  }
  for (int i = 0; i < g.n_ref; i++, r0 += 3)
    for (int c = 0; c < 3; c++) {
      res[r0 + c] = g.ref[i].err[c];
      if (J) J[(size_t)(r0 + c) * Dabi + 3 * g.ref[i].v + c] = 1.0;
    }
  for (int i = 0; i < g.n_curv; i++, r0++) {
    EdgeCurv *e = &g.curv[i];
    res[r0] = e->err;
    if (J) {
      curv_linearize(e);
      for (int a = 0; a < e->nv; a++)
        for (int c = 0; c < 3; c++) J[(size_t)r0 * Dabi + 3 * e->v[a] + c] += e->J[3 * a + c];
    }
  }
  for (int i = 0; i < g.n_str; i++, r0++) {
    EdgeStretch *e = &g.str[i];
    res[r0] = e->err;
    if (J) {
      stretch_linearize(&g, e);
      for (int c = 0; c < 3; c++) {
        J[(size_t)r0 * Dabi + 3 * e->a + c] += e->Ja[c];
        J[(size_t)r0 * Dabi + 3 * e->b + c] -= e->Ja[c];
      }
    }
  }
  oracle_graph_free(&g);
  return rows;
}

/* Apply an update vector (ABI order) to the state: nodes += d, pose = exp(dc)*pose.
 * Used by the finite-difference tests to perturb exactly as the solver does. */
int oracle_sft_apply_update(const defslam_sft_problem *p, const double *d, double *node_xyz_out, float *T_cw_out,
                            double *q_out, double *t_out) {
  const int n = p->tmpl_desc->n_nodes;
  double q[4], t[3];
  pose_from_Tcw(p->T_cw, q, t);
  pose_oplus(q, t, &d[3 * n]);
  for (int i = 0; i < 3 * n; i++) node_xyz_out[i] = p->node_xyz[i] + d[i];
  if (T_cw_out) pose_to_Tcw(q, t, T_cw_out);
  if (q_out) memcpy(q_out, q, sizeof(q));
  if (t_out) memcpy(t_out, t, sizeof(t));
  return 0;
}

/* residuals with an explicit fp64 pose (q,t) instead of the f32 T_cw, so that
 * finite differences are not quantised by the f32 pose */
int oracle_sft_residuals_pose(const defslam_sft_problem *p, const double *q, const double *t, const double *node_xyz,
                              double *res, int max_rows) {
  Graph g;
  int rc = oracle_graph_build(&g, p);
  if (rc) { oracle_graph_free(&g); return rc; }
  memcpy(g.q, q, sizeof(g.q)); memcpy(g.t, t, sizeof(g.t));
  memcpy(g.x, node_xyz, sizeof(double) * 3 * g.n_nodes);
  compute_active_errors(&g);
  const int rows = 2 * g.n_rep + 3 * g.n_ref + g.n_curv + g.n_str;
  if (rows > max_rows) { oracle_graph_free(&g); return rows; }
  int r0 = 0;
  for (int i = 0; i < g.n_rep; i++) { res[r0++] = g.rep[i].err[0]; res[r0++] = g.rep[i].err[1]; }
  for (int i = 0; i < g.n_ref; i++) for (int c = 0; c < 3; c++) res[r0++] = g.ref[i].err[c];
  for (int i = 0; i < g.n_curv; i++) res[r0++] = g.curv[i].err;
  for (int i = 0; i < g.n_str; i++) res[r0++] = g.str[i].err;
  oracle_graph_free(&g);
  return rows;
}

/* DefMapPoint::RecalculatePosition  Modules/Common/DefMapPoint.cc:129-147 */
int oracle_mappoints_recalculate(int32_t n_nodes, const double *node_xyz, int32_t n_points,
                                 const int32_t *point_nodes, const double *point_bary, float *out) {
  (void)n_nodes;
  for (int i = 0; i < n_points; i++)
    for (int c = 0; c < 3; c++)
      out[3 * i + c] = (float)(point_bary[3 * i] * node_xyz[3 * point_nodes[3 * i] + c] +
                               point_bary[3 * i + 1] * node_xyz[3 * point_nodes[3 * i + 1] + c] +
                               point_bary[3 * i + 2] * node_xyz[3 * point_nodes[3 * i + 2] + c]);
  return 0;
}
