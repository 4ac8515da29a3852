/*
 * mesh_core.h -- template (mesh) construction on the device.
 *
 * Replaces:
 *   Facet/Edge constructors (edge discovery)      Modules/Template/Facet.cc:32-62, Edge.cc:29-59
 *   Node::GetNeighbours                           Modules/Template/Node.cc:114-129
 *   LaplacianMesh::ExtractMeanCurvatures          Modules/Template/LaplacianMesh.cc:53-148
 *   Template::getEdgeMeanSize (the median)        Modules/Template/Template.cc:158-175
 *   TriangularMesh::calculateFeaturesCoordinates  Modules/Template/TriangularMesh.cc:133-200
 *   TriangularMesh::pointInTriangle (fp32)        Modules/Template/TriangularMesh.cc:207-236
 *   DefMapPoint::RecalculatePosition              Modules/Common/DefMapPoint.cc:129-147
 *   Surface::getVertex                            Modules/Mapping/Surface.cc:125-161
 *
 * Order conventions (the reference orders by heap address, quirk C13): nodes by
 * index, neighbours ascending, edges in discovery order (facet order, edges
 * (v1,v2),(v2,v3),(v1,v3) of each facet) stored as (min,max).
 */
#ifndef DS_MESH_CORE_H_
#define DS_MESH_CORE_H_
#include "bbs_core.h"
#include "ds_common.h"

namespace ds {

struct MeshLapArgs {
  int n, nf, max_ring;
  const double *X;     /* [3n]  */
  const int *facets;   /* [3nf] */
  int *nbr_cnt;        /* [n] */
  int *nbr_idx;        /* [n*max_ring] */
  double *nbr_w;       /* [n*max_ring] */
  uint8_t *boundary;   /* [n] */
  double *kappa0;      /* [n] */
  int *n_edges;        /* [1] */
  int *edge_ab;        /* [2*3nf] */
  double *edge_len0;   /* [3nf] */
  double *median;      /* [1] */
  int *status;         /* [1] */
  int *cand_first;     /* scratch [3nf] */
  int *cand_pos;       /* scratch [3nf] */
};

DS_FN void cand_pair(const int *facets, int c, int &a, int &b) {
  const int f = c / 3, e = c - 3 * f;
  const int v0 = facets[3 * f], v1 = facets[3 * f + 1], v2 = facets[3 * f + 2];
  const int p = e == 0 ? v0 : (e == 1 ? v1 : v0);
  const int q = e == 0 ? v1 : (e == 1 ? v2 : v2);
  a = p < q ? p : q;
  b = p < q ? q : p;
}

DS_FN bool ring_has(const int *idx, int cnt, int v) {
  for (int k = 0; k < cnt; k++)
    if (idx[k] == v) return true;
  return false;
}

/* one CTA */
DS_FN void mesh_laplacian_team(const Team &team, const MeshLapArgs &A, double *red) {
  const int n = A.n, nf = A.nf, C = 3 * nf, R = A.max_ring;
  if (team.tid == 0) *A.status = 0;
  DS_FOR(v, n) { A.nbr_cnt[v] = 0; A.boundary[v] = 0; A.kappa0[v] = 0.0; }
  DS_FOR(i, n * R) { A.nbr_idx[i] = -1; A.nbr_w[i] = 0.0; }
  team.sync();
  /* edge discovery: a candidate is an edge iff no earlier candidate joins the same nodes */
  int bad = 0;
  DS_FOR(c, C) {
    int a, b;
    cand_pair(A.facets, c, a, b);
    if (a < 0 || b >= n || a == b) bad = 1;
    int first = 1;
    for (int c2 = 0; c2 < c && first; c2++) {
      int a2, b2;
      cand_pair(A.facets, c2, a2, b2);
      if (a2 == a && b2 == b) first = 0;
    }
    A.cand_first[c] = first;
  }
  bad = team_sum_int(team, bad, red);
  if (bad) { if (team.tid == 0) *A.status = DEFSLAM_EBADARG; return; }
  if (team.tid == 0) {
    int s = 0;
    for (int c = 0; c < C; c++) { A.cand_pos[c] = s; s += A.cand_first[c]; }
    *A.n_edges = s;
  }
  team.sync();
  const int ne = *A.n_edges;
  DS_FOR(c, C) {
    if (!A.cand_first[c]) continue;
    int a, b;
    cand_pair(A.facets, c, a, b);
    const int e = A.cand_pos[c];
    A.edge_ab[2 * e] = a;
    A.edge_ab[2 * e + 1] = b;
    /* Edge::InitialDist = v1->distanceto(v2)  Edge.cc:52, Node.cc:71-76 */
    A.edge_len0[e] = sqrt(pow(A.X[3 * a] - A.X[3 * b], 2) + pow(A.X[3 * a + 1] - A.X[3 * b + 1], 2) +
                          pow(A.X[3 * a + 2] - A.X[3 * b + 2], 2));
  }
  team.sync();
  /* 1-rings, ascending */
  int over = 0;
  DS_FOR(v, n) {
    int cnt = 0;
    int *ring = &A.nbr_idx[v * R];
    for (int e = 0; e < ne; e++) {
      const int a = A.edge_ab[2 * e], b = A.edge_ab[2 * e + 1];
      const int o = a == v ? b : (b == v ? a : -1);
      if (o < 0) continue;
      if (cnt >= R) { over = 1; break; }
      int j = cnt - 1;
      while (j >= 0 && ring[j] > o) { ring[j + 1] = ring[j]; j--; }
      ring[j + 1] = o;
      cnt++;
    }
    A.nbr_cnt[v] = cnt;
  }
  over = team_sum_int(team, over, red);
  if (over) { if (team.tid == 0) *A.status = DEFSLAM_ETOOLARGE; return; }
  /* median edge length: element ne/2 of the sorted lengths (Template.cc:158-175) */
  if (ne == 0 && team.tid == 0) *A.median = 0.10;
  DS_FOR(e, ne) {
    const double l = A.edge_len0[e];
    int rank = 0;
    for (int e2 = 0; e2 < ne; e2++) {
      const double l2 = A.edge_len0[e2];
      rank += (l2 < l || (l2 == l && e2 < e)) ? 1 : 0;
    }
    if (rank == ne / 2) *A.median = l;
  }
  /* mean-value weights and boundary flags (LaplacianMesh.cc:55-115) */
  DS_FOR(i, n) {
    const double *Ni = &A.X[3 * i];
    const int *ri = &A.nbr_idx[i * R];
    const int ci = A.nbr_cnt[i];
    for (int kj = 0; kj < ci; kj++) {
      const int j = ri[kj];
      const double *Nj = &A.X[3 * j];
      int com[2] = {0, 0}, nc = 0;
      for (int kk = 0; kk < A.nbr_cnt[j]; kk++) {
        const int c = A.nbr_idx[j * R + kk];
        if (ring_has(ri, ci, c)) {
          if (nc < 2) com[nc] = c;
          nc++;
        }
      }
      if (nc == 1) {
        A.boundary[j] = 1; /* flags the NEIGHBOUR (:90-93); same value from every writer */
      } else if (nc >= 2) {
        const double *Nj1 = &A.X[3 * com[0]], *Nj_1 = &A.X[3 * com[1]];
        double a1[3], a2[3], bj[3];
        for (int c = 0; c < 3; c++) { a1[c] = Nj_1[c] - Ni[c]; a2[c] = Nj1[c] - Ni[c]; bj[c] = Nj[c] - Ni[c]; }
        const double c1[3] = {a1[1] * bj[2] - a1[2] * bj[1], a1[2] * bj[0] - a1[0] * bj[2], a1[0] * bj[1] - a1[1] * bj[0]};
        const double c2[3] = {a2[1] * bj[2] - a2[2] * bj[1], a2[2] * bj[0] - a2[0] * bj[2], a2[0] * bj[1] - a2[1] * bj[0]};
        const double tn1 = sqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]) / (a1[0] * bj[0] + a1[1] * bj[1] + a1[2] * bj[2]);
        const double tn2 = sqrt(c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2]) / (a2[0] * bj[0] + a2[1] * bj[1] + a2[2] * bj[2]);
        const double nij = sqrt(bj[0] * bj[0] + bj[1] * bj[1] + bj[2] * bj[2]);
        A.nbr_w[i * R + kj] = (tan(fabs(atan(tn1)) / 2) + tan(fabs(atan(tn2)) / 2)) / nij;
      }
    }
  }
  team.sync();
  /* Laplacian coordinates of the non-boundary nodes (:117-147) */
  DS_FOR(i, n) {
    if (A.boundary[i] || A.nbr_cnt[i] <= 1) continue;
    double L0 = 0, L1 = 0, L2 = 0, sw = 0.0;
    for (int kj = 0; kj < A.nbr_cnt[i]; kj++) {
      const int j = A.nbr_idx[i * R + kj];
      const double w = A.nbr_w[i * R + kj];
      L0 = L0 + w * A.X[3 * j]; L1 = L1 + w * A.X[3 * j + 1]; L2 = L2 + w * A.X[3 * j + 2];
      sw = sw + w;
    }
    const double d0 = A.X[3 * i] - (L0 / sw), d1 = A.X[3 * i + 1] - (L1 / sw), d2 = A.X[3 * i + 2] - (L2 / sw);
    A.kappa0[i] = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
  }
}

/* TriangularMesh::pointInTriangle, all fp32 (TriangularMesh.cc:207-236).  The
 * explicit __f*_rn forms keep nvcc from contracting into FMAs, so that the
 * barycentrics are bit-identical to a plain fp32 evaluation. */
#if DS_CUDA
#define DS_FMUL(a, b) __fmul_rn((a), (b))
#define DS_FADD(a, b) __fadd_rn((a), (b))
#define DS_FSUB(a, b) __fsub_rn((a), (b))
#else
#define DS_FMUL(a, b) ((float)((a) * (b)))
#define DS_FADD(a, b) ((float)((a) + (b)))
#define DS_FSUB(a, b) ((float)((a) - (b)))
#endif

DS_FN float dot3f(const float a[3], const float b[3]) {
  return DS_FADD(DS_FADD(DS_FMUL(a[0], b[0]), DS_FMUL(a[1], b[1])), DS_FMUL(a[2], b[2]));
}
DS_FN void cross3f(const float a[3], const float b[3], float o[3]) {
  o[0] = DS_FSUB(DS_FMUL(a[1], b[2]), DS_FMUL(a[2], b[1]));
  o[1] = DS_FSUB(DS_FMUL(a[2], b[0]), DS_FMUL(a[0], b[2]));
  o[2] = DS_FSUB(DS_FMUL(a[0], b[1]), DS_FMUL(a[1], b[0]));
}

DS_FN bool point_in_triangle(const float q[3], const float p0[3], const float p1[3], const float p2[3], float bary[3]) {
  float u[3], v[3], w[3], nrm[3], uw[3], wv[3];
  for (int c = 0; c < 3; c++) { u[c] = DS_FSUB(p1[c], p0[c]); v[c] = DS_FSUB(p2[c], p0[c]); w[c] = DS_FSUB(q[c], p0[c]); }
  cross3f(u, v, nrm);
  cross3f(u, w, uw);
  cross3f(w, v, wv);
  const float nn = dot3f(nrm, nrm);
  const float gamma = dot3f(uw, nrm) / nn;
  const float beta = dot3f(wv, nrm) / nn;
  const float alpha = DS_FSUB(DS_FSUB(1.f, gamma), beta);
  bary[0] = alpha; bary[1] = beta; bary[2] = gamma;
  float d2 = 0.f;
  for (int c = 0; c < 3; c++) {
    const float np = DS_FADD(DS_FADD(DS_FMUL(p0[c], alpha), DS_FMUL(p1[c], beta)), DS_FMUL(p2[c], gamma));
    const float df = DS_FSUB(np, q[c]);
    d2 = DS_FADD(d2, DS_FMUL(df, df));
  }
  if (d2 > 1E-1) return false;
  return (0 <= alpha) && (alpha <= 1) && (0 <= beta) && (beta <= 1) && (0 <= gamma) && (gamma <= 1);
}

/* one map point */
DS_FN void embed_point(int n, const double *X, int nf, const int *facets, const float *mp, int *out_facet,
                       int *out_nodes, float *out_bary) {
  *out_facet = -1;
  for (int c = 0; c < 3; c++) { out_nodes[c] = -1; out_bary[c] = 0.f; }
  int closest = -1;
  double best = 100;
  for (int v = 0; v < n; v++) {
    const double dist = sqrt(pow(X[3 * v] - mp[0], 2) + pow(X[3 * v + 1] - mp[1], 2) + pow(X[3 * v + 2] - mp[2], 2));
    if (dist < best) { closest = v; best = dist; }
  }
  if (closest < 0) return;
  for (int f = 0; f < nf; f++) {
    int v[3] = {facets[3 * f], facets[3 * f + 1], facets[3 * f + 2]};
    if (v[0] != closest && v[1] != closest && v[2] != closest) continue;
    /* ascending order: Facet::getNodes() is a std::set */
    if (v[0] > v[1]) { const int t = v[0]; v[0] = v[1]; v[1] = t; }
    if (v[1] > v[2]) { const int t = v[1]; v[1] = v[2]; v[2] = t; }
    if (v[0] > v[1]) { const int t = v[0]; v[0] = v[1]; v[1] = t; }
    float p[3][3], bary[3];
    for (int k = 0; k < 3; k++)
      for (int c = 0; c < 3; c++) p[k][c] = (float)X[3 * v[k] + c];
    if (point_in_triangle(mp, p[0], p[1], p[2], bary)) {
      *out_facet = f;
      for (int k = 0; k < 3; k++) { out_nodes[k] = v[k]; out_bary[k] = bary[k]; }
      return;
    }
  }
}

/* DefMapPoint::RecalculatePosition: fp64 combination stored fp32 */
DS_FN void mappoint_position(const double *X, const int *nodes, const double *bary, float *out) {
  for (int c = 0; c < 3; c++)
    out[c] = (float)(bary[0] * X[3 * nodes[0] + c] + bary[1] * X[3 * nodes[1] + c] + bary[2] * X[3 * nodes[2] + c]);
}

/* Surface::getVertex (Surface.cc:125-161): sample the depth spline on an
 * xs x ys grid inset by 0.03, x (u) outer, y (v) inner; fp32 output */
DS_FN void surface_vertex(const BbsView &s, const double *ctrl, int xs, int ys, int idx, float *out) {
  const int x = idx / ys, y = idx - x * ys;
  const double t = 0.03;
  const double u = ((s.umax - s.umin - 2 * t) * x) / (xs - 1) + (s.umin + t);
  const double v = ((s.vmax - s.vmin - 2 * t) * y) / (ys - 1) + (s.vmin + t);
  double d;
  bbs_eval_site(s, ctrl, u, v, 0, 0, &d);
  out[0] = (float)(u * d);
  out[1] = (float)(v * d);
  out[2] = (float)d;
}

}  // namespace ds
#endif
