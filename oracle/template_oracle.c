/*
 * template_oracle.c -- CPU restatement of the template (mesh) construction the
 * SfT solve depends on.  TEST INFRASTRUCTURE ONLY (see sft_oracle.c header).
 * The Laplacian part (mean-value weights, boundary flags, kappa0) is PINNED TO THE REFERENCE'S OWN CODE:
 * LaplacianMesh.cc:53-148,157-162 are compiled by oracle/Makefile behind oracle/template_ref_harness.cc and
 * tests/test_oracle_sft_ref.py compares this file with them bit for bit (regular and Delaunay meshes, live and
 * through tests/golden/sft_ref.npz).  Embedding (pointInTriangle) and triangulation remain restatements
 * (they need OpenCV types); no reference tests/fixtures exist for them.
 *
 * Follows (paths under the DefSLAM tree):
 *   TriangularMesh::regularTriangulation     Modules/Template/TriangularMesh.cc:92-107
 *   Facet ctor / Edge ctor (edge discovery)  Modules/Template/Facet.cc:32-62, Edge.cc:29-59
 *   Node::GetNeighbours / distanceto         Modules/Template/Node.cc:114-129,71-76
 *   LaplacianMesh::ExtractMeanCurvatures     Modules/Template/LaplacianMesh.cc:53-148
 *   Template::getEdgeMeanSize (median)       Modules/Template/Template.cc:158-175
 *   calculateFeaturesCoordinates/pointInTriangle  TriangularMesh.cc:133-236
 *
 * Order conventions (the reference orders by heap address, quirk C13): nodes
 * by index, neighbours ascending, facets by index, edges in discovery order
 * with (a,b) = (min,max).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sft_oracle.h"

/* TriangularMesh.cc:92-107 */
int oracle_regular_triangulation(int nodesVer, int nodesHor, int32_t *facets) {
  int f = 0;
  for (int j = 0; j < nodesHor - 1; j++)
    for (int i = 0; i < nodesVer - 1; i++) {
      facets[3 * f + 0] = i + nodesHor * j;
      facets[3 * f + 1] = i + nodesHor * j + 1;
      facets[3 * f + 2] = (nodesHor * (j + 1)) + i;
      f++;
      facets[3 * f + 0] = i + nodesHor * j + 1;
      facets[3 * f + 1] = (nodesHor * (j + 1)) + i;
      facets[3 * f + 2] = (nodesHor * (j + 1)) + i + 1;
      f++;
    }
  return f;
}

static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}
static int cmp_dbl(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

static int ring_has(const int32_t *idx, int cnt, int v) {
  for (int k = 0; k < cnt; k++)
    if (idx[k] == v) return 1;
  return 0;
}

int oracle_mesh_laplacian(int32_t n, const double *X, int32_t nf, const int32_t *facets, int32_t max_ring,
                          int32_t *nbr_cnt, int32_t *nbr_idx, double *nbr_w, uint8_t *boundary, double *kappa0,
                          int32_t *n_edges_out, int32_t *edge_ab, double *edge_len0, double *edge_median_len) {
  memset(nbr_cnt, 0, sizeof(int32_t) * n);
  memset(boundary, 0, n);
  memset(kappa0, 0, sizeof(double) * n);
  for (int i = 0; i < n * max_ring; i++) { nbr_idx[i] = -1; nbr_w[i] = 0.0; }
  /* Facet ctor: edges (v1,v2),(v2,v3),(v1,v3) unless already present (Facet.cc:46-58) */
  int ne = 0;
  for (int f = 0; f < nf; f++) {
    const int v[3] = {facets[3 * f], facets[3 * f + 1], facets[3 * f + 2]};
    const int pr[3][2] = {{v[0], v[1]}, {v[1], v[2]}, {v[0], v[2]}};
    for (int e = 0; e < 3; e++) {
      const int a = pr[e][0] < pr[e][1] ? pr[e][0] : pr[e][1];
      const int b = pr[e][0] < pr[e][1] ? pr[e][1] : pr[e][0];
      if (a < 0 || b >= n || a == b) return DEFSLAM_EBADARG;
      if (ring_has(&nbr_idx[a * max_ring], nbr_cnt[a], b)) continue;
      if (nbr_cnt[a] >= max_ring || nbr_cnt[b] >= max_ring) return DEFSLAM_ETOOLARGE;
      nbr_idx[a * max_ring + nbr_cnt[a]++] = b;
      nbr_idx[b * max_ring + nbr_cnt[b]++] = a;
      edge_ab[2 * ne] = a;
      edge_ab[2 * ne + 1] = b;
      /* InitialDist = v1->distanceto(v2)  Edge.cc:52, Node.cc:71-76 */
      edge_len0[ne] = sqrt(pow(X[3 * a] - X[3 * b], 2) + pow(X[3 * a + 1] - X[3 * b + 1], 2) +
                           pow(X[3 * a + 2] - X[3 * b + 2], 2));
      ne++;
    }
  }
  *n_edges_out = ne;
  for (int v = 0; v < n; v++) qsort(&nbr_idx[v * max_ring], nbr_cnt[v], sizeof(int32_t), cmp_int);

  /* Template::getEdgeMeanSize: sorted lengths, element size/2 (Template.cc:158-175) */
  if (ne > 0) {
    double *d = (double *)malloc(sizeof(double) * ne);
    memcpy(d, edge_len0, sizeof(double) * ne);
    qsort(d, ne, sizeof(double), cmp_dbl);
    *edge_median_len = d[ne / 2];
    free(d);
  } else {
    *edge_median_len = 0.10;
  }

  /* ExtractMeanCurvatures, first loop (LaplacianMesh.cc:55-115) */
  for (int i = 0; i < n; i++) {
    const double *Ni = &X[3 * i];
    for (int kj = 0; kj < nbr_cnt[i]; kj++) {
      const int j = nbr_idx[i * max_ring + kj];
      const double *Nj = &X[3 * j];
      /* neighbours of j that are also neighbours of i, ascending */
      int com[2], nc = 0;
      for (int kk = 0; kk < nbr_cnt[j]; kk++) {
        const int c = nbr_idx[j * max_ring + kk];
        if (ring_has(&nbr_idx[i * max_ring], nbr_cnt[i], c)) {
          if (nc < 2) com[nc] = c;
          nc++;
        }
      }
      if (nc == 0) {
        /* (*ite)->setBadFlag(): does not occur on a triangulated mesh; ignored */
      } else if (nc == 1) {
        boundary[j] = 1; /* (*ite)->setBoundary() -- flags j, not i (:90-93) */
      } else {
        const double *Nj1 = &X[3 * com[0]], *Nj_1 = &X[3 * com[1]];
        double a1[3], a2[3], bj[3];
        for (int c = 0; c < 3; c++) { a1[c] = Nj_1[c] - Ni[c]; a2[c] = Nj1[c] - Ni[c]; bj[c] = Nj[c] - Ni[c]; }
        const double c1[3] = {a1[1] * bj[2] - a1[2] * bj[1], a1[2] * bj[0] - a1[0] * bj[2], a1[0] * bj[1] - a1[1] * bj[0]};
        const double c2[3] = {a2[1] * bj[2] - a2[2] * bj[1], a2[2] * bj[0] - a2[0] * bj[2], a2[0] * bj[1] - a2[1] * bj[0]};
        const double tn1 = sqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]) / (a1[0] * bj[0] + a1[1] * bj[1] + a1[2] * bj[2]);
        const double tn2 = sqrt(c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2]) / (a2[0] * bj[0] + a2[1] * bj[1] + a2[2] * bj[2]);
        const double nij = sqrt(bj[0] * bj[0] + bj[1] * bj[1] + bj[2] * bj[2]);
        nbr_w[i * max_ring + kj] = (tan(fabs(atan(tn1)) / 2) + tan(fabs(atan(tn2)) / 2)) / nij; /* :104-112 */
      }
    }
  }
  /* second loop: Laplacian coordinates of non-boundary nodes (:117-147) */
  for (int i = 0; i < n; i++) {
    if (boundary[i] || nbr_cnt[i] <= 1) continue;
    double L[3] = {0, 0, 0}, sw = 0.0;
    for (int kj = 0; kj < nbr_cnt[i]; kj++) {
      const int j = nbr_idx[i * max_ring + kj];
      const double w = nbr_w[i * max_ring + kj];
      for (int c = 0; c < 3; c++) L[c] = L[c] + w * X[3 * j + c];
      sw = sw + w;
    }
    double d[3];
    for (int c = 0; c < 3; c++) d[c] = X[3 * i + c] - (L[c] / sw);
    kappa0[i] = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); /* GetMeanCurvatureInitial :157-162 */
  }
  return 0;
}

/* TriangularMesh::pointInTriangle  TriangularMesh.cc:207-236 (all fp32) */
static int point_in_triangle(const float q[3], const float p0[3], const float p1[3], const float p2[3], float bary[3]) {
  float u[3], v[3], w[3], nrm[3];
  for (int c = 0; c < 3; c++) { u[c] = p1[c] - p0[c]; v[c] = p2[c] - p0[c]; w[c] = q[c] - p0[c]; }
  nrm[0] = u[1] * v[2] - u[2] * v[1]; nrm[1] = u[2] * v[0] - u[0] * v[2]; nrm[2] = u[0] * v[1] - u[1] * v[0];
  const float uw[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
  const float wv[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
  const float nn = nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2];
  const float gamma = (uw[0] * nrm[0] + uw[1] * nrm[1] + uw[2] * nrm[2]) / nn;
  const float beta = (wv[0] * nrm[0] + wv[1] * nrm[1] + wv[2] * nrm[2]) / nn;
  const float alpha = 1 - gamma - beta;
  bary[0] = alpha; bary[1] = beta; bary[2] = gamma;
  float d2 = 0.f;
  for (int c = 0; c < 3; c++) {
    const float np = p0[c] * alpha + p1[c] * beta + p2[c] * gamma;
    d2 += (np - q[c]) * (np - q[c]);
  }
  if (d2 > 1E-1) return 0;
  return (0 <= alpha) && (alpha <= 1) && (0 <= beta) && (beta <= 1) && (0 <= gamma) && (gamma <= 1);
}

/* TriangularMesh::calculateFeaturesCoordinates  TriangularMesh.cc:133-200 */
int oracle_embed_points(int32_t n, const double *X, int32_t nf, const int32_t *facets, int32_t npts,
                        const float *P, int32_t *out_facet, int32_t *out_nodes, float *out_bary) {
  for (int i = 0; i < npts; i++) {
    out_facet[i] = -1;
    for (int c = 0; c < 3; c++) { out_nodes[3 * i + c] = -1; out_bary[3 * i + c] = 0.f; }
    const float *mp = &P[3 * i];
    int closest = -1;
    double bestdist = 100;
    for (int v = 0; v < n; v++) {
      const double dist = sqrt(pow(X[3 * v] - mp[0], 2) + pow(X[3 * v + 1] - mp[1], 2) + pow(X[3 * v + 2] - mp[2], 2));
      if (dist < bestdist) { closest = v; bestdist = dist; }
    }
    if (closest < 0) continue;
    for (int f = 0; f < nf; f++) {
      int v[3] = {facets[3 * f], facets[3 * f + 1], facets[3 * f + 2]};
      if (v[0] != closest && v[1] != closest && v[2] != closest) continue;
      qsort(v, 3, sizeof(int), cmp_int); /* Facet::getNodes() is a std::set */
      float p[3][3], bary[3];
      for (int k = 0; k < 3; k++)
        for (int c = 0; c < 3; c++) p[k][c] = (float)X[3 * v[k] + c];
      if (point_in_triangle(mp, p[0], p[1], p[2], bary)) {
        out_facet[i] = f;
        for (int k = 0; k < 3; k++) { out_nodes[3 * i + k] = v[k]; out_bary[3 * i + k] = bary[k]; }
        break;
      }
    }
  }
  return 0;
}
