/* Internal to oracle/: the graph the oracle builds from a defslam_sft_problem (the restatement of
 * DefOptimizer.cc:251-507), shared with the reference-pinning harness (g2o_ref_harness.cc) so that the
 * reference's own edge classes are fed exactly the edges, weights and vertex ids the oracle uses.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef DEFSLAM_ORACLE_GRAPH_H_
#define DEFSLAM_ORACLE_GRAPH_H_
#include <stdint.h>
#include "../include/defslam_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int m;          /* match index */
  int v[3];       /* node ids */
  double bary[3]; /* barycentrics */
  double obs[2];
  double info;    /* invSigma2 / N */
  double err[2];
  double Jc[12];    /* 2x6 */
  double Jn[3][6];  /* 2x3 each */
} EdgeReproj;

typedef struct {
  int v;
  double meas[3];
  double err[3];
} EdgeRef;

typedef struct {
  int nv;        /* 1 + #neighbours */
  int *v;        /* v[0] centre, then neighbours */
  double *w;     /* weights, same order as neighbours */
  double len;    /* lenghtEdge_ */
  double kappa0; /* measurement */
  double err;
  double mc[3], mcn, sumw; /* meanCurvature_, its norm, sumWeights_ */
  double *J;     /* [nv*3] */
} EdgeCurv;

typedef struct {
  int a, b;
  double len0;
  double err;
  double Ja[3];
} EdgeStretch;

typedef struct {
  /* sizes */
  int n_nodes, n_matches;
  /* state */
  double q[4], t[3];
  double *x; /* [n*3] */
  /* camera */
  double fx, fy, cx, cy;
  /* free-variable map: idx[v] = first dense row of node v, -1 if fixed.
   * camera occupies dense rows 0..5 (g2o: vertex id 0 first) */
  int *idx;
  int D;
  /* edges */
  int n_rep, n_ref, n_curv, n_str;
  EdgeReproj *rep;
  EdgeRef *ref;
  EdgeCurv *curv;
  EdgeStretch *str;
  double info_ref, info_curv, info_str;
  double huber_delta, huber_dsqr;
  uint8_t *viewed, *optlap;
  int n_optlap, n_viewed;
  int variant;    /* 1: the matches-given overload (DefOptimizer.cc:582-837) */
  int n_curv_den; /* denominator of the curvature information: |OptLap| (= |Viewed| in the overload) */
  /* solver */
  double *H, *b, *dx, *Hwork, *diag_backup;
} Graph;

int oracle_graph_build(Graph *g, const defslam_sft_problem *p);
void oracle_graph_free(Graph *g);
/* dense LDL^T used by the oracle (no pivoting); returns 1 if every pivot is positive */
int oracle_dense_ldlt_solve(int D, const double *H, double *L, const double *b, double *x);

#ifdef __cplusplus
}
#endif
#endif
