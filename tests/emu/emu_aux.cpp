/*
 * emu_aux.cpp -- TEST INFRASTRUCTURE ONLY (see emu_sft.cpp).  Host-compiled copies of
 * the device functions in mesh_core.h / bbs_core.h behind the same argument lists as the
 * C ABI, so that the CPU-only test tier can check them against the oracle and against the
 * reference's own bbs.cc.
 */
#define DS_EMULATE 1
#include <vector>

#include "../../include/defslam_b200.h"
#include "../../defslam_b200/csrc/mesh_core.h"
#include "../../defslam_b200/csrc/newpts_core.h"
#include "../../defslam_b200/csrc/match_core.h"

using namespace ds;

static BbsView view_of(const defslam_bbs *b) {
  BbsView s;
  s.umin = b->umin; s.umax = b->umax; s.vmin = b->vmin; s.vmax = b->vmax;
  s.nptsu = b->nptsu; s.nptsv = b->nptsv; s.valdim = b->valdim;
  return s;
}

extern "C" {

int emu_mesh_laplacian(int32_t n_nodes, const double *node_xyz, int32_t n_facets, const int32_t *facets,
                       int32_t max_ring, int32_t *nbr_cnt, int32_t *nbr_idx, double *nbr_w, uint8_t *node_boundary,
                       double *node_kappa0, int32_t *n_edges_out, int32_t *edge_ab, double *edge_len0,
                       double *edge_median_len) {
  std::vector<int> cf(3 * (size_t)n_facets + 1), cp(3 * (size_t)n_facets + 1);
  int status = 0;
  double red[40];
  MeshLapArgs A;
  A.n = n_nodes; A.nf = n_facets; A.max_ring = max_ring;
  A.X = node_xyz; A.facets = facets;
  A.nbr_cnt = nbr_cnt; A.nbr_idx = nbr_idx; A.nbr_w = nbr_w; A.boundary = node_boundary; A.kappa0 = node_kappa0;
  A.n_edges = n_edges_out; A.edge_ab = edge_ab; A.edge_len0 = edge_len0; A.median = edge_median_len;
  A.status = &status; A.cand_first = cf.data(); A.cand_pos = cp.data();
  Team team;
  team.tid = 0; team.nthr = 1;
  mesh_laplacian_team(team, A, red);
  return status;
}

int emu_embed_points(int32_t n_nodes, const double *node_xyz, int32_t n_facets, const int32_t *facets,
                     int32_t n_points, const float *point_xyz, int32_t *out_facet, int32_t *out_nodes,
                     float *out_bary) {
  for (int i = 0; i < n_points; i++)
    embed_point(n_nodes, node_xyz, n_facets, facets, &point_xyz[3 * i], &out_facet[i], &out_nodes[3 * i],
                &out_bary[3 * i]);
  return 0;
}

int emu_mappoints_recalculate(int32_t n_nodes, const double *node_xyz, int32_t n_points, const int32_t *point_nodes,
                              const double *point_bary, float *out) {
  (void)n_nodes;
  for (int i = 0; i < n_points; i++) mappoint_position(node_xyz, &point_nodes[3 * i], &point_bary[3 * i], &out[3 * i]);
  return 0;
}

int emu_bbs_eval(const defslam_bbs *bbs, const double *ctrl, int32_t nsites, const double *u, const double *v,
                 int32_t du, int32_t dv, double *val) {
  const BbsView s = view_of(bbs);
  for (int i = 0; i < nsites; i++) bbs_eval_site(s, ctrl, u[i], v[i], du, dv, &val[(size_t)i * s.valdim]);
  return 0;
}

int emu_bbs_coloc(const defslam_bbs *bbs, int32_t nsites, const double *u, const double *v, int32_t du, int32_t dv,
                  double *Cm) {
  const BbsView s = view_of(bbs);
  const size_t NC = (size_t)s.nptsu * s.nptsv;
  for (size_t i = 0; i < (size_t)nsites * NC; i++) Cm[i] = 0.0;
  int bad = 0;
  for (int i = 0; i < nsites; i++)
    if (!bbs_coloc_row(s, u[i], v[i], du, dv, &Cm[(size_t)i * NC])) bad = 1;
  if (bad) /* like the ABI wrapper: the reference aborts leaving an empty matrix */
    for (size_t i = 0; i < (size_t)nsites * NC; i++) Cm[i] = 0.0;
  return bad ? DEFSLAM_EBADARG : 0;
}

int emu_bbs_bending(const defslam_bbs *bbs, double *B) {
  const BbsView s = view_of(bbs);
  const int NC = s.nptsu * s.nptsv;
  for (int i = 0; i < NC; i++)
    for (int j = 0; j < NC; j++) B[(size_t)i * NC + j] = bbs_bending_entry(s, i, j);
  return 0;
}

int emu_surface_vertices(const defslam_bbs *bbs, const double *ctrl, int32_t xs, int32_t ys, float *out) {
  const BbsView s = view_of(bbs);
  for (int i = 0; i < xs * ys; i++) surface_vertex(s, ctrl, xs, ys, i, &out[3 * i]);
  return 0;
}
/* same predicate and arithmetic as new_map_points_kernel, one keypoint at a time */
int emu_new_map_points(const defslam_newpoints_problem *p, uint8_t *action, float *world, int32_t *n_new) {
  const int n = p->n_keypoints, ksz = p->cols / 20, anc = ksz / 2;
  int cnt = 0;
  for (int i = 0; i < n; i++) {
    const int st = p->kp_state[i];
    const int cx = (int)p->kp_xy[2 * i], cy = (int)p->kp_xy[2 * i + 1];
    bool occupied = false;
    if (st == 0)
      for (int j = 0; j < n && !occupied; j++) {
        if (p->kp_state[j] != 1) continue;
        const int mx = (int)p->kp_xy[2 * j], my = (int)p->kp_xy[2 * j + 1];
        occupied = window_hits(cx, mx, p->cols, ksz, anc) && window_hits(cy, my, p->rows, ksz, anc);
      }
    const int act = st == 1 ? 1 : (st == 0 && !occupied ? 2 : 0);
    action[i] = (uint8_t)act;
    if (world) {
      float w[3] = {0.f, 0.f, 0.f};
      if (act != 0) surface_point_to_world(p->T_wc, &p->surf_xyz[3 * i], w);
      world[3 * i] = w[0]; world[3 * i + 1] = w[1]; world[3 * i + 2] = w[2];
    }
    if (act == 2) cnt++;
  }
  *n_new = cnt;
  return 0;
}
/* the kernels of match_cuda.cu in sequence: cells, candidates (count + ranked keys), in-order resolve */
int emu_search_by_projection(const defslam_projsearch_problem *p, int32_t *match_out, int32_t *nmatches_out) {
  ProjView V;
  V.n_last = p->n_last; V.n_cur = p->n_cur; V.n_levels = p->n_levels;
  V.last_state = p->last_state; V.last_has_obs = p->last_has_obs; V.last_desc = p->last_desc; V.cur_desc = p->cur_desc;
  V.cur_taken = p->cur_taken; V.last_xyz = p->last_world_xyz; V.last_angle = p->last_angle; V.cur_xy = p->cur_xy;
  V.cur_angle = p->cur_angle; V.cur_uright = p->cur_uright; V.scale = p->scale_factors; V.last_octave = p->last_octave;
  V.cur_octave = p->cur_octave;
  for (int k = 0; k < 16; k++) V.Tcw[k] = p->T_cw[k];
  V.fx = p->fx; V.fy = p->fy; V.cx = p->cx; V.cy = p->cy; V.mbf = p->mbf;
  V.min_x = p->min_x; V.max_x = p->max_x; V.min_y = p->min_y; V.max_y = p->max_y;
  V.gwi = p->grid_width_inv; V.ghi = p->grid_height_inv; V.th = p->th; V.th_high = p->th_high;
  V.check_orientation = p->check_orientation;
  float twc[3];
  for (int a = 0; a < 3; a++) twc[a] = -(p->T_cw[a] * p->T_cw[3] + p->T_cw[4 + a] * p->T_cw[7] + p->T_cw[8 + a] * p->T_cw[11]);
  const float tlc2 = p->T_lw[8] * twc[0] + p->T_lw[9] * twc[1] + p->T_lw[10] * twc[2] + p->T_lw[11];
  V.forward = tlc2 > p->mb && !p->mono; V.backward = -tlc2 > p->mb && !p->mono;
  std::vector<int> cell(p->n_cur);
  for (int j = 0; j < p->n_cur; j++) { cell[j] = keypoint_cell(V, j); match_out[j] = -1; }
  std::vector<uint8_t> taken(p->cur_taken, p->cur_taken + p->n_cur);
  std::vector<int> acc_i, acc_j;
  int nmatches = 0;
  for (int i = 0; i < p->n_last; i++) {
    const Proj r = project_point(V, i);
    if (!r.ok) continue;
    uint64_t best = ~0ull;
    for (int j = 0; j < p->n_cur; j++) {
      if (!candidate_ok(V, r, j, cell[j]) || taken[j]) continue;
      const uint64_t k = cand_key(hamming256(&p->last_desc[32 * (size_t)i], &p->cur_desc[32 * (size_t)j]), cell[j], j);
      if (k < best) best = k;
    }
    if (best != ~0ull && key_dist(best) < 256 && key_dist(best) <= p->th_high) {
      const int j = key_index(best);
      match_out[j] = i; taken[j] = p->last_has_obs[i];
      acc_i.push_back(i); acc_j.push_back(j);
      nmatches++;
    }
  }
  if (p->check_orientation) {
    int hist[HISTO_LENGTH] = {0}, i1, i2, i3;
    for (size_t a = 0; a < acc_i.size(); a++) hist[rotation_bin(p->last_angle[acc_i[a]], p->cur_angle[acc_j[a]])]++;
    three_maxima(hist, HISTO_LENGTH, i1, i2, i3);
    for (size_t a = 0; a < acc_i.size(); a++) {
      const int bin = rotation_bin(p->last_angle[acc_i[a]], p->cur_angle[acc_j[a]]);
      if (bin != i1 && bin != i2 && bin != i3) { match_out[acc_j[a]] = -1; nmatches--; }
    }
  }
  *nmatches_out = nmatches;
  return 0;
}
int emu_search_by_schwarp(const defslam_warpsearch_problem *p, int32_t *match12_out, int32_t *nmatches_out) {
  const size_t NC = (size_t)p->bbs.nptsu * p->bbs.nptsv;
  std::vector<double> ctrl(2 * NC);
  for (size_t l = 0; l < NC; l++) { ctrl[2 * l] = p->x[l]; ctrl[2 * l + 1] = p->x[NC + l]; }
  WarpView W;
  W.bbs.umin = p->bbs.umin; W.bbs.umax = p->bbs.umax; W.bbs.vmin = p->bbs.vmin; W.bbs.vmax = p->bbs.vmax;
  W.bbs.nptsu = p->bbs.nptsu; W.bbs.nptsv = p->bbs.nptsv; W.bbs.valdim = 2;
  W.ctrl = ctrl.data(); W.n1 = p->n1; W.n2 = p->n2; W.kp1 = p->kp1_norm; W.kp2 = p->kp2_xy; W.st1 = p->kp1_state;
  W.d1 = p->kp1_desc; W.has2 = p->kp2_has_mp; W.d2 = p->kp2_desc; W.fx = p->fx; W.fy = p->fy; W.cx = p->cx; W.cy = p->cy;
  W.min_x = p->min_x; W.max_x = p->max_x; W.min_y = p->min_y; W.max_y = p->max_y; W.gwi = p->grid_width_inv;
  W.ghi = p->grid_height_inv; W.radius = p->radius; W.th_low = p->th_low;
  std::vector<int> cell2(p->n2 > 0 ? p->n2 : 1);
  for (int j = 0; j < p->n2; j++) cell2[j] = cell_of_xy(W.kp2[2 * j], W.kp2[2 * j + 1], W.min_x, W.min_y, W.gwi, W.ghi);
  int n = 0;
  for (int i = 0; i < p->n1; i++) { match12_out[i] = warp_search_one(W, cell2.data(), i); n += match12_out[i] >= 0; }
  *nmatches_out = n;
  return 0;
}
}
