/* Oracle entry points.  TEST INFRASTRUCTURE ONLY -- see sft_oracle.c header. */
#ifndef DEFSLAM_ORACLE_H_
#define DEFSLAM_ORACLE_H_
#include "../include/defslam_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- SfT (sft_oracle.c) ---- */
int oracle_sft_solve(const defslam_sft_problem *p, defslam_sft_result *r);
int oracle_sft_normal_equations(const defslam_sft_problem *p, double *H_dense, double *b, double *chi2);
int oracle_sft_residuals(const defslam_sft_problem *p, double *res, double *J, int max_rows);
int oracle_sft_residuals_pose(const defslam_sft_problem *p, const double *q, const double *t,
                              const double *node_xyz, double *res, int max_rows);
int oracle_sft_apply_update(const defslam_sft_problem *p, const double *d, double *node_xyz_out,
                            float *T_cw_out, double *q_out, double *t_out);
int oracle_mappoints_recalculate(int32_t n_nodes, const double *node_xyz, int32_t n_points,
                                 const int32_t *point_nodes, const double *point_bary, float *out);

/* ---- template (template_oracle.c) ---- */
int oracle_regular_triangulation(int nodes_ver, int nodes_hor, int32_t *facets /*[2*(v-1)*(h-1)*3]*/);
int oracle_mesh_laplacian(int32_t n_nodes, const double *node_xyz, int32_t n_facets, const int32_t *facets,
                          int32_t max_ring, int32_t *nbr_cnt, int32_t *nbr_idx, double *nbr_w,
                          uint8_t *node_boundary, double *node_kappa0, int32_t *n_edges_out,
                          int32_t *edge_ab, double *edge_len0, double *edge_median_len);
int oracle_embed_points(int32_t n_nodes, const double *node_xyz, int32_t n_facets, const int32_t *facets,
                        int32_t n_points, const float *point_xyz, int32_t *out_facet, int32_t *out_nodes,
                        float *out_bary);

/* ---- bicubic B-splines (bbs_oracle.c) ---- */
int oracle_bbs_eval(const defslam_bbs *s, const double *ctrl, int32_t nsites, const double *u, const double *v,
                    int32_t du, int32_t dv, double *val);
int oracle_bbs_coloc(const defslam_bbs *s, int32_t nsites, const double *u, const double *v, int32_t du, int32_t dv,
                     double *C);
int oracle_bbs_bending(const defslam_bbs *s, double *B);
int oracle_surface_vertices(const defslam_bbs *s, const double *ctrl, int32_t xs, int32_t ys, float *out);

/* ---- NRSfM stages (nrsfm_oracle.c) ---- */
int oracle_schwarp_evaluate(const defslam_schwarp_problem *p, double *r, double *J);
int oracle_schwarp_init(const defslam_schwarp_problem *p, double *x0);
int oracle_schwarp_fit(const defslam_schwarp_problem *p, defslam_diffprop *out);
int oracle_schwarp_initial(const defslam_schwarp_problem *p, uint8_t *keep_out, double *err_out);
int oracle_polysolver_coefficients(int32_t npairs, const float *J12, const float *H12, const float *I1,
                                   const float *I2, double *eq1, double *eq2);
int oracle_normals_batched(const defslam_normals_problem *p, double *k_out, double *cov_out, float *normal_out,
                           uint8_t *status_out, int32_t *iters_out, float *pair_normal_out,
                           uint8_t *pair_valid_out);
int oracle_sfn_system(const defslam_sfn_problem *p, double *A, double *b);
int oracle_sfn_solve(const defslam_sfn_problem *p);

/* ---- Sim(3) surface registration (sim3_oracle.c) ---- */
int oracle_sim3_register_batched(int32_t nprob, const defslam_sim3_problem *p, defslam_sim3_result *out, int32_t dev);
int oracle_sim3_jacobian(const defslam_sim3_problem *p, int i, double *J21);
int oracle_scale_min_median(int32_t n, const float *mono, const float *stereo, uint64_t seed, float *scale_out);

/* ---- new map points / exploration test (newpts_oracle.c) ---- */
int oracle_new_map_points(const defslam_newpoints_problem *p, uint8_t *action_out, float *world_xyz_out,
                          int32_t *n_new_out);

/* ---- projection search (match_oracle.c) ---- */
int oracle_search_by_projection(const defslam_projsearch_problem *p, int32_t *match_out, int32_t *nmatches_out);
int oracle_search_by_schwarp(const defslam_warpsearch_problem *p, int32_t *match12_out, int32_t *nmatches_out);

#ifdef __cplusplus
}
#endif
#endif
