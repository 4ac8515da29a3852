/*
 * defslam_b200.h -- C ABI of the B200-native DefSLAM deformable hot path.
 *
 * This is the drop-in boundary.  The reference has no
 * FFI of its own: the path sits behind C++ free functions / classes.  Every
 * entry point below names the reference interface it replaces (file:line under
 * the DefSLAM tree).  All arrays are caller-owned, row-major, plain pointers
 * and sizes; no C++ / torch types cross this boundary.
 *
 * Return codes (all entry points):
 *    0  DEFSLAM_OK
 *   -1  DEFSLAM_EBADARG   null pointer / inconsistent sizes / match references
 *                         a node triple that is not coupled in the template
 *   -2  DEFSLAM_ECUDA     CUDA runtime error or no CUDA device (there is NO
 *                         CPU fallback: the call fails)
 *   -3  DEFSLAM_ENUMERIC  non-finite input/outcome; outputs untouched, which
 *                         mirrors the reference's silent "keep the previous
 *                         estimate" behaviour
 *   -4  DEFSLAM_ETOOLARGE problem does not fit the on-chip working set
 *   -5  DEFSLAM_ENOTIMPL  entry point declared but not built yet in this round
 */
#ifndef DEFSLAM_B200_H_
#define DEFSLAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEFSLAM_OK 0
#define DEFSLAM_EBADARG (-1)
#define DEFSLAM_ECUDA (-2)
#define DEFSLAM_ENUMERIC (-3)
#define DEFSLAM_ETOOLARGE (-4)
#define DEFSLAM_ENOTIMPL (-5)

/* ------------------------------------------------------------------------- *
 *  Template (mesh) description
 *  replaces the data the SfT solve reads out of defSLAM::Template / Node /
 *  Edge / Facet / LaplacianMesh:
 *    Modules/Template/Node.cc:114-129,193-204  (1-ring, rest pose, boundary)
 *    Modules/Template/Edge.cc:29-59,72         (edge list, rest length)
 *    Modules/Template/Facet.cc:32-62,77-80     (facet node triples)
 *    Modules/Template/LaplacianMesh.cc:53-162  (mean-value weights, kappa0)
 *    Modules/Template/Template.cc:158-175      (median edge length)
 *  Node order is the caller's (grid order for the regular mesh); the
 *  reference's std::set<Node*> address order is not reproducible.
 * ------------------------------------------------------------------------- */
typedef struct defslam_template_desc {
  int32_t n_nodes;
  int32_t n_edges;
  int32_t n_facets;
  const double *node_rest_xyz;  /* [n_nodes*3]  Node::xO,yO,zO                   */
  const uint8_t *node_boundary; /* [n_nodes]    Node::isBoundary()               */
  const int32_t *nbr_ptr;       /* [n_nodes+1]  CSR of Node::GetNeighbours()     */
  const int32_t *nbr_idx;       /* [nbr_ptr[n]]                                  */
  const double *nbr_w;          /* [nbr_ptr[n]] Node::weights (mean-value w_ij)  */
  const double *node_kappa0;    /* [n_nodes]    |LaplacianCoords| (0 on boundary)*/
  const int32_t *edge_ab;       /* [n_edges*2]                                   */
  const double *edge_len0;      /* [n_edges]    Edge::getDist()                  */
  const int32_t *facets;        /* [n_facets*3] Facet::getNodesArray()           */
  double edge_median_len;       /* Template::getEdgeMeanSize() (the median)      */
} defslam_template_desc;

/* Opaque device-resident plan of a template (topology + Laplacian constants +
 * the normal-matrix sparsity plan).  Lifetime mirrors defSLAM::Template as owned
 * by DefMap (Modules/Common/DefMap.cc:55-82): create when the template is
 * (re)built, destroy when it is replaced. */
typedef struct defslam_template defslam_template;

/* replaces: TemplateGenerator::LaplacianMeshCreate (TemplateGenerator.h:49-50)
 * as far as the solver-visible state goes.  device < 0: current device. */
int defslam_template_create(const defslam_template_desc *desc, int device,
                            defslam_template **out);
void defslam_template_destroy(defslam_template *t);

/* Solver-side facts of a plan (any pointer may be NULL): scalar half-bandwidth of
 * the node part of the normal matrix in the caller's node order, band row
 * length, padded node dimension, number of coupled 3x3 blocks, and the dynamic
 * shared memory one solve needs. */
int defslam_template_info(const defslam_template *t, int32_t *bandwidth, int32_t *band_ld,
                          int32_t *dn_pad, int32_t *n_blocks, int32_t *smem_bytes);

/* ------------------------------------------------------------------------- *
 *  Mesh Laplacian set-up (runs on the GPU)
 *  replaces: LaplacianMesh::ExtractMeanCurvatures  LaplacianMesh.cc:53-148
 *            Edge ctor rest lengths                Edge.cc:52
 *            Template::getEdgeMeanSize             Template.cc:158-175
 *  in : node xyz, facets.   out: everything a defslam_template_desc needs.
 *  nbr_* use a fixed stride of max_ring entries per node (nbr_cnt[n] valid).
 * ------------------------------------------------------------------------- */
int defslam_mesh_laplacian(int32_t n_nodes, const double *node_xyz,
                           int32_t n_facets, const int32_t *facets,
                           int32_t max_ring,
                           int32_t *nbr_cnt,       /* [n_nodes]              */
                           int32_t *nbr_idx,       /* [n_nodes*max_ring]     */
                           double *nbr_w,          /* [n_nodes*max_ring]     */
                           uint8_t *node_boundary, /* [n_nodes]              */
                           double *node_kappa0,    /* [n_nodes]              */
                           int32_t *n_edges_out,
                           int32_t *edge_ab,  /* [cap_edges*2], cap = 3*n_facets */
                           double *edge_len0, /* [cap_edges]                 */
                           double *edge_median_len);

/* ------------------------------------------------------------------------- *
 *  Shape-from-Template solve
 *  replaces: defSLAM::Optimizer::DefPoseOptimization(Frame*,Map*,RegLap,
 *            RegInex,RegTemp,NeighboursLayers)
 *            Modules/Tracking/DefOptimizer.cc:251-578  (DefOptimizer.h:51-53)
 *  and everything below it: sft_types.h edges, g2o LM + BlockSolverX +
 *  LinearSolverDense.
 * ------------------------------------------------------------------------- */
typedef struct defslam_sft_problem {
  /* template: either a plan handle, or (tmpl == NULL) an inline description
   * from which a temporary plan is built for this call */
  const defslam_template *tmpl;
  const defslam_template_desc *tmpl_desc;

  const double *node_xyz;         /* [n_nodes*3] current Node::x,y,z            */
  int32_t n_matches;              /* matched map points that have a facet       */
  int32_t n_frame_keypoints;      /* Frame::N -- the 1/N in Omega (:340)        */
  const int32_t *match_nodes;     /* [n_matches*3] Facet::getNodes() order      */
  const double *match_bary;       /* [n_matches*3] DefMapPoint::b1,b2,b3        */
  const float *match_uv;          /* [n_matches*2] Frame::mvKeysUn[i].pt        */
  const float *match_inv_sigma2;  /* [n_matches]   mvInvLevelSigma2[octave]     */
  double fx, fy, cx, cy;          /* Frame::fx,fy,cx,cy                         */
  float T_cw[16];                 /* Frame::mTcw, row-major 4x4 f32 (cv::Mat)   */
  double reg_lap;                 /* RegLap                                     */
  double reg_inex;                /* RegInex                                    */
  double reg_temp;                /* RegTemp                                    */
  int32_t neighbour_layers;       /* NeighboursLayers (>=1 behaves as 1, C3)    */
  int32_t max_iterations;         /* optimizer.optimize(50) (:513); <=0 -> 50   */
  /* matches_given = 1 selects the second overload,
   *   DefPoseOptimization(const vector<vector<double>> &matches, Frame*, Map*, vector<bool> &outlier,
   *                       RegLap, RegInex, RegTemp)      DefOptimizer.h:58-61, DefOptimizer.cc:582-837
   * whose graph differs: every node is free (:615-620); Omega = I / #matches (:655) and the Huber width
   * is 0.5 (:625); the temporal edges are built but never added (:676-689); curvature edges exist for the
   * non-boundary VIEWED nodes only, one per neighbour, information RegLap / #viewed (:693-758); stretch
   * edges for ALL mesh edges, information RegInex / #edges (:764-797); outlier <=> |e| > 0.5 px (:806-818).
   * match_inv_sigma2 / n_frame_keypoints / neighbour_layers are ignored.  The reference leaves
   * EdgeMeanCurvature::lenghtEdge_ uninitialised there (no setDistanceEdges, quirk C8): the caller says
   * what value it stands for in curv_edge_len (> 0).  The reference does not write the pose back in this
   * overload (only updateNodes, :822); T_cw_out holds the optimised pose all the same. */
  int32_t matches_given;
  double curv_edge_len;
} defslam_sft_problem;

typedef struct defslam_sft_result {
  double *node_xyz_out;  /* [n_nodes*3]   Node::setXYZ values        (:570)    */
  uint8_t *outlier_out;  /* [n_matches]   Frame::mvbOutlier[idx]     (:515-537)*/
  uint8_t *node_role_out;/* [n_nodes] optional: bit0 viewed, bit1 in OptLap    */
  float T_cw_out[16];    /* Frame::SetPose(Converter::toCvMat(SE3)) (:562-566) */
  float rep_error;       /* Frame::repError                          (:559)    */
  int32_t n_inliers;     /* return value nInitialCorrespondences-nBad (:577)   */
  int32_t lm_iterations; /* outer iterations executed (optimize() return)      */
  int32_t lm_trials;     /* total inner trials (linear solves)                 */
  double chi2_initial;   /* activeRobustChi2 before the first iteration        */
  double chi2_final;     /* currentChi after the last accepted step            */
  double lambda_final;
  /* optional LM trace for parity tests: one row per outer iteration
   * {chi2 at start, lambda at start, trials, chi2 at end}; NULL to skip */
  double *trace;         /* [trace_capacity*4]                                 */
  int32_t trace_capacity;
  int32_t status;        /* per-problem return code (batched call)             */
} defslam_sft_result;

int defslam_sft_solve(const defslam_sft_problem *p, defslam_sft_result *r);

/* Batched-frames mode: nprob independent solves in one launch.  device = -1:
 * current device.  Multi-GPU sharding is one process per GPU above this call
 * (SURVEY.md section 8(e)). */
int defslam_sft_solve_batched(int32_t nprob, const defslam_sft_problem *p,
                              defslam_sft_result *r, int device);

/* Resident batch: the frames of a batch are marshalled and uploaded once
 * (create), solved by one kernel launch per run() with inputs already in HBM,
 * and results are copied back on demand (fetch).  This is what a tracking loop
 * that keeps its template and match tables on the device uses; it is also how
 * bench.py separates kernel time from host<->device traffic.  No reference
 * counterpart: the reference solves one frame at a time on the host. */
typedef struct defslam_sft_batch defslam_sft_batch;
int defslam_sft_batch_create(int32_t nprob, const defslam_sft_problem *p, int device,
                             defslam_sft_batch **out);
int defslam_sft_batch_run(defslam_sft_batch *b);   /* launch + wait; sets defslam_last_kernel_ms */
int defslam_sft_batch_fetch(defslam_sft_batch *b, defslam_sft_result *r);
int defslam_sft_batch_info(const defslam_sft_batch *b, int32_t *grid, int32_t *threads,
                           int32_t *smem_bytes, int64_t *h2d_bytes, int64_t *d2h_bytes,
                           double *last_kernel_ms);
void defslam_sft_batch_destroy(defslam_sft_batch *b);

/* Normal equations at the current state (one LM linearisation, no step).
 * replaces: SparseOptimizer::computeActiveErrors + activeRobustChi2
 *           (sparse_optimizer.cpp:104-120) and BlockSolver::buildSystem
 *           (block_solver.hpp:502-560) for the graph DefOptimizer.cc builds.
 * Variable order: node 0 xyz, node 1 xyz, ..., then the 6 camera dofs
 * (omega, upsilon).  Rows/cols of nodes outside OptLap are identity/zero.
 * H_dense: [D*D] full symmetric, b: [D], D = 3*n_nodes + 6. */
int defslam_sft_normal_equations(const defslam_sft_problem *p, double *H_dense,
                                 double *b, double *chi2);

/* Map-point write-back: x = sum_k b_k * node_k, stored fp32.
 * replaces: DefMapPoint::RecalculatePosition  Modules/Common/DefMapPoint.cc:129-147 */
int defslam_mappoints_recalculate(int32_t n_nodes, const double *node_xyz,
                                  int32_t n_points, const int32_t *point_nodes,
                                  const double *point_bary, float *point_xyz_out);

/* ------------------------------------------------------------------------- *
 *  Template construction helpers
 * ------------------------------------------------------------------------- */
/* Barycentric embedding of map points (fp32 arithmetic as in the reference).
 * replaces: TriangularMesh::calculateFeaturesCoordinates / pointInTriangle
 *           Modules/Template/TriangularMesh.cc:133-236
 * out_facet[i] = facet index or -1; out_bary in the facet's ascending-node order */
int defslam_embed_points(int32_t n_nodes, const double *node_xyz,
                         int32_t n_facets, const int32_t *facets,
                         int32_t n_points, const float *point_xyz,
                         int32_t *out_facet, int32_t *out_nodes /*[n*3]*/,
                         float *out_bary /*[n*3]*/);

/* ------------------------------------------------------------------------- *
 *  Bicubic B-spline (BBS) evaluation
 *  replaces: BBS::EvalEigen   Thirdparty/BBS/bbs_coloc.cc:610-653
 *            BBS::eval        Thirdparty/BBS/bbs.cc:155-195
 * ------------------------------------------------------------------------- */
typedef struct defslam_bbs {
  double umin, umax;
  int32_t nptsu;
  double vmin, vmax;
  int32_t nptsv;
  int32_t valdim;
} defslam_bbs;

/* ctrl: [valdim * nptsu * nptsv], index valdim*((iu)*nptsv + iv) + d
 * val : [nsites * valdim]  (site-major, like bbs.cc eval)                    */
int defslam_bbs_eval(const defslam_bbs *bbs, const double *ctrl, int32_t nsites,
                     const double *u, const double *v, int32_t du, int32_t dv,
                     double *val);

/* All six derivative orders (0,0),(1,0),(0,1),(2,0),(1,1),(0,2) in one pass.
 * replaces: the 6x Warp::getEstimates calls  SchwarpDatabase.cc:243-264
 * val6: [6 * nsites * valdim] */
int defslam_bbs_eval6(const defslam_bbs *bbs, const double *ctrl, int32_t nsites,
                      const double *u, const double *v, double *val6);

/* Dense collocation matrix rows, 16 non-zeros per site.
 * replaces: BBS::colocEigen / coloc_derivEigen  bbs_coloc.cc:76-207
 * C: [nsites * nptsu*nptsv] row-major dense */
int defslam_bbs_coloc(const defslam_bbs *bbs, int32_t nsites, const double *u,
                      const double *v, int32_t du, int32_t dv, double *C);

/* Dense bending-energy matrix (lambda = 1).
 * replaces: BBS::BendingEigen  bbs_coloc.cc:406-507 / bending_ur bbs.cc:563-640
 * B: [NC*NC], NC = nptsu*nptsv */
int defslam_bbs_bending(const defslam_bbs *bbs, double *B);

/* ------------------------------------------------------------------------- *
 *  NRSfM stages (mapping thread).  All three are batched: the units (keyframe
 *  pairs, map points, keyframes) are independent.
 * ------------------------------------------------------------------------- */
/* Schwarp fit between two keyframes.
 * replaces: SchwarpDatabase::calculateSchwarps  Modules/Mapping/SchwarpDatabase.cc:145-349
 *           Warps::Warp / Warps::Schwarzian     Modules/Mapping/Schwarp.cc
 *           Warps::Warp::initialize             Modules/Mapping/Schwarp.cc:99-160
 * The trust-region iteration is Ceres' Levenberg-Marquardt with its default
 * options (Jacobi scaling, initial radius 1e4, min_relative_decrease 1e-3,
 * tolerances 1e-6/1e-8/1e-10) and the HuberLoss(5.77) corrector on the data
 * block; Ceres itself is an unpinned external dependency of the reference.
 * Reference quirk C6 (Schwarp.cc:291-299) is reproduced: the Jacobian rows of
 * the y residuals are the rows of the x residuals, without the 1/sigma factor. */
typedef struct defslam_schwarp_problem {
  defslam_bbs bbs;            /* domain of KF1 (DefKeyFrame umin..vmax), valdim=2 */
  int32_t n_matches;
  const float *kp1;           /* [n*2] normalised keypoints in KF1               */
  const float *kp2;           /* [n*2] normalised keypoints in KF2               */
  const float *inv_sigma;     /* [n]   sqrt(invLevelSigma2) per match            */
  double lambda;              /* LocalMapping.Schwarp.Regularizer                */
  double fx, fy;              /* as handed to Warps::Warp (the reference passes
                                 (KF->fy, KF->fx), SchwarpDatabase.cc:199-201)   */
  double px_fx, px_fy;        /* KF->fx, KF->fy of the 10 px unlink test
                                 SchwarpDatabase.cc:283-293                      */
  int32_t max_iterations;     /* 3 in the reference                              */
  int32_t initialize;         /* 1: x0 = (C'C + lambda*B)^-1 C' q2 (Warp::initialize) */
  double *x;                  /* in/out [2*NC] control points, [all x; all y]    */
} defslam_schwarp_problem;

typedef struct defslam_diffprop {
  /* per match, fp32 like Modules/Mapping/diffProp.h:52-88 */
  float *warp_uv;   /* [n*2] warped position of kp1                             */
  float *J12;       /* [n*4] a,b,c,d = du/du, dv/du, du/dv, dv/dv               */
  float *J21;       /* [n*4] a,b,c,d of the inverse                             */
  float *H12;       /* [n*6] uux,uuy,uvx,uvy,vvx,vvy                            */
  uint8_t *keep;    /* [n]   0 if the warp error exceeds 10 px (unlinked)       */
  double cost_initial, cost_final; /* Ceres cost = 0.5*sum rho                  */
  int32_t iterations;              /* trust-region steps taken (<= max)         */
  int32_t accepted;                /* how many of them were accepted            */
} defslam_diffprop;

int defslam_schwarp_fit(const defslam_schwarp_problem *p, defslam_diffprop *out);
/* nprob independent keyframe pairs, one CTA each. device: CUDA ordinal, -1 = current */
int defslam_schwarp_fit_batched(int32_t nprob, const defslam_schwarp_problem *p,
                                defslam_diffprop *out, int32_t device);

/* Initial warp between two keyframes and the match filter that goes with it.
 * replaces: DefORBmatcher::CalculateInitialSchwarp  Modules/Matching/DefORBmatcher.cc:111-187
 *           (the first half of DefORBmatcher::findbyWarp :47-71; the second half is defslam_search_by_schwarp)
 *   x <- Warps::Warp::initialize (:146-147), NaN entries among the first 2*NCu*NCu of them set to 0 (:149-154;
 *   the reference scrubs NCu*NCu*2 = 338 of its 390 entries, quirk C9);
 *   residuals of the Warp cost at x as ceres::Problem::Evaluate returns them with HuberLoss(5.77) on the block
 *   (:156-169; default EvaluateOptions apply the loss: every residual is scaled by sqrt(rho'), rho' = 1 if the
 *   block's squared norm s <= 5.77^2 else 5.77/sqrt(s));
 *   match i is dropped when residuals[2i]^2 + residuals[2i+1]^2 > 20 (:171-186).  Kept bug-compatible: the
 *   residual vector is laid out [all x residuals; all y residuals] (Schwarp.cc:275-280), so entries 2i, 2i+1 are
 *   the x (i < n/2) or y residuals of matches 2i, 2i+1 (mod n), not the two residuals of match i.
 * p->fx, p->fy: as handed to Warps::Warp here, (KF->fx, KF->fy) (:159-160).  p->initialize / max_iterations are
 * ignored.  p->x: out [2*NC].  keep_out: [n] 1 = match kept.  err_out: [n] the tested quantity, or NULL. */
int defslam_schwarp_initial(const defslam_schwarp_problem *p, uint8_t *keep_out, double *err_out);

/* Residuals / Jacobian of the Schwarp cost at p->x, before the loss corrector (parity hook).
 * replaces: Warp::Evaluate Schwarp.cc:235-303 + Schwarzian::Evaluate :368-543
 * r: [2n + 4NC]; J: [(2n+4NC) * 2NC] row-major dense or NULL                 */
int defslam_schwarp_evaluate(const defslam_schwarp_problem *p, double *r, double *J);

/* Batched isometric-NRSfM normal estimation: one 2-unknown LM per map point,
 * then the transfer of the normal to the second keyframe of every pair.
 * replaces: NormalEstimator::ObtainK1K2  Modules/Mapping/NormalEstimator.cc:38-229
 *           PolySolver::getCoefficients/Evaluate Modules/Mapping/PolySolver.cc:50-193
 * Pair data (the DiffProp records of WarpDatabase) is CSR by point. */
typedef struct defslam_normals_problem {
  int32_t n_points;
  const int32_t *pair_ptr; /* [n_points+1]                                      */
  const float *J12;        /* [npairs*4] a,b,c,d                                */
  const float *J21;        /* [npairs*4] a,b,c,d of the inverse (transfer)      */
  const float *H12;        /* [npairs*6] uux,uuy,uvx,uvy,vvx,vvy                */
  const float *I1;         /* [npairs*2] point in the pair's first KF (u,v)     */
  const float *I2;         /* [npairs*2] point in the pair's second KF (u,v)    */
  const uint8_t *pair_from_ref; /* [npairs] 1: first KF == the point's reference
                                   KF (the pair enters the polynomial system)   */
  const float *k_first;    /* [npairs*2] (k1,k2) stored for the point in the
                              pair's first KF, read when pair_from_ref==0;
                              NaN = no normal there (pair skipped); may be NULL */
  const double *k_init;    /* [n_points*2] initial (k1,k2): last estimate or 0  */
  const float *ref_uv;     /* [n_points*2] normalised keypoint in the ref KF    */
  int32_t max_iterations;  /* 200                                               */
  int32_t corrected_t2;    /* 0 = reference-compatible (default): the polynomial
                              build takes t2 = (c*H12vvy - d*H12vvx)/2
                              (NormalEstimator.cc:96-97), which vanishes for any
                              projective warp; 1 = the definition of the transfer
                              step t2 = (d*H12uux - c*H12uuy)/2 (:210), with which
                              the polynomials vanish on an isometric pair (quirk C7) */
} defslam_normals_problem;

/* status_out[i]: 0 no equation for the point (nothing estimated), 1 estimated,
 *                2 covariance rank deficient (reference `continue`s: no normal, no transfer)
 * pair_valid_out[j]: 1 if pair_normal_out[j*3..] was written                  */
int defslam_normals_batched(const defslam_normals_problem *p,
                            double *k_out /*[n*2]*/, double *cov_out /*[n*4]*/,
                            float *normal_out /*[n*3]*/, uint8_t *status_out /*[n]*/,
                            int32_t *iters_out /*[n]*/,
                            float *pair_normal_out /*[npairs*3]*/,
                            uint8_t *pair_valid_out /*[npairs]*/);

/* PolySolver::getCoefficients for one pair (parity hook)  PolySolver.cc:50-149 */
int defslam_polysolver_coefficients(int32_t npairs, const float *J12, const float *H12,
                                    const float *I1, const float *I2,
                                    double *eq1 /*[npairs*10]*/, double *eq2 /*[npairs*10]*/);

/* Shape-from-normals depth-spline solve.
 * replaces: ShapeFromNormals::{ShapeFromNormals,obtainM,estimate}
 *           Modules/Mapping/ShapeFromNormals.cc:38-76,178-260,81-171
 * Least squares [M; bending*B; 1'] X = [0; 0; NC*mean_depth]; the reference
 * uses a dense Householder QR, here: normal equations + Cholesky with two
 * corrected-semi-normal-equation refinement sweeps (same minimiser). */
typedef struct defslam_sfn_problem {
  defslam_bbs bbs;          /* valdim = 1                                       */
  int32_t n_normals;
  const float *uv;          /* [n*2] normalised keypoint                        */
  const float *normals;     /* [n*3] (not normalised)                           */
  double bending;           /* LocalMapping.Bending                             */
  double mean_depth;        /* DefKeyFrame::accMean (=1)                        */
  int32_t n_eval;           /* sites where depth is evaluated afterwards        */
  const float *eval_uv;     /* [n_eval*2]                                       */
  double *ctrl_out;         /* [NC] control depths after the median rescale     */
  float *xyz_out;           /* [n_eval*3] (u d, v d, d)                         */
} defslam_sfn_problem;

/* DEFSLAM_ENUMERIC when the solution is not finite (reference: estimate() == false) */
int defslam_sfn_solve(const defslam_sfn_problem *p);
int defslam_sfn_solve_batched(int32_t nprob, const defslam_sfn_problem *p, int32_t *rc_out /*[nprob] or NULL*/,
                              int32_t device);
/* The stacked system itself (parity hook): A [(2n+NC+1) * NC] row-major, b [2n+NC+1] */
int defslam_sfn_system(const defslam_sfn_problem *p, double *A, double *b);

/* Sim(3) surface registration (last stage of DefLocalMapping::NRSfM).
 * replaces: Optimizer::OptimizeHorn          Modules/Tracking/DefOptimizer.cc:840-922
 *           EdgeSim3Simple / VertexSim3ExpmapNoProj
 *                                            Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h:96-188
 *           Sim3(update) / map / operator*   Thirdparty/g2o/g2o/types/sim3.h:70-160,270-277
 *           called from SurfaceRegistration::registerSurfaces
 *                                            Modules/Mapping/SurfaceRegistration.cc:48-153
 * Two runs of g2o's Levenberg-Marquardt (<= max_iterations each) over the 7-dof vertex with one
 * Huber edge e = p2 - S.map(p1) per point.  The reference differentiates numerically (central
 * differences, delta 1e-9, base_unary_edge.hpp:82-118); the kernel uses the analytic Jacobian those
 * differences approximate. */
typedef struct defslam_sim3_problem {
  int32_t n_points;
  const float *pts1;      /* [n*3] cloud moved by the transform (surface points, world frame) */
  const float *pts2;      /* [n*3] target cloud (stored map-point positions)                  */
  double rot[4];          /* initial Sim3: unit quaternion x,y,z,w                            */
  double trans[3];
  double scale;           /* GroundTruthTools::scaleMinMedian                                 */
  double chi;             /* chiLimit^2                                                       */
  double huber;           /* 0.01; the kernel width is sqrt(huber)                            */
  int32_t max_iterations; /* 50                                                               */
} defslam_sim3_problem;

typedef struct defslam_sim3_result {
  double rot[4], trans[3], scale; /* estimate after the FIRST run: what the reference copies
                                     back into g2oS12 (DefOptimizer.cc:896)                   */
  double chi2;                    /* optimizer.chi2() after the second run                    */
  int32_t inliers;                /* edges with chi2 <= chi after the first run               */
  int32_t acceptable;             /* chi2 finite and chi2 / inliers < chi                     */
  int32_t iterations[2];
} defslam_sim3_result;

int defslam_sim3_register_batched(int32_t nprob, const defslam_sim3_problem *p, defslam_sim3_result *out,
                                  int32_t device);

/* Min-median scale between two clouds.
 * replaces: GroundTruthTools::scaleMinMedian  Modules/GroundTruth/GroundTruthCalculator.cc:54-159
 * The reference subsamples candidates and residuals with unseeded rand() (quirk C10); here the
 * same 25 % Bernoulli draws come from a counter-based hash of (seed, i, j), so the result is
 * reproducible.  Returns the scale in *scale_out (0 when a candidate has no sampled residual,
 * like the reference). */
int defslam_scale_min_median(int32_t n, const float *mono_xyz, const float *stereo_xyz, uint64_t seed,
                             float *scale_out);

/* New map points from the estimated surface, and the exploration test.
 * replaces: DefLocalMapping::CreateNewMapPoints  Modules/Mapping/DefLocalMapping.cc:240-347
 *           DefLocalMapping::needNewTemplate     Modules/Mapping/DefLocalMapping.cc:355-403
 * The reference paints a rows x cols 8-bit mask with 255 at (int)pt.y,(int)pt.x of every keypoint
 * that owns a good map point, box-filters it with cv::filter2D (ones kernel of edge cols/20, anchor at
 * its centre, BORDER_REFLECT_101, saturating) and thresholds at 1: a pixel is "occupied" when a marked
 * pixel lies in its window.  Here the same predicate is evaluated per keypoint against the list of
 * marked pixels (no image).  Keypoints outside the image are EBADARG (the reference indexes the mask
 * out of bounds there).
 *   kp_state[i]: 0 = no map point, 1 = good map point, 2 = bad map point (isBad())
 *   action_out[i]: 0 = leave, 1 = move the existing map point to world_xyz_out[i],
 *                  2 = create a map point at world_xyz_out[i]
 *   world_xyz_out[i] = (Twc * [surf_xyz[i]; 1])(0..2), fp32 products summed in fp32 like cv::gemm
 *   *n_new_out = number of state-0 keypoints on unoccupied pixels (needNewTemplate's newPoints)
 * surf_xyz / T_wc / world_xyz_out may be NULL when only the count is wanted. */
typedef struct defslam_newpoints_problem {
  int32_t n_keypoints, rows, cols;
  const float *kp_xy;       /* [n*2] KeyFrame::mvKeysUn[i].pt (x, y)                          */
  const uint8_t *kp_state;  /* [n]                                                            */
  const float *surf_xyz;    /* [n*3] Surface::get3DSurfacePoint (camera frame of the keyframe) */
  const float *T_wc;        /* [16] row-major KeyFrame::GetPoseInverse()                      */
} defslam_newpoints_problem;

int defslam_new_map_points(const defslam_newpoints_problem *p, uint8_t *action_out, float *world_xyz_out,
                           int32_t *n_new_out);

/* Match production feeding the SfT solve: projection search of the last frame's template points.
 * replaces: DefORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)
 *             Modules/Matching/DefORBmatcher.cc:296-451
 *           Frame::GetFeaturesInArea / PosInGrid / AssignFeaturesToGrid
 *             Thirdparty/ORBSLAM_2/src/Frame.cc:421-496,294-309
 *           ORBmatcher::DescriptorDistance / ComputeThreeMaxima
 *             Thirdparty/ORBSLAM_2/src/ORBmatcher.cc:1645-1707
 * The reference loops over the last frame's keypoints in order and lets earlier assignments hide
 * keypoints from later map points; here candidates and Hamming distances are computed in parallel
 * (one thread per map point, every keypoint of the current frame tested against the reference's
 * cell window) and one warp then replays the assignments in the reference's order, ties broken by
 * the order GetFeaturesInArea would have listed the candidates (cell column, cell row, index).
 *   last_state[i]: 1 = map point usable (non-null, not an outlier, not bad, has a facet), else 0
 *   last_has_obs[i]: MapPoint::Observations() > 0 (a keypoint assigned to such a point is hidden from
 *                    the points after it)
 *   cur_taken[j]: CurrentFrame.mvpMapPoints[j] already holds a point with observations
 *   match_out[j]: index i of the last-frame keypoint whose map point keypoint j received, or -1
 *   *nmatches_out: the reference's return value */
typedef struct defslam_projsearch_problem {
  int32_t n_last, n_cur, n_levels;
  const uint8_t *last_state;        /* [n_last]      */
  const uint8_t *last_has_obs;      /* [n_last]      */
  const float *last_world_xyz;      /* [n_last*3]  MapPoint::GetWorldPos()              */
  const uint8_t *last_desc;         /* [n_last*32] MapPoint::GetDescriptor()            */
  const int32_t *last_octave;       /* [n_last]    LastFrame.mvKeys[i].octave           */
  const float *last_angle;          /* [n_last]    LastFrame.mvKeysUn[i].angle          */
  const float *cur_xy;              /* [n_cur*2]   CurrentFrame.mvKeysUn[j].pt          */
  const int32_t *cur_octave;        /* [n_cur]                                          */
  const float *cur_angle;           /* [n_cur]                                          */
  const uint8_t *cur_desc;          /* [n_cur*32]  CurrentFrame.mDescriptors            */
  const float *cur_uright;          /* [n_cur]     mvuRight (negative: monocular)       */
  const uint8_t *cur_taken;         /* [n_cur]                                          */
  const float *scale_factors;       /* [n_levels]  mvScaleFactors                       */
  float T_cw[16];                   /* CurrentFrame.mTcw, row-major                     */
  float T_lw[16];                   /* LastFrame.mTcw                                   */
  float fx, fy, cx, cy, mb, mbf;
  float min_x, max_x, min_y, max_y; /* mnMinX ... mnMaxY                                */
  float grid_width_inv, grid_height_inv; /* mfGridElementWidthInv / HeightInv (64 x 48 cells) */
  float th;
  int32_t mono;                     /* bMono                                            */
  int32_t th_high;                  /* ORBmatcher::TH_HIGH (75)                         */
  int32_t check_orientation;        /* mbCheckOrientation                               */
} defslam_projsearch_problem;

int defslam_search_by_projection(const defslam_projsearch_problem *p, int32_t *match_out, int32_t *nmatches_out);

/* Warp-guided search of map points between two keyframes.
 * replaces: DefORBmatcher::searchBySchwarp  Modules/Matching/DefORBmatcher.cc:190-293
 *           (Warps::Warp::getEstimates  Modules/Mapping/Schwarp.cc:162-233, KeyFrame::GetFeaturesInArea / IsInImage
 *            Thirdparty/ORBSLAM_2/src/KeyFrame.cc:618-668)
 * Every keypoint of keyframe 1 that owns a usable map point not yet seen in keyframe 2 is sent through the
 * bicubic B-spline warp x (the reference's layout: NC u-coordinates, then NC v-coordinates), converted to pixels
 * of keyframe 2, and matched to the closest descriptor (< th_low) among the keypoints of keyframe 2 without a map
 * point within `radius` pixels; first minimum in the grid order of GetFeaturesInArea.  No order dependence.
 *   kp1_state[i]: 1 = candidate (map point non-null, not bad, not in keyframe 2), else 0
 *   match12_out[i]: index of the keypoint of keyframe 2, or -1;  *nmatches_out = number of candidates matched */
typedef struct defslam_warpsearch_problem {
  defslam_bbs bbs;                  /* domain of keyframe 1, NCu x NCv, valdim 2          */
  const double *x;                  /* [2*NC] warp control points                          */
  int32_t n1, n2;
  const float *kp1_norm;            /* [n1*2] DefKeyFrame::mpKeypointNorm[i].pt            */
  const uint8_t *kp1_state;         /* [n1]                                                */
  const uint8_t *kp1_desc;          /* [n1*32] mDescriptors of keyframe 1                  */
  const float *kp2_xy;              /* [n2*2] mvKeysUn of keyframe 2 (pixels)              */
  const uint8_t *kp2_has_mp;        /* [n2]   GetMapPoint(j) != NULL                       */
  const uint8_t *kp2_desc;          /* [n2*32]                                             */
  float fx, fy, cx, cy;             /* keyframe 2                                          */
  float min_x, max_x, min_y, max_y, grid_width_inv, grid_height_inv;
  float radius;                     /* th = 2                                              */
  int32_t th_low;                   /* ORBmatcher::TH_LOW (50)                             */
} defslam_warpsearch_problem;

int defslam_search_by_schwarp(const defslam_warpsearch_problem *p, int32_t *match12_out, int32_t *nmatches_out);

/* Surface -> template nodes.
 * replaces: Surface::getVertex  Modules/Mapping/Surface.cc:125-161
 * nodes_out: [xs*ys*3] fp32 (u d, v d, d), x-major outer loop */
int defslam_surface_vertices(const defslam_bbs *bbs, const double *ctrl_depth,
                             int32_t xs, int32_t ys, float *nodes_out);

/* ------------------------------------------------------------------------- *
 *  Library information
 * ------------------------------------------------------------------------- */
const char *defslam_version(void);
/* number of kernel launches issued by this process through the library */
int64_t defslam_kernel_launch_count(void);
int defslam_device_count(void);
/* device time (ms) of the kernels of the last batched solve on this thread,
 * measured with CUDA events on the library's stream */
double defslam_last_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* DEFSLAM_B200_H_ */
