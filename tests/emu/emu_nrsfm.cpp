/*
 * emu_nrsfm.cpp -- TEST INFRASTRUCTURE ONLY (see emu_sft.cpp).  The device code of
 * nrsfm_core.h compiled with g++ as a one-thread team behind the argument lists of the
 * C ABI, so the CPU-only test tier can check the kernels' own arithmetic, indexing and
 * LM control flow against the oracle.  Never loaded by the package.
 */
#define DS_EMULATE 1
#include <vector>

#include "../../include/defslam_b200.h"
#include "../../defslam_b200/csrc/nrsfm_core.h"
#include "../../defslam_b200/csrc/sim3_core.h"

using namespace ds;

static BbsView view2(const defslam_bbs *b) {
  BbsView s;
  s.umin = b->umin; s.umax = b->umax; s.vmin = b->vmin; s.vmax = b->vmax;
  s.nptsu = b->nptsu; s.nptsv = b->nptsv; s.valdim = b->valdim;
  return s;
}

static SchwarpProb to_prob(const defslam_schwarp_problem *p, defslam_diffprop *out, double *scalars) {
  SchwarpProb P;
  P.bbs = view2(&p->bbs);
  P.n = p->n_matches;
  P.kp1 = p->kp1; P.kp2 = p->kp2; P.isig = p->inv_sigma;
  P.lambda = p->lambda; P.fx = p->fx; P.fy = p->fy; P.px_fx = p->px_fx; P.px_fy = p->px_fy;
  P.max_iterations = p->max_iterations; P.initialize = p->initialize;
  P.x = p->x;
  P.warp_uv = out ? out->warp_uv : nullptr; P.J12 = out ? out->J12 : nullptr; P.J21 = out ? out->J21 : nullptr;
  P.H12 = out ? out->H12 : nullptr; P.keep = out ? out->keep : nullptr;
  P.scalars = scalars;
  return P;
}

extern "C" {

int emu_schwarp_fit(const defslam_schwarp_problem *p, defslam_diffprop *out) {
  double scalars[5] = {0, 0, 0, 0, 0};
  SchwarpProb P = to_prob(p, out, scalars);
  const int nu = P.bbs.nptsu, nv = P.bbs.nptsv;
  const SchwarpSizes z = schwarp_ws_sizes(nu, nv, P.n);
  std::vector<uint8_t> wsb(z.total + 256);
  uint8_t *b = wsb.data();
  SchwarpWs ws;
  ws.cell = (int *)(b + z.cell); ws.cstart = (int *)(b + z.cstart); ws.perm = (int *)(b + z.perm);
  ws.taps = (double *)(b + z.taps); ws.CtC = (double *)(b + z.CtC); ws.Hb = (double *)(b + z.Hb);
  ws.Lb = (double *)(b + z.Lb); ws.Js = (double *)(b + z.Js); ws.rdata = (double *)(b + z.rdata);
  ws.sdv = (double *)(b + z.sdv);
  std::vector<double> sh(schwarp_smem(nu, nv).total + 8);
  Team team;
  team.tid = 0; team.nthr = 1;
  schwarp_fit_one(team, P, ws, sh.data());
  if (out) {
    out->cost_initial = scalars[0]; out->cost_final = scalars[1];
    out->iterations = (int)scalars[2]; out->accepted = (int)scalars[3];
  }
  const int st = (int)scalars[4];
  return st == SCHWARP_OK ? 0 : (st == SCHWARP_OUT_OF_DOMAIN ? DEFSLAM_EBADARG : DEFSLAM_ENUMERIC);
}

int emu_schwarp_evaluate(const defslam_schwarp_problem *p, double *r, double *J) {
  double scalars[5];
  SchwarpProb P = to_prob(p, nullptr, scalars);
  const int NC = P.bbs.nptsu * P.bbs.nptsv, NR = 2 * P.n + 4 * NC;
  if (J) for (size_t i = 0; i < (size_t)NR * 2 * NC; i++) J[i] = 0.0;
  for (int row = 0; row < NR; row++) schwarp_row(P, row, r, J);
  return 0;
}

int emu_polysolver_coefficients(int32_t npairs, const float *J12, const float *H12, const float *I1, const float *I2,
                                double *eq1, double *eq2) {
  for (int i = 0; i < npairs; i++)
    pair_polynomials(J12 + 4 * i, H12 + 6 * i, I1 + 2 * i, I2 + 2 * i, 0, eq1 + 10 * i, eq2 + 10 * i, 1);
  return 0;
}

int emu_normals_batched(const defslam_normals_problem *p, double *k_out, double *cov_out, float *normal_out,
                        uint8_t *status_out, int32_t *iters_out, float *pair_normal_out, uint8_t *pair_valid_out) {
  NormalsProb P;
  P.n_points = p->n_points; P.npairs = p->pair_ptr[p->n_points];
  P.pair_ptr = p->pair_ptr; P.J12 = p->J12; P.J21 = p->J21; P.H12 = p->H12; P.I1 = p->I1; P.I2 = p->I2;
  P.k_first = p->k_first; P.ref_uv = p->ref_uv; P.from_ref = p->pair_from_ref; P.k_init = p->k_init;
  P.max_iterations = p->max_iterations; P.corrected_t2 = p->corrected_t2;
  std::vector<double> Q((size_t)20 * (P.npairs + 1));
  P.Q = Q.data();
  P.k_out = k_out; P.cov_out = cov_out; P.normal_out = normal_out; P.pair_normal_out = pair_normal_out;
  P.status_out = status_out; P.pair_valid_out = pair_valid_out; P.iters_out = iters_out;
  for (int i = 0; i < P.n_points; i++) normals_point(P, i);
  return 0;
}

static SfnProb to_sfn(const defslam_sfn_problem *p, int *rc) {
  SfnProb P;
  P.bbs = view2(&p->bbs);
  P.n = p->n_normals; P.n_eval = p->n_eval;
  P.uv = p->uv; P.normals = p->normals; P.eval_uv = p->eval_uv;
  P.bending = p->bending; P.mean_depth = p->mean_depth;
  P.ctrl_out = p->ctrl_out; P.xyz_out = p->xyz_out; P.rc_out = rc;
  return P;
}

int emu_sfn_solve(const defslam_sfn_problem *p) {
  int rc = DEFSLAM_ENUMERIC;
  SfnProb P = to_sfn(p, &rc);
  const int nu = P.bbs.nptsu, nv = P.bbs.nptsv, NC = nu * nv;
  const SfnSizes z = sfn_ws_sizes(nu, nv, P.n);
  std::vector<uint8_t> wsb(z.total + 256);
  uint8_t *b = wsb.data();
  SfnWs ws;
  ws.cell = (int *)(b + z.cell); ws.cstart = (int *)(b + z.cstart); ws.perm = (int *)(b + z.perm);
  ws.taps = (double *)(b + z.taps); ws.mrow = (double *)(b + z.mrow); ws.B = (double *)(b + z.B);
  ws.N = (double *)(b + z.N); ws.res = (double *)(b + z.res); ws.G = (double *)(b + z.G);
  std::vector<double> sh(sfn_smem_fixed(NC) + (size_t)NC * (NC + 1) / 2 + 8);
  Team team;
  team.tid = 0; team.nthr = 1;
  sfn_solve_one(team, P, ws, sh.data(), true);
  return rc;
}

int emu_sfn_system(const defslam_sfn_problem *p, double *A, double *b) {
  int rc = 0;
  SfnProb P = to_sfn(p, &rc);
  const int NC = P.bbs.nptsu * P.bbs.nptsv;
  double ci[48];
  Team team;
  team.tid = 0; team.nthr = 1;
  fill_cell_integrals(team, ci);
  for (int row = 0; row < 2 * P.n + NC + 1; row++) sfn_system_row(P, ci, row, A, b);
  return 0;
}

int emu_sim3_register_batched(int32_t nprob, const defslam_sim3_problem *p, defslam_sim3_result *out, int32_t) {
  Team team;
  team.tid = 0; team.nthr = 1;
  for (int i = 0; i < nprob; i++) {
    double o[16] = {0}, red[200];
    Sim3Prob P;
    P.n = p[i].n_points; P.p1 = p[i].pts1; P.p2 = p[i].pts2;
    for (int k = 0; k < 4; k++) P.init.q[k] = p[i].rot[k];
    for (int k = 0; k < 3; k++) P.init.t[k] = p[i].trans[k];
    P.init.s = p[i].scale; P.chi = p[i].chi; P.huber = p[i].huber; P.max_iterations = p[i].max_iterations;
    P.out = o;
    sim3_register_one(team, P, red);
    for (int k = 0; k < 4; k++) out[i].rot[k] = o[k];
    for (int k = 0; k < 3; k++) out[i].trans[k] = o[4 + k];
    out[i].scale = o[7]; out[i].chi2 = o[8]; out[i].inliers = (int)o[9]; out[i].acceptable = (int)o[10];
    out[i].iterations[0] = (int)o[11]; out[i].iterations[1] = (int)o[12];
  }
  return 0;
}

int emu_scale_min_median(int32_t n, const float *mono, const float *stereo, uint64_t seed, float *scale_out) {
  std::vector<float> buf(n + 1);
  int cnt = 0;
  double sc[4];
  Team team;
  team.tid = 0; team.nthr = 1;
  scale_min_median_team(team, n, mono, stereo, seed, buf.data(), &cnt, sc, scale_out);
  return 0;
}
}
