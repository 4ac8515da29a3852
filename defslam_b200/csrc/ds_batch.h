/*
 * ds_batch.h -- host-side marshalling of a batch of SfT problems.
 *
 * Turns the caller's defslam_sft_problem array (what a DefOptimizer.cc-shaped
 * adapter fills from Frame / Map, DefOptimizer.cc:293-361) into one contiguous
 * input arena + one output arena, so that a whole batch moves with one
 * host->device and one device->host copy.  Pure host C++ (no CUDA calls): the
 * CUDA library and the test-only emulation build share it.
 */
#ifndef DS_BATCH_H_
#define DS_BATCH_H_

#include <vector>

#include "ds_host.h"
#include "sft_core.h"

namespace ds {

struct ProbSlot {
  size_t in_off, out_off;
  size_t o_nodes, o_bary, o_mnodes, o_uv, o_isig;          /* within the input slot  */
  size_t o_out_nodes, o_trace, o_res, o_outlier, o_role, o_H, o_b; /* within the output slot */
  int n_nodes, M, trace_cap;
};

struct BatchMarshal {
  std::vector<ProbSlot> slots;
  std::vector<ProbView> views;
  size_t in_bytes = 0, out_bytes = 0;
  /* workspace / launch sizing (max over the batch) */
  size_t ws_band = 0, ws_dinv = 0, ws_cg = 0, ws_F = 0, ws_S = 0, ws_M = 0, ws_nf = 0, ws_dp = 0;
  int smem_doubles = 0;
  bool any_e_global = false;
  int row_mode = 1;          /* 0: never use the row-owner factorisation (DEFSLAM_ROW_MODE=0, A/B runs) */
  bool any_x_global = false; /* then EVERY problem of the batch keeps x/dx in the workspace (one kernel variant per launch) */

  WorkspaceSizes ws_sizes() const {
    WorkspaceSizes z;
    z.band = ws_band; z.dinv = ws_dinv; z.cg = ws_cg; z.F = ws_F; z.S = ws_S; z.M = ws_M; z.nf = ws_nf; z.dp = ws_dp;
    return z;
  }

  static size_t al(size_t x) { return (x + 15) & ~(size_t)15; }

  /* smem_limit_doubles: largest dynamic shared memory a CTA may use */
  template <class Resolve>
  int plan(int nprob, const defslam_sft_problem *p, int mode, int smem_limit_doubles, Resolve resolve) {
    slots.assign(nprob, ProbSlot());
    views.assign(nprob, ProbView());
    in_bytes = out_bytes = 0;
    ws_band = ws_dinv = ws_cg = ws_F = ws_S = ws_M = ws_nf = ws_dp = 0;
    smem_doubles = 0;
    any_e_global = false;
    any_x_global = false;
    std::vector<const PlanView *> hvs(nprob);
    for (int i = 0; i < nprob; i++) {
      const PlanView *hv = nullptr, *dv = nullptr;
      const int rc = resolve(p[i], &hv, &dv);
      if (rc != 0) return rc;
      hvs[i] = hv;
      if (!p[i].node_xyz || p[i].n_matches < 0 || (p[i].n_frame_keypoints <= 0 && !p[i].matches_given)) return DEFSLAM_EBADARG;
      if (p[i].n_matches > 0 && (!p[i].match_nodes || !p[i].match_bary || !p[i].match_uv ||
                                 (!p[i].match_inv_sigma2 && !p[i].matches_given)))
        return DEFSLAM_EBADARG;
      ProbSlot &s = slots[i];
      const size_t n = hv->n_nodes, M = p[i].n_matches;
      s.n_nodes = (int)n; s.M = (int)M;
      size_t o = 0;
      s.o_nodes = o;  o += al(3 * n * sizeof(double));
      s.o_bary = o;   o += al(3 * M * sizeof(double));
      s.o_mnodes = o; o += al(3 * M * sizeof(int));
      s.o_uv = o;     o += al(2 * M * sizeof(float));
      s.o_isig = o;   o += al(M * sizeof(float));
      s.in_off = in_bytes; in_bytes += o;
      s.trace_cap = 64;
      o = 0;
      s.o_out_nodes = o; o += al(3 * n * sizeof(double));
      s.o_trace = o;     o += al(4 * (size_t)s.trace_cap * sizeof(double));
      s.o_res = o;       o += al(sizeof(ResultScalars));
      s.o_outlier = o;   o += al(M);
      s.o_role = o;      o += al(n);
      const size_t D = 3 * n + 6;
      s.o_H = o; if (mode == MODE_NORMAL_EQ) o += al(D * D * sizeof(double));
      s.o_b = o; if (mode == MODE_NORMAL_EQ) o += al(D * sizeof(double));
      s.out_off = out_bytes; out_bytes += o;

      ProbView &v = views[i];
      v.plan = dv;
      v.mode = mode;
      v.n_matches = (int)M;
      v.n_kp = p[i].n_frame_keypoints;
      v.max_it = p[i].max_iterations;
      v.layers = p[i].neighbour_layers;
      v.variant = p[i].matches_given ? 1 : 0;
      v.curv_len = p[i].curv_edge_len;
      if (v.variant == 1 && !(p[i].curv_edge_len > 0.0)) return DEFSLAM_EBADARG;
      v.trace_cap = s.trace_cap;
      v.fx = p[i].fx; v.fy = p[i].fy; v.cx = p[i].cx; v.cy = p[i].cy;
      v.reg_lap = p[i].reg_lap; v.reg_inex = p[i].reg_inex; v.reg_temp = p[i].reg_temp;
      for (int k = 0; k < 16; k++) v.Tcw[k] = p[i].T_cw[k];

      /* shared-memory placement, most resident first: border rows and x/dx in smem; border rows in
       * the global workspace; x/dx there too (large meshes: the window alone fills the SM).  Meshes whose
       * band has a supported tile count and whose ring fits take the row-owner factorisation (sft_rows.h). */
      SmemLayout L = smem_layout(hv->n_nodes, hv->n_edges, hv->Dn_pad, hv->bwp, hv->ld, hv->Wr, hv->ES, true, true);
      v.e_in_smem = 1;
      v.x_in_smem = 1;
      v.row_nt = 0;
      {
        const int nt = hv->bwp / NB + 1;
        if (row_mode != 0 && mode == MODE_SOLVE && row_mode_nt_supported(nt)) {
          const SmemLayout Lr = smem_layout(hv->n_nodes, hv->n_edges, hv->Dn_pad, hv->bwp, hv->ld, hv->Wr, hv->ES, false, true, nt);
          if (Lr.total + CTX_DOUBLES <= smem_limit_doubles) {
            L = Lr;
            v.row_nt = nt;
            v.e_in_smem = 0;
            ws_band = std::max(ws_band, row_mode_factor_doubles(hv->nblk, nt));
          }
        }
      }
      if (L.total + CTX_DOUBLES > smem_limit_doubles) {
        L = smem_layout(hv->n_nodes, hv->n_edges, hv->Dn_pad, hv->bwp, hv->ld, hv->Wr, hv->ES, false, true);
        v.e_in_smem = 0;
        any_e_global = true;
      }
      if (L.total + CTX_DOUBLES > smem_limit_doubles) {
        L = smem_layout(hv->n_nodes, hv->n_edges, hv->Dn_pad, hv->bwp, hv->ld, hv->Wr, hv->ES, false, false);
        v.x_in_smem = 0;
        any_x_global = true;
        if (L.total + CTX_DOUBLES > smem_limit_doubles) return DEFSLAM_ETOOLARGE;
      }
      if (L.total + CTX_DOUBLES > smem_doubles) smem_doubles = L.total + CTX_DOUBLES;
      ws_band = std::max(ws_band, (size_t)hv->Dn_pad * hv->ld);
      ws_dinv = std::max(ws_dinv, (size_t)hv->nblk * 64);
      ws_cg = std::max(ws_cg, (size_t)8 * hv->ES);
      ws_F = std::max(ws_F, (size_t)NFACC * hv->n_facets);
      ws_S = std::max(ws_S, (size_t)NMSCR * M);
      ws_M = std::max(ws_M, M);
      ws_nf = std::max(ws_nf, (size_t)hv->n_facets + 1);
      ws_dp = std::max(ws_dp, (size_t)hv->Dn_pad);
    }
    if (any_x_global) {
      smem_doubles = 0;
      for (int i = 0; i < nprob; i++) {
        const PlanView *hv = hvs[i];
        views[i].x_in_smem = 0;
        views[i].row_nt = 0;
        const SmemLayout L = smem_layout(hv->n_nodes, hv->n_edges, hv->Dn_pad, hv->bwp, hv->ld, hv->Wr, hv->ES,
                                         views[i].e_in_smem != 0, false);
        if (L.total + CTX_DOUBLES > smem_doubles) smem_doubles = L.total + CTX_DOUBLES;
      }
    }
    return 0;
  }

  void pack_inputs(const defslam_sft_problem *p, uint8_t *host_in) const {
    /* ~52 KB per C2 frame: threads once the batch is a few MB */
    const size_t per_frame = slots.empty() ? 1 : std::max<size_t>(1, in_bytes / slots.size());
    host_parallel_for(slots.size(), std::max<size_t>(1, ((size_t)4 << 20) / per_frame), [&](size_t lo_, size_t hi_) {
    for (size_t i = lo_; i < hi_; i++) {
      const ProbSlot &s = slots[i];
      uint8_t *b = host_in + s.in_off;
      const size_t n = s.n_nodes, M = s.M;
      memcpy(b + s.o_nodes, p[i].node_xyz, 3 * n * sizeof(double));
      if (M) {
        memcpy(b + s.o_bary, p[i].match_bary, 3 * M * sizeof(double));
        memcpy(b + s.o_mnodes, p[i].match_nodes, 3 * M * sizeof(int));
        memcpy(b + s.o_uv, p[i].match_uv, 2 * M * sizeof(float));
        if (p[i].match_inv_sigma2) memcpy(b + s.o_isig, p[i].match_inv_sigma2, M * sizeof(float));
        else memset(b + s.o_isig, 0, M * sizeof(float));
      }
    }
    });
  }

  /* resolve the arena-relative pointers of every view */
  void bind(uint8_t *dev_in, uint8_t *dev_out) {
    for (size_t i = 0; i < slots.size(); i++) {
      const ProbSlot &s = slots[i];
      ProbView &v = views[i];
      const uint8_t *bi = dev_in + s.in_off;
      uint8_t *bo = dev_out + s.out_off;
      v.node_xyz = (const double *)(bi + s.o_nodes);
      v.match_bary = (const double *)(bi + s.o_bary);
      v.match_nodes = (const int *)(bi + s.o_mnodes);
      v.match_uv = (const float *)(bi + s.o_uv);
      v.match_isig = (const float *)(bi + s.o_isig);
      v.out_nodes = (double *)(bo + s.o_out_nodes);
      v.out_trace = (double *)(bo + s.o_trace);
      v.out_res = (ResultScalars *)(bo + s.o_res);
      v.out_outlier = (uint8_t *)(bo + s.o_outlier);
      v.out_role = (uint8_t *)(bo + s.o_role);
      v.out_H = v.mode == MODE_NORMAL_EQ ? (double *)(bo + s.o_H) : nullptr;
      v.out_b = v.mode == MODE_NORMAL_EQ ? (double *)(bo + s.o_b) : nullptr;
    }
  }

  /* scatter the output arena into the caller's result structs.  A problem
   * whose status is not OK leaves the caller's arrays untouched. */
  int unpack(const uint8_t *host_out, defslam_sft_result *r) const {
    int worst = 0;
    for (size_t i = 0; i < slots.size(); i++) {
      const ProbSlot &s = slots[i];
      const uint8_t *b = host_out + s.out_off;
      const ResultScalars *rs = (const ResultScalars *)(b + s.o_res);
      r[i].status = rs->status;
      if (rs->status != 0) { if (worst == 0) worst = rs->status; continue; }
      if (r[i].node_xyz_out) memcpy(r[i].node_xyz_out, b + s.o_out_nodes, 3 * (size_t)s.n_nodes * sizeof(double));
      if (r[i].outlier_out && s.M) memcpy(r[i].outlier_out, b + s.o_outlier, s.M);
      if (r[i].node_role_out) memcpy(r[i].node_role_out, b + s.o_role, s.n_nodes);
      memcpy(r[i].T_cw_out, rs->Tcw, sizeof(rs->Tcw));
      r[i].rep_error = rs->rep_error;
      r[i].n_inliers = rs->n_inliers;
      r[i].lm_iterations = rs->lm_iterations;
      r[i].lm_trials = rs->lm_trials;
      r[i].chi2_initial = rs->chi2_initial;
      r[i].chi2_final = rs->chi2_final;
      r[i].lambda_final = rs->lambda_final;
      if (r[i].trace && r[i].trace_capacity > 0) {
        const int rows = std::min(r[i].trace_capacity, std::min(s.trace_cap, rs->lm_iterations));
        memcpy(r[i].trace, b + s.o_trace, 4 * (size_t)rows * sizeof(double));
      }
    }
    return worst;
  }
};

}  // namespace ds
#endif
