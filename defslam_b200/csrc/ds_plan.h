/*
 * ds_plan.h -- the solver-side plan of a template (mesh).
 *
 * A plan flattens what DefOptimizer.cc reads out of the Template pointer graph
 * every frame (Node::GetNeighbours/getEdges/weights, Edge::getDist,
 * Facet::getNodes, LaplacianMesh::GetMeanCurvatureInitial,
 * Template::getEdgeMeanSize -- Modules/Tracking/DefOptimizer.cc:293-507) into
 * flat arrays, and adds the sparsity plan of the normal matrix:
 *
 *   - node order = caller's order; unknown order = node 0 xyz, node 1 xyz, ...,
 *     then the 6 camera dofs LAST (arrowhead: band + 6-row border, no fill
 *     outside band and border);
 *   - two nodes are coupled iff they share a facet/edge (reprojection, stretch)
 *     or are both in {i} U N(i) of one curvature centre i (2-ring);
 *   - bw = scalar half-bandwidth of the node part in that order;
 *   - for every coupled block (p >= q) the lists of facets / curvature centres /
 *     edge that contribute to it, so that the assembly is a pure gather: every
 *     entry of H is produced by exactly one thread in a fixed order
 *     (deterministic, no atomics).
 *
 * Host code only builds and owns the arrays; PlanView is the POD the kernels see.
 */
#ifndef DS_PLAN_H_
#define DS_PLAN_H_

#include "ds_common.h"

namespace ds {

struct PlanView {
  int n_nodes, n_edges, n_facets, n_blk;
  int Dn;      /* 3*n_nodes                                              */
  int Dn_pad;  /* Dn rounded up to NB (padding rows are identity)        */
  int bw;      /* scalar half bandwidth                                   */
  int bwE;     /* bw rounded up to even: column of the diagonal in a band row */
  int ld;      /* band row length: >= bwE+1 and == 9 (mod 16), see sft_core.h */
  int ES;      /* row stride of the 8 border rows: >= Dn_pad, == 8 (mod 16)   */
  int bwp;     /* bw rounded up to TILE: extent of the trailing update    */
  int Wr;      /* rows of the sliding window (multiple of NB, >= NB+bwp)  */
  int nblk;    /* Dn_pad / NB                                             */
  int max_deg;
  double median_len;
  const double *rest;      /* [3n]  */
  const double *kappa0;    /* [n]   */
  const double *inv_len2;  /* [n]   sum over incident edges of 1/len0^2 */
  const double *nbr_w;     /* [nnz] mean-value weights                  */
  const double *nbr_c;     /* [nnz] w_ij / W_i                          */
  const double *sum_w;     /* [n]   W_i                                 */
  const double *edge_len0; /* [ne]  */
  const uint8_t *boundary; /* [n]   */
  const int *nbr_ptr, *nbr_idx;
  const int *edge_ab;      /* [2ne] */
  const int *facets;       /* [3nf] ascending node ids */
  const int *nf_ptr, *nf_ent;   /* node -> (facet<<2 | slot)                 */
  const int *ne_ptr, *ne_ent;   /* node -> (edge<<1 | node_is_b)             */
  const int *nc_ptr, *nc_ent;   /* node -> 2 ints: centre, cidx (-1: centre) */
  const int *blk_pq;            /* [2*n_blk] p >= q                          */
  const int *blk_fac_ptr, *blk_fac; /* (facet<<4 | slot_p<<2 | slot_q)       */
  const int *blk_ctr_ptr, *blk_ctr; /* 3 ints: centre, cidx_p, cidx_q        */
  const int *blk_edge;          /* [n_blk] edge joining p,q (p != q) or -1   */
  /* packed form of the same block plan, one dependent level shallower (diagonal blocks first):
   * blk_hdr[8*bi] = p, q, edge, ctr_ptr, n_ctr, n_fac, (F index 0 | pointer into blk_fidx when n_fac > 2), F index 1;
   * F index = pair_plane*n_facets + facet;  per centre entry: the centre and the two coefficients
   * (1 for the centre itself, -w/W for a neighbour) */
  const int *blk_hdr, *blk_fidx, *blk_ctr_i;
  const double *blk_ctr_pq;     /* [2*entries] cp, cq */
};

}  // namespace ds

#ifndef DS_DEVICE_ONLY
#include <algorithm>
#include <map>
#include <set>
#include <vector>

#include "../../include/defslam_b200.h"

namespace ds {

struct PlanHost {
  std::vector<double> dbl;
  std::vector<int> i32;
  std::vector<uint8_t> u8;
  PlanView v;          /* offsets stored as pointers relative to NULL until bind() */
  size_t o_rest, o_kappa0, o_inv_len2, o_nbr_w, o_nbr_c, o_sum_w, o_edge_len0;
  size_t o_nbr_ptr, o_nbr_idx, o_edge_ab, o_facets, o_nf_ptr, o_nf_ent, o_ne_ptr, o_ne_ent, o_nc_ptr, o_nc_ent,
      o_blk_pq, o_blk_fac_ptr, o_blk_fac, o_blk_ctr_ptr, o_blk_ctr, o_blk_edge, o_blk_hdr, o_blk_fidx, o_blk_ctr_i,
      o_blk_ctr_pq;

  /* returns DEFSLAM_OK or an error code */
  int build(const defslam_template_desc *d) {
    if (!d || d->n_nodes <= 0 || d->n_edges < 0 || d->n_facets < 0) return DEFSLAM_EBADARG;
    if (!d->node_rest_xyz || !d->node_boundary || !d->nbr_ptr || !d->node_kappa0) return DEFSLAM_EBADARG;
    if (d->n_edges > 0 && (!d->edge_ab || !d->edge_len0)) return DEFSLAM_EBADARG;
    if (d->n_facets > 0 && !d->facets) return DEFSLAM_EBADARG;
    const int n = d->n_nodes, ne = d->n_edges, nf = d->n_facets;
    const int nnz = d->nbr_ptr[n];
    if (nnz < 0 || (nnz > 0 && (!d->nbr_idx || !d->nbr_w))) return DEFSLAM_EBADARG;
    for (int i = 0; i < n; i++)
      if (d->nbr_ptr[i + 1] < d->nbr_ptr[i]) return DEFSLAM_EBADARG;
    for (int k = 0; k < nnz; k++)
      if (d->nbr_idx[k] < 0 || d->nbr_idx[k] >= n) return DEFSLAM_EBADARG;
    for (int e = 0; e < 2 * ne; e++)
      if (d->edge_ab[e] < 0 || d->edge_ab[e] >= n) return DEFSLAM_EBADARG;
    for (int f = 0; f < 3 * nf; f++)
      if (d->facets[f] < 0 || d->facets[f] >= n) return DEFSLAM_EBADARG;

    memset(&v, 0, sizeof(v));
    v.n_nodes = n; v.n_edges = ne; v.n_facets = nf;
    v.Dn = 3 * n;
    v.median_len = d->edge_median_len;
    dbl.clear(); i32.clear(); u8.clear();

    auto push_d = [&](const double *p, size_t cnt) { size_t o = dbl.size(); dbl.insert(dbl.end(), p, p + cnt); return o; };
    auto push_i = [&](const std::vector<int> &a) { size_t o = i32.size(); i32.insert(i32.end(), a.begin(), a.end()); return o; };

    o_rest = push_d(d->node_rest_xyz, 3 * (size_t)n);
    o_kappa0 = push_d(d->node_kappa0, n);
    o_nbr_w = push_d(d->nbr_w, nnz);
    o_edge_len0 = push_d(d->edge_len0, ne);

    /* W_i in neighbour order (sft_types.h:262-283 accumulates in vertex order) */
    std::vector<double> sumw(n, 0.0), nbrc(nnz, 0.0), invl2(n, 0.0);
    int max_deg = 0;
    for (int i = 0; i < n; i++) {
      double s = 0.0;
      for (int k = d->nbr_ptr[i]; k < d->nbr_ptr[i + 1]; k++) s = s + d->nbr_w[k];
      sumw[i] = s;
      for (int k = d->nbr_ptr[i]; k < d->nbr_ptr[i + 1]; k++) nbrc[k] = d->nbr_w[k] / s;
      max_deg = std::max(max_deg, d->nbr_ptr[i + 1] - d->nbr_ptr[i]);
    }
    v.max_deg = max_deg;

    /* node -> incident edges */
    std::vector<std::vector<int>> inc(n);
    for (int e = 0; e < ne; e++) {
      const int a = d->edge_ab[2 * e], b = d->edge_ab[2 * e + 1];
      if (a == b) return DEFSLAM_EBADARG;
      inc[a].push_back(e << 1);
      inc[b].push_back((e << 1) | 1);
    }
    for (int i = 0; i < n; i++)
      for (int ent : inc[i]) {
        const double l = d->edge_len0[ent >> 1];
        invl2[i] += 1.0 / (l * l);
      }
    o_sum_w = push_d(sumw.data(), n);
    o_nbr_c = push_d(nbrc.data(), nnz);
    o_inv_len2 = push_d(invl2.data(), n);

    u8.assign(d->node_boundary, d->node_boundary + n);

    /* facets with ascending node ids (Facet::getNodes() is a std::set) */
    std::vector<int> fac(3 * (size_t)nf);
    std::vector<std::vector<int>> nfac(n);
    for (int f = 0; f < nf; f++) {
      int t[3] = {d->facets[3 * f], d->facets[3 * f + 1], d->facets[3 * f + 2]};
      std::sort(t, t + 3);
      if (t[0] == t[1] || t[1] == t[2]) return DEFSLAM_EBADARG;
      for (int k = 0; k < 3; k++) {
        fac[3 * f + k] = t[k];
        nfac[t[k]].push_back((f << 2) | k);
      }
    }

    /* coupled blocks (p >= q) */
    struct Blk { std::vector<int> fac, ctr; int edge = -1; };
    std::map<std::pair<int, int>, Blk> blocks;
    for (int p = 0; p < n; p++) blocks[{p, p}];
    for (int f = 0; f < nf; f++)
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          const int p = fac[3 * f + a], q = fac[3 * f + b];
          if (p < q) continue;
          blocks[{p, q}].fac.push_back((f << 4) | (a << 2) | b);
        }
    for (int e = 0; e < ne; e++) {
      const int a = d->edge_ab[2 * e], b = d->edge_ab[2 * e + 1];
      blocks[{std::max(a, b), std::min(a, b)}].edge = e;
    }
    std::vector<std::vector<int>> nctr(n);
    for (int i = 0; i < n; i++) {
      /* vertex list of the curvature edge: centre, then neighbours */
      std::vector<std::pair<int, int>> vs;  /* (node, cidx) */
      vs.push_back({i, -1});
      for (int k = d->nbr_ptr[i]; k < d->nbr_ptr[i + 1]; k++) vs.push_back({d->nbr_idx[k], k});
      for (auto &a : vs) {
        nctr[a.first].push_back(i);
        nctr[a.first].push_back(a.second);
        for (auto &b : vs) {
          if (a.first < b.first) continue;
          if (a.first == b.first && a.second != b.second) return DEFSLAM_EBADARG; /* node twice in a ring */
          Blk &B = blocks[{a.first, b.first}];
          B.ctr.push_back(i); B.ctr.push_back(a.second); B.ctr.push_back(b.second);
        }
      }
    }
    int bwn = 0;
    std::vector<int> blk_pq, blk_fac_ptr(1, 0), blk_fac, blk_ctr_ptr(1, 0), blk_ctr, blk_edge;
    /* diagonal blocks first: they carry the longer dependent chain (stretch edges of the node), so
     * they share warps with each other instead of stalling warps of off-diagonal blocks */
    std::vector<std::pair<std::pair<int, int>, const Blk *>> order;
    for (auto &kv : blocks) if (kv.first.first == kv.first.second) order.push_back({kv.first, &kv.second});
    for (auto &kv : blocks) if (kv.first.first != kv.first.second) order.push_back({kv.first, &kv.second});
    std::vector<int> blk_hdr, blk_fidx, blk_ctr_i;
    std::vector<double> blk_ctr_pq;
    for (auto &ov : order) {
      struct { std::pair<int, int> first; const Blk &second; } kv = {ov.first, *ov.second};
      {
        int hdr[8] = {kv.first.first, kv.first.second, kv.second.edge, (int)blk_ctr_i.size(), (int)kv.second.ctr.size() / 3,
                      (int)kv.second.fac.size(), -1, -1};
        std::vector<int> fi;
        for (int ent : kv.second.fac) {
          const int f = ent >> 4, sp = (ent >> 2) & 3, sq = ent & 3;
          const int lo = std::min(sp, sq), hi = std::max(sp, sq);
          const int idx = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
          fi.push_back(idx * nf + f);
        }
        if (fi.size() <= 2) { for (size_t k = 0; k < fi.size(); k++) hdr[6 + k] = fi[k]; }
        else { hdr[6] = (int)blk_fidx.size(); blk_fidx.insert(blk_fidx.end(), fi.begin(), fi.end()); }
        blk_hdr.insert(blk_hdr.end(), hdr, hdr + 8);
        for (size_t k = 0; k + 2 < kv.second.ctr.size() + 0; k += 3) {
          const int ip = kv.second.ctr[k + 1], iq = kv.second.ctr[k + 2];
          blk_ctr_i.push_back(kv.second.ctr[k]);
          blk_ctr_pq.push_back(ip < 0 ? 1.0 : -nbrc[ip]);
          blk_ctr_pq.push_back(iq < 0 ? 1.0 : -nbrc[iq]);
        }
      }
      blk_pq.push_back(kv.first.first);
      blk_pq.push_back(kv.first.second);
      bwn = std::max(bwn, kv.first.first - kv.first.second);
      blk_fac.insert(blk_fac.end(), kv.second.fac.begin(), kv.second.fac.end());
      blk_fac_ptr.push_back((int)blk_fac.size());
      blk_ctr.insert(blk_ctr.end(), kv.second.ctr.begin(), kv.second.ctr.end());
      blk_ctr_ptr.push_back((int)blk_ctr.size() / 3);
      blk_edge.push_back(kv.second.edge);
    }
    v.n_blk = (int)blk_edge.size();
    v.Dn_pad = round_up(v.Dn, NB);
    v.bw = std::min(3 * bwn + 2, v.Dn - 1);
    if (v.bw < 2) v.bw = 2;
    v.bwE = round_up(v.bw, 2);
    v.ld = v.bwE + 1;
    while ((v.ld & 15) != 9) v.ld++;
    v.ES = (v.Dn_pad & 15) == 8 ? v.Dn_pad : v.Dn_pad + 8;
    v.bwp = round_up(v.bw, TILE);
    v.Wr = round_up(NB + v.bwp, NB);
    v.nblk = v.Dn_pad / NB;

    auto csr = [&](const std::vector<std::vector<int>> &l, int per, std::vector<int> &ptr, std::vector<int> &ent) {
      ptr.assign(1, 0);
      ent.clear();
      for (auto &r : l) {
        ent.insert(ent.end(), r.begin(), r.end());
        ptr.push_back((int)ent.size() / per);
      }
    };
    std::vector<int> p1, e1;
    o_nbr_ptr = push_i(std::vector<int>(d->nbr_ptr, d->nbr_ptr + n + 1));
    o_nbr_idx = push_i(std::vector<int>(d->nbr_idx, d->nbr_idx + nnz));
    o_edge_ab = push_i(std::vector<int>(d->edge_ab, d->edge_ab + 2 * (size_t)ne));
    o_facets = push_i(fac);
    csr(nfac, 1, p1, e1); o_nf_ptr = push_i(p1); o_nf_ent = push_i(e1);
    csr(inc, 1, p1, e1);  o_ne_ptr = push_i(p1); o_ne_ent = push_i(e1);
    csr(nctr, 2, p1, e1); o_nc_ptr = push_i(p1); o_nc_ent = push_i(e1);
    o_blk_pq = push_i(blk_pq);
    o_blk_fac_ptr = push_i(blk_fac_ptr); o_blk_fac = push_i(blk_fac);
    o_blk_ctr_ptr = push_i(blk_ctr_ptr); o_blk_ctr = push_i(blk_ctr);
    o_blk_edge = push_i(blk_edge);
    if (blk_fidx.empty()) blk_fidx.push_back(0);
    while (i32.size() & 3) i32.push_back(0);           /* headers are read as two 16-byte loads */
    o_blk_hdr = push_i(blk_hdr);
    o_blk_fidx = push_i(blk_fidx);
    o_blk_ctr_i = push_i(blk_ctr_i);
    o_blk_ctr_pq = push_d(blk_ctr_pq.data(), blk_ctr_pq.size());
    return DEFSLAM_OK;
  }

  /* a view whose pointers address the given copies of the three arenas */
  PlanView bind(const double *D, const int *I, const uint8_t *U) const {
    PlanView w = v;
    w.rest = D + o_rest; w.kappa0 = D + o_kappa0; w.inv_len2 = D + o_inv_len2;
    w.nbr_w = D + o_nbr_w; w.nbr_c = D + o_nbr_c; w.sum_w = D + o_sum_w; w.edge_len0 = D + o_edge_len0;
    w.boundary = U;
    w.nbr_ptr = I + o_nbr_ptr; w.nbr_idx = I + o_nbr_idx; w.edge_ab = I + o_edge_ab; w.facets = I + o_facets;
    w.nf_ptr = I + o_nf_ptr; w.nf_ent = I + o_nf_ent; w.ne_ptr = I + o_ne_ptr; w.ne_ent = I + o_ne_ent;
    w.nc_ptr = I + o_nc_ptr; w.nc_ent = I + o_nc_ent;
    w.blk_pq = I + o_blk_pq; w.blk_fac_ptr = I + o_blk_fac_ptr; w.blk_fac = I + o_blk_fac;
    w.blk_ctr_ptr = I + o_blk_ctr_ptr; w.blk_ctr = I + o_blk_ctr; w.blk_edge = I + o_blk_edge;
    w.blk_hdr = I + o_blk_hdr; w.blk_fidx = I + o_blk_fidx; w.blk_ctr_i = I + o_blk_ctr_i;
    w.blk_ctr_pq = D + o_blk_ctr_pq;
    return w;
  }
  PlanView host_view() const { return bind(dbl.data(), i32.data(), u8.data()); }
};

}  // namespace ds
#endif /* DS_DEVICE_ONLY */
#endif
