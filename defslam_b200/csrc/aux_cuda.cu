/*
 * aux_cuda.cu -- CUDA kernels + C ABI of the template-construction helpers and
 * the bicubic B-spline (BBS) primitives.  Device arithmetic lives in
 * mesh_core.h / bbs_core.h; see include/defslam_b200.h for what each entry
 * point replaces in the reference.
 */
#include <vector>

#include "ds_runtime.h"
#include "mesh_core.h"
#include "newpts_core.h"

using namespace ds;

namespace {

/* grow-only device scratch of the calling thread: inputs and outputs of one
 * call are packed into it, one H2D and one D2H copy per call */
struct Scratch {
  DevBuf dev, host;
  Scratch() { host.pinned = true; }
};
Scratch &tl_scratch(int device) {
  static thread_local std::map<int, std::unique_ptr<Scratch>> tl;
  auto &s = tl[device];
  if (!s) s.reset(new Scratch);
  return *s;
}

struct Packer {
  size_t total = 0;
  size_t add(size_t bytes) { size_t o = total; total += (bytes + 255) & ~(size_t)255; return o; }
};

__global__ void mesh_laplacian_kernel(MeshLapArgs A) {
  __shared__ double red[40];
  Team team;
  team.tid = threadIdx.x;
  team.nthr = blockDim.x;
  mesh_laplacian_team(team, A, red);
}

__global__ void embed_points_kernel(int n, const double *X, int nf, const int *facets, int npts, const float *P,
                                    int *out_facet, int *out_nodes, float *out_bary) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x)
    embed_point(n, X, nf, facets, &P[3 * i], &out_facet[i], &out_nodes[3 * i], &out_bary[3 * i]);
}

__global__ void mappoints_kernel(const double *X, int npts, const int *nodes, const double *bary, float *out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x)
    mappoint_position(X, &nodes[3 * i], &bary[3 * i], &out[3 * i]);
}

__global__ void surface_vertices_kernel(BbsView s, const double *ctrl, int xs, int ys, float *out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < xs * ys; i += gridDim.x * blockDim.x)
    surface_vertex(s, ctrl, xs, ys, i, &out[3 * i]);
}

/* One thread per keypoint; the pixels of the keypoints that own a good map point are staged
 * through shared memory in tiles and every candidate tests its window against them. */
__global__ void new_map_points_kernel(int n, int rows, int cols, const float *kp_xy, const uint8_t *state,
                                      const float *surf, const float *Twc, uint8_t *action, float *world,
                                      int *n_new) {
  __shared__ int mx[256], my[256];
  __shared__ float T[16];
  const int ksz = cols / 20, anc = ksz / 2;
  if (Twc != nullptr && threadIdx.x < 16) T[threadIdx.x] = Twc[threadIdx.x];
  const int nblk_iter = (n + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
  for (int it = 0; it < nblk_iter; it++) {
    const int i = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int st = valid ? state[i] : 2;
    const int cx = valid ? (int)kp_xy[2 * i] : 0, cy = valid ? (int)kp_xy[2 * i + 1] : 0;
    bool occupied = false;
    for (int base = 0; base < n; base += blockDim.x) {
      __syncthreads();
      const int j = base + threadIdx.x;
      int px = -1, py = -1;
      if (j < n && state[j] == 1) { px = (int)kp_xy[2 * j]; py = (int)kp_xy[2 * j + 1]; }
      mx[threadIdx.x] = px; my[threadIdx.x] = py;
      __syncthreads();
      if (st == 0 && !occupied) {
        const int lim = n - base < (int)blockDim.x ? n - base : (int)blockDim.x;
        for (int t = 0; t < lim; t++) {
          if (mx[t] < 0) continue;
          if (window_hits(cx, mx[t], cols, ksz, anc) && window_hits(cy, my[t], rows, ksz, anc)) { occupied = true; break; }
        }
      }
    }
    if (valid) {
      const int act = st == 1 ? 1 : (st == 0 && !occupied ? 2 : 0);
      action[i] = (uint8_t)act;
      if (world != nullptr) {
        float w[3] = {0.f, 0.f, 0.f};
        if (act != 0) surface_point_to_world(T, &surf[3 * i], w);
        world[3 * i] = w[0]; world[3 * i + 1] = w[1]; world[3 * i + 2] = w[2];
      }
      if (act == 2) atomicAdd(n_new, 1);
    }
  }
}

/* nord derivative orders per site; val laid out [ord][site][valdim] */
__global__ void bbs_eval_kernel(BbsView s, const double *ctrl, int nsites, const double *u, const double *v, int nord,
                                int du0, int dv0, double *val) {
  const int du6[6] = {0, 1, 0, 2, 1, 0}, dv6[6] = {0, 0, 1, 0, 1, 2};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nsites; i += gridDim.x * blockDim.x) {
    const double uu = u[i], vv = v[i];
    for (int o = 0; o < nord; o++) {
      const int du = nord == 1 ? du0 : du6[o], dv = nord == 1 ? dv0 : dv6[o];
      double tmp[4];
      if (s.valdim <= 4) {
        bbs_eval_site(s, ctrl, uu, vv, du, dv, tmp);
        for (int d = 0; d < s.valdim; d++) val[((size_t)o * nsites + i) * s.valdim + d] = tmp[d];
      } else {
        bbs_eval_site(s, ctrl, uu, vv, du, dv, &val[((size_t)o * nsites + i) * s.valdim]);
      }
    }
  }
}

__global__ void bbs_coloc_kernel(BbsView s, int nsites, const double *u, const double *v, int du, int dv, double *Cm,
                                 int *err) {
  const int NC = s.nptsu * s.nptsv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nsites; i += gridDim.x * blockDim.x)
    if (!bbs_coloc_row(s, u[i], v[i], du, dv, &Cm[(size_t)i * NC])) atomicExch(err, 1);
}

__global__ void bbs_bending_kernel(BbsView s, double *B) {
  const int NC = s.nptsu * s.nptsv;
  const long long tot = (long long)NC * NC;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < tot; k += (long long)gridDim.x * blockDim.x)
    B[k] = bbs_bending_entry(s, (int)(k / NC), (int)(k % NC));
}

BbsView to_view(const defslam_bbs *b) {
  BbsView s;
  s.umin = b->umin; s.umax = b->umax; s.vmin = b->vmin; s.vmax = b->vmax;
  s.nptsu = b->nptsu; s.nptsv = b->nptsv; s.valdim = b->valdim;
  return s;
}
bool bbs_ok(const defslam_bbs *b) {
  return b && b->nptsu >= 4 && b->nptsv >= 4 && b->valdim >= 1 && b->umax > b->umin && b->vmax > b->vmin;
}
int grid_for(int n, int sm) {
  int g = (n + 255) / 256;
  if (g > sm * 8) g = sm * 8;
  return g < 1 ? 1 : g;
}

}  // namespace

extern "C" {

int defslam_mesh_laplacian(int32_t n_nodes, const double *node_xyz, int32_t n_facets, const int32_t *facets,
                           int32_t max_ring, int32_t *nbr_cnt, int32_t *nbr_idx, double *nbr_w,
                           uint8_t *node_boundary, double *node_kappa0, int32_t *n_edges_out, int32_t *edge_ab,
                           double *edge_len0, double *edge_median_len) {
  if (n_nodes <= 0 || n_facets < 0 || max_ring <= 0 || !node_xyz || (n_facets > 0 && !facets) || !nbr_cnt ||
      !nbr_idx || !nbr_w || !node_boundary || !node_kappa0 || !n_edges_out || !edge_ab || !edge_len0 ||
      !edge_median_len)
    return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  const size_t n = n_nodes, nf = n_facets, R = max_ring, C = 3 * nf;
  Packer in, out;
  const size_t o_X = in.add(3 * n * 8), o_F = in.add(3 * nf * 4 + 4);
  const size_t o_cnt = out.add(n * 4), o_idx = out.add(n * R * 4), o_w = out.add(n * R * 8), o_bd = out.add(n),
               o_k0 = out.add(n * 8), o_ne = out.add(16), o_ab = out.add(2 * C * 4 + 4), o_len = out.add(C * 8 + 8),
               o_med = out.add(8), o_st = out.add(8);
  Packer scr;
  const size_t o_cf = scr.add(C * 4 + 4), o_cp = scr.add(C * 4 + 4);
  Scratch &S = tl_scratch(ctx->device);
  const size_t host_need = in.total > out.total ? in.total : out.total;
  int rc;
  if ((rc = S.host.ensure(host_need)) || (rc = S.dev.ensure(in.total + out.total + scr.total))) return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total, *d_scr = d_out + out.total;
  memcpy(h + o_X, node_xyz, 3 * n * 8);
  if (nf) memcpy(h + o_F, facets, 3 * nf * 4);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  MeshLapArgs A;
  A.n = n_nodes; A.nf = n_facets; A.max_ring = max_ring;
  A.X = (const double *)(d_in + o_X); A.facets = (const int *)(d_in + o_F);
  A.nbr_cnt = (int *)(d_out + o_cnt); A.nbr_idx = (int *)(d_out + o_idx); A.nbr_w = (double *)(d_out + o_w);
  A.boundary = d_out + o_bd; A.kappa0 = (double *)(d_out + o_k0); A.n_edges = (int *)(d_out + o_ne);
  A.edge_ab = (int *)(d_out + o_ab); A.edge_len0 = (double *)(d_out + o_len); A.median = (double *)(d_out + o_med);
  A.status = (int *)(d_out + o_st);
  A.cand_first = (int *)(d_scr + o_cf); A.cand_pos = (int *)(d_scr + o_cp);
  mesh_laplacian_kernel<<<1, 1024, 0, ctx->stream>>>(A);
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, out.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  const int st = *(const int *)(h + o_st);
  if (st) return st;
  const int ne = *(const int *)(h + o_ne);
  memcpy(nbr_cnt, h + o_cnt, n * 4);
  memcpy(nbr_idx, h + o_idx, n * R * 4);
  memcpy(nbr_w, h + o_w, n * R * 8);
  memcpy(node_boundary, h + o_bd, n);
  memcpy(node_kappa0, h + o_k0, n * 8);
  *n_edges_out = ne;
  memcpy(edge_ab, h + o_ab, 2 * (size_t)ne * 4);
  memcpy(edge_len0, h + o_len, (size_t)ne * 8);
  *edge_median_len = *(const double *)(h + o_med);
  return DEFSLAM_OK;
}

int defslam_embed_points(int32_t n_nodes, const double *node_xyz, int32_t n_facets, const int32_t *facets,
                         int32_t n_points, const float *point_xyz, int32_t *out_facet, int32_t *out_nodes,
                         float *out_bary) {
  if (n_nodes <= 0 || n_facets < 0 || n_points < 0 || !node_xyz || (n_facets > 0 && !facets) ||
      (n_points > 0 && (!point_xyz || !out_facet || !out_nodes || !out_bary)))
    return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  if (n_points == 0) return DEFSLAM_OK;
  const size_t n = n_nodes, nf = n_facets, np = n_points;
  Packer in, out;
  const size_t o_X = in.add(3 * n * 8), o_F = in.add(3 * nf * 4 + 4), o_P = in.add(3 * np * 4);
  const size_t o_f = out.add(np * 4), o_n = out.add(3 * np * 4), o_b = out.add(3 * np * 4);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > out.total ? in.total : out.total)) || (rc = S.dev.ensure(in.total + out.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  memcpy(h + o_X, node_xyz, 3 * n * 8);
  if (nf) memcpy(h + o_F, facets, 3 * nf * 4);
  memcpy(h + o_P, point_xyz, 3 * np * 4);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  embed_points_kernel<<<grid_for(n_points, ctx->sm_count), 256, 0, ctx->stream>>>(
      n_nodes, (const double *)(d_in + o_X), n_facets, (const int *)(d_in + o_F), n_points,
      (const float *)(d_in + o_P), (int *)(d_out + o_f), (int *)(d_out + o_n), (float *)(d_out + o_b));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, out.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(out_facet, h + o_f, np * 4);
  memcpy(out_nodes, h + o_n, 3 * np * 4);
  memcpy(out_bary, h + o_b, 3 * np * 4);
  return DEFSLAM_OK;
}

int defslam_mappoints_recalculate(int32_t n_nodes, const double *node_xyz, int32_t n_points,
                                  const int32_t *point_nodes, const double *point_bary, float *point_xyz_out) {
  if (n_nodes <= 0 || n_points < 0 || !node_xyz || (n_points > 0 && (!point_nodes || !point_bary || !point_xyz_out)))
    return DEFSLAM_EBADARG;
  for (int i = 0; i < 3 * n_points; i++)
    if (point_nodes[i] < 0 || point_nodes[i] >= n_nodes) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  if (n_points == 0) return DEFSLAM_OK;
  const size_t n = n_nodes, np = n_points;
  Packer in, out;
  const size_t o_X = in.add(3 * n * 8), o_N = in.add(3 * np * 4), o_B = in.add(3 * np * 8);
  const size_t o_o = out.add(3 * np * 4);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > out.total ? in.total : out.total)) || (rc = S.dev.ensure(in.total + out.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  memcpy(h + o_X, node_xyz, 3 * n * 8);
  memcpy(h + o_N, point_nodes, 3 * np * 4);
  memcpy(h + o_B, point_bary, 3 * np * 8);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  mappoints_kernel<<<grid_for(n_points, ctx->sm_count), 256, 0, ctx->stream>>>(
      (const double *)(d_in + o_X), n_points, (const int *)(d_in + o_N), (const double *)(d_in + o_B),
      (float *)(d_out + o_o));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, out.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(point_xyz_out, h + o_o, 3 * np * 4);
  return DEFSLAM_OK;
}

static int bbs_eval_common(const defslam_bbs *bbs, const double *ctrl, int32_t nsites, const double *u,
                           const double *v, int nord, int du, int dv, double *val) {
  if (!bbs_ok(bbs) || !ctrl || nsites < 0 || (nsites > 0 && (!u || !v || !val))) return DEFSLAM_EBADARG;
  if (nord == 1 && (du < 0 || du > 2 || dv < 0 || dv > 2)) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  if (nsites == 0) return DEFSLAM_OK;
  const size_t NC = (size_t)bbs->nptsu * bbs->nptsv, ns = nsites, vd = bbs->valdim;
  Packer in, out;
  const size_t o_c = in.add(NC * vd * 8), o_u = in.add(ns * 8), o_v = in.add(ns * 8);
  const size_t o_o = out.add((size_t)nord * ns * vd * 8);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > out.total ? in.total : out.total)) || (rc = S.dev.ensure(in.total + out.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  memcpy(h + o_c, ctrl, NC * vd * 8);
  memcpy(h + o_u, u, ns * 8);
  memcpy(h + o_v, v, ns * 8);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  bbs_eval_kernel<<<grid_for(nsites, ctx->sm_count), 256, 0, ctx->stream>>>(
      to_view(bbs), (const double *)(d_in + o_c), nsites, (const double *)(d_in + o_u), (const double *)(d_in + o_v),
      nord, du, dv, (double *)(d_out + o_o));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, out.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(val, h + o_o, (size_t)nord * ns * vd * 8);
  return DEFSLAM_OK;
}

int defslam_bbs_eval(const defslam_bbs *bbs, const double *ctrl, int32_t nsites, const double *u, const double *v,
                     int32_t du, int32_t dv, double *val) {
  return bbs_eval_common(bbs, ctrl, nsites, u, v, 1, du, dv, val);
}

int defslam_bbs_eval6(const defslam_bbs *bbs, const double *ctrl, int32_t nsites, const double *u, const double *v,
                      double *val6) {
  return bbs_eval_common(bbs, ctrl, nsites, u, v, 6, 0, 0, val6);
}

int defslam_bbs_coloc(const defslam_bbs *bbs, int32_t nsites, const double *u, const double *v, int32_t du,
                      int32_t dv, double *Cm) {
  if (!bbs_ok(bbs) || nsites < 0 || (nsites > 0 && (!u || !v || !Cm)) || du < 0 || du > 2 || dv < 0 || dv > 2)
    return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  if (nsites == 0) return DEFSLAM_OK;
  const size_t NC = (size_t)bbs->nptsu * bbs->nptsv, ns = nsites;
  Packer in, out;
  const size_t o_u = in.add(ns * 8), o_v = in.add(ns * 8);
  const size_t o_C = out.add(ns * NC * 8), o_e = out.add(8);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > out.total ? in.total : out.total)) || (rc = S.dev.ensure(in.total + out.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  memcpy(h + o_u, u, ns * 8);
  memcpy(h + o_v, v, ns * 8);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  DS_CUDA_TRY(cudaMemsetAsync(d_out, 0, out.total, ctx->stream));
  bbs_coloc_kernel<<<grid_for(nsites, ctx->sm_count), 256, 0, ctx->stream>>>(
      to_view(bbs), nsites, (const double *)(d_in + o_u), (const double *)(d_in + o_v), du, dv, (double *)(d_out + o_C),
      (int *)(d_out + o_e));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, out.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (*(const int *)(h + o_e)) { /* a site outside the domain: the reference aborts leaving zeros */
    memset(Cm, 0, ns * NC * 8);
    return DEFSLAM_EBADARG;
  }
  memcpy(Cm, h + o_C, ns * NC * 8);
  return DEFSLAM_OK;
}

int defslam_bbs_bending(const defslam_bbs *bbs, double *B) {
  if (!bbs_ok(bbs) || !B) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  const size_t NC = (size_t)bbs->nptsu * bbs->nptsv;
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(NC * NC * 8)) || (rc = S.dev.ensure(NC * NC * 8))) return rc;
  bbs_bending_kernel<<<grid_for((int)(NC * NC), ctx->sm_count), 256, 0, ctx->stream>>>(to_view(bbs), (double *)S.dev.p);
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(S.host.p, S.dev.p, NC * NC * 8, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(B, S.host.p, NC * NC * 8);
  return DEFSLAM_OK;
}

int defslam_surface_vertices(const defslam_bbs *bbs, const double *ctrl_depth, int32_t xs, int32_t ys,
                             float *nodes_out) {
  if (!bbs_ok(bbs) || bbs->valdim != 1 || !ctrl_depth || xs < 2 || ys < 2 || !nodes_out) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  const size_t NC = (size_t)bbs->nptsu * bbs->nptsv, nv = (size_t)xs * ys;
  Packer in, out;
  const size_t o_c = in.add(NC * 8);
  const size_t o_o = out.add(nv * 3 * 4);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > out.total ? in.total : out.total)) || (rc = S.dev.ensure(in.total + out.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  memcpy(h + o_c, ctrl_depth, NC * 8);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  surface_vertices_kernel<<<grid_for((int)nv, ctx->sm_count), 256, 0, ctx->stream>>>(
      to_view(bbs), (const double *)(d_in + o_c), xs, ys, (float *)(d_out + o_o));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, out.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(nodes_out, h + o_o, nv * 3 * 4);
  return DEFSLAM_OK;
}

int defslam_new_map_points(const defslam_newpoints_problem *p, uint8_t *action_out, float *world_xyz_out,
                           int32_t *n_new_out) {
  if (!p || p->n_keypoints < 0 || p->rows <= 0 || p->cols <= 0 || !action_out || !n_new_out) return DEFSLAM_EBADARG;
  const int n = p->n_keypoints, ksz = p->cols / 20;
  if (n > 0 && (!p->kp_xy || !p->kp_state)) return DEFSLAM_EBADARG;
  if (ksz < 1 || ksz > p->rows || ksz > p->cols) return DEFSLAM_EBADARG;
  const bool place = world_xyz_out != nullptr;
  if (place && (!p->surf_xyz || !p->T_wc)) return DEFSLAM_EBADARG;
  for (int i = 0; i < n; i++) {
    const int x = (int)p->kp_xy[2 * i], y = (int)p->kp_xy[2 * i + 1];
    if (p->kp_state[i] > 2) return DEFSLAM_EBADARG;
    /* the reference indexes its mask with these: outside the image is out of bounds there */
    if (p->kp_state[i] != 2 && (x < 0 || x >= p->cols || y < 0 || y >= p->rows || !(p->kp_xy[2 * i] > -1.f) ||
                                !(p->kp_xy[2 * i + 1] > -1.f)))
      return DEFSLAM_EBADARG;
  }
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  *n_new_out = 0;
  if (n == 0) return DEFSLAM_OK;
  const size_t N = (size_t)n;
  Packer in, out;
  const size_t o_xy = in.add(N * 8), o_st = in.add(N), o_sf = in.add(place ? N * 12 : 0), o_T = in.add(64);
  const size_t o_cnt = out.add(4), o_act = out.add(N), o_w = out.add(place ? N * 12 : 0);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > out.total ? in.total : out.total)) || (rc = S.dev.ensure(in.total + out.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  memcpy(h + o_xy, p->kp_xy, N * 8);
  memcpy(h + o_st, p->kp_state, N);
  if (place) { memcpy(h + o_sf, p->surf_xyz, N * 12); memcpy(h + o_T, p->T_wc, 64); }
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  DS_CUDA_TRY(cudaMemsetAsync(d_out + o_cnt, 0, 4, ctx->stream));
  new_map_points_kernel<<<grid_for(n, ctx->sm_count), 256, 0, ctx->stream>>>(
      n, p->rows, p->cols, (const float *)(d_in + o_xy), d_in + o_st, place ? (const float *)(d_in + o_sf) : nullptr,
      place ? (const float *)(d_in + o_T) : nullptr, d_out + o_act, place ? (float *)(d_out + o_w) : nullptr,
      (int *)(d_out + o_cnt));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, out.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(n_new_out, h + o_cnt, 4);
  memcpy(action_out, h + o_act, N);
  if (place) memcpy(world_xyz_out, h + o_w, N * 12);
  return DEFSLAM_OK;
}

}  // extern "C"
