/*
 * nrsfm_cuda.cu -- CUDA kernels + C ABI of the NRSfM mapping stages
 * (Schwarp fit, isometric normals, shape-from-normals).  Device arithmetic lives
 * in nrsfm_core.h; include/defslam_b200.h says what each entry point replaces.
 *
 * Every batched call packs its inputs into one pinned arena (one H2D copy),
 * launches one kernel and reads one output arena back (one D2H copy).  Units
 * (keyframe pairs / map points / keyframes) are independent: persistent CTAs
 * take the next unit from an atomic counter.
 */
#include <string.h>

#include <vector>

#include "ds_host.h"
#include "ds_runtime.h"
#include "nrsfm_core.h"
#include "sim3_core.h"

using namespace ds;

namespace {

struct Scratch {
  DevBuf dev, host, ws;
  /* two side streams + events for calls that pipeline their batch in chunks (normals) */
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_done[16] = {}, ev_k0[16] = {}, ev_k1[16] = {};
  bool side_ok = false;
  Scratch() { host.pinned = true; }
  int ensure_side() {
    if (side_ok) return 0;
    for (int i = 0; i < 2; i++) DS_CUDA_TRY(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
    for (int i = 0; i < 16; i++) {
      DS_CUDA_TRY(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
      DS_CUDA_TRY(cudaEventCreate(&ev_k0[i]));
      DS_CUDA_TRY(cudaEventCreate(&ev_k1[i]));
    }
    side_ok = true;
    return 0;
  }
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

constexpr int NRSFM_THREADS = NRSFM_THREADS_;

BbsView to_view(const defslam_bbs *b) {
  BbsView s;
  s.umin = b->umin; s.umax = b->umax; s.vmin = b->vmin; s.vmax = b->vmax;
  s.nptsu = b->nptsu; s.nptsv = b->nptsv; s.valdim = b->valdim;
  return s;
}
bool bbs_ok(const defslam_bbs *b, int valdim) {
  return b->nptsu >= 4 && b->nptsv >= 4 && b->valdim == valdim && b->umax > b->umin && b->vmax > b->vmin &&
         b->nptsu <= 64 && b->nptsv <= 64;
}

/* ------------------------------------------------------------------ kernels ---------- */

__global__ void __launch_bounds__(NRSFM_THREADS, 1)
schwarp_fit_kernel(const SchwarpProb *probs, int nprob, uint8_t *ws_base, size_t ws_stride, int ws_nu, int ws_nv,
                   int ws_nmax, int *counter) {
  extern __shared__ double sh[];
  __shared__ int s_next;
  Team team;
  team.tid = threadIdx.x;
  team.nthr = blockDim.x;
  uint8_t *b = ws_base + (size_t)blockIdx.x * ws_stride;
  const SchwarpSizes z = schwarp_ws_sizes(ws_nu, ws_nv, ws_nmax);
  SchwarpWs ws;
  ws.cell = (int *)(b + z.cell); ws.cstart = (int *)(b + z.cstart); ws.perm = (int *)(b + z.perm);
  ws.taps = (double *)(b + z.taps); ws.CtC = (double *)(b + z.CtC); ws.Hb = (double *)(b + z.Hb);
  ws.Lb = (double *)(b + z.Lb); ws.Js = (double *)(b + z.Js); ws.rdata = (double *)(b + z.rdata);
  ws.sdv = (double *)(b + z.sdv);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_next = atomicAdd(counter, 1);
    __syncthreads();
    const int i = s_next;
    if (i >= nprob) break;
    schwarp_fit_one(team, probs[i], ws, sh);
  }
}

__global__ void schwarp_rows_kernel(SchwarpProb P, int nrows, double *r, double *J) {
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += gridDim.x * blockDim.x)
    schwarp_row(P, row, r, J);
}

/* NaN scrub of the first nscrub control coordinates (DefORBmatcher.cc:149-154) */
__global__ void schwarp_scrub_kernel(double *x, int nscrub) {
  for (int i = threadIdx.x; i < nscrub; i += blockDim.x)
    if (isnan(x[i])) x[i] = 0.0;
}

/* ceres::Problem::Evaluate residuals of the one Warp block under HuberLoss(5.77) (Corrector: scale sqrt(rho')),
 * then the reference's test residuals[2i]^2 + residuals[2i+1]^2 > 20 (DefORBmatcher.cc:166-180).  One CTA; the
 * block's squared norm is summed in a fixed order. */
__global__ void schwarp_initial_filter_kernel(const double *r, int n, uint8_t *keep, double *err) {
  __shared__ double part[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) s += r[i] * r[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  const double sq = part[0], delta = 5.77;
  const double rho1 = sq <= delta * delta ? 1.0 : delta / sqrt(sq);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double e = rho1 * (r[2 * i] * r[2 * i] + r[2 * i + 1] * r[2 * i + 1]);
    err[i] = e;
    keep[i] = e > 20.0 ? 0 : 1;
  }
}

/* map points [point0, point1) of the problem */
__global__ void normals_kernel(NormalsProb P, int point0, int point1) {
  for (int i = point0 + blockIdx.x * blockDim.x + threadIdx.x; i < point1; i += gridDim.x * blockDim.x) normals_point(P, i);
}

__global__ void poly_kernel(int npairs, const float *J12, const float *H12, const float *I1, const float *I2,
                            double *eq1, double *eq2) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npairs; i += gridDim.x * blockDim.x)
    pair_polynomials(J12 + 4 * i, H12 + 6 * i, I1 + 2 * i, I2 + 2 * i, 0, eq1 + 10 * i, eq2 + 10 * i, 1);
}

__global__ void __launch_bounds__(NRSFM_THREADS, 1)
sfn_solve_kernel(const SfnProb *probs, int nprob, uint8_t *ws_base, size_t ws_stride, int ws_nu, int ws_nv,
                 int ws_nmax, int n_in_smem, int *counter) {
  extern __shared__ double sh[];
  __shared__ int s_next;
  Team team;
  team.tid = threadIdx.x;
  team.nthr = blockDim.x;
  uint8_t *b = ws_base + (size_t)blockIdx.x * ws_stride;
  const SfnSizes z = sfn_ws_sizes(ws_nu, ws_nv, ws_nmax);
  SfnWs ws;
  ws.cell = (int *)(b + z.cell); ws.cstart = (int *)(b + z.cstart); ws.perm = (int *)(b + z.perm);
  ws.taps = (double *)(b + z.taps); ws.mrow = (double *)(b + z.mrow); ws.B = (double *)(b + z.B);
  ws.N = (double *)(b + z.N); ws.res = (double *)(b + z.res); ws.G = (double *)(b + z.G);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_next = atomicAdd(counter, 1);
    __syncthreads();
    const int i = s_next;
    if (i >= nprob) break;
    sfn_solve_one(team, probs[i], ws, sh, n_in_smem);
  }
}

__global__ void sfn_rows_kernel(SfnProb P, int nrows, double *A, double *b) {
  __shared__ double ci[48];
  Team team;
  team.tid = threadIdx.x;
  team.nthr = blockDim.x;
  fill_cell_integrals(team, ci);
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += gridDim.x * blockDim.x)
    sfn_system_row(P, ci, row, A, b);
}

constexpr int SIM3_THREADS = 256;

__global__ void __launch_bounds__(SIM3_THREADS)
sim3_register_kernel(const Sim3Prob *probs, int nprob) {
  __shared__ double red[36 + 36 * (SIM3_THREADS / 32) + 40];
  Team team;
  team.tid = threadIdx.x;
  team.nthr = blockDim.x;
  for (int i = blockIdx.x; i < nprob; i += gridDim.x) {
    __syncthreads();
    sim3_register_one(team, probs[i], red);
  }
}

__global__ void __launch_bounds__(512)
scale_min_median_kernel(int n, const float *mono, const float *stereo, uint64_t seed, float *out) {
  extern __shared__ float mm_buf[];
  __shared__ int cnt;
  __shared__ double sc[4];
  Team team;
  team.tid = threadIdx.x;
  team.nthr = blockDim.x;
  scale_min_median_team(team, n, mono, stereo, seed, mm_buf, &cnt, sc, out);
}

int grid_for(int n, int sm) {
  int g = (n + 255) / 256;
  if (g > sm * 8) g = sm * 8;
  return g < 1 ? 1 : g;
}

int persistent_grid(int nprob, int sm) { return nprob < sm ? nprob : sm; }

}  // namespace

extern "C" {

/* ------------------------------------------------------------------ Schwarp ---------- */

int defslam_schwarp_fit_batched(int32_t nprob, const defslam_schwarp_problem *p, defslam_diffprop *out,
                                int32_t device) {
  if (nprob < 0 || (nprob > 0 && (!p || !out))) return DEFSLAM_EBADARG;
  int nu = 0, nv = 0, nmax = 0;
  for (int i = 0; i < nprob; i++) {
    if (!bbs_ok(&p[i].bbs, 2) || p[i].n_matches <= 0 || !p[i].kp1 || !p[i].kp2 || !p[i].inv_sigma || !p[i].x ||
        p[i].max_iterations < 0)
      return DEFSLAM_EBADARG;
    if (p[i].bbs.nptsu > nu) nu = p[i].bbs.nptsu;
    if (p[i].bbs.nptsv > nv) nv = p[i].bbs.nptsv;
    if (p[i].n_matches > nmax) nmax = p[i].n_matches;
  }
  DeviceGuard device_guard_;
  DevCtx *ctx = get_ctx(device);
  if (!ctx) return DEFSLAM_ECUDA;
  if (nprob == 0) return DEFSLAM_OK;
  const size_t smem = sizeof(double) * (size_t)schwarp_smem(nu, nv).total;
  if (smem > (size_t)ctx->smem_optin) return DEFSLAM_ETOOLARGE;

  /* arenas */
  Packer in, outp;
  std::vector<size_t> o_kp1(nprob), o_kp2(nprob), o_sig(nprob), o_x(nprob), o_uv(nprob), o_j12(nprob), o_j21(nprob),
      o_h12(nprob), o_keep(nprob), o_sc(nprob);
  const size_t o_probs = in.add(sizeof(SchwarpProb) * nprob);
  const size_t o_counter = in.add(sizeof(int) * 16); /* one work counter per chunk */
  /* x travels in and comes back: all x first in the output arena, so that the upload is one
   * contiguous span (inputs + x) */
  for (int i = 0; i < nprob; i++) o_x[i] = outp.add(16 * (size_t)p[i].bbs.nptsu * p[i].bbs.nptsv);
  const size_t x_total = outp.total;
  for (int i = 0; i < nprob; i++) {
    const size_t n = p[i].n_matches;
    o_kp1[i] = in.add(8 * n); o_kp2[i] = in.add(8 * n); o_sig[i] = in.add(4 * n);
    o_uv[i] = outp.add(8 * n); o_j12[i] = outp.add(16 * n); o_j21[i] = outp.add(16 * n); o_h12[i] = outp.add(24 * n);
    o_keep[i] = outp.add(n); o_sc[i] = outp.add(8 * 8);
  }
  const int grid = persistent_grid(nprob, ctx->sm_count);
  const size_t ws_stride = schwarp_ws_sizes(nu, nv, nmax).total;
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total + outp.total)) || (rc = S.dev.ensure(in.total + outp.total)) ||
      (rc = S.ws.ensure(ws_stride * grid)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total, *h_out = h + in.total;
  SchwarpProb *hp = (SchwarpProb *)(h + o_probs);
  /* Large batches run as a pipeline of chunks of whole waves (one fit per CTA): the host packs chunk c+1 and unpacks
   * chunk c-1 (several threads each) while the device works on chunk c; uploads, kernels (one stream: they share the
   * per-CTA workspace) and downloads on three streams.  Serially the call was 1.7x slower than its kernel. */
  const int nchunk = (nprob >= 2 * ctx->sm_count && getenv("DEFSLAM_NO_PIPELINE") == nullptr)
                         ? std::min(16, (nprob + ctx->sm_count - 1) / ctx->sm_count) : 1;
  if (nchunk > 1 && (rc = S.ensure_side())) return rc;
  auto chunk_lo = [&](int c) { return (int)((long long)nprob * c / nchunk); };
  auto pack = [&](int i_lo, int i_hi) {
    host_parallel_for((size_t)(i_hi - i_lo), 32, [&](size_t a_, size_t b_) {
      for (size_t i = i_lo + a_; i < i_lo + b_; i++) {
        const size_t n = p[i].n_matches, NC = (size_t)p[i].bbs.nptsu * p[i].bbs.nptsv;
        memcpy(h + o_kp1[i], p[i].kp1, 8 * n);
        memcpy(h + o_kp2[i], p[i].kp2, 8 * n);
        memcpy(h + o_sig[i], p[i].inv_sigma, 4 * n);
        memcpy(h_out + o_x[i], p[i].x, 16 * NC);
        SchwarpProb &P = hp[i];
        P.bbs = to_view(&p[i].bbs);
        P.n = p[i].n_matches;
        P.kp1 = (const float *)(d_in + o_kp1[i]); P.kp2 = (const float *)(d_in + o_kp2[i]);
        P.isig = (const float *)(d_in + o_sig[i]);
        P.lambda = p[i].lambda; P.fx = p[i].fx; P.fy = p[i].fy; P.px_fx = p[i].px_fx; P.px_fy = p[i].px_fy;
        P.max_iterations = p[i].max_iterations; P.initialize = p[i].initialize;
        P.x = (double *)(d_out + o_x[i]);
        P.warp_uv = (float *)(d_out + o_uv[i]); P.J12 = (float *)(d_out + o_j12[i]); P.J21 = (float *)(d_out + o_j21[i]);
        P.H12 = (float *)(d_out + o_h12[i]); P.keep = d_out + o_keep[i]; P.scalars = (double *)(d_out + o_sc[i]);
      }
    });
  };
  auto unpack = [&](int i_lo, int i_hi) {
    host_parallel_for((size_t)(i_hi - i_lo), 32, [&](size_t a_, size_t b_) {
      for (size_t i = i_lo + a_; i < i_lo + b_; i++) {
        const size_t n = p[i].n_matches, NC = (size_t)p[i].bbs.nptsu * p[i].bbs.nptsv;
        /* a failed fit leaves its outputs untouched, like the reference keeping its previous estimate */
        if ((int)((const double *)(h_out + o_sc[i]))[4] != SCHWARP_OK) continue;
        memcpy(p[i].x, h_out + o_x[i], 16 * NC);
        if (out[i].warp_uv) memcpy(out[i].warp_uv, h_out + o_uv[i], 8 * n);
        if (out[i].J12) memcpy(out[i].J12, h_out + o_j12[i], 16 * n);
        if (out[i].J21) memcpy(out[i].J21, h_out + o_j21[i], 16 * n);
        if (out[i].H12) memcpy(out[i].H12, h_out + o_h12[i], 24 * n);
        if (out[i].keep) memcpy(out[i].keep, h_out + o_keep[i], n);
      }
    });
  };
  memset(h + o_counter, 0, sizeof(int) * 16);
  DS_CUDA_TRY(raise_dynamic_smem((const void *)schwarp_fit_kernel, ctx->device, (int)smem));
  /* spans of problems [i_lo, i_hi) in the arenas (the per-problem blocks are laid out in problem order) */
  auto span_in = [&](int i_lo, int i_hi, size_t &off, size_t &len) {
    off = o_kp1[i_lo];
    len = (i_hi < nprob ? o_kp1[i_hi] : in.total) - off;
  };
  auto span_x = [&](int i_lo, int i_hi, size_t &off, size_t &len) {
    off = o_x[i_lo];
    len = (i_hi < nprob ? o_x[i_hi] : x_total) - off;
  };
  auto span_out = [&](int i_lo, int i_hi, size_t &off, size_t &len) {
    off = o_uv[i_lo];
    len = (i_hi < nprob ? o_uv[i_hi] : outp.total) - off;
  };
  float ms_total = 0.f;
  if (nchunk == 1) {
    pack(0, nprob);
    DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total + x_total, cudaMemcpyHostToDevice, ctx->stream));
    DS_CUDA_TRY(cudaEventRecord(ctx->e0, ctx->stream));
    schwarp_fit_kernel<<<grid, NRSFM_THREADS, smem, ctx->stream>>>((const SchwarpProb *)(d_in + o_probs), nprob,
                                                                   (uint8_t *)S.ws.p, ws_stride, nu, nv, nmax,
                                                                   (int *)(d_in + o_counter));
    DS_CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    DS_CUDA_TRY(cudaEventRecord(ctx->e1, ctx->stream));
    DS_CUDA_TRY(cudaMemcpyAsync(h_out, d_out, outp.total, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ms_total, ctx->e0, ctx->e1);
    unpack(0, nprob);
  } else {
    cudaStream_t s_up = S.side[0], s_dn = S.side[1], s_k = ctx->stream;
    /* the problem descriptors and the counters go up once the first chunk is packed; later chunks only rewrite
     * their own descriptors, so all of them are packed first (cheap: no array copies) -- done inside pack() per
     * chunk, uploaded per chunk below */
    for (int c = 0; c < nchunk; c++) {
      const int i_lo = chunk_lo(c), i_hi = chunk_lo(c + 1);
      pack(i_lo, i_hi);
      size_t off, len;
      if (c == 0) DS_CUDA_TRY(cudaMemcpyAsync(d_in + o_counter, h + o_counter, sizeof(int) * 16, cudaMemcpyHostToDevice, s_up));
      DS_CUDA_TRY(cudaMemcpyAsync(d_in + o_probs + sizeof(SchwarpProb) * i_lo, h + o_probs + sizeof(SchwarpProb) * i_lo,
                                  sizeof(SchwarpProb) * (i_hi - i_lo), cudaMemcpyHostToDevice, s_up));
      span_in(i_lo, i_hi, off, len);
      DS_CUDA_TRY(cudaMemcpyAsync(d_in + off, h + off, len, cudaMemcpyHostToDevice, s_up));
      span_x(i_lo, i_hi, off, len);
      DS_CUDA_TRY(cudaMemcpyAsync(d_out + off, h_out + off, len, cudaMemcpyHostToDevice, s_up));
      DS_CUDA_TRY(cudaEventRecord(S.ev_done[c], s_up)); /* (reused below for the download of the same chunk) */
      DS_CUDA_TRY(cudaStreamWaitEvent(s_k, S.ev_done[c], 0));
      DS_CUDA_TRY(cudaEventRecord(S.ev_k0[c], s_k));
      const int g = persistent_grid(i_hi - i_lo, ctx->sm_count);
      schwarp_fit_kernel<<<g, NRSFM_THREADS, smem, s_k>>>((const SchwarpProb *)(d_in + o_probs) + i_lo, i_hi - i_lo,
                                                          (uint8_t *)S.ws.p, ws_stride, nu, nv, nmax,
                                                          (int *)(d_in + o_counter) + c);
      DS_CUDA_TRY(cudaGetLastError());
      g_launches.fetch_add(1);
      DS_CUDA_TRY(cudaEventRecord(S.ev_k1[c], s_k));
      DS_CUDA_TRY(cudaStreamWaitEvent(s_dn, S.ev_k1[c], 0));
      span_x(i_lo, i_hi, off, len);
      DS_CUDA_TRY(cudaMemcpyAsync(h_out + off, d_out + off, len, cudaMemcpyDeviceToHost, s_dn));
      span_out(i_lo, i_hi, off, len);
      DS_CUDA_TRY(cudaMemcpyAsync(h_out + off, d_out + off, len, cudaMemcpyDeviceToHost, s_dn));
      if (c >= 1) {
        DS_CUDA_TRY(cudaEventSynchronize(S.ev_done[c - 1]));
        unpack(chunk_lo(c - 1), chunk_lo(c));
      }
      DS_CUDA_TRY(cudaEventRecord(S.ev_done[c], s_dn));
    }
    DS_CUDA_TRY(cudaEventSynchronize(S.ev_done[nchunk - 1]));
    unpack(chunk_lo(nchunk - 1), nprob);
    for (int c = 0; c < nchunk; c++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, S.ev_k0[c], S.ev_k1[c]);
      ms_total += ms;
    }
  }
  g_last_kernel_ms = ms_total;
  int worst = DEFSLAM_OK;
  for (int i = 0; i < nprob; i++) {
    const double *sc = (const double *)(h_out + o_sc[i]);
    const int st = (int)sc[4];
    out[i].cost_initial = sc[0]; out[i].cost_final = sc[1];
    out[i].iterations = (int)sc[2]; out[i].accepted = (int)sc[3];
    if (st != SCHWARP_OK && worst == DEFSLAM_OK) worst = st == SCHWARP_OUT_OF_DOMAIN ? DEFSLAM_EBADARG : DEFSLAM_ENUMERIC;
  }
  return worst;
}

int defslam_schwarp_fit(const defslam_schwarp_problem *p, defslam_diffprop *out) {
  return defslam_schwarp_fit_batched(1, p, out, -1);
}

int defslam_schwarp_evaluate(const defslam_schwarp_problem *p, double *r, double *J) {
  if (!p || !r || !bbs_ok(&p->bbs, 2) || p->n_matches < 0 || !p->x || (p->n_matches > 0 && (!p->kp1 || !p->kp2 || !p->inv_sigma)))
    return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  const size_t n = p->n_matches, NC = (size_t)p->bbs.nptsu * p->bbs.nptsv, NR = 2 * n + 4 * NC, NP = 2 * NC;
  Packer in, outp;
  const size_t o_kp1 = in.add(8 * n + 8), o_kp2 = in.add(8 * n + 8), o_sig = in.add(4 * n + 8), o_x = in.add(8 * NP);
  const size_t o_r = outp.add(8 * NR), o_J = outp.add(J ? 8 * NR * NP : 8);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > outp.total ? in.total : outp.total)) || (rc = S.dev.ensure(in.total + outp.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  if (n) { memcpy(h + o_kp1, p->kp1, 8 * n); memcpy(h + o_kp2, p->kp2, 8 * n); memcpy(h + o_sig, p->inv_sigma, 4 * n); }
  memcpy(h + o_x, p->x, 8 * NP);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  if (J) DS_CUDA_TRY(cudaMemsetAsync(d_out + o_J, 0, 8 * NR * NP, ctx->stream));
  SchwarpProb P;
  memset(&P, 0, sizeof(P));
  P.bbs = to_view(&p->bbs);
  P.n = p->n_matches;
  P.kp1 = (const float *)(d_in + o_kp1); P.kp2 = (const float *)(d_in + o_kp2); P.isig = (const float *)(d_in + o_sig);
  P.lambda = p->lambda; P.fx = p->fx; P.fy = p->fy;
  P.x = (double *)(d_in + o_x);
  schwarp_rows_kernel<<<grid_for((int)NR, ctx->sm_count), 256, 0, ctx->stream>>>(
      P, (int)NR, (double *)(d_out + o_r), J ? (double *)(d_out + o_J) : nullptr);
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, outp.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(r, h + o_r, 8 * NR);
  if (J) memcpy(J, h + o_J, 8 * NR * NP);
  return DEFSLAM_OK;
}

/* DefORBmatcher::CalculateInitialSchwarp (DefORBmatcher.cc:111-187): Warp::initialize on the device (the fit
 * kernel with zero trust-region steps), then NaN scrub, data residuals, loss corrector and the > 20 filter in two
 * more launches. */
int defslam_schwarp_initial(const defslam_schwarp_problem *p, uint8_t *keep_out, double *err_out) {
  if (!p || !keep_out || !bbs_ok(&p->bbs, 2) || p->n_matches <= 0 || !p->kp1 || !p->kp2 || !p->inv_sigma || !p->x)
    return DEFSLAM_EBADARG;
  defslam_schwarp_problem q = *p;
  q.initialize = 1;
  q.max_iterations = 0;
  defslam_diffprop dp;
  memset(&dp, 0, sizeof(dp));
  int rc = defslam_schwarp_fit_batched(1, &q, &dp, -1); /* x <- (C'C + lambda B)^-1 C' q2 */
  if (rc) return rc;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  const size_t n = p->n_matches, NC = (size_t)p->bbs.nptsu * p->bbs.nptsv, NP = 2 * NC;
  Packer in, outp;
  const size_t o_kp1 = in.add(8 * n + 8), o_kp2 = in.add(8 * n + 8), o_sig = in.add(4 * n + 8);
  const size_t o_x = outp.add(8 * NP), o_keep = outp.add(n + 8), o_err = outp.add(8 * n), o_r = outp.add(16 * n);
  Scratch &S = tl_scratch(ctx->device);
  if ((rc = S.host.ensure(in.total + outp.total)) || (rc = S.dev.ensure(in.total + outp.total))) return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total, *h_out = h + in.total;
  memcpy(h + o_kp1, p->kp1, 8 * n); memcpy(h + o_kp2, p->kp2, 8 * n); memcpy(h + o_sig, p->inv_sigma, 4 * n);
  memcpy(h_out + o_x, p->x, 8 * NP);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total + o_keep, cudaMemcpyHostToDevice, ctx->stream));
  SchwarpProb P;
  memset(&P, 0, sizeof(P));
  P.bbs = to_view(&p->bbs);
  P.n = p->n_matches;
  P.kp1 = (const float *)(d_in + o_kp1); P.kp2 = (const float *)(d_in + o_kp2); P.isig = (const float *)(d_in + o_sig);
  P.lambda = p->lambda; P.fx = p->fx; P.fy = p->fy;
  P.x = (double *)(d_out + o_x);
  int nscrub = 2 * p->bbs.nptsu * p->bbs.nptsu; /* quirk C9: NCu*NCu*2, not NCu*NCv*2 */
  if (nscrub > (int)NP) nscrub = (int)NP;
  schwarp_scrub_kernel<<<1, 256, 0, ctx->stream>>>(P.x, nscrub);
  schwarp_rows_kernel<<<grid_for((int)(2 * n), ctx->sm_count), 256, 0, ctx->stream>>>(P, (int)(2 * n), (double *)(d_out + o_r), nullptr);
  schwarp_initial_filter_kernel<<<1, 256, 0, ctx->stream>>>((const double *)(d_out + o_r), (int)n, d_out + o_keep,
                                                            (double *)(d_out + o_err));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(3);
  DS_CUDA_TRY(cudaMemcpyAsync(h_out, d_out, o_r, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(p->x, h_out + o_x, 8 * NP);
  memcpy(keep_out, h_out + o_keep, n);
  if (err_out) memcpy(err_out, h_out + o_err, 8 * n);
  return DEFSLAM_OK;
}

/* ------------------------------------------------------------------ normals ---------- */

int defslam_normals_batched(const defslam_normals_problem *p, double *k_out, double *cov_out, float *normal_out,
                            uint8_t *status_out, int32_t *iters_out, float *pair_normal_out,
                            uint8_t *pair_valid_out) {
  if (!p || p->n_points < 0 || (p->n_points > 0 && (!p->pair_ptr || !p->ref_uv)) || p->max_iterations < 0)
    return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  const size_t n = p->n_points;
  if (n == 0) return DEFSLAM_OK;
  const size_t np = (size_t)p->pair_ptr[n];
  for (size_t i = 0; i < n; i++)
    if (p->pair_ptr[i + 1] < p->pair_ptr[i]) return DEFSLAM_EBADARG;
  if (np > 0 && (!p->J12 || !p->J21 || !p->H12 || !p->I1 || !p->I2)) return DEFSLAM_EBADARG;
  Packer in, outp, scr;
  const size_t o_ptr = in.add(4 * (n + 1)), o_j12 = in.add(16 * np + 16), o_j21 = in.add(16 * np + 16),
               o_h12 = in.add(24 * np + 24), o_i1 = in.add(8 * np + 8), o_i2 = in.add(8 * np + 8),
               o_fr = in.add(np + 8), o_kf = in.add(8 * np + 8), o_ki = in.add(16 * n), o_uv = in.add(8 * n);
  const size_t o_k = outp.add(16 * n), o_cov = outp.add(32 * n), o_nrm = outp.add(12 * n), o_st = outp.add(n),
               o_it = outp.add(4 * n), o_pn = outp.add(12 * np + 12), o_pv = outp.add(np + 8);
  const size_t o_Q = scr.add(160 * np + 160);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > outp.total ? in.total : outp.total)) ||
      (rc = S.dev.ensure(in.total + outp.total)) || (rc = S.ws.ensure(scr.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  /* The pinned arena holds the inputs first and is reused for the outputs (in.total vs outp.total: separate halves
   * when the batch is pipelined). */
  NormalsProb P;
  P.n_points = (int)n; P.npairs = (int)np;
  P.pair_ptr = (const int *)(d_in + o_ptr);
  P.J12 = (const float *)(d_in + o_j12); P.J21 = (const float *)(d_in + o_j21); P.H12 = (const float *)(d_in + o_h12);
  P.I1 = (const float *)(d_in + o_i1); P.I2 = (const float *)(d_in + o_i2);
  P.from_ref = p->pair_from_ref ? d_in + o_fr : nullptr;
  P.k_first = p->k_first ? (const float *)(d_in + o_kf) : nullptr;
  P.k_init = p->k_init ? (const double *)(d_in + o_ki) : nullptr;
  P.ref_uv = (const float *)(d_in + o_uv);
  P.max_iterations = p->max_iterations; P.corrected_t2 = p->corrected_t2;
  P.Q = (double *)((uint8_t *)S.ws.p + o_Q);
  P.k_out = (double *)(d_out + o_k); P.cov_out = (double *)(d_out + o_cov); P.normal_out = (float *)(d_out + o_nrm);
  P.status_out = d_out + o_st; P.iters_out = (int *)(d_out + o_it);
  P.pair_normal_out = (float *)(d_out + o_pn); P.pair_valid_out = d_out + o_pv;

  /* Large batches run as a pipeline of chunks of map points (their pairs are contiguous: CSR) over two side streams:
   * while the device works on chunk c (upload, kernel, download), the host packs chunk c+1 into the pinned arena and
   * unpacks chunk c-1 from it, both on several threads.  Serially the call was 13x slower than its kernel (pack 82 MB,
   * upload, 1.5 ms of kernel, download 44 MB, unpack). */
  /* (chunks stay large: a chunk's kernel lasts as long as its slowest point, up to max_iterations trips) */
  size_t nchunk = n >= 65536 ? std::max<size_t>(2, std::min<size_t>(6, n / 98304)) : 1;
  if (getenv("DEFSLAM_NO_PIPELINE") != nullptr) nchunk = 1;
  if (const char *e = getenv("DEFSLAM_NORMALS_CHUNKS")) { /* 1: one pass (the kernel timed alone), 2..16: forced */
    const int v = atoi(e);
    if (v >= 1) nchunk = std::min<size_t>((size_t)std::min(v, 16), std::max<size_t>(1, n / 128));
  }
  uint8_t *h_in = h, *h_out = h;
  if (nchunk > 1) {
    if ((rc = S.host.ensure(in.total + outp.total)) || (rc = S.ensure_side())) return rc;
    h = (uint8_t *)S.host.p; h_in = h; h_out = h + in.total;
  }
  struct Slice { size_t i0, i1, p0, p1; };
  auto slice_of = [&](size_t c) {
    Slice sl;
    sl.i0 = n * c / nchunk; sl.i1 = n * (c + 1) / nchunk;
    sl.p0 = (size_t)p->pair_ptr[sl.i0]; sl.p1 = (size_t)p->pair_ptr[sl.i1];
    return sl;
  };
  /* byte ranges of the arrays of one slice: {arena offset, caller pointer, element bytes, first, count} */
  struct Seg { size_t off; const void *src; size_t eb, first, count; };
  auto in_segs = [&](const Slice &sl, Seg *sg) {
    int k = 0;
    const size_t ni = sl.i1 - sl.i0, nq = sl.p1 - sl.p0;
    sg[k++] = Seg{o_ptr, p->pair_ptr, 4, sl.i0, ni + 1}; /* the kernel reads pair_ptr[i1] as well */
    if (nq) {
      sg[k++] = Seg{o_j12, p->J12, 16, sl.p0, nq}; sg[k++] = Seg{o_j21, p->J21, 16, sl.p0, nq};
      sg[k++] = Seg{o_h12, p->H12, 24, sl.p0, nq};
      sg[k++] = Seg{o_i1, p->I1, 8, sl.p0, nq}; sg[k++] = Seg{o_i2, p->I2, 8, sl.p0, nq};
      if (p->pair_from_ref) sg[k++] = Seg{o_fr, p->pair_from_ref, 1, sl.p0, nq};
      if (p->k_first) sg[k++] = Seg{o_kf, p->k_first, 8, sl.p0, nq};
    }
    if (p->k_init) sg[k++] = Seg{o_ki, p->k_init, 16, sl.i0, ni};
    sg[k++] = Seg{o_uv, p->ref_uv, 8, sl.i0, ni};
    return k;
  };
  auto out_segs = [&](const Slice &sl, Seg *sg) {
    int k = 0;
    const size_t ni = sl.i1 - sl.i0, nq = sl.p1 - sl.p0;
    sg[k++] = Seg{o_k, nullptr, 16, sl.i0, ni}; sg[k++] = Seg{o_cov, nullptr, 32, sl.i0, ni};
    sg[k++] = Seg{o_nrm, nullptr, 12, sl.i0, ni}; sg[k++] = Seg{o_st, nullptr, 1, sl.i0, ni};
    sg[k++] = Seg{o_it, nullptr, 4, sl.i0, ni};
    if (nq) { sg[k++] = Seg{o_pn, nullptr, 12, sl.p0, nq}; sg[k++] = Seg{o_pv, nullptr, 1, sl.p0, nq}; }
    return k;
  };
  auto pack = [&](const Slice &sl) {
    Seg sg[12];
    const int k = in_segs(sl, sg);
    size_t bytes = 0;
    for (int a = 0; a < k; a++) bytes += sg[a].eb * sg[a].count;
    /* every thread copies its share of every array */
    host_parallel_for(bytes, (size_t)2 << 20, [&](size_t b, size_t e) {
      const double f0 = (double)b / (double)bytes, f1 = (double)e / (double)bytes;
      for (int a = 0; a < k; a++) {
        const size_t len = sg[a].eb * sg[a].count;
        const size_t lo = e == bytes && b == 0 ? 0 : (size_t)(f0 * (double)len) & ~(size_t)15;
        const size_t hi = e == bytes ? len : (size_t)(f1 * (double)len) & ~(size_t)15;
        if (hi > lo) memcpy(h_in + sg[a].off + sg[a].eb * sg[a].first + lo, (const uint8_t *)sg[a].src + sg[a].eb * sg[a].first + lo, hi - lo);
      }
    });
  };
  auto unpack = [&](const Slice &sl) {
    const uint8_t *st = h_out + o_st, *pv = h_out + o_pv;
    const size_t ni = sl.i1 - sl.i0, nq = sl.p1 - sl.p0;
    host_parallel_for(ni + nq, 65536, [&](size_t b, size_t e) {
      /* points [i0 + b', ...) and pairs, split proportionally */
      const size_t ib = sl.i0 + (size_t)((double)b / (double)(ni + nq) * (double)ni);
      const size_t ie = e == ni + nq ? sl.i1 : sl.i0 + (size_t)((double)e / (double)(ni + nq) * (double)ni);
      const size_t qb = sl.p0 + (size_t)((double)b / (double)(ni + nq) * (double)nq);
      const size_t qe = e == ni + nq ? sl.p1 : sl.p0 + (size_t)((double)e / (double)(ni + nq) * (double)nq);
      if (k_out) memcpy(k_out + 2 * ib, h_out + o_k + 16 * ib, 16 * (ie - ib));
      if (status_out) memcpy(status_out + ib, st + ib, ie - ib);
      if (iters_out) memcpy(iters_out + ib, h_out + o_it + 4 * ib, 4 * (ie - ib));
      /* covariance / normal only where estimated (the reference leaves the rest untouched) */
      for (size_t i = ib; i < ie; i++) {
        if (st[i] != 1) continue;
        if (cov_out) memcpy(cov_out + 4 * i, h_out + o_cov + 32 * i, 32);
        if (normal_out) memcpy(normal_out + 3 * i, h_out + o_nrm + 12 * i, 12);
      }
      if (pair_valid_out && qe > qb) memcpy(pair_valid_out + qb, pv + qb, qe - qb);
      if (pair_normal_out)
        for (size_t j = qb; j < qe; j++)
          if (pv[j]) memcpy(pair_normal_out + 3 * j, h_out + o_pn + 12 * j, 12);
    });
  };
  auto enqueue = [&](const Slice &sl, cudaStream_t st, cudaEvent_t k0, cudaEvent_t k1) -> int {
    Seg sg[12];
    int k = in_segs(sl, sg);
    for (int a = 0; a < k; a++) {
      const size_t off = sg[a].off + sg[a].eb * sg[a].first, len = sg[a].eb * sg[a].count;
      if (len) DS_CUDA_TRY(cudaMemcpyAsync(d_in + off, h_in + off, len, cudaMemcpyHostToDevice, st));
    }
    DS_CUDA_TRY(cudaEventRecord(k0, st));
    const size_t ni = sl.i1 - sl.i0;
    if (ni) {
      const int g = (int)((ni + 127) / 128);
      normals_kernel<<<g, 128, 0, st>>>(P, (int)sl.i0, (int)sl.i1);
      DS_CUDA_TRY(cudaGetLastError());
      g_launches.fetch_add(1);
    }
    DS_CUDA_TRY(cudaEventRecord(k1, st));
    k = out_segs(sl, sg);
    for (int a = 0; a < k; a++) {
      const size_t off = sg[a].off + sg[a].eb * sg[a].first, len = sg[a].eb * sg[a].count;
      if (len) DS_CUDA_TRY(cudaMemcpyAsync(h_out + off, d_out + off, len, cudaMemcpyDeviceToHost, st));
    }
    return DEFSLAM_OK;
  };
  float ms_total = 0.f;
  if (nchunk == 1) {
    const Slice sl = slice_of(0);
    pack(sl);
    if ((rc = enqueue(sl, ctx->stream, ctx->e0, ctx->e1))) return rc;
    DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ms_total, ctx->e0, ctx->e1);
    unpack(sl);
  } else {
    for (size_t c = 0; c < nchunk; c++) {
      const Slice sl = slice_of(c);
      pack(sl);
      if ((rc = enqueue(sl, S.side[c & 1], S.ev_k0[c], S.ev_k1[c]))) return rc;
      DS_CUDA_TRY(cudaEventRecord(S.ev_done[c], S.side[c & 1]));
      if (c >= 1) {
        DS_CUDA_TRY(cudaEventSynchronize(S.ev_done[c - 1]));
        unpack(slice_of(c - 1));
      }
    }
    DS_CUDA_TRY(cudaEventSynchronize(S.ev_done[nchunk - 1]));
    unpack(slice_of(nchunk - 1));
    for (size_t c = 0; c < nchunk; c++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, S.ev_k0[c], S.ev_k1[c]);
      ms_total += ms;
    }
  }
  g_last_kernel_ms = ms_total;
  return DEFSLAM_OK;
}

int defslam_polysolver_coefficients(int32_t npairs, const float *J12, const float *H12, const float *I1,
                                    const float *I2, double *eq1, double *eq2) {
  if (npairs < 0 || (npairs > 0 && (!J12 || !H12 || !I1 || !I2 || !eq1 || !eq2))) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  if (npairs == 0) return DEFSLAM_OK;
  const size_t np = npairs;
  Packer in, outp;
  const size_t o_j = in.add(16 * np), o_h = in.add(24 * np), o_1 = in.add(8 * np), o_2 = in.add(8 * np);
  const size_t o_e1 = outp.add(80 * np), o_e2 = outp.add(80 * np);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > outp.total ? in.total : outp.total)) || (rc = S.dev.ensure(in.total + outp.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  memcpy(h + o_j, J12, 16 * np); memcpy(h + o_h, H12, 24 * np); memcpy(h + o_1, I1, 8 * np); memcpy(h + o_2, I2, 8 * np);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  poly_kernel<<<grid_for(npairs, ctx->sm_count), 256, 0, ctx->stream>>>(
      npairs, (const float *)(d_in + o_j), (const float *)(d_in + o_h), (const float *)(d_in + o_1),
      (const float *)(d_in + o_2), (double *)(d_out + o_e1), (double *)(d_out + o_e2));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, outp.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(eq1, h + o_e1, 80 * np);
  memcpy(eq2, h + o_e2, 80 * np);
  return DEFSLAM_OK;
}

/* ------------------------------------------------------------------ shape from normals */

static int sfn_check(const defslam_sfn_problem *p) {
  if (!bbs_ok(&p->bbs, 1) || p->n_normals < 0 || p->n_eval < 0 || (p->n_normals > 0 && (!p->uv || !p->normals)) ||
      (p->n_eval > 0 && (!p->eval_uv || !p->xyz_out)))
    return DEFSLAM_EBADARG;
  return 0;
}

int defslam_sfn_solve_batched(int32_t nprob, const defslam_sfn_problem *p, int32_t *rc_out, int32_t device) {
  if (nprob < 0 || (nprob > 0 && !p)) return DEFSLAM_EBADARG;
  int nu = 0, nv = 0, nmax = 0;
  for (int i = 0; i < nprob; i++) {
    if (sfn_check(&p[i])) return DEFSLAM_EBADARG;
    if (p[i].bbs.nptsu > nu) nu = p[i].bbs.nptsu;
    if (p[i].bbs.nptsv > nv) nv = p[i].bbs.nptsv;
    if (p[i].n_normals > nmax) nmax = p[i].n_normals;
  }
  DeviceGuard device_guard_;
  DevCtx *ctx = get_ctx(device);
  if (!ctx) return DEFSLAM_ECUDA;
  if (nprob == 0) return DEFSLAM_OK;
  const size_t NCmax = (size_t)nu * nv;
  const size_t fixed = sizeof(double) * (size_t)sfn_smem_fixed((int)NCmax);
  const size_t packed = sizeof(double) * NCmax * (NCmax + 1) / 2;
  if (fixed > (size_t)ctx->smem_optin) return DEFSLAM_ETOOLARGE;
  /* N: as tiles in shared memory for the tensor-core factorisation (2), else packed in shared memory (1), else packed
   * in the workspace (0); DEFSLAM_SFN_MODE caps it (A/B runs) */
  const size_t tiled = sizeof(double) * (size_t)sfn_tile_doubles((int)NCmax) + 16;
  int n_in_smem = fixed + tiled <= (size_t)ctx->smem_optin ? 2 : (fixed + packed <= (size_t)ctx->smem_optin ? 1 : 0);
  if (const char *e = getenv("DEFSLAM_SFN_MODE")) { const int v = atoi(e); if (v >= 0 && v < n_in_smem && (v != 1 || fixed + packed <= (size_t)ctx->smem_optin)) n_in_smem = v; }
  const size_t smem = fixed + (n_in_smem == 2 ? tiled : (n_in_smem ? packed : 0));
  Packer in, outp;
  std::vector<size_t> o_uv(nprob), o_nr(nprob), o_ev(nprob), o_ct(nprob), o_xyz(nprob), o_rc(nprob);
  const size_t o_probs = in.add(sizeof(SfnProb) * nprob), o_counter = in.add(sizeof(int));
  for (int i = 0; i < nprob; i++) {
    const size_t n = p[i].n_normals, NC = (size_t)p[i].bbs.nptsu * p[i].bbs.nptsv, ne = p[i].n_eval;
    o_uv[i] = in.add(8 * n + 8); o_nr[i] = in.add(12 * n + 12); o_ev[i] = in.add(8 * ne + 8);
    o_ct[i] = outp.add(8 * NC); o_xyz[i] = outp.add(12 * ne + 12); o_rc[i] = outp.add(8);
  }
  const int grid = persistent_grid(nprob, ctx->sm_count);
  const size_t ws_stride = sfn_ws_sizes(nu, nv, nmax).total;
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > outp.total ? in.total : outp.total)) ||
      (rc = S.dev.ensure(in.total + outp.total)) || (rc = S.ws.ensure(ws_stride * grid)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  SfnProb *hp = (SfnProb *)(h + o_probs);
  for (int i = 0; i < nprob; i++) {
    const size_t n = p[i].n_normals, ne = p[i].n_eval;
    if (n) { memcpy(h + o_uv[i], p[i].uv, 8 * n); memcpy(h + o_nr[i], p[i].normals, 12 * n); }
    if (ne) memcpy(h + o_ev[i], p[i].eval_uv, 8 * ne);
    SfnProb &P = hp[i];
    P.bbs = to_view(&p[i].bbs);
    P.n = p[i].n_normals; P.n_eval = p[i].n_eval;
    P.uv = (const float *)(d_in + o_uv[i]); P.normals = (const float *)(d_in + o_nr[i]);
    P.eval_uv = (const float *)(d_in + o_ev[i]);
    P.bending = p[i].bending; P.mean_depth = p[i].mean_depth;
    P.ctrl_out = (double *)(d_out + o_ct[i]); P.xyz_out = (float *)(d_out + o_xyz[i]); P.rc_out = (int *)(d_out + o_rc[i]);
  }
  *(int *)(h + o_counter) = 0;
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  DS_CUDA_TRY(cudaMemsetAsync(d_out, 0xff, outp.total, ctx->stream));
  DS_CUDA_TRY(raise_dynamic_smem((const void *)sfn_solve_kernel, ctx->device, (int)smem));
  DS_CUDA_TRY(cudaEventRecord(ctx->e0, ctx->stream));
  sfn_solve_kernel<<<grid, NRSFM_THREADS, smem, ctx->stream>>>((const SfnProb *)(d_in + o_probs), nprob,
                                                               (uint8_t *)S.ws.p, ws_stride, nu, nv, nmax, n_in_smem,
                                                               (int *)(d_in + o_counter));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaEventRecord(ctx->e1, ctx->stream));
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, outp.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->e0, ctx->e1);
  g_last_kernel_ms = ms;
  int worst = DEFSLAM_OK;
  for (int i = 0; i < nprob; i++) {
    const size_t NC = (size_t)p[i].bbs.nptsu * p[i].bbs.nptsv, ne = p[i].n_eval;
    int r = *(const int *)(h + o_rc[i]);
    if (r != 0 && r != DEFSLAM_EBADARG && r != DEFSLAM_ENUMERIC) r = DEFSLAM_ECUDA;
    if (rc_out) rc_out[i] = r;
    if (r) { if (worst == DEFSLAM_OK) worst = r; continue; }
    if (p[i].ctrl_out) memcpy(p[i].ctrl_out, h + o_ct[i], 8 * NC);
    if (ne) memcpy(p[i].xyz_out, h + o_xyz[i], 12 * ne);
  }
  return worst;
}

int defslam_sfn_solve(const defslam_sfn_problem *p) {
  if (!p) return DEFSLAM_EBADARG;
  return defslam_sfn_solve_batched(1, p, nullptr, -1);
}

int defslam_sfn_system(const defslam_sfn_problem *p, double *A, double *b) {
  if (!p || !A || !b || sfn_check(p)) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  const size_t n = p->n_normals, NC = (size_t)p->bbs.nptsu * p->bbs.nptsv, rows = 2 * n + NC + 1;
  Packer in, outp;
  const size_t o_uv = in.add(8 * n + 8), o_nr = in.add(12 * n + 12);
  const size_t o_A = outp.add(8 * rows * NC), o_b = outp.add(8 * rows);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > outp.total ? in.total : outp.total)) || (rc = S.dev.ensure(in.total + outp.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  if (n) { memcpy(h + o_uv, p->uv, 8 * n); memcpy(h + o_nr, p->normals, 12 * n); }
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  SfnProb P;
  memset(&P, 0, sizeof(P));
  P.bbs = to_view(&p->bbs);
  P.n = p->n_normals;
  P.uv = (const float *)(d_in + o_uv); P.normals = (const float *)(d_in + o_nr);
  P.bending = p->bending; P.mean_depth = p->mean_depth;
  sfn_rows_kernel<<<grid_for((int)rows, ctx->sm_count), 256, 0, ctx->stream>>>(P, (int)rows, (double *)(d_out + o_A),
                                                                                (double *)(d_out + o_b));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, outp.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(A, h + o_A, 8 * rows * NC);
  memcpy(b, h + o_b, 8 * rows);
  return DEFSLAM_OK;
}

/* ------------------------------------------------------------------ Sim(3) registration */

int defslam_sim3_register_batched(int32_t nprob, const defslam_sim3_problem *p, defslam_sim3_result *out,
                                  int32_t device) {
  if (nprob < 0 || (nprob > 0 && (!p || !out))) return DEFSLAM_EBADARG;
  for (int i = 0; i < nprob; i++)
    if (p[i].n_points <= 0 || !p[i].pts1 || !p[i].pts2 || p[i].max_iterations < 0 || !(p[i].huber > 0)) return DEFSLAM_EBADARG;
  DeviceGuard device_guard_;
  DevCtx *ctx = get_ctx(device);
  if (!ctx) return DEFSLAM_ECUDA;
  if (nprob == 0) return DEFSLAM_OK;
  Packer in, outp;
  std::vector<size_t> o1(nprob), o2(nprob), oo(nprob);
  const size_t o_probs = in.add(sizeof(Sim3Prob) * nprob);
  for (int i = 0; i < nprob; i++) {
    const size_t n = p[i].n_points;
    o1[i] = in.add(12 * n); o2[i] = in.add(12 * n); oo[i] = outp.add(16 * 8);
  }
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > outp.total ? in.total : outp.total)) || (rc = S.dev.ensure(in.total + outp.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  Sim3Prob *hp = (Sim3Prob *)(h + o_probs);
  for (int i = 0; i < nprob; i++) {
    const size_t n = p[i].n_points;
    memcpy(h + o1[i], p[i].pts1, 12 * n);
    memcpy(h + o2[i], p[i].pts2, 12 * n);
    Sim3Prob &P = hp[i];
    P.n = p[i].n_points;
    P.p1 = (const float *)(d_in + o1[i]); P.p2 = (const float *)(d_in + o2[i]);
    for (int k = 0; k < 4; k++) P.init.q[k] = p[i].rot[k];
    for (int k = 0; k < 3; k++) P.init.t[k] = p[i].trans[k];
    P.init.s = p[i].scale;
    P.chi = p[i].chi; P.huber = p[i].huber; P.max_iterations = p[i].max_iterations;
    P.out = (double *)(d_out + oo[i]);
  }
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  DS_CUDA_TRY(cudaEventRecord(ctx->e0, ctx->stream));
  const int grid = nprob < ctx->sm_count * 4 ? nprob : ctx->sm_count * 4;
  sim3_register_kernel<<<grid, SIM3_THREADS, 0, ctx->stream>>>((const Sim3Prob *)(d_in + o_probs), nprob);
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaEventRecord(ctx->e1, ctx->stream));
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, outp.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->e0, ctx->e1);
  g_last_kernel_ms = ms;
  for (int i = 0; i < nprob; i++) {
    const double *o = (const double *)(h + oo[i]);
    for (int k = 0; k < 4; k++) out[i].rot[k] = o[k];
    for (int k = 0; k < 3; k++) out[i].trans[k] = o[4 + k];
    out[i].scale = o[7]; out[i].chi2 = o[8];
    out[i].inliers = (int)o[9]; out[i].acceptable = (int)o[10];
    out[i].iterations[0] = (int)o[11]; out[i].iterations[1] = (int)o[12];
  }
  return DEFSLAM_OK;
}

int defslam_scale_min_median(int32_t n, const float *mono_xyz, const float *stereo_xyz, uint64_t seed,
                             float *scale_out) {
  if (n <= 0 || !mono_xyz || !stereo_xyz || !scale_out) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  if ((size_t)n * 4 > (size_t)ctx->smem_optin - 1024) return DEFSLAM_ETOOLARGE;
  Packer in, outp;
  const size_t om = in.add(12 * (size_t)n), os = in.add(12 * (size_t)n), oo = outp.add(16);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > outp.total ? in.total : outp.total)) || (rc = S.dev.ensure(in.total + outp.total)))
    return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d_in = (uint8_t *)S.dev.p, *d_out = d_in + in.total;
  memcpy(h + om, mono_xyz, 12 * (size_t)n);
  memcpy(h + os, stereo_xyz, 12 * (size_t)n);
  DS_CUDA_TRY(cudaMemcpyAsync(d_in, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  const size_t smem = 4 * (size_t)n;
  if (smem > 48 * 1024)
    DS_CUDA_TRY(raise_dynamic_smem((const void *)scale_min_median_kernel, ctx->device, (int)smem));
  scale_min_median_kernel<<<1, 512, smem, ctx->stream>>>(n, (const float *)(d_in + om), (const float *)(d_in + os), seed,
                                                         (float *)(d_out + oo));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaMemcpyAsync(h, d_out, outp.total, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *scale_out = *(const float *)(h + oo);
  return DEFSLAM_OK;
}

}
