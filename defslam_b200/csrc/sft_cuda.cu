/*
 * sft_cuda.cu -- CUDA kernel wrapper + C ABI of the Shape-from-Template solve.
 *
 * One persistent CTA per frame runs the whole Levenberg-Marquardt solve
 * (sft_core.h); a batch of frames is one launch.  Entry points replace
 * defSLAM::Optimizer::DefPoseOptimization (Modules/Tracking/DefOptimizer.cc:251-578)
 * and everything below it; see include/defslam_b200.h.
 */
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ds_batch.h"
#include "ds_runtime.h"

using namespace ds;

namespace {

constexpr int SFT_THREADS = 256;

/* XS: node positions and LM step in shared memory (every mesh up to ~21 x 21); the other variant
 * keeps them in the CTA's global workspace for meshes whose factorisation window fills the SM. */
template <bool XS>
#ifndef DS_MIN_CTAS
#define DS_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(SFT_THREADS, DS_MIN_CTAS)
sft_lm_kernel(const ProbView *__restrict__ probs, int nprob, uint8_t *ws_base, size_t ws_stride, WorkspaceSizes z,
              long long *prof, int *work_counter) {
  extern __shared__ __align__(16) double smem[];
  Team team;
  team.tid = threadIdx.x;
  team.nthr = blockDim.x;
  uint8_t *ws = ws_base + (size_t)blockIdx.x * ws_stride;
  /* frames are handed out dynamically: LM trial counts differ per frame */
  __shared__ int next_problem;
  bool first = true;
  int pi = blockIdx.x;
  while (pi < nprob) {
    sft_run_problem<XS>(team, probs[pi], smem, ws, z, first, blockIdx.x == 0 ? prof : nullptr);
    first = false;
    __syncthreads();
    if (threadIdx.x == 0) next_problem = (int)gridDim.x + atomicAdd(work_counter, 1);
    __syncthreads();
    pi = next_problem;
  }
}

}  // namespace

/* ------------------------------------------------------------ templates -- */

struct defslam_template {
  PlanHost host;
  PlanView hview;       /* host-addressable */
  PlanView dview_host;  /* device pointers, host copy */
  int device = -1;
  double *d_dbl = nullptr;
  int *d_i32 = nullptr;
  uint8_t *d_u8 = nullptr;
  PlanView *d_view = nullptr;
};

static void template_free(defslam_template *t) {
  if (!t) return;
  if (t->d_dbl || t->d_i32 || t->d_u8 || t->d_view) {
    int cur = 0;
    cudaGetDevice(&cur);
    if (t->device >= 0) cudaSetDevice(t->device);
    cudaFree(t->d_dbl); cudaFree(t->d_i32); cudaFree(t->d_u8); cudaFree(t->d_view);
    cudaSetDevice(cur);
  }
  delete t;
}

static int template_make(const defslam_template_desc *desc, DevCtx *ctx, defslam_template **out) {
  std::unique_ptr<defslam_template> t(new defslam_template);
  const int rc = t->host.build(desc);
  if (rc) return rc;
  t->hview = t->host.host_view();
  t->device = ctx->device;
  defslam_template *raw = t.release();
  auto fail = [&](int code) { template_free(raw); return code; };
  const size_t nd = raw->host.dbl.size() * sizeof(double), ni = raw->host.i32.size() * sizeof(int),
               nu = raw->host.u8.size();
  if (cudaMalloc(&raw->d_dbl, nd + 8) != cudaSuccess || cudaMalloc(&raw->d_i32, ni + 8) != cudaSuccess ||
      cudaMalloc(&raw->d_u8, nu + 8) != cudaSuccess || cudaMalloc(&raw->d_view, sizeof(PlanView)) != cudaSuccess) {
    cudaGetLastError();
    return fail(DEFSLAM_ECUDA);
  }
  raw->dview_host = raw->host.bind(raw->d_dbl, raw->d_i32, raw->d_u8);
  if (cudaMemcpyAsync(raw->d_dbl, raw->host.dbl.data(), nd, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(raw->d_i32, raw->host.i32.data(), ni, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(raw->d_u8, raw->host.u8.data(), nu, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(raw->d_view, &raw->dview_host, sizeof(PlanView), cudaMemcpyHostToDevice, ctx->stream) !=
          cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    cudaGetLastError();
    return fail(DEFSLAM_ECUDA);
  }
  *out = raw;
  return DEFSLAM_OK;
}

/* --------------------------------------------------------------- batches -- */

struct defslam_sft_batch {
  DevCtx *ctx = nullptr;
  BatchMarshal bm;
  DevBuf h_in, h_out, d_in, d_out, d_views, d_ws, d_prof, d_counter;
  std::vector<defslam_template *> temps;
  int nprob = 0, grid = 0, smem_bytes = 0, mode = MODE_SOLVE;
  size_t ws_stride = 0;
  float last_ms = 0.f;
  cudaEvent_t k0 = nullptr, k1 = nullptr, done = nullptr, up = nullptr; /* pipelined host path: per-chunk events */
  int ensure_events() {
    if (k0) return 0;
    if (cudaEventCreate(&k0) != cudaSuccess || cudaEventCreate(&k1) != cudaSuccess ||
        cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&up, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      return DEFSLAM_ECUDA;
    }
    return 0;
  }
  defslam_sft_batch() { h_in.pinned = true; h_out.pinned = true; }
  void drop_temps() {
    for (auto *t : temps) template_free(t);
    temps.clear();
  }
  ~defslam_sft_batch() {
    drop_temps();
    h_in.release(); h_out.release(); d_in.release(); d_out.release(); d_views.release(); d_ws.release();
    d_prof.release(); d_counter.release();
  }
};

/* marshal + upload; after this the batch is resident on the device */
static int batch_load(defslam_sft_batch *B, int nprob, const defslam_sft_problem *p, int mode,
                      cudaStream_t copy_stream = nullptr) {
  DevCtx *ctx = B->ctx;
  if (!copy_stream) copy_stream = ctx->stream;
  B->drop_temps();
  B->nprob = nprob;
  B->mode = mode;
  std::map<const defslam_template_desc *, defslam_template *> by_desc;
  auto resolve = [&](const defslam_sft_problem &q, const PlanView **hv, const PlanView **dv) -> int {
    const defslam_template *t = q.tmpl;
    if (!t) {
      if (!q.tmpl_desc) return DEFSLAM_EBADARG;
      auto it = by_desc.find(q.tmpl_desc);
      if (it == by_desc.end()) {
        defslam_template *nt = nullptr;
        const int rc = template_make(q.tmpl_desc, ctx, &nt);
        if (rc) return rc;
        B->temps.push_back(nt);
        it = by_desc.emplace(q.tmpl_desc, nt).first;
      }
      t = it->second;
    } else if (t->device != ctx->device) {
      return DEFSLAM_EBADARG;
    }
    *hv = &t->hview;
    *dv = t->d_view;
    return 0;
  };
  int smem_limit = (ctx->smem_optin - 1024) / (int)sizeof(double);
  if (const char *e = getenv("DEFSLAM_SMEM_LIMIT")) { /* experiments: force the planner's fallback placements */
    const int lim = atoi(e) / (int)sizeof(double);
    if (lim > 0 && lim < smem_limit) smem_limit = lim;
  }
  if (const char *e = getenv("DEFSLAM_ROW_MODE")) B->bm.row_mode = atoi(e); /* 0: sliding-window factorisation only */
  int rc = B->bm.plan(nprob, p, mode, smem_limit, resolve);
  if (rc) return rc;
  if ((rc = B->h_in.ensure(B->bm.in_bytes)) || (rc = B->h_out.ensure(B->bm.out_bytes)) ||
      (rc = B->d_in.ensure(B->bm.in_bytes)) || (rc = B->d_out.ensure(B->bm.out_bytes)) ||
      (rc = B->d_views.ensure(sizeof(ProbView) * (size_t)nprob)))
    return rc;
  B->bm.pack_inputs(p, (uint8_t *)B->h_in.p);
  B->bm.bind((uint8_t *)B->d_in.p, (uint8_t *)B->d_out.p);

  B->smem_bytes = B->bm.smem_doubles * (int)sizeof(double);
  const bool xs = !B->bm.any_x_global;
  int occ = 0;
  if (xs) {
    DS_CUDA_TRY(raise_dynamic_smem((const void *)sft_lm_kernel<true>, ctx->device, B->smem_bytes));
    DS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sft_lm_kernel<true>, SFT_THREADS, B->smem_bytes));
  } else {
    DS_CUDA_TRY(raise_dynamic_smem((const void *)sft_lm_kernel<false>, ctx->device, B->smem_bytes));
    DS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sft_lm_kernel<false>, SFT_THREADS, B->smem_bytes));
  }
  if (occ < 1) return DEFSLAM_ETOOLARGE;
  B->grid = ctx->sm_count * occ;
  if (B->grid > nprob) B->grid = nprob;
  const WorkspaceSizes z = B->bm.ws_sizes();
  B->ws_stride = workspace_bytes(z);
  if ((rc = B->d_ws.ensure(B->ws_stride * (size_t)B->grid))) return rc;

  DS_CUDA_TRY(cudaMemcpyAsync(B->d_in.p, B->h_in.p, B->bm.in_bytes, cudaMemcpyHostToDevice, copy_stream));
  DS_CUDA_TRY(cudaMemcpyAsync(B->d_views.p, B->bm.views.data(), sizeof(ProbView) * (size_t)nprob,
                              cudaMemcpyHostToDevice, copy_stream));
  return 0;
}

static int batch_launch(defslam_sft_batch *B, cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr,
                        cudaStream_t st = nullptr) {
  DevCtx *ctx = B->ctx;
  if (!ev0) { ev0 = ctx->e0; ev1 = ctx->e1; }
  if (!st) st = ctx->stream;
  const WorkspaceSizes z = B->bm.ws_sizes();
  long long *prof = nullptr;
  if (getenv("DEFSLAM_PROFILE")) { /* diagnostics: per-phase cycles of CTA 0 */
    int rc = B->d_prof.ensure(sizeof(long long) * PF_TOTAL);
    if (rc) return rc;
    prof = (long long *)B->d_prof.p;
    DS_CUDA_TRY(cudaMemsetAsync(prof, 0, sizeof(long long) * PF_TOTAL, st));
  }
  {
    int rc = B->d_counter.ensure(sizeof(int));
    if (rc) return rc;
    DS_CUDA_TRY(cudaMemsetAsync(B->d_counter.p, 0, sizeof(int), st));
  }
  DS_CUDA_TRY(cudaEventRecord(ev0, st));
  if (!B->bm.any_x_global)
    sft_lm_kernel<true><<<B->grid, SFT_THREADS, B->smem_bytes, st>>>(
        (const ProbView *)B->d_views.p, B->nprob, (uint8_t *)B->d_ws.p, B->ws_stride, z, prof, (int *)B->d_counter.p);
  else
    sft_lm_kernel<false><<<B->grid, SFT_THREADS, B->smem_bytes, st>>>(
        (const ProbView *)B->d_views.p, B->nprob, (uint8_t *)B->d_ws.p, B->ws_stride, z, prof, (int *)B->d_counter.p);
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  DS_CUDA_TRY(cudaEventRecord(ev1, st));
  return 0;
}

static int batch_wait_kernel(defslam_sft_batch *B) {
  DevCtx *ctx = B->ctx;
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  DS_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->e0, ctx->e1));
  B->last_ms = ms;
  g_last_kernel_ms = ms;
  if (getenv("DEFSLAM_PROFILE") && B->d_prof.p) {
    long long h[PF_TOTAL];
    DS_CUDA_TRY(cudaMemcpy(h, B->d_prof.p, sizeof(h), cudaMemcpyDeviceToHost));
    static const char *names[PF_COUNT] = {"prologue", "eval_store", "build", "fs_init", "S1", "S1_wait", "S2", "S3",
                                          "schur", "bwd_init", "bwd", "update", "eval_trial", "lm_scalar", "finalize"};
    long long tot = 0;
    for (int i = 0; i < PF_COUNT; i++) tot += h[i];
    fprintf(stderr, "[defslam profile] CTA0 cycles total %lld:", tot);
    for (int i = 0; i < PF_COUNT; i++) fprintf(stderr, " %s=%.1f%%", names[i], 100.0 * (double)h[i] / (double)(tot ? tot : 1));
    fprintf(stderr, "\n");
    const double steps = (double)(h[PF_X_STEPS] ? h[PF_X_STEPS] : 1);
    fprintf(stderr, "[defslam profile] per step (%lld steps): S3 phase %.0f, look-ahead factor %.0f; S3 busy per warp:",
            h[PF_X_STEPS], (double)h[PF_S3] / steps, (double)h[PF_X_DIAG] / steps);
    for (int w = 0; w < 16; w++) if (h[PF_X_WARP + w]) fprintf(stderr, " %.0f", (double)h[PF_X_WARP + w] / steps);
    fprintf(stderr, "; S2 phase %.0f busy per warp:", (double)h[PF_S2] / steps);
    for (int w = 0; w < 16; w++) if (h[PF_X_S2W + w]) fprintf(stderr, " %.0f", (double)h[PF_X_S2W + w] / steps);
    fprintf(stderr, "\n");
    if (h[PF_X_STEPS] == 0 && h[PF_X_WARP + 1]) { /* row-owner factorisation: chain warp and owners, per block row */
      fprintf(stderr, "[defslam profile] row owner path, cycles summed over the launch: chain warp waiting %lld, factoring %lld; "
                      "owners (reading their band of H / the block-row steps incl. waits):", h[PF_X_WARP], h[PF_X_WARP + 1]);
      for (int w = 0; w < 5; w++) fprintf(stderr, " %lld/%lld", h[PF_X_WARP + 2 + 2 * w], h[PF_X_WARP + 3 + 2 * w]);
      fprintf(stderr, "\n");
    }
    const double nb = (double)(h[PF_X_BUILDS] ? h[PF_X_BUILDS] : 1);
    fprintf(stderr, "[defslam profile] per build (%lld builds): phase %.0f; facet sums %.0f, camera reduce %.0f, block gather %.0f, "
            "per-node %.0f, max-diag reduce %.0f (thread 0)\n", h[PF_X_BUILDS], (double)h[PF_BUILD] / nb,
            (double)h[PF_X_BUILD] / nb, (double)h[PF_X_BUILD + 1] / nb, (double)h[PF_X_BUILD + 2] / nb,
            (double)h[PF_X_BUILD + 3] / nb, (double)h[PF_X_BUILD + 4] / nb);
    fprintf(stderr, "[defslam profile] facet sums per build (thread 0): chunk setup+barrier %.0f, staging copy %.0f, barrier %.0f, sums %.0f\n",
            (double)h[PF_X_FS] / nb, (double)h[PF_X_FS + 1] / nb, (double)h[PF_X_FS + 2] / nb, (double)h[PF_X_FS + 3] / nb);
    fprintf(stderr, "[defslam profile] backward sweep (thread 0; sliding window: per step, row owners: summed): phase %.0f; sliding window: TMA wait / row owners: copy issue %.0f, block solve / row arithmetic %.0f, update / late copy %.0f, barrier %.0f\n",
            (double)h[PF_BWD] / steps, (double)h[PF_X_BWD] / steps, (double)h[PF_X_BWD + 1] / steps,
            (double)h[PF_X_BWD + 2] / steps, (double)h[PF_X_BWD + 3] / steps);
  }
  return 0;
}

static int batch_download(defslam_sft_batch *B) {
  DevCtx *ctx = B->ctx;
  DS_CUDA_TRY(cudaMemcpyAsync(B->h_out.p, B->d_out.p, B->bm.out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static defslam_sft_batch *tl_batch(DevCtx *ctx, int which = 0) {
  /* raw pointers, never deleted: a thread_local destructor would call cudaFree / cudaFreeHost at thread or
   * process exit, possibly after the CUDA runtime shut down (same reason as DevCtx::~DevCtx) */
  static thread_local std::map<int, defslam_sft_batch *> tl;
  defslam_sft_batch *&b = tl[ctx->device * 4 + which];
  if (!b) { b = new defslam_sft_batch; b->ctx = ctx; }
  return b;
}

/* Large host batches are cut into chunks that ping-pong between two batch objects: while the
 * kernel of chunk c runs on the library's stream, the host marshals chunk c+1 into the other pinned
 * arena, a copy stream uploads it and brings the results of chunk c-1 back, and the host scatters
 * them -- only the first marshal+upload and the last download+scatter are exposed. */
static cudaStream_t tl_copy_stream(DevCtx *ctx, int which) {
  static thread_local std::map<int, cudaStream_t> tl;
  const int key = ctx->device * 4 + which;
  auto it = tl.find(key);
  if (it != tl.end()) return it->second;
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  tl[key] = s;
  return s;
}
static int solve_pipelined(DevCtx *ctx, int nprob, const defslam_sft_problem *p, defslam_sft_result *r, int chunk) {
  defslam_sft_batch *Bs[2] = {tl_batch(ctx, 1), tl_batch(ctx, 2)};
  cudaStream_t cs = tl_copy_stream(ctx, 0), cs_down = tl_copy_stream(ctx, 1); /* uploads / downloads */
  /* kernels of consecutive chunks go to two compute streams: the next chunk's CTAs move onto the SMs
   * the current chunk's last frames leave idle (frames differ in LM trial count) */
  cudaStream_t comp[2] = {ctx->stream, tl_copy_stream(ctx, 2)};
  if (!cs || !cs_down || !comp[1]) return DEFSLAM_ECUDA;
  int rc;
  for (int k = 0; k < 2; k++)
    if ((rc = Bs[k]->ensure_events())) return rc;
  const int nchunks = (nprob + chunk - 1) / chunk;
  int first_err = DEFSLAM_OK;
  float total_ms = 0.f;
  auto finish = [&](int c) -> int {
    defslam_sft_batch *B = Bs[c & 1];
    DS_CUDA_TRY(cudaEventSynchronize(B->done));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, B->k0, B->k1) == cudaSuccess) total_ms += ms;
    const int e = B->bm.unpack((const uint8_t *)B->h_out.p, r + (size_t)c * chunk);
    if (e && first_err == DEFSLAM_OK) first_err = e;
    B->drop_temps();
    return 0;
  };
  for (int c = 0; c < nchunks; c++) {
    defslam_sft_batch *B = Bs[c & 1];
    if (c >= 2 && (rc = finish(c - 2))) return rc;
    const int off = c * chunk, cnt = nprob - off < chunk ? nprob - off : chunk;
    if ((rc = batch_load(B, cnt, p + off, MODE_SOLVE, cs))) return rc;
    DS_CUDA_TRY(cudaEventRecord(B->up, cs));
    DS_CUDA_TRY(cudaStreamWaitEvent(comp[c & 1], B->up, 0));
    if ((rc = batch_launch(B, B->k0, B->k1, comp[c & 1]))) return rc;
    DS_CUDA_TRY(cudaStreamWaitEvent(cs_down, B->k1, 0));
    DS_CUDA_TRY(cudaMemcpyAsync(B->h_out.p, B->d_out.p, B->bm.out_bytes, cudaMemcpyDeviceToHost, cs_down));
    DS_CUDA_TRY(cudaEventRecord(B->done, cs_down));
  }
  for (int c = nchunks >= 2 ? nchunks - 2 : 0; c < nchunks; c++)
    if ((rc = finish(c))) return rc;
  g_last_kernel_ms = total_ms;
  return first_err;
}

/* ------------------------------------------------------------------ ABI -- */

extern "C" {

int defslam_template_create(const defslam_template_desc *desc, int device, defslam_template **out) {
  if (!desc || !out) return DEFSLAM_EBADARG;
  *out = nullptr;
  DeviceGuard device_guard_;
  DevCtx *ctx = get_ctx(device);
  if (!ctx) return DEFSLAM_ECUDA;
  return template_make(desc, ctx, out);
}

void defslam_template_destroy(defslam_template *t) { template_free(t); }

int defslam_template_info(const defslam_template *t, int32_t *bandwidth, int32_t *band_ld, int32_t *dn_pad,
                          int32_t *n_blocks, int32_t *smem_bytes) {
  if (!t) return DEFSLAM_EBADARG;
  const PlanView &v = t->hview;
  if (bandwidth) *bandwidth = v.bw;
  if (band_ld) *band_ld = v.ld;
  if (dn_pad) *dn_pad = v.Dn_pad;
  if (n_blocks) *n_blocks = v.n_blk;
  if (smem_bytes)
    *smem_bytes = (int32_t)sizeof(double) *
                  (CTX_DOUBLES + smem_layout(v.n_nodes, v.n_edges, v.Dn_pad, v.bwp, v.ld, v.Wr, v.ES, true).total);
  return DEFSLAM_OK;
}

int defslam_sft_solve_batched(int32_t nprob, const defslam_sft_problem *p, defslam_sft_result *r, int device) {
  if (nprob < 0 || (nprob > 0 && (!p || !r))) return DEFSLAM_EBADARG;
  DeviceGuard device_guard_;
  DevCtx *ctx = get_ctx(device);
  if (!ctx) return DEFSLAM_ECUDA;
  if (nprob == 0) return DEFSLAM_OK;
  {
    const int chunk = 4 * ctx->sm_count; /* two waves at two CTAs per SM */
    if (nprob >= 3 * chunk && getenv("DEFSLAM_NO_PIPELINE") == nullptr) return solve_pipelined(ctx, nprob, p, r, chunk);
  }
  defslam_sft_batch *B = tl_batch(ctx);
  int rc;
  if ((rc = batch_load(B, nprob, p, MODE_SOLVE))) return rc;
  if ((rc = batch_launch(B))) return rc;
  if ((rc = batch_download(B))) return rc;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->e0, ctx->e1) == cudaSuccess) { B->last_ms = ms; g_last_kernel_ms = ms; }
  rc = B->bm.unpack((const uint8_t *)B->h_out.p, r);
  B->drop_temps();
  return rc;
}

int defslam_sft_solve(const defslam_sft_problem *p, defslam_sft_result *r) {
  if (!p || !r) return DEFSLAM_EBADARG;
  return defslam_sft_solve_batched(1, p, r, -1);
}

int defslam_sft_normal_equations(const defslam_sft_problem *p, double *H_dense, double *b, double *chi2) {
  if (!p) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  defslam_sft_batch *B = tl_batch(ctx);
  int rc;
  if ((rc = batch_load(B, 1, p, MODE_NORMAL_EQ))) return rc;
  if ((rc = batch_launch(B))) return rc;
  if ((rc = batch_download(B))) return rc;
  const ProbSlot &s = B->bm.slots[0];
  const uint8_t *o = (const uint8_t *)B->h_out.p + s.out_off;
  const ResultScalars *rs = (const ResultScalars *)(o + s.o_res);
  B->drop_temps();
  if (rs->status) return rs->status;
  const size_t D = 3 * (size_t)s.n_nodes + 6;
  if (H_dense) memcpy(H_dense, o + s.o_H, D * D * sizeof(double));
  if (b) memcpy(b, o + s.o_b, D * sizeof(double));
  if (chi2) *chi2 = rs->chi2_initial;
  return DEFSLAM_OK;
}

/* resident batch: inputs uploaded once, solve re-runnable, results fetched on demand */
int defslam_sft_batch_create(int32_t nprob, const defslam_sft_problem *p, int device, defslam_sft_batch **out) {
  if (!out || nprob <= 0 || !p) return DEFSLAM_EBADARG;
  *out = nullptr;
  DeviceGuard device_guard_;
  DevCtx *ctx = get_ctx(device);
  if (!ctx) return DEFSLAM_ECUDA;
  std::unique_ptr<defslam_sft_batch> B(new defslam_sft_batch);
  B->ctx = ctx;
  int rc = batch_load(B.get(), nprob, p, MODE_SOLVE);
  if (rc) return rc;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return DEFSLAM_ECUDA; }
  *out = B.release();
  return DEFSLAM_OK;
}

int defslam_sft_batch_run(defslam_sft_batch *B) {
  if (!B) return DEFSLAM_EBADARG;
  if (cudaSetDevice(B->ctx->device) != cudaSuccess) { cudaGetLastError(); return DEFSLAM_ECUDA; }
  int rc = batch_launch(B);
  if (rc) return rc;
  return batch_wait_kernel(B);
}

int defslam_sft_batch_fetch(defslam_sft_batch *B, defslam_sft_result *r) {
  if (!B || !r) return DEFSLAM_EBADARG;
  if (cudaSetDevice(B->ctx->device) != cudaSuccess) { cudaGetLastError(); return DEFSLAM_ECUDA; }
  int rc = batch_download(B);
  if (rc) return rc;
  return B->bm.unpack((const uint8_t *)B->h_out.p, r);
}

int defslam_sft_batch_info(const defslam_sft_batch *B, int32_t *grid, int32_t *threads, int32_t *smem_bytes,
                           int64_t *h2d_bytes, int64_t *d2h_bytes, double *last_kernel_ms) {
  if (!B) return DEFSLAM_EBADARG;
  if (grid) *grid = B->grid;
  if (threads) *threads = SFT_THREADS;
  if (smem_bytes) *smem_bytes = B->smem_bytes;
  if (h2d_bytes) *h2d_bytes = (int64_t)(B->bm.in_bytes + sizeof(ProbView) * (size_t)B->nprob);
  if (d2h_bytes) *d2h_bytes = (int64_t)B->bm.out_bytes;
  if (last_kernel_ms) *last_kernel_ms = B->last_ms;
  return DEFSLAM_OK;
}

void defslam_sft_batch_destroy(defslam_sft_batch *B) {
  if (!B) return;
  cudaSetDevice(B->ctx->device);
  delete B;
}

}  // extern "C"
