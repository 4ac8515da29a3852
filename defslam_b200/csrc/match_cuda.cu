/*
 * match_cuda.cu -- CUDA kernels + C ABI of the projection search (match production for the SfT
 * solve).  Device arithmetic lives in match_core.h; see include/defslam_b200.h for what the entry
 * point replaces in the reference.
 */
#include <stdlib.h>

#include <vector>

#include "ds_runtime.h"
#include "match_core.h"

using namespace ds;

namespace {

struct Scratch {
  DevBuf dev, host, keys;
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
int grid_for(int n, int sm) {
  int g = (n + 127) / 128;
  if (g > sm * 8) g = sm * 8;
  return g < 1 ? 1 : g;
}

__global__ void cell_kernel(ProjView P, int *cell) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.n_cur; j += gridDim.x * blockDim.x) cell[j] = keypoint_cell(P, j);
}

/* FILL = false: projection + candidate count per map point; FILL = true: ranked candidate keys */
template <bool FILL>
__global__ void candidates_kernel(ProjView P, const int *cell, Proj *proj, int *cnt, const int *off, uint64_t *keys) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_last; i += gridDim.x * blockDim.x) {
    Proj r;
    if (FILL) r = proj[i];
    else { r = project_point(P, i); proj[i] = r; }
    int c = 0;
    if (r.ok) {
      uint64_t *out = FILL ? keys + off[i] : nullptr;
      const uint8_t *d = &P.last_desc[32 * (size_t)i];
      for (int j = 0; j < P.n_cur; j++) {
        const int cj = cell[j];
        if (!candidate_ok(P, r, j, cj)) continue;
        if (FILL) out[c] = cand_key(hamming256(d, &P.cur_desc[32 * (size_t)j]), cj, j);
        c++;
      }
    }
    if (!FILL) cnt[i] = c;
  }
}

__global__ void warp_cell_kernel(WarpView W, int *cell2) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < W.n2; j += gridDim.x * blockDim.x)
    cell2[j] = cell_of_xy(W.kp2[2 * j], W.kp2[2 * j + 1], W.min_x, W.min_y, W.gwi, W.ghi);
}

/* ---- low-latency path: one warp per map point, fixed-capacity candidate lists ------------------
 * The lanes of a warp stride over the keypoints of the current frame; candidates take slots in
 * keys[i*CAND_CAP ..) by ballot (their order does not matter: the resolve pass ranks by key).  A map
 * point with more than CAND_CAP candidates raises `overflow` and the host repeats the search with the
 * exact two-pass lists. */
constexpr int CAND_CAP = 64;

__global__ void candidates_warp_kernel(ProjView P, const int *cell, int *cnt, uint64_t *keys, int *overflow) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < P.n_last; i += gridDim.x * wpb) {
    const Proj r = project_point(P, i);
    int base = 0;
    if (r.ok) {
      const uint8_t *d = &P.last_desc[32 * (size_t)i];
      uint64_t *out = keys + (size_t)i * CAND_CAP;
      for (int j0 = 0; j0 < P.n_cur; j0 += 32) {
        const int j = j0 + lane;
        bool ok = false;
        int cj = -1;
        if (j < P.n_cur) { cj = cell[j]; ok = candidate_ok(P, r, j, cj); }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int pos = base + __popc(m & ((1u << lane) - 1u));
          if (pos < CAND_CAP) out[pos] = cand_key(hamming256(d, &P.cur_desc[32 * (size_t)j]), cj, j);
        }
        base += __popc(m);
      }
    }
    if (lane == 0) {
      cnt[i] = base < CAND_CAP ? base : CAND_CAP;
      if (base > CAND_CAP) atomicOr(overflow, 1);
    }
  }
}

/* resolve pass for the fixed-capacity lists: `taken` lives in shared memory, the first 32 keys of the
 * next map point are in flight while the current one is reduced */
__global__ void resolve_fixed_kernel(ProjView P, const int *cnt, const uint64_t *keys, const uint8_t *taken_in,
                                     int *match, int *acc_i, int *acc_j, int *nmatches_out) {
  extern __shared__ uint8_t taken[];
  const int lane = threadIdx.x;
  for (int j = lane; j < P.n_cur; j += 32) taken[j] = taken_in[j];
  __syncwarp();
  int nacc = 0, nmatches = 0;
  uint64_t knext = keys[lane];
  for (int i0 = 0; i0 < P.n_last; i0 += 32) {
    const int my_cnt = i0 + lane < P.n_last ? cnt[i0 + lane] : 0;
    const int tend = P.n_last - i0 < 32 ? P.n_last - i0 : 32;
    for (int t = 0; t < tend; t++) {
      const int i = i0 + t;
      const int n = __shfl_sync(0xffffffffu, my_cnt, t);
      uint64_t k = knext;
      if (i + 1 < P.n_last) knext = keys[(size_t)(i + 1) * CAND_CAP + lane];
      if (n == 0) continue;
      if (lane >= n || taken[key_index(k)]) k = ~0ull;
      uint64_t best = k;
      if (n > 32) {
        uint64_t k2 = ~0ull;
        if (32 + lane < n) {
          k2 = keys[(size_t)i * CAND_CAP + 32 + lane];
          if (taken[key_index(k2)]) k2 = ~0ull;
        }
        best = k2 < best ? k2 : best;
      }
      {
        /* 64-bit minimum by two hardware warp reductions: (distance, cell) first, then the index */
        const unsigned hi = best == ~0ull ? 0xffffffffu : (unsigned)(best >> 32);
        const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
        const unsigned lo = hi == mhi ? (unsigned)(best & 0xffffffffu) : 0xffffffffu;
        const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
        best = mhi == 0xffffffffu ? ~0ull : (((uint64_t)mhi << 32) | mlo);
      }
      if (best != ~0ull && key_dist(best) < 256 && key_dist(best) <= P.th_high) {
        const int j = key_index(best);
        if (lane == 0) {
          match[j] = i;
          taken[j] = P.last_has_obs[i];
          acc_i[nacc] = i;
          acc_j[nacc] = j;
        }
        nacc++;
        nmatches++;
      }
      __syncwarp();
    }
  }
  if (P.check_orientation) {
    __shared__ int hist[HISTO_LENGTH];
    if (lane < HISTO_LENGTH) hist[lane] = 0;
    __syncwarp();
    if (lane == 0) {
      for (int a = 0; a < nacc; a++) hist[rotation_bin(P.last_angle[acc_i[a]], P.cur_angle[acc_j[a]])]++;
      int i1, i2, i3;
      three_maxima(hist, HISTO_LENGTH, i1, i2, i3);
      for (int a = 0; a < nacc; a++) {
        const int bin = rotation_bin(P.last_angle[acc_i[a]], P.cur_angle[acc_j[a]]);
        if (bin != i1 && bin != i2 && bin != i3) { match[acc_j[a]] = -1; nmatches--; }
      }
    }
  }
  if (lane == 0) *nmatches_out = nmatches;
}


/* ---- fused path: ONE launch (one CTA of 1024 threads) for cells, projections, candidate lists and the resolve.
 * The reference loop is order dependent only through the taken flags: point i takes its best-ranked candidate that no
 * EARLIER point with observations took.  That is the unique fixed point of
 *     choice[i] = best candidate j of i with owner[j] >= i,   owner[j] = min { i : has_obs[i], choice[i] = j }
 * (induction on i: choice[i] depends on the choices of smaller indices only), so instead of replaying 1200 points
 * one after the other the CTA iterates the two maps in parallel until nothing changes -- as many rounds as the
 * longest chain of displaced points (a handful), not as many as there are points.  What the loop leaves in match[]
 * is reproduced exactly: the last writer of match[j] is the taker if there is one, else the largest chooser without
 * observations; nmatches counts every point that chose, and the rotation filter clears match[j] as soon as ANY point
 * that chose j falls outside the three dominant bins (DefORBmatcher.cc:424-446). */
constexpr int FUSED_THREADS = 1024;
__global__ void __launch_bounds__(FUSED_THREADS, 1)
search_fused_kernel(ProjView P, uint64_t *keys, int *cnt, int *choice, int *match,
                    int *tail /* nmatches, overflow, CTAs done */) {
  extern __shared__ int sm_i[];
  const int NC = P.n_cur, NL = P.n_last;
  int *cell = sm_i, *owner = cell + NC, *owner_new = owner + NC, *lastw = owner_new + NC;
  uint8_t *kill = (uint8_t *)(lastw + NC);
  __shared__ int hist[HISTO_LENGTH];
  __shared__ int s_flag, s_nm, s_keep[3], s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = FUSED_THREADS / 32;
  for (int j = tid; j < NC; j += FUSED_THREADS) cell[j] = keypoint_cell(P, j);
  __syncthreads();
  /* candidate lists, all CTAs: warp per map point, lanes stride over the keypoints (slots by ballot) */
  for (int i = blockIdx.x * nwarp + warp; i < NL; i += gridDim.x * nwarp) {
    const Proj r = project_point(P, i);
    int base = 0;
    if (r.ok) {
      const uint8_t *d = &P.last_desc[32 * (size_t)i];
      uint64_t *out = keys + (size_t)i * CAND_CAP;
      for (int j0 = 0; j0 < NC; j0 += 32) {
        const int j = j0 + lane;
        bool ok = false;
        int cj = -1;
        if (j < NC) { cj = cell[j]; ok = candidate_ok(P, r, j, cj); }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int pos = base + __popc(m & ((1u << lane) - 1u));
          if (pos < CAND_CAP) out[pos] = cand_key(hamming256(d, &P.cur_desc[32 * (size_t)j]), cj, j);
        }
        base += __popc(m);
      }
    }
    if (lane == 0) {
      cnt[i] = base < CAND_CAP ? base : CAND_CAP;
      if (base > CAND_CAP) atomicOr(&tail[1], 1);
    }
  }
  /* the CTA that finishes last resolves (its loads of the other CTAs' lists come after the fence + atomic) */
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&tail[2], 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (*(volatile int *)&tail[1]) return; /* overflow: the host repeats the search with exact lists */
  for (int j = tid; j < NC; j += FUSED_THREADS) {
    owner[j] = P.cur_taken[j] ? -1 : 0x7fffffff; /* -1: taken before the search started, blocks every point */
    lastw[j] = -1;
    kill[j] = 0;
  }
  if (tid < HISTO_LENGTH) hist[tid] = 0;
  if (tid == 0) { s_flag = 0; s_nm = 0; }
  __syncthreads();
  for (int i = tid; i < NL; i += FUSED_THREADS) choice[i] = -2; /* not evaluated yet */
  /* fixed-point rounds */
  for (int round = 0; round < NL + 2; round++) {
    for (int j = tid; j < NC; j += FUSED_THREADS) owner_new[j] = P.cur_taken[j] ? -1 : 0x7fffffff;
    __syncthreads();
    if (tid == 0) s_flag = 0;
    __syncthreads();
    bool changed = false;
    for (int i = tid; i < NL; i += FUSED_THREADS) {
      const int n = __ldcg(&cnt[i]);
      uint64_t best = ~0ull;
      const uint64_t *k = keys + (size_t)i * CAND_CAP;
      for (int c = 0; c < n; c++) {
        const uint64_t key = __ldcg((const unsigned long long *)&k[c]);
        if (owner[key_index(key)] >= i && key < best) best = key;
      }
      int ch = -1;
      if (best != ~0ull && key_dist(best) < 256 && key_dist(best) <= P.th_high) ch = key_index(best);
      if (ch != choice[i]) { choice[i] = ch; changed = true; }
      if (ch >= 0 && P.last_has_obs[i]) atomicMin(&owner_new[ch], i);
    }
    if (changed) s_flag = 1;
    __syncthreads();
    const bool again = s_flag != 0;
    for (int j = tid; j < NC; j += FUSED_THREADS) owner[j] = owner_new[j];
    __syncthreads();
    if (!again) break;
  }
  /* what the sequential loop leaves behind */
  int mine = 0;
  for (int i = tid; i < NL; i += FUSED_THREADS) {
    const int j = choice[i];
    if (j < 0) continue;
    mine++;
    if (owner[j] == 0x7fffffff) atomicMax(&lastw[j], i); /* no taker: the largest chooser wrote last */
    if (P.check_orientation) atomicAdd(&hist[rotation_bin(P.last_angle[i], P.cur_angle[j])], 1);
  }
  if (mine) atomicAdd(&s_nm, mine);
  __syncthreads();
  if (P.check_orientation) {
    if (tid == 0) { int i1, i2, i3; three_maxima(hist, HISTO_LENGTH, i1, i2, i3); s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3; }
    __syncthreads();
    int dropped = 0;
    for (int i = tid; i < NL; i += FUSED_THREADS) {
      const int j = choice[i];
      if (j < 0) continue;
      const int bin = rotation_bin(P.last_angle[i], P.cur_angle[j]);
      if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { kill[j] = 1; dropped++; }
    }
    if (dropped) atomicSub(&s_nm, dropped);
    __syncthreads();
  }
  for (int j = tid; j < NC; j += FUSED_THREADS) {
    const int o = owner[j];
    match[j] = kill[j] ? -1 : (o != 0x7fffffff && o >= 0 ? o : lastw[j]);
  }
  if (tid == 0) tail[0] = s_nm;
}

/* warp per keypoint of keyframe 1: lanes stride over the keypoints of keyframe 2 */
__global__ void warp_search_warp_kernel(WarpView W, const int *cell2, int *match12, int *nmatches) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < W.n1; i += gridDim.x * wpb) {
    const uint64_t best = warp_search_lanes(W, cell2, i, lane, 32);
    uint64_t b = best;
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, b, o);
      b = other < b ? other : b;
    }
    if (lane == 0) {
      match12[i] = b == ~0ull ? -1 : key_index(b);
      if (b != ~0ull) atomicAdd(nmatches, 1);
    }
  }
}

/* exclusive scan of cnt[0..n) by one CTA; total -> off[n] */
__global__ void scan_kernel(const int *cnt, int *off, int n) {
  __shared__ int part[1024];
  const int chunk = (n + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * chunk, e = min(b + chunk, n);
  int s = 0;
  for (int i = b; i < e; i++) s += cnt[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int t = 0; t < (int)blockDim.x; t++) { const int v = part[t]; part[t] = acc; acc += v; }
    off[n] = acc;
  }
  __syncthreads();
  int acc = part[threadIdx.x];
  for (int i = b; i < e; i++) { off[i] = acc; acc += cnt[i]; }
}

/* The order-dependent tail, by one warp: map points in keypoint order, each taking its best-ranked
 * candidate that is still free; then the rotation-consistency filter. */
__global__ void resolve_kernel(ProjView P, const int *cnt, const int *off, const uint64_t *keys, uint8_t *taken,
                               int *match, int *acc_i, int *acc_j, int *nmatches_out) {
  const int lane = threadIdx.x;
  int nacc = 0, nmatches = 0;
  for (int i = 0; i < P.n_last; i++) {
    const int n = cnt[i];
    if (n == 0) continue;
    uint64_t best = ~0ull;
    for (int b = 0; b < n; b += 32) {
      uint64_t k = ~0ull;
      if (b + lane < n) {
        k = keys[off[i] + b + lane];
        if (taken[key_index(k)]) k = ~0ull;
      }
      best = k < best ? k : best;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    /* bestDist starts at 256 and only a strictly smaller distance replaces it; accept when <= TH_HIGH */
    if (best != ~0ull && key_dist(best) < 256 && key_dist(best) <= P.th_high) {
      const int j = key_index(best);
      if (lane == 0) {
        match[j] = i;
        taken[j] = P.last_has_obs[i];
        acc_i[nacc] = i;
        acc_j[nacc] = j;
      }
      nacc++;
      nmatches++;
    }
    __syncwarp();
  }
  if (P.check_orientation) {
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int keep[3];
    if (lane < HISTO_LENGTH) hist[lane] = 0;
    __syncwarp();
    if (lane == 0) {
      for (int a = 0; a < nacc; a++) hist[rotation_bin(P.last_angle[acc_i[a]], P.cur_angle[acc_j[a]])]++;
      int i1, i2, i3;
      three_maxima(hist, HISTO_LENGTH, i1, i2, i3);
      keep[0] = i1; keep[1] = i2; keep[2] = i3;
      for (int a = 0; a < nacc; a++) {
        const int bin = rotation_bin(P.last_angle[acc_i[a]], P.cur_angle[acc_j[a]]);
        if (bin != keep[0] && bin != keep[1] && bin != keep[2]) { match[acc_j[a]] = -1; nmatches--; }
      }
    }
  }
  if (lane == 0) *nmatches_out = nmatches;
}

}  // namespace

extern "C" {

int defslam_search_by_projection(const defslam_projsearch_problem *p, int32_t *match_out, int32_t *nmatches_out) {
  if (!p || !match_out || !nmatches_out || p->n_last < 0 || p->n_cur < 0 || p->n_levels <= 0 || !p->scale_factors)
    return DEFSLAM_EBADARG;
  if (p->n_last > 0 && (!p->last_state || !p->last_has_obs || !p->last_world_xyz || !p->last_desc || !p->last_octave ||
                        !p->last_angle))
    return DEFSLAM_EBADARG;
  if (p->n_cur > 0 && (!p->cur_xy || !p->cur_octave || !p->cur_angle || !p->cur_desc || !p->cur_uright || !p->cur_taken))
    return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  *nmatches_out = 0;
  const size_t NL = (size_t)p->n_last, NC = (size_t)p->n_cur;
  for (size_t j = 0; j < NC; j++) match_out[j] = -1;
  if (NL == 0 || NC == 0) return DEFSLAM_OK;
  Packer in, wk;
  const size_t o_ls = in.add(NL), o_lo = in.add(NL), o_lx = in.add(NL * 12), o_ld = in.add(NL * 32), o_loc = in.add(NL * 4),
               o_la = in.add(NL * 4), o_cx = in.add(NC * 8), o_co = in.add(NC * 4), o_ca = in.add(NC * 4),
               o_cd = in.add(NC * 32), o_cu = in.add(NC * 4), o_ct = in.add(NC), o_sc = in.add((size_t)p->n_levels * 4);
  const size_t w_cell = wk.add(NC * 4), w_proj = wk.add(NL * sizeof(Proj)), w_cnt = wk.add(NL * 4),
               w_off = wk.add((NL + 1) * 4), w_ai = wk.add(NL * 4), w_aj = wk.add(NL * 4), w_match = wk.add(NC * 4),
               w_n = wk.add(16);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > wk.total ? in.total : wk.total)) || (rc = S.dev.ensure(in.total + wk.total))) return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d = (uint8_t *)S.dev.p, *w = d + in.total;
  memcpy(h + o_ls, p->last_state, NL); memcpy(h + o_lo, p->last_has_obs, NL);
  memcpy(h + o_lx, p->last_world_xyz, NL * 12); memcpy(h + o_ld, p->last_desc, NL * 32);
  memcpy(h + o_loc, p->last_octave, NL * 4); memcpy(h + o_la, p->last_angle, NL * 4);
  memcpy(h + o_cx, p->cur_xy, NC * 8); memcpy(h + o_co, p->cur_octave, NC * 4); memcpy(h + o_ca, p->cur_angle, NC * 4);
  memcpy(h + o_cd, p->cur_desc, NC * 32); memcpy(h + o_cu, p->cur_uright, NC * 4); memcpy(h + o_ct, p->cur_taken, NC);
  memcpy(h + o_sc, p->scale_factors, (size_t)p->n_levels * 4);
  DS_CUDA_TRY(cudaMemcpyAsync(d, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  ProjView V;
  V.n_last = p->n_last; V.n_cur = p->n_cur; V.n_levels = p->n_levels;
  V.last_state = d + o_ls; V.last_has_obs = d + o_lo; V.last_desc = d + o_ld; V.cur_desc = d + o_cd; V.cur_taken = d + o_ct;
  V.last_xyz = (const float *)(d + o_lx); V.last_angle = (const float *)(d + o_la); V.cur_xy = (const float *)(d + o_cx);
  V.cur_angle = (const float *)(d + o_ca); V.cur_uright = (const float *)(d + o_cu); V.scale = (const float *)(d + o_sc);
  V.last_octave = (const int *)(d + o_loc); V.cur_octave = (const int *)(d + o_co);
  for (int k = 0; k < 16; k++) V.Tcw[k] = p->T_cw[k];
  V.fx = p->fx; V.fy = p->fy; V.cx = p->cx; V.cy = p->cy; V.mbf = p->mbf;
  V.min_x = p->min_x; V.max_x = p->max_x; V.min_y = p->min_y; V.max_y = p->max_y;
  V.gwi = p->grid_width_inv; V.ghi = p->grid_height_inv; V.th = p->th;
  V.th_high = p->th_high; V.check_orientation = p->check_orientation;
  {
    /* twc = -Rcw^T tcw; tlc = Rlw twc + tlw (fp32 cv::Mat products); only the sign tests on tlc.z matter */
    float twc[3], tlc2;
    for (int a = 0; a < 3; a++)
      twc[a] = -(p->T_cw[a] * p->T_cw[3] + p->T_cw[4 + a] * p->T_cw[7] + p->T_cw[8 + a] * p->T_cw[11]);
    tlc2 = p->T_lw[8] * twc[0] + p->T_lw[9] * twc[1] + p->T_lw[10] * twc[2] + p->T_lw[11];
    V.forward = tlc2 > p->mb && !p->mono;
    V.backward = -tlc2 > p->mb && !p->mono;
  }
  int *cell = (int *)(w + w_cell), *cnt = (int *)(w + w_cnt), *off = (int *)(w + w_off);
  Proj *proj = (Proj *)(w + w_proj);
  const int wgrid = (p->n_last + 3) / 4 < ctx->sm_count * 16 ? (p->n_last + 3) / 4 : ctx->sm_count * 16;
  bool exact = NC > 48 * 1024; /* the resolve pass keeps the taken flags in shared memory */
  /* fused path: one launch, one synchronisation (cells, owners and flags of the current frame in shared memory) */
  const size_t fused_smem = NC * 17;
  if (!exact && fused_smem <= (size_t)ctx->smem_optin - 2048 && getenv("DEFSLAM_MATCH_UNFUSED") == nullptr) {
    if ((rc = S.keys.ensure(NL * CAND_CAP * 8))) return rc;
    DS_CUDA_TRY(raise_dynamic_smem((const void *)search_fused_kernel, ctx->device, (int)fused_smem));
    DS_CUDA_TRY(cudaMemsetAsync(w + w_n, 0, 16, ctx->stream));
    int fgrid = (p->n_last + 31) / 32;
    if (fgrid > ctx->sm_count) fgrid = ctx->sm_count;
    search_fused_kernel<<<fgrid, FUSED_THREADS, fused_smem, ctx->stream>>>(V, (uint64_t *)S.keys.p, cnt, (int *)(w + w_ai),
                                                                           (int *)(w + w_match), (int *)(w + w_n));
    DS_CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    /* matches and the two tail words in one copy (w_match and w_n are adjacent arena slots) */
    DS_CUDA_TRY(cudaMemcpyAsync(h, w + w_match, (w_n - w_match) + 16, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const int *tail = (const int *)(h + (w_n - w_match));
    if (!tail[1]) {
      memcpy(match_out, h, NC * 4);
      *nmatches_out = tail[0];
      return DEFSLAM_OK;
    }
    exact = true; /* a map point had more than CAND_CAP candidates: exact lists */
  }
  cell_kernel<<<grid_for(p->n_cur, ctx->sm_count), 128, 0, ctx->stream>>>(V, cell);
  if (!exact) {
    /* low-latency path: three launches, one synchronisation */
    if ((rc = S.keys.ensure(NL * CAND_CAP * 8))) return rc;
    DS_CUDA_TRY(cudaMemsetAsync(w + w_match, 0xff, NC * 4, ctx->stream));
    DS_CUDA_TRY(cudaMemsetAsync(w + w_n, 0, 8, ctx->stream));
    candidates_warp_kernel<<<wgrid, 128, 0, ctx->stream>>>(V, cell, cnt, (uint64_t *)S.keys.p, (int *)(w + w_n) + 1);
    resolve_fixed_kernel<<<1, 32, NC, ctx->stream>>>(V, cnt, (const uint64_t *)S.keys.p, d + o_ct, (int *)(w + w_match),
                                                     (int *)(w + w_ai), (int *)(w + w_aj), (int *)(w + w_n));
    DS_CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(3);
    int tail[2] = {0, 0};
    DS_CUDA_TRY(cudaMemcpyAsync(h, w + w_match, NC * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA_TRY(cudaMemcpyAsync(tail, w + w_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (!tail[1]) {
      memcpy(match_out, h, NC * 4);
      *nmatches_out = tail[0];
      return DEFSLAM_OK;
    }
    exact = true; /* a map point had more than CAND_CAP candidates: exact lists */
  }
  DS_CUDA_TRY(cudaMemsetAsync(w + w_match, 0xff, NC * 4, ctx->stream));
  candidates_kernel<false><<<grid_for(p->n_last, ctx->sm_count), 128, 0, ctx->stream>>>(V, cell, proj, cnt, nullptr, nullptr);
  scan_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, off, p->n_last);
  DS_CUDA_TRY(cudaGetLastError());
  int total = 0;
  DS_CUDA_TRY(cudaMemcpyAsync(&total, off + NL, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if ((rc = S.keys.ensure((size_t)(total > 0 ? total : 1) * 8))) return rc;
  candidates_kernel<true><<<grid_for(p->n_last, ctx->sm_count), 128, 0, ctx->stream>>>(V, cell, proj, cnt, off,
                                                                                      (uint64_t *)S.keys.p);
  /* cur_taken is an input copy in the arena: the resolve pass updates it in place */
  resolve_kernel<<<1, 32, 0, ctx->stream>>>(V, cnt, off, (const uint64_t *)S.keys.p, d + o_ct, (int *)(w + w_match),
                                            (int *)(w + w_ai), (int *)(w + w_aj), (int *)(w + w_n));
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(5);
  DS_CUDA_TRY(cudaMemcpyAsync(h, w + w_match, NC * 4, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaMemcpyAsync(h + ((NC * 4 + 255) & ~(size_t)255), w + w_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(match_out, h, NC * 4);
  memcpy(nmatches_out, h + ((NC * 4 + 255) & ~(size_t)255), 4);
  return DEFSLAM_OK;
}

int defslam_search_by_schwarp(const defslam_warpsearch_problem *p, int32_t *match12_out, int32_t *nmatches_out) {
  if (!p || !match12_out || !nmatches_out || p->n1 < 0 || p->n2 < 0 || !p->x) return DEFSLAM_EBADARG;
  if (p->bbs.nptsu < 4 || p->bbs.nptsv < 4 || p->bbs.valdim != 2 || !(p->bbs.umax > p->bbs.umin) || !(p->bbs.vmax > p->bbs.vmin))
    return DEFSLAM_EBADARG;
  if (p->n1 > 0 && (!p->kp1_norm || !p->kp1_state || !p->kp1_desc)) return DEFSLAM_EBADARG;
  if (p->n2 > 0 && (!p->kp2_xy || !p->kp2_has_mp || !p->kp2_desc)) return DEFSLAM_EBADARG;
  DevCtx *ctx = get_ctx(-1);
  if (!ctx) return DEFSLAM_ECUDA;
  *nmatches_out = 0;
  const size_t N1 = (size_t)p->n1, N2 = (size_t)p->n2, NC = (size_t)p->bbs.nptsu * p->bbs.nptsv;
  for (size_t i = 0; i < N1; i++) match12_out[i] = -1;
  if (N1 == 0 || N2 == 0) return DEFSLAM_OK;
  Packer in, wk;
  const size_t o_c = in.add(NC * 16), o_k1 = in.add(N1 * 8), o_s1 = in.add(N1), o_d1 = in.add(N1 * 32), o_k2 = in.add(N2 * 8),
               o_h2 = in.add(N2), o_d2 = in.add(N2 * 32);
  const size_t w_cell = wk.add(N2 * 4), w_m = wk.add(N1 * 4), w_n = wk.add(4);
  Scratch &S = tl_scratch(ctx->device);
  int rc;
  if ((rc = S.host.ensure(in.total > wk.total ? in.total : wk.total)) || (rc = S.dev.ensure(in.total + wk.total))) return rc;
  uint8_t *h = (uint8_t *)S.host.p, *d = (uint8_t *)S.dev.p, *w = d + in.total;
  double *ctrl = (double *)(h + o_c); /* Array[valdim*l + n] = ControlPoints(l, n), Schwarp.cc:185-193 */
  for (size_t l = 0; l < NC; l++) { ctrl[2 * l] = p->x[l]; ctrl[2 * l + 1] = p->x[NC + l]; }
  memcpy(h + o_k1, p->kp1_norm, N1 * 8); memcpy(h + o_s1, p->kp1_state, N1); memcpy(h + o_d1, p->kp1_desc, N1 * 32);
  memcpy(h + o_k2, p->kp2_xy, N2 * 8); memcpy(h + o_h2, p->kp2_has_mp, N2); memcpy(h + o_d2, p->kp2_desc, N2 * 32);
  DS_CUDA_TRY(cudaMemcpyAsync(d, h, in.total, cudaMemcpyHostToDevice, ctx->stream));
  WarpView W;
  W.bbs.umin = p->bbs.umin; W.bbs.umax = p->bbs.umax; W.bbs.vmin = p->bbs.vmin; W.bbs.vmax = p->bbs.vmax;
  W.bbs.nptsu = p->bbs.nptsu; W.bbs.nptsv = p->bbs.nptsv; W.bbs.valdim = 2;
  W.ctrl = (const double *)(d + o_c); W.n1 = p->n1; W.n2 = p->n2;
  W.kp1 = (const float *)(d + o_k1); W.kp2 = (const float *)(d + o_k2); W.st1 = d + o_s1; W.d1 = d + o_d1; W.has2 = d + o_h2;
  W.d2 = d + o_d2; W.fx = p->fx; W.fy = p->fy; W.cx = p->cx; W.cy = p->cy; W.min_x = p->min_x; W.max_x = p->max_x;
  W.min_y = p->min_y; W.max_y = p->max_y; W.gwi = p->grid_width_inv; W.ghi = p->grid_height_inv; W.radius = p->radius;
  W.th_low = p->th_low;
  DS_CUDA_TRY(cudaMemsetAsync(w + w_n, 0, 4, ctx->stream));
  warp_cell_kernel<<<grid_for(p->n2, ctx->sm_count), 128, 0, ctx->stream>>>(W, (int *)(w + w_cell));
  {
    const int wgrid = (p->n1 + 3) / 4 < ctx->sm_count * 16 ? (p->n1 + 3) / 4 : ctx->sm_count * 16;
    warp_search_warp_kernel<<<wgrid, 128, 0, ctx->stream>>>(W, (const int *)(w + w_cell), (int *)(w + w_m), (int *)(w + w_n));
  }
  DS_CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(2);
  DS_CUDA_TRY(cudaMemcpyAsync(h, w + w_m, (N1 * 4 + 255 & ~(size_t)255) + 4, cudaMemcpyDeviceToHost, ctx->stream));
  DS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(match12_out, h, N1 * 4);
  memcpy(nmatches_out, h + ((N1 * 4 + 255) & ~(size_t)255), 4);
  return DEFSLAM_OK;
}

}  // extern "C"
