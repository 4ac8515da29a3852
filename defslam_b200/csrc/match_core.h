/*
 * match_core.h -- projection search of template points (DefORBmatcher::SearchByProjection,
 * Modules/Matching/DefORBmatcher.cc:296-451) as data-parallel pieces:
 *
 *   project_point()    the per-map-point preamble (:323-357): fp32 cv::Mat arithmetic, image bounds,
 *                      search radius, the cell window of Frame::GetFeaturesInArea (Frame.cc:428-448)
 *   candidate_ok()     membership of one current-frame keypoint in that window (Frame.cc:452-478 via
 *                      PosInGrid :484-496, the level test, the |dx|,|dy| < r box, the stereo test :377-383)
 *   hamming256()       ORBmatcher::DescriptorDistance (ORBmatcher.cc:1691-1707)
 *   resolve_in_order() the order-dependent part: assignments replayed in keypoint order, rotation
 *                      histogram and ComputeThreeMaxima (:404-447, ORBmatcher.cc:1645-1687)
 *
 * A candidate is ranked by (distance, cell column, cell row, keypoint index): the reference keeps
 * the FIRST minimum of a list that GetFeaturesInArea fills cell column by cell column, row by row,
 * each cell in keypoint order.
 */
#ifndef DS_MATCH_CORE_H_
#define DS_MATCH_CORE_H_

#include "bbs_core.h"
#include "ds_common.h"
#include "newpts_core.h" /* mul_rn / add_rn */

namespace ds {

constexpr int GRID_COLS = 64, GRID_ROWS = 48, HISTO_LENGTH = 30;

struct ProjView {
  int n_last, n_cur, n_levels;
  const uint8_t *last_state, *last_has_obs, *last_desc, *cur_desc, *cur_taken;
  const float *last_xyz, *last_angle, *cur_xy, *cur_angle, *cur_uright, *scale;
  const int *last_octave, *cur_octave;
  float Tcw[16];
  float fx, fy, cx, cy, mbf;
  float min_x, max_x, min_y, max_y, gwi, ghi, th;
  int forward, backward, th_high, check_orientation;
};

struct Proj {
  float u, v, radius, invzc;
  int cx0, cx1, cy0, cy1; /* cell window, empty when cx0 > cx1 */
  int lmin, lmax;         /* level window of GetFeaturesInArea */
  int ok;
};

DS_FN int floor_to_int(float a) { return (int)floorf(a); }
DS_FN int ceil_to_int(float a) { return (int)ceilf(a); }

/* Rcw * x3Dw + tcw on CV_32F cv::Mat (cv::gemm small-matrix path: fp32 products summed left to
 * right, then the addend), then the pinhole projection in fp32 */
DS_FN Proj project_point(const ProjView &P, int i) {
  Proj r;
  r.ok = 0; r.cx0 = 1; r.cx1 = 0; r.cy0 = 1; r.cy1 = 0; r.u = r.v = r.radius = r.invzc = 0.f; r.lmin = -1; r.lmax = -1;
  if (!P.last_state[i]) return r;
  const float *X = &P.last_xyz[3 * i];
  float c[3];
  for (int a = 0; a < 3; a++) {
    float s = mul_rn(P.Tcw[4 * a], X[0]);
    s = add_rn(s, mul_rn(P.Tcw[4 * a + 1], X[1]));
    s = add_rn(s, mul_rn(P.Tcw[4 * a + 2], X[2]));
    c[a] = add_rn(s, P.Tcw[4 * a + 3]);
  }
  const float invzc = (float)(1.0 / (double)c[2]);
  if (invzc < 0) return r;
  const float u = add_rn(mul_rn(mul_rn(P.fx, c[0]), invzc), P.cx);
  const float v = add_rn(mul_rn(mul_rn(P.fy, c[1]), invzc), P.cy);
  if (u < P.min_x || u > P.max_x) return r;
  if (v < P.min_y || v > P.max_y) return r;
  const int oct = P.last_octave[i];
  if (oct < 0 || oct >= P.n_levels) return r;
  const float radius = mul_rn(P.th, P.scale[oct]);
  if (P.forward) { r.lmin = oct; r.lmax = -1; }
  else if (P.backward) { r.lmin = 0; r.lmax = oct; }
  else { r.lmin = oct - 1; r.lmax = oct + 1; }
  r.u = u; r.v = v; r.radius = radius; r.invzc = invzc;
  /* Frame::GetFeaturesInArea, Frame.cc:428-448 */
  int a0 = floor_to_int(mul_rn(add_rn(add_rn(u, -P.min_x), -radius), P.gwi));
  if (a0 < 0) a0 = 0;
  if (a0 >= GRID_COLS) return r;
  int a1 = ceil_to_int(mul_rn(add_rn(add_rn(u, -P.min_x), radius), P.gwi));
  if (a1 > GRID_COLS - 1) a1 = GRID_COLS - 1;
  if (a1 < 0) return r;
  int b0 = floor_to_int(mul_rn(add_rn(add_rn(v, -P.min_y), -radius), P.ghi));
  if (b0 < 0) b0 = 0;
  if (b0 >= GRID_ROWS) return r;
  int b1 = ceil_to_int(mul_rn(add_rn(add_rn(v, -P.min_y), radius), P.ghi));
  if (b1 > GRID_ROWS - 1) b1 = GRID_ROWS - 1;
  if (b1 < 0) return r;
  r.cx0 = a0; r.cx1 = a1; r.cy0 = b0; r.cy1 = b1;
  r.ok = 1;
  return r;
}

/* Frame::PosInGrid (Frame.cc:484-496): cell of a current-frame keypoint, -1 when outside the grid */
DS_FN int keypoint_cell(const ProjView &P, int j) {
  const int px = (int)roundf(mul_rn(add_rn(P.cur_xy[2 * j], -P.min_x), P.gwi));
  const int py = (int)roundf(mul_rn(add_rn(P.cur_xy[2 * j + 1], -P.min_y), P.ghi));
  if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) return -1;
  return px * GRID_ROWS + py;
}

/* would GetFeaturesInArea list keypoint j (cell `cell`) for this projection, and does it pass the
 * stereo test of the matcher?  (the `taken` test is order-dependent: resolve_in_order) */
DS_FN bool candidate_ok(const ProjView &P, const Proj &r, int j, int cell) {
  if (cell < 0) return false;
  const int px = cell / GRID_ROWS, py = cell - px * GRID_ROWS;
  if (px < r.cx0 || px > r.cx1 || py < r.cy0 || py > r.cy1) return false;
  const bool check_levels = (r.lmin > 0) || (r.lmax >= 0);
  if (check_levels) {
    const int oc = P.cur_octave[j];
    if (oc < r.lmin) return false;
    if (r.lmax >= 0 && oc > r.lmax) return false;
  }
  const float dx = add_rn(P.cur_xy[2 * j], -r.u), dy = add_rn(P.cur_xy[2 * j + 1], -r.v);
  if (!(fabsf(dx) < r.radius && fabsf(dy) < r.radius)) return false;
  if (P.cur_uright[j] > 0) {
    const float ur = add_rn(r.u, -mul_rn(P.mbf, r.invzc));
    const float er = fabsf(add_rn(ur, -P.cur_uright[j]));
    if (er > r.radius) return false;
  }
  return true;
}

DS_FN int popc32(uint32_t v) {
#if DS_CUDA
  return __popc(v);
#else
  v = v - ((v >> 1) & 0x55555555u);
  v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
  return (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
#endif
}

DS_FN int hamming256(const uint8_t *a, const uint8_t *b) {
  const uint32_t *pa = (const uint32_t *)a, *pb = (const uint32_t *)b;
  int d = 0;
  for (int k = 0; k < 8; k++) d += popc32(pa[k] ^ pb[k]);
  return d;
}

/* rank of a candidate: smaller wins.  dist < 512, cell < 4096, index < 2^31 */
DS_FN uint64_t cand_key(int dist, int cell, int j) {
  return ((uint64_t)(uint32_t)dist << 44) | ((uint64_t)(uint32_t)cell << 32) | (uint64_t)(uint32_t)j;
}
DS_FN int key_dist(uint64_t k) { return (int)(k >> 44); }
DS_FN int key_index(uint64_t k) { return (int)(uint32_t)(k & 0xffffffffu); }

/* rotation bin of a match (DefORBmatcher.cc:409-419) */
DS_FN int rotation_bin(float angle_last, float angle_cur) {
  float rot = add_rn(angle_last, -angle_cur);
  if (rot < 0.0f) rot = add_rn(rot, 360.0f);
  int bin = (int)roundf(mul_rn(rot, 1.0f / HISTO_LENGTH));
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

/* ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:1645-1687) on bin sizes */
DS_FN void three_maxima(const int *size, int L, int &ind1, int &ind2, int &ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  ind1 = ind2 = ind3 = -1;
  for (int i = 0; i < L; i++) {
    const int s = size[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

/* ---- warp-guided search (DefORBmatcher::searchBySchwarp, DefORBmatcher.cc:190-293) ---- */

struct WarpView {
  BbsView bbs;
  const double *ctrl; /* interleaved [NC][2] like Warp::getEstimates builds it */
  int n1, n2;
  const float *kp1, *kp2;
  const uint8_t *st1, *d1, *has2, *d2;
  float fx, fy, cx, cy, min_x, max_x, min_y, max_y, gwi, ghi, radius;
  int th_low;
};

DS_FN int cell_of_xy(float x, float y, float min_x, float min_y, float gwi, float ghi) {
  const int px = (int)roundf(mul_rn(add_rn(x, -min_x), gwi));
  const int py = (int)roundf(mul_rn(add_rn(y, -min_y), ghi));
  if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) return -1;
  return px * GRID_ROWS + py;
}

/* best-ranked candidate key among the keypoints j = first, first+stride, ... of keyframe 2 for keypoint i
 * of keyframe 1 (~0 when none): one thread scans everything with (0, 1), a warp splits the scan by lane */
DS_FN uint64_t warp_search_lanes(const WarpView &W, const int *cell2, int i, int first, int stride) {
  if (!W.st1[i]) return ~0ull;
  double val[2];
  if (!bbs_eval_site(W.bbs, W.ctrl, (double)W.kp1[2 * i], (double)W.kp1[2 * i + 1], 0, 0, val)) return ~0ull;
  const float ex = (float)val[0], ey = (float)val[1]; /* cv::KeyPoint stores floats */
  const float x = add_rn(mul_rn(ex, W.fx), W.cx), y = add_rn(mul_rn(ey, W.fy), W.cy);
  if (!(x >= W.min_x && x < W.max_x && y >= W.min_y && y < W.max_y)) return ~0ull; /* KeyFrame::IsInImage */
  const float r = W.radius;
  int a0 = floor_to_int(mul_rn(add_rn(add_rn(x, -W.min_x), -r), W.gwi));
  if (a0 < 0) a0 = 0;
  if (a0 >= GRID_COLS) return ~0ull;
  int a1 = ceil_to_int(mul_rn(add_rn(add_rn(x, -W.min_x), r), W.gwi));
  if (a1 > GRID_COLS - 1) a1 = GRID_COLS - 1;
  if (a1 < 0) return ~0ull;
  int b0 = floor_to_int(mul_rn(add_rn(add_rn(y, -W.min_y), -r), W.ghi));
  if (b0 < 0) b0 = 0;
  if (b0 >= GRID_ROWS) return ~0ull;
  int b1 = ceil_to_int(mul_rn(add_rn(add_rn(y, -W.min_y), r), W.ghi));
  if (b1 > GRID_ROWS - 1) b1 = GRID_ROWS - 1;
  if (b1 < 0) return ~0ull;
  uint64_t best = ~0ull;
  for (int j = first; j < W.n2; j += stride) {
    const int cj = cell2[j];
    if (cj < 0 || W.has2[j]) continue;
    const int px = cj / GRID_ROWS, py = cj - px * GRID_ROWS;
    if (px < a0 || px > a1 || py < b0 || py > b1) continue;
    const float dx = add_rn(W.kp2[2 * j], -x), dy = add_rn(W.kp2[2 * j + 1], -y);
    if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;
    const int dist = hamming256(&W.d1[32 * (size_t)i], &W.d2[32 * (size_t)j]);
    if (dist >= W.th_low) continue;
    const uint64_t k = cand_key(dist, cj, j);
    if (k < best) best = k;
  }
  return best;
}

/* best keypoint of keyframe 2 for keypoint i of keyframe 1, or -1 */
DS_FN int warp_search_one(const WarpView &W, const int *cell2, int i) {
  const uint64_t best = warp_search_lanes(W, cell2, i, 0, 1);
  return best == ~0ull ? -1 : key_index(best);
}

}  // namespace ds
#endif
