/*
 * match_oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the projection search
 *   DefORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)   Modules/Matching/DefORBmatcher.cc:296-451
 *   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea          Thirdparty/ORBSLAM_2/src/Frame.cc:294-309,421-496
 *   ORBmatcher::DescriptorDistance / ComputeThreeMaxima                  Thirdparty/ORBSLAM_2/src/ORBmatcher.cc:1645-1707
 * in the reference's own control flow: the 64 x 48 grid of keypoint lists is built, the last
 * frame's keypoints are visited in order, every visit lists the features in the area cell by cell
 * and keeps the first minimum, assignments hide keypoints from later visits, rotation histogram last.
 *
 * Parity unpinned by the reference: these translation units need OpenCV (cv::Mat, cv::KeyPoint) and
 * cannot be compiled here, and the reference ships no test vectors.  The fp32 arithmetic is written
 * one rounding per operation (volatile), the form checked against cv2.gemm for the 3x3 * 3x1 + 3x1
 * product (tests/golden/newpts_cv.npz pins the same summation order for 4x4 * 4x1).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/defslam_b200.h"

#define GRID_COLS 64
#define GRID_ROWS 48
#define HISTO_LENGTH 30

static float fmul(float a, float b) { volatile float t = a * b; return t; }
static float fadd(float a, float b) { volatile float t = a + b; return t; }

static int descriptor_distance(const uint8_t *a, const uint8_t *b) {
  const int32_t *pa = (const int32_t *)a, *pb = (const int32_t *)b;
  int dist = 0;
  for (int i = 0; i < 8; i++, pa++, pb++) {
    unsigned int v = *pa ^ *pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

typedef struct { int *idx; int n, cap; } cell_t;

static void cell_push(cell_t *c, int v) {
  if (c->n == c->cap) { c->cap = c->cap ? 2 * c->cap : 4; c->idx = (int *)realloc(c->idx, sizeof(int) * c->cap); }
  c->idx[c->n++] = v;
}

static void three_maxima(const int *size, int L, int *ind1, int *ind2, int *ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = size[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
    else if (s > max3) { max3 = s; *ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

int oracle_search_by_projection(const defslam_projsearch_problem *p, int32_t *match_out, int32_t *nmatches_out) {
  const int NL = p->n_last, NC = p->n_cur;
  cell_t *grid = (cell_t *)calloc(GRID_COLS * GRID_ROWS, sizeof(cell_t));
  uint8_t *taken = (uint8_t *)malloc(NC > 0 ? NC : 1);
  int *hist_idx = (int *)malloc(sizeof(int) * (NL > 0 ? NL : 1)), *hist_bin = (int *)malloc(sizeof(int) * (NL > 0 ? NL : 1));
  int *vIndices2 = (int *)malloc(sizeof(int) * (NC > 0 ? NC : 1));
  int nhist = 0, nmatches = 0;
  for (int j = 0; j < NC; j++) { match_out[j] = -1; taken[j] = p->cur_taken[j]; }
  /* Frame::AssignFeaturesToGrid */
  for (int j = 0; j < NC; j++) {
    const int posX = (int)roundf(fmul(fadd(p->cur_xy[2 * j], -p->min_x), p->grid_width_inv));
    const int posY = (int)roundf(fmul(fadd(p->cur_xy[2 * j + 1], -p->min_y), p->grid_height_inv));
    if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) continue;
    cell_push(&grid[posX * GRID_ROWS + posY], j);
  }
  const float *Tc = p->T_cw, *Tl = p->T_lw;
  float twc[3];
  for (int a = 0; a < 3; a++) twc[a] = -(Tc[a] * Tc[3] + Tc[4 + a] * Tc[7] + Tc[8 + a] * Tc[11]);
  const float tlc2 = Tl[8] * twc[0] + Tl[9] * twc[1] + Tl[10] * twc[2] + Tl[11];
  const int bForward = tlc2 > p->mb && !p->mono, bBackward = -tlc2 > p->mb && !p->mono;
  const float factor = 1.0f / HISTO_LENGTH;
  for (int i = 0; i < NL; i++) {
    if (!p->last_state[i]) continue;
    const float *X = &p->last_world_xyz[3 * i];
    float c[3];
    for (int a = 0; a < 3; a++) {
      float s = fmul(Tc[4 * a], X[0]);
      s = fadd(s, fmul(Tc[4 * a + 1], X[1]));
      s = fadd(s, fmul(Tc[4 * a + 2], X[2]));
      c[a] = fadd(s, Tc[4 * a + 3]);
    }
    const float xc = c[0], yc = c[1];
    const float invzc = (float)(1.0 / c[2]);
    if (invzc < 0) continue;
    const float u = fadd(fmul(fmul(p->fx, xc), invzc), p->cx);
    const float v = fadd(fmul(fmul(p->fy, yc), invzc), p->cy);
    if (u < p->min_x || u > p->max_x) continue;
    if (v < p->min_y || v > p->max_y) continue;
    const int nLastOctave = p->last_octave[i];
    if (nLastOctave < 0 || nLastOctave >= p->n_levels) continue;
    const float radius = fmul(p->th, p->scale_factors[nLastOctave]);
    int minLevel, maxLevel;
    if (bForward) { minLevel = nLastOctave; maxLevel = -1; }
    else if (bBackward) { minLevel = 0; maxLevel = nLastOctave; }
    else { minLevel = nLastOctave - 1; maxLevel = nLastOctave + 1; }
    /* Frame::GetFeaturesInArea(u, v, radius, minLevel, maxLevel) */
    int nv = 0;
    do {
      int nMinCellX = (int)floorf(fmul(fadd(fadd(u, -p->min_x), -radius), p->grid_width_inv));
      if (nMinCellX < 0) nMinCellX = 0;
      if (nMinCellX >= GRID_COLS) break;
      int nMaxCellX = (int)ceilf(fmul(fadd(fadd(u, -p->min_x), radius), p->grid_width_inv));
      if (nMaxCellX > GRID_COLS - 1) nMaxCellX = GRID_COLS - 1;
      if (nMaxCellX < 0) break;
      int nMinCellY = (int)floorf(fmul(fadd(fadd(v, -p->min_y), -radius), p->grid_height_inv));
      if (nMinCellY < 0) nMinCellY = 0;
      if (nMinCellY >= GRID_ROWS) break;
      int nMaxCellY = (int)ceilf(fmul(fadd(fadd(v, -p->min_y), radius), p->grid_height_inv));
      if (nMaxCellY > GRID_ROWS - 1) nMaxCellY = GRID_ROWS - 1;
      if (nMaxCellY < 0) break;
      const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
      for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
          const cell_t *vCell = &grid[ix * GRID_ROWS + iy];
          for (int k = 0; k < vCell->n; k++) {
            const int j = vCell->idx[k];
            if (bCheckLevels) {
              if (p->cur_octave[j] < minLevel) continue;
              if (maxLevel >= 0 && p->cur_octave[j] > maxLevel) continue;
            }
            const float distx = fadd(p->cur_xy[2 * j], -u), disty = fadd(p->cur_xy[2 * j + 1], -v);
            if (fabsf(distx) < radius && fabsf(disty) < radius) vIndices2[nv++] = j;
          }
        }
    } while (0);
    if (nv == 0) continue;
    const uint8_t *dMP = &p->last_desc[32 * (size_t)i];
    int bestDist = 256, bestIdx2 = -1;
    for (int k = 0; k < nv; k++) {
      const int i2 = vIndices2[k];
      if (taken[i2]) continue; /* mvpMapPoints[i2] && mvpMapPoints[i2]->Observations() > 0 */
      if (p->cur_uright[i2] > 0) {
        const float ur = fadd(u, -fmul(p->mbf, invzc));
        const float er = fabsf(fadd(ur, -p->cur_uright[i2]));
        if (er > radius) continue;
      }
      const int dist = descriptor_distance(dMP, &p->cur_desc[32 * (size_t)i2]);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= p->th_high) {
      match_out[bestIdx2] = i;
      taken[bestIdx2] = p->last_has_obs[i];
      nmatches++;
      if (p->check_orientation) {
        float rot = fadd(p->last_angle[i], -p->cur_angle[bestIdx2]);
        if (rot < 0.0) rot = fadd(rot, 360.0f);
        int bin = (int)roundf(fmul(rot, factor));
        if (bin == HISTO_LENGTH) bin = 0;
        hist_idx[nhist] = bestIdx2; hist_bin[nhist] = bin; nhist++;
      }
    }
  }
  if (p->check_orientation) {
    int size[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
    memset(size, 0, sizeof(size));
    for (int a = 0; a < nhist; a++) size[hist_bin[a]]++;
    three_maxima(size, HISTO_LENGTH, &ind1, &ind2, &ind3);
    for (int a = 0; a < nhist; a++)
      if (hist_bin[a] != ind1 && hist_bin[a] != ind2 && hist_bin[a] != ind3) { match_out[hist_idx[a]] = -1; nmatches--; }
  }
  for (int k = 0; k < GRID_COLS * GRID_ROWS; k++) free(grid[k].idx);
  free(grid); free(taken); free(hist_idx); free(hist_bin); free(vIndices2);
  *nmatches_out = nmatches;
  return DEFSLAM_OK;
}

/* ------------------------------------------------------------------------------------------------
 * DefORBmatcher::searchBySchwarp  (Modules/Matching/DefORBmatcher.cc:190-293), in the reference's own
 * control flow: list of keypoints with a usable map point, Warp::getEstimates (one BBS evaluation of
 * the warp, Schwarp.cc:162-233), pixel conversion, KeyFrame::IsInImage, KeyFrame::GetFeaturesInArea
 * (KeyFrame.cc:618-663) over the 64 x 48 grid of keyframe 2, best descriptor below TH_LOW among the
 * features without a map point.
 * ---------------------------------------------------------------------------------------------- */
int oracle_bbs_eval(const defslam_bbs *s, const double *ctrl, int32_t nsites, const double *u, const double *v,
                    int32_t du, int32_t dv, double *val);

int oracle_search_by_schwarp(const defslam_warpsearch_problem *p, int32_t *match12_out, int32_t *nmatches_out) {
  const int N1 = p->n1, N2 = p->n2, NC = p->bbs.nptsu * p->bbs.nptsv;
  cell_t *grid = (cell_t *)calloc(GRID_COLS * GRID_ROWS, sizeof(cell_t));
  for (int j = 0; j < N2; j++) {       /* KeyFrame grid: built like Frame::AssignFeaturesToGrid */
    const int posX = (int)roundf(fmul(fadd(p->kp2_xy[2 * j], -p->min_x), p->grid_width_inv));
    const int posY = (int)roundf(fmul(fadd(p->kp2_xy[2 * j + 1], -p->min_y), p->grid_height_inv));
    if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) continue;
    cell_push(&grid[posX * GRID_ROWS + posY], j);
  }
  int *list = (int *)malloc(sizeof(int) * (N1 > 0 ? N1 : 1)), nl = 0;
  double *u = (double *)malloc(sizeof(double) * (N1 > 0 ? N1 : 1)), *v = (double *)malloc(sizeof(double) * (N1 > 0 ? N1 : 1));
  double *val = (double *)malloc(sizeof(double) * 2 * (N1 > 0 ? N1 : 1)), *Array = (double *)malloc(sizeof(double) * 2 * NC);
  for (int i = 0; i < N1; i++) {
    match12_out[i] = -1;
    if (!p->kp1_state[i]) continue;
    list[nl] = i; u[nl] = p->kp1_norm[2 * i]; v[nl] = p->kp1_norm[2 * i + 1]; nl++;
  }
  int nmatches = 0;
  if (nl > 0) {
    for (int n = 0; n < 2; n++)
      for (int l = 0; l < NC; l++) Array[2 * l + n] = p->x[n * NC + l];
    defslam_bbs b = p->bbs;
    b.valdim = 2;
    oracle_bbs_eval(&b, Array, nl, u, v, 0, 0, val);
    for (int k = 0; k < nl; k++) {
      const int i = list[k];
      const float ex = (float)val[2 * k], ey = (float)val[2 * k + 1];
      if (ex != ex || ey != ey) continue; /* outside the spline domain */
      const float x = fadd(fmul(ex, p->fx), p->cx), y = fadd(fmul(ey, p->fy), p->cy);
      if (!(x >= p->min_x && x < p->max_x && y >= p->min_y && y < p->max_y)) continue;
      const float r = p->radius;
      int bestDist = p->th_low, bestIdx2 = -1;
      do {
        int nMinCellX = (int)floorf(fmul(fadd(fadd(x, -p->min_x), -r), p->grid_width_inv));
        if (nMinCellX < 0) nMinCellX = 0;
        if (nMinCellX >= GRID_COLS) break;
        int nMaxCellX = (int)ceilf(fmul(fadd(fadd(x, -p->min_x), r), p->grid_width_inv));
        if (nMaxCellX > GRID_COLS - 1) nMaxCellX = GRID_COLS - 1;
        if (nMaxCellX < 0) break;
        int nMinCellY = (int)floorf(fmul(fadd(fadd(y, -p->min_y), -r), p->grid_height_inv));
        if (nMinCellY < 0) nMinCellY = 0;
        if (nMinCellY >= GRID_ROWS) break;
        int nMaxCellY = (int)ceilf(fmul(fadd(fadd(y, -p->min_y), r), p->grid_height_inv));
        if (nMaxCellY > GRID_ROWS - 1) nMaxCellY = GRID_ROWS - 1;
        if (nMaxCellY < 0) break;
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
          for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const cell_t *vCell = &grid[ix * GRID_ROWS + iy];
            for (int c = 0; c < vCell->n; c++) {
              const int j = vCell->idx[c];
              const float distx = fadd(p->kp2_xy[2 * j], -x), disty = fadd(p->kp2_xy[2 * j + 1], -y);
              if (!(fabsf(distx) < r && fabsf(disty) < r)) continue;
              if (p->kp2_has_mp[j]) continue;
              const int dist = descriptor_distance(&p->kp1_desc[32 * (size_t)i], &p->kp2_desc[32 * (size_t)j]);
              if (dist < p->th_low && dist < bestDist) { bestIdx2 = j; bestDist = dist; }
            }
          }
      } while (0);
      if (bestIdx2 >= 0) { match12_out[i] = bestIdx2; nmatches++; }
    }
  }
  for (int k = 0; k < GRID_COLS * GRID_ROWS; k++) free(grid[k].idx);
  free(grid); free(list); free(u); free(v); free(val); free(Array);
  *nmatches_out = nmatches;
  return DEFSLAM_OK;
}
