/*
 * MatcherB200.h -- host side above the C ABI for the match production that feeds the SfT solve:
 *
 *   int DefORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th,
 *                                         const bool bMono)      Modules/Matching/DefORBmatcher.cc:296-451
 *
 * A template over the reference's types, written against the member names the reference body uses
 * (mvpMapPoints, mvbOutlier, mvKeys, mvKeysUn, mvuRight, mvScaleFactors, fx..mbf, mnMinX..mnMaxY,
 * mfGridElementWidthInv/HeightInv, N).  cv::Mat never crosses the adapter; the frame exposes
 *   getPoseRowMajor(float[16])            copy of mTcw
 *   descriptorRow(i) -> const uint8_t*    mDescriptors.ptr<uint8_t>(i)
 * and the map point
 *   getWorldPosXYZ(float[3])              GetWorldPos()
 *   descriptorPtr() -> const uint8_t*     GetDescriptor().ptr<uint8_t>()
 * (one-line wrappers).  On success CurrentFrame.mvpMapPoints receives the same assignments as the
 * reference loop and the return value is its nmatches; a failing C-ABI call assigns nothing and
 * returns 0 (the caller then widens the window or declares tracking lost, DefTracking.cc).
 */
#ifndef DEFSLAM_B200_MATCHER_ADAPTER_H_
#define DEFSLAM_B200_MATCHER_ADAPTER_H_

#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "../include/defslam_b200.h"

namespace defslam_b200 {

template <class Frame, class DefMapPoint>
int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono,
                       int TH_HIGH = 75, bool mbCheckOrientation = true) {
  const int NL = LastFrame.N, NC = CurrentFrame.N;
  std::vector<uint8_t> state(NL + 1, 0), has_obs(NL + 1, 0), ldesc(32 * (size_t)NL + 32, 0), taken(NC + 1, 0),
      cdesc(32 * (size_t)NC + 32, 0);
  std::vector<float> world(3 * (size_t)NL + 3, 0.f), langle(NL + 1, 0.f), cxy(2 * (size_t)NC + 2), cangle(NC + 1), uright(NC + 1);
  std::vector<int32_t> loct(NL + 1, 0), coct(NC + 1);
  for (int i = 0; i < NL; i++) {
    auto *pMP = LastFrame.mvpMapPoints[i];
    loct[i] = LastFrame.mvKeys[i].octave;
    langle[i] = LastFrame.mvKeysUn[i].angle;
    if (!pMP || LastFrame.mvbOutlier[i] || pMP->isBad()) continue;                 /* :325-331 */
    if (!static_cast<DefMapPoint *>(pMP)->getFacet()) continue;                    /* :332-333 */
    state[i] = 1;
    has_obs[i] = pMP->Observations() > 0;
    pMP->getWorldPosXYZ(&world[3 * (size_t)i]);
    std::memcpy(&ldesc[32 * (size_t)i], pMP->descriptorPtr(), 32);
  }
  for (int j = 0; j < NC; j++) {
    cxy[2 * (size_t)j] = CurrentFrame.mvKeysUn[j].pt.x; cxy[2 * (size_t)j + 1] = CurrentFrame.mvKeysUn[j].pt.y;
    coct[j] = CurrentFrame.mvKeysUn[j].octave; cangle[j] = CurrentFrame.mvKeysUn[j].angle;
    uright[j] = CurrentFrame.mvuRight[j];
    auto *q = CurrentFrame.mvpMapPoints[j];
    taken[j] = q && q->Observations() > 0;                                          /* :372-374 */
    std::memcpy(&cdesc[32 * (size_t)j], CurrentFrame.descriptorRow(j), 32);
  }
  defslam_projsearch_problem p;
  p.n_last = NL; p.n_cur = NC; p.n_levels = (int32_t)CurrentFrame.mvScaleFactors.size();
  p.last_state = state.data(); p.last_has_obs = has_obs.data(); p.last_world_xyz = world.data(); p.last_desc = ldesc.data();
  p.last_octave = loct.data(); p.last_angle = langle.data(); p.cur_xy = cxy.data(); p.cur_octave = coct.data();
  p.cur_angle = cangle.data(); p.cur_desc = cdesc.data(); p.cur_uright = uright.data(); p.cur_taken = taken.data();
  p.scale_factors = CurrentFrame.mvScaleFactors.data();
  CurrentFrame.getPoseRowMajor(p.T_cw);
  LastFrame.getPoseRowMajor(p.T_lw);
  p.fx = CurrentFrame.fx; p.fy = CurrentFrame.fy; p.cx = CurrentFrame.cx; p.cy = CurrentFrame.cy;
  p.mb = CurrentFrame.mb; p.mbf = CurrentFrame.mbf;
  p.min_x = CurrentFrame.mnMinX; p.max_x = CurrentFrame.mnMaxX; p.min_y = CurrentFrame.mnMinY; p.max_y = CurrentFrame.mnMaxY;
  p.grid_width_inv = CurrentFrame.mfGridElementWidthInv; p.grid_height_inv = CurrentFrame.mfGridElementHeightInv;
  p.th = th; p.mono = bMono; p.th_high = TH_HIGH; p.check_orientation = mbCheckOrientation;
  std::vector<int32_t> match(NC + 1, -1);
  int32_t nmatches = 0;
  if (defslam_search_by_projection(&p, match.data(), &nmatches) != DEFSLAM_OK) return 0;
  for (int j = 0; j < NC; j++)
    if (match[j] >= 0) CurrentFrame.mvpMapPoints[j] = LastFrame.mvpMapPoints[match[j]];
  return nmatches;
}

/* ---------------------------------------------------------------------------------------------------
 *   void DefORBmatcher::CalculateInitialSchwarp(KeyFrame *Kf1, KeyFrame *Kf2, vector<pair<size_t,size_t>> &vMatchedIndices,
 *                                               double (&x)[NCu*NCv*2], double lambda)   DefORBmatcher.cc:111-187
 *   int  DefORBmatcher::searchBySchwarp(KeyFrame *pKF1, KeyFrame *pKF2, double (&x)[...],
 *                                       vector<pair<size_t,size_t>> &vMatchedPairs)      DefORBmatcher.cc:190-293
 *   void DefORBmatcher::findbyWarp(KeyFrame *Kf1, KeyFrame *Kf2, vector<pair<size_t,size_t>> &vMatchedIndices,
 *                                  double (&x)[...], double lambda)                       DefORBmatcher.cc:47-71
 * against the member names of DefKeyFrame (mpKeypointNorm, mvKeysUn, mvInvLevelSigma2, umin..vmax, NCu, NCv, valdim,
 * fx, fy, cx, cy, mnMinX..mnMaxY, mfGridElementWidthInv/HeightInv, GetMapPoint, EraseMapPointMatch, addMapPoint);
 * descriptorRow(i) -> const uint8_t* stands for mDescriptors.row(i).  x is a plain double* of 2*NCu*NCv entries. */
template <class KeyFrameT>
void CalculateInitialSchwarp(KeyFrameT *KF, KeyFrameT *KF2, std::vector<std::pair<size_t, size_t>> &vMatchedIndices,
                             double *x, double lambda) {
  const size_t n = vMatchedIndices.size();
  if (n == 0) return;
  std::vector<float> kp1(2 * n), kp2(2 * n), isig(n);
  for (size_t i = 0; i < n; i++) {                                                   /* :125-137 */
    const size_t idx1 = vMatchedIndices[i].first, idx2 = vMatchedIndices[i].second;
    kp1[2 * i] = KF->mpKeypointNorm[idx1].pt.x; kp1[2 * i + 1] = KF->mpKeypointNorm[idx1].pt.y;
    kp2[2 * i] = KF2->mpKeypointNorm[idx2].pt.x; kp2[2 * i + 1] = KF2->mpKeypointNorm[idx2].pt.y;
    isig[i] = std::sqrt(KF->mvInvLevelSigma2[KF->mvKeysUn[idx1].octave]);
  }
  defslam_schwarp_problem p;
  std::memset(&p, 0, sizeof(p));
  p.bbs.umin = KF->umin; p.bbs.umax = KF->umax; p.bbs.nptsu = KF->NCu;
  p.bbs.vmin = KF->vmin; p.bbs.vmax = KF->vmax; p.bbs.nptsv = KF->NCv; p.bbs.valdim = KF->valdim;
  p.n_matches = (int32_t)n; p.kp1 = kp1.data(); p.kp2 = kp2.data(); p.inv_sigma = isig.data();
  p.lambda = lambda; p.fx = KF->fx; p.fy = KF->fy;                                   /* (KF->fx, KF->fy) :159-160 */
  p.x = x;
  std::vector<uint8_t> keep(n, 1);
  if (defslam_schwarp_initial(&p, keep.data(), nullptr) != DEFSLAM_OK) return;       /* x and the matches untouched */
  std::vector<std::pair<size_t, size_t>> kept;
  for (size_t i = 0; i < n; i++) {                                                   /* :171-186 */
    if (keep[i]) kept.push_back(vMatchedIndices[i]);
    else KF2->EraseMapPointMatch(vMatchedIndices[i].second);
  }
  vMatchedIndices.swap(kept);
}

template <class KeyFrameT>
int searchBySchwarp(KeyFrameT *dKF, KeyFrameT *dKF2, const double *x, std::vector<std::pair<size_t, size_t>> &vMatchedPairs,
                    int TH_LOW = 50) {
  const int n1 = (int)dKF->mpKeypointNorm.size(), n2 = (int)dKF2->mvKeysUn.size();
  std::vector<float> k1(2 * (size_t)n1 + 2), k2(2 * (size_t)n2 + 2);
  std::vector<uint8_t> st(n1 + 1, 0), d1(32 * (size_t)n1 + 32), has2(n2 + 1, 0), d2(32 * (size_t)n2 + 32);
  int ncand = 0;
  for (int i = 0; i < n1; i++) {                                                     /* :200-212 */
    k1[2 * (size_t)i] = dKF->mpKeypointNorm[i].pt.x; k1[2 * (size_t)i + 1] = dKF->mpKeypointNorm[i].pt.y;
    std::memcpy(&d1[32 * (size_t)i], dKF->descriptorRow(i), 32);
    auto *pMP = dKF->GetMapPoint(i);
    if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(dKF2)) continue;
    st[i] = 1;
    ncand++;
  }
  vMatchedPairs.clear();
  if (ncand < 1) return 0;                                                           /* :214-215 */
  for (int j = 0; j < n2; j++) {
    k2[2 * (size_t)j] = dKF2->mvKeysUn[j].pt.x; k2[2 * (size_t)j + 1] = dKF2->mvKeysUn[j].pt.y;
    has2[j] = dKF2->GetMapPoint(j) != nullptr;                                       /* :260-262 */
    std::memcpy(&d2[32 * (size_t)j], dKF2->descriptorRow(j), 32);
  }
  defslam_warpsearch_problem p;
  std::memset(&p, 0, sizeof(p));
  p.bbs.umin = dKF->umin; p.bbs.umax = dKF->umax; p.bbs.nptsu = dKF->NCu;
  p.bbs.vmin = dKF->vmin; p.bbs.vmax = dKF->vmax; p.bbs.nptsv = dKF->NCv; p.bbs.valdim = dKF->valdim;
  p.x = x; p.n1 = n1; p.n2 = n2;
  p.kp1_norm = k1.data(); p.kp1_state = st.data(); p.kp1_desc = d1.data();
  p.kp2_xy = k2.data(); p.kp2_has_mp = has2.data(); p.kp2_desc = d2.data();
  p.fx = dKF2->fx; p.fy = dKF2->fy; p.cx = dKF2->cx; p.cy = dKF2->cy;               /* :244-245 */
  p.min_x = dKF2->mnMinX; p.max_x = dKF2->mnMaxX; p.min_y = dKF2->mnMinY; p.max_y = dKF2->mnMaxY;
  p.grid_width_inv = dKF2->mfGridElementWidthInv; p.grid_height_inv = dKF2->mfGridElementHeightInv;
  p.radius = 2.f;                                                                    /* th = 2  :253 */
  p.th_low = TH_LOW;
  std::vector<int32_t> m12(n1 + 1, -1);
  int32_t nmatches = 0;
  if (defslam_search_by_schwarp(&p, m12.data(), &nmatches) != DEFSLAM_OK) return 0;
  vMatchedPairs.reserve(nmatches);
  for (int i = 0; i < n1; i++)                                                       /* :282-290 */
    if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));
  return nmatches;
}

template <class KeyFrameT>
void findbyWarp(KeyFrameT *Kf1, KeyFrameT *Kf2, std::vector<std::pair<size_t, size_t>> &vMatchedIndices, double *x,
                double lambda) {
  CalculateInitialSchwarp(Kf1, Kf2, vMatchedIndices, x, lambda);                     /* :54 */
  std::vector<std::pair<size_t, size_t>> vMatchedIndices2;
  searchBySchwarp(Kf1, Kf2, x, vMatchedIndices2);                                    /* :58 */
  for (size_t i = 0; i < vMatchedIndices2.size(); i++) {                             /* :60-68 */
    auto *pMP = Kf1->GetMapPoint(vMatchedIndices2[i].first);
    if (pMP) {
      pMP->AddObservation(Kf2, vMatchedIndices2[i].second);
      Kf2->addMapPoint(pMP, vMatchedIndices2[i].second);
    }
  }
  vMatchedIndices.insert(vMatchedIndices.end(), vMatchedIndices2.begin(), vMatchedIndices2.end());  /* :70-71 */
}

}  // namespace defslam_b200
#endif
