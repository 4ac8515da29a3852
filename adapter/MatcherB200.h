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

#include <cstdint>
#include <cstring>
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

}  // namespace defslam_b200
#endif
