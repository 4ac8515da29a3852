/*
 * NrsfmB200.h -- host side above the C ABI for the NRSfM mapping stages: header-only adapters with
 * the arguments and side effects of the three reference stages that DefLocalMapping::NRSfM chains
 * (Modules/Mapping/DefLocalMapping.cc:172-234):
 *
 *   SchwarpDatabase::calculateSchwarps(KFi, KF2i, vMatchedIndices, x, lambda)
 *                         Modules/Mapping/SchwarpDatabase.cc:145-349   -> calculateSchwarps()
 *   NormalEstimator::ObtainK1K2()
 *                         Modules/Mapping/NormalEstimator.cc:38-229    -> ObtainK1K2()
 *   ShapeFromNormals::estimate()  (+ ctor / obtainM)
 *                         Modules/Mapping/ShapeFromNormals.cc:38-260   -> estimateSurface()
 *   SurfaceRegistration::registerSurfaces()  (+ Optimizer::OptimizeHorn, scaleMinMedian)
 *                         Modules/Mapping/SurfaceRegistration.cc:48-153 -> registerSurfaces()
 *   DefLocalMapping::needNewTemplate() / CreateNewMapPoints()
 *                         Modules/Mapping/DefLocalMapping.cc:355-403,240-347 -> needNewTemplate(), CreateNewMapPoints()
 *
 * Like DefOptimizerB200.h they are templates over the reference's types, written against the member
 * names the reference bodies use, so they compile unchanged against the DefSLAM headers
 * (DefKeyFrame, MapPoint, DiffProp, Surface, WarpDatabase) and against the mock types of
 * tests/cpp/test_adapter_nrsfm.cc.  cv::Vec3f / cv::KeyPoint never cross the adapter: a Vec3f is
 * anything indexable with (i) holding floats, a keypoint anything with .pt.x/.pt.y/.octave.
 *
 * Error behaviour (mirrors the reference): a failing C-ABI call leaves every reference object
 * untouched; calculateSchwarps/ObtainK1K2 return without writing records, estimateSurface returns
 * false (DefLocalMapping.cc:210-214 then aborts NRSfM for this keyframe).
 */
#ifndef DEFSLAM_B200_NRSFM_ADAPTER_H_
#define DEFSLAM_B200_NRSFM_ADAPTER_H_

#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "../include/defslam_b200.h"

namespace defslam_b200 {

template <class DefKeyFrame>
inline defslam_bbs keyframe_bbs(const DefKeyFrame *KF, int valdim) {
  defslam_bbs b;
  b.umin = KF->umin; b.umax = KF->umax; b.nptsu = KF->NCu;
  b.vmin = KF->vmin; b.vmax = KF->vmax; b.nptsv = KF->NCv;
  b.valdim = valdim;
  return b;
}

/* SchwarpDatabase::calculateSchwarps.  `db_points` / `db_new` are WarpDatabase::mapPointsDB_ and
 * newInformation_ (the subclass passes its protected members).  x: [2*NCu*NCv] in/out, the initial
 * warp from DefORBmatcher::findbyWarp.  Returns the C-ABI code. */
template <class DefKeyFrame, class KeyFrame, class MapPoint, class DiffProp>
int calculateSchwarps(KeyFrame *KFi, KeyFrame *KF2i, std::vector<std::pair<size_t, size_t>> &vMatchedIndices, double *x,
                      double lambda, std::map<MapPoint *, std::vector<std::shared_ptr<DiffProp>>> &db_points,
                      std::map<MapPoint *, bool> &db_new) {
  DefKeyFrame *KF = static_cast<DefKeyFrame *>(KFi);
  DefKeyFrame *KF2 = static_cast<DefKeyFrame *>(KF2i);
  const size_t n = vMatchedIndices.size();
  std::vector<float> kp1(2 * n), kp2(2 * n), isig(n);
  for (size_t i = 0; i < n; i++) {                                   /* :162-183 */
    const size_t idx1 = vMatchedIndices[i].first, idx2 = vMatchedIndices[i].second;
    kp1[2 * i] = KF->mpKeypointNorm[idx1].pt.x; kp1[2 * i + 1] = KF->mpKeypointNorm[idx1].pt.y;
    kp2[2 * i] = KF2->mpKeypointNorm[idx2].pt.x; kp2[2 * i + 1] = KF2->mpKeypointNorm[idx2].pt.y;
    isig[i] = std::sqrt(KF->mvInvLevelSigma2[KF->mvKeysUn[idx1].octave]);
  }
  defslam_schwarp_problem p;
  p.bbs = keyframe_bbs(KF, KF->valdim);
  p.n_matches = (int32_t)n;
  p.kp1 = kp1.data(); p.kp2 = kp2.data(); p.inv_sigma = isig.data();
  p.lambda = lambda;
  p.fx = double(KF->fy); p.fy = double(KF->fx);                      /* the reference's argument order, :199-201 */
  p.px_fx = KF->fx; p.px_fy = KF->fy;
  p.max_iterations = 3;                                              /* :214 */
  p.initialize = 0;                                                  /* x comes from findbyWarp */
  p.x = x;
  std::vector<float> uv(2 * n), J12(4 * n), J21(4 * n), H12(6 * n);
  std::vector<uint8_t> keep(n);
  defslam_diffprop d;
  d.warp_uv = uv.data(); d.J12 = J12.data(); d.J21 = J21.data(); d.H12 = H12.data(); d.keep = keep.data();
  const int rc = defslam_schwarp_fit(&p, &d);
  if (rc != DEFSLAM_OK) return rc;
  for (size_t i = 0; i < n; i++) {                                   /* :268-345 */
    const size_t idx1 = vMatchedIndices[i].first, idx2 = vMatchedIndices[i].second;
    MapPoint *mapPoint = KF->GetMapPoint(idx1);
    MapPoint *mapPoint2 = KF2->GetMapPoint(idx2);
    if (!mapPoint || !mapPoint2) continue;
    if (mapPoint->isBad() || mapPoint2->isBad()) continue;
    if (!keep[i]) {                                                  /* > 10 px: unlink, :288-293 */
      mapPoint2->EraseObservation(KF2i);
      KF2->EraseMapPointMatch(idx2);
      continue;
    }
    if (mapPoint->GetReferenceKeyFrame() != KFi) continue;
    db_points[mapPoint].push_back(std::shared_ptr<DiffProp>(new DiffProp()));
    std::shared_ptr<DiffProp> r = db_points[mapPoint].back();
    r->KFToKF = std::pair<KeyFrame *, KeyFrame *>(KFi, KF2i);
    r->idx1 = idx1; r->idx2 = idx2;
    r->I1u = kp1[2 * i]; r->I1v = kp1[2 * i + 1]; r->I2u = kp2[2 * i]; r->I2v = kp2[2 * i + 1];
    r->J12a = J12[4 * i]; r->J12b = J12[4 * i + 1]; r->J12c = J12[4 * i + 2]; r->J12d = J12[4 * i + 3];
    r->J21a = J21[4 * i]; r->J21b = J21[4 * i + 1]; r->J21c = J21[4 * i + 2]; r->J21d = J21[4 * i + 3];
    r->H12uux = H12[6 * i]; r->H12uuy = H12[6 * i + 1]; r->H12uvx = H12[6 * i + 2];
    r->H12uvy = H12[6 * i + 3]; r->H12vvx = H12[6 * i + 4]; r->H12vvy = H12[6 * i + 5];
    db_new[mapPoint] = true;
  }
  return DEFSLAM_OK;
}

/* NormalEstimator::ObtainK1K2 over the whole database in ONE batched call.  Vec3f: cv::Vec3f. */
template <class DefKeyFrame, class KeyFrame, class MapPoint, class DiffProp, class Vec3f>
int ObtainK1K2(std::map<MapPoint *, std::vector<std::shared_ptr<DiffProp>>> &diffDB,
               std::map<MapPoint *, bool> &toProccess) {
  std::vector<MapPoint *> pts;
  std::vector<int32_t> ptr(1, 0);
  std::vector<float> J12, J21, H12, I1, I2, kfirst, refuv;
  std::vector<uint8_t> from_ref;
  std::vector<double> kinit;
  std::vector<std::shared_ptr<DiffProp>> recs;
  for (auto &process : toProccess) {                                 /* :49-75 */
    if (!process.second) continue;
    process.second = false;
    MapPoint *mp = process.first;
    if (!mp || mp->isBad()) continue;
    auto &kf2kf = diffDB[mp];
    if (kf2kf.size() < 1) continue;
    KeyFrame *refKF = mp->GetReferenceKeyFrame();
    DefKeyFrame *ref = static_cast<DefKeyFrame *>(refKF);
    size_t idx = mp->GetIndexInKeyFrame(refKF);
    for (auto &r : kf2kf) {
      const bool fr = refKF == r->KFToKF.first;
      if (fr) idx = r->idx1;
      from_ref.push_back(fr ? 1 : 0);
      const float j12[4] = {r->J12a, r->J12b, r->J12c, r->J12d}, j21[4] = {r->J21a, r->J21b, r->J21c, r->J21d};
      const float h[6] = {r->H12uux, r->H12uuy, r->H12uvx, r->H12uvy, r->H12vvx, r->H12vvy};
      J12.insert(J12.end(), j12, j12 + 4); J21.insert(J21.end(), j21, j21 + 4); H12.insert(H12.end(), h, h + 6);
      I1.push_back(r->I1u); I1.push_back(r->I1v); I2.push_back(r->I2u); I2.push_back(r->I2v);
      Vec3f Ni;                                                      /* :184-196 */
      if (!fr && static_cast<DefKeyFrame *>(r->KFToKF.first)->surface->getNormalSurfacePoint(r->idx1, Ni)) {
        kfirst.push_back(Ni(0)); kfirst.push_back(Ni(1));
      } else {
        kfirst.push_back(NAN); kfirst.push_back(NAN);
      }
      recs.push_back(r);
    }
    Vec3f N0;                                                        /* :120-131 */
    if (ref->surface->getNormalSurfacePoint(idx, N0)) { kinit.push_back(N0(0)); kinit.push_back(N0(1)); }
    else { kinit.push_back(0.0); kinit.push_back(-0.0); }
    refuv.push_back(ref->mpKeypointNorm[idx].pt.x); refuv.push_back(ref->mpKeypointNorm[idx].pt.y);
    pts.push_back(mp);
    ptr.push_back((int32_t)recs.size());
  }
  const size_t n = pts.size(), np = recs.size();
  if (n == 0) return DEFSLAM_OK;
  defslam_normals_problem p;
  p.n_points = (int32_t)n; p.pair_ptr = ptr.data();
  p.J12 = J12.data(); p.J21 = J21.data(); p.H12 = H12.data(); p.I1 = I1.data(); p.I2 = I2.data();
  p.pair_from_ref = from_ref.data(); p.k_first = kfirst.data(); p.k_init = kinit.data(); p.ref_uv = refuv.data();
  p.max_iterations = 200; p.corrected_t2 = 0;
  std::vector<double> k(2 * n), cov(4 * n);
  std::vector<float> nrm(3 * n), pn(3 * np + 3);
  std::vector<uint8_t> st(n), pv(np + 1);
  std::vector<int32_t> it(n);
  const int rc = defslam_normals_batched(&p, k.data(), cov.data(), nrm.data(), st.data(), it.data(), pn.data(), pv.data());
  if (rc != DEFSLAM_OK) return rc;
  for (size_t i = 0; i < n; i++) {
    MapPoint *mp = pts[i];
    if (st[i] == 1) {                                                /* :150-174 */
      for (int c = 0; c < 4; c++) mp->covNorm[c] = cov[4 * i + c];
      DefKeyFrame *ref = static_cast<DefKeyFrame *>(mp->GetReferenceKeyFrame());
      size_t idx = mp->GetIndexInKeyFrame(mp->GetReferenceKeyFrame());
      for (int j = ptr[i]; j < ptr[i + 1]; j++)
        if (from_ref[j]) idx = recs[j]->idx1;
      Vec3f normal;
      normal(0) = nrm[3 * i]; normal(1) = nrm[3 * i + 1]; normal(2) = nrm[3 * i + 2];
      ref->surface->setNormalSurfacePoint(idx, normal);
    }
    for (int j = ptr[i]; j < ptr[i + 1]; j++) {                      /* :176-223 */
      if (!pv[j]) continue;
      Vec3f normal_i;
      normal_i(0) = pn[3 * j]; normal_i(1) = pn[3 * j + 1]; normal_i(2) = pn[3 * j + 2];
      static_cast<DefKeyFrame *>(recs[j]->KFToKF.second)->surface->setNormalSurfacePoint(recs[j]->idx2, normal_i);
    }
  }
  return DEFSLAM_OK;
}

/* ShapeFromNormals(refKf, bendingWeight).estimate().  BbsT: BBS::bbs_t (same field order as
 * defslam_bbs), passed through to Surface::saveArray. */
template <class DefKeyFrame, class KeyFrame, class Vec3f, class BbsT>
bool estimateSurface(KeyFrame *refKf_, double bendingWeight_) {
  DefKeyFrame *kf = static_cast<DefKeyFrame *>(refKf_);
  const size_t N = refKf_->mvKeysUn.size();
  if (N == 0) return false;                                          /* :108-109 */
  std::vector<float> uv, nrm, all(2 * N);
  for (size_t var = 0; var < N; var++) {                             /* obtainM :193-213 */
    all[2 * var] = kf->mpKeypointNorm[var].pt.x; all[2 * var + 1] = kf->mpKeypointNorm[var].pt.y;
    auto mp = refKf_->GetMapPoint(var);
    if (!mp || mp->isBad()) continue;
    Vec3f Normal;
    if (!kf->surface->getNormalSurfacePoint(var, Normal)) continue;
    if (!(Normal(0) == Normal(0) && Normal(1) == Normal(1) && Normal(2) == Normal(2))) continue;
    uv.push_back(all[2 * var]); uv.push_back(all[2 * var + 1]);
    nrm.push_back(Normal(0)); nrm.push_back(Normal(1)); nrm.push_back(Normal(2));
  }
  defslam_sfn_problem p;
  p.bbs = keyframe_bbs(kf, 1);
  p.n_normals = (int32_t)(uv.size() / 2);
  p.uv = uv.data(); p.normals = nrm.data();
  p.bending = bendingWeight_; p.mean_depth = kf->accMean;
  p.n_eval = (int32_t)N; p.eval_uv = all.data();
  std::vector<double> ctrl((size_t)p.bbs.nptsu * p.bbs.nptsv);
  std::vector<float> xyz(3 * N);
  p.ctrl_out = ctrl.data(); p.xyz_out = xyz.data();
  if (defslam_sfn_solve(&p) != DEFSLAM_OK) return false;             /* "nan fail"/"inf fail" :111-122 */
  for (size_t i = 0; i < N; i++) {                                   /* :153-162 */
    Vec3f X3d;
    X3d(0) = xyz[3 * i]; X3d(1) = xyz[3 * i + 1]; X3d(2) = xyz[3 * i + 2];
    kf->surface->set3DSurfacePoint(i, X3d);
  }
  BbsT bbs;
  bbs.umin = p.bbs.umin; bbs.umax = p.bbs.umax; bbs.nptsu = p.bbs.nptsu;
  bbs.vmin = p.bbs.vmin; bbs.vmax = p.bbs.vmax; bbs.nptsv = p.bbs.nptsv; bbs.valdim = 1;
  kf->surface->saveArray(ctrl, bbs);                                 /* :164 */
  return true;
}

/* SurfaceRegistration(refKF, chiLimit, check_chi).registerSurfaces()
 * (Modules/Mapping/SurfaceRegistration.cc:48-153).  cv::Mat never crosses the adapter: the keyframe
 * exposes getPoseInverseRowMajor(float[16]) / SetPoseRowMajor(const float[16]) (one-line wrappers of
 * GetPoseInverse / SetPose) and the map point getPositionInKeyframe(KeyFrame*, float[3]) (the
 * PosesKeyframes[refKF] entry, false when empty).  `seed` replaces the reference's unseeded rand(). */
template <class DefKeyFrame, class KeyFrame, class DefMapPoint, class Vec3f>
bool registerSurfaces(KeyFrame *refKF, double chiLimit_, bool check_chi, uint64_t seed = 1) {
  DefKeyFrame *kf = static_cast<DefKeyFrame *>(refKF);
  float Twc[16];
  refKF->getPoseInverseRowMajor(Twc);
  std::vector<float> cloud1, cloud2; /* cloud1pc: stored map points; cloud2pc: surface points in the world frame */
  for (size_t i = 0; i < refKF->mvKeysUn.size(); i++) {               /* :60-103 */
    auto *pMP = refKF->GetMapPoint(i);
    if (!pMP || pMP->isBad()) continue;
    DefMapPoint *dmp = static_cast<DefMapPoint *>(pMP);
    if (!dmp->getFacet()) continue;
    float pos[3];
    if (!dmp->getPositionInKeyframe(refKF, pos)) continue;
    Vec3f x;
    kf->surface->get3DSurfacePoint(i, x);
    cloud1.insert(cloud1.end(), pos, pos + 3);
    for (int r = 0; r < 3; r++) cloud2.push_back(Twc[4 * r] * x(0) + Twc[4 * r + 1] * x(1) + Twc[4 * r + 2] * x(2) + Twc[4 * r + 3]);
  }
  const int n = (int)(cloud1.size() / 3);
  if (n < 15) return false;                                            /* :105-106 */
  float scale = 0.f;
  if (defslam_scale_min_median(n, cloud2.data(), cloud1.data(), seed, &scale) != DEFSLAM_OK) return false;
  defslam_sim3_problem p;
  p.n_points = n; p.pts1 = cloud2.data(); p.pts2 = cloud1.data();
  p.rot[0] = p.rot[1] = p.rot[2] = 0.0; p.rot[3] = 1.0;
  p.trans[0] = p.trans[1] = p.trans[2] = 0.0;
  p.scale = scale; p.chi = chiLimit_ * chiLimit_; p.huber = 0.01; p.max_iterations = 50;
  defslam_sim3_result r;
  if (defslam_sim3_register_batched(1, &p, &r, -1) != DEFSLAM_OK) return false;
  if (!r.acceptable && check_chi) return false;                        /* :129-130 */
  /* mScw = [s R | t] in fp32 (Converter::toCvMat(g2o::Sim3)); Twc' = mScw * Twc  (:132-136) */
  const double x = r.rot[0], y = r.rot[1], z = r.rot[2], w = r.rot[3];
  const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                       2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                       2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
  float S[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1}, T[16];
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++) S[4 * a + b] = (float)(r.scale * R[3 * a + b]);
    S[4 * a + 3] = (float)r.trans[a];
  }
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      float acc = 0.f;
      for (int k = 0; k < 4; k++) acc += S[4 * a + k] * Twc[4 * k + b];
      T[4 * a + b] = acc;
    }
  const double s22 = std::sqrt((double)(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]));  /* :138-140 */
  kf->surface->applyScale(s22);
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) T[4 * a + b] = (float)(T[4 * a + b] / s22);
  /* Tcw = inverse of the rigid Twc' (:142-145) */
  float Tcw[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++) Tcw[4 * a + b] = T[4 * b + a];
    Tcw[4 * a + 3] = -(T[a] * T[3] + T[4 + a] * T[7] + T[8 + a] * T[11]);
  }
  refKF->SetPoseRowMajor(Tcw);
  return true;
}

/* Keypoint table of a keyframe in the C ABI's form: pixel positions and map-point state
 * (0 none, 1 good, 2 bad) -- what both mask loops of DefLocalMapping.cc read. */
template <class KeyFrame>
inline void keyframe_keypoint_states(KeyFrame *kf, std::vector<float> &xy, std::vector<uint8_t> &state) {
  const size_t nval = kf->mvKeysUn.size();
  xy.resize(2 * nval); state.resize(nval);
  for (size_t i = 0; i < nval; i++) {
    xy[2 * i] = kf->mvKeysUn[i].pt.x; xy[2 * i + 1] = kf->mvKeysUn[i].pt.y;
    auto *pMP = kf->GetMapPoint(i);
    state[i] = !pMP ? 0 : (pMP->isBad() ? 2 : 1);
  }
}

/* DefLocalMapping::needNewTemplate (DefLocalMapping.cc:355-403): rows/cols are imGray.rows/cols.
 * A failing C-ABI call answers false (no template change), like a frame without new points. */
template <class KeyFrame>
bool needNewTemplate(KeyFrame *mpCurrentKeyFrame, int rows, int cols, int pointsToTemplate_, int *newPoints_out = nullptr) {
  std::vector<float> xy; std::vector<uint8_t> state;
  keyframe_keypoint_states(mpCurrentKeyFrame, xy, state);
  defslam_newpoints_problem p;
  p.n_keypoints = (int32_t)state.size(); p.rows = rows; p.cols = cols;
  p.kp_xy = xy.data(); p.kp_state = state.data(); p.surf_xyz = nullptr; p.T_wc = nullptr;
  std::vector<uint8_t> action(state.size() + 1);
  int32_t newPoints = 0;
  if (defslam_new_map_points(&p, action.data(), nullptr, &newPoints) != DEFSLAM_OK) return false;
  if (newPoints_out) *newPoints_out = newPoints;
  return newPoints > pointsToTemplate_;                               /* :399 */
}

/* DefLocalMapping::CreateNewMapPoints (DefLocalMapping.cc:240-347).  Existing good map points move
 * onto the surface (SetWorldPosXYZ(const float[3]), the cv::Mat-free form of SetWorldPos); for every
 * keypoint without a map point on a free pixel `create_point(i, x3w)` runs the reference's
 * object-management tail (:329-341: new DefMapPoint, AddObservation, addMapPoint,
 * ComputeDistinctiveDescriptors, UpdateNormalAndDepth, Map::addMapPoint, mlpRecentAddedMapPoints).
 * Returns the number of points created, or -1 with nothing touched when the C-ABI call fails. */
template <class DefKeyFrame, class KeyFrame, class DefMapPoint, class Vec3f, class Create>
int CreateNewMapPoints(KeyFrame *referenceKF_, int rows, int cols, Create create_point) {
  DefKeyFrame *kf = static_cast<DefKeyFrame *>(referenceKF_);
  std::vector<float> xy; std::vector<uint8_t> state;
  keyframe_keypoint_states(referenceKF_, xy, state);
  const size_t nval = state.size();
  std::vector<float> surf(3 * nval + 3, 0.f), world(3 * nval + 3);
  for (size_t i = 0; i < nval; i++) {
    if (state[i] == 2) continue;
    Vec3f x3c;
    kf->surface->get3DSurfacePoint(i, x3c);
    for (int c = 0; c < 3; c++) surf[3 * i + c] = x3c(c);
  }
  float Twc[16];
  referenceKF_->getPoseInverseRowMajor(Twc);
  defslam_newpoints_problem p;
  p.n_keypoints = (int32_t)nval; p.rows = rows; p.cols = cols;
  p.kp_xy = xy.data(); p.kp_state = state.data(); p.surf_xyz = surf.data(); p.T_wc = Twc;
  std::vector<uint8_t> action(nval + 1);
  int32_t n_new = 0;
  if (defslam_new_map_points(&p, action.data(), world.data(), &n_new) != DEFSLAM_OK) return -1;
  int created = 0;
  for (size_t i = 0; i < nval; i++) {
    if (action[i] == 1) {                                             /* :281-311 */
      DefMapPoint *defMP = static_cast<DefMapPoint *>(referenceKF_->GetMapPoint(i));
      float pos[3];
      const bool known = defMP->getPositionInKeyframe(referenceKF_, pos);
      defMP->SetWorldPosXYZ(&world[3 * i]);
      if (!known) defMP->lastincorporasion = false;
    } else if (action[i] == 2) {                                      /* :313-342 */
      create_point(i, (const float *)&world[3 * i]);
      created++;
    }
  }
  return created;
}

}  // namespace defslam_b200
#endif
