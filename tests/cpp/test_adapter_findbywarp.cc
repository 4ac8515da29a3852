// Compiles the findbyWarp family of adapter/MatcherB200.h (CalculateInitialSchwarp, searchBySchwarp, findbyWarp:
// DefORBmatcher.cc:47-71,111-187,190-293) against mock keyframe types carrying the member names the reference bodies
// use, runs it on a synthetic keyframe pair and checks it against the CPU oracle fed with independently marshalled
// data.  Exit code 0 = pass.  Without a CUDA device the library must fail loudly: nothing is modified.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <set>
#include <vector>

#include "../../adapter/MatcherB200.h"
#include "../../oracle/sft_oracle.h"

struct KeyPoint { struct { float x, y; } pt; int octave = 0; };
struct KeyFrame;
struct MapPoint {
  bool bad = false; std::set<const KeyFrame *> seen; std::vector<std::pair<KeyFrame *, size_t>> added;
  bool isBad() const { return bad; }
  bool IsInKeyFrame(const KeyFrame *k) const { return seen.count(k) > 0; }
  void AddObservation(KeyFrame *k, size_t idx) { added.push_back({k, idx}); }
};
struct KeyFrame {
  std::vector<KeyPoint> mpKeypointNorm, mvKeysUn; std::vector<float> mvInvLevelSigma2; std::vector<MapPoint *> mps; std::vector<uint8_t> desc;
  double umin, umax, vmin, vmax; int NCu = 13, NCv = 15, valdim = 2;
  float fx = 435.2047f, fy = 435.2047f, cx = 367.4517f, cy = 252.2009f;
  float mnMinX = 0, mnMaxX = 640, mnMinY = 0, mnMaxY = 480, mfGridElementWidthInv = 64.f / 640.f, mfGridElementHeightInv = 48.f / 480.f;
  std::vector<size_t> erased;
  MapPoint *GetMapPoint(size_t i) const { return mps[i]; }
  void EraseMapPointMatch(size_t i) { erased.push_back(i); }
  void addMapPoint(MapPoint *p, size_t i) { mps[i] = p; }
  const uint8_t *descriptorRow(int i) const { return &desc[32 * (size_t)i]; }
};

int main() {
  const int N1 = 500, NX = 150;
  std::mt19937 rng(3); std::uniform_real_distribution<float> U(0, 1); std::normal_distribution<float> G(0, 1);
  KeyFrame k1, k2;
  k1.mvInvLevelSigma2 = {1.0f, 0.694444f, 0.482253f, 0.334898f, 0.232568f, 0.161506f}; k2.mvInvLevelSigma2 = k1.mvInvLevelSigma2;
  std::vector<MapPoint> pts(N1);
  float umin = 0.75f, umax = -0.75f, vmin = 0.75f, vmax = -0.75f;
  for (int i = 0; i < N1; i++) {
    KeyPoint px; px.pt.x = 15 + U(rng) * 610; px.pt.y = 15 + U(rng) * 450; px.octave = (int)(rng() % 6);
    KeyPoint q; q.pt.x = (px.pt.x - k1.cx) / k1.fx; q.pt.y = (px.pt.y - k1.cy) / k1.fy; q.octave = px.octave;
    k1.mvKeysUn.push_back(px); k1.mpKeypointNorm.push_back(q);
    if (q.pt.x - 0.1f < umin) umin = q.pt.x - 0.1f; if (q.pt.x + 0.1f > umax) umax = q.pt.x + 0.1f;   // DefKeyFrame.cc:116-131
    if (q.pt.y - 0.1f < vmin) vmin = q.pt.y - 0.1f; if (q.pt.y + 0.1f > vmax) vmax = q.pt.y + 0.1f;
    uint8_t d[32]; for (int b = 0; b < 32; b++) d[b] = (uint8_t)(rng() & 255);
    k1.desc.insert(k1.desc.end(), d, d + 32);
    pts[i].bad = U(rng) < 0.03f;
    k1.mps.push_back(U(rng) < 0.9f ? &pts[i] : nullptr);
    // keyframe 2: smooth displacement + noise; descriptor with a few flipped bits
    const float wu = q.pt.x + 0.02f * std::sin(2.f * q.pt.x) * std::cos(1.5f * q.pt.y) + 0.01f, wv = q.pt.y + 0.015f * std::cos(1.7f * q.pt.x) - 0.005f;
    KeyPoint p2; p2.pt.x = wu * k2.fx + k2.cx + 0.4f * G(rng); p2.pt.y = wv * k2.fy + k2.cy + 0.4f * G(rng); p2.octave = px.octave;
    KeyPoint q2; q2.pt.x = (p2.pt.x - k2.cx) / k2.fx; q2.pt.y = (p2.pt.y - k2.cy) / k2.fy;
    k2.mvKeysUn.push_back(p2); k2.mpKeypointNorm.push_back(q2);
    for (int f = 0; f < 15; f++) { const int b = rng() % 256; d[b >> 3] ^= (uint8_t)(1 << (b & 7)); }
    k2.desc.insert(k2.desc.end(), d, d + 32);
    k2.mps.push_back(nullptr);
  }
  for (int i = 0; i < NX; i++) {  // clutter in keyframe 2
    KeyPoint p2; p2.pt.x = U(rng) * 640; p2.pt.y = U(rng) * 480; KeyPoint q2; q2.pt.x = (p2.pt.x - k2.cx) / k2.fx; q2.pt.y = (p2.pt.y - k2.cy) / k2.fy;
    k2.mvKeysUn.push_back(p2); k2.mpKeypointNorm.push_back(q2); k2.mps.push_back(nullptr);
    for (int b = 0; b < 32; b++) k2.desc.push_back((uint8_t)(rng() & 255));
  }
  k1.umin = umin; k1.umax = umax; k1.vmin = vmin; k1.vmax = vmax; k2.umin = umin; k2.umax = umax; k2.vmin = vmin; k2.vmax = vmax;
  // 220 known matches (every other keypoint with a map point), 8 of them wrong; those map points are already in keyframe 2
  std::vector<std::pair<size_t, size_t>> matched;
  for (int i = 0; i < N1 && matched.size() < 220; i += 2)
    if (k1.mps[i] && !pts[i].bad) { const size_t j = matched.size() < 8 ? (size_t)((i + 37) % N1) : (size_t)i; matched.push_back({(size_t)i, j}); pts[i].seen.insert(&k2); k2.mps[j] = &pts[i]; }
  const std::vector<std::pair<size_t, size_t>> matched0 = matched;
  const std::vector<MapPoint *> mps2_0 = k2.mps;
  std::vector<double> x(2 * 13 * 15, 0.0);
  defslam_b200::findbyWarp(&k1, &k2, matched, x.data(), 0.001);
  if (defslam_device_count() <= 0) {
    bool untouched = matched == matched0 && k2.erased.empty() && k2.mps == mps2_0;
    for (double v : x) untouched = untouched && v == 0.0;
    printf("findbyWarp, no CUDA device: %s\n", untouched ? "untouched" : "MODIFIED");
    return untouched ? 0 : 1;
  }
  // ---- oracle, marshalled independently
  const size_t n = matched0.size();
  std::vector<float> a(2 * n), b(2 * n), sg(n);
  for (size_t i = 0; i < n; i++) {
    a[2 * i] = k1.mpKeypointNorm[matched0[i].first].pt.x; a[2 * i + 1] = k1.mpKeypointNorm[matched0[i].first].pt.y;
    b[2 * i] = k2.mpKeypointNorm[matched0[i].second].pt.x; b[2 * i + 1] = k2.mpKeypointNorm[matched0[i].second].pt.y;
    sg[i] = std::sqrt(k1.mvInvLevelSigma2[k1.mvKeysUn[matched0[i].first].octave]);
  }
  std::vector<double> xo(2 * 13 * 15, 0.0); std::vector<uint8_t> keep(n);
  defslam_schwarp_problem sp; memset(&sp, 0, sizeof(sp));
  sp.bbs.umin = umin; sp.bbs.umax = umax; sp.bbs.nptsu = 13; sp.bbs.vmin = vmin; sp.bbs.vmax = vmax; sp.bbs.nptsv = 15; sp.bbs.valdim = 2;
  sp.n_matches = (int)n; sp.kp1 = a.data(); sp.kp2 = b.data(); sp.inv_sigma = sg.data(); sp.lambda = 0.001; sp.fx = k1.fx; sp.fy = k1.fy; sp.x = xo.data();
  if (oracle_schwarp_initial(&sp, keep.data(), nullptr)) { printf("oracle_schwarp_initial failed\n"); return 1; }
  double xerr = 0; for (size_t i = 0; i < xo.size(); i++) xerr = std::max(xerr, std::fabs(xo[i] - x[i]));
  std::vector<std::pair<size_t, size_t>> expect; std::vector<size_t> exp_erased;
  for (size_t i = 0; i < n; i++) { if (keep[i]) expect.push_back(matched0[i]); else exp_erased.push_back(matched0[i].second); }
  // searchBySchwarp on the state CalculateInitialSchwarp left (mock EraseMapPointMatch only records)
  const int n1 = N1, n2 = N1 + NX;
  std::vector<float> k1n(2 * n1), k2p(2 * n2); std::vector<uint8_t> st(n1), has2(n2); std::vector<int32_t> m12(n1, -1); int32_t nm = 0;
  for (int i = 0; i < n1; i++) { k1n[2 * i] = k1.mpKeypointNorm[i].pt.x; k1n[2 * i + 1] = k1.mpKeypointNorm[i].pt.y; MapPoint *q = k1.mps[i]; st[i] = q && !q->bad && !q->seen.count(&k2); }
  for (int j = 0; j < n2; j++) { k2p[2 * j] = k2.mvKeysUn[j].pt.x; k2p[2 * j + 1] = k2.mvKeysUn[j].pt.y; has2[j] = mps2_0[j] != nullptr; }
  defslam_warpsearch_problem wp; memset(&wp, 0, sizeof(wp));
  wp.bbs = sp.bbs; wp.x = xo.data(); wp.n1 = n1; wp.n2 = n2; wp.kp1_norm = k1n.data(); wp.kp1_state = st.data(); wp.kp1_desc = k1.desc.data();
  wp.kp2_xy = k2p.data(); wp.kp2_has_mp = has2.data(); wp.kp2_desc = k2.desc.data(); wp.fx = k2.fx; wp.fy = k2.fy; wp.cx = k2.cx; wp.cy = k2.cy;
  wp.min_x = 0; wp.max_x = 640; wp.min_y = 0; wp.max_y = 480; wp.grid_width_inv = k2.mfGridElementWidthInv; wp.grid_height_inv = k2.mfGridElementHeightInv;
  wp.radius = 2.f; wp.th_low = 50;
  if (oracle_search_by_schwarp(&wp, m12.data(), &nm)) { printf("oracle_search_by_schwarp failed\n"); return 1; }
  for (int i = 0; i < n1; i++) if (m12[i] >= 0) expect.push_back({(size_t)i, (size_t)m12[i]});
  int added_ok = 0, right = 0;
  for (int i = 0; i < n1; i++) if (m12[i] >= 0) { added_ok += k2.mps[m12[i]] == k1.mps[i] && !pts[i].added.empty(); right += m12[i] == i; }
  const bool same = matched == expect && k2.erased == exp_erased;
  printf("findbyWarp adapter: %zu known matches -> %zu kept, %d new (oracle %d, %d correct), x err %.2g, lists %s, %d map points added\n",
         n, n - exp_erased.size(), (int)(matched.size() - (n - exp_erased.size())), (int)nm, right, xerr, same ? "identical" : "DIFFER", added_ok);
  return (same && xerr < 1e-9 && added_ok == nm && nm > 50 && !exp_erased.empty()) ? 0 : 1;
}
