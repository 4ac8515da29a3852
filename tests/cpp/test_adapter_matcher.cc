// Compiles adapter/MatcherB200.h against mock types carrying the member names the reference body of
// DefORBmatcher::SearchByProjection uses (Frame, MapPoint), runs it on a synthetic pair of frames and
// checks the assignments against the CPU oracle fed with the same data.  Exit code 0 = pass.  Without a
// CUDA device the library must fail loudly: the adapter then assigns nothing and returns 0.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../adapter/MatcherB200.h"
#include "../../oracle/sft_oracle.h"

struct KeyPoint { struct { float x, y; } pt; int octave = 0; float angle = 0; };
struct MapPoint {
  float pos[3]; uint8_t desc[32]; bool bad = false, facet = true; int nobs = 1;
  bool isBad() const { return bad; }
  bool getFacet() const { return facet; }
  int Observations() const { return nobs; }
  void getWorldPosXYZ(float *o) const { memcpy(o, pos, sizeof(pos)); }
  const uint8_t *descriptorPtr() const { return desc; }
};
struct Frame {
  int N = 0;
  std::vector<MapPoint *> mvpMapPoints; std::vector<bool> mvbOutlier; std::vector<KeyPoint> mvKeys, mvKeysUn;
  std::vector<float> mvuRight, mvScaleFactors; std::vector<uint8_t> desc;
  float Tcw[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float fx = 435.2047f, fy = 435.2047f, cx = 367.4517f, cy = 252.2009f, mb = 0, mbf = 0;
  float mnMinX = 0, mnMaxX = 640, mnMinY = 0, mnMaxY = 480, mfGridElementWidthInv = 64.f / 640.f, mfGridElementHeightInv = 48.f / 480.f;
  void getPoseRowMajor(float *o) const { memcpy(o, Tcw, sizeof(Tcw)); }
  const uint8_t *descriptorRow(int i) const { return &desc[32 * (size_t)i]; }
};

int main() {
  const int NL = 600, NX = 200;
  std::mt19937 rng(11); std::uniform_real_distribution<float> U(0, 1); std::normal_distribution<float> G(0, 1);
  std::vector<MapPoint> mps(NL);
  Frame last, cur;
  for (int l = 0; l < 8; l++) { last.mvScaleFactors.push_back(std::pow(1.2f, l)); cur.mvScaleFactors.push_back(std::pow(1.2f, l)); }
  cur.Tcw[3] = 0.01f; cur.Tcw[7] = -0.004f;
  for (int i = 0; i < NL; i++) {
    const float px = 10 + U(rng) * 620, py = 10 + U(rng) * 460, z = 0.8f + 0.5f * U(rng);
    mps[i].pos[0] = (px - last.cx) / last.fx * z; mps[i].pos[1] = (py - last.cy) / last.fy * z; mps[i].pos[2] = z;
    for (int b = 0; b < 32; b++) mps[i].desc[b] = (uint8_t)(rng() & 255);
    mps[i].bad = U(rng) < 0.05f; mps[i].facet = U(rng) > 0.05f; mps[i].nobs = U(rng) < 0.03f ? 0 : 2;
    KeyPoint k; k.pt.x = px; k.pt.y = py; k.octave = (int)(rng() % 8); k.angle = U(rng) * 360;
    last.mvKeys.push_back(k); last.mvKeysUn.push_back(k);
    last.mvpMapPoints.push_back(U(rng) < 0.9f ? &mps[i] : nullptr); last.mvbOutlier.push_back(U(rng) < 0.05f);
    last.mvuRight.push_back(-1); last.desc.insert(last.desc.end(), mps[i].desc, mps[i].desc + 32);
  }
  last.N = NL;
  for (int i = 0; i < NL + NX; i++) {            // current frame: re-observations (bits flipped) + clutter
    KeyPoint k; uint8_t d[32];
    if (i < NL) {
      const float *X = mps[i].pos;
      const float xc = X[0] + cur.Tcw[3], yc = X[1] + cur.Tcw[7], zc = X[2];
      k.pt.x = cur.fx * xc / zc + cur.cx + 2 * G(rng); k.pt.y = cur.fy * yc / zc + cur.cy + 2 * G(rng);
      k.octave = last.mvKeys[i].octave; k.angle = std::fmod(last.mvKeysUn[i].angle - 5 + 360, 360.f);
      memcpy(d, mps[i].desc, 32);
      for (int f = 0; f < 20; f++) { const int b = rng() % 256; d[b >> 3] ^= (uint8_t)(1 << (b & 7)); }
    } else {
      k.pt.x = U(rng) * 640; k.pt.y = U(rng) * 480; k.octave = (int)(rng() % 8); k.angle = U(rng) * 360;
      for (int b = 0; b < 32; b++) d[b] = (uint8_t)(rng() & 255);
    }
    cur.mvKeys.push_back(k); cur.mvKeysUn.push_back(k); cur.mvpMapPoints.push_back(nullptr); cur.mvbOutlier.push_back(false);
    cur.mvuRight.push_back(-1); cur.desc.insert(cur.desc.end(), d, d + 32);
  }
  cur.N = NL + NX;
  const int n = defslam_b200::SearchByProjection<Frame, MapPoint>(cur, last, 15.f, true);
  if (defslam_device_count() <= 0) {
    bool untouched = n == 0;
    for (auto *q : cur.mvpMapPoints) untouched = untouched && q == nullptr;
    printf("matcher, no CUDA device: %s\n", untouched ? "untouched" : "MODIFIED");
    return untouched ? 0 : 1;
  }
  // oracle on the same data, marshalled independently
  std::vector<uint8_t> st(NL), ho(NL), tk(cur.N, 0); std::vector<float> w(3 * NL), la(NL), cxy(2 * cur.N), ca(cur.N), ur(cur.N, -1.f);
  std::vector<int32_t> lo(NL), co(cur.N), mo(cur.N); int32_t no = 0;
  for (int i = 0; i < NL; i++) {
    MapPoint *q = last.mvpMapPoints[i];
    st[i] = q && !last.mvbOutlier[i] && !q->bad && q->facet; ho[i] = q && q->nobs > 0;
    memcpy(&w[3 * i], mps[i].pos, 12); lo[i] = last.mvKeys[i].octave; la[i] = last.mvKeysUn[i].angle;
  }
  for (int j = 0; j < cur.N; j++) { cxy[2 * j] = cur.mvKeysUn[j].pt.x; cxy[2 * j + 1] = cur.mvKeysUn[j].pt.y; co[j] = cur.mvKeysUn[j].octave; ca[j] = cur.mvKeysUn[j].angle; }
  defslam_projsearch_problem p; memset(&p, 0, sizeof(p));
  p.n_last = NL; p.n_cur = cur.N; p.n_levels = 8; p.last_state = st.data(); p.last_has_obs = ho.data(); p.last_world_xyz = w.data();
  p.last_desc = last.desc.data(); p.last_octave = lo.data(); p.last_angle = la.data(); p.cur_xy = cxy.data(); p.cur_octave = co.data();
  p.cur_angle = ca.data(); p.cur_desc = cur.desc.data(); p.cur_uright = ur.data(); p.cur_taken = tk.data(); p.scale_factors = cur.mvScaleFactors.data();
  memcpy(p.T_cw, cur.Tcw, 64); memcpy(p.T_lw, last.Tcw, 64);
  p.fx = cur.fx; p.fy = cur.fy; p.cx = cur.cx; p.cy = cur.cy; p.min_x = 0; p.max_x = 640; p.min_y = 0; p.max_y = 480;
  p.grid_width_inv = cur.mfGridElementWidthInv; p.grid_height_inv = cur.mfGridElementHeightInv; p.th = 15.f; p.mono = 1; p.th_high = 75; p.check_orientation = 1;
  if (oracle_search_by_projection(&p, mo.data(), &no)) { printf("oracle failed\n"); return 1; }
  int bad = 0, right = 0;
  for (int j = 0; j < cur.N; j++) {
    MapPoint *expect = mo[j] >= 0 ? last.mvpMapPoints[mo[j]] : nullptr;
    if (cur.mvpMapPoints[j] != expect) bad++;
    if (expect && j < NL && expect == &mps[j]) right++;
  }
  printf("matcher adapter: %d matches (oracle %d), %d mismatches, %d correct re-observations\n", n, (int)no, bad, right);
  return (bad == 0 && n == no && right > NL / 2) ? 0 : 1;
}
