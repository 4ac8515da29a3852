// Compiles adapter/NrsfmB200.h against mock types carrying the member names the reference bodies use
// (DefKeyFrame, MapPoint, DiffProp, Surface), runs the three NRSfM stages the way
// DefLocalMapping::NRSfM chains them, and checks every stage against the CPU oracle fed with the same
// data.  Exit code 0 = pass.  Without a CUDA device the library must fail loudly: the adapters then
// leave the mock objects untouched.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <vector>

#include "../../adapter/NrsfmB200.h"
#include "../../oracle/sft_oracle.h"

struct Vec3f { float v[3] = {0, 0, 0}; float &operator()(int i) { return v[i]; } float operator()(int i) const { return v[i]; } };
struct KeyPoint { struct { float x, y; } pt; int octave = 0; };
struct BbsT { double umin, umax; int nptsu; double vmin, vmax; int nptsv; int valdim; };
struct Surface {
  std::vector<Vec3f> normals, pts; std::vector<bool> has; std::vector<double> ctrl; bool saved = false; double applied = 0;
  void get3DSurfacePoint(size_t i, Vec3f &x) { x = pts[i]; }
  void applyScale(double s) { applied = s; for (auto &p : pts) for (int c = 0; c < 3; c++) p(c) = (float)(p(c) * s); }
  explicit Surface(size_t n) : normals(n), pts(n), has(n, false) {}
  bool getNormalSurfacePoint(size_t i, Vec3f &N) { if (!has[i]) return false; N = normals[i]; return true; }
  void setNormalSurfacePoint(size_t i, Vec3f &N) { normals[i] = N; has[i] = true; }
  void set3DSurfacePoint(size_t i, Vec3f &x) { pts[i] = x; }
  void saveArray(const std::vector<double> &a, BbsT &) { ctrl = a; saved = true; }
};
struct KeyFrame;
struct MapPoint {
  bool bad = false; KeyFrame *ref = nullptr; std::map<KeyFrame *, size_t> obs; double covNorm[4] = {0, 0, 0, 0};
  float kfpos[3] = {0, 0, 0}; bool facet = true; bool known = true, lastincorporasion = true; float wpos[3] = {0, 0, 0};
  bool getFacet() const { return facet; }
  bool getPositionInKeyframe(KeyFrame *, float *o) { memcpy(o, kfpos, sizeof(kfpos)); return known; }
  void SetWorldPosXYZ(const float *p) { memcpy(wpos, p, sizeof(wpos)); }
  bool isBad() const { return bad; }
  KeyFrame *GetReferenceKeyFrame() { return ref; }
  size_t GetIndexInKeyFrame(KeyFrame *k) { return obs[k]; }
  void EraseObservation(KeyFrame *k) { obs.erase(k); }
};
struct KeyFrame {
  std::vector<KeyPoint> mvKeysUn; std::vector<MapPoint *> mps; float Twc[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, Tcw[16];
  void getPoseInverseRowMajor(float *o) const { memcpy(o, Twc, sizeof(Twc)); }
  void SetPoseRowMajor(const float *i) { memcpy(Tcw, i, sizeof(Tcw)); }
  MapPoint *GetMapPoint(size_t i) { return mps[i]; }
  void EraseMapPointMatch(size_t i) { mps[i] = nullptr; }
  virtual ~KeyFrame() {}
};
struct DefKeyFrame : KeyFrame {
  std::vector<KeyPoint> mpKeypointNorm; std::vector<float> mvInvLevelSigma2;
  double umin, umax, vmin, vmax, accMean = 1.0; int NCu = 13, NCv = 15, valdim = 2; float fx = 435.2047f, fy = 435.2047f;
  Surface *surface = nullptr;
};
struct DiffProp {
  std::pair<KeyFrame *, KeyFrame *> KFToKF; size_t idx1, idx2; float I1u, I1v, I2u, I2v;
  float J12a, J12b, J12c, J12d, J21a, J21b, J21c, J21d, H12uux, H12uuy, H12uvx, H12uvy, H12vvx, H12vvy;
};

// DefLocalMapping::needNewTemplate / CreateNewMapPoints through the adapter, against the oracle
static int test_new_points(bool have_device) {
  const int N = 500, rows = 480, cols = 640;
  std::mt19937 rng(4); std::uniform_real_distribution<float> U(0, 1);
  DefKeyFrame kf; Surface sf(N); kf.surface = &sf;
  std::vector<MapPoint> mps(N);
  const float T[16] = {0.98f, 0.02f, 0.f, 0.1f, -0.02f, 0.98f, 0.01f, -0.05f, 0.f, -0.01f, 0.99f, 0.2f, 0, 0, 0, 1};
  memcpy(kf.Twc, T, sizeof(T));
  std::vector<float> xy(2 * N), surf(3 * N); std::vector<uint8_t> st(N);
  for (int i = 0; i < N; i++) {
    KeyPoint k; k.pt.x = U(rng) * (cols - 1); k.pt.y = U(rng) * (rows - 1);
    const float r = U(rng);
    st[i] = r < 0.5f ? 0 : (r < 0.9f ? 1 : 2);
    if (st[i] == 1) k.pt.x *= 0.5f;   // map points cover the left half only
    kf.mvKeysUn.push_back(k); xy[2 * i] = k.pt.x; xy[2 * i + 1] = k.pt.y;
    mps[i].bad = st[i] == 2; mps[i].known = (i % 3) != 0;
    kf.mps.push_back(st[i] ? &mps[i] : nullptr);
    for (int c = 0; c < 3; c++) { sf.pts[i](c) = U(rng) + (c == 2 ? 1.f : -0.5f); surf[3 * i + c] = sf.pts[i](c); }
  }
  std::vector<std::pair<size_t, Vec3f>> made;
  auto create = [&](size_t i, const float *x) { Vec3f v; for (int c = 0; c < 3; c++) v(c) = x[c]; made.push_back({i, v}); };
  int newPoints = -7;
  const bool need = defslam_b200::needNewTemplate((KeyFrame *)&kf, rows, cols, 50, &newPoints);
  const int created = defslam_b200::CreateNewMapPoints<DefKeyFrame, KeyFrame, MapPoint, Vec3f>((KeyFrame *)&kf, rows, cols, create);
  if (!have_device) {
    bool untouched = !need && newPoints == -7 && created == -1 && made.empty();
    for (auto &m : mps) untouched = untouched && m.wpos[2] == 0.f && m.lastincorporasion;
    printf("new points, no CUDA device: %s\n", untouched ? "untouched" : "MODIFIED");
    return untouched ? 0 : 1;
  }
  defslam_newpoints_problem p; p.n_keypoints = N; p.rows = rows; p.cols = cols; p.kp_xy = xy.data(); p.kp_state = st.data();
  p.surf_xyz = surf.data(); p.T_wc = T;
  std::vector<uint8_t> act(N); std::vector<float> w(3 * N); int32_t nn = 0;
  if (oracle_new_map_points(&p, act.data(), w.data(), &nn)) { printf("oracle new points failed\n"); return 1; }
  size_t mi = 0; int bad = 0;
  for (int i = 0; i < N; i++) {
    if (act[i] == 1) {
      if (memcmp(mps[i].wpos, &w[3 * i], 12) || mps[i].lastincorporasion != mps[i].known) bad++;
    } else if (act[i] == 2) {
      if (mi >= made.size() || made[mi].first != (size_t)i || memcmp(made[mi].second.v, &w[3 * i], 12)) bad++;
      mi++;
    } else if (st[i] && mps[i].wpos[2] != 0.f) bad++;
  }
  printf("new points: %d created (oracle %d), needNewTemplate %d, %d mismatches\n", created, (int)nn, (int)need, bad);
  return (bad == 0 && created == nn && mi == made.size() && newPoints == nn && need == (nn > 50) && nn > 0) ? 0 : 1;
}

static double depth(double u, double v) { return 1.0 + 0.05 * std::sin(2.0 * u) * std::cos(2.5 * v); }

int main() {
  const int N = 400, NC = 13 * 15;
  std::mt19937 rng(9); std::uniform_real_distribution<double> U(0, 1); std::normal_distribution<double> G(0, 1);
  DefKeyFrame kf1, kf2;
  Surface s1(N), s2(N);
  kf1.surface = &s1; kf2.surface = &s2;
  std::vector<MapPoint> mps(N);
  for (DefKeyFrame *k : {&kf1, &kf2}) {
    k->mvInvLevelSigma2.resize(6);
    for (int l = 0; l < 6; l++) k->mvInvLevelSigma2[l] = (float)(1.0 / std::pow(1.2, 2 * l));
    k->umin = k->vmin = 0.75; k->umax = k->vmax = -0.75;
  }
  // small rigid motion + mild bending between the two keyframes
  const double ang = 0.06, ca = std::cos(ang), sa = std::sin(ang);
  for (int i = 0; i < N; i++) {
    const double u = -0.7 + 1.3 * U(rng), v = -0.5 + 1.0 * U(rng), d = depth(u, v);
    const double X[3] = {u * d, v * d, d + 0.01 * std::sin(3 * u)};
    const double Y[3] = {ca * X[0] + sa * X[2] + 0.03, X[1] - 0.02, -sa * X[0] + ca * X[2] + 0.01};
    KeyPoint a, b; a.octave = b.octave = i % 6;
    a.pt.x = (float)u; a.pt.y = (float)v;
    b.pt.x = (float)(Y[0] / Y[2] + G(rng) * 0.3 / 435.0); b.pt.y = (float)(Y[1] / Y[2] + G(rng) * 0.3 / 435.0);
    kf1.mpKeypointNorm.push_back(a); kf1.mvKeysUn.push_back(a); kf2.mpKeypointNorm.push_back(b); kf2.mvKeysUn.push_back(b);
    mps[i].ref = &kf1; mps[i].obs[&kf1] = i; mps[i].obs[&kf2] = i;
    kf1.mps.push_back(&mps[i]); kf2.mps.push_back(&mps[i]);
    for (auto pr : {std::make_pair(&kf1, a), std::make_pair(&kf2, b)}) {       // DefKeyFrame.cc:116-131
      DefKeyFrame *k = pr.first; const KeyPoint &p = pr.second;
      if (p.pt.x < k->umin) k->umin = p.pt.x - 0.10; if (p.pt.x > k->umax) k->umax = p.pt.x + 0.10;
      if (p.pt.y < k->vmin) k->vmin = p.pt.y - 0.10; if (p.pt.y > k->vmax) k->vmax = p.pt.y + 0.10;
    }
  }
  std::vector<std::pair<size_t, size_t>> matches;
  for (int i = 0; i < N; i++) matches.push_back({(size_t)i, (size_t)i});

  // the initial warp handed over by DefORBmatcher::findbyWarp: Warp::initialize (oracle)
  std::vector<float> kp1(2 * N), kp2(2 * N), isig(N);
  for (int i = 0; i < N; i++) {
    kp1[2 * i] = kf1.mpKeypointNorm[i].pt.x; kp1[2 * i + 1] = kf1.mpKeypointNorm[i].pt.y;
    kp2[2 * i] = kf2.mpKeypointNorm[i].pt.x; kp2[2 * i + 1] = kf2.mpKeypointNorm[i].pt.y;
    isig[i] = std::sqrt(kf1.mvInvLevelSigma2[kf1.mvKeysUn[i].octave]);
  }
  defslam_schwarp_problem sp;
  sp.bbs = defslam_b200::keyframe_bbs(&kf1, 2);
  sp.n_matches = N; sp.kp1 = kp1.data(); sp.kp2 = kp2.data(); sp.inv_sigma = isig.data();
  sp.lambda = 0.05; sp.fx = kf1.fy; sp.fy = kf1.fx; sp.px_fx = kf1.fx; sp.px_fy = kf1.fy; sp.max_iterations = 3; sp.initialize = 0;
  std::vector<double> x0(2 * NC), x(2 * NC), xo(2 * NC);
  sp.x = x0.data();
  if (oracle_schwarp_init(&sp, x0.data())) { printf("oracle init failed\n"); return 1; }
  x = x0; xo = x0;

  std::map<MapPoint *, std::vector<std::shared_ptr<DiffProp>>> db;
  std::map<MapPoint *, bool> fresh;
  const int rc = defslam_b200::calculateSchwarps<DefKeyFrame, KeyFrame, MapPoint, DiffProp>(&kf1, &kf2, matches, x.data(), 0.05, db, fresh);
  if (defslam_device_count() <= 0) {
    const bool untouched = rc != 0 && db.empty() && fresh.empty() && x == x0;
    const int rc2 = defslam_b200::ObtainK1K2<DefKeyFrame, KeyFrame, MapPoint, DiffProp, Vec3f>(db, fresh);
    const bool ok3 = !defslam_b200::estimateSurface<DefKeyFrame, KeyFrame, Vec3f, BbsT>((KeyFrame *)&kf1, 0.7) && !s1.saved &&
                     !defslam_b200::registerSurfaces<DefKeyFrame, KeyFrame, MapPoint, Vec3f>((KeyFrame *)&kf1, 0.07, true) && s1.applied == 0;
    printf("no CUDA device: rc=%d rc2=%d state %s\n", rc, rc2, untouched && ok3 ? "untouched" : "MODIFIED");
    return (untouched && ok3 ? 0 : 1) | test_new_points(false);
  }
  if (rc) { printf("calculateSchwarps rc=%d\n", rc); return 1; }

  // ---- oracle on the same data
  std::vector<float> ouv(2 * N), oJ12(4 * N), oJ21(4 * N), oH12(6 * N); std::vector<uint8_t> okeep(N);
  defslam_diffprop od; od.warp_uv = ouv.data(); od.J12 = oJ12.data(); od.J21 = oJ21.data(); od.H12 = oH12.data(); od.keep = okeep.data();
  sp.x = xo.data();
  if (oracle_schwarp_fit(&sp, &od)) { printf("oracle fit failed\n"); return 1; }
  double ex = 0, ej = 0; size_t nrec = 0;
  for (int i = 0; i < 2 * NC; i++) ex = std::fmax(ex, std::fabs(x[i] - xo[i]));
  for (int i = 0; i < N; i++) {
    if (!okeep[i]) { if (db.count(&mps[i])) { printf("record for an unlinked match\n"); return 1; } continue; }
    auto &r = db[&mps[i]]; if (r.size() != 1) { printf("missing record %d\n", i); return 1; }
    nrec++;
    ej = std::fmax(ej, std::fabs(r[0]->J12a - oJ12[4 * i])); ej = std::fmax(ej, std::fabs(r[0]->H12vvy - oH12[6 * i + 5]) * 1e-2);
    ej = std::fmax(ej, std::fabs(r[0]->J21c - oJ21[4 * i + 2]));
  }
  printf("schwarp: %zu records, max ctrl err %.3e, max record err %.3e\n", nrec, ex, ej);
  if (ex > 1e-9 || ej > 1e-5) return 1;

  // ---- normals
  std::vector<int32_t> ptr(1, 0); std::vector<float> J12, J21, H12, I1, I2, kfst, ruv; std::vector<uint8_t> fr; std::vector<double> ki; std::vector<int> who;
  for (auto &kv : fresh) {                      // same (pointer) order as the adapter iterates
    if (!kv.second) continue;
    const int i = (int)(kv.first - &mps[0]);
    for (int c = 0; c < 4; c++) { J12.push_back(oJ12[4 * i + c]); J21.push_back(oJ21[4 * i + c]); }
    for (int c = 0; c < 6; c++) H12.push_back(oH12[6 * i + c]);
    I1.push_back(kp1[2 * i]); I1.push_back(kp1[2 * i + 1]); I2.push_back(kp2[2 * i]); I2.push_back(kp2[2 * i + 1]);
    fr.push_back(1); kfst.push_back(NAN); kfst.push_back(NAN); ki.push_back(0); ki.push_back(0);
    ruv.push_back(kp1[2 * i]); ruv.push_back(kp1[2 * i + 1]); who.push_back(i); ptr.push_back((int)fr.size());
  }
  defslam_normals_problem np; np.n_points = (int)who.size(); np.pair_ptr = ptr.data(); np.J12 = J12.data(); np.J21 = J21.data();
  np.H12 = H12.data(); np.I1 = I1.data(); np.I2 = I2.data(); np.pair_from_ref = fr.data(); np.k_first = kfst.data();
  np.k_init = ki.data(); np.ref_uv = ruv.data(); np.max_iterations = 200; np.corrected_t2 = 0;
  std::vector<double> ok(2 * who.size()), ocov(4 * who.size()); std::vector<float> onrm(3 * who.size()), opn(3 * who.size());
  std::vector<uint8_t> ost(who.size()), opv(who.size()); std::vector<int32_t> oit(who.size());
  oracle_normals_batched(&np, ok.data(), ocov.data(), onrm.data(), ost.data(), oit.data(), opn.data(), opv.data());
  if (defslam_b200::ObtainK1K2<DefKeyFrame, KeyFrame, MapPoint, DiffProp, Vec3f>(db, fresh)) { printf("ObtainK1K2 failed\n"); return 1; }
  double en = 0; int nn = 0;
  for (size_t q = 0; q < who.size(); q++) {
    const int i = who[q];
    if (ost[q] != 1) { if (s1.has[i]) { printf("normal without estimate\n"); return 1; } continue; }
    if (!s1.has[i] || !s2.has[i]) { printf("missing normal %d\n", i); return 1; }
    if (std::fabs(ocov[4 * q]) > 1e6 || oit[q] >= 200) continue;   // ill-conditioned point
    nn++;
    for (int c = 0; c < 3; c++) { en = std::fmax(en, std::fabs(s1.normals[i](c) - onrm[3 * q + c])); en = std::fmax(en, std::fabs(s2.normals[i](c) - opn[3 * q + c])); }
  }
  for (auto &kv : fresh) if (kv.second) { printf("toProccess flag not cleared\n"); return 1; }
  printf("normals: %d compared, max err %.3e\n", nn, en);
  if (en > 1e-5 || nn < N / 2) return 1;

  // ---- shape from normals
  if (!defslam_b200::estimateSurface<DefKeyFrame, KeyFrame, Vec3f, BbsT>((KeyFrame *)&kf1, 0.7) || !s1.saved) { printf("estimateSurface failed\n"); return 1; }
  std::vector<float> uv, nr, all(2 * N);
  for (int i = 0; i < N; i++) {
    all[2 * i] = kp1[2 * i]; all[2 * i + 1] = kp1[2 * i + 1];
    if (!s1.has[i]) continue;
    uv.push_back(kp1[2 * i]); uv.push_back(kp1[2 * i + 1]);
    for (int c = 0; c < 3; c++) nr.push_back(s1.normals[i](c));
  }
  defslam_sfn_problem fp; fp.bbs = defslam_b200::keyframe_bbs(&kf1, 1); fp.n_normals = (int)uv.size() / 2; fp.uv = uv.data(); fp.normals = nr.data();
  fp.bending = 0.7; fp.mean_depth = 1.0; fp.n_eval = N; fp.eval_uv = all.data();
  std::vector<double> octrl(NC); std::vector<float> oxyz(3 * N); fp.ctrl_out = octrl.data(); fp.xyz_out = oxyz.data();
  if (oracle_sfn_solve(&fp)) { printf("oracle sfn failed\n"); return 1; }
  double ec = 0, ep = 0;
  for (int i = 0; i < NC; i++) ec = std::fmax(ec, std::fabs(octrl[i] - s1.ctrl[i]));
  for (int i = 0; i < N; i++) for (int c = 0; c < 3; c++) ep = std::fmax(ep, std::fabs(oxyz[3 * i + c] - s1.pts[i](c)));
  printf("sfn: max ctrl err %.3e, max point err %.3e\n", ec, ep);
  if (ec > 1e-7 || ep > 1e-5) return 1;
  // ---- Sim(3) registration: stored map points = 1.3 x the estimated surface, shifted
  for (int i = 0; i < N; i++) for (int c = 0; c < 3; c++) mps[i].kfpos[c] = 1.3f * s1.pts[i](c) + (c == 2 ? 0.02f : 0.f);
  const std::vector<Vec3f> before = s1.pts;
  if (!defslam_b200::registerSurfaces<DefKeyFrame, KeyFrame, MapPoint, Vec3f>((KeyFrame *)&kf1, 0.07, true)) { printf("registerSurfaces failed\n"); return 1; }
  printf("registration: surface scaled by %.6f, camera moved to (%.4f %.4f %.4f)\n", s1.applied, kf1.Tcw[3], kf1.Tcw[7], kf1.Tcw[11]);
  if (std::fabs(s1.applied - 1.3) > 1e-3 || std::fabs(kf1.Tcw[11] + 0.02) > 2e-3) return 1;
  if (test_new_points(true)) return 1;
  printf("nrsfm adapter ok\n");
  return 0;
}
