// Compiles adapter/TemplateB200.h (TemplateGenerator::LaplacianMeshCreate -> LaplacianMesh / TriangularMesh constructors)
// against mock types carrying the member names the reference bodies use, builds a template from a synthetic keyframe
// surface and checks nodes, Laplacian state and map-point embedding against the CPU oracle.  Exit code 0 = pass.
// Without a CUDA device the library must fail loudly: the adapter returns -1 and creates nothing.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <set>
#include <vector>

#include "../../adapter/TemplateB200.h"
#include "../../oracle/sft_oracle.h"

struct Node { double x, y, z; unsigned idx; bool boundary = false; std::map<Node *, double> weights; void setBoundary() { boundary = true; } };
struct Facet { Node *n[3]; };
struct Edge { Node *a, *b; double d; };
struct Template {
  std::vector<Node *> nodes; std::vector<Facet *> facets; std::vector<Edge *> edges; std::map<Node *, double> kappa; double median = 0;
  Node *addNode(double x, double y, double z, unsigned i) { Node *n = new Node{x, y, z, i}; nodes.push_back(n); return n; }
  Facet *addFacet(Node *a, Node *b, Node *c) { Facet *f = new Facet{{a, b, c}}; facets.push_back(f); return f; }
  Edge *addEdge(Node *a, Node *b, double d) { Edge *e = new Edge{a, b, d}; edges.push_back(e); return e; }
  void setLaplacianCoordNorm(Node *n, double k) { kappa[n] = k; }
  void setEdgeMedian(double m) { median = m; }
};
struct MapPoint { virtual ~MapPoint() {} float pos[3]; bool bad = false; bool isBad() const { return bad; } void getWorldPosXYZ(float *o) const { memcpy(o, pos, 12); } };
struct DefMapPoint : MapPoint {
  bool lastincorporasion = true; Facet *facet = nullptr; float b[3] = {0, 0, 0}; int reposed = 0;
  void SetCoordinates(float a, float bb, float c) { b[0] = a; b[1] = bb; b[2] = c; }
  void SetFacet(Facet *f) { facet = f; }
  void Repose() { reposed++; }
};
struct Surface { defslam_bbs b; std::vector<double> depth; defslam_bbs bbs() const { return b; } const double *controlDepth() const { return depth.data(); } };
struct KeyFrame { Surface *surface; float Twc[16]; void getPoseInverseRowMajor(float *o) const { memcpy(o, Twc, 64); } };

int main() {
  std::mt19937 rng(9); std::uniform_real_distribution<float> U(0, 1);
  Surface S; S.b.umin = -0.95; S.b.umax = 0.72; S.b.nptsu = 13; S.b.vmin = -0.68; S.b.vmax = 0.62; S.b.nptsv = 15; S.b.valdim = 1;
  S.depth.resize(13 * 15);
  for (int i = 0; i < 13; i++) for (int j = 0; j < 15; j++) S.depth[i * 15 + j] = 1.0 + 0.08 * std::sin(0.5 * i) * std::cos(0.4 * j);
  KeyFrame kf; kf.surface = &S;
  const float c = std::cos(0.05f), s = std::sin(0.05f);
  const float T[16] = {c, 0, s, 0.02f, 0, 1, 0, -0.01f, -s, 0, c, 0.03f, 0, 0, 0, 1}; memcpy(kf.Twc, T, 64);
  const int G = 10, n = G * G;
  // oracle: vertices -> world -> Laplacian
  std::vector<float> cam(3 * n); if (oracle_surface_vertices(&S.b, S.depth.data(), G, G, cam.data())) return 2;
  std::vector<double> X(3 * n);
  for (int i = 0; i < n; i++) for (int r = 0; r < 3; r++) { float a = T[4 * r] * cam[3 * i]; a += T[4 * r + 1] * cam[3 * i + 1]; a += T[4 * r + 2] * cam[3 * i + 2]; a += T[4 * r + 3]; X[3 * i + r] = a; }
  std::vector<int32_t> F(2 * (G - 1) * (G - 1) * 3); const int nf = oracle_regular_triangulation(G, G, F.data());
  const int R = 8; std::vector<int32_t> cnt(n), idx(n * R), ab(2 * 3 * nf); std::vector<double> w(n * R), k0(n), l0(3 * nf); std::vector<uint8_t> bd(n); int32_t ne = 0; double med = 0;
  if (oracle_mesh_laplacian(n, X.data(), nf, F.data(), R, cnt.data(), idx.data(), w.data(), bd.data(), k0.data(), &ne, ab.data(), l0.data(), &med)) return 3;
  // map points: on the facets (random barycentrics, small offset), some far away, some bad, one null
  const int NP = 300; std::vector<DefMapPoint> mps(NP); std::set<MapPoint *> msp; std::vector<float> P(3 * NP);
  for (int i = 0; i < NP; i++) {
    const int f = rng() % nf; float a = U(rng), b = U(rng); if (a + b > 1) { a = 1 - a; b = 1 - b; }
    for (int r = 0; r < 3; r++) mps[i].pos[r] = (float)((1 - a - b) * X[3 * F[3 * f] + r] + a * X[3 * F[3 * f + 1] + r] + b * X[3 * F[3 * f + 2] + r]) + (i % 11 == 0 ? 0.5f : 1e-4f * (U(rng) - 0.5f));
    mps[i].bad = i % 17 == 0;
    msp.insert(&mps[i]);
  }
  msp.insert(nullptr);
  Template tmpl;
  const int emb = defslam_b200::LaplacianMeshCreate<Template, Node, Facet, MapPoint, DefMapPoint, KeyFrame>(&tmpl, msp, &kf, G, G);
  if (defslam_device_count() <= 0) {
    const bool untouched = emb == -1 && tmpl.nodes.empty() && tmpl.facets.empty() && mps[1].facet == nullptr && mps[1].lastincorporasion;
    printf("template adapter, no CUDA device: %s\n", untouched ? "untouched" : "MODIFIED");
    return untouched ? 0 : 1;
  }
  // oracle embedding of the same (non-null, not bad) points in the set's order
  std::vector<MapPoint *> pts; for (MapPoint *p : msp) if (p && !p->isBad()) pts.push_back(p);
  const int np = (int)pts.size(); std::vector<float> Q(3 * np); for (int i = 0; i < np; i++) memcpy(&Q[3 * i], pts[i]->pos, 12);
  std::vector<int32_t> of(np), on(3 * np); std::vector<float> ob(3 * np);
  if (oracle_embed_points(n, X.data(), nf, F.data(), np, Q.data(), of.data(), on.data(), ob.data())) return 4;
  double nerr = 0, werr = 0, kerr = 0; int bdiff = 0, ediff = 0, pdiff = 0, exp_emb = 0;
  if ((int)tmpl.nodes.size() != n || (int)tmpl.facets.size() != nf || (int)tmpl.edges.size() != ne) { printf("sizes differ\n"); return 5; }
  for (int i = 0; i < n; i++) {
    Node *nd = tmpl.nodes[i];
    nerr = std::max(nerr, std::max(std::fabs(nd->x - X[3 * i]), std::max(std::fabs(nd->y - X[3 * i + 1]), std::fabs(nd->z - X[3 * i + 2]))));
    bdiff += nd->boundary != (bd[i] != 0);
    for (int k = 0; k < cnt[i]; k++) werr = std::max(werr, std::fabs(nd->weights[tmpl.nodes[idx[i * R + k]]] - w[i * R + k]));
    if (!bd[i]) kerr = std::max(kerr, std::fabs(tmpl.kappa[nd] - k0[i]));
  }
  for (int e = 0; e < ne; e++) ediff += !(tmpl.edges[e]->a == tmpl.nodes[ab[2 * e]] && tmpl.edges[e]->b == tmpl.nodes[ab[2 * e + 1]] && std::fabs(tmpl.edges[e]->d - l0[e]) < 1e-15);
  for (int i = 0; i < np; i++) {
    DefMapPoint *mp = static_cast<DefMapPoint *>(pts[i]);
    const bool in = of[i] >= 0; exp_emb += in;
    if (in) pdiff += !(mp->facet == tmpl.facets[of[i]] && mp->b[0] == ob[3 * i] && mp->b[1] == ob[3 * i + 1] && mp->b[2] == ob[3 * i + 2] && mp->reposed == 1);
    else pdiff += mp->facet != nullptr;
    pdiff += mp->lastincorporasion;
  }
  for (int i = 0; i < NP; i++) if (mps[i].bad) pdiff += !(mps[i].facet == nullptr && mps[i].lastincorporasion);
  printf("template adapter: %d nodes %d facets %d edges, node err %.2g, weight err %.2g, kappa err %.2g, boundary diff %d, edge diff %d, "
         "median %.17g (oracle %.17g), embedded %d (oracle %d), map point diff %d\n", n, nf, (int)ne, nerr, werr, kerr, bdiff, ediff, tmpl.median, med, emb, exp_emb, pdiff);
  return (nerr == 0 && werr < 1e-13 && kerr < 1e-13 && bdiff == 0 && ediff == 0 && tmpl.median == med && emb == exp_emb && pdiff == 0 && exp_emb > 100) ? 0 : 6;
}
