/*
 * TemplateB200.h -- host side above the C ABI for the template (re)build,
 *
 *   Template *TemplateGenerator::LaplacianMeshCreate(std::set<MapPoint *> &mspMapPoints, Map *map, KeyFrame *kF)
 *     Modules/Template/TemplateGenerator.h:49-50  ->  new LaplacianMesh(mspMapPoints, map, kF)  LaplacianMesh.cc:38-43
 *     -> TriangularMesh(mspMapPoints, map, kf)  TriangularMesh.cc:57-89  ->  ExtractMeanCurvatures  LaplacianMesh.cc:53-148
 *
 * written against the member names those bodies use, as templates over the reference's types (or the mock types
 * of tests/cpp/test_adapter_template.cc).  What runs where, line for line with the reference:
 *   TriangularMesh.cc:62-65   RefSurface->getVertex(NodesSurface, 10, 10)          -> defslam_surface_vertices (GPU)
 *   :66                       regularTriangulation(nodVer, nodHor)                 -> index arithmetic, here
 *   :71-83                    x3wh = Twc * NodesSurface[i] (CV_32F)                -> here, fp32 left-to-right like cv::gemm
 *   :84                       setNodes(vertexW, facets): Node / Facet / Edge objects -> tmpl->addNode / addFacet / addEdge
 *   :85                       calculateFeaturesCoordinates()                       -> defslam_embed_points (GPU),
 *                             SetCoordinates / SetFacet / Repose per map point
 *   LaplacianMesh.cc:53-148   ExtractMeanCurvatures: weights, boundary flags, Laplacian coordinates
 *                                                                                  -> defslam_mesh_laplacian (GPU)
 * The surface exposes   bbs() -> defslam_bbs   and   controlDepth() -> const double* (Surface::bbs, nodesDepth_);
 * the keyframe          getPoseInverseRowMajor(float[16])  (GetPoseInverse(), CV_32F 4x4).
 * The template object receives   addNode(x,y,z,index) -> Node*,  addFacet(Node*,Node*,Node*) -> Facet*,
 * addEdge(Node*,Node*,dist) -> Edge*,  setLaplacianCoordNorm(Node*, kappa0),  setEdgeMedian(m)
 * and each node   weights[Node*] = w,  setBoundary()  -- the state LaplacianMesh leaves behind.
 * Returns the number of map points embedded, or -1 when a C-ABI call fails (nothing is modified then).
 */
#ifndef DEFSLAM_B200_TEMPLATE_ADAPTER_H_
#define DEFSLAM_B200_TEMPLATE_ADAPTER_H_

#include <cstdint>
#include <set>
#include <vector>

#include "../include/defslam_b200.h"

namespace defslam_b200 {

template <class TemplateT, class NodeT, class FacetT, class MapPointT, class DefMapPointT, class KeyFrameT>
int LaplacianMeshCreate(TemplateT *tmpl, std::set<MapPointT *> &mspMapPoints, KeyFrameT *kf, int nodVer = 10,
                        int nodHor = 10) {
  /* ---- everything is computed first; the objects are touched only when every call succeeded */
  const int n = nodVer * nodHor;
  const defslam_bbs bbs = kf->surface->bbs();
  std::vector<float> cam(3 * (size_t)n);
  if (defslam_surface_vertices(&bbs, kf->surface->controlDepth(), nodVer, nodHor, cam.data()) != DEFSLAM_OK) return -1;
  float Twc[16];
  kf->getPoseInverseRowMajor(Twc);
  std::vector<double> X(3 * (size_t)n);
  for (int i = 0; i < n; i++)
    for (int r = 0; r < 3; r++) {  /* (Twc * [x y z 1]')(r) in fp32, summed left to right */
      float s = Twc[4 * r] * cam[3 * i];
      s += Twc[4 * r + 1] * cam[3 * i + 1];
      s += Twc[4 * r + 2] * cam[3 * i + 2];
      s += Twc[4 * r + 3] * 1.f;
      X[3 * (size_t)i + r] = (double)s;  /* Node(x,y,z) takes the floats as doubles */
    }
  std::vector<int32_t> F;                  /* regularTriangulation  TriangularMesh.cc:92-107 */
  for (int j = 0; j < nodHor - 1; j++)
    for (int i = 0; i < nodVer - 1; i++) {
      const int f1[3] = {i + nodHor * j, i + nodHor * j + 1, (nodHor * (j + 1)) + i};
      const int f2[3] = {i + nodHor * j + 1, (nodHor * (j + 1)) + i, (nodHor * (j + 1)) + i + 1};
      F.insert(F.end(), f1, f1 + 3);
      F.insert(F.end(), f2, f2 + 3);
    }
  const int nf = (int)F.size() / 3, R = 8;
  std::vector<int32_t> cnt(n), idx((size_t)n * R), ab(2 * 3 * (size_t)nf);
  std::vector<double> w((size_t)n * R), k0(n), l0(3 * (size_t)nf);
  std::vector<uint8_t> bd(n);
  int32_t ne = 0;
  double med = 0.0;
  if (defslam_mesh_laplacian(n, X.data(), nf, F.data(), R, cnt.data(), idx.data(), w.data(), bd.data(), k0.data(), &ne,
                             ab.data(), l0.data(), &med) != DEFSLAM_OK)
    return -1;
  std::vector<MapPointT *> pts;            /* calculateFeaturesCoordinates: non-null, not bad  :139-143 */
  for (MapPointT *p : mspMapPoints)
    if (p && !p->isBad()) pts.push_back(p);
  const int np = (int)pts.size();
  std::vector<float> P(3 * (size_t)np + 3);
  for (int i = 0; i < np; i++) pts[i]->getWorldPosXYZ(&P[3 * (size_t)i]);
  std::vector<int32_t> pf(np + 1), pn(3 * (size_t)np + 3);
  std::vector<float> pb(3 * (size_t)np + 3);
  if (np > 0 && defslam_embed_points(n, X.data(), nf, F.data(), np, P.data(), pf.data(), pn.data(), pb.data()) != DEFSLAM_OK)
    return -1;
  /* ---- setNodes :84 + the state ExtractMeanCurvatures leaves on the nodes */
  std::vector<NodeT *> nodes(n);
  for (int i = 0; i < n; i++) nodes[i] = tmpl->addNode(X[3 * (size_t)i], X[3 * (size_t)i + 1], X[3 * (size_t)i + 2], (unsigned)i);
  std::vector<FacetT *> facets(nf);
  for (int f = 0; f < nf; f++) facets[f] = tmpl->addFacet(nodes[F[3 * f]], nodes[F[3 * f + 1]], nodes[F[3 * f + 2]]);
  for (int e = 0; e < ne; e++) tmpl->addEdge(nodes[ab[2 * e]], nodes[ab[2 * e + 1]], l0[e]);
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < cnt[i]; k++) nodes[i]->weights[nodes[idx[(size_t)i * R + k]]] = w[(size_t)i * R + k];
    if (bd[i]) nodes[i]->setBoundary();
    else tmpl->setLaplacianCoordNorm(nodes[i], k0[i]);
  }
  tmpl->setEdgeMedian(med);
  int embedded = 0;
  for (int i = 0; i < np; i++) {
    DefMapPointT *mp = static_cast<DefMapPointT *>(pts[i]);
    mp->lastincorporasion = false;                                                   /* :145 */
    if (pf[i] < 0) continue;
    mp->SetCoordinates(pb[3 * (size_t)i], pb[3 * (size_t)i + 1], pb[3 * (size_t)i + 2]);  /* :188-190 */
    mp->SetFacet(facets[pf[i]]);
    mp->Repose();
    embedded++;
  }
  return embedded;
}

}  // namespace defslam_b200
#endif
