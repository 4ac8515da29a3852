/*
 * DefOptimizerB200.h -- the host side above the C ABI: a header-only adapter with the signature
 * and side effects of
 *
 *   int defSLAM::Optimizer::DefPoseOptimization(Frame *pFrame, Map *mMap, double RegLap,
 *                                               double RegInex, double RegTemp,
 *                                               unsigned int NeighboursLayers)
 *   (Modules/Tracking/DefOptimizer.h:51-53, body Modules/Tracking/DefOptimizer.cc:251-578)
 *
 * It is written against the *member names* the reference body uses, as templates, so that it
 * compiles unchanged against the real DefSLAM headers (Frame, DefMap, Template, Node, Facet,
 * DefMapPoint) and against the small mock types of tests/cpp/test_adapter.cc in this repository
 * (the reference's own headers need OpenCV/Eigen, absent from this image).
 *
 * What the adapter does, line for line with the reference body:
 *   :268-273   camera vertex <- pFrame->mTcw ; every node is a vertex            -> problem.T_cw, node_xyz
 *   :293-361   one reprojection edge per non-outlier match whose map point has a
 *              facet: facet nodes (std::set order), b1..b3, undistorted keypoint,
 *              mvInvLevelSigma2[octave] / N                                       -> match_* arrays
 *   :363-507   temporal / curvature / stretch edges from the template             -> the template plan
 *   :511-513   optimize(50)                                                       -> defslam_sft_solve
 *   :515-559   mvbOutlier, repError                                               -> outlier_out, rep_error
 *   :562-566   pFrame->SetPose(float 4x4)                                         -> T_cw_out
 *   :570-576   updateNodes ; RecalculatePosition of every map point with a facet  -> node_xyz_out + loop
 *   :577       return nInitialCorrespondences - nBad                              -> n_inliers
 *
 * The device-resident plan of the template is cached on the template object's address and
 * rebuilt when DefMap swaps the template (DefMap.cc:55-82).
 */
#ifndef DEFSLAM_B200_DEF_OPTIMIZER_H_
#define DEFSLAM_B200_DEF_OPTIMIZER_H_

#include <algorithm>
#include <cstdint>
#include <map>
#include <set>
#include <vector>

#include "../include/defslam_b200.h"

namespace defslam_b200 {

/* Flat copy of what DefOptimizer.cc reads out of a Template (Node::GetNeighbours / weights /
 * getInitialPose / isBoundary, Edge::getDist, Facet::getNodes, GetMeanCurvatureInitial,
 * getEdgeMeanSize).  Node order = Node::getIndex() order, made dense. */
struct TemplateArrays {
  std::vector<double> rest, kappa0, nbr_w, edge_len0;
  std::vector<uint8_t> boundary;
  std::vector<int32_t> nbr_ptr, nbr_idx, edge_ab, facets;
  double median = 0.0;
  defslam_template_desc desc() const {
    defslam_template_desc d;
    d.n_nodes = (int32_t)boundary.size();
    d.n_edges = (int32_t)edge_len0.size();
    d.n_facets = (int32_t)facets.size() / 3;
    d.node_rest_xyz = rest.data();
    d.node_boundary = boundary.data();
    d.nbr_ptr = nbr_ptr.data();
    d.nbr_idx = nbr_idx.data();
    d.nbr_w = nbr_w.data();
    d.node_kappa0 = kappa0.data();
    d.edge_ab = edge_ab.data();
    d.edge_len0 = edge_len0.data();
    d.facets = facets.data();
    d.edge_median_len = median;
    return d;
  }
};

template <class TemplateT, class NodeT>
struct PlanCache {
  const TemplateT *owner = nullptr;
  defslam_template *plan = nullptr;
  std::vector<NodeT *> nodes;               /* dense index -> Node*          */
  std::map<const NodeT *, int32_t> index;   /* Node* -> dense index          */
  ~PlanCache() { defslam_template_destroy(plan); }
};

/* Flatten a Template (any type with the reference's accessors). */
template <class TemplateT, class NodeT>
int extract_template(TemplateT *tmpl, TemplateArrays &A, std::vector<NodeT *> &nodes,
                     std::map<const NodeT *, int32_t> &index) {
  nodes.clear();
  index.clear();
  for (NodeT *n : tmpl->get_nodes()) nodes.push_back(n);
  /* deterministic order: the node's own index (setMeshNodes numbers vertices by it, :926-952) */
  std::sort(nodes.begin(), nodes.end(), [](NodeT *a, NodeT *b) { return a->getIndex() < b->getIndex(); });
  for (size_t i = 0; i < nodes.size(); i++) index[nodes[i]] = (int32_t)i;
  const size_t n = nodes.size();
  A = TemplateArrays();
  A.rest.resize(3 * n); A.kappa0.assign(n, 0.0); A.boundary.assign(n, 0); A.nbr_ptr.assign(n + 1, 0);
  for (size_t i = 0; i < n; i++) {
    NodeT *nd = nodes[i];
    const auto rest = nd->getInitialPose();          /* xO, yO, zO (Node.cc:193-198) */
    A.rest[3 * i] = rest[0]; A.rest[3 * i + 1] = rest[1]; A.rest[3 * i + 2] = rest[2];
    A.boundary[i] = nd->isBoundary() ? 1 : 0;
    if (!nd->isBoundary()) A.kappa0[i] = tmpl->GetMeanCurvatureInitial(nd);   /* LaplacianMesh.cc:157-162 */
    std::vector<std::pair<int32_t, double>> ring;
    for (NodeT *nb : nd->GetNeighbours()) ring.push_back({index.at(nb), nd->weights.count(nb) ? nd->weights.at(nb) : 0.0});
    std::sort(ring.begin(), ring.end());
    for (auto &r : ring) { A.nbr_idx.push_back(r.first); A.nbr_w.push_back(r.second); }
    A.nbr_ptr[i + 1] = (int32_t)A.nbr_idx.size();
  }
  for (auto *e : tmpl->get_edges()) {
    auto pr = e->get_pair_nodes();
    A.edge_ab.push_back(index.at(pr.first));
    A.edge_ab.push_back(index.at(pr.second));
    A.edge_len0.push_back(e->getDist());                /* Edge.cc:72 */
  }
  for (auto *f : tmpl->get_facets())
    for (NodeT *nd : f->getNodes()) A.facets.push_back(index.at(nd));
  A.median = tmpl->getEdgeMeanSize();                   /* the median, Template.cc:158-175 */
  return 0;
}

/* updateNodes (DefOptimizer.cc:955-968) after the roles the graph construction set on the nodes
 * (setViewed :332, setLocal :433): the reference calls Node::update(), which latches the role into the
 * viewed/local flags the drawers read (Node.cc:142-165), then resetRole(), then setXYZ. */
template <class TemplateT, class NodeT>
void write_back_nodes(PlanCache<TemplateT, NodeT> &cache, const std::vector<double> &node_out,
                      const std::vector<uint8_t> &role) {
  for (size_t i = 0; i < cache.nodes.size(); i++) {
    NodeT *nd = cache.nodes[i];
    nd->resetRole();
    if (role[i] & 1) nd->setViewed();
    else if (role[i] & 2) nd->setLocal();
    nd->update();
    nd->resetRole();
    nd->setXYZ(node_out[3 * i], node_out[3 * i + 1], node_out[3 * i + 2]);
  }
}

/* plan of the map's current template (rebuilt when DefMap swapped the template) */
template <class MapT, class TemplateT, class NodeT>
bool ensure_plan(MapT *mMap, PlanCache<TemplateT, NodeT> &cache) {
  TemplateT *tmpl = mMap->GetTemplate();
  if (!tmpl) return false;
  if (cache.owner != tmpl || !cache.plan) {
    defslam_template_destroy(cache.plan);
    cache.plan = nullptr;
    TemplateArrays A;
    extract_template<TemplateT, NodeT>(tmpl, A, cache.nodes, cache.index);
    const defslam_template_desc d = A.desc();
    if (defslam_template_create(&d, -1, &cache.plan) != DEFSLAM_OK) return false;
    cache.owner = tmpl;
  }
  return true;
}

/* The second overload,
 *   int DefPoseOptimization(const std::vector<std::vector<double>> &matches, Frame *pFrame, Map *mMap,
 *                           std::vector<bool> &outlier, double RegLap, double RegInex, double RegTemp)
 *   (Modules/Tracking/DefOptimizer.h:58-61, body DefOptimizer.cc:582-837).
 * matches[i] = {n0, n1, n2, b0, b1, b2, u, v} with n* positions in Template::nodeArray_ (:631-650).
 * The reference leaves EdgeMeanCurvature::lenghtEdge_ uninitialised in this overload (quirk C8); the
 * caller states the value here (curvEdgeLen; the median edge length is the natural choice).  Like the
 * reference: outlier[] is APPENDED to (:806-818), the pose is not written back, the return value is 0. */
template <class FrameT, class MapT, class TemplateT, class NodeT>
int DefPoseOptimization(const std::vector<std::vector<double>> &matches, FrameT *pFrame, MapT *mMap,
                        PlanCache<TemplateT, NodeT> &cache, std::vector<bool> &outlier, double RegLap,
                        double RegInex, double RegTemp, double curvEdgeLen) {
  if (!ensure_plan(mMap, cache)) return 0;
  TemplateT *tmpl = mMap->GetTemplate();
  const int n_nodes = (int)cache.nodes.size();
  std::vector<double> node_xyz(3 * (size_t)n_nodes), node_out(3 * (size_t)n_nodes);
  for (int i = 0; i < n_nodes; i++) {
    double x, y, z;
    cache.nodes[i]->getXYZ(x, y, z);
    node_xyz[3 * i] = x; node_xyz[3 * i + 1] = y; node_xyz[3 * i + 2] = z;
  }
  const int M = (int)matches.size();
  std::vector<int32_t> mnodes(3 * (size_t)M); std::vector<double> mbary(3 * (size_t)M); std::vector<float> muv(2 * (size_t)M);
  for (int i = 0; i < M; i++) {
    for (int k = 0; k < 3; k++) {
      mnodes[3 * i + k] = cache.index.at(tmpl->nodeArray_[(size_t)matches[i][k]]);   /* Nodes[matches[i][ui-1]] :640 */
      mbary[3 * i + k] = matches[i][3 + k];
    }
    muv[2 * i] = (float)matches[i][6]; muv[2 * i + 1] = (float)matches[i][7];
  }
  defslam_sft_problem p = {};
  p.tmpl = cache.plan;
  p.node_xyz = node_xyz.data();
  p.n_matches = M;
  p.match_nodes = mnodes.data(); p.match_bary = mbary.data(); p.match_uv = muv.data();
  p.fx = pFrame->fx; p.fy = pFrame->fy; p.cx = pFrame->cx; p.cy = pFrame->cy;
  pFrame->getPoseRowMajor(p.T_cw);
  p.reg_lap = RegLap; p.reg_inex = RegInex; p.reg_temp = RegTemp;
  p.max_iterations = 50;
  p.matches_given = 1;
  p.curv_edge_len = curvEdgeLen;
  std::vector<uint8_t> outl((size_t)std::max(M, 1)), role((size_t)n_nodes);
  defslam_sft_result r = {};
  r.node_xyz_out = node_out.data(); r.outlier_out = outl.data(); r.node_role_out = role.data();
  if (defslam_sft_solve(&p, &r) != DEFSLAM_OK) return 0;
  for (int m = 0; m < M; m++) outlier.push_back(outl[m] != 0);                   /* :806-818 */
  /* roles: every matched node was setViewed (:647), the curvature centres setLocal afterwards (:711) */
  for (int i = 0; i < n_nodes; i++) role[i] = (role[i] & 1) ? ((cache.nodes[i]->isBoundary()) ? 1 : 2) : 0;
  write_back_nodes(cache, node_out, role);                                       /* :822 */
  return 0;
}

/* The drop-in.  FrameT/MapT/... are the reference's types (or the test mocks). */
template <class FrameT, class MapT, class TemplateT, class NodeT, class DefMapPointT>
int DefPoseOptimization(FrameT *pFrame, MapT *mMap, PlanCache<TemplateT, NodeT> &cache, double RegLap = 5000,
                        double RegInex = 5000, double RegTemp = 0, unsigned int NeighboursLayers = 1) {
  if (!ensure_plan(mMap, cache)) return 0; /* (re)built when DefMap::createTemplate swapped the template */
  const int n_nodes = (int)cache.nodes.size();
  std::vector<double> node_xyz(3 * (size_t)n_nodes), node_out(3 * (size_t)n_nodes);
  for (int i = 0; i < n_nodes; i++) {
    double x, y, z;
    cache.nodes[i]->getXYZ(x, y, z);
    node_xyz[3 * i] = x; node_xyz[3 * i + 1] = y; node_xyz[3 * i + 2] = z;
  }
  const int N = pFrame->N;
  std::vector<int32_t> mnodes; std::vector<double> mbary; std::vector<float> muv, misig; std::vector<int> midx;
  for (int i = 0; i < N; i++) {
    if (pFrame->mvbOutlier[i]) continue;
    auto *pMP = pFrame->mvpMapPoints[i];
    if (!pMP || pMP->isBad()) continue;
    DefMapPointT *dmp = static_cast<DefMapPointT *>(pMP);
    if (!dmp->getFacet()) continue;
    for (NodeT *nd : dmp->getFacet()->getNodes()) mnodes.push_back(cache.index.at(nd));   /* std::set order */
    mbary.push_back(dmp->b1); mbary.push_back(dmp->b2); mbary.push_back(dmp->b3);
    const auto &kpUn = pFrame->mvKeysUn[i];
    muv.push_back(kpUn.pt.x); muv.push_back(kpUn.pt.y);
    misig.push_back(pFrame->mvInvLevelSigma2[kpUn.octave]);
    midx.push_back(i);
  }
  const int M = (int)midx.size();
  defslam_sft_problem p = {};
  p.tmpl = cache.plan;
  p.node_xyz = node_xyz.data();
  p.n_matches = M;
  p.n_frame_keypoints = N;
  p.match_nodes = mnodes.data(); p.match_bary = mbary.data(); p.match_uv = muv.data(); p.match_inv_sigma2 = misig.data();
  p.fx = pFrame->fx; p.fy = pFrame->fy; p.cx = pFrame->cx; p.cy = pFrame->cy;
  pFrame->getPoseRowMajor(p.T_cw);                 /* cv::Mat mTcw (CV_32F 4x4) -> 16 floats */
  p.reg_lap = RegLap; p.reg_inex = RegInex; p.reg_temp = RegTemp;
  p.neighbour_layers = (int32_t)NeighboursLayers;
  p.max_iterations = 50;
  std::vector<uint8_t> outl((size_t)std::max(M, 1)), role((size_t)n_nodes);
  defslam_sft_result r = {};
  r.node_xyz_out = node_out.data(); r.outlier_out = outl.data(); r.node_role_out = role.data();
  const int rc = defslam_sft_solve(&p, &r);
  if (rc != DEFSLAM_OK) return 0;               /* keep the previous estimate, like a failed g2o solve */
  for (int m = 0; m < M; m++) pFrame->mvbOutlier[midx[m]] = outl[m] != 0;        /* :515-537 */
  pFrame->repError = r.rep_error;                                              /* :559 */
  pFrame->SetPoseRowMajor(r.T_cw_out);                                         /* :562-566 */
  write_back_nodes(cache, node_out, role);                                     /* updateNodes :955-968 */
  for (auto *pMP : mMap->GetAllMapPoints())                                    /* :570-576 */
    if (static_cast<DefMapPointT *>(pMP)->getFacet()) static_cast<DefMapPointT *>(pMP)->RecalculatePosition();
  return r.n_inliers;                                                          /* :577 */
}

}  // namespace defslam_b200
#endif
