// template_ref_harness.cc -- runs the REFERENCE'S OWN LaplacianMesh::ExtractMeanCurvatures
// (Modules/Template/LaplacianMesh.cc:53-148, extracted by oracle/Makefile into oracle/_ref/laplacian_extract.inc)
// and GetMeanCurvatureInitial (:157-162) so that oracle/template_oracle.c (mean-value weights, boundary flags,
// kappa0) can be pinned against it.  TEST INFRASTRUCTURE ONLY; built into oracle/_ref/libg2o_sft_ref.so.
//
// Restated here: only the containers those lines walk -- a Node with the members they touch (Node.h:138-141:
// x,y,z, weights, NodesjJ_1J; GetNeighbours / setBoundary / isBoundary / setBadFlag as in Node.cc:114-129,200-204)
// and the edge discovery of the Facet constructor (Facet.cc:46-58).  The reference orders std::set<Node*> by
// heap address; the nodes here live in one array, so address order = index order (the oracle's convention).
#include <cmath>
#include <cstdint>
#include <map>
#include <set>
#include <vector>

#include "ref_shim/inc/mini_eigen.h"

namespace Eigen { typedef Matrix<double, 1, 1> Vector1d; }  // LaplacianMesh.h:30-33

namespace defSLAM {
struct Node {
  std::map<Node *, double> weights;
  std::map<Node *, std::pair<Node *, Node *> > NodesjJ_1J;
  double x, y, z;
  std::set<Node *> nbr;
  bool Boundary, bad;
  Node() : x(0), y(0), z(0), Boundary(false), bad(false) {}
  std::set<Node *> GetNeighbours() { return nbr; }
  void setBoundary() { Boundary = true; }
  bool isBoundary() { return Boundary; }
  void setBadFlag() { bad = true; }
};
class LaplacianMesh {
 public:
  void ExtractMeanCurvatures();
  const Eigen::Vector1d GetMeanCurvatureInitial(Node *n);
  std::set<Node *> nodes_;
  std::map<Node *, Eigen::Vector3d> LaplacianCoords;
};
#include "_ref/laplacian_extract.inc"   // LaplacianMesh.cc:53-148
#include "_ref/laplacian_kappa.inc"     // LaplacianMesh.cc:157-162
}  // namespace defSLAM

// same signature as oracle_mesh_laplacian for the outputs the reference lines produce (edges/median are not theirs)
extern "C" int ref_mesh_laplacian(int32_t n, const double *X, int32_t nf, const int32_t *facets, int32_t max_ring,
                                  int32_t *nbr_cnt, int32_t *nbr_idx, double *nbr_w, uint8_t *boundary,
                                  double *kappa0, int32_t *n_bad) {
  std::vector<defSLAM::Node> nodes(n);
  for (int i = 0; i < n; i++) { nodes[i].x = X[3 * i]; nodes[i].y = X[3 * i + 1]; nodes[i].z = X[3 * i + 2]; }
  for (int f = 0; f < nf; f++) {  // Facet.cc:46-58: an edge per side of every facet
    const int v[3] = {facets[3 * f], facets[3 * f + 1], facets[3 * f + 2]};
    const int pr[3][2] = {{v[0], v[1]}, {v[1], v[2]}, {v[0], v[2]}};
    for (int e = 0; e < 3; e++) {
      if (pr[e][0] < 0 || pr[e][0] >= n || pr[e][1] < 0 || pr[e][1] >= n) return -1;
      nodes[pr[e][0]].nbr.insert(&nodes[pr[e][1]]);
      nodes[pr[e][1]].nbr.insert(&nodes[pr[e][0]]);
    }
  }
  defSLAM::LaplacianMesh mesh;
  for (int i = 0; i < n; i++) mesh.nodes_.insert(&nodes[i]);
  mesh.ExtractMeanCurvatures();
  int bad = 0;
  for (int i = 0; i < n; i++) {
    defSLAM::Node &nd = nodes[i];
    nbr_cnt[i] = (int32_t)nd.nbr.size();
    if (nbr_cnt[i] > max_ring) return -4;
    int k = 0;
    for (std::set<defSLAM::Node *>::iterator it = nd.nbr.begin(); it != nd.nbr.end(); ++it, ++k) {
      nbr_idx[i * max_ring + k] = (int32_t)(*it - &nodes[0]);
      nbr_w[i * max_ring + k] = nd.weights.count(*it) ? nd.weights[*it] : 0.0;
    }
    for (; k < max_ring; k++) { nbr_idx[i * max_ring + k] = -1; nbr_w[i * max_ring + k] = 0.0; }
    boundary[i] = nd.Boundary ? 1 : 0;
    kappa0[i] = mesh.LaplacianCoords.count(&nd) ? mesh.GetMeanCurvatureInitial(&nd)(0) : 0.0;
    bad += nd.bad ? 1 : 0;
  }
  if (n_bad) *n_bad = bad;
  return 0;
}
