// g2o_ref_harness.cc -- runs the REFERENCE'S OWN SfT code on a defslam_sft_problem so that the C oracle
// (sft_oracle.c) can be pinned against it.  TEST INFRASTRUCTURE ONLY; built by oracle/Makefile into
// oracle/_ref/libg2o_sft_ref.so when /root/reference is present (never committed, never linked by the product).
//
// Reference code that runs here (compiled where it lies, or extracted by line range into oracle/_ref/*.inc):
//   * the four SfT edges                    Thirdparty/g2o/g2o/types/sft_types.h            (verbatim #include)
//   * SE3Quat exp / map / product           Thirdparty/g2o/g2o/types/se3quat.h, se3_ops.*   (verbatim #include)
//   * quadratic forms of unary / binary / multi edges, Hessian-block mapping
//                                           Thirdparty/g2o/g2o/core/base_{unary,binary,multi}_edge.hpp (verbatim)
//   * Huber kernel                          core/robust_kernel_impl.cpp:65-91               (extracted)
//   * the Levenberg-Marquardt driver        core/optimization_algorithm_levenberg.cpp:43-189 (extracted)
//   * activeRobustChi2 / update             core/sparse_optimizer.cpp:104-120,477-491       (extracted)
//   * vertex oplus                          types_six_dof_expmap.h:73-76, types_sba.h:52-56 (extracted)
// Restated here (glue without arithmetic of its own, or arithmetic that lives in Eigen, which the image lacks):
//   * graph construction from the problem   Modules/Tracking/DefOptimizer.cc:251-513 -- the oracle's graph
//                                           (oracle_graph_build) supplies vertex ids, weights, measurements
//                                           and information values; this file turns them into reference objects
//   * BlockSolver bookkeeping               core/block_solver.hpp:143-236,502-604 (block allocation, b copy, lambda)
//   * LinearSolverDense                     solvers/linear_solver_dense.h:65-113: dense copy as written there, then
//                                           Eigen::LDLT restated (diagonal pivoting, isPositive = no negative pivot)
//   * SparseOptimizer::optimize loop        core/sparse_optimizer.cpp:403-475
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "ref_shim/inc/g2o_shim.h"
#include <types/sft_types.h>  // the reference's own file (-I<ref>/Thirdparty/g2o/g2o)

#include <types/sim3.h>       // the reference's own file (Sim3 exp / map / product)

#include "sft_oracle_graph.h"

namespace g2o {
using namespace std;

// ---- types_seven_dof_expmap.h:96-126 (VertexSim3ExpmapNoProj) and :159-188 (EdgeSim3Simple), whole classes ----
#include "_ref/vertex_sim3_noproj.inc"
#include "_ref/edge_sim3_simple.inc"

// ---- robust_kernel_impl.cpp:65-91 (setDelta, setDeltaSqr, robustify of RobustKernelHuber) ----
#include "_ref/robust_kernel_huber.inc"

// ---- things the extracted LM body refers to ----
struct G2OBatchStatistics {
  double timeResiduals, timeQuadraticForm, timeLinearSolution, timeUpdate;
  int levenbergIterations;
  static G2OBatchStatistics *globalStats() { return 0; }
};
static inline double get_monotonic_time() { return 0.0; }
template <typename T> class Property {
 public:
  explicit Property(const T &v) : _v(v) {}
  const T &value() const { return _v; }
  void setValue(const T &v) { _v = v; }
 private:
  T _v;
};
struct PropertyMap {
  template <typename P, typename T> P *makeProperty(const std::string &, const T &v) { return new P(v); }
};
template <typename T> static bool arrayHasNaN(const T *a, int n) {
  for (int i = 0; i < n; i++) if (g2o_isnan(a[i])) return true;
  return false;
}

// SparseOptimizer: containers as in sparse_optimizer.h; computeActiveErrors is the loop of
// sparse_optimizer.cpp:75-79; activeRobustChi2 and update are the reference's lines.
class SparseOptimizer {
 public:
  typedef std::vector<OptimizableGraph::Vertex *> VertexContainer;
  typedef std::vector<OptimizableGraph::Edge *> EdgeContainer;
  const VertexContainer &indexMapping() const { return _ivMap; }
  const EdgeContainer &activeEdges() const { return _activeEdges; }
  JacobianWorkspace &jacobianWorkspace() { return _jw; }
  void computeActiveErrors() {
    for (int k = 0; k < static_cast<int>(_activeEdges.size()); ++k) _activeEdges[k]->computeError();
  }
  double activeRobustChi2() const;
  void update(const double *update);
  void push() { for (size_t i = 0; i < _ivMap.size(); i++) _ivMap[i]->push(); }        // sparse_optimizer.h push(_ivMap)
  void pop() { for (size_t i = 0; i < _ivMap.size(); i++) _ivMap[i]->pop(); }
  void discardTop() { for (size_t i = 0; i < _ivMap.size(); i++) _ivMap[i]->discardTop(); }
  bool terminate() { return false; }
  VertexContainer _ivMap;
  EdgeContainer _activeEdges;
  JacobianWorkspace _jw;
};
#include "_ref/sparse_optimizer_chi2.inc"
#include "_ref/sparse_optimizer_update.inc"

// BlockSolverX + LinearSolverDense, bookkeeping restated (see header comment)
class Solver {
 public:
  explicit Solver(SparseOptimizer *o) : _optimizer(o), _size(0) {}
  SparseOptimizer *optimizer() const { return _optimizer; }
  double *x() { return _x.data(); }
  double *b() { return _b.data(); }
  size_t vectorSize() const { return _size; }
  bool buildStructure() {  // block_solver.hpp:143-236 (all free vertices are "poses": none is marginalized)
    const SparseOptimizer::VertexContainer &iv = _optimizer->indexMapping();
    int size = 0;
    _base.clear();
    for (size_t i = 0; i < iv.size(); i++) { iv[i]->setColInHessian(size); _base.push_back(size); size += iv[i]->dimension(); }
    _size = size;
    _x.assign(size, 0.0); _b.assign(size, 0.0);
    _blocks.clear();
    for (size_t i = 0; i < iv.size(); i++) iv[i]->mapHessianMemory(block((int)i, (int)i));
    const SparseOptimizer::EdgeContainer &ae = _optimizer->activeEdges();
    for (size_t k = 0; k < ae.size(); k++) {
      OptimizableGraph::Edge *e = ae[k];
      for (size_t viIdx = 0; viIdx < e->vertices().size(); ++viIdx) {
        OptimizableGraph::Vertex *v1 = (OptimizableGraph::Vertex *)e->vertex(viIdx);
        int ind1 = v1->hessianIndex();
        if (ind1 == -1) continue;
        int indexV1Bak = ind1;
        for (size_t vjIdx = viIdx + 1; vjIdx < e->vertices().size(); ++vjIdx) {
          OptimizableGraph::Vertex *v2 = (OptimizableGraph::Vertex *)e->vertex(vjIdx);
          int ind2 = v2->hessianIndex();
          if (ind2 == -1) continue;
          ind1 = indexV1Bak;
          bool transposedBlock = ind1 > ind2;
          if (transposedBlock) swap(ind1, ind2);
          e->mapHessianMemory(block(ind1, ind2), (int)viIdx, (int)vjIdx, transposedBlock);
        }
      }
    }
    return true;
  }
  bool buildSystem() {  // block_solver.hpp:502-560
    const SparseOptimizer::VertexContainer &iv = _optimizer->indexMapping();
    for (size_t i = 0; i < iv.size(); i++) iv[i]->clearQuadraticForm();
    for (std::map<std::pair<int, int>, std::vector<double> >::iterator it = _blocks.begin(); it != _blocks.end(); ++it)
      std::fill(it->second.begin(), it->second.end(), 0.0);
    JacobianWorkspace &jw = _optimizer->jacobianWorkspace();
    const SparseOptimizer::EdgeContainer &ae = _optimizer->activeEdges();
    for (size_t k = 0; k < ae.size(); k++) {
      ae[k]->linearizeOplus(jw);
      ae[k]->constructQuadraticForm();
    }
    for (size_t i = 0; i < iv.size(); i++) iv[i]->copyB(_b.data() + iv[i]->colInHessian());
    return 0;
  }
  bool setLambda(double lambda, bool backup) {  // block_solver.hpp:564-589
    const SparseOptimizer::VertexContainer &iv = _optimizer->indexMapping();
    if (backup) _diagBackup.assign(_size, 0.0);
    for (size_t i = 0; i < iv.size(); i++) {
      const int d = iv[i]->dimension();
      double *m = block((int)i, (int)i);
      for (int k = 0; k < d; k++) {
        if (backup) _diagBackup[_base[i] + k] = m[k * d + k];
        m[k * d + k] += lambda;
      }
    }
    return true;
  }
  void restoreDiagonal() {  // block_solver.hpp:591-604
    const SparseOptimizer::VertexContainer &iv = _optimizer->indexMapping();
    for (size_t i = 0; i < iv.size(); i++) {
      const int d = iv[i]->dimension();
      double *m = block((int)i, (int)i);
      for (int k = 0; k < d; k++) m[k * d + k] = _diagBackup[_base[i] + k];
    }
  }
  // linear_solver_dense.h:65-113: upper blocks copied, lower triangle mirrored, LDLT, fail unless positive
  void dense(std::vector<double> &H) const {
    const int n = (int)_size;
    const SparseOptimizer::VertexContainer &iv = _optimizer->indexMapping();
    H.assign((size_t)n * n, 0.0);
    for (std::map<std::pair<int, int>, std::vector<double> >::const_iterator it = _blocks.begin(); it != _blocks.end(); ++it) {
      const int bi = it->first.first, bj = it->first.second;
      const int r0 = _base[bi], c0 = _base[bj], rs = iv[bi]->dimension(), cs = iv[bj]->dimension();
      for (int i = 0; i < rs; i++)
        for (int j = 0; j < cs; j++) {
          const double v = it->second[(size_t)j * rs + i];  // column-major block
          H[(size_t)(r0 + i) * n + c0 + j] = v;
          if (r0 != c0) H[(size_t)(c0 + j) * n + r0 + i] = v;
        }
    }
  }
  bool solve() {
    const int n = (int)_size;
    std::vector<double> H;
    dense(H);
    return ldlt_solve(n, H, _b.data(), _x.data());
  }
  // Eigen::LDLT<MatrixXd> (Eigen/src/Cholesky/LDLT.h, unblocked, lower): at step k the largest remaining
  // |diagonal| is swapped into place; isPositive() = no negative pivot was seen; solve = P^T L^-T D^-1 L^-1 P b
  // with pivots below the smallest normalised double treated as zero.
  static bool ldlt_solve(int n, std::vector<double> &A, const double *b, double *x) {
    std::vector<int> tr(n);
    int sign = 0;  // 0 zero, +1 positive semidef, -1 negative semidef, 2 indefinite
    std::vector<double> tmp(n);
    for (int k = 0; k < n; k++) {
      int p = k;
      double big = fabs(A[(size_t)k * n + k]);
      for (int i = k + 1; i < n; i++) if (fabs(A[(size_t)i * n + i]) > big) { big = fabs(A[(size_t)i * n + i]); p = i; }
      tr[k] = p;
      if (p != k) {
        const int s = n - p - 1;
        for (int j = 0; j < k; j++) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
        for (int i = 0; i < s; i++) std::swap(A[(size_t)(p + 1 + i) * n + k], A[(size_t)(p + 1 + i) * n + p]);
        std::swap(A[(size_t)k * n + k], A[(size_t)p * n + p]);
        for (int i = k + 1; i < p; i++) std::swap(A[(size_t)i * n + k], A[(size_t)p * n + i]);
      }
      const int rs = n - k - 1;
      if (k > 0) {
        for (int j = 0; j < k; j++) tmp[j] = A[(size_t)j * n + j] * A[(size_t)k * n + j];
        double s = 0;
        for (int j = 0; j < k; j++) s += A[(size_t)k * n + j] * tmp[j];
        A[(size_t)k * n + k] -= s;
        for (int i = 0; i < rs; i++) {
          double t = 0;
          const double *Ai = &A[(size_t)(k + 1 + i) * n];
          for (int j = 0; j < k; j++) t += Ai[j] * tmp[j];
          A[(size_t)(k + 1 + i) * n + k] -= t;
        }
      }
      const double akk = A[(size_t)k * n + k];
      if (rs > 0 && fabs(akk) > 0.0) for (int i = 0; i < rs; i++) A[(size_t)(k + 1 + i) * n + k] /= akk;
      if (sign == 1) { if (akk < 0) sign = 2; }
      else if (sign == -1) { if (akk > 0) sign = 2; }
      else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
    }
    if (!(sign == 1 || sign == 0)) return false;
    std::vector<double> y(b, b + n);
    for (int k = 0; k < n; k++) std::swap(y[k], y[tr[k]]);
    for (int i = 0; i < n; i++) { double s = y[i]; for (int j = 0; j < i; j++) s -= A[(size_t)i * n + j] * y[j]; y[i] = s; }
    for (int i = 0; i < n; i++) { const double d = A[(size_t)i * n + i]; y[i] = fabs(d) > DBL_MIN ? y[i] / d : 0.0; }
    for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int j = i + 1; j < n; j++) s -= A[(size_t)j * n + i] * y[j]; y[i] = s; }
    for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[tr[k]]);
    for (int i = 0; i < n; i++) x[i] = y[i];
    return true;
  }
  double *block(int i, int j) {
    const SparseOptimizer::VertexContainer &iv = _optimizer->indexMapping();
    std::vector<double> &m = _blocks[std::make_pair(i, j)];
    if (m.empty()) m.assign((size_t)iv[i]->dimension() * iv[j]->dimension(), 0.0);
    return m.data();
  }
  SparseOptimizer *_optimizer;
  size_t _size;
  std::vector<double> _x, _b, _diagBackup;
  std::vector<int> _base;
  std::map<std::pair<int, int>, std::vector<double> > _blocks;
};

// skeletons of optimization_algorithm.h / _with_hessian.h / _levenberg.h:40-88 (members only)
class OptimizationAlgorithm {
 public:
  enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
  OptimizationAlgorithm() : _optimizer(0) {}
  virtual ~OptimizationAlgorithm() {}
  void setOptimizer(SparseOptimizer *o) { _optimizer = o; }
 protected:
  SparseOptimizer *_optimizer;
  PropertyMap _properties;
};
class OptimizationAlgorithmWithHessian : public OptimizationAlgorithm {
 public:
  explicit OptimizationAlgorithmWithHessian(Solver *solver) : _solver(solver) {}
 protected:
  Solver *_solver;
};
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithmWithHessian {
 public:
  explicit OptimizationAlgorithmLevenberg(Solver *solver);
  virtual ~OptimizationAlgorithmLevenberg();
  virtual SolverResult solve(int iteration, bool online = false);
  double currentLambda() const { return _currentLambda; }
  int levenbergIteration() { return _levenbergIterations; }
 protected:
  Property<int> *_maxTrialsAfterFailure;
  Property<double> *_userLambdaInit;
  double _currentLambda;
  double _tau;
  double _goodStepLowerScale;
  double _goodStepUpperScale;
  double _ni;
  int _levenbergIterations;
  int _nBad;
  double computeLambdaInit() const;
  double computeScale() const;
};
// ---- optimization_algorithm_levenberg.cpp:43-189 (ctor, dtor, solve, computeLambdaInit, computeScale) ----
#include "_ref/levenberg_body.inc"

}  // namespace g2o

// ------------------------------------------------------------------------------------------------
namespace {

struct RefGraph {
  g2o::SparseOptimizer opt;
  g2o::VertexSE3Expmap *cam;
  std::vector<g2o::VertexSBAPointXYZ *> nodes;
  std::vector<g2o::EdgeNodesCamera *> rep;
  std::vector<g2o::EdgesReference *> ref;
  std::vector<g2o::EdgeMeanCurvature *> curv;
  std::vector<g2o::EdgesStreching *> str;
  std::vector<g2o::RobustKernelHuber *> kernels;
  ~RefGraph() {
    for (size_t i = 0; i < rep.size(); i++) delete rep[i];
    for (size_t i = 0; i < ref.size(); i++) delete ref[i];
    for (size_t i = 0; i < curv.size(); i++) delete curv[i];
    for (size_t i = 0; i < str.size(); i++) delete str[i];
    for (size_t i = 0; i < kernels.size(); i++) delete kernels[i];
    for (size_t i = 0; i < nodes.size(); i++) delete nodes[i];
    delete cam;
  }
};

// DefOptimizer.cc:266-507 with the oracle's graph as the source of ids / weights / measurements
void build_ref_graph(RefGraph &R, const Graph &g, const defslam_sft_problem *p) {
  const int n = g.n_nodes;
  // vSE3->setEstimate(Converter::toSE3Quat(pFrame->mTcw))  :268-271 ; Converter.cc: R,t doubles from the f32 cv::Mat
  Eigen::Matrix3d Rm;
  Eigen::Vector3d tv;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) Rm(i, j) = (double)p->T_cw[i * 4 + j];
    tv(i) = (double)p->T_cw[i * 4 + 3];
  }
  R.cam = new g2o::VertexSE3Expmap();
  R.cam->setEstimate(g2o::SE3Quat(Rm, tv));
  R.cam->setId(0);
  R.cam->setFixed(false);
  // setMeshNodes :926-952, freed nodes :414-419
  R.nodes.resize(n);
  for (int v = 0; v < n; v++) {
    g2o::VertexSBAPointXYZ *vn = new g2o::VertexSBAPointXYZ();
    Eigen::Vector3d x;
    x << g.x[3 * v], g.x[3 * v + 1], g.x[3 * v + 2];
    vn->setEstimate(x);
    vn->setFixed(!g.optlap[v]);
    vn->setMarginalized(!g.optlap[v]);
    vn->setId(v + 1);
    R.nodes[v] = vn;
  }
  // initializeOptimization: free vertices in id order get hessian indices (sparse_optimizer.cpp:181-301)
  int hi = 0;
  R.cam->setHessianIndex(hi++);
  R.opt._ivMap.push_back(R.cam);
  for (int v = 0; v < n; v++)
    if (g.optlap[v]) { R.nodes[v]->setHessianIndex(hi++); R.opt._ivMap.push_back(R.nodes[v]); }

  const float deltaMono = g.variant ? 0.5 : sqrt(5.991);  // :286 ; :625 in the matches-given overload
  for (int i = 0; i < g.n_rep; i++) {   // :305-357
    const EdgeReproj &o = g.rep[i];
    Eigen::Matrix<double, 2, 1> obs;
    obs << o.obs[0], o.obs[1];
    g2o::EdgeNodesCamera *e = new g2o::EdgeNodesCamera();
    e->resize(4);
    e->setVertex(0, R.cam);
    Eigen::Vector3d bary;
    bary << o.bary[0], o.bary[1], o.bary[2];
    for (int k = 0; k < 3; k++) e->setVertex(k + 1, R.nodes[o.v[k]]);
    e->setBarycentric(bary);
    e->setMeasurement(obs);
    const float invSigma2 = g.variant ? 0.f : p->match_inv_sigma2[o.m];
    const int N = p->n_frame_keypoints;
    if (g.variant) e->setInformation(Eigen::Matrix2d::Identity() / double(p->n_matches));  // :655
    else e->setInformation(Eigen::Matrix2d::Identity() * invSigma2 / N);
    g2o::RobustKernelHuber *rk = new g2o::RobustKernelHuber;
    R.kernels.push_back(rk);
    e->setRobustKernel(rk);
    rk->setDelta(deltaMono);
    e->fx = p->fx; e->fy = p->fy; e->cx = p->cx; e->cy = p->cy;
    e->computeError();
    R.rep.push_back(e);
    R.opt._activeEdges.push_back(e);
  }
  const double m = p->tmpl_desc->edge_median_len;  // getEdgeMeanSize :366
  for (int i = 0; i < g.n_ref; i++) {              // :367-382
    g2o::EdgesReference *e = new g2o::EdgesReference();
    e->setVertex(0, R.nodes[g.ref[i].v]);
    Eigen::Vector3d v;
    v << g.ref[i].meas[0], g.ref[i].meas[1], g.ref[i].meas[2];
    e->setMeasurement(v);
    e->setInformation(p->reg_temp * Eigen::Matrix3d::Identity() / pow(m, 2));
    e->computeError();
    R.ref.push_back(e);
    R.opt._activeEdges.push_back(e);
  }
  for (int i = 0; i < g.n_curv; i++) {  // :427-461
    const EdgeCurv &o = g.curv[i];
    g2o::EdgeMeanCurvature *e = new g2o::EdgeMeanCurvature;
    e->resize(o.nv);
    e->SetNeighbourgEdge(0);
    e->setVertex(0, R.nodes[o.v[0]]);
    std::vector<double> weights;
    for (int k = 1; k < o.nv; k++) { e->setVertex(k, R.nodes[o.v[k]]); weights.push_back(o.w[k - 1]); }
    e->setDistanceEdges(o.len);
    e->setWeights(weights);
    Eigen::Vector1D InitialMeanCurvature;
    InitialMeanCurvature << o.kappa0;
    e->setMeasurement(InitialMeanCurvature);
    e->computeError();
    e->setInformation(p->reg_lap * Eigen::Vector1D::Identity() / (double)(size_t)g.n_curv_den);
    R.curv.push_back(e);
    R.opt._activeEdges.push_back(e);
  }
  for (int i = 0; i < g.n_str; i++) {  // :482-507
    const EdgeStretch &o = g.str[i];
    g2o::EdgesStreching *e = new g2o::EdgesStreching;
    e->setVertex(0, R.nodes[o.a]);
    e->setVertex(1, R.nodes[o.b]);
    Eigen::Vector1D s;
    s << o.len0;
    e->setMeasurement(s);
    e->setInformation(p->reg_inex * Eigen::Vector1D::Identity() / (double)(size_t)g.n_str);
    e->computeError();
    R.str.push_back(e);
    R.opt._activeEdges.push_back(e);
  }
}

}  // namespace

extern "C" {

// Per-edge errors and Jacobians of the reference's edge classes at the problem's state.
// Layout = oracle_sft_residuals: rows 2*n_rep, 3*n_ref, n_curv, n_str; J dense [rows x (3n+6)] in the ABI
// variable order (nodes first, camera last); J may be NULL.  Returns the number of rows.
int ref_sft_residuals(const defslam_sft_problem *p, double *res, double *J, int max_rows) {
  Graph g;
  int rc = oracle_graph_build(&g, p);
  if (rc) { oracle_graph_free(&g); return rc; }
  RefGraph R;
  build_ref_graph(R, g, p);
  const int n = g.n_nodes, Dabi = 3 * n + 6;
  const int rows = 2 * g.n_rep + 3 * g.n_ref + g.n_curv + g.n_str;
  if (rows > max_rows) { oracle_graph_free(&g); return rows; }
  if (J) memset(J, 0, sizeof(double) * (size_t)rows * Dabi);
  g2o::JacobianWorkspace &jw = R.opt.jacobianWorkspace();
  int r0 = 0;
  for (size_t k = 0; k < R.opt._activeEdges.size(); k++) {
    g2o::OptimizableGraph::Edge *e = R.opt._activeEdges[k];
    e->computeError();
    const int d = e->dimension();
    for (int r = 0; r < d; r++) res[r0 + r] = e->errorData()[r];
    if (J) {
      e->linearizeOplus(jw);
      for (size_t vi = 0; vi < e->vertices().size(); vi++) {
        g2o::OptimizableGraph::Vertex *v = (g2o::OptimizableGraph::Vertex *)e->vertex(vi);
        const int vd = v->dimension();
        const int col0 = v->id() == 0 ? 3 * n : 3 * (v->id() - 1);
        const double *w = jw.workspaceForVertex((int)vi);  // column-major d x vd
        for (int r = 0; r < d; r++)
          for (int c = 0; c < vd; c++) J[(size_t)(r0 + r) * Dabi + col0 + c] += w[c * d + r];
      }
    }
    r0 += d;
  }
  oracle_graph_free(&g);
  return rows;
}

// H, b, chi2 built by the reference's own linearizeOplus + constructQuadraticForm + Huber, in the ABI order of
// defslam_sft_normal_equations (nodes first, camera last; rows of nodes outside OptLap are identity / zero).
int ref_sft_normal_equations(const defslam_sft_problem *p, double *H_dense, double *b, double *chi2) {
  Graph g;
  int rc = oracle_graph_build(&g, p);
  if (rc) { oracle_graph_free(&g); return rc; }
  RefGraph R;
  build_ref_graph(R, g, p);
  g2o::Solver solver(&R.opt);
  R.opt.computeActiveErrors();
  if (chi2) *chi2 = R.opt.activeRobustChi2();
  solver.buildStructure();
  solver.buildSystem();
  std::vector<double> H;
  solver.dense(H);
  const int n = g.n_nodes, Dabi = 3 * n + 6, D = (int)solver.vectorSize();
  std::vector<int> map(Dabi, -1);
  for (int v = 0; v < n; v++)
    if (g.optlap[v]) for (int c = 0; c < 3; c++) map[3 * v + c] = R.nodes[v]->colInHessian() + c;
  for (int c = 0; c < 6; c++) map[3 * n + c] = c;
  for (int i = 0; i < Dabi; i++) {
    if (b) b[i] = map[i] < 0 ? 0.0 : solver.b()[map[i]];
    if (H_dense)
      for (int j = 0; j < Dabi; j++) {
        double v = 0.0;
        if (map[i] >= 0 && map[j] >= 0) v = H[(size_t)map[i] * D + map[j]];
        else if (i == j) v = 1.0;
        H_dense[(size_t)i * Dabi + j] = v;
      }
  }
  oracle_graph_free(&g);
  return 0;
}

// The whole solve: the reference's LM driver on the reference's edges; post-processing as DefOptimizer.cc:515-577.
int ref_sft_solve(const defslam_sft_problem *p, defslam_sft_result *r) {
  Graph g;
  int rc = oracle_graph_build(&g, p);
  if (rc) { oracle_graph_free(&g); return rc; }
  RefGraph R;
  build_ref_graph(R, g, p);
  g2o::Solver solver(&R.opt);
  g2o::OptimizationAlgorithmLevenberg lm(&solver);
  lm.setOptimizer(&R.opt);
  const int maxit = p->max_iterations > 0 ? p->max_iterations : 50;
  // SparseOptimizer::optimize sparse_optimizer.cpp:403-475.  For the trace this loop evaluates the errors at the
  // top of every iteration -- at exactly the state solve() evaluates them again first thing, so nothing changes.
  int cjIterations = 0, trials = 0;
  bool ok = true;
  double chi_first = 0.0, lam_prev = 0.0;
  for (int i = 0; i < maxit && ok; i++) {
    R.opt.computeActiveErrors();
    const double chi0 = R.opt.activeRobustChi2();
    if (i == 0) {
      chi_first = chi0;
      // lambda the driver will compute (computeLambdaInit): tau * max |diag H|, evaluated here only for the trace
      solver.buildStructure();
      solver.buildSystem();
      double maxDiag = 0.0;
      for (size_t k = 0; k < R.opt._ivMap.size(); k++)
        for (int j = 0; j < R.opt._ivMap[k]->dimension(); j++) maxDiag = std::max(fabs(R.opt._ivMap[k]->hessian(j, j)), maxDiag);
      lam_prev = 1e-5 * maxDiag;
    }
    if (r->trace && i > 0 && i - 1 < r->trace_capacity) r->trace[4 * (i - 1) + 3] = chi0;
    g2o::OptimizationAlgorithm::SolverResult result = lm.solve(i, false);
    ok = (result == g2o::OptimizationAlgorithm::OK);
    trials += lm.levenbergIteration();
    if (r->trace && i < r->trace_capacity) {
      r->trace[4 * i + 0] = chi0;
      r->trace[4 * i + 1] = lam_prev;
      r->trace[4 * i + 2] = (double)lm.levenbergIteration();
      r->trace[4 * i + 3] = 0.0;
    }
    lam_prev = lm.currentLambda();
    ++cjIterations;
  }
  const float deltaMonoV = 0.5;
  // outliers, reprojection error: DefOptimizer.cc:515-559
  int nBad = 0;
  std::vector<uint8_t> outl(g.n_rep > 0 ? g.n_rep : 1, 0);
  for (int i = 0; i < g.n_rep; i++) {
    const float chi2 = R.rep[i]->chi2();
    const double *a = R.rep[i]->errorData();
    const bool out = g.variant ? (deltaMonoV < sqrt(pow(a[0], 2) + pow(a[1], 2))) : (chi2 > 5.991);  // :806-818
    if (out) { outl[i] = 1; nBad++; }
  }
  double sumError = 0.0;
  unsigned cnt = 0;
  for (int i = 0; i < g.n_rep; i++)
    if (!outl[i]) {
      R.rep[i]->computeError();
      const double *a = R.rep[i]->errorData();
      sumError += sqrt(pow(a[0], 2) + pow(a[1], 2));
      cnt++;
    }
  r->rep_error = (float)(sumError / cnt);
  // Converter::toCvMat(SE3Quat): to_homogeneous_matrix cast to float
  Eigen::Matrix<double, 4, 4> Th = R.cam->estimate().to_homogeneous_matrix();
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r->T_cw_out[i * 4 + j] = (float)Th(i, j);
  const int n = g.n_nodes;
  if (r->node_xyz_out)
    for (int v = 0; v < n; v++) for (int c = 0; c < 3; c++) r->node_xyz_out[3 * v + c] = R.nodes[v]->estimate()(c);
  if (r->outlier_out) memcpy(r->outlier_out, outl.data(), g.n_rep);
  if (r->node_role_out) for (int v = 0; v < n; v++) r->node_role_out[v] = (uint8_t)(g.viewed[v] | (g.optlap[v] << 1));
  r->n_inliers = g.n_rep - nBad;
  r->lm_iterations = cjIterations;
  r->lm_trials = trials;
  r->chi2_initial = chi_first;
  R.opt.computeActiveErrors();  // state after the last accepted step (the driver popped any rejected one)
  r->chi2_final = R.opt.activeRobustChi2();
  if (r->trace && cjIterations > 0 && cjIterations - 1 < r->trace_capacity) r->trace[4 * (cjIterations - 1) + 3] = r->chi2_final;
  r->lambda_final = lm.currentLambda();
  r->status = 0;
  oracle_graph_free(&g);
  return 0;
}

// SE3Quat::exp(update) * SE3Quat(q,t): the reference's pose update, for pinning the oracle's se3_exp / pose_oplus.
// q = (x,y,z,w).
void ref_se3_oplus(const double *q_in, const double *t_in, const double *update6, double *q_out, double *t_out) {
  Eigen::Quaterniond q(q_in[3], q_in[0], q_in[1], q_in[2]);
  Eigen::Vector3d t(t_in[0], t_in[1], t_in[2]);
  g2o::VertexSE3Expmap v;
  v.setEstimate(g2o::SE3Quat(q, t));
  v.oplus(update6);
  const g2o::SE3Quat &e = v.estimate();
  q_out[0] = e.rotation().x(); q_out[1] = e.rotation().y(); q_out[2] = e.rotation().z(); q_out[3] = e.rotation().w();
  for (int i = 0; i < 3; i++) t_out[i] = e.translation()(i);
}

// Optimizer::OptimizeHorn (Modules/Tracking/DefOptimizer.cc:840-922) on the reference's own Sim3, EdgeSim3Simple,
// VertexSim3ExpmapNoProj, numeric Jacobians (base_unary_edge.hpp:82-118), Huber kernel and Levenberg driver.
// In/out as oracle_sim3_register_batched for one problem.
int ref_sim3_optimize_horn(const defslam_sim3_problem *p, defslam_sim3_result *out) {
  g2o::SparseOptimizer opt;
  g2o::Solver solver(&opt);
  g2o::OptimizationAlgorithmLevenberg lm(&solver);
  lm.setOptimizer(&opt);
  Eigen::Quaterniond q0(p->rot[3], p->rot[0], p->rot[1], p->rot[2]);
  Eigen::Vector3d t0(p->trans[0], p->trans[1], p->trans[2]);
  g2o::Sim3 g2oS12(q0, t0, p->scale);
  g2o::VertexSim3ExpmapNoProj *vert0 = new g2o::VertexSim3ExpmapNoProj;
  vert0->setEstimate(g2oS12);
  vert0->setId(0);
  vert0->setHessianIndex(0);
  opt._ivMap.push_back(vert0);
  const int N = p->n_points;
  const float deltaHuber = sqrt(p->huber);
  std::vector<g2o::EdgeSim3Simple *> edgesSimple;
  std::vector<g2o::RobustKernelHuber *> kernels;
  for (int i = 0; i < N; i++) {
    g2o::EdgeSim3Simple *simple = new g2o::EdgeSim3Simple;
    simple->setVertex(0, vert0);
    Eigen::Matrix<double, 3, 1> v1;
    v1 << p->pts1[3 * i], p->pts1[3 * i + 1], p->pts1[3 * i + 2];
    Eigen::Matrix<double, 3, 1> v2;
    v2 << p->pts2[3 * i], p->pts2[3 * i + 1], p->pts2[3 * i + 2];
    simple->setPoints(v1, v2);
    simple->setInformation(Eigen::Matrix3d::Identity());
    g2o::RobustKernelHuber *rk = new g2o::RobustKernelHuber;
    simple->setRobustKernel(rk);
    rk->setDelta(deltaHuber);
    kernels.push_back(rk);
    edgesSimple.push_back(simple);
    opt._activeEdges.push_back(simple);
  }
  for (int run = 0; run < 2; run++) {
    int cj = 0;
    bool ok = true;
    for (int i = 0; i < p->max_iterations && ok; i++) {
      ok = lm.solve(i, false) == g2o::OptimizationAlgorithm::OK;
      ++cj;
    }
    out->iterations[run] = cj;
    if (run == 0) {
      const g2o::Sim3 &e = vert0->estimate();
      out->rot[0] = e.rotation().x(); out->rot[1] = e.rotation().y(); out->rot[2] = e.rotation().z(); out->rot[3] = e.rotation().w();
      for (int c = 0; c < 3; c++) out->trans[c] = e.translation()(c);
      out->scale = e.scale();
      int count = 0;
      for (int i = 0; i < N; i++) if (!(edgesSimple[i]->chi2() > p->chi)) count++;
      out->inliers = count;
    }
  }
  double chi2 = 0.0;  // SparseOptimizer::chi2 = activeChi2 (sparse_optimizer.cpp:93-102)
  for (int i = 0; i < N; i++) chi2 += edgesSimple[i]->chi2();
  out->chi2 = chi2;
  out->acceptable = std::isfinite(chi2) && (chi2 / out->inliers < p->chi);
  for (int i = 0; i < N; i++) { delete edgesSimple[i]; delete kernels[i]; }
  delete vert0;
  return 0;
}

// RobustKernelHuber::setDelta + robustify (the reference's lines), delta passed as the reference passes it
// (a float promoted to double)
void ref_huber(float delta, double e2, double *rho3) {
  g2o::RobustKernelHuber rk;
  rk.setDelta(delta);
  Eigen::Vector3d rho;
  rk.robustify(e2, rho);
  rho3[0] = rho[0]; rho3[1] = rho[1]; rho3[2] = rho[2];
}

}  // extern "C"
