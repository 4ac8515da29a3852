// g2o_shim.h -- class DECLARATIONS standing in for the g2o core headers that DefSLAM's SfT edge types
// (Thirdparty/g2o/g2o/types/sft_types.h) are written against.  TEST INFRASTRUCTURE ONLY (oracle/).
//
// What is the reference's own code, compiled where it lies under /root/reference (nothing copied):
//   types/sft_types.h                 verbatim  (EdgeNodesCamera, EdgeMeanCurvature, EdgesStreching, EdgesReference)
//   types/se3quat.h, se3_ops.h/.hpp   verbatim  (SE3Quat::exp / map / operator*)
//   core/base_vertex.hpp              verbatim  (ctor, clearQuadraticForm, mapHessianMemory)
//   core/base_unary_edge.hpp          verbatim  (constructQuadraticForm, numeric linearizeOplus)
//   core/base_binary_edge.hpp         verbatim  (constructQuadraticForm, mapHessianMemory)
//   core/base_multi_edge.hpp          verbatim  (constructQuadraticForm, computeQuadraticForm, mapHessianMemory)
//   core/robust_kernel_impl.cpp:65-91, core/optimization_algorithm_levenberg.cpp:43-189,
//   core/sparse_optimizer.cpp:61-74,104-120,477-491, types_six_dof_expmap.h:73-76, types_sba.h:52-56
//                                     extracted by oracle/Makefile into oracle/_ref/*.inc (git-ignored)
// What is restated here: only the class skeletons those bodies are members of (member names and types as
// in core/base_edge.h:40-104, base_vertex.h:51-112, base_unary_edge.h:42-92, base_binary_edge.h:42-112,
// base_multi_edge.h:51-105, optimizable_graph.h) -- no arithmetic apart from BaseEdge::chi2 (base_edge.h:58-61)
// and robustInformation (base_edge.h:96-102), which are one-liners inside a header that cannot be included.
#ifndef DEFSLAM_ORACLE_G2O_SHIM_H_
#define DEFSLAM_ORACLE_G2O_SHIM_H_
#include <cstring>
#include <iostream>
#include <limits>
#include <set>
#include <stack>
#include <vector>

#include "mini_eigen.h"
#include <types/se3quat.h>  // the reference's own file (-I<ref>/Thirdparty/g2o/g2o)

#define g2o_isnan(x) std::isnan(x)
#define g2o_isfinite(x) std::isfinite(x)

namespace g2o {
using namespace Eigen;

typedef Eigen::Matrix<double, 2, 1> Vector2D;
typedef Eigen::Matrix<double, 3, 1> Vector3D;

// jacobian_workspace.h:82-86 -- memory the edge Jacobians are mapped onto
class JacobianWorkspace {
 public:
  JacobianWorkspace() : _w(32, std::vector<double>(64, 0.0)) {}
  double *workspaceForVertex(int i) { return _w[i].data(); }
 private:
  std::vector<std::vector<double> > _w;
};

class RobustKernel {  // robust_kernel.h:55-78
 public:
  RobustKernel() : _delta(1.) {}
  virtual ~RobustKernel() {}
  virtual void robustify(double squaredError, Eigen::Vector3d &rho) const = 0;
  virtual void setDelta(double delta) { _delta = delta; }
  double delta() const { return _delta; }
 protected:
  double _delta;
};

class RobustKernelHuber : public RobustKernel {  // robust_kernel_impl.h:76-85 (note: dsqr is a FLOAT)
 public:
  virtual void setDelta(double delta);
  virtual void setDeltaSqr(const double &delta, const double &deltaSqr);
  virtual void robustify(double e2, Eigen::Vector3d &rho) const;
 private:
  float dsqr;
};

struct HyperGraph {
  class Vertex {
   public:
    Vertex() : _id(-1) {}
    virtual ~Vertex() {}
    int id() const { return _id; }
    void setId(int id) { _id = id; }
   protected:
    int _id;
  };
  typedef std::set<Vertex *> VertexSet;
  class Edge {
   public:
    Edge() : _id(-1) {}
    virtual ~Edge() {}
    virtual void resize(size_t size) { _vertices.resize(size, 0); }
    const std::vector<Vertex *> &vertices() const { return _vertices; }
    std::vector<Vertex *> &vertices() { return _vertices; }
    const Vertex *vertex(size_t i) const { return _vertices[i]; }
    Vertex *vertex(size_t i) { return _vertices[i]; }
    void setVertex(size_t i, Vertex *v) { _vertices[i] = v; }
    int id() const { return _id; }
   protected:
    std::vector<Vertex *> _vertices;
    int _id;
  };
};

struct OptimizableGraph {
  typedef HyperGraph::VertexSet VertexSet;
  class Vertex : public HyperGraph::Vertex {  // optimizable_graph.h
   public:
    Vertex() : _hessianIndex(-1), _fixed(false), _marginalized(false), _dimension(0), _colInHessian(-1) {}
    virtual const double &hessian(int i, int j) const = 0;
    virtual double *hessianData() = 0;
    virtual void mapHessianMemory(double *d) = 0;
    virtual int copyB(double *b_) const = 0;
    virtual double *bData() = 0;
    virtual void clearQuadraticForm() = 0;
    virtual void push() = 0;
    virtual void pop() = 0;
    virtual void discardTop() = 0;
    void oplus(const double *v) { oplusImpl(v); updateCache(); }  // optimizable_graph.h: oplus = oplusImpl + updateCache
    int hessianIndex() const { return _hessianIndex; }
    void setHessianIndex(int ti) { _hessianIndex = ti; }
    bool fixed() const { return _fixed; }
    void setFixed(bool fixed) { _fixed = fixed; }
    bool marginalized() const { return _marginalized; }
    void setMarginalized(bool m) { _marginalized = m; }
    int dimension() const { return _dimension; }
    void setColInHessian(int c) { _colInHessian = c; }
    int colInHessian() const { return _colInHessian; }
    void lockQuadraticForm() {}
    void unlockQuadraticForm() {}
   protected:
    virtual void oplusImpl(const double *v) = 0;
    void updateCache() {}
    int _hessianIndex;
    bool _fixed, _marginalized;
    int _dimension;
    int _colInHessian;
  };
  class Edge : public HyperGraph::Edge {
   public:
    Edge() : _dimension(-1), _robustKernel(0) {}
    virtual ~Edge() {}
    virtual void computeError() = 0;
    virtual double chi2() const = 0;
    virtual void constructQuadraticForm() = 0;
    virtual void mapHessianMemory(double *d, int i, int j, bool rowMajor) = 0;
    virtual void linearizeOplus(JacobianWorkspace &jacobianWorkspace) = 0;
    virtual const double *errorData() const = 0;
    RobustKernel *robustKernel() const { return _robustKernel; }
    void setRobustKernel(RobustKernel *ptr) { _robustKernel = ptr; }
    int dimension() const { return _dimension; }
   protected:
    int _dimension;
    RobustKernel *_robustKernel;
  };
};

// ---- base_vertex.h:51-112 ----
template <int D, typename T> class BaseVertex : public OptimizableGraph::Vertex {
 public:
  typedef T EstimateType;
  typedef std::stack<EstimateType, std::vector<EstimateType, Eigen::aligned_allocator<EstimateType> > > BackupStackType;
  static const int Dimension = D;
  typedef Eigen::Map<Matrix<double, D, D>, Matrix<double, D, D>::Flags & AlignedBit ? Aligned : Unaligned> HessianBlockType;

  BaseVertex();
  virtual const double &hessian(int i, int j) const { assert(i < D && j < D); return const_cast<HessianBlockType &>(_hessian)(i, j); }
  virtual double &hessian(int i, int j) { assert(i < D && j < D); return _hessian(i, j); }
  virtual double *hessianData() { return const_cast<double *>(_hessian.data()); }
  virtual void mapHessianMemory(double *d);
  virtual int copyB(double *b_) const { memcpy(b_, _b.data(), Dimension * sizeof(double)); return Dimension; }
  virtual double *bData() { return _b.data(); }
  virtual void clearQuadraticForm();
  double solveDirect(double lambda = 0);  // declared only (never instantiated: needs determinant()/llt())
  Matrix<double, D, 1> &b() { return _b; }
  const Matrix<double, D, 1> &b() const { return _b; }
  HessianBlockType &A() { return _hessian; }
  const HessianBlockType &A() const { return _hessian; }
  virtual void push() { _backup.push(_estimate); }
  virtual void pop() { assert(!_backup.empty()); _estimate = _backup.top(); _backup.pop(); updateCache(); }
  virtual void discardTop() { assert(!_backup.empty()); _backup.pop(); }
  const EstimateType &estimate() const { return _estimate; }
  void setEstimate(const EstimateType &et) { _estimate = et; updateCache(); }
 protected:
  HessianBlockType _hessian;
  Matrix<double, D, 1> _b;
  EstimateType _estimate;
  BackupStackType _backup;
};
#include <core/base_vertex.hpp>  // the reference's own file

// ---- base_edge.h:40-104 ----
template <int D, typename E> class BaseEdge : public OptimizableGraph::Edge {
 public:
  static const int Dimension = D;
  typedef E Measurement;
  typedef Matrix<double, D, 1> ErrorVector;
  typedef Matrix<double, D, D> InformationType;
  BaseEdge() : OptimizableGraph::Edge() { _dimension = D; }
  virtual ~BaseEdge() {}
  virtual double chi2() const { return _error.dot(information() * _error); }  // base_edge.h:58-61
  virtual const double *errorData() const { return _error.data(); }
  const ErrorVector &error() const { return _error; }
  ErrorVector &error() { return _error; }
  const InformationType &information() const { return _information; }
  InformationType &information() { return _information; }
  void setInformation(const InformationType &information) { _information = information; }
  const Measurement &measurement() const { return _measurement; }
  virtual void setMeasurement(const Measurement &m) { _measurement = m; }
 protected:
  Measurement _measurement;
  InformationType _information;
  ErrorVector _error;
  InformationType robustInformation(const Eigen::Vector3d &rho) {  // base_edge.h:96-102
    InformationType result = rho[1] * _information;
    return result;
  }
};

// ---- base_unary_edge.h:42-92 ----
template <int D, typename E, typename VertexXi> class BaseUnaryEdge : public BaseEdge<D, E> {
 public:
  static const int Dimension = BaseEdge<D, E>::Dimension;
  typedef typename BaseEdge<D, E>::Measurement Measurement;
  typedef VertexXi VertexXiType;
  typedef typename Matrix<double, D, VertexXiType::Dimension>::AlignedMapType JacobianXiOplusType;
  typedef typename BaseEdge<D, E>::ErrorVector ErrorVector;
  typedef typename BaseEdge<D, E>::InformationType InformationType;
  BaseUnaryEdge() : BaseEdge<D, E>(), _jacobianOplusXi(0, D, VertexXiType::Dimension) { _vertices.resize(1); }
  virtual void resize(size_t size);
  virtual bool allVerticesFixed() const;
  virtual void linearizeOplus(JacobianWorkspace &jacobianWorkspace);
  virtual void linearizeOplus();
  const JacobianXiOplusType &jacobianOplusXi() const { return _jacobianOplusXi; }
  virtual void constructQuadraticForm();
  virtual void initialEstimate(const OptimizableGraph::VertexSet &from, OptimizableGraph::Vertex *to);
  virtual void mapHessianMemory(double *, int, int, bool) { assert(0 && "BaseUnaryEdge does not map memory of the Hessian"); }
  using BaseEdge<D, E>::computeError;
 protected:
  using BaseEdge<D, E>::_measurement;
  using BaseEdge<D, E>::_information;
  using BaseEdge<D, E>::_error;
  using BaseEdge<D, E>::_vertices;
  using BaseEdge<D, E>::_dimension;
  JacobianXiOplusType _jacobianOplusXi;
};
#include <core/base_unary_edge.hpp>  // the reference's own file

// ---- base_binary_edge.h:42-112 ----
template <int D, typename E, typename VertexXi, typename VertexXj> class BaseBinaryEdge : public BaseEdge<D, E> {
 public:
  typedef VertexXi VertexXiType;
  typedef VertexXj VertexXjType;
  static const int Di = VertexXiType::Dimension;
  static const int Dj = VertexXjType::Dimension;
  static const int Dimension = BaseEdge<D, E>::Dimension;
  typedef typename BaseEdge<D, E>::Measurement Measurement;
  typedef typename Matrix<double, D, Di>::AlignedMapType JacobianXiOplusType;
  typedef typename Matrix<double, D, Dj>::AlignedMapType JacobianXjOplusType;
  typedef typename BaseEdge<D, E>::ErrorVector ErrorVector;
  typedef typename BaseEdge<D, E>::InformationType InformationType;
  typedef Eigen::Map<Matrix<double, Di, Dj>, Matrix<double, Di, Dj>::Flags & AlignedBit ? Aligned : Unaligned> HessianBlockType;
  typedef Eigen::Map<Matrix<double, Dj, Di>, Matrix<double, Dj, Di>::Flags & AlignedBit ? Aligned : Unaligned> HessianBlockTransposedType;
  BaseBinaryEdge()
      : BaseEdge<D, E>(), _hessianRowMajor(false), _hessian(0, VertexXiType::Dimension, VertexXjType::Dimension),
        _hessianTransposed(0, VertexXjType::Dimension, VertexXiType::Dimension), _jacobianOplusXi(0, D, Di),
        _jacobianOplusXj(0, D, Dj) {
    _vertices.resize(2);
  }
  virtual OptimizableGraph::Vertex *createFrom();
  virtual OptimizableGraph::Vertex *createTo();
  virtual void resize(size_t size);
  virtual bool allVerticesFixed() const;
  virtual void linearizeOplus(JacobianWorkspace &jacobianWorkspace);
  virtual void linearizeOplus();
  const JacobianXiOplusType &jacobianOplusXi() const { return _jacobianOplusXi; }
  const JacobianXjOplusType &jacobianOplusXj() const { return _jacobianOplusXj; }
  virtual void constructQuadraticForm();
  virtual void mapHessianMemory(double *d, int i, int j, bool rowMajor);
  using BaseEdge<D, E>::resize;
  using BaseEdge<D, E>::computeError;
 protected:
  using BaseEdge<D, E>::_measurement;
  using BaseEdge<D, E>::_information;
  using BaseEdge<D, E>::_error;
  using BaseEdge<D, E>::_vertices;
  using BaseEdge<D, E>::_dimension;
  bool _hessianRowMajor;
  HessianBlockType _hessian;
  HessianBlockTransposedType _hessianTransposed;
  JacobianXiOplusType _jacobianOplusXi;
  JacobianXjOplusType _jacobianOplusXj;
};
#include <core/base_binary_edge.hpp>  // the reference's own file

// ---- base_multi_edge.h:51-105 ----
template <int D, typename E> class BaseMultiEdge : public BaseEdge<D, E> {
 public:
  struct HessianHelper {
    Eigen::Map<MatrixXd> matrix;
    bool transposed;
    HessianHelper() : matrix(0, 0, 0), transposed(false) {}
  };
  static const int Dimension = BaseEdge<D, E>::Dimension;
  typedef typename BaseEdge<D, E>::Measurement Measurement;
  typedef MatrixXd::MapType JacobianType;
  typedef typename BaseEdge<D, E>::ErrorVector ErrorVector;
  typedef typename BaseEdge<D, E>::InformationType InformationType;
  typedef Eigen::Map<MatrixXd, MatrixXd::Flags & AlignedBit ? Aligned : Unaligned> HessianBlockType;
  BaseMultiEdge() : BaseEdge<D, E>() {}
  virtual void linearizeOplus(JacobianWorkspace &jacobianWorkspace);
  virtual void linearizeOplus();
  virtual void resize(size_t size);
  virtual bool allVerticesFixed() const;
  virtual void constructQuadraticForm();
  virtual void mapHessianMemory(double *d, int i, int j, bool rowMajor);
  using BaseEdge<D, E>::computeError;
 protected:
  using BaseEdge<D, E>::_measurement;
  using BaseEdge<D, E>::_information;
  using BaseEdge<D, E>::_error;
  using BaseEdge<D, E>::_vertices;
  using BaseEdge<D, E>::_dimension;
  std::vector<HessianHelper> _hessian;
  std::vector<JacobianType, aligned_allocator<JacobianType> > _jacobianOplus;
  void computeQuadraticForm(const InformationType &omega, const ErrorVector &weightedError);
};
#include <core/base_multi_edge.hpp>  // the reference's own file

// ---- vertices of the SfT graph: skeletons of types_six_dof_expmap.h:60-77 and types_sba.h:40-57;
//      the oplusImpl bodies are the reference's lines ----
class VertexSE3Expmap : public BaseVertex<6, SE3Quat> {
 public:
  VertexSE3Expmap() {}
#include "../../_ref/vertex_se3_oplus.inc"  // types_six_dof_expmap.h:73-76
};
class VertexSBAPointXYZ : public BaseVertex<3, Vector3d> {
 public:
  VertexSBAPointXYZ() {}
#include "../../_ref/vertex_xyz_oplus.inc"  // types_sba.h:52-56
};

}  // namespace g2o
#endif
