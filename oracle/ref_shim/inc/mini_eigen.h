// mini_eigen.h -- the small subset of the Eigen API that DefSLAM's own g2o sources use on the SfT path,
// so that those sources (Thirdparty/g2o/g2o/types/sft_types.h, se3quat.h, se3_ops.h verbatim; the bodies
// of optimization_algorithm_levenberg.cpp, base_*_edge.hpp, robust_kernel_impl.cpp extracted at build
// time) compile in an image that has no Eigen.  TEST INFRASTRUCTURE ONLY (oracle/): it exists to pin the
// C oracle against the reference's own code.  Not an Eigen copy: value semantics, no expression
// templates, no vectorisation -- every operator evaluates straight away in the order the reference's
// expression is written (left to right, coefficient sums k = 0,1,2,... like Eigen's lazy small products).
// The pieces of arithmetic that live INSIDE Eigen and that the reference relies on are restated from
// Eigen's documented algorithms: Quaternion(Matrix3), Quaternion::toRotationMatrix, quaternion * vector,
// quaternion product, normalize (division by the norm).
#ifndef DEFSLAM_ORACLE_MINI_EIGEN_H_
#define DEFSLAM_ORACLE_MINI_EIGEN_H_
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <memory>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3

namespace Eigen {

enum { Dynamic = -1 };
enum { ColMajor = 0, RowMajor = 1 };
enum { Unaligned = 0, Aligned = 16 };
enum { AlignedBit = 0x80 };
enum { Upper = 1, Lower = 2 };

template <typename T> using aligned_allocator = std::allocator<T>;

template <typename Scalar, int R, int C, int Options = 0> class Matrix;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;

template <typename Derived> class MatrixBase;

// filled row by row, whatever the storage order (Eigen's CommaInitializer)
template <typename Derived> class CommaInitializer {
 public:
  CommaInitializer(Derived &m) : m_(m), k_(0) {}
  CommaInitializer &operator,(double v) { put(v); return *this; }
  template <typename O> CommaInitializer &operator,(const MatrixBase<O> &o) { putm(o); return *this; }
  void put(double v) {
    const int c = m_.cols();
    assert(k_ < m_.rows() * c);
    m_(k_ / c, k_ % c) = v;
    k_++;
  }
  template <typename O> void putm(const MatrixBase<O> &o) {
    // only whole-matrix / vector stacking is used by the reference
    if (o.rows() == m_.rows() && o.cols() == m_.cols()) {
      for (int i = 0; i < o.rows(); i++)
        for (int j = 0; j < o.cols(); j++) m_(i, j) = o(i, j);
      k_ = m_.rows() * m_.cols();
    } else {
      for (int i = 0; i < o.rows(); i++)
        for (int j = 0; j < o.cols(); j++) put(o(i, j));
    }
  }
 private:
  Derived &m_;
  int k_;
};

// a writable rectangular view (block(), col(), head())
template <typename Derived> class BlockRef : public MatrixBase<BlockRef<Derived> > {
 public:
  BlockRef(Derived &m, int r0, int c0, int nr, int nc) : m_(m), r0_(r0), c0_(c0), nr_(nr), nc_(nc) {}
  int rows() const { return nr_; }
  int cols() const { return nc_; }
  double &coeffRef(int i, int j) { return m_.coeffRef(r0_ + i, c0_ + j); }
  double coeff(int i, int j) const { return const_cast<Derived &>(m_).coeffRef(r0_ + i, c0_ + j); }
  template <typename O> BlockRef &operator=(const MatrixBase<O> &o) {
    assert(o.rows() == nr_ && o.cols() == nc_);
    for (int i = 0; i < nr_; i++)
      for (int j = 0; j < nc_; j++) coeffRef(i, j) = o(i, j);
    return *this;
  }
  BlockRef &operator=(const BlockRef &o) { return operator=<BlockRef>(o); }
  BlockRef head(int n) { return nc_ == 1 ? BlockRef(m_, r0_, c0_, n, 1) : BlockRef(m_, r0_, c0_, 1, n); }
 private:
  Derived &m_;
  int r0_, c0_, nr_, nc_;
};

template <typename Derived> class MatrixBase {
 public:
  Derived &derived() { return *static_cast<Derived *>(this); }
  const Derived &derived() const { return *static_cast<const Derived *>(this); }
  int rows() const { return derived().rows(); }
  int cols() const { return derived().cols(); }
  int size() const { return rows() * cols(); }
  double operator()(int i, int j) const { return derived().coeff(i, j); }
  double &operator()(int i, int j) { return derived().coeffRef(i, j); }
  // single index: vectors (and 1-row / 1-column dynamic matrices)
  double operator()(int i) const { return cols() == 1 ? derived().coeff(i, 0) : derived().coeff(0, i); }
  double &operator()(int i) { return cols() == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
  double operator[](int i) const { return (*this)(i); }
  double &operator[](int i) { return (*this)(i); }

  Derived &noalias() { return derived(); }
  MatrixXd transpose() const;
  MatrixXd inverse() const;  // 1x1 only (all the reference inverts on this path)
  // PartialPivLU stand-in: only Sim3::log() (sim3.h, not on the tested path) needs it to compile
  struct LuSolver {
    std::vector<double> a; int n;
    template <typename O> Matrix<double, Dynamic, Dynamic> solve(const MatrixBase<O> &b) const;
  };
  LuSolver lu() const {
    LuSolver l; l.n = rows(); l.a.resize((size_t)l.n * l.n);
    for (int i = 0; i < l.n; i++) for (int j = 0; j < l.n; j++) l.a[(size_t)i * l.n + j] = (*this)(i, j);
    return l;
  }
  double squaredNorm() const {
    double s = 0;
    bool first = true;
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) {
        const double v = (*this)(i, j);
        if (first) { s = v * v; first = false; } else s += v * v;
      }
    return s;
  }
  double norm() const { return std::sqrt(squaredNorm()); }
  template <typename O> double dot(const MatrixBase<O> &o) const {
    assert(size() == o.size());
    double s = 0;
    for (int i = 0; i < size(); i++) { const double p = (*this)(i) * o(i); s = i ? s + p : p; }
    return s;
  }
  template <typename O> Vector3d cross(const MatrixBase<O> &o) const;
  double trace() const { double s = 0; for (int i = 0; i < rows(); i++) s = i ? s + (*this)(i, i) : (*this)(i, i); return s; }

  void fill(double v) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) (*this)(i, j) = v; }
  Derived &setZero() { fill(0.0); return derived(); }
  Derived &setIdentity() { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) (*this)(i, j) = i == j ? 1.0 : 0.0; return derived(); }

  CommaInitializer<Derived> operator<<(double v) { CommaInitializer<Derived> c(derived()); c.put(v); return c; }
  template <typename O> CommaInitializer<Derived> operator<<(const MatrixBase<O> &o) {
    CommaInitializer<Derived> c(derived()); c.putm(o); return c;
  }

  template <typename O> Derived &operator+=(const MatrixBase<O> &o) {
    assert(rows() == o.rows() && cols() == o.cols());
    for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) (*this)(i, j) += o(i, j);
    return derived();
  }
  template <typename O> Derived &operator-=(const MatrixBase<O> &o) {
    assert(rows() == o.rows() && cols() == o.cols());
    for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) (*this)(i, j) -= o(i, j);
    return derived();
  }
  Derived &operator*=(double s) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) (*this)(i, j) *= s; return derived(); }
  Derived &operator/=(double s) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) (*this)(i, j) /= s; return derived(); }

  BlockRef<Derived> block(int r0, int c0, int nr, int nc) { return BlockRef<Derived>(derived(), r0, c0, nr, nc); }
  BlockRef<Derived> col(int j) { return BlockRef<Derived>(derived(), 0, j, rows(), 1); }
  BlockRef<Derived> row(int i) { return BlockRef<Derived>(derived(), i, 0, 1, cols()); }
};

template <int R, int C> struct mini_storage {
  double v[R * C];
  mini_storage() { for (int i = 0; i < R * C; i++) v[i] = 0.0; }
  void resize(int r, int c) { assert(r == R && c == C); (void)r; (void)c; }
  int rows() const { return R; }
  int cols() const { return C; }
  double *data() { return v; }
  const double *data() const { return v; }
};
template <int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)> struct mini_storage_sel { typedef mini_storage<R, C> type; };
struct mini_dyn_storage {
  std::vector<double> v;
  int r, c;
  mini_dyn_storage() : r(0), c(0) {}
  void resize(int rr, int cc) { if (rr != r || cc != c) { r = rr; c = cc; v.assign((size_t)rr * cc, 0.0); } }
  int rows() const { return r; }
  int cols() const { return c; }
  double *data() { return v.data(); }
  const double *data() const { return v.data(); }
};
template <int R, int C> struct mini_storage_sel<R, C, true> { typedef mini_dyn_storage type; };

template <typename PlainType, int MapOptions = 0> class Map;

template <typename Scalar, int R, int C, int Options> class Matrix : public MatrixBase<Matrix<Scalar, R, C, Options> > {
  typedef MatrixBase<Matrix<Scalar, R, C, Options> > Base;
 public:
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, Flags = 0 };
  typedef Map<Matrix> MapType;
  typedef Map<Matrix> AlignedMapType;
  typedef Map<Matrix> ConstMapType;

  Matrix() { if (R != Dynamic && C != Dynamic) s_.resize(R, C); else if (C == 1) s_.resize(0, 1); }
  Matrix(const Matrix &o) : Base(), s_(o.s_) {}
  template <typename O> Matrix(const MatrixBase<O> &o) { assign(o); }
  // (rows, cols) for dynamic matrices, (x, y) for 2-vectors -- as in Eigen
  Matrix(double a, double b) {
    if (R == 2 && C == 1) { s_.resize(2, 1); s_.data()[0] = a; s_.data()[1] = b; }
    else s_.resize((int)a, (int)b);
  }
  explicit Matrix(int n) { s_.resize(C == 1 ? n : 1, C == 1 ? 1 : n); }
  Matrix(double x, double y, double z) { s_.resize(3, 1); s_.data()[0] = x; s_.data()[1] = y; s_.data()[2] = z; }
  Matrix(double x, double y, double z, double w) { s_.resize(4, 1); double *d = s_.data(); d[0] = x; d[1] = y; d[2] = z; d[3] = w; }

  Matrix &operator=(const Matrix &o) { s_ = o.s_; return *this; }
  template <typename O> Matrix &operator=(const MatrixBase<O> &o) { assign(o); return *this; }

  int rows() const { return s_.rows(); }
  int cols() const { return s_.cols(); }
  void resize(int r, int c) { s_.resize(r, c); }
  double *data() { return s_.data(); }
  const double *data() const { return s_.data(); }
  // column-major unless RowMajor is asked for
  double &coeffRef(int i, int j) {
    assert(i >= 0 && i < rows() && j >= 0 && j < cols());
    return (Options & RowMajor) ? s_.data()[(size_t)i * cols() + j] : s_.data()[(size_t)j * rows() + i];
  }
  double coeff(int i, int j) const { return const_cast<Matrix *>(this)->coeffRef(i, j); }

  static Matrix Zero() { Matrix m; m.fill(0.0); return m; }
  static Matrix Zero(int r, int c) { Matrix m; m.resize(r, c); m.fill(0.0); return m; }
  static Matrix Ones(int r, int c) { Matrix m; m.resize(r, c); m.fill(1.0); return m; }
  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
  static Matrix Identity(int r, int c) { Matrix m; m.resize(r, c); m.setIdentity(); return m; }

 private:
  template <typename O> void assign(const MatrixBase<O> &o) {
    // a 1x1 result may be assigned to any 1x1, a vector to a vector of the same length
    if (R != Dynamic && C != Dynamic) {
      assert((o.rows() == R && o.cols() == C) || (o.size() == R * C && (R == 1 || C == 1)));
      s_.resize(R, C);
      if (o.rows() == R) { for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) coeffRef(i, j) = o(i, j); }
      else { for (int i = 0; i < R * C; i++) s_.data()[i] = o(i); }
    } else {
      // evaluate into a temporary first: o may alias *this
      const int r = o.rows(), c = o.cols();
      std::vector<double> t((size_t)r * c);
      for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) t[(size_t)j * r + i] = o(i, j);
      s_.resize(r, c);
      for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) coeffRef(i, j) = t[(size_t)j * r + i];
    }
  }
  typename mini_storage_sel<R, C>::type s_;
};

// view of caller-owned column-major memory
template <typename PlainType, int MapOptions> class Map : public MatrixBase<Map<PlainType, MapOptions> > {
 public:
  Map(const double *d, int r, int c) : d_(const_cast<double *>(d)), r_(r), c_(c) {}
  Map(const double *d, int n) : d_(const_cast<double *>(d)), r_(PlainType::ColsAtCompileTime == 1 ? n : 1), c_(PlainType::ColsAtCompileTime == 1 ? 1 : n) {}
  explicit Map(const double *d) : d_(const_cast<double *>(d)), r_(PlainType::RowsAtCompileTime), c_(PlainType::ColsAtCompileTime) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  double *data() { return d_; }
  const double *data() const { return d_; }
  double &coeffRef(int i, int j) { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return d_[(size_t)j * r_ + i]; }
  double coeff(int i, int j) const { return d_[(size_t)j * r_ + i]; }
  void resize(int r, int c) { assert(r == r_ && c == c_); (void)r; (void)c; }
  template <typename O> Map &operator=(const MatrixBase<O> &o) {
    assert(o.rows() == r_ && o.cols() == c_);
    for (int i = 0; i < r_; i++) for (int j = 0; j < c_; j++) coeffRef(i, j) = o(i, j);
    return *this;
  }
  Map &operator=(const Map &o) { return operator=<Map>(o); }
 private:
  double *d_;
  int r_, c_;
};
// Map<const T>: read-only use in the reference; same view
template <typename PlainType, int MapOptions> class Map<const PlainType, MapOptions> : public Map<PlainType, MapOptions> {
 public:
  Map(const double *d, int r, int c) : Map<PlainType, MapOptions>(d, r, c) {}
  Map(const double *d, int n) : Map<PlainType, MapOptions>(d, n) {}
  explicit Map(const double *d) : Map<PlainType, MapOptions>(d) {}
};

// ---- operators: evaluated immediately, results are dynamic matrices ----
template <typename A, typename B> MatrixXd operator+(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  MatrixXd r(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r(i, j) = a(i, j) + b(i, j);
  return r;
}
template <typename A, typename B> MatrixXd operator-(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  MatrixXd r(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r(i, j) = a(i, j) - b(i, j);
  return r;
}
template <typename A> MatrixXd operator-(const MatrixBase<A> &a) {
  MatrixXd r(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r(i, j) = -a(i, j);
  return r;
}
template <typename A> MatrixXd operator*(double s, const MatrixBase<A> &a) {
  MatrixXd r(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r(i, j) = s * a(i, j);
  return r;
}
template <typename A> MatrixXd operator*(const MatrixBase<A> &a, double s) {
  MatrixXd r(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r(i, j) = a(i, j) * s;
  return r;
}
template <typename A> MatrixXd operator/(const MatrixBase<A> &a, double s) {
  MatrixXd r(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r(i, j) = a(i, j) / s;
  return r;
}
template <typename A, typename B> MatrixXd operator*(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  assert(a.cols() == b.rows());
  MatrixXd r(a.rows(), b.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < b.cols(); j++) {
      double s = 0;
      for (int k = 0; k < a.cols(); k++) { const double p = a(i, k) * b(k, j); s = k ? s + p : p; }
      r(i, j) = s;
    }
  return r;
}
template <typename D> std::ostream &operator<<(std::ostream &os, const MatrixBase<D> &m) {
  for (int i = 0; i < m.rows(); i++) { for (int j = 0; j < m.cols(); j++) os << (j ? " " : "") << m(i, j); os << "\n"; }
  return os;
}

template <typename Derived> MatrixXd MatrixBase<Derived>::transpose() const {
  MatrixXd r(cols(), rows());
  for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) r(j, i) = (*this)(i, j);
  return r;
}
template <typename Derived> MatrixXd MatrixBase<Derived>::inverse() const {
  assert(rows() == cols() && rows() == 1 && "mini_eigen: only the 1x1 inverse the reference uses");
  MatrixXd r(1, 1);
  r(0, 0) = 1.0 / (*this)(0, 0);
  return r;
}
template <typename Derived> template <typename O>
MatrixXd MatrixBase<Derived>::LuSolver::solve(const MatrixBase<O> &b) const {
  std::vector<double> A(a), x(n);
  for (int i = 0; i < n; i++) x[i] = b(i);
  for (int k = 0; k < n; k++) {  // Gaussian elimination with partial pivoting
    int p = k;
    for (int i = k + 1; i < n; i++) if (std::fabs(A[(size_t)i * n + k]) > std::fabs(A[(size_t)p * n + k])) p = i;
    if (p != k) { for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]); std::swap(x[k], x[p]); }
    for (int i = k + 1; i < n; i++) {
      const double f = A[(size_t)i * n + k] / A[(size_t)k * n + k];
      for (int j = k; j < n; j++) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
      x[i] -= f * x[k];
    }
  }
  MatrixXd r(n, 1);
  for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int j = i + 1; j < n; j++) s -= A[(size_t)i * n + j] * r(j, 0); r(i, 0) = s / A[(size_t)i * n + i]; }
  return r;
}
template <typename Derived> template <typename O> Vector3d MatrixBase<Derived>::cross(const MatrixBase<O> &o) const {
  const MatrixBase &a = *this;
  Vector3d r;
  r(0) = a(1) * o(2) - a(2) * o(1);
  r(1) = a(2) * o(0) - a(0) * o(2);
  r(2) = a(0) * o(1) - a(1) * o(0);
  return r;
}

// ---- Quaternion (coeffs stored x,y,z,w like Eigen) ----
class Quaterniond {
 public:
  Quaterniond() {}
  Quaterniond(double w, double x, double y, double z) { c_(0) = x; c_(1) = y; c_(2) = z; c_(3) = w; }
  // Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other,3,3> (Ken Shoemake's method)
  template <typename O> explicit Quaterniond(const MatrixBase<O> &mat) {
    assert(mat.rows() == 3 && mat.cols() == 3);
    double t = mat.trace();
    if (t > 0.0) {
      t = std::sqrt(t + 1.0);
      w() = 0.5 * t;
      t = 0.5 / t;
      x() = (mat(2, 1) - mat(1, 2)) * t;
      y() = (mat(0, 2) - mat(2, 0)) * t;
      z() = (mat(1, 0) - mat(0, 1)) * t;
    } else {
      int i = 0;
      if (mat(1, 1) > mat(0, 0)) i = 1;
      if (mat(2, 2) > mat(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + 1.0);
      c_(i) = 0.5 * t;
      t = 0.5 / t;
      w() = (mat(k, j) - mat(j, k)) * t;
      c_(j) = (mat(j, i) + mat(i, j)) * t;
      c_(k) = (mat(k, i) + mat(i, k)) * t;
    }
  }
  double &x() { return c_(0); }
  double &y() { return c_(1); }
  double &z() { return c_(2); }
  double &w() { return c_(3); }
  double x() const { return c_(0); }
  double y() const { return c_(1); }
  double z() const { return c_(2); }
  double w() const { return c_(3); }
  Vector4d &coeffs() { return c_; }
  const Vector4d &coeffs() const { return c_; }
  void setIdentity() { c_(0) = c_(1) = c_(2) = 0.0; c_(3) = 1.0; }
  double squaredNorm() const { return c_.squaredNorm(); }
  double norm() const { return c_.norm(); }
  void normalize() { c_ /= norm(); }
  Quaterniond conjugate() const { return Quaterniond(w(), -x(), -y(), -z()); }
  Quaterniond operator*(const Quaterniond &b) const {
    const Quaterniond &a = *this;
    return Quaterniond(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                       a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                       a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                       a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaterniond &operator*=(const Quaterniond &b) { *this = *this * b; return *this; }
  // QuaternionBase::_transformVector: v + w*uv + vec x uv with uv = 2 vec x v
  template <typename O> Vector3d operator*(const MatrixBase<O> &v) const {
    Vector3d q; q(0) = x(); q(1) = y(); q(2) = z();
    Vector3d uv = q.cross(v);
    uv += uv;
    Vector3d r = q.cross(uv);
    Vector3d out;
    for (int i = 0; i < 3; i++) out(i) = v(i) + w() * uv(i) + r(i);
    return out;
  }
  // QuaternionBase::toRotationMatrix
  Matrix3d toRotationMatrix() const {
    Matrix3d res;
    const double tx = 2.0 * x(), ty = 2.0 * y(), tz = 2.0 * z();
    const double twx = tx * w(), twy = ty * w(), twz = tz * w();
    const double txx = tx * x(), txy = ty * x(), txz = tz * x();
    const double tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res(0, 0) = 1.0 - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
    res(1, 0) = txy + twz; res(1, 1) = 1.0 - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = 1.0 - (txx + tyy);
    return res;
  }
 private:
  Vector4d c_;
};

class Isometry3d {
 public:
  Isometry3d() {}
  explicit Isometry3d(const Quaterniond &q) : R_(q.toRotationMatrix()) {}
  Vector3d &translation() { return t_; }
  const Matrix3d &linear() const { return R_; }
 private:
  Matrix3d R_;
  Vector3d t_;
};

}  // namespace Eigen
#endif
