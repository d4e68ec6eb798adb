// mini_eigen.hpp -- STAND-IN for the slice of Eigen3 that graph_core's planner sources use.
//
// TEST INFRASTRUCTURE ONLY (oracle/): this image has no Eigen, so the reference's own
// translation units are compiled against this header to obtain oracle/_ref, the executable form of
// the reference that pins oracle/gc_oracle.cpp.  Nothing under graph_core_b200/ includes it.
//
// Dense, dynamically sized, column-major, double only, eager evaluation (no expression templates).
// Arithmetic conventions: every elementwise operation is one IEEE fp64 operation; reductions
// (norm, dot, matrix products) accumulate sequentially in index order.  Real Eigen vectorises its
// reductions (packet-wise partial sums), so a result can differ from real Eigen in the last ulp; the
// control flow, tie-breaking and visiting order of the reference -- what oracle/_ref is for -- do
// not depend on that (DESIGN.md "oracle pinning").
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <initializer_list>
#include <ostream>
#include <vector>

#include "shim_random.hpp"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 0
#define MINI_EIGEN_STANDIN 1

namespace Eigen
{
typedef std::ptrdiff_t Index;
const int Dynamic = -1;

class MatrixXd;
class VectorXd;

// writable view of one column of a MatrixXd
class ColXpr
{
public:
  ColXpr(MatrixXd& m, Index c) : m_(m), c_(c) {}
  inline Index size() const;
  inline Index rows() const { return size(); }
  inline double& operator()(Index i);
  inline double operator()(Index i) const;
  inline ColXpr& operator=(const MatrixXd& v);
  inline ColXpr& operator=(const ColXpr& v);
  inline ColXpr& operator-=(const MatrixXd& v);
  inline ColXpr& operator/=(double s);
  inline double norm() const;
  inline double dot(const MatrixXd& v) const;
  inline double dot(const ColXpr& v) const;
  inline operator MatrixXd() const;

private:
  MatrixXd& m_;
  Index c_;
};

class MatrixXd
{
public:
  enum { ColsAtCompileTime = Dynamic, RowsAtCompileTime = Dynamic };
  MatrixXd() : r_(0), c_(0) {}
  MatrixXd(Index r, Index c) : r_(r), c_(c), v_((size_t)(r * c), 0.0) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  void resize(Index r, Index c)
  {
    r_ = r;
    c_ = c;
    v_.resize((size_t)(r * c));
  }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(Index i, Index j) { return v_[(size_t)(j * r_ + i)]; }
  const double& operator()(Index i, Index j) const { return v_[(size_t)(j * r_ + i)]; }
  double& operator()(Index i) { return v_[(size_t)i]; }
  const double& operator()(Index i) const { return v_[(size_t)i]; }
  double& operator[](Index i) { return v_[(size_t)i]; }
  const double& operator[](Index i) const { return v_[(size_t)i]; }

  ColXpr col(Index c) { return ColXpr(*this, c); }
  MatrixXd col(Index c) const
  {
    MatrixXd o(r_, 1);
    for (Index i = 0; i < r_; i++)
      o(i) = (*this)(i, c);
    return o;
  }

  // uniform in [-1, 1] like Eigen's default scalar random for double.  Real Eigen draws from std::rand(); the
  // stand-in draws from a private generator (shim_random.hpp) so that a seeded run is reproducible whatever other
  // threads of the process (CUDA driver, BLAS pools ...) do with the C library's global stream
  static double randomScalar() { return 2.0 * (double)(gc_shim_rng::next() >> 11) * (1.0 / 9007199254740992.0) - 1.0; }
  static MatrixXd Random(Index r, Index c)
  {
    MatrixXd o(r, c);
    for (auto& x : o.v_)
      x = randomScalar();
    return o;
  }
  MatrixXd& setRandom()
  {
    for (auto& x : v_)
      x = randomScalar();
    return *this;
  }
  MatrixXd& setRandom(Index r, Index c)
  {
    resize(r, c);
    return setRandom();
  }
  MatrixXd& setConstant(double s)
  {
    for (auto& x : v_)
      x = s;
    return *this;
  }
  MatrixXd& setZero() { return setConstant(0.0); }
  MatrixXd& setOnes() { return setConstant(1.0); }
  MatrixXd& setOnes(Index r, Index c)
  {
    resize(r, c);
    return setOnes();
  }
  MatrixXd& setIdentity()
  {
    setZero();
    for (Index i = 0; i < (r_ < c_ ? r_ : c_); i++)
      (*this)(i, i) = 1.0;
    return *this;
  }

  double squaredNorm() const
  {
    double s = 0.0;
    for (double x : v_)
      s += x * x;
    return s;
  }
  double norm() const { return std::sqrt(squaredNorm()); }
  void normalize() { *this /= norm(); }
  MatrixXd normalized() const
  {
    MatrixXd o(*this);
    o /= norm();
    return o;
  }
  double dot(const MatrixXd& o) const
  {
    assert(o.size() == size());
    double s = 0.0;
    for (size_t i = 0; i < v_.size(); i++)
      s += v_[i] * o.v_[i];
    return s;
  }
  double dot(const ColXpr& o) const
  {
    double s = 0.0;
    for (Index i = 0; i < size(); i++)
      s += v_[(size_t)i] * o(i);
    return s;
  }
  double maxCoeff() const
  {
    double m = v_[0];
    for (double x : v_)
      if (x > m)
        m = x;
    return m;
  }
  double minCoeff() const
  {
    double m = v_[0];
    for (double x : v_)
      if (x < m)
        m = x;
    return m;
  }
  double sum() const
  {
    double s = 0.0;
    for (double x : v_)
      s += x;
    return s;
  }
  MatrixXd cwiseAbs() const
  {
    MatrixXd o(*this);
    for (auto& x : o.v_)
      x = std::fabs(x);
    return o;
  }
  MatrixXd cwiseInverse() const
  {
    MatrixXd o(*this);
    for (auto& x : o.v_)
      x = 1.0 / x;
    return o;
  }
  MatrixXd cwiseProduct(const MatrixXd& b) const
  {
    assert(b.size() == size());
    MatrixXd o(*this);
    for (size_t i = 0; i < v_.size(); i++)
      o.v_[i] = v_[i] * b.v_[i];
    return o;
  }
  MatrixXd transpose() const
  {
    MatrixXd o(c_, r_);
    for (Index i = 0; i < r_; i++)
      for (Index j = 0; j < c_; j++)
        o(j, i) = (*this)(i, j);
    return o;
  }
  MatrixXd asDiagonal() const
  {
    MatrixXd o(size(), size());
    for (Index i = 0; i < size(); i++)
      o(i, i) = v_[(size_t)i];
    return o;
  }

  MatrixXd& operator+=(const MatrixXd& b)
  {
    assert(b.size() == size());
    for (size_t i = 0; i < v_.size(); i++)
      v_[i] += b.v_[i];
    return *this;
  }
  MatrixXd& operator-=(const MatrixXd& b)
  {
    assert(b.size() == size());
    for (size_t i = 0; i < v_.size(); i++)
      v_[i] -= b.v_[i];
    return *this;
  }
  MatrixXd& operator*=(double s)
  {
    for (auto& x : v_)
      x *= s;
    return *this;
  }
  MatrixXd& operator/=(double s)
  {
    for (auto& x : v_)
      x /= s;
    return *this;
  }
  bool operator==(const MatrixXd& b) const { return r_ == b.r_ && c_ == b.c_ && v_ == b.v_; }
  bool operator!=(const MatrixXd& b) const { return !(*this == b); }

protected:
  Index r_, c_;
  std::vector<double> v_;
};

class VectorXd : public MatrixXd
{
public:
  enum { ColsAtCompileTime = 1 };
  VectorXd() : MatrixXd() {}
  VectorXd(Index n) : MatrixXd(n, 1) {}
  VectorXd(int n) : MatrixXd(n, 1) {}
  VectorXd(unsigned int n) : MatrixXd((Index)n, 1) {}
  VectorXd(size_t n) : MatrixXd((Index)n, 1) {}
  VectorXd(const MatrixXd& m) : MatrixXd(m) { assert(m.cols() == 1 || m.size() == 0); }
  VectorXd(const ColXpr& c) : MatrixXd((MatrixXd)c) {}
  using MatrixXd::resize;
  void resize(Index n) { MatrixXd::resize(n, 1); }
  using MatrixXd::setConstant;
  using MatrixXd::setOnes;
  using MatrixXd::setRandom;
  VectorXd& setConstant(Index n, double s)
  {
    resize(n);
    MatrixXd::setConstant(s);
    return *this;
  }
};

template <typename S, int R, int C>
struct MatrixSel
{
  typedef MatrixXd type;
};
template <typename S, int R, int C>
using Matrix = typename MatrixSel<S, R, C>::type;

inline MatrixXd operator+(const MatrixXd& a, const MatrixXd& b)
{
  MatrixXd o(a);
  o += b;
  return o;
}
inline MatrixXd operator-(const MatrixXd& a, const MatrixXd& b)
{
  MatrixXd o(a);
  o -= b;
  return o;
}
inline MatrixXd operator-(const MatrixXd& a)
{
  MatrixXd o(a);
  for (Index i = 0; i < o.size(); i++)
    o(i) = -o(i);
  return o;
}
inline MatrixXd operator*(const MatrixXd& a, double s)
{
  MatrixXd o(a);
  o *= s;
  return o;
}
inline MatrixXd operator*(double s, const MatrixXd& a)
{
  MatrixXd o(a);
  for (Index i = 0; i < o.size(); i++)
    o(i) = s * o(i);
  return o;
}
inline MatrixXd operator/(const MatrixXd& a, double s)
{
  MatrixXd o(a);
  o /= s;
  return o;
}
inline MatrixXd operator*(const MatrixXd& a, const MatrixXd& b)
{
  assert(a.cols() == b.rows());
  MatrixXd o(a.rows(), b.cols());
  for (Index j = 0; j < b.cols(); j++)
    for (Index i = 0; i < a.rows(); i++)
    {
      double s = 0.0;
      for (Index k = 0; k < a.cols(); k++)
        s += a(i, k) * b(k, j);
      o(i, j) = s;
    }
  return o;
}
inline MatrixXd operator*(double s, const ColXpr& c) { return s * (MatrixXd)c; }
inline MatrixXd operator*(const ColXpr& c, double s) { return (MatrixXd)c * s; }

inline std::ostream& operator<<(std::ostream& os, const MatrixXd& m)
{
  for (Index i = 0; i < m.rows(); i++)
  {
    for (Index j = 0; j < m.cols(); j++)
      os << (j ? " " : "") << m(i, j);
    if (i + 1 < m.rows())
      os << "\n";
  }
  return os;
}

// ---- ColXpr members ----
inline Index ColXpr::size() const { return m_.rows(); }
inline double& ColXpr::operator()(Index i) { return m_(i, c_); }
inline double ColXpr::operator()(Index i) const { return m_(i, c_); }
inline ColXpr& ColXpr::operator=(const MatrixXd& v)
{
  for (Index i = 0; i < size(); i++)
    m_(i, c_) = v(i);
  return *this;
}
inline ColXpr& ColXpr::operator=(const ColXpr& v)
{
  MatrixXd t = v;
  return (*this = t);
}
inline ColXpr& ColXpr::operator-=(const MatrixXd& v)
{
  for (Index i = 0; i < size(); i++)
    m_(i, c_) -= v(i);
  return *this;
}
inline ColXpr& ColXpr::operator/=(double s)
{
  for (Index i = 0; i < size(); i++)
    m_(i, c_) /= s;
  return *this;
}
inline double ColXpr::norm() const
{
  double s = 0.0;
  for (Index i = 0; i < size(); i++)
    s += m_(i, c_) * m_(i, c_);
  return std::sqrt(s);
}
inline double ColXpr::dot(const MatrixXd& v) const
{
  double s = 0.0;
  for (Index i = 0; i < size(); i++)
    s += m_(i, c_) * v(i);
  return s;
}
inline double ColXpr::dot(const ColXpr& v) const
{
  double s = 0.0;
  for (Index i = 0; i < size(); i++)
    s += m_(i, c_) * v(i);
  return s;
}
inline ColXpr::operator MatrixXd() const
{
  MatrixXd o(size(), 1);
  for (Index i = 0; i < size(); i++)
    o(i) = m_(i, c_);
  return o;
}
}  // namespace Eigen
