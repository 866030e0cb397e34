// mini_eigen.h -- TEST INFRASTRUCTURE.  A small, eager stand-in for the subset of Eigen3 that rosdyn_core's headers use.
//
// Why it exists: the reference (CNR-STIIMA-IRAS/rosdyn, rosdyn_core) is header-only C++ on top of Eigen3, urdfdom and roscpp, none of
// which is installed in this image (no network).  With this directory first on the include path the reference's OWN sources
// (primitives.h / internal/primitives_impl.h / spacevect_algebra.h / urdf_parser.h / friction_polynomial*.h / ideal_spring.h, compiled
// where they lie under /root/reference) build into oracle/_ref/librosdyn_ref.so, so the restatement oracle and the CUDA engine can be
// checked against the reference's real code path.  What is NOT the reference here is the dense arithmetic Eigen would perform: every
// product / sum / cross / transpose below is a plain nested loop in the textbook order (no expression templates, no vectorisation, no
// FMA contraction beyond what the compiler does), exact up to rounding order.  Nothing under rosdyn_b200/ includes this file.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3

namespace Eigen
{
constexpr int Dynamic = -1;
enum
{
  ColMajor = 0,
  RowMajor = 1
};
typedef std::ptrdiff_t Index;

template <class T>
using aligned_allocator = std::allocator<T>;

template <class T, int R, int C, int O = 0, int MR = R, int MC = C>
class Matrix;
template <class P, int BR, int BC>
class Block;
template <class D>
struct MatrixBase;
template <class D>
class CommaInit;

template <class D>
struct traits;
template <class T, int R, int C, int O, int MR, int MC>
struct traits<Matrix<T, R, C, O, MR, MC>>
{
  static constexpr int Rows = R, Cols = C;
};
template <class P, int BR, int BC>
struct traits<Block<P, BR, BC>>
{
  static constexpr int Rows = BR, Cols = BC;
};

constexpr int pick(int a, int b) { return a != Dynamic ? a : b; }

// ------------------------------------------------------------------------------------------------ base (CRTP)
template <class D>
struct MatrixBase
{
  static constexpr int RowsAtCompileTime = traits<D>::Rows, ColsAtCompileTime = traits<D>::Cols;
  D& derived() { return *static_cast<D*>(this); }
  const D& derived() const { return *static_cast<const D*>(this); }
  Index rows() const { return derived().rows_(); }
  Index cols() const { return derived().cols_(); }
  Index size() const { return rows() * cols(); }
  double coeff(Index i, Index j) const { return derived().get(i, j); }

  double operator()(Index i, Index j) const { return derived().get(i, j); }
  double& operator()(Index i, Index j) { return derived().ref(i, j); }
  double operator()(Index i) const { return cols() == 1 ? derived().get(i, 0) : derived().get(0, i); }
  double& operator()(Index i) { return cols() == 1 ? derived().ref(i, 0) : derived().ref(0, i); }
  double operator[](Index i) const { return (*this)(i); }
  double& operator[](Index i) { return (*this)(i); }
  double x() const { return (*this)(0); }
  double y() const { return (*this)(1); }
  double z() const { return (*this)(2); }
  double& x() { return (*this)(0); }
  double& y() { return (*this)(1); }
  double& z() { return (*this)(2); }

  D& matrix() { return derived(); }
  const D& matrix() const { return derived(); }

  // ---- blocks: writable proxies on non-const objects, evaluated copies on const ones
  Block<D, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) { return Block<D, Dynamic, Dynamic>(derived(), i, j, r, c); }
  Matrix<double, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) const;
  Block<D, RowsAtCompileTime, 1> col(Index j) { return Block<D, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
  Matrix<double, RowsAtCompileTime, 1> col(Index j) const;
  Block<D, 1, ColsAtCompileTime> row(Index i) { return Block<D, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
  Matrix<double, 1, ColsAtCompileTime> row(Index i) const;
  Block<D, Dynamic, 1> head(Index n) { return Block<D, Dynamic, 1>(derived(), 0, 0, n, 1); }
  Matrix<double, Dynamic, 1> head(Index n) const;
  Block<D, Dynamic, 1> tail(Index n) { return Block<D, Dynamic, 1>(derived(), rows() - n, 0, n, 1); }
  Matrix<double, Dynamic, 1> tail(Index n) const;
  Block<D, RowsAtCompileTime, Dynamic> rightCols(Index n) { return Block<D, RowsAtCompileTime, Dynamic>(derived(), 0, cols() - n, rows(), n); }
  Matrix<double, RowsAtCompileTime, Dynamic> rightCols(Index n) const;

  // ---- in-place
  D& setZero()
  {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) derived().ref(i, j) = 0.0;
    return derived();
  }
  D& setIdentity()
  {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) derived().ref(i, j) = i == j ? 1.0 : 0.0;
    return derived();
  }
  D& setConstant(double v)
  {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) derived().ref(i, j) = v;
    return derived();
  }
  template <class E>
  D& operator+=(const MatrixBase<E>& o)
  {
    check_same(o);
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) derived().ref(i, j) += o.coeff(i, j);
    return derived();
  }
  template <class E>
  D& operator-=(const MatrixBase<E>& o)
  {
    check_same(o);
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) derived().ref(i, j) -= o.coeff(i, j);
    return derived();
  }
  D& operator*=(double s)
  {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) derived().ref(i, j) *= s;
    return derived();
  }
  D& operator/=(double s)
  {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) derived().ref(i, j) /= s;
    return derived();
  }
  template <class E>
  void check_same(const MatrixBase<E>& o) const
  {
    if (rows() != o.rows() || cols() != o.cols()) throw std::logic_error("mini_eigen: size mismatch");
  }

  // ---- reductions / products
  double norm() const { return std::sqrt(squaredNorm()); }
  double squaredNorm() const
  {
    double s = 0;
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) s += coeff(i, j) * coeff(i, j);
    return s;
  }
  template <class E>
  double dot(const MatrixBase<E>& o) const
  {
    if (size() != o.size()) throw std::logic_error("mini_eigen: dot size mismatch");
    double s = 0;
    for (Index i = 0; i < size(); i++) s += (*this)(i) * o(i);
    return s;
  }
  template <class E>
  Matrix<double, 3, 1> cross(const MatrixBase<E>& o) const;
  Matrix<double, ColsAtCompileTime, RowsAtCompileTime> transpose() const;
  Matrix<double, RowsAtCompileTime, ColsAtCompileTime> inverse() const;
  template <class E>
  Matrix<double, pick(RowsAtCompileTime, traits<E>::Rows), pick(ColsAtCompileTime, traits<E>::Cols)> cwiseProduct(const MatrixBase<E>& o) const;
  Matrix<double, Dynamic, Dynamic> asDiagonal() const;
  Matrix<double, RowsAtCompileTime, ColsAtCompileTime> eval() const;

  template <class E>
  bool operator==(const MatrixBase<E>& o) const
  {
    if (rows() != o.rows() || cols() != o.cols()) return false;
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++)
        if (coeff(i, j) != o.coeff(i, j)) return false;
    return true;
  }
  template <class E>
  bool operator!=(const MatrixBase<E>& o) const
  {
    return !(*this == o);
  }

  // ---- comma initialiser
  CommaInit<D> operator<<(double v);
  template <class E>
  CommaInit<D> operator<<(const MatrixBase<E>& m);
};

// ------------------------------------------------------------------------------------------------ storage
template <class T, int R, int C, int O, int MR, int MC>
class Matrix : public MatrixBase<Matrix<T, R, C, O, MR, MC>>
{
  static_assert(std::is_same<T, double>::value, "mini_eigen: double only");
  static constexpr bool kFixed = R != Dynamic && C != Dynamic;
  typedef typename std::conditional<kFixed, double[kFixed ? (R * C > 0 ? R * C : 1) : 1], std::vector<double>>::type Store;
  Store d_;
  Index r_ = R == Dynamic ? 0 : R, c_ = C == Dynamic ? 0 : C;

  void alloc(Index r, Index c)
  {
    if ((R != Dynamic && r != R) || (C != Dynamic && c != C)) throw std::logic_error("mini_eigen: resize of a fixed dimension");
    r_ = r;
    c_ = c;
    if constexpr (!kFixed) d_.assign((size_t)(r * c), 0.0);
  }

public:
  typedef MatrixBase<Matrix> Base;
  using Base::operator();
  Index rows_() const { return r_; }
  Index cols_() const { return c_; }
  double get(Index i, Index j) const { return d_[(size_t)(j * r_ + i)]; }
  double& ref(Index i, Index j) { return d_[(size_t)(j * r_ + i)]; }
  double* data() { return &d_[0]; }
  const double* data() const { return &d_[0]; }

  Matrix()
  {
    if constexpr (kFixed)
      for (int k = 0; k < R * C; k++) d_[k] = 0.0;  // Eigen leaves these uninitialised; zero is a valid instance of "anything"
  }
  explicit Matrix(Index n)
  {
    if (C == 1 || C == Dynamic) alloc(R == Dynamic ? n : R, C == Dynamic ? 1 : C);
    else alloc(1, n);
    if (R != Dynamic && C != Dynamic) { /* fixed: the argument is a size hint only */ }
  }
  Matrix(Index r, Index c) { alloc(r, c); }
  // fixed-size vectors from coefficients
  Matrix(double a, double b, double c)
  {
    alloc(R == Dynamic ? 3 : R, C == Dynamic ? 1 : C);
    (*this)(0) = a;
    (*this)(1) = b;
    (*this)(2) = c;
  }
  Matrix(const Matrix&) = default;
  Matrix& operator=(const Matrix&) = default;
  template <class E>
  Matrix(const MatrixBase<E>& o)
  {
    assign(o);
  }
  template <class E>
  Matrix& operator=(const MatrixBase<E>& o)
  {
    assign(o);
    return *this;
  }
  template <class E>
  void assign(const MatrixBase<E>& o)
  {
    Index r = o.rows(), c = o.cols();
    // Eigen lets a column vector receive a row vector of the same length (and vice versa) only through transposition; the
    // reference never relies on it, so sizes must agree apart from dynamic dimensions.
    if (r_ != r || c_ != c) alloc(r, c);
    for (Index j = 0; j < c; j++)
      for (Index i = 0; i < r; i++) ref(i, j) = o.coeff(i, j);
  }
  void resize(Index n)
  {
    if (C == 1) alloc(n, 1);
    else if (R == 1) alloc(1, n);
    else throw std::logic_error("mini_eigen: resize(n) on a matrix");
  }
  void resize(Index r, Index c) { alloc(r, c); }
  void conservativeResize(Index r, Index c)
  {
    Matrix old(*this);
    alloc(r, c);
    for (Index j = 0; j < std::min(c, old.cols()); j++)
      for (Index i = 0; i < std::min(r, old.rows()); i++) ref(i, j) = old.get(i, j);
  }

  // a 1x1 result converts to its coefficient (Eigen does this for inner products)
  template <class U, typename std::enable_if<std::is_same<U, double>::value && R == 1 && C == 1, int>::type = 0>
  operator U() const
  {
    return get(0, 0);
  }

  static Matrix Zero()
  {
    Matrix m;
    m.setZero();
    return m;
  }
  static Matrix Zero(Index n)
  {
    Matrix m(n);
    m.setZero();
    return m;
  }
  static Matrix Zero(Index r, Index c)
  {
    Matrix m(r, c);
    m.setZero();
    return m;
  }
  static Matrix Identity()
  {
    Matrix m;
    m.setIdentity();
    return m;
  }
  static Matrix Identity(Index r, Index c)
  {
    Matrix m(r, c);
    m.setIdentity();
    return m;
  }
  static Matrix Constant(Index r, Index c, double v)
  {
    Matrix m(r, c);
    m.setConstant(v);
    return m;
  }
  static Matrix Constant(Index n, double v)
  {
    Matrix m(n);
    m.setConstant(v);
    return m;
  }
  static Matrix Unit(int k)
  {
    Matrix m;
    m.setZero();
    m(k) = 1.0;
    return m;
  }
  static Matrix UnitX() { return Unit(0); }
  static Matrix UnitY() { return Unit(1); }
  static Matrix UnitZ() { return Unit(2); }
};

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;

// writable view of a rectangular part of P (P outlives the view; all uses in the reference are within one statement)
template <class P, int BR, int BC>
class Block : public MatrixBase<Block<P, BR, BC>>
{
  P* p_;
  Index i0_, j0_, r_, c_;

public:
  typedef MatrixBase<Block> Base;
  using Base::operator();
  Block(P& p, Index i, Index j, Index r, Index c) : p_(&p), i0_(i), j0_(j), r_(r), c_(c)
  {
    if (i < 0 || j < 0 || r < 0 || c < 0 || i + r > p.rows() || j + c > p.cols()) throw std::out_of_range("mini_eigen: block out of range");
  }
  Block(const Block&) = default;
  Index rows_() const { return r_; }
  Index cols_() const { return c_; }
  double get(Index i, Index j) const { return static_cast<const P*>(p_)->coeff(i0_ + i, j0_ + j); }
  double& ref(Index i, Index j) { return (*p_)(i0_ + i, j0_ + j); }
  template <class E>
  Block& operator=(const MatrixBase<E>& o)
  {
    Matrix<double, Dynamic, Dynamic> tmp(o);  // evaluate first: the right-hand side may alias the parent
    if (tmp.rows() != r_ || tmp.cols() != c_) throw std::logic_error("mini_eigen: block assignment size mismatch");
    for (Index j = 0; j < c_; j++)
      for (Index i = 0; i < r_; i++) ref(i, j) = tmp.get(i, j);
    return *this;
  }
  Block& operator=(const Block& o) { return this->template operator=<Block>(static_cast<const MatrixBase<Block>&>(o)); }
};

// ------------------------------------------------------------------------------------------------ deferred members
template <class D>
Matrix<double, Dynamic, Dynamic> MatrixBase<D>::block(Index i, Index j, Index r, Index c) const
{
  if (i < 0 || j < 0 || i + r > rows() || j + c > cols()) throw std::out_of_range("mini_eigen: block out of range");
  Matrix<double, Dynamic, Dynamic> m(r, c);
  for (Index b = 0; b < c; b++)
    for (Index a = 0; a < r; a++) m(a, b) = coeff(i + a, j + b);
  return m;
}
template <class D>
Matrix<double, MatrixBase<D>::RowsAtCompileTime, 1> MatrixBase<D>::col(Index j) const
{
  return block(0, j, rows(), 1);
}
template <class D>
Matrix<double, 1, MatrixBase<D>::ColsAtCompileTime> MatrixBase<D>::row(Index i) const
{
  return block(i, 0, 1, cols());
}
template <class D>
Matrix<double, Dynamic, 1> MatrixBase<D>::head(Index n) const
{
  return block(0, 0, n, 1);
}
template <class D>
Matrix<double, Dynamic, 1> MatrixBase<D>::tail(Index n) const
{
  return block(rows() - n, 0, n, 1);
}
template <class D>
Matrix<double, MatrixBase<D>::RowsAtCompileTime, Dynamic> MatrixBase<D>::rightCols(Index n) const
{
  return block(0, cols() - n, rows(), n);
}
template <class D>
Matrix<double, MatrixBase<D>::RowsAtCompileTime, MatrixBase<D>::ColsAtCompileTime> MatrixBase<D>::eval() const
{
  return Matrix<double, RowsAtCompileTime, ColsAtCompileTime>(*this);
}
template <class D>
template <class E>
Matrix<double, 3, 1> MatrixBase<D>::cross(const MatrixBase<E>& o) const
{
  if (size() != 3 || o.size() != 3) throw std::logic_error("mini_eigen: cross needs 3-vectors");
  const double a0 = (*this)(0), a1 = (*this)(1), a2 = (*this)(2), b0 = o(0), b1 = o(1), b2 = o(2);
  return Matrix<double, 3, 1>(a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0);
}
template <class D>
Matrix<double, MatrixBase<D>::ColsAtCompileTime, MatrixBase<D>::RowsAtCompileTime> MatrixBase<D>::transpose() const
{
  Matrix<double, ColsAtCompileTime, RowsAtCompileTime> m(cols(), rows());
  for (Index j = 0; j < cols(); j++)
    for (Index i = 0; i < rows(); i++) m(j, i) = coeff(i, j);
  return m;
}
template <class D>
template <class E>
Matrix<double, pick(MatrixBase<D>::RowsAtCompileTime, traits<E>::Rows), pick(MatrixBase<D>::ColsAtCompileTime, traits<E>::Cols)>
MatrixBase<D>::cwiseProduct(const MatrixBase<E>& o) const
{
  check_same(o);
  Matrix<double, pick(RowsAtCompileTime, traits<E>::Rows), pick(ColsAtCompileTime, traits<E>::Cols)> m(rows(), cols());
  for (Index j = 0; j < cols(); j++)
    for (Index i = 0; i < rows(); i++) m(i, j) = coeff(i, j) * o.coeff(i, j);
  return m;
}
template <class D>
Matrix<double, Dynamic, Dynamic> MatrixBase<D>::asDiagonal() const
{
  Matrix<double, Dynamic, Dynamic> m(size(), size());
  for (Index i = 0; i < size(); i++) m(i, i) = (*this)(i);
  return m;
}
// general inverse by Gauss-Jordan elimination with partial pivoting (the reference only inverts rotation matrices, off the hot path)
template <class D>
Matrix<double, MatrixBase<D>::RowsAtCompileTime, MatrixBase<D>::ColsAtCompileTime> MatrixBase<D>::inverse() const
{
  const Index n = rows();
  if (n != cols()) throw std::logic_error("mini_eigen: inverse of a non-square matrix");
  Matrix<double, Dynamic, Dynamic> a(*this), b = Matrix<double, Dynamic, Dynamic>::Identity(n, n);
  for (Index k = 0; k < n; k++)
  {
    Index piv = k;
    for (Index i = k + 1; i < n; i++)
      if (std::fabs(a(i, k)) > std::fabs(a(piv, k))) piv = i;
    for (Index j = 0; j < n; j++)
    {
      std::swap(a(k, j), a(piv, j));
      std::swap(b(k, j), b(piv, j));
    }
    const double d = a(k, k);
    for (Index j = 0; j < n; j++)
    {
      a(k, j) /= d;
      b(k, j) /= d;
    }
    for (Index i = 0; i < n; i++)
      if (i != k)
      {
        const double f = a(i, k);
        for (Index j = 0; j < n; j++)
        {
          a(i, j) -= f * a(k, j);
          b(i, j) -= f * b(k, j);
        }
      }
  }
  return Matrix<double, RowsAtCompileTime, ColsAtCompileTime>(b);
}

// ------------------------------------------------------------------------------------------------ operators
template <class A, class B>
Matrix<double, pick(traits<A>::Rows, traits<B>::Rows), pick(traits<A>::Cols, traits<B>::Cols)> operator+(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
  a.check_same(b);
  Matrix<double, pick(traits<A>::Rows, traits<B>::Rows), pick(traits<A>::Cols, traits<B>::Cols)> m(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) m(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return m;
}
template <class A, class B>
Matrix<double, pick(traits<A>::Rows, traits<B>::Rows), pick(traits<A>::Cols, traits<B>::Cols)> operator-(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
  a.check_same(b);
  Matrix<double, pick(traits<A>::Rows, traits<B>::Rows), pick(traits<A>::Cols, traits<B>::Cols)> m(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) m(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return m;
}
template <class A>
Matrix<double, traits<A>::Rows, traits<A>::Cols> operator-(const MatrixBase<A>& a)
{
  Matrix<double, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) m(i, j) = -a.coeff(i, j);
  return m;
}
template <class A, class B>
Matrix<double, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
  if (a.cols() != b.rows()) throw std::logic_error("mini_eigen: product size mismatch");
  Matrix<double, traits<A>::Rows, traits<B>::Cols> m(a.rows(), b.cols());
  for (Index j = 0; j < b.cols(); j++)
    for (Index i = 0; i < a.rows(); i++)
    {
      double s = 0;
      for (Index k = 0; k < a.cols(); k++) s += a.coeff(i, k) * b.coeff(k, j);
      m(i, j) = s;
    }
  return m;
}
template <class A>
Matrix<double, traits<A>::Rows, traits<A>::Cols> operator*(const MatrixBase<A>& a, double s)
{
  Matrix<double, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) m(i, j) = a.coeff(i, j) * s;
  return m;
}
template <class A>
Matrix<double, traits<A>::Rows, traits<A>::Cols> operator*(double s, const MatrixBase<A>& a)
{
  Matrix<double, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) m(i, j) = s * a.coeff(i, j);
  return m;
}
template <class A>
Matrix<double, traits<A>::Rows, traits<A>::Cols> operator/(const MatrixBase<A>& a, double s)
{
  Matrix<double, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) m(i, j) = a.coeff(i, j) / s;
  return m;
}

// ------------------------------------------------------------------------------------------------ comma initialiser
template <class D>
class CommaInit
{
  D* m_;
  Index row_ = 0, col_ = 0, blockRows_ = 1;

public:
  explicit CommaInit(D& m) : m_(&m) {}
  void put(double v)
  {
    if (col_ == m_->cols())
    {
      row_ += blockRows_;
      col_ = 0;
      blockRows_ = 1;
    }
    (*m_)(row_, col_) = v;
    col_ += 1;
  }
  template <class E>
  void put(const MatrixBase<E>& o)
  {
    if (col_ == m_->cols() || (col_ == 0 && row_ == 0 && false))
    {
      row_ += blockRows_;
      col_ = 0;
    }
    blockRows_ = o.rows();
    for (Index j = 0; j < o.cols(); j++)
      for (Index i = 0; i < o.rows(); i++) (*m_)(row_ + i, col_ + j) = o.coeff(i, j);
    col_ += o.cols();
  }
  CommaInit& operator,(double v)
  {
    put(v);
    return *this;
  }
  template <class E>
  CommaInit& operator,(const MatrixBase<E>& o)
  {
    put(o);
    return *this;
  }
};
template <class D>
CommaInit<D> MatrixBase<D>::operator<<(double v)
{
  CommaInit<D> c(derived());
  c.put(v);
  return c;
}
template <class D>
template <class E>
CommaInit<D> MatrixBase<D>::operator<<(const MatrixBase<E>& m)
{
  CommaInit<D> c(derived());
  c.put(m);
  return c;
}

// ------------------------------------------------------------------------------------------------ Ref
// Ref<T> aliases an existing T (what every call site in the reference passes); anything else is copied into owned storage.
template <class T>
class Ref;
template <class T>
struct traits<Ref<T>>
{
  static constexpr int Rows = traits<T>::Rows, Cols = traits<T>::Cols;
};
template <class T>
class Ref : public MatrixBase<Ref<T>>
{
  T* p_;
  std::shared_ptr<T> own_;

public:
  typedef MatrixBase<Ref> Base;
  using Base::operator();
  Ref(T& t) : p_(&t) {}
  Ref(const T& t) : p_(const_cast<T*>(&t)) {}
  template <class E, typename std::enable_if<!std::is_same<E, T>::value, int>::type = 0>
  Ref(const MatrixBase<E>& o) : own_(std::make_shared<T>(o))
  {
    p_ = own_.get();
  }
  Ref(const Ref&) = default;
  Index rows_() const { return p_->rows(); }
  Index cols_() const { return p_->cols(); }
  double get(Index i, Index j) const { return static_cast<const T*>(p_)->coeff(i, j); }
  double& ref(Index i, Index j) { return (*p_)(i, j); }
  template <class E>
  Ref& operator=(const MatrixBase<E>& o)
  {
    *p_ = o;
    return *this;
  }
};

// ------------------------------------------------------------------------------------------------ geometry
class AngleAxisd;

class Quaterniond
{
  Matrix<double, 4, 1> c_;  // x y z w

public:
  Quaterniond() {}
  Quaterniond(double w, double x, double y, double z)
  {
    c_(0) = x;
    c_(1) = y;
    c_(2) = z;
    c_(3) = w;
  }
  // rotation matrix -> quaternion (the branch structure of Eigen's quaternionbase_assign_impl<Other,3,3>)
  template <class E>
  explicit Quaterniond(const MatrixBase<E>& mat)
  {
    if (mat.rows() != 3 || mat.cols() != 3) throw std::logic_error("mini_eigen: Quaterniond needs a 3x3 matrix");
    double t = mat(0, 0) + mat(1, 1) + mat(2, 2);
    if (t > 0.0)
    {
      t = std::sqrt(t + 1.0);
      c_(3) = 0.5 * t;
      t = 0.5 / t;
      c_(0) = (mat(2, 1) - mat(1, 2)) * t;
      c_(1) = (mat(0, 2) - mat(2, 0)) * t;
      c_(2) = (mat(1, 0) - mat(0, 1)) * t;
    }
    else
    {
      int i = 0;
      if (mat(1, 1) > mat(0, 0)) i = 1;
      if (mat(2, 2) > mat(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + 1.0);
      c_(i) = 0.5 * t;
      t = 0.5 / t;
      c_(3) = (mat(k, j) - mat(j, k)) * t;
      c_(j) = (mat(j, i) + mat(i, j)) * t;
      c_(k) = (mat(k, i) + mat(i, k)) * t;
    }
  }
  double x() const { return c_(0); }
  double y() const { return c_(1); }
  double z() const { return c_(2); }
  double w() const { return c_(3); }
  double& x() { return c_(0); }
  double& y() { return c_(1); }
  double& z() { return c_(2); }
  double& w() { return c_(3); }
  Block<Matrix<double, 4, 1>, 3, 1> vec() { return Block<Matrix<double, 4, 1>, 3, 1>(c_, 0, 0, 3, 1); }
  Matrix<double, 3, 1> vec() const { return Matrix<double, 3, 1>(c_(0), c_(1), c_(2)); }
  double norm() const { return c_.norm(); }
  void normalize() { c_ /= c_.norm(); }
  // Eigen::QuaternionBase::toRotationMatrix, same operation order (no normalisation)
  Matrix3d toRotationMatrix() const
  {
    Matrix3d res;
    const double tx = 2.0 * x(), ty = 2.0 * y(), tz = 2.0 * z();
    const double twx = tx * w(), twy = ty * w(), twz = tz * w();
    const double txx = tx * x(), txy = ty * x(), txz = tz * x();
    const double tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res(0, 0) = 1.0 - (tyy + tzz);
    res(0, 1) = txy - twz;
    res(0, 2) = txz + twy;
    res(1, 0) = txy + twz;
    res(1, 1) = 1.0 - (txx + tzz);
    res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy;
    res(2, 1) = tyz + twx;
    res(2, 2) = 1.0 - (txx + tyy);
    return res;
  }
};

class AngleAxisd
{
  double angle_ = 0;
  Vector3d axis_ = Vector3d::UnitX();

public:
  AngleAxisd() {}
  template <class E>
  AngleAxisd(double angle, const MatrixBase<E>& axis) : angle_(angle), axis_(axis)
  {
  }
  explicit AngleAxisd(const Quaterniond& q) { from_quat(q); }
  template <class E>
  explicit AngleAxisd(const MatrixBase<E>& rot)
  {
    from_quat(Quaterniond(rot));
  }
  void from_quat(const Quaterniond& q)
  {
    double n = q.vec().norm();
    if (n < 2.2250738585072014e-308) n = q.vec().norm();
    if (n != 0.0)
    {
      angle_ = 2.0 * std::atan2(n, std::fabs(q.w()));
      if (q.w() < 0) n = -n;
      axis_ = q.vec() / n;
    }
    else
    {
      angle_ = 0.0;
      axis_ = Vector3d::UnitX();
    }
  }
  double angle() const { return angle_; }
  const Vector3d& axis() const { return axis_; }
  Matrix3d toRotationMatrix() const
  {
    Matrix3d res;
    const double s = std::sin(angle_), c = std::cos(angle_);
    const Vector3d sin_axis = s * axis_;
    const Vector3d cos1_axis = (1.0 - c) * axis_;
    double tmp;
    tmp = cos1_axis.x() * axis_.y();
    res(0, 1) = tmp - sin_axis.z();
    res(1, 0) = tmp + sin_axis.z();
    tmp = cos1_axis.x() * axis_.z();
    res(0, 2) = tmp + sin_axis.y();
    res(2, 0) = tmp - sin_axis.y();
    tmp = cos1_axis.y() * axis_.z();
    res(1, 2) = tmp - sin_axis.x();
    res(2, 1) = tmp + sin_axis.x();
    res(0, 0) = cos1_axis.x() * axis_.x() + c;
    res(1, 1) = cos1_axis.y() * axis_.y() + c;
    res(2, 2) = cos1_axis.z() * axis_.z() + c;
    return res;
  }
};
template <class A>
Matrix3d operator*(const MatrixBase<A>& a, const AngleAxisd& r)
{
  return a * r.toRotationMatrix();
}

enum TransformTraits
{
  Isometry = 1,
  Affine = 2,
  AffineCompact = 3,
  Projective = 4
};

// Transform<double,3,Affine>: a 4x4 homogeneous matrix
class Affine3d
{
  Matrix4d m_ = Matrix4d::Identity();

public:
  Affine3d() {}
  Affine3d(const Affine3d&) = default;
  Affine3d& operator=(const Affine3d&) = default;
  template <class E>
  explicit Affine3d(const MatrixBase<E>& m) : m_(m)
  {
  }
  Affine3d& operator=(const Quaterniond& q)
  {
    m_.setIdentity();
    linear() = q.toRotationMatrix();
    return *this;
  }
  void setIdentity() { m_.setIdentity(); }
  static Affine3d Identity() { return Affine3d(); }
  Matrix4d& matrix() { return m_; }
  const Matrix4d& matrix() const { return m_; }
  Block<Matrix4d, 3, 3> linear() { return Block<Matrix4d, 3, 3>(m_, 0, 0, 3, 3); }
  Matrix3d linear() const { return Matrix3d(m_.block(0, 0, 3, 3)); }
  Block<Matrix4d, 3, 3> rotation() { return linear(); }
  Matrix3d rotation() const { return linear(); }
  Block<Matrix4d, 3, 1> translation() { return Block<Matrix4d, 3, 1>(m_, 0, 3, 3, 1); }
  Vector3d translation() const { return Vector3d(m_.block(0, 3, 3, 1)); }
  Affine3d operator*(const Affine3d& o) const { return Affine3d(m_ * o.m_); }
  template <class E>
  Vector3d operator*(const MatrixBase<E>& v) const
  {
    return linear() * v + translation();
  }
  Affine3d inverse() const
  {
    Affine3d r;
    const Matrix3d li = linear().inverse();
    r.linear() = li;
    r.translation() = -(li * translation());
    return r;
  }
};
typedef Affine3d Isometry3d;

}  // namespace Eigen
