// ORACLE — TEST INFRASTRUCTURE ONLY.
// A minimal stand-in for the subset of Eigen 3 that the reference's inertial code uses (common/so3_extra.h, src/Odom/NavState.h, the
// bodies of IMUPreIntegratorBase::update and of the inertial edges of src/Odom/g2otypes.h / g2otypes.cpp).  Eigen itself is a
// third-party dependency of the reference that is absent from this image (no network); this header lets those reference sources
// compile UNCHANGED so the oracle's restatement of them can be checked against the reference's own text.  Fixed-size, column-major,
// every expression evaluated eagerly into a temporary.  Reductions (dot products, norms, product coefficients) add their terms in the
// order of Eigen's unrolled scalar reduction; where Eigen would vectorise or call its GEMM kernels instead (the 9 x 9 covariance
// products, quaternion norms) the order is Eigen's own business, so comparisons in double carry a 1e-12 relative tolerance.
// Quaternion <-> matrix conversions, the quaternion product and _transformVector restate Eigen 3.3's published formulas.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <utility>
#include <list>
#include <ostream>
#include <type_traits>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_DEVICE_FUNC

namespace Eigen {

enum { ComputeFullU = 4, ComputeFullV = 16 };

template <class S, int R, int C>
class Matrix;
template <class P, int R, int C>
class Block;
template <class D>
struct traits;

// Eigen's fully unrolled scalar reduction (redux_novec_unroller): the terms [start, start + len) are summed as
// sum(first half) + sum(second half), e.g. t0 + (t1 + t2) for three terms — the order of sum(), dot(), squaredNorm() and of the
// coefficients of a small fixed-size product
template <class F>
auto redux_halves(const F& term, int start, int len) -> decltype(term(0)) {
  if (len == 1) return term(start);
  const int h = len / 2;
  return redux_halves(term, start, h) + redux_halves(term, start + h, len - h);
}

template <class D>
class MatrixBase {
 public:
  typedef typename traits<D>::Scalar Scalar;
  enum { Rows = traits<D>::Rows, Cols = traits<D>::Cols };
  typedef Matrix<Scalar, Rows, Cols> Plain;
  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  Scalar coeff(int i, int j) const { return derived().coeff(i, j); }
  Scalar operator()(int i, int j) const { return coeff(i, j); }
  Scalar operator()(int i) const {
    static_assert(Cols == 1 || Rows == 1, "vector access");
    return Cols == 1 ? coeff(i, 0) : coeff(0, i);
  }
  Scalar operator[](int i) const { return (*this)(i); }
  Scalar x() const { return (*this)(0); }
  Scalar y() const { return (*this)(1); }
  Scalar z() const { return (*this)(2); }
  int rows() const { return Rows; }
  int cols() const { return Cols; }
  int size() const { return Rows * Cols; }
  const D& matrix() const { return derived(); }
  Matrix<Scalar, Cols, Rows> transpose() const {
    Matrix<Scalar, Cols, Rows> t;
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) t(j, i) = coeff(i, j);
    return t;
  }
  Scalar squaredNorm() const {
    return redux_halves([this](int k) { return coeff(k % Rows, k / Rows) * coeff(k % Rows, k / Rows); }, 0, Rows * Cols);
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Plain normalized() const {  // Eigen 3.3: a zero vector is returned unchanged
    Plain r(*this);
    const Scalar n2 = squaredNorm();
    if (n2 > Scalar(0)) {
      const Scalar n = std::sqrt(n2);
      for (int k = 0; k < Rows * Cols; ++k) r.data()[k] /= n;
    }
    return r;
  }
  template <class O>
  Scalar dot(const MatrixBase<O>& o) const {
    return redux_halves([this, &o](int k) { return (*this)(k) * o(k); }, 0, Rows * Cols);
  }
  template <class O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    return Matrix<Scalar, 3, 1>(y() * o.z() - z() * o.y(), z() * o.x() - x() * o.z(), x() * o.y() - y() * o.x());
  }
  template <class T>
  Matrix<T, Rows, Cols> cast() const {
    Matrix<T, Rows, Cols> r;
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) r(i, j) = (T)coeff(i, j);
    return r;
  }
  bool allFinite() const {
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j)
        if (!std::isfinite(coeff(i, j))) return false;
    return true;
  }
  // fixed-size 3 x 3 inverse the way Eigen computes it (cofactors of the first column -> determinant -> adjugate * 1 / det)
  Plain inverse() const {
    static_assert((int)Rows == 3 && (int)Cols == 3, "only the 3 x 3 inverse is restated");
    auto cof = [this](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return coeff(i1, j1) * coeff(i2, j2) - coeff(i1, j2) * coeff(i2, j1);
    };
    const Scalar c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const Scalar det = c0 * coeff(0, 0) + (c1 * coeff(1, 0) + c2 * coeff(2, 0));
    const Scalar invdet = Scalar(1) / det;
    Plain r;
    r(0, 0) = c0 * invdet; r(0, 1) = c1 * invdet; r(0, 2) = c2 * invdet;
    r(1, 0) = cof(0, 1) * invdet; r(1, 1) = cof(1, 1) * invdet; r(1, 2) = cof(2, 1) * invdet;
    r(2, 0) = cof(0, 2) * invdet; r(2, 1) = cof(1, 2) * invdet; r(2, 2) = cof(2, 2) * invdet;
    return r;
  }
  struct PartialPivLU {  // Gaussian elimination with row pivoting on the largest magnitude (what Eigen's lu() of a fixed-size matrix is)
    Plain lu;
    int perm[Rows];
    template <class B>
    Matrix<Scalar, Rows, traits<B>::Cols> solve(const MatrixBase<B>& b) const {
      Matrix<Scalar, Rows, traits<B>::Cols> x;
      for (int c = 0; c < (int)traits<B>::Cols; ++c) {
        for (int i = 0; i < Rows; ++i) {
          Scalar v = b.coeff(perm[i], c);
          for (int k = 0; k < i; ++k) v -= lu.coeff(i, k) * x(k, c);
          x(i, c) = v;
        }
        for (int i = Rows - 1; i >= 0; --i) {
          Scalar v = x(i, c);
          for (int k = i + 1; k < Rows; ++k) v -= lu.coeff(i, k) * x(k, c);
          x(i, c) = v / lu.coeff(i, i);
        }
      }
      return x;
    }
  };
  PartialPivLU lu() const {
    static_assert((int)Rows == (int)Cols, "square");
    PartialPivLU f;
    f.lu = *this;
    for (int i = 0; i < Rows; ++i) f.perm[i] = i;
    for (int k = 0; k < Rows; ++k) {
      int piv = k;
      for (int i = k + 1; i < Rows; ++i)
        if (std::abs(f.lu.coeff(i, k)) > std::abs(f.lu.coeff(piv, k))) piv = i;
      if (piv != k) {
        for (int j = 0; j < Cols; ++j) std::swap(f.lu.coeffRef(k, j), f.lu.coeffRef(piv, j));
        std::swap(f.perm[k], f.perm[piv]);
      }
      for (int i = k + 1; i < Rows; ++i) {
        const Scalar m = f.lu.coeffRef(i, k) /= f.lu.coeff(k, k);
        for (int j = k + 1; j < Cols; ++j) f.lu.coeffRef(i, j) -= m * f.lu.coeff(k, j);
      }
    }
    return f;
  }
  // read-only sub-blocks are copies
  template <int N>
  Matrix<Scalar, N, 1> segment(int i) const {
    Matrix<Scalar, N, 1> r;
    for (int k = 0; k < N; ++k) r(k) = (*this)(i + k);
    return r;
  }
  template <int BR, int BC>
  Matrix<Scalar, BR, BC> block(int i, int j) const {
    Matrix<Scalar, BR, BC> r;
    for (int a = 0; a < BR; ++a)
      for (int b = 0; b < BC; ++b) r(a, b) = coeff(i + a, j + b);
    return r;
  }
};

// writable layer: needs D::coeffRef(i, j)
template <class D>
class CommaInit {
  D& m_;
  int k_ = 0;

 public:
  CommaInit(D& m, typename traits<D>::Scalar v) : m_(m) { put(v); }
  void put(typename traits<D>::Scalar v) {
    const int C = traits<D>::Cols;
    assert(k_ < traits<D>::Rows * C);
    m_.coeffRef(k_ / C, k_ % C) = v;  // row by row, whatever the storage order
    ++k_;
  }
  CommaInit& operator,(typename traits<D>::Scalar v) {
    put(v);
    return *this;
  }
  D finished() const { return m_; }
};

// a run-time-sized right-hand side (what Eigen::MatrixXd is to the reference's getHessian() members): any type deriving from this tag
// with rows(), cols(), coeff(i, j) can be assigned / added to a fixed-size matrix or block of the same size
struct DynamicRhsTag {};

template <class D>
class Writable : public MatrixBase<D> {
 public:
  template <class X, class = typename std::enable_if<std::is_base_of<DynamicRhsTag, X>::value>::type>
  D& assignDyn(const X& m, bool add) {
    assert(m.rows() == (int)MatrixBase<D>::Rows && m.cols() == (int)MatrixBase<D>::Cols);
    for (int i = 0; i < (int)MatrixBase<D>::Rows; ++i)
      for (int j = 0; j < (int)MatrixBase<D>::Cols; ++j)
        this->derived().coeffRef(i, j) = add ? this->derived().coeff(i, j) + m.coeff(i, j) : m.coeff(i, j);
    return this->derived();
  }
  template <class X, class = typename std::enable_if<std::is_base_of<DynamicRhsTag, X>::value>::type>
  D& operator+=(const X& m) {
    return assignDyn(m, true);
  }
  typedef MatrixBase<D> Base;
  typedef typename Base::Scalar Scalar;
  using Base::coeff;
  using Base::derived;
  using Base::operator();
  using Base::operator[];
  using Base::x;
  using Base::y;
  using Base::z;
  using Base::segment;
  using Base::block;
  enum { Rows = Base::Rows, Cols = Base::Cols };
  Scalar& operator()(int i, int j) { return derived().coeffRef(i, j); }
  Scalar& operator()(int i) { return Cols == 1 ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
  Scalar& operator[](int i) { return (*this)(i); }
  Scalar& x() { return (*this)(0); }
  Scalar& y() { return (*this)(1); }
  Scalar& z() { return (*this)(2); }
  template <class O>
  D& assign(const MatrixBase<O>& o) {
    static_assert((int)traits<O>::Rows == (int)Rows && (int)traits<O>::Cols == (int)Cols, "size mismatch");
    typename Base::Plain t;  // the right-hand side may alias the destination
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) t.coeffRef(i, j) = o.coeff(i, j);
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) derived().coeffRef(i, j) = t.coeff(i, j);
    return derived();
  }
  D& setZero() {
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) derived().coeffRef(i, j) = Scalar(0);
    return derived();
  }
  D& fill(Scalar v) {
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) derived().coeffRef(i, j) = v;
    return derived();
  }
  Block<D, Rows, 1> col(int j) { return Block<D, Rows, 1>(derived(), 0, j); }
  D& setIdentity() {
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) derived().coeffRef(i, j) = Scalar(i == j);
    return derived();
  }
  template <class O>
  D& operator+=(const MatrixBase<O>& o) {
    return assign(derived() + o);
  }
  template <class O>
  D& operator-=(const MatrixBase<O>& o) {
    return assign(derived() - o);
  }
  D& operator*=(Scalar s) {
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) derived().coeffRef(i, j) *= s;
    return derived();
  }
  template <class O>
  D& operator*=(const MatrixBase<O>& o) {
    return assign(derived() * o);
  }
  D& transposeInPlace() {
    static_assert((int)Rows == (int)Cols, "square");
    return assign(this->transpose());
  }
  D& operator/=(Scalar s) {
    for (int i = 0; i < Rows; ++i)
      for (int j = 0; j < Cols; ++j) derived().coeffRef(i, j) /= s;
    return derived();
  }
  CommaInit<D> operator<<(Scalar v) { return CommaInit<D>(derived(), v); }
  template <int N>
  Block<D, N, 1> segment(int i) {
    static_assert(Cols == 1, "segment of a column vector");
    return Block<D, N, 1>(derived(), i, 0);
  }
  template <int BR, int BC>
  Block<D, BR, BC> block(int i, int j) {
    return Block<D, BR, BC>(derived(), i, j);
  }
};

template <class S, int R, int C>
struct traits<Matrix<S, R, C>> {
  typedef S Scalar;
  enum { Rows = R, Cols = C };
};
template <class P, int R, int C>
struct traits<Block<P, R, C>> {
  typedef typename traits<P>::Scalar Scalar;
  enum { Rows = R, Cols = C };
};

template <class S, int R, int C>
class Matrix : public Writable<Matrix<S, R, C>> {
  S m_[R * C];

 public:
  typedef S Scalar;
  typedef Writable<Matrix<S, R, C>> W;
  using W::operator();
  Matrix() {
    for (int k = 0; k < R * C; ++k) m_[k] = S(0);
  }
  Matrix(S a, S b) {
    static_assert(R * C == 2, "two coefficients");
    m_[0] = a, m_[1] = b;
  }
  Matrix(S a, S b, S c) {
    static_assert(R * C == 3, "three coefficients");
    m_[0] = a, m_[1] = b, m_[2] = c;
  }
  explicit Matrix(const S* p) {
    for (int k = 0; k < R * C; ++k) m_[k] = p[k];
  }
  template <class O>
  Matrix(const MatrixBase<O>& o) {
    this->assign(o);
  }
  template <class O>
  Matrix& operator=(const MatrixBase<O>& o) {
    return this->assign(o);
  }
  S coeff(int i, int j) const {
    assert(i >= 0 && i < R && j >= 0 && j < C);
    return m_[j * R + i];
  }
  S& coeffRef(int i, int j) {
    assert(i >= 0 && i < R && j >= 0 && j < C);
    return m_[j * R + i];
  }
  // a 1 x 1 result (an inner product written as a matrix product) converts to its scalar
  template <class T, class = typename std::enable_if<std::is_same<T, S>::value && R == 1 && C == 1>::type>
  operator T() const {
    return m_[0];
  }
  S* data() { return m_; }
  const S* data() const { return m_; }
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() {
    Matrix m;
    m.setIdentity();
    return m;
  }
};

template <class P, int R, int C>
class Block : public Writable<Block<P, R, C>> {
  P& p_;
  int i0_, j0_;

 public:
  typedef typename traits<P>::Scalar Scalar;
  Block(P& p, int i0, int j0) : p_(p), i0_(i0), j0_(j0) {
    assert(i0 >= 0 && i0 + R <= (int)traits<P>::Rows && j0 >= 0 && j0 + C <= (int)traits<P>::Cols);
  }
  Scalar coeff(int i, int j) const { return const_cast<const P&>(p_).coeff(i0_ + i, j0_ + j); }
  Scalar& coeffRef(int i, int j) { return p_.coeffRef(i0_ + i, j0_ + j); }
  template <class O>
  Block& operator=(const MatrixBase<O>& o) {
    return this->assign(o);
  }
  Block& operator=(const Block& o) { return this->assign(o); }
  Block& operator=(const Matrix<Scalar, R, C>& o) { return this->assign(o); }
  template <class X, class = typename std::enable_if<std::is_base_of<DynamicRhsTag, X>::value>::type>
  Block& operator=(const X& m) {
    return this->assignDyn(m, false);
  }
};

// read-only view of a raw array (the only form the reference's vertex updates use)
template <class M>
class Map;
template <class S, int R, int C>
struct traits<Map<const Matrix<S, R, C>>> {
  typedef S Scalar;
  enum { Rows = R, Cols = C };
};
template <class S, int R, int C>
class Map<const Matrix<S, R, C>> : public MatrixBase<Map<const Matrix<S, R, C>>> {
  const S* p_;

 public:
  explicit Map(const S* p) : p_(p) {}
  S coeff(int i, int j) const { return p_[j * R + i]; }
};

template <class S, int R, int C>
struct traits<Map<Matrix<S, R, C>>> {
  typedef S Scalar;
  enum { Rows = R, Cols = C };
};
template <class S, int R, int C>
class Map<Matrix<S, R, C>> : public Writable<Map<Matrix<S, R, C>>> {
  S* p_;

 public:
  explicit Map(S* p) : p_(p) {}
  S coeff(int i, int j) const { return p_[j * R + i]; }
  S& coeffRef(int i, int j) { return p_[j * R + i]; }
};

template <class D>
std::ostream& operator<<(std::ostream& os, const MatrixBase<D>& m) {
  for (int i = 0; i < (int)traits<D>::Rows; ++i) {
    for (int j = 0; j < (int)traits<D>::Cols; ++j) os << (j ? " " : "") << m.coeff(i, j);
    if (i + 1 < (int)traits<D>::Rows) os << "\n";
  }
  return os;
}

// ---- arithmetic (eager) ------------------------------------------------------------------------------------------------------
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  static_assert((int)traits<A>::Rows == (int)traits<B>::Rows && (int)traits<A>::Cols == (int)traits<B>::Cols, "size mismatch");
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> r;
  for (int i = 0; i < (int)traits<A>::Rows; ++i)
    for (int j = 0; j < (int)traits<A>::Cols; ++j) r(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  static_assert((int)traits<A>::Rows == (int)traits<B>::Rows && (int)traits<A>::Cols == (int)traits<B>::Cols, "size mismatch");
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> r;
  for (int i = 0; i < (int)traits<A>::Rows; ++i)
    for (int j = 0; j < (int)traits<A>::Cols; ++j) r(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator-(const MatrixBase<A>& a) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> r;
  for (int i = 0; i < (int)traits<A>::Rows; ++i)
    for (int j = 0; j < (int)traits<A>::Cols; ++j) r(i, j) = -a.coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  static_assert((int)traits<A>::Cols == (int)traits<B>::Rows, "inner dimensions");
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> r;
  for (int i = 0; i < (int)traits<A>::Rows; ++i)
    for (int j = 0; j < (int)traits<B>::Cols; ++j) {
      r(i, j) = redux_halves([&a, &b, i, j](int k) { return a.coeff(i, k) * b.coeff(k, j); }, 0, (int)traits<A>::Cols);
    }
  return r;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator*(const MatrixBase<A>& a, typename traits<A>::Scalar s) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> r;
  for (int i = 0; i < (int)traits<A>::Rows; ++i)
    for (int j = 0; j < (int)traits<A>::Cols; ++j) r(i, j) = a.coeff(i, j) * s;
  return r;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator*(typename traits<A>::Scalar s, const MatrixBase<A>& a) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> r;
  for (int i = 0; i < (int)traits<A>::Rows; ++i)
    for (int j = 0; j < (int)traits<A>::Cols; ++j) r(i, j) = s * a.coeff(i, j);
  return r;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator/(const MatrixBase<A>& a, typename traits<A>::Scalar s) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> r;
  for (int i = 0; i < (int)traits<A>::Rows; ++i)
    for (int j = 0; j < (int)traits<A>::Cols; ++j) r(i, j) = a.coeff(i, j) / s;
  return r;
}

// 1 x 1 results used as scalars
template <class S>
S operator+(const Matrix<S, 1, 1>& a, typename std::common_type<S>::type b) {
  return a.coeff(0, 0) + b;
}
template <class S>
S operator+(typename std::common_type<S>::type a, const Matrix<S, 1, 1>& b) {
  return a + b.coeff(0, 0);
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 3, 3> Matrix3f;

// ---- quaternion (coefficients stored x, y, z, w like Eigen) ------------------------------------------------------------------
template <class D>
struct QuaternionBase {
  const D& derived() const { return *static_cast<const D*>(this); }
};
template <class S, int Options = 0>
class Quaternion : public QuaternionBase<Quaternion<S, Options>> {
  Matrix<S, 4, 1> c_;

 public:
  typedef S Scalar;
  Quaternion() {}
  Quaternion(S w, S x, S y, S z) {
    c_(0) = x, c_(1) = y, c_(2) = z, c_(3) = w;
  }
  template <class D>
  Quaternion(const QuaternionBase<D>& o) : c_(o.derived().coeffs()) {}
  // Eigen 3.3 quaternionbase_assign_impl<Other, 3, 3>
  explicit Quaternion(const Matrix<S, 3, 3>& m) {
    S t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > S(0)) {
      t = std::sqrt(t + S(1.0));
      c_(3) = S(0.5) * t;
      t = S(0.5) / t;
      c_(0) = (m(2, 1) - m(1, 2)) * t;
      c_(1) = (m(0, 2) - m(2, 0)) * t;
      c_(2) = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + S(1.0));
      c_(i) = S(0.5) * t;
      t = S(0.5) / t;
      c_(3) = (m(k, j) - m(j, k)) * t;
      c_(j) = (m(j, i) + m(i, j)) * t;
      c_(k) = (m(k, i) + m(i, k)) * t;
    }
  }
  S x() const { return c_(0); }
  S y() const { return c_(1); }
  S z() const { return c_(2); }
  S w() const { return c_(3); }
  S& x() { return c_(0); }
  S& y() { return c_(1); }
  S& z() { return c_(2); }
  S& w() { return c_(3); }
  Matrix<S, 3, 1> vec() const { return Matrix<S, 3, 1>(c_(0), c_(1), c_(2)); }
  const Matrix<S, 4, 1>& coeffs() const { return c_; }
  Matrix<S, 4, 1>& coeffs() { return c_; }
  S squaredNorm() const { return c_.squaredNorm(); }
  S norm() const { return c_.norm(); }
  void normalize() { c_ /= c_.norm(); }
  Quaternion normalized() const {
    Quaternion q(*this);
    q.normalize();
    return q;
  }
  Quaternion& setIdentity() {
    c_(0) = c_(1) = c_(2) = S(0), c_(3) = S(1);
    return *this;
  }
  static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }
  Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
  Quaternion operator*(const Quaternion& b) const {
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaternion& operator*=(const Quaternion& b) {
    *this = *this * b;
    return *this;
  }
  template <class V>
  Matrix<S, 3, 1> _transformVector(const MatrixBase<V>& v) const {
    Matrix<S, 3, 1> uv = vec().cross(v);
    uv += uv;
    return v + w() * uv + vec().cross(uv);
  }
  template <class V>
  Matrix<S, 3, 1> operator*(const MatrixBase<V>& v) const {
    return _transformVector(v);
  }
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> r;
    const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
    const S twx = tx * w(), twy = ty * w(), twz = tz * w();
    const S txx = tx * x(), txy = ty * x(), txz = tz * x();
    const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    r(0, 0) = S(1) - (tyy + tzz);
    r(0, 1) = txy - twz;
    r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;
    r(1, 1) = S(1) - (txx + tzz);
    r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;
    r(2, 1) = tyz + twx;
    r(2, 2) = S(1) - (txx + tyy);
    return r;
  }
  template <class T>
  Quaternion<T> cast() const {
    return Quaternion<T>((T)w(), (T)x(), (T)y(), (T)z());
  }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

// named by common/so3_extra.h in a function nothing on this path calls
template <class M>
struct JacobiSVD {
  JacobiSVD(const M&, unsigned) { std::abort(); }
  M matrixU() const { return M(); }
  M matrixV() const { return M(); }
};

template <class T>
using aligned_list = std::list<T>;

}  // namespace Eigen
