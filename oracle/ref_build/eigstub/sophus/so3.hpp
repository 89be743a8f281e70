// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the part of Sophus::SO3 (third-party, absent from this image; the reference builds with USE_SOPHUS_NEWEST, i.e. its
// SO3ex derives from Sophus' quaternion-backed SO3) that common/so3_extra.h and the inertial code reach: construction from a
// quaternion (normalised), group product with Sophus' first-order renormalisation, action on a point, inverse, matrix(), hat(),
// cast() and Sophus' atan-based log().  Restated from the library's published algorithms; see ../mini_eigen.h for the tolerance note.
#pragma once
#include "../mini_eigen.h"

#define SOPHUS_FUNC

namespace Sophus {

template <class D>
struct SO3Base {
  const D& derived() const { return *static_cast<const D*>(this); }
};

template <class Scalar_, int Options = 0>
class SO3 : public SO3Base<SO3<Scalar_, Options>> {
 public:
  typedef Scalar_ Scalar;
  typedef Eigen::Quaternion<Scalar, Options> QuaternionMember;
  typedef Eigen::Matrix<Scalar, 3, 1> Tangent;
  typedef Eigen::Matrix<Scalar, 3, 1> Point;
  typedef Eigen::Matrix<Scalar, 3, 3> Transformation;
  template <class Other>
  using ReturnScalar = Scalar;
  template <class Other>
  using SO3Product = SO3<Scalar>;
  template <class PointDerived>
  using PointProduct = Eigen::Matrix<Scalar, 3, 1>;

  SO3() { unit_quaternion_.setIdentity(); }
  SO3(const SO3& o) = default;
  SO3& operator=(const SO3& o) = default;
  template <class D>
  SO3(const SO3Base<D>& o) : unit_quaternion_(o.derived().unit_quaternion()) {}
  template <class D>
  explicit SO3(const Eigen::QuaternionBase<D>& q) : unit_quaternion_(q) {
    normalize();
  }
  explicit SO3(const Transformation& R) : unit_quaternion_(R) {}

  const QuaternionMember& unit_quaternion() const { return unit_quaternion_; }
  void normalize() {
    const Scalar length = unit_quaternion_.norm();
    unit_quaternion_.coeffs() /= length;
  }
  template <class T>
  SO3<T> cast() const {
    return SO3<T>(unit_quaternion_.template cast<T>());
  }
  SO3 inverse() const { return SO3(unit_quaternion_.conjugate()); }
  Transformation matrix() const { return unit_quaternion_.toRotationMatrix(); }
  template <class D>
  SO3 operator*(const SO3Base<D>& o) const {
    const QuaternionMember &a = unit_quaternion_, &b = o.derived().unit_quaternion();
    SO3 r;
    r.unit_quaternion_ = QuaternionMember(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                                          a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                                          a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    const Scalar sq = r.unit_quaternion_.squaredNorm();
    if (sq != Scalar(1.0)) {  // first-order renormalisation
      const Scalar scale = Scalar(2.0) / (Scalar(1.0) + sq);
      r.unit_quaternion_.coeffs() *= scale;
    }
    return r;
  }
  template <class D>
  SO3& operator*=(const SO3Base<D>& o) {
    unit_quaternion_ = (*this * o).unit_quaternion_;
    return *this;
  }
  template <class P>
  Point operator*(const Eigen::MatrixBase<P>& p) const {
    Point uv = unit_quaternion_.vec().cross(p);
    uv += uv;
    return p + unit_quaternion_.w() * uv + unit_quaternion_.vec().cross(uv);
  }
  static Transformation hat(const Tangent& omega) {
    Transformation Omega;
    Omega << Scalar(0), -omega(2), omega(1), omega(2), Scalar(0), -omega(0), -omega(1), omega(0), Scalar(0);
    return Omega;
  }
  Tangent log() const {
    const Scalar eps = Scalar(1e-10);
    const Scalar squared_n = unit_quaternion_.vec().squaredNorm();
    const Scalar w = unit_quaternion_.w();
    Scalar two_atan_nbyw_by_n;
    if (squared_n < eps * eps) {
      const Scalar squared_w = w * w;
      two_atan_nbyw_by_n = Scalar(2) / w - Scalar(2.0 / 3.0) * squared_n / (w * squared_w);
    } else {
      const Scalar n = std::sqrt(squared_n);
      if (std::abs(w) < eps)
        two_atan_nbyw_by_n = (w > Scalar(0) ? Scalar(M_PI) : Scalar(-M_PI)) / n;
      else
        two_atan_nbyw_by_n = Scalar(2) * std::atan(n / w) / n;
    }
    return two_atan_nbyw_by_n * unit_quaternion_.vec();
  }

 protected:
  QuaternionMember& unit_quaternion_nonconst() { return unit_quaternion_; }
  QuaternionMember unit_quaternion_;
};

}  // namespace Sophus
