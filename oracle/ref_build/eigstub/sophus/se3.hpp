// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the part of Sophus::SE3 (third-party, absent from this image) the compiled reference functions reach: (SO3, translation)
// with Sophus' product, inverse, cast, action on a point, rotationMatrix() and translation().  Restated from the library's published
// formulas over the SO3 stand-in beside it.
#pragma once
#include "so3.hpp"

namespace Sophus {
template <class T>
class SE3 {  // sophus/se3.hpp: (SO3, translation); T1 * T2 = (R1 R2, t1 + R1 t2); T^-1 = (R^-1, R^-1 * (t * -1))
 public:
  SO3<T> so3_;
  Eigen::Matrix<T, 3, 1> t_;
  SE3() {
    for (int i = 0; i < 3; ++i) t_(i) = T(0);
  }
  SE3(const SO3<T>& r, const Eigen::Matrix<T, 3, 1>& t) : so3_(r), t_(t) {}
  SE3 inverse() const {
    const SO3<T> inv = so3_.inverse();
    const Eigen::Matrix<T, 3, 1> nt = t_ * T(-1);
    return SE3(inv, inv * nt);
  }
  SE3 operator*(const SE3& o) const {
    const Eigen::Matrix<T, 3, 1> rt = so3_ * o.t_;
    return SE3(so3_ * o.so3_, t_ + rt);
  }
  template <class U>
  SE3<U> cast() const {
    return SE3<U>(so3_.template cast<U>(), t_.template cast<U>());
  }
  Eigen::Matrix<T, 3, 3> rotationMatrix() const { return so3_.matrix(); }
  const Eigen::Matrix<T, 3, 1>& translation() const { return t_; }
  template <class P>
  Eigen::Matrix<T, 3, 1> operator*(const Eigen::MatrixBase<P>& p) const {
    const Eigen::Matrix<T, 3, 1> rp = so3_ * p;
    return rp + t_;
  }
};
typedef SE3<double> SE3d;
}  // namespace Sophus
