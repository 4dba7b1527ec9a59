#pragma once
#include "../base.h"
namespace gtsam {
class Rot3 {
 public:
  Rot3() = default;
  explicit Rot3(const Matrix3& m) : m_(m) {}
  const Matrix3& matrix() const { return m_; }

 private:
  Matrix3 m_;
};
class Pose3 {
 public:
  Pose3() = default;
  Pose3(const Rot3& r, const Vector3& t) : r_(r), t_(t) {}
  const Rot3& rotation() const { return r_; }
  const Vector3& translation() const { return t_; }

 private:
  Rot3 r_;
  Vector3 t_;
};
}  // namespace gtsam
