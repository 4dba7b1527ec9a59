#pragma once
#include "../base.h"
namespace gtsam {
class Unit3 {
 public:
  Unit3() = default;
  Unit3(double x, double y, double z) { v_(0) = x, v_(1) = y, v_(2) = z; }
  Vector3 unitVector() const { return v_; }

 private:
  Vector3 v_;
};
}  // namespace gtsam
