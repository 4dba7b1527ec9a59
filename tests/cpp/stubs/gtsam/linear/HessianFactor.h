#pragma once
#include "../base.h"
#include "../nonlinear/NonlinearFactor.h"
namespace gtsam {
// HessianFactor(j, G, g, f): error 0.5 x^T G x - x^T g + 0.5 f (gtsam/linear/HessianFactor.h, unary constructor)
class HessianFactor : public GaussianFactor {
 public:
  HessianFactor(Key j, const Matrix& G, const Vector& g, double f) : key_(j), G_(G), g_(g), f_(f) {}
  Key key() const { return key_; }
  const Matrix& information() const { return G_; }
  const Vector& linearTerm() const { return g_; }
  double constantTerm() const { return f_; }

 private:
  Key key_;
  Matrix G_;
  Vector g_;
  double f_;
};
}  // namespace gtsam
