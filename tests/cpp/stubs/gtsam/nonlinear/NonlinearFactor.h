#pragma once
#include "../base.h"
#include "Values.h"
namespace gtsam {
class GaussianFactor {
 public:
  virtual ~GaussianFactor() {}
};
class NonlinearFactor {
 public:
  using shared_ptr = std::shared_ptr<NonlinearFactor>;
  NonlinearFactor() = default;
  explicit NonlinearFactor(const std::vector<Key>& keys) : keys_(keys) {}
  virtual ~NonlinearFactor() {}
  const std::vector<Key>& keys() const { return keys_; }
  virtual size_t dim() const = 0;
  virtual double error(const Values& c) const = 0;
  virtual shared_ptr clone() const = 0;
  virtual std::shared_ptr<GaussianFactor> linearize(const Values& c) const = 0;

 protected:
  std::vector<Key> keys_;
};
}  // namespace gtsam
