// MINIMAL STAND-INS for the few GTSAM types the factor adapter touches (test infrastructure; no GTSAM exists in this
// image).  Only the members the adapter and its test call; element access like Eigen's: m(i, j), v(i).
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <typeindex>
#include <vector>

namespace gtsam {
using Key = std::uint64_t;

class Matrix {
 public:
  Matrix() = default;
  Matrix(size_t r, size_t c) : r_(r), c_(c), d_(r * c, 0.0) {}
  double& operator()(size_t i, size_t j) { return d_[i * c_ + j]; }
  double operator()(size_t i, size_t j) const { return d_[i * c_ + j]; }
  size_t rows() const { return r_; }
  size_t cols() const { return c_; }

 private:
  size_t r_ = 0, c_ = 0;
  std::vector<double> d_;
};
class Vector {
 public:
  Vector() = default;
  explicit Vector(size_t n) : d_(n, 0.0) {}
  double& operator()(size_t i) { return d_[i]; }
  double operator()(size_t i) const { return d_[i]; }
  size_t size() const { return d_.size(); }

 private:
  std::vector<double> d_;
};
struct Matrix3 {
  std::array<double, 9> d{1, 0, 0, 0, 1, 0, 0, 0, 1};
  double& operator()(size_t i, size_t j) { return d[3 * i + j]; }
  double operator()(size_t i, size_t j) const { return d[3 * i + j]; }
};
struct Vector3 {
  std::array<double, 3> d{0, 0, 0};
  double& operator()(size_t i) { return d[i]; }
  double operator()(size_t i) const { return d[i]; }
};
}  // namespace gtsam
