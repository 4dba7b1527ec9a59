// MINIMAL STAND-IN for pcl::PointCloud (test infrastructure): the members the factor adapter reads.
#pragma once
#include <cstddef>
#include <vector>
namespace pcl {
template <typename PointT>
class PointCloud {
 public:
  std::vector<PointT> points;
  size_t size() const { return points.size(); }
  void push_back(const PointT& p) { points.push_back(p); }
};
}  // namespace pcl
