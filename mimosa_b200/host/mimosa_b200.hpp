// mimosa_b200.hpp — header-only C++17 mirror of mimosa's LiDAR geometric-factor interface over the C ABI
// (include/mimosa_b200.h).  Same class roles, method names and error behaviour (exceptions) as the reference:
//
//   mimosa_b200::RegistrationConfig       mimosa/include/mimosa/lidar/geometric_config.hpp:17-33
//   mimosa_b200::IncrementalVoxelMapB200  mimosa::lidar::IncrementalVoxelMapPCL, incremental_voxel_map.hpp:22-51
//   mimosa_b200::ICPFactorB200            mimosa::lidar::ICPFactor (unary), geometric_factor.hpp:25-563
//
// It depends on nothing but the C ABI and the standard library; the GTSAM / PCL glue a mimosa build adds on top
// (deriving from gtsam::NonlinearFactor, taking pcl::PointCloud<Point>) is shown in INTEGRATION.md.
// All compute runs in libmimosa_b200.so on the GPU.  Exceptions replace the reference's logCriticalException
// (mimosa/include/mimosa/utils.hpp:300-306).
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mimosa_b200.h"

namespace mimosa_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) {
  if (rc != MB_OK) throw Error(rc, std::string("mimosa_b200: ") + mb_last_error());
}

// Field-for-field the reference struct (defaults included), convertible to the ABI POD.
struct RegistrationConfig {
  float source_voxel_grid_filter_leaf_size = 0.5f;
  float source_voxel_grid_min_dist_in_voxel = 0.1f;
  float target_ivox_map_leaf_size = 0.5f;
  float target_ivox_map_min_dist_in_voxel = 0.1f;
  size_t num_corres_points = 5;
  float max_corres_distance = 2.24f;
  float plane_validity_distance = 0.04f;
  float lidar_point_noise_std_dev = 0.02f;
  bool use_huber = true;
  float huber_threshold = 1.345f;
  bool reg_4_dof = false;
  bool project_on_degneneracy = true;
  float degen_thresh_rot = 10.f;
  float degen_thresh_trans = 15.f;

  mb_icp_config to_c() const {
    mb_icp_config c{};
    c.source_voxel_grid_filter_leaf_size = source_voxel_grid_filter_leaf_size;
    c.source_voxel_grid_min_dist_in_voxel = source_voxel_grid_min_dist_in_voxel;
    c.target_ivox_map_leaf_size = target_ivox_map_leaf_size;
    c.target_ivox_map_min_dist_in_voxel = target_ivox_map_min_dist_in_voxel;
    c.num_corres_points = num_corres_points;
    c.max_corres_distance = max_corres_distance;
    c.plane_validity_distance = plane_validity_distance;
    c.lidar_point_noise_std_dev = lidar_point_noise_std_dev;
    c.use_huber = use_huber;
    c.huber_threshold = huber_threshold;
    c.reg_4_dof = reg_4_dof;
    c.project_on_degneneracy = project_on_degneneracy;
    c.degen_thresh_rot = degen_thresh_rot;
    c.degen_thresh_trans = degen_thresh_trans;
    return c;
  }
};

// 32-byte record with mimosa::lidar::Point's layout (mimosa/include/mimosa/lidar/point.hpp:18-39).
struct alignas(16) Point {
  float x, y, z, pad;
  float intensity;
  uint32_t t, idx;
  float range;
};
static_assert(sizeof(Point) == MB_POINT_STRIDE, "lidar::Point is 32 bytes");

class Context {
 public:
  explicit Context(int device = 0) { check(mb_init(device, &ctx_)); }
  ~Context() { mb_shutdown(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  mb_ctx* get() const { return ctx_; }
  void sync() const { check(mb_sync(ctx_)); }

  // Geometric::downsample (geometric.cpp:55-126): indices of the kept points in the reference's output order.
  std::vector<uint32_t> downsample(const Point* pts, size_t n, float leaf, size_t cap, float min_dist) const {
    std::vector<uint32_t> out(n ? n : 1);
    size_t kept = 0;
    check(mb_downsample(ctx_, &pts->x, n, sizeof(Point), leaf, cap, min_dist, out.data(), &kept));
    out.resize(kept);
    return out;
  }

 private:
  mb_ctx* ctx_ = nullptr;
};

class IncrementalVoxelMapB200 {
 public:
  using Ptr = std::shared_ptr<IncrementalVoxelMapB200>;

  // IncrementalVoxelMapPCL(leaf) followed by the three setters at geometric.cpp:25-28.
  IncrementalVoxelMapB200(const Context& ctx, float leaf_size, float min_dist_in_cell = 0.1f, int neighbor_voxel_mode = 7,
                          size_t lru_horizon = 100, int max_points_per_cell = 20)
      : ctx_(&ctx) {
    check(mb_map_create(ctx.get(), leaf_size, min_dist_in_cell, max_points_per_cell, neighbor_voxel_mode, lru_horizon, &map_));
  }
  // Deep copy, as the reference's copy constructor (incremental_voxel_map.hpp:34-43; used at geometric.cpp:494).
  IncrementalVoxelMapB200(const IncrementalVoxelMapB200& other) : ctx_(other.ctx_) { check(mb_map_snapshot(other.map_, &map_)); }
  IncrementalVoxelMapB200& operator=(const IncrementalVoxelMapB200&) = delete;
  ~IncrementalVoxelMapB200() { mb_map_release(map_); }

  void insert(const Point* pts, size_t n) { check(mb_map_insert(map_, &pts->x, n, sizeof(Point))); }
  void insert_xyz(const float* xyz, size_t n) { check(mb_map_insert(map_, xyz, n, 3 * sizeof(float))); }

  // Same contract as IncrementalVoxelMapPCL::knn_search: true iff exactly k neighbours were found.
  bool knn_search(const double point[3], size_t k, std::vector<size_t>& indices, std::vector<double>& sq_dists) const {
    indices.resize(k);
    sq_dists.resize(k);
    std::vector<uint64_t> idx(k);
    uint8_t ok = 0;
    check(mb_map_knn(map_, point, 1, (int)k, idx.data(), sq_dists.data(), &ok));
    for (size_t i = 0; i < k; ++i) indices[i] = (size_t)idx[i];
    return ok != 0;
  }
  // underlying()->point(i)
  std::array<double, 3> point(size_t index) const {
    const uint64_t i = index;
    std::array<double, 3> p{};
    check(mb_map_points(map_, &i, 1, p.data()));
    return p;
  }
  // getCloud(): every stored point, voxel order then in-voxel order.
  std::vector<std::array<float, 3>> getCloud() const {
    size_t nv = 0, np = 0;
    check(mb_map_size(map_, &nv, &np, nullptr));
    std::vector<int32_t> counts(nv);
    std::vector<float> pts(nv * 20 * 3);
    check(mb_map_download(map_, nullptr, counts.data(), nullptr, pts.data()));
    std::vector<std::array<float, 3>> out;
    out.reserve(np);
    for (size_t v = 0; v < nv; ++v)
      for (int j = 0; j < counts[v]; ++j) out.push_back({pts[(v * 20 + j) * 3], pts[(v * 20 + j) * 3 + 1], pts[(v * 20 + j) * 3 + 2]});
    return out;
  }
  size_t size() const {
    size_t np = 0;
    check(mb_map_size(map_, nullptr, &np, nullptr));
    return np;
  }
  mb_map* get() const { return map_; }
  const Context& context() const { return *ctx_; }

 private:
  const Context* ctx_;
  mb_map* map_ = nullptr;
};

class ICPFactorB200 {
 public:
  using Ptr = std::shared_ptr<ICPFactorB200>;
  enum class RejectStatus : uint8_t {  // geometric_factor.hpp:35-46
    Unprocessed = 0,
    InsufficientCorresPoints,
    CorresMaxDist,
    EigenSolverFail,
    MinEigenValueLow,
    Line,
    CorresPlaneInvalid,
    MaxError,
    Valid
  };

  // ICPFactor(key_source, ivox_target, cloud_source, config): the scan is copied, the map is shared.
  ICPFactorB200(IncrementalVoxelMapB200::Ptr ivox_target, const Point* cloud_source, size_t n, const RegistrationConfig& config)
      : target_(std::move(ivox_target)), n_(n) {
    const mb_icp_config c = config.to_c();
    check(mb_factor_create(target_->context().get(), target_->get(), cloud_source, n, sizeof(Point), &c, 0, n, &f_));
  }
  ICPFactorB200(const ICPFactorB200&) = delete;
  ICPFactorB200& operator=(const ICPFactorB200&) = delete;
  ~ICPFactorB200() { mb_factor_release(f_); }

  size_t dim() const { return 6; }

  // linearize(values): R (row-major) and t of values.at<Pose3>(key), the unit gravity direction of
  // values.at<Unit3>(G(0)).  Returns the HessianFactor terms G = H, g, f (geometric_factor.hpp:559-560).
  const mb_linearization& linearize(const double R[9], const double t[3], const double gravity_unit[3]) {
    check(mb_factor_linearize(f_, R, t, gravity_unit, &last_));
    return last_;
  }

  std::vector<RejectStatus> getStatuses() const {
    std::vector<RejectStatus> s(n_);
    check(mb_factor_download_state(f_, reinterpret_cast<uint8_t*>(s.data()), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    return s;
  }
  std::vector<std::array<double, 3>> getCorresMeansTarget() const { return vec(1); }
  std::vector<std::array<double, 3>> getCorresNormalsTarget() const { return vec(2); }
  void getLocalizabilities(double trans_comp[3], double rot_comp[3], double trans_final[3], double rot_final[3],
                           double eigenvectors_trans[9], double eigenvectors_rot[9]) const {
    for (int i = 0; i < 3; ++i) {
      trans_comp[i] = last_.loc_trans_comp[i];
      rot_comp[i] = last_.loc_rot_comp[i];
      trans_final[i] = last_.loc_trans_final[i];
      rot_final[i] = last_.loc_rot_final[i];
    }
    for (int i = 0; i < 9; ++i) {
      eigenvectors_trans[i] = last_.eigvec_trans[i];
      eigenvectors_rot[i] = last_.eigvec_rot[i];
    }
  }
  void getDegenInfo(double rot[3], double eigenvectors_rot[9], double trans[3], double eigenvectors_trans[9]) const {
    for (int i = 0; i < 3; ++i) {
      rot[i] = last_.degen_rot[i];
      trans[i] = last_.degen_trans[i];
    }
    for (int i = 0; i < 9; ++i) {
      eigenvectors_rot[i] = last_.degen_eigvec_rot[i];
      eigenvectors_trans[i] = last_.degen_eigvec_trans[i];
    }
  }
  int getLinearizeCount() const { return last_.linearize_count; }
  mb_factor* get() const { return f_; }

 private:
  std::vector<std::array<double, 3>> vec(int which) const {
    std::vector<std::array<double, 3>> v(n_);
    double* ptrs[3] = {nullptr, nullptr, nullptr};
    ptrs[which] = v.empty() ? nullptr : v[0].data();
    check(mb_factor_download_state(f_, nullptr, ptrs[0], ptrs[1], ptrs[2], nullptr, nullptr, nullptr));
    return v;
  }
  IncrementalVoxelMapB200::Ptr target_;
  size_t n_;
  mb_factor* f_ = nullptr;
  mb_linearization last_{};
};

}  // namespace mimosa_b200
