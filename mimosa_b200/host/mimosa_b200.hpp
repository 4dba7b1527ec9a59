// mimosa_b200.hpp — header-only C++17 mirror of mimosa's LiDAR geometric-factor interface over the C ABI
// (include/mimosa_b200.h).  Same class roles, method names and error behaviour (exceptions) as the reference:
//
//   mimosa_b200::RegistrationConfig       mimosa/include/mimosa/lidar/geometric_config.hpp:17-33
//   mimosa_b200::IncrementalVoxelMapB200  mimosa::lidar::IncrementalVoxelMapPCL, incremental_voxel_map.hpp:22-51
//   mimosa_b200::ICPFactorB200            mimosa::lidar::ICPFactor (unary), geometric_factor.hpp:25-563
//   mimosa_b200::GeometricB200            mimosa::lidar::Geometric's preprocess / getFactors / updateMap, geometric.hpp:80-88
//   mimosa_b200::KeyframeGate             the keyframe rule inside Geometric::updateMap, geometric.cpp:437-478
//   mimosa_b200::degeneracyInfo           the degeneracy flags of Geometric::getFactors, geometric.cpp:208-228
//
// It depends on nothing but the C ABI and the standard library; the GTSAM / PCL glue a mimosa build adds on top
// (deriving from gtsam::NonlinearFactor, taking pcl::PointCloud<Point>) is shown in INTEGRATION.md.
// All compute runs in libmimosa_b200.so on the GPU.  Exceptions replace the reference's logCriticalException
// (mimosa/include/mimosa/utils.hpp:300-306).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
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
  // How long the linearisation kernel stays resident after ICPFactorB200::linearize (0: every call launches).
  void set_resident_window(unsigned microseconds) const { check(mb_set_resident_window(ctx_, microseconds)); }

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
      : ctx_(&ctx), cap_(max_points_per_cell) {
    check(mb_map_create(ctx.get(), leaf_size, min_dist_in_cell, max_points_per_cell, neighbor_voxel_mode, lru_horizon, &map_));
  }
  // Deep copy, as the reference's copy constructor (incremental_voxel_map.hpp:34-43; used at geometric.cpp:494).
  IncrementalVoxelMapB200(const IncrementalVoxelMapB200& other) : ctx_(other.ctx_), cap_(other.cap_) {
    check(mb_map_snapshot(other.map_, &map_));
  }
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
    const size_t cap = (size_t)cap_;  // mb_map_download writes nv * cap * 3 floats
    std::vector<int32_t> counts(nv);
    std::vector<float> pts(nv * cap * 3);
    check(mb_map_download(map_, nullptr, counts.data(), nullptr, pts.data()));
    std::vector<std::array<float, 3>> out;
    out.reserve(np);
    for (size_t v = 0; v < nv; ++v)
      for (size_t j = 0; j < (size_t)counts[v]; ++j)
        out.push_back({pts[(v * cap + j) * 3], pts[(v * cap + j) * 3 + 1], pts[(v * cap + j) * 3 + 2]});
    return out;
  }
  size_t size() const {
    size_t np = 0;
    check(mb_map_size(map_, nullptr, &np, nullptr));
    return np;
  }
  // updateMap's insertion (geometric.cpp:483-495): W = R_W_Be p + t_W_Be in float for a device scan, then insert
  void insert_scan(mb_scan* scan, const float R[9], const float t[3]) { check(mb_map_insert_scan(map_, scan, R, t)); }
  int max_points_per_cell() const { return cap_; }
  mb_map* get() const { return map_; }
  const Context& context() const { return *ctx_; }

 private:
  const Context* ctx_;
  int cap_ = 20;
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
  // the same from a device-resident scan (Geometric::getFactors builds the factor from sm_Be_cloud_ds_, geometric.cpp:194)
  ICPFactorB200(IncrementalVoxelMapB200::Ptr ivox_target, mb_scan* cloud_source, const RegistrationConfig& config)
      : target_(std::move(ivox_target)) {
    size_t stride = 0;
    check(mb_scan_size(cloud_source, &n_, &stride));
    const mb_icp_config c = config.to_c();
    check(mb_factor_create_from_scan(target_->context().get(), target_->get(), cloud_source, &c, 0, n_, &f_));
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
  // the nine-bin status histogram Geometric::getFactors builds from getStatuses() (geometric.cpp:280-323), as counted
  // on the device by the last linearize()
  std::array<int64_t, 9> getStatusHistogram() const {
    std::array<int64_t, 9> h{};
    for (int i = 0; i < 9; ++i) h[i] = last_.counts[i];
    return h;
  }
  const mb_linearization& last() const { return last_; }
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

// ---- lidar::Geometric: the caller either side of the factor ------------------------------------------------------

// What Geometric::getFactors derives from the first linearisation (geometric.cpp:208-228): blkdiag(V_rot, V_trans)
// and the six booleans "component localizability below its threshold" (rotations first).
struct DegeneracyInfo {
  std::array<double, 36> eigenvectors_block_matrix{};  // row-major 6x6
  std::array<double, 6> degen_directions{};            // 1.0 / 0.0, as the reference's V6D
};
inline DegeneracyInfo degeneracyInfo(const mb_linearization& lin, const RegistrationConfig& cfg) {
  DegeneracyInfo d;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      d.eigenvectors_block_matrix[6 * r + c] = lin.eigvec_rot[3 * r + c];
      d.eigenvectors_block_matrix[6 * (r + 3) + 3 + c] = lin.eigvec_trans[3 * r + c];
    }
  for (int a = 0; a < 3; ++a) {
    d.degen_directions[a] = lin.loc_rot_comp[a] < (double)cfg.degen_thresh_rot ? 1.0 : 0.0;
    d.degen_directions[3 + a] = lin.loc_trans_comp[a] < (double)cfg.degen_thresh_trans ? 1.0 : 0.0;
  }
  return d;
}

// Pose as the reference's gtsam::Pose3: R row-major, t.
struct Pose {
  std::array<double, 9> R{1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::array<double, 3> t{0, 0, 0};
};

// The keyframe rule of Geometric::updateMap (geometric.cpp:437-478): compare with the NEAREST (by translation, float,
// first minimum) pose already in the map; update when that distance exceeds map_keyframe_trans_thresh, or when the
// largest |yaw|, |pitch|, |roll| of  R_B_L^-1 * (R_kf^-1 R_now) * R_B_L  exceeds DEG2RAD(map_keyframe_rot_thresh_deg)
// (PCL's macro: x * 0.017453293); the first initial_clouds_to_force_map_update calls always update.  The Euler
// angles are GTSAM's Rot3::ypr(), i.e. its RQ decomposition (gtsam/geometry/Rot3.cpp, external to the reference).
class KeyframeGate {
 public:
  KeyframeGate(float trans_thresh, float rot_thresh_deg, size_t initial_clouds_to_force_map_update,
               const std::array<double, 9>& R_B_L = {1, 0, 0, 0, 1, 0, 0, 0, 1})
      : trans_thresh_(trans_thresh), rot_thresh_deg_(rot_thresh_deg), forced_left_(initial_clouds_to_force_map_update), R_B_L_(R_B_L) {}

  // geometric.cpp:439-478; like the reference, a forced update consumes one of the initial clouds whatever the rule said
  bool shouldUpdate(const Pose& T_W_Be) {
    bool update_map = true;
    if (!map_poses_.empty()) {
      float min_diff_trans = std::numeric_limits<float>::max();
      size_t min_diff_index = 0;
      for (size_t i = 0; i < map_poses_.size(); ++i) {
        const double dx = map_poses_[i].t[0] - T_W_Be.t[0], dy = map_poses_[i].t[1] - T_W_Be.t[1], dz = map_poses_[i].t[2] - T_W_Be.t[2];
        const float diff_trans = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
        if (diff_trans < min_diff_trans) {
          min_diff_trans = diff_trans;
          min_diff_index = i;
        }
      }
      const std::array<double, 9> between = mulT(map_poses_[min_diff_index].R, T_W_Be.R);  // R_kf^T R_now
      const std::array<double, 9> rot_diff = mul(mulT(R_B_L_, between), R_B_L_);
      const std::array<double, 3> xyz = rq_xyz(rot_diff);
      const double max_abs = std::max(std::fabs(xyz[0]), std::max(std::fabs(xyz[1]), std::fabs(xyz[2])));
      if (min_diff_trans > trans_thresh_) {
        update_map = true;
      } else if (max_abs > (double)rot_thresh_deg_ * 0.017453293) {
        update_map = true;
      } else {
        update_map = false;
      }
    }
    if (forced_left_ > 0) {
      update_map = true;
      --forced_left_;
    }
    return update_map;
  }
  void addKeyframe(const Pose& T_W_Be) { map_poses_.push_back(T_W_Be); }  // map_poses_.push_back, geometric.cpp:499
  const std::vector<Pose>& keyframes() const { return map_poses_; }

  // GTSAM's RQ(A): the angles (x, y, z) with A = Rz(z) Ry(y) Rx(x); Rot3::ypr() returns them as (z, y, x).
  static std::array<double, 3> rq_xyz(const std::array<double, 9>& A) {
    const double x = -std::atan2(-A[7], A[8]);
    const std::array<double, 9> B = mul(A, rx(-x));
    const double y = -std::atan2(B[6], B[8]);
    const std::array<double, 9> C = mul(B, ry(-y));
    const double z = -std::atan2(-C[3], C[4]);
    return {x, y, z};
  }

 private:
  static std::array<double, 9> mul(const std::array<double, 9>& a, const std::array<double, 9>& b) {
    std::array<double, 9> c{};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    return c;
  }
  static std::array<double, 9> mulT(const std::array<double, 9>& a, const std::array<double, 9>& b) {  // a^T b
    std::array<double, 9> c{};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) c[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
    return c;
  }
  static std::array<double, 9> rx(double t) { return {1, 0, 0, 0, std::cos(t), -std::sin(t), 0, std::sin(t), std::cos(t)}; }
  static std::array<double, 9> ry(double t) { return {std::cos(t), 0, std::sin(t), 0, 1, 0, -std::sin(t), 0, std::cos(t)}; }
  float trans_thresh_, rot_thresh_deg_;
  size_t forced_left_;
  std::array<double, 9> R_B_L_;
  std::vector<Pose> map_poses_;
};

// The fields of lidar::GeometricConfig this path reads (geometric_config.hpp:37-57), defaults included.
struct GeometricConfig {
  Pose T_B_L;
  float map_keyframe_trans_thresh = 0.1f;
  float map_keyframe_rot_thresh_deg = 10.f;
  size_t initial_clouds_to_force_map_update = 10;
  size_t lru_horizon = 100;
  size_t neighbor_voxel_mode = 7;
  RegistrationConfig scan_to_map;
};

// lidar::Geometric's three per-scan steps (geometric.hpp:80-88) on device-resident scans:
//   preprocess  (geometric.cpp:128-183)  points_deskewed[idxs] -> T_B_L in float -> Be_cloud_ -> downsample (cap 20)
//   getFactors  (geometric.cpp:185-328)  ICPFactor on the current map, first linearisation, degeneracy info, histogram
//   updateMap   (geometric.cpp:427-513)  keyframe rule, deep map copy, insertion of the FULL Be_cloud_ in float
// The caller keeps the graph side (graph.add(factor), the smoother's later linearize() calls on the returned factor).
class GeometricB200 {
 public:
  GeometricB200(const Context& ctx, const GeometricConfig& config)
      : ctx_(&ctx), config(config),
        gate_(config.map_keyframe_trans_thresh, config.map_keyframe_rot_thresh_deg, config.initial_clouds_to_force_map_update, config.T_B_L.R) {
    // geometric.cpp:23-28
    ivox_map_ = std::make_shared<IncrementalVoxelMapB200>(ctx, config.scan_to_map.target_ivox_map_leaf_size,
                                                          config.scan_to_map.target_ivox_map_min_dist_in_voxel,
                                                          (int)config.neighbor_voxel_mode, config.lru_horizon, 20);
  }
  ~GeometricB200() {
    mb_scan_release(Be_cloud_);
    mb_scan_release(sm_Be_cloud_ds_);
  }
  GeometricB200(const GeometricB200&) = delete;
  GeometricB200& operator=(const GeometricB200&) = delete;

  void preprocess(mb_scan* points_deskewed, const uint32_t* idxs, size_t n_idxs) {
    mb_scan_release(Be_cloud_);
    mb_scan_release(sm_Be_cloud_ds_);
    Be_cloud_ = sm_Be_cloud_ds_ = nullptr;
    check(mb_scan_gather(points_deskewed, idxs, n_idxs, &Be_cloud_));
    float R[9], t[3];
    for (int a = 0; a < 9; ++a) R[a] = (float)config.T_B_L.R[a];
    for (int a = 0; a < 3; ++a) t[a] = (float)config.T_B_L.t[a];
    check(mb_scan_transform(Be_cloud_, R, t));
    check(mb_scan_downsample(Be_cloud_, config.scan_to_map.source_voxel_grid_filter_leaf_size, 20,
                             config.scan_to_map.source_voxel_grid_min_dist_in_voxel, &sm_Be_cloud_ds_));
  }

  ICPFactorB200::Ptr getFactors(const Pose& T_W_Be, const double gravity_unit[3], DegeneracyInfo& degen) {
    factor_ = std::make_shared<ICPFactorB200>(ivox_map_, sm_Be_cloud_ds_, config.scan_to_map);
    const mb_linearization& lin = factor_->linearize(T_W_Be.R.data(), T_W_Be.t.data(), gravity_unit);
    degen = degeneracyInfo(lin, config.scan_to_map);
    return factor_;
  }

  // returns whether the map was updated
  bool updateMap(const Pose& T_W_Be) {
    if (!gate_.shouldUpdate(T_W_Be)) return false;
    float R[9], t[3];
    for (int a = 0; a < 9; ++a) R[a] = (float)T_W_Be.R[a];
    for (int a = 0; a < 3; ++a) t[a] = (float)T_W_Be.t[a];
    ivox_map_ = std::make_shared<IncrementalVoxelMapB200>(*ivox_map_);  // factors in the smoother keep their snapshot
    ivox_map_->insert_scan(Be_cloud_, R, t);
    gate_.addKeyframe(T_W_Be);
    return true;
  }

  const IncrementalVoxelMapB200::Ptr& map() const { return ivox_map_; }
  const KeyframeGate& gate() const { return gate_; }
  mb_scan* Be_cloud() const { return Be_cloud_; }
  mb_scan* sm_Be_cloud_ds() const { return sm_Be_cloud_ds_; }

 private:
  const Context* ctx_;

 public:
  const GeometricConfig config;

 private:
  KeyframeGate gate_;
  IncrementalVoxelMapB200::Ptr ivox_map_;
  ICPFactorB200::Ptr factor_;
  mb_scan* Be_cloud_ = nullptr;
  mb_scan* sm_Be_cloud_ds_ = nullptr;
};

}  // namespace mimosa_b200
