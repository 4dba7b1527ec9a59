// geometric_factor_b200.hpp — the GTSAM / PCL side of the drop-in: what a mimosa maintainer adds next to
// mimosa/include/mimosa/lidar/geometric_factor.hpp so that lidar::Geometric can build the B200 factor instead of
// ICPFactor without touching graph::Manager or ISAM2.
//
//   mimosa::lidar::ICPFactorB200 : gtsam::NonlinearFactor
//     constructor        same signature as ICPFactor's unary one, geometric_factor.hpp:119-129 (the cloud is copied, :125)
//     clone / dim / error   :160-174 (error() returns 0 like the reference; a clone shares the per-point caches the way the
//                        reference's copy shares ivox_target_)
//     linearize          :231-562 -> gtsam::HessianFactor(key, G = J^T J, g = -J^T e, f = sum e^2), :559-560; reads the pose
//                        of keys()[0] (:247) and, unconditionally like the reference, the gravity direction G(0) (:257)
//     accessors          getStatuses / getCorresMeansTarget / getCorresNormalsTarget / getLocalizabilities / getDegenInfo /
//                        getLinearizeCount, :48-72, consumed by Geometric::getFactors (geometric.cpp:208-228, 280-323)
//
// Needs only what the reference's factor already includes: gtsam (NonlinearFactor, HessianFactor, Pose3, Unit3, Values,
// Symbol) and pcl::PointCloud.  It is written against the element-access subset of those APIs — matrix(i, j), vector(i),
// Matrix(rows, cols), Vector(n) — so that this repository can compile and run it against the minimal stand-ins under
// tests/cpp/stubs/ (no GTSAM / PCL / Eigen exists in the build image); with the real libraries the same source compiles
// unchanged.  MIMOSA_B200_POINT / MIMOSA_B200_CONFIG name the reference's lidar::Point and RegistrationConfig types.
#pragma once
#include <gtsam/geometry/Pose3.h>
#include <gtsam/geometry/Unit3.h>
#include <gtsam/inference/Symbol.h>
#include <gtsam/linear/HessianFactor.h>
#include <gtsam/nonlinear/NonlinearFactor.h>
#include <gtsam/nonlinear/Values.h>
#include <pcl/point_cloud.h>

#include <memory>
#include <vector>

#include "../mimosa_b200.hpp"

#ifndef MIMOSA_B200_POINT
#define MIMOSA_B200_POINT ::mimosa::lidar::Point
#endif
#ifndef MIMOSA_B200_CONFIG
#define MIMOSA_B200_CONFIG ::mimosa::lidar::RegistrationConfig
#endif

namespace mimosa {
namespace lidar {

class ICPFactorB200 : public gtsam::NonlinearFactor {
 public:
  using Base = gtsam::NonlinearFactor;
  using This = ICPFactorB200;
  using Ptr = std::shared_ptr<ICPFactorB200>;
  using RejectStatus = mimosa_b200::ICPFactorB200::RejectStatus;  // same enumerators, same values (:35-46)
  using MapPtr = mimosa_b200::IncrementalVoxelMapB200::Ptr;

  ICPFactorB200(const gtsam::Key key_source, MapPtr ivox_target, const pcl::PointCloud<MIMOSA_B200_POINT>& cloud_source,
                const MIMOSA_B200_CONFIG& config)
      : Base(std::vector<gtsam::Key>{key_source}),
        impl_(std::make_shared<mimosa_b200::ICPFactorB200>(
            std::move(ivox_target), reinterpret_cast<const mimosa_b200::Point*>(cloud_source.points.data()), cloud_source.size(),
            to_b200(config))) {
    static_assert(sizeof(MIMOSA_B200_POINT) == sizeof(mimosa_b200::Point), "lidar::Point is the 32-byte record of point.hpp:18-39");
  }
  ~ICPFactorB200() override {}

  gtsam::NonlinearFactor::shared_ptr clone() const override {
    return std::static_pointer_cast<gtsam::NonlinearFactor>(gtsam::NonlinearFactor::shared_ptr(new This(*this)));
  }
  size_t dim() const override { return 6; }
  double error(const gtsam::Values&) const override { return 0.0; }

  std::shared_ptr<gtsam::GaussianFactor> linearize(const gtsam::Values& c) const override {
    const gtsam::Pose3& T_W_B = c.at<gtsam::Pose3>(keys()[0]);
    const gtsam::Unit3& gravity = c.at<gtsam::Unit3>(gtsam::symbol_shorthand::G(0));  // throws when absent, like the reference
    const auto Rm = T_W_B.rotation().matrix();
    const auto tv = T_W_B.translation();
    const auto gu = gravity.unitVector();
    double R[9], t[3], g[3];
    for (int r = 0; r < 3; ++r) {
      for (int col = 0; col < 3; ++col) R[3 * r + col] = Rm(r, col);
      t[r] = tv(r);
      g[r] = gu(r);
    }
    const mb_linearization& L = impl_->linearize(R, t, g);
    gtsam::Matrix G(6, 6);
    gtsam::Vector gv(6);
    for (int r = 0; r < 6; ++r) {
      for (int col = 0; col < 6; ++col) G(r, col) = L.H[6 * r + col];
      gv(r) = L.g[r];  // already -J^T e
    }
    return std::make_shared<gtsam::HessianFactor>(keys()[0], G, gv, L.f);
  }

  std::vector<RejectStatus> getStatuses() const { return impl_->getStatuses(); }
  std::vector<std::array<double, 3>> getCorresMeansTarget() const { return impl_->getCorresMeansTarget(); }
  std::vector<std::array<double, 3>> getCorresNormalsTarget() const { return impl_->getCorresNormalsTarget(); }
  // V3D / M33 outputs: anything with (i) / (i, j) element access
  template <typename V3, typename M3>
  void getLocalizabilities(V3& trans_comp, V3& rot_comp, V3& trans_final, V3& rot_final, M3& eigenvectors_trans,
                           M3& eigenvectors_rot) const {
    double tc[3], rc[3], tf[3], rf[3], et[9], er[9];
    impl_->getLocalizabilities(tc, rc, tf, rf, et, er);
    for (int a = 0; a < 3; ++a) {
      trans_comp(a) = tc[a], rot_comp(a) = rc[a], trans_final(a) = tf[a], rot_final(a) = rf[a];
      for (int b = 0; b < 3; ++b) eigenvectors_trans(a, b) = et[3 * a + b], eigenvectors_rot(a, b) = er[3 * a + b];
    }
  }
  template <typename V3, typename M3>
  void getDegenInfo(V3& degen_rot, M3& degen_eigenvectors_rot, V3& degen_trans, M3& degen_eigenvectors_trans) const {
    double r[3], t[3], er[9], et[9];
    impl_->getDegenInfo(r, er, t, et);
    for (int a = 0; a < 3; ++a) {
      degen_rot(a) = r[a], degen_trans(a) = t[a];
      for (int b = 0; b < 3; ++b) degen_eigenvectors_rot(a, b) = er[3 * a + b], degen_eigenvectors_trans(a, b) = et[3 * a + b];
    }
  }
  int getLinearizeCount() const { return impl_->getLinearizeCount(); }
  const mimosa_b200::ICPFactorB200& impl() const { return *impl_; }

 private:
  static mimosa_b200::RegistrationConfig to_b200(const MIMOSA_B200_CONFIG& c) {  // field by field, geometric_config.hpp:17-33
    mimosa_b200::RegistrationConfig o;
    o.source_voxel_grid_filter_leaf_size = c.source_voxel_grid_filter_leaf_size;
    o.source_voxel_grid_min_dist_in_voxel = c.source_voxel_grid_min_dist_in_voxel;
    o.target_ivox_map_leaf_size = c.target_ivox_map_leaf_size;
    o.target_ivox_map_min_dist_in_voxel = c.target_ivox_map_min_dist_in_voxel;
    o.num_corres_points = c.num_corres_points;
    o.max_corres_distance = c.max_corres_distance;
    o.plane_validity_distance = c.plane_validity_distance;
    o.lidar_point_noise_std_dev = c.lidar_point_noise_std_dev;
    o.use_huber = c.use_huber;
    o.huber_threshold = c.huber_threshold;
    o.reg_4_dof = c.reg_4_dof;
    o.project_on_degneneracy = c.project_on_degneneracy;
    o.degen_thresh_rot = c.degen_thresh_rot;
    o.degen_thresh_trans = c.degen_thresh_trans;
    return o;
  }
  // shared: a GTSAM clone (ISAM2 clones factors it keeps) works on the same device-side caches, and `linearize` is
  // const-but-mutating exactly like the reference's (`mutable` members, geometric_factor.hpp:79-116)
  std::shared_ptr<mimosa_b200::ICPFactorB200> impl_;
};

}  // namespace lidar
}  // namespace mimosa
