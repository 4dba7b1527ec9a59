// The GTSAM / PCL adapter (mimosa_b200/host/adapters/geometric_factor_b200.hpp) compiled against the minimal stand-ins
// under tests/cpp/stubs/ and driven the way ISAM2 drives a factor: through gtsam::NonlinearFactor pointers and clones.
// Exit codes: 0 ok, 2 no device (expected on the CPU box), 1 failure.
#include <cmath>
#include <cstdio>
#include <random>

// the reference's own types, as far as the adapter needs them (point.hpp:18-39, geometric_config.hpp:17-33)
namespace mimosa {
namespace lidar {
struct alignas(16) Point {
  float x, y, z, pad;
  float intensity;
  unsigned t, idx;
  float range;
};
struct RegistrationConfig {
  float source_voxel_grid_filter_leaf_size = 0.5f, source_voxel_grid_min_dist_in_voxel = 0.1f;
  float target_ivox_map_leaf_size = 0.5f, target_ivox_map_min_dist_in_voxel = 0.1f;
  size_t num_corres_points = 5;
  float max_corres_distance = 2.24f, plane_validity_distance = 0.04f, lidar_point_noise_std_dev = 0.02f;
  bool use_huber = true;
  float huber_threshold = 1.345f;
  bool reg_4_dof = false, project_on_degneneracy = true;
  float degen_thresh_rot = 10.f, degen_thresh_trans = 15.f;
};
}  // namespace lidar
}  // namespace mimosa

#include "adapters/geometric_factor_b200.hpp"

using namespace mimosa::lidar;
using gtsam::symbol_shorthand::G;
using gtsam::symbol_shorthand::X;

int main() {
  try {
    mimosa_b200::Context ctx(0);
    RegistrationConfig cfg;  // hornbill values
    cfg.source_voxel_grid_filter_leaf_size = cfg.target_ivox_map_leaf_size = 1.f;
    cfg.source_voxel_grid_min_dist_in_voxel = cfg.target_ivox_map_min_dist_in_voxel = 0.2f;
    cfg.max_corres_distance = 1.f;
    cfg.plane_validity_distance = cfg.lidar_point_noise_std_dev = 0.07f;
    cfg.project_on_degneneracy = false;
    auto map = std::make_shared<mimosa_b200::IncrementalVoxelMapB200>(ctx, 1.f, 0.2f, 19, 1000);
    std::mt19937 rng(2);
    std::uniform_real_distribution<float> u(-10.f, 10.f);
    std::normal_distribution<float> nz(0.f, 0.01f);
    std::vector<mimosa_b200::Point> world(40000);
    for (auto& p : world) p = mimosa_b200::Point{u(rng), u(rng), -1.37f + nz(rng), 1.f, 0.f, 0u, 0u, 0.f};
    map->insert(world.data(), world.size());
    pcl::PointCloud<Point> scan;
    for (int i = 0; i < 3000; ++i) scan.push_back(Point{0.8f * u(rng), 0.8f * u(rng), -1.37f + nz(rng), 1.f, 0.f, 0u, 0u, 0.f});

    // the factor as Geometric::getFactors builds it (geometric.cpp:194), then held like the graph holds it
    gtsam::NonlinearFactor::shared_ptr factor = std::make_shared<ICPFactorB200>(X(7), map, scan, cfg);
    if (factor->dim() != 6 || factor->keys().size() != 1 || factor->keys()[0] != X(7)) return 1;
    gtsam::Values values;
    gtsam::Vector3 t;
    t(2) = 0.02;
    values.insert(X(7), gtsam::Pose3(gtsam::Rot3(), t));
    bool threw = false;
    try {
      factor->linearize(values);  // G(0) missing: must throw like the reference's c.at<Unit3>(G(0)), :257
    } catch (const std::out_of_range&) {
      threw = true;
    }
    if (!threw) return 1;
    values.insert(G(0), gtsam::Unit3(0, 0, -1));
    if (factor->error(values) != 0.0) return 1;
    auto gf = std::dynamic_pointer_cast<gtsam::HessianFactor>(factor->linearize(values));
    if (!gf || gf->key() != X(7) || gf->information().rows() != 6) return 1;
    const auto& icp = static_cast<const ICPFactorB200&>(*factor);
    const mb_linearization& L = icp.impl().last();
    for (int r = 0; r < 6; ++r) {
      if (gf->linearTerm()(r) != L.g[r]) return 1;
      for (int c = 0; c < 6; ++c)
        if (gf->information()(r, c) != L.H[6 * r + c] || gf->information()(r, c) != gf->information()(c, r)) return 1;
    }
    if (gf->constantTerm() != L.f || !(L.f > 0) || !(gf->information()(5, 5) > 0) || !(gf->linearTerm()(5) < 0)) return 1;
    // a clone (ISAM2 keeps clones) shares the caches: the second linearisation at the same pose searches nothing
    gtsam::NonlinearFactor::shared_ptr copy = factor->clone();
    auto gf2 = std::dynamic_pointer_cast<gtsam::HessianFactor>(copy->linearize(values));
    const auto& icp2 = static_cast<const ICPFactorB200&>(*copy);
    if (icp2.getLinearizeCount() != 2 || icp2.impl().last().n_searched != 0) return 1;
    if (gf2->constantTerm() != gf->constantTerm()) return 1;
    gtsam::Vector3 tc, rc, tf, rf;
    gtsam::Matrix3 et, er;
    icp2.getLocalizabilities(tc, rc, tf, rf, et, er);
    long valid = 0;
    for (auto s : icp2.getStatuses()) valid += s == ICPFactorB200::RejectStatus::Valid;
    std::printf("adapter ok: valid %ld of %zu, f = %.3f, loc_trans_comp z = %.1f\n", valid, scan.size(), gf->constantTerm(), tc(2));
    if (valid < 1000 || !(tc(2) > 100.0) || !(tf(2) > 0)) return 1;
    return 0;
  } catch (const mimosa_b200::Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return e.code == MB_ERR_NO_DEVICE ? 2 : 1;
  }
}
