// Compiles against the C++ host mirror exactly as a mimosa translation unit would, and (on a GPU box) runs a
// tiny scan-to-map linearisation through it.  Exit codes: 0 ok, 2 no device (expected on the CPU box), 1 failure.
#include <cmath>
#include <cstdio>
#include <random>

#include "mimosa_b200.hpp"

using namespace mimosa_b200;

int main() {
  try {
    Context ctx(0);
    RegistrationConfig cfg;  // hornbill values, mimosa/config/hornbill/params.yaml:86-102
    cfg.source_voxel_grid_filter_leaf_size = cfg.target_ivox_map_leaf_size = 1.f;
    cfg.source_voxel_grid_min_dist_in_voxel = cfg.target_ivox_map_min_dist_in_voxel = 0.2f;
    cfg.max_corres_distance = 1.f;
    cfg.plane_validity_distance = 0.07f;
    cfg.lidar_point_noise_std_dev = 0.07f;
    cfg.project_on_degneneracy = false;
    auto map = std::make_shared<IncrementalVoxelMapB200>(ctx, cfg.target_ivox_map_leaf_size, cfg.target_ivox_map_min_dist_in_voxel, 19, 1000);
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> u(-10.f, 10.f);
    std::normal_distribution<float> nz(0.f, 0.01f);
    std::vector<Point> cloud(40000);
    for (auto& p : cloud) p = Point{u(rng), u(rng), -1.37f + nz(rng), 1.f, 0.f, 0u, 0u, 0.f};
    map->insert(cloud.data(), cloud.size());
    auto snapshot = std::make_shared<IncrementalVoxelMapB200>(*map);  // geometric.cpp:494
    snapshot->insert(cloud.data(), 1000);                              // re-inserting changes nothing
    if (snapshot->size() != map->size()) return 1;
    std::vector<Point> scan(2000);
    for (auto& p : scan) p = Point{0.8f * u(rng), 0.8f * u(rng), -1.37f + nz(rng), 1.f, 0.f, 0u, 0u, 0.f};
    ICPFactorB200 factor(map, scan.data(), scan.size(), cfg);
    const double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0.02}, g[3] = {0, 0, -1};
    const mb_linearization& L = factor.linearize(R, t, g);
    const auto st = factor.getStatuses();
    long valid = 0;
    for (auto s : st) valid += s == ICPFactorB200::RejectStatus::Valid;
    std::printf("valid %ld of %zu, H[5][5] = %.3f, g[5] = %.3f, f = %.3f, count %d\n", valid, st.size(), L.H[35], L.g[5], L.f,
                factor.getLinearizeCount());
    if (valid != L.counts[8] || valid < 500 || !(L.H[35] > 0) || !(L.g[5] < 0)) return 1;
    std::vector<size_t> idx;
    std::vector<double> d2;
    const double q[3] = {0.1, 0.2, -1.37};
    if (!map->knn_search(q, 5, idx, d2) || !(d2[0] <= d2[4])) return 1;
    const auto p0 = map->point(idx[0]);
    if (std::fabs((p0[0] - q[0]) * (p0[0] - q[0]) + (p0[1] - q[1]) * (p0[1] - q[1]) + (p0[2] - q[2]) * (p0[2] - q[2]) - d2[0]) > 1e-12) return 1;
    return 0;
  } catch (const Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return e.code == MB_ERR_NO_DEVICE ? 2 : 1;
  }
}
