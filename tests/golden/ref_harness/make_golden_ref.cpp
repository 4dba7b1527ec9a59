// make_golden_ref.cpp — PIN-READINESS HARNESS.  NOT BUILT OR RUN IN THIS REPOSITORY'S IMAGE (no ROS / PCL / GTSAM /
// gtsam_points / Eigen here); written so that anybody with the reference's CI image
// (/root/reference/.github/docker/ci-base.Dockerfile: ROS Noetic + ntnu-arl/gtsam feature/imu_factor_with_gravity +
// ntnu-arl/gtsam_points minimal_updated) can replace "parity unpinned" by "parity pinned" in an afternoon.
//
// It links the REAL mimosa (lidar::ICPFactor, lidar::IncrementalVoxelMapPCL -> gtsam_points::iVox) and reruns config C1
// on the inputs export_c1_inputs.py writes, through the reference's own code path:
//   * the map is built by the reference's insert (the same point chunks in the same order) and dumped voxel by voxel —
//     checks the oracle's iVox assumptions 1-3, 6, 7 of oracle/ivox_ref.hpp (coordinates, creation-order ids, cap before
//     distance test, (voxel << 32) | point indices, deep copy);
//   * every scan point's 5 nearest neighbours are queried through IncrementalVoxelMapPCL::knn_search
//     (incremental_voxel_map.cpp:26-32) — assumptions 4, 5 (neighbour order, strict '<' ties);
//   * five iterations of ICPFactor::linearize (geometric_factor.hpp:231-562) with the harness' Gauss-Newton step
//     delta = (H + lambda I)^-1 g, T <- T * Pose3::Expmap(delta), each iteration's H, g, f, statuses, means, normals.
// Outputs are raw little-endian arrays in <dir>/ref_*.bin; convert_ref_dump.py turns them into an .npz with the schema of
// tests/golden/c1_golden.npz, and `MB_GOLDEN=<that file> pytest tests/test_golden_oracle.py tests/test_gpu_parity.py -k golden`
// then compares the oracle and the CUDA path with the REFERENCE's numbers.
//
// Build inside the CI image, from the catkin workspace that contains mimosa (see README.md next to this file):
//   g++ -O2 -std=c++17 make_golden_ref.cpp -I<ws>/src/mimosa/mimosa/include $(pkg-config --cflags eigen3) \
//       -I/opt/ros/noetic/include -I<ws>/install/include -L<ws>/install/lib -lmimosa -lgtsam -lgtsam_points \
//       $(pkg-config --libs pcl_common) -fopenmp -o make_golden_ref
#include <gtsam/geometry/Pose3.h>
#include <gtsam/geometry/Unit3.h>
#include <gtsam/inference/Symbol.h>
#include <gtsam/linear/HessianFactor.h>
#include <gtsam/nonlinear/Values.h>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "mimosa/lidar/geometric_config.hpp"
#include "mimosa/lidar/geometric_factor.hpp"
#include "mimosa/lidar/incremental_voxel_map.hpp"

using gtsam::symbol_shorthand::G;
using gtsam::symbol_shorthand::X;
using namespace mimosa::lidar;

template <typename T>
static std::vector<T> read_bin(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("cannot open " + path);
  const size_t bytes = (size_t)f.tellg();
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), bytes);
  return v;
}
template <typename T>
static void write_bin(const std::string& path, const std::vector<T>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(T));
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: make_golden_ref <dir with c1_*.bin from export_c1_inputs.py>\n");
    return 2;
  }
  const std::string dir = std::string(argv[1]) + "/";
  const auto chunk_sizes = read_bin<int64_t>(dir + "c1_chunk_sizes.bin");  // points per insert() call
  const auto chunk_xyz = read_bin<float>(dir + "c1_chunks_xyz.bin");       // all chunks, xyz float32
  const auto scan_xyz = read_bin<float>(dir + "c1_scan_xyz.bin");          // n x 3 float32, body frame
  const auto pose0 = read_bin<double>(dir + "c1_pose0.bin");               // R row-major (9), t (3), iters, lambda

  // ---- the map, exactly as Geometric's constructor and updateMap build it (geometric.cpp:22-28, 487-495) ------------
  RegistrationConfig cfg;  // hornbill values, mimosa/config/hornbill/params.yaml:86-102
  cfg.source_voxel_grid_filter_leaf_size = cfg.target_ivox_map_leaf_size = 1.0f;
  cfg.source_voxel_grid_min_dist_in_voxel = cfg.target_ivox_map_min_dist_in_voxel = 0.2f;
  cfg.num_corres_points = 5;
  cfg.max_corres_distance = 1.0f;
  cfg.plane_validity_distance = 0.07f;
  cfg.lidar_point_noise_std_dev = 0.07f;
  cfg.use_huber = true;
  cfg.huber_threshold = 1.345f;
  cfg.reg_4_dof = false;
  cfg.project_on_degneneracy = false;
  auto map = std::make_shared<IncrementalVoxelMapPCL>(cfg.target_ivox_map_leaf_size);
  map->underlying()->set_lru_horizon(1000);
  map->underlying()->set_neighbor_voxel_mode(19);
  map->underlying()->voxel_insertion_setting().set_min_dist_in_cell(cfg.target_ivox_map_min_dist_in_voxel);
  size_t off = 0;
  for (int64_t n : chunk_sizes) {
    std::vector<Eigen::Vector3f> pts((size_t)n);
    for (int64_t i = 0; i < n; ++i) pts[i] << chunk_xyz[3 * (off + i)], chunk_xyz[3 * (off + i) + 1], chunk_xyz[3 * (off + i) + 2];
    off += (size_t)n;
    auto frame = std::make_shared<gtsam_points::PointCloudCPU>(pts);
    map = std::make_shared<IncrementalVoxelMapPCL>(*map);  // the deep copy of geometric.cpp:494 before every insert
    map->underlying()->insert(*frame);
  }
  {  // every stored point in index order: getCloud() walks voxel_data() (incremental_voxel_map.cpp:34-38)
    auto cloud = map->getCloud();
    std::vector<float> xyz;
    for (const auto& p : cloud->points) xyz.insert(xyz.end(), {p.x, p.y, p.z});
    write_bin(dir + "ref_map_points_xyz.bin", xyz);
  }

  // ---- k-NN of every transformed scan point at the start pose, through the wrapper ---------------------------------
  const size_t n = scan_xyz.size() / 3;
  gtsam::Matrix3 R0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R0(r, c) = pose0[3 * r + c];
  gtsam::Pose3 T(gtsam::Rot3(R0), gtsam::Point3(pose0[9], pose0[10], pose0[11]));
  const int iters = (int)pose0[12];
  const double lambda = pose0[13];
  {
    std::vector<uint64_t> idx(n * 5, ~0ull);
    std::vector<double> d2(n * 5, 0.0);
    std::vector<uint8_t> ok(n, 0);
    std::vector<double> nn_xyz(n * 5 * 3, 0.0);
    for (size_t i = 0; i < n; ++i) {
      const Eigen::Vector3d q = T.transformFrom(Eigen::Vector3d(scan_xyz[3 * i], scan_xyz[3 * i + 1], scan_xyz[3 * i + 2]));
      std::vector<size_t> k_idx(5);
      std::vector<double> k_d2(5);
      ok[i] = map->knn_search(q, 5, k_idx, k_d2) ? 1 : 0;
      for (int j = 0; j < 5; ++j) {
        idx[5 * i + j] = ok[i] ? (uint64_t)k_idx[j] : ~0ull;
        d2[5 * i + j] = k_d2[j];
        if (ok[i]) {
          const auto p = map->underlying()->point(k_idx[j]);  // geometric_factor.hpp:184
          for (int a = 0; a < 3; ++a) nn_xyz[(5 * i + j) * 3 + a] = p[a];
        }
      }
    }
    write_bin(dir + "ref_knn_idx.bin", idx);
    write_bin(dir + "ref_knn_d2.bin", d2);
    write_bin(dir + "ref_knn_ok.bin", ok);
    write_bin(dir + "ref_knn_points.bin", nn_xyz);
  }

  // ---- the factor and the Gauss-Newton harness ------------------------------------------------------------------------
  pcl::PointCloud<Point> cloud;
  cloud.resize(n);
  for (size_t i = 0; i < n; ++i) {
    cloud.points[i].x = scan_xyz[3 * i];
    cloud.points[i].y = scan_xyz[3 * i + 1];
    cloud.points[i].z = scan_xyz[3 * i + 2];
  }
  ICPFactor factor(X(0), map, cloud, cfg);
  std::vector<double> H_all, g_all, f_all, pose_all;
  std::vector<uint8_t> status_all;
  std::vector<double> mean_all, normal_all;
  for (int it = 0; it < iters; ++it) {
    gtsam::Values values;
    values.insert(X(0), T);
    values.insert(G(0), gtsam::Unit3(0, 0, -1));
    auto gf = std::dynamic_pointer_cast<gtsam::HessianFactor>(factor.linearize(values));
    const gtsam::Matrix H = gf->information();  // 6x6 = J^T J
    const gtsam::Vector g = gf->linearTerm();   // = -J^T e
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) H_all.push_back(H(r, c));
    for (int r = 0; r < 6; ++r) g_all.push_back(g(r));
    f_all.push_back(gf->constantTerm());
    for (auto s : factor.getStatuses()) status_all.push_back((uint8_t)s);
    for (const auto& m : factor.getCorresMeansTarget())
      for (int a = 0; a < 3; ++a) mean_all.push_back(m(a));
    for (const auto& nrm : factor.getCorresNormalsTarget())
      for (int a = 0; a < 3; ++a) normal_all.push_back(nrm(a));
    gtsam::Matrix A = H;
    for (int r = 0; r < 6; ++r) A(r, r) += lambda;
    const gtsam::Vector delta = A.ldlt().solve(g);
    T = T * gtsam::Pose3::Expmap(delta);
    const gtsam::Matrix3 Rn = T.rotation().matrix();
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) pose_all.push_back(Rn(r, c));
    for (int a = 0; a < 3; ++a) pose_all.push_back(T.translation()(a));
  }
  write_bin(dir + "ref_tr_H.bin", H_all);
  write_bin(dir + "ref_tr_g.bin", g_all);
  write_bin(dir + "ref_tr_f.bin", f_all);
  write_bin(dir + "ref_tr_pose.bin", pose_all);
  write_bin(dir + "ref_status.bin", status_all);
  write_bin(dir + "ref_mean.bin", mean_all);
  write_bin(dir + "ref_normal.bin", normal_all);
  std::printf("wrote reference dump for %zu scan points, %d iterations to %s\n", n, iters, dir.c_str());
  return 0;
}
