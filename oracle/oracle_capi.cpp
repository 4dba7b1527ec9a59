// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the CPU restatement so tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs can drive it through ctypes.  Never linked into the
// product library.  PARITY UNPINNED (see icp_factor_ref.hpp).
//
// Struct layouts deliberately equal include/mimosa_b200.h's mb_icp_config / mb_linearization /
// mb_icp_trace so one ctypes definition serves both sides of a parity test.
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>

#include "icp_factor_ref.hpp"

using namespace mimosa_oracle;

namespace {
struct MapHandle {
  std::shared_ptr<IVoxRef> map;
};
struct FactorHandle {
  std::unique_ptr<IcpFactorRef> f;
};
Pose make_pose(const double* R, const double* t) {
  Pose T;
  for (int i = 0; i < 9; ++i) T.R.m[i] = R[i];
  T.t = V3{t[0], t[1], t[2]};
  return T;
}
}  // namespace

extern "C" {

void* orc_map_create(float leaf, float min_dist, int cap, int nbr_mode, uint64_t lru_horizon) {
  auto* h = new MapHandle;
  h->map = std::make_shared<IVoxRef>((double)leaf, (double)min_dist, cap, nbr_mode, lru_horizon);
  return h;
}
void orc_map_release(void* m) { delete (MapHandle*)m; }
void* orc_map_snapshot(void* m) {
  auto* h = new MapHandle;
  h->map = std::make_shared<IVoxRef>(*((MapHandle*)m)->map);
  return h;
}
void orc_map_insert(void* m, const float* xyz, size_t n, size_t stride_bytes) {
  ((MapHandle*)m)->map->insert(xyz, n, stride_bytes);
}
size_t orc_map_num_voxels(void* m) { return ((MapHandle*)m)->map->num_voxels(); }
size_t orc_map_num_points(void* m) { return ((MapHandle*)m)->map->num_points(); }
uint64_t orc_map_lru_counter(void* m) { return ((MapHandle*)m)->map->lru_counter(); }
// coords: n_vox*3 int32, counts: n_vox int32, lru: n_vox uint32, pts: n_vox*cap*3 float (zero padded).
void orc_map_download(void* m, int32_t* coords, int32_t* counts, uint32_t* lru, float* pts) {
  const IVoxRef& map = *((MapHandle*)m)->map;
  const int cap = map.cap();
  size_t v = 0;
  for (const VoxelRef& vr : map.voxels()) {
    coords[3 * v] = vr.coord.x;
    coords[3 * v + 1] = vr.coord.y;
    coords[3 * v + 2] = vr.coord.z;
    counts[v] = (int32_t)vr.pts.size();
    if (lru) lru[v] = (uint32_t)vr.lru;
    for (int j = 0; j < cap; ++j) {
      float* f = pts + (v * cap + j) * 3;
      if (j < (int)vr.pts.size()) {
        f[0] = (float)vr.pts[j].x;
        f[1] = (float)vr.pts[j].y;
        f[2] = (float)vr.pts[j].z;
      } else {
        f[0] = f[1] = f[2] = 0.f;
      }
    }
    ++v;
  }
}
void orc_map_load_raw(void* m, const int32_t* coords, const int32_t* counts, const uint32_t* lru,
                      const float* pts, size_t n_vox, uint64_t lru_counter) {
  ((MapHandle*)m)->map->load_raw(coords, counts, lru, pts, n_vox, lru_counter);
}
// ok[i] = (found == k)  (incremental_voxel_map.cpp:31)
void orc_map_knn(void* m, const double* q, size_t nq, int k, uint64_t* idx, double* d2, uint8_t* ok,
                 int n_threads) {
  const IVoxRef& map = *((MapHandle*)m)->map;
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(static)
  for (long i = 0; i < (long)nq; ++i) {
    const int found = map.knn(V3{q[3 * i], q[3 * i + 1], q[3 * i + 2]}, k, idx + (size_t)i * k, d2 + (size_t)i * k);
    ok[i] = found == k;
  }
}

void* orc_factor_create(void* m, const void* pts, size_t n, size_t stride_bytes, const RegistrationConfigRef* cfg) {
  auto* h = new FactorHandle;
  h->f = std::make_unique<IcpFactorRef>(((MapHandle*)m)->map, pts, n, stride_bytes, *cfg);
  return h;
}
void orc_factor_release(void* f) { delete (FactorHandle*)f; }
void orc_factor_reset(void* f) { ((FactorHandle*)f)->f->reset_state(); }
void orc_factor_linearize(void* f, const double* R, const double* t, const double* gravity_unit,
                          LinearizationRef* out, int n_threads) {
  const V3 g{gravity_unit[0], gravity_unit[1], gravity_unit[2]};
  ((FactorHandle*)f)->f->linearize(make_pose(R, t), g, *out, n_threads > 0 ? n_threads : 4, n_threads > 0);
}
// Any pointer may be null.  vectors are n*3 doubles, idx is n*k uint64.
void orc_factor_download_state(void* f, uint8_t* status, double* p_da, double* mean, double* normal,
                               double* loc_rot, double* loc_trans, uint64_t* knn_idx) {
  const IcpFactorRef& F = *((FactorHandle*)f)->f;
  const size_t n = F.size();
  if (status) std::memcpy(status, F.status().data(), n);
  auto cp = [n](double* dst, const std::vector<V3>& src) {
    if (!dst) return;
    for (size_t i = 0; i < n; ++i) {
      dst[3 * i] = src[i].x;
      dst[3 * i + 1] = src[i].y;
      dst[3 * i + 2] = src[i].z;
    }
  };
  cp(p_da, F.p_da());
  cp(mean, F.mean());
  cp(normal, F.normal());
  cp(loc_rot, F.loc_rot());
  cp(loc_trans, F.loc_trans());
  if (knn_idx) std::memcpy(knn_idx, F.knn_idx().data(), F.knn_idx().size() * sizeof(uint64_t));
}
// T (R row-major 9, t 3) is updated in place.  Returns wall seconds of the loop.
double orc_icp_run(void* f, double* R, double* t, int iters, double lambda, IcpTraceRef* trace, int n_threads) {
  Pose T = make_pose(R, t);
  const auto t0 = std::chrono::steady_clock::now();
  icp_run_ref(*((FactorHandle*)f)->f, T, iters, lambda, trace, n_threads > 0 ? n_threads : 4, n_threads > 0);
  const auto t1 = std::chrono::steady_clock::now();
  for (int i = 0; i < 9; ++i) R[i] = T.R.m[i];
  t[0] = T.t.x;
  t[1] = T.t.y;
  t[2] = T.t.z;
  return std::chrono::duration<double>(t1 - t0).count();
}

// out_idx has room for n entries; returns the number kept.  (geometric.cpp:55-126)
size_t orc_downsample(const float* xyz, size_t n, size_t stride_bytes, float leaf, size_t cap, float min_dist,
                      uint32_t* out_idx) {
  const auto v = downsample_ref(xyz, n, stride_bytes, (double)leaf, cap, (double)min_dist);
  std::memcpy(out_idx, v.data(), v.size() * sizeof(uint32_t));
  return v.size();
}

// lidar::Manager::deskewPoints per-point part (mimosa/src/lidar/manager.cpp:494-509) and the float transforms of
// Geometric::preprocess (geometric.cpp:153-160) / updateMap (geometric.cpp:483-490): p <- R p + t in binary32,
// 3-term sums as a0 + (a1 + a2).  poses: n_poses x 12 floats (R row-major, t); pose_index may be null (pose 0).
void orc_transform_f32(void* pts, size_t n, size_t stride_bytes, const uint32_t* pose_index, const float* poses) {
  for (size_t i = 0; i < n; ++i) {
    float* f = (float*)((char*)pts + i * stride_bytes);
    const float* P = poses + (pose_index ? (size_t)pose_index[i] * 12 : 0);
    const float x = f[0], y = f[1], z = f[2];
    f[0] = (P[0] * x + (P[1] * y + P[2] * z)) + P[9];
    f[1] = (P[3] * x + (P[4] * y + P[5] * z)) + P[10];
    f[2] = (P[6] * x + (P[7] * y + P[8] * z)) + P[11];
  }
}

// ---- small dense helpers exposed for the known-answer tests ------------------------------------------
int orc_eigh3(const double* A, double* lam, double* V) {
  M3 a, v;
  for (int i = 0; i < 9; ++i) a.m[i] = A[i];
  const bool ok = eigh3(a, lam, v);
  for (int i = 0; i < 9; ++i) V[i] = v.m[i];
  return ok ? 1 : 0;
}
void orc_se3_expmap(const double* xi, double* R, double* t) {
  const Pose P = se3_expmap(xi);
  for (int i = 0; i < 9; ++i) R[i] = P.R.m[i];
  t[0] = P.t.x;
  t[1] = P.t.y;
  t[2] = P.t.z;
}
void orc_se3_retract(double* R, double* t, const double* xi) {
  const Pose O = se3_retract(make_pose(R, t), xi);
  for (int i = 0; i < 9; ++i) R[i] = O.R.m[i];
  t[0] = O.t.x;
  t[1] = O.t.y;
  t[2] = O.t.z;
}
int orc_solve6(const double* H, double lambda, const double* rhs, double* x) {
  return solve6_ldlt(H, lambda, rhs, x) ? 1 : 0;
}
int orc_fast_floor(double x) { return fast_floor1(x); }
int orc_max_threads() { return omp_get_max_threads(); }
size_t orc_sizeof(int which) {
  return which == 0 ? sizeof(RegistrationConfigRef) : which == 1 ? sizeof(LinearizationRef) : which == 2 ? sizeof(IcpTraceRef) : 0;
}

}  // extern "C"
