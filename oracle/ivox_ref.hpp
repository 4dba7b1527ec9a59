// ORACLE — TEST INFRASTRUCTURE ONLY (see linalg_ref.hpp for the rules).  PARITY UNPINNED.
//
// CPU restatement of the incremental hashed voxel map mimosa uses as its ICP target:
//   mimosa::lidar::IncrementalVoxelMapPCL  (mimosa/include/mimosa/lidar/incremental_voxel_map.hpp:22-51,
//   mimosa/src/lidar/incremental_voxel_map.cpp:14-62) wrapping gtsam_points::iVox
//   (= IncrementalVoxelMap<FlatContainer>), which is NOT vendored under /root/reference:
//   dependency github.com/ntnu-arl/gtsam_points, branch `minimal_updated`, no commit pinned
//   (README.md:47, .github/docker/ci-base.Dockerfile:41).
// The algorithm below restates the published upstream behaviour (koide3/gtsam_points
// include/gtsam_points/ann/{incremental_voxelmap.hpp,impl/incremental_voxelmap_impl.hpp,
// flat_container.hpp,knn_result.hpp}), anchored on mimosa's own in-tree copies of the helpers:
//   fast_floor                 mimosa/include/mimosa/lidar/utils.hpp:218-222
//   FlatContainerMinimal::add  mimosa/include/mimosa/lidar/utils.hpp:260-278  (cap check BEFORE the
//                              distance check, strict '<' on squared distance, first come wins)
//   XORVector3iHash            mimosa/include/mimosa/lidar/utils.hpp:228-238  (not parity relevant)
// and on mimosa's call sites: geometric.cpp:23-28 (leaf, lru_horizon, neighbor_voxel_mode,
// min_dist_in_cell), geometric.cpp:483-495 (f32 world transform, snapshot copy, insert),
// incremental_voxel_map.cpp:26-32 (knn_search returns found == k), geometric_factor.hpp:184 (point(i)).
//
// Assumption set — cannot be checked against the fork offline:
//  (1) coord = fast_floor(p * (1/leaf)), voxel created on first touch, id = index in a flat vector;
//  (2) add(): reject when count >= cap(20); reject when any stored point is closer than min_dist
//      (squared, strict '<'); else append;
//  (3) insert() refreshes voxel.lru = lru_counter on every touched voxel, then ++lru_counter and every
//      lru_clear_cycle(=10) inserts removes voxels with lru + lru_horizon < lru_counter, compacting the
//      flat vector (ids shift down, order preserved) and rebuilding the hash;
//  (4) neighbour offsets: mode 1 = centre; 7 = centre,+x,-x,+y,-y,+z,-z; 19 = nested i,j,k in -1..1
//      skipping |i|=|j|=|k|=1; 27 = nested i,j,k, all;
//  (5) k-NN visits offsets in that order and stored points in order; a candidate with
//      d2 >= current worst (initially max_sq_dist) is ignored, otherwise it is insertion-sorted
//      ascending with strict '<' (equal distances keep visiting order);
//  (6) global point index = (voxel_id << 32) | point_id;
//  (7) squared distances are evaluated on homogeneous 4-vectors (w = 1 - 1 = 0) with Eigen's
//      vectorised fixed-size reduction order (x^2 + z^2) + (y^2 + w^2);
//  (8) copying the map deep-copies voxel contents (the evident intent of geometric.cpp:494).
#pragma once
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <unordered_map>
#include <vector>

#include "linalg_ref.hpp"

namespace mimosa_oracle {

struct Coord {
  int32_t x, y, z;
  bool operator==(const Coord& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct CoordHash {  // Teschner et al. XOR hash, utils.hpp:228-238
  size_t operator()(const Coord& c) const {
    const size_t p1 = 9132043225175502913ull, p2 = 7277549399757405689ull, p3 = 6673468629021231217ull;
    return (size_t)((size_t)(int64_t)c.x * p1) ^ ((size_t)(int64_t)c.y * p2) ^ ((size_t)(int64_t)c.z * p3);
  }
};

// utils.hpp:218-222: int(x) - (x < int(x)).
inline int32_t fast_floor1(double x) {
  const int32_t i = (int32_t)x;
  return i - (x < (double)i ? 1 : 0);
}

// (x^2 + z^2) + (y^2 + 0): assumption (7).
inline double sqdist4(const V3& a, const V3& b) {
  const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return (dx * dx + dz * dz) + (dy * dy + 0.0);
}

inline std::vector<Coord> neighbor_offsets(int mode) {
  std::vector<Coord> o;
  if (mode == 1) {
    o.push_back({0, 0, 0});
  } else if (mode == 7) {
    o = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  } else {
    for (int i = -1; i <= 1; ++i)
      for (int j = -1; j <= 1; ++j)
        for (int k = -1; k <= 1; ++k) {
          if (mode == 19 && i != 0 && j != 0 && k != 0) continue;
          o.push_back({i, j, k});
        }
  }
  return o;
}

struct VoxelRef {
  Coord coord;
  uint64_t lru;
  std::vector<V3> pts;  // stored as double, values are f32-exact (incremental_voxel_map.cpp:40-48)
};

class IVoxRef {
 public:
  IVoxRef(double leaf, double min_dist, int cap, int nbr_mode, uint64_t lru_horizon)
      : inv_leaf_(1.0 / leaf),
        min_sq_dist_(min_dist * min_dist),
        cap_(cap),
        lru_horizon_(lru_horizon),
        offsets_(neighbor_offsets(nbr_mode)) {}

  // Deep snapshot (assumption 8).
  IVoxRef(const IVoxRef& o) = default;

  void insert(const float* xyz, size_t n, size_t stride_bytes) {
    for (size_t i = 0; i < n; ++i) {
      const float* f = (const float*)((const char*)xyz + i * stride_bytes);
      const V3 p{(double)f[0], (double)f[1], (double)f[2]};
      const Coord c{fast_floor1(p.x * inv_leaf_), fast_floor1(p.y * inv_leaf_), fast_floor1(p.z * inv_leaf_)};
      auto it = index_.find(c);
      if (it == index_.end()) {
        it = index_.emplace(c, (uint32_t)flat_.size()).first;
        flat_.push_back(VoxelRef{c, lru_counter_, {}});
        flat_.back().pts.reserve(cap_);
      }
      VoxelRef& v = flat_[it->second];
      v.lru = lru_counter_;
      if ((int)v.pts.size() >= cap_) continue;
      bool too_close = false;
      for (const V3& q : v.pts)
        if (sqdist4(q, p) < min_sq_dist_) {
          too_close = true;
          break;
        }
      if (!too_close) v.pts.push_back(p);
    }
    if ((++lru_counter_) % lru_clear_cycle_ == 0) {
      size_t w = 0;
      for (size_t r = 0; r < flat_.size(); ++r) {
        if (flat_[r].lru + lru_horizon_ < lru_counter_) continue;
        if (w != r) flat_[w] = std::move(flat_[r]);
        ++w;
      }
      if (w != flat_.size()) {
        flat_.resize(w);
        index_.clear();
        for (size_t i = 0; i < flat_.size(); ++i) index_[flat_[i].coord] = (uint32_t)i;
      }
    }
  }

  // Returns the number found; idx/d2 have room for k entries (unfilled: ~0 / max_sq_dist).
  int knn(const V3& q, int k, uint64_t* idx, double* d2,
          double max_sq_dist = std::numeric_limits<double>::max()) const {
    for (int i = 0; i < k; ++i) {
      idx[i] = ~0ull;
      d2[i] = max_sq_dist;
    }
    int found = 0;
    const Coord c{fast_floor1(q.x * inv_leaf_), fast_floor1(q.y * inv_leaf_), fast_floor1(q.z * inv_leaf_)};
    for (const Coord& o : offsets_) {
      const auto it = index_.find(Coord{c.x + o.x, c.y + o.y, c.z + o.z});
      if (it == index_.end()) continue;
      const VoxelRef& v = flat_[it->second];
      for (size_t j = 0; j < v.pts.size(); ++j) {
        const double d = sqdist4(v.pts[j], q);
        if (d >= d2[k - 1]) continue;
        int loc = found < k - 1 ? found : k - 1;
        for (; loc > 0 && d < d2[loc - 1]; --loc) {
          idx[loc] = idx[loc - 1];
          d2[loc] = d2[loc - 1];
        }
        idx[loc] = ((uint64_t)it->second << 32) | (uint64_t)j;
        d2[loc] = d;
        if (found < k) ++found;
      }
    }
    return found;
  }

  const V3& point(uint64_t gidx) const { return flat_[gidx >> 32].pts[gidx & 0xffffffffull]; }

  size_t num_voxels() const { return flat_.size(); }
  size_t num_points() const {
    size_t n = 0;
    for (const auto& v : flat_) n += v.pts.size();
    return n;
  }
  uint64_t lru_counter() const { return lru_counter_; }
  int cap() const { return cap_; }
  const std::vector<VoxelRef>& voxels() const { return flat_; }

  // Test-infrastructure loader: rebuild the structure from a dump (id order preserved).
  void load_raw(const int32_t* coords, const int32_t* counts, const uint32_t* lru, const float* pts,
                size_t n_vox, uint64_t lru_counter) {
    flat_.clear();
    index_.clear();
    flat_.reserve(n_vox);
    index_.reserve(n_vox * 2);
    for (size_t v = 0; v < n_vox; ++v) {
      VoxelRef vr{Coord{coords[3 * v], coords[3 * v + 1], coords[3 * v + 2]}, lru ? lru[v] : 0, {}};
      vr.pts.reserve(cap_);
      for (int j = 0; j < counts[v]; ++j) {
        const float* f = pts + ((size_t)v * cap_ + j) * 3;
        vr.pts.push_back(V3{(double)f[0], (double)f[1], (double)f[2]});
      }
      index_[vr.coord] = (uint32_t)v;
      flat_.push_back(std::move(vr));
    }
    lru_counter_ = lru_counter;
  }

 private:
  double inv_leaf_, min_sq_dist_;
  int cap_;
  uint64_t lru_horizon_;
  uint64_t lru_clear_cycle_ = 10;
  uint64_t lru_counter_ = 0;
  std::vector<Coord> offsets_;
  std::vector<VoxelRef> flat_;
  std::unordered_map<Coord, uint32_t, CoordHash> index_;
};

// Geometric::downsample (mimosa/src/lidar/geometric.cpp:55-126): one-shot greedy voxel thinning of the
// scan with the same add() rule; output = indices into the input, voxel creation order then
// in-voxel order.
inline std::vector<uint32_t> downsample_ref(const float* xyz, size_t n, size_t stride_bytes, double leaf,
                                            size_t cap, double min_dist) {
  const double inv_leaf = 1.0 / leaf, min_sq = min_dist * min_dist;
  struct Cell {
    std::vector<V3> pts;
    std::vector<uint32_t> ids;
  };
  std::vector<Cell> cells;
  std::unordered_map<Coord, uint32_t, CoordHash> index;
  index.reserve(n / 2 + 1);
  for (size_t i = 0; i < n; ++i) {
    const float* f = (const float*)((const char*)xyz + i * stride_bytes);
    const V3 p{(double)f[0], (double)f[1], (double)f[2]};
    const Coord c{fast_floor1(p.x * inv_leaf), fast_floor1(p.y * inv_leaf), fast_floor1(p.z * inv_leaf)};
    auto it = index.find(c);
    if (it == index.end()) {
      it = index.emplace(c, (uint32_t)cells.size()).first;
      cells.emplace_back();
    }
    Cell& cell = cells[it->second];
    if (cell.pts.size() >= cap) continue;
    bool too_close = false;
    for (const V3& q : cell.pts) {
      // utils.hpp:267: (existing - p).squaredNorm() on Vector3d -> fixed-size 3 reduction order.
      const V3 d = q - p;
      if (sqnorm(d) < min_sq) {
        too_close = true;
        break;
      }
    }
    if (too_close) continue;
    cell.pts.push_back(p);
    cell.ids.push_back((uint32_t)i);
  }
  std::vector<uint32_t> out;
  for (const Cell& c : cells) out.insert(out.end(), c.ids.begin(), c.ids.end());
  return out;
}

}  // namespace mimosa_oracle
