// Device-resident incremental hashed voxel map: build (sort + per-voxel greedy insert), LRU eviction,
// snapshot, restricted k-NN.  See mb_map.cuh for the layout and include/mimosa_b200.h for the reference
// interfaces each entry point replaces.
//
// Insert reproduces the reference's *sequential* semantics (gtsam_points::IncrementalVoxelMap::insert as
// used at mimosa/src/lidar/geometric.cpp:495; the per-voxel rule is restated in-tree at
// mimosa/include/mimosa/lidar/utils.hpp:260-278) on a parallel machine:
//   * points only interact inside one voxel, so a STABLE radix sort by voxel key turns the input into one
//     run per voxel with the input order preserved inside the run;
//   * voxel ids are creation order = order of each new voxel's first point in the input, recovered with a
//     flag-and-scan over the original positions;
//   * one warp walks each run in order applying "full? -> reject; any stored point closer than min_dist?
//     -> reject; else append" with the stored points spread over the lanes.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "mb_map.cuh"

namespace mb {

namespace {

constexpr int kCoordBias = 1 << 20;

__device__ __forceinline__ uint64_t pack_key(int x, int y, int z) {
  return ((uint64_t)(uint32_t)(x + kCoordBias) << 42) | ((uint64_t)(uint32_t)(y + kCoordBias) << 21) |
         (uint64_t)(uint32_t)(z + kCoordBias);
}
__device__ __forceinline__ int3 unpack_key(uint64_t k) {
  return make_int3((int)((k >> 42) & 0x1fffff) - kCoordBias, (int)((k >> 21) & 0x1fffff) - kCoordBias,
                   (int)(k & 0x1fffff) - kCoordBias);
}

__global__ void k_make_keys(const unsigned char* __restrict__ raw, size_t n, size_t stride, double inv_leaf,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int* __restrict__ err) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* f = (const float*)(raw + i * stride);
  const int cx = fast_floor((double)f[0] * inv_leaf), cy = fast_floor((double)f[1] * inv_leaf),
            cz = fast_floor((double)f[2] * inv_leaf);
  if (abs(cx) >= kCoordBias || abs(cy) >= kCoordBias || abs(cz) >= kCoordBias || !isfinite(f[0]) ||
      !isfinite(f[1]) || !isfinite(f[2]))
    *err = 1;
  keys[i] = pack_key(cx, cy, cz);
  vals[i] = (uint32_t)i;
}

__global__ void k_mark_heads(const uint64_t* __restrict__ keys, size_t n, uint32_t* __restrict__ head) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void k_run_starts(const uint32_t* __restrict__ head, const uint32_t* __restrict__ run_id, size_t n,
                             uint32_t* __restrict__ run_start) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (head[i]) run_start[run_id[i]] = (uint32_t)i;
  // exclusive scan: a head sees its own run index, a non-head already counts its run's head
  if (i == n - 1) run_start[run_id[i] + head[i]] = (uint32_t)n;  // sentinel at index n_runs
}

// counters[0] = number of runs, counters[1] = number of new voxels
__global__ void k_count2(const uint32_t* __restrict__ a_flag, const uint32_t* __restrict__ a_scan,
                         const uint32_t* __restrict__ b_flag, const uint32_t* __restrict__ b_scan, size_t n,
                         uint32_t* __restrict__ counters) {
  if (a_flag) counters[0] = a_flag[n - 1] + a_scan[n - 1];
  if (b_flag) counters[1] = b_flag[n - 1] + b_scan[n - 1];
}

__global__ void k_lookup_runs(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                              const uint32_t* __restrict__ run_start, uint32_t n_runs, const int4* __restrict__ table,
                              uint32_t mask, uint32_t* __restrict__ run_packed, uint32_t* __restrict__ new_flag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const uint32_t s = run_start[r];
  const int3 c = unpack_key(keys[s]);
  const uint32_t p = table_find(table, mask, c.x, c.y, c.z);
  run_packed[r] = p;
  if (p == kEmpty) new_flag[vals[s]] = 1u;  // stable sort: the run's first element is its earliest input point
}

__device__ __forceinline__ uint32_t table_claim(int4* table, uint32_t mask, int x, int y, int z, uint32_t packed) {
  uint32_t h = hash_coord(x, y, z) & mask;
  while (true) {
    const unsigned old = atomicCAS((unsigned*)&table[h].w, kEmpty, packed);
    if (old == kEmpty) {
      table[h].x = x;
      table[h].y = y;
      table[h].z = z;
      return h;
    }
    h = (h + 1) & mask;
  }
}

__global__ void k_create_voxels(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                const uint32_t* __restrict__ run_start, uint32_t n_runs,
                                const uint32_t* __restrict__ run_packed, const uint32_t* __restrict__ new_rank,
                                uint32_t base_id, int lru, int4* __restrict__ table, uint32_t mask,
                                int4* __restrict__ info, int32_t* __restrict__ count, uint32_t* __restrict__ epos,
                                uint32_t* __restrict__ run_slot) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const uint32_t p = run_packed[r];
  if (p != kEmpty) {
    run_slot[r] = p >> kCountBits;
    return;
  }
  const uint32_t s = run_start[r];
  const int3 c = unpack_key(keys[s]);
  const uint32_t id = base_id + new_rank[vals[s]];
  info[id] = make_int4(c.x, c.y, c.z, lru);
  count[id] = 0;
  epos[id] = table_claim(table, mask, c.x, c.y, c.z, id << kCountBits);
  run_slot[r] = id;
}

// One warp per run.  Lane q holds stored point q of the voxel (as doubles).
// The greedy accept rule (FlatContainerMinimal::add, mimosa/include/mimosa/lidar/utils.hpp:240-294: reject when the
// voxel is full, reject when a kept point lies within min_dist, else append) is sequential in the order of the
// run, but almost every candidate of a long run is rejected by the points kept BEFORE it.  So a warp takes 32
// consecutive candidates at a time: every lane tests its candidate against the kept set as it stood at the start of
// the chunk (shared memory, broadcast reads), and only the survivors are resolved one by one in run order — the
// first survivor is accepted, the later ones are re-tested against it, and so on.  Same result as the one-by-one
// loop (a 10 000-point run under the sensor costs ~300 chunk steps instead of 10 000 dependent steps).
constexpr int kRunWarps = 8;  // warps per block of the two run kernels (256 threads)
struct KeptSet {
  double x[32], y[32], z[32];
};

// Resolve one chunk: `alive` = this lane's candidate exists and is not close to any point kept before the chunk.
// Accepted candidates are appended to `ks` (count c) in lane order; returns the lanes whose candidates were accepted.
// kHomogeneous selects the squared-distance expression of the rule being reproduced: the map's sqdist4
// ((dx^2 + dz^2) + dy^2, homogeneous 4-vectors) or the scan downsample's 3-vector norm (dx^2 + (dy^2 + dz^2)).
template <bool kHomogeneous>
__device__ __forceinline__ double run_sqdist(double kx, double ky, double kz, double px, double py, double pz) {
  return kHomogeneous ? sqdist4(kx, ky, kz, px, py, pz) : sqnorm3(sub3(mk3(kx, ky, kz), mk3(px, py, pz)));
}
template <bool kHomogeneous>
__device__ __forceinline__ unsigned resolve_chunk(KeptSet& ks, int& c, int cap, double min_sq, bool alive, double px, double py,
                                                  double pz, int lane) {
  unsigned accepted = 0;
  unsigned mask = __ballot_sync(kFull, alive);
  while (mask != 0 && c < cap) {
    const int s = __ffs(mask) - 1;
    if (lane == s) {
      ks.x[c] = px;
      ks.y[c] = py;
      ks.z[c] = pz;
    }
    __syncwarp();
    accepted |= 1u << s;
    if (alive && lane > s) alive = !(run_sqdist<kHomogeneous>(ks.x[c], ks.y[c], ks.z[c], px, py, pz) < min_sq);
    ++c;
    mask = __ballot_sync(kFull, alive && lane > s);
  }
  return accepted;
}

__global__ void __launch_bounds__(kRunWarps * 32)
    k_insert_runs(const unsigned char* __restrict__ raw, size_t stride, const uint32_t* __restrict__ vals,
                  const uint32_t* __restrict__ run_start, uint32_t n_runs, const uint32_t* __restrict__ run_slot, int cap,
                  double min_sq, int lru, float4* __restrict__ pts, int4* __restrict__ info, int32_t* __restrict__ count,
                  const uint32_t* __restrict__ epos, int4* __restrict__ table) {
  __shared__ KeptSet s_kept[kRunWarps];
  const int lane = threadIdx.x & 31;
  KeptSet& ks = s_kept[threadIdx.x >> 5];
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_runs; r += warps) {
    const uint32_t slot = run_slot[r];
    int c = count[slot];
    if (lane < c) {  // the voxel's stored points are the kept set the run starts from
      const float4 p = pts[(size_t)slot * cap + lane];
      ks.x[lane] = p.x;
      ks.y[lane] = p.y;
      ks.z[lane] = p.z;
    }
    __syncwarp();
    const uint32_t t1 = run_start[r + 1];
    for (uint32_t t = run_start[r]; t < t1 && c < cap; t += 32) {
      const bool have = t + lane < t1;
      float fx = 0.f, fy = 0.f, fz = 0.f;
      if (have) {
        const float* f = (const float*)(raw + (size_t)vals[t + lane] * stride);
        fx = f[0], fy = f[1], fz = f[2];
      }
      const double px = (double)fx, py = (double)fy, pz = (double)fz;
      bool alive = have;
      for (int j = 0; j < c; ++j) alive = alive && !(run_sqdist<true>(ks.x[j], ks.y[j], ks.z[j], px, py, pz) < min_sq);
      const int c0 = c;
      const unsigned acc = resolve_chunk<true>(ks, c, cap, min_sq, alive, px, py, pz, lane);
      if ((acc >> lane) & 1u) pts[(size_t)slot * cap + c0 + __popc(acc & ((1u << lane) - 1u))] = make_float4(fx, fy, fz, 0.f);
      __syncwarp();
    }
    if (lane == 0) {
      count[slot] = c;
      info[slot].w = lru;
      table[epos[slot]].w = (int)((slot << kCountBits) | (uint32_t)c);
    }
    __syncwarp();
  }
}

__global__ void k_rehash(const int4* __restrict__ info, const int32_t* __restrict__ count, uint32_t n_vox,
                         int4* __restrict__ table, uint32_t mask, uint32_t* __restrict__ epos) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_vox) return;
  const int4 c = info[id];
  epos[id] = table_claim(table, mask, c.x, c.y, c.z, (id << kCountBits) | (uint32_t)count[id]);
}

__global__ void k_flag_keep(const int4* __restrict__ info, uint32_t n_vox, long long horizon, long long counter,
                            uint32_t* __restrict__ keep) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_vox) return;
  keep[id] = ((long long)info[id].w + horizon < counter) ? 0u : 1u;
}

// One warp per surviving voxel moves its bucket to the compacted id.
__global__ void k_compact(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ new_id, uint32_t n_vox,
                          int cap, const float4* __restrict__ pts, const int4* __restrict__ info,
                          const int32_t* __restrict__ count, float4* __restrict__ pts2, int4* __restrict__ info2,
                          int32_t* __restrict__ count2) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; id < n_vox; id += warps) {
    if (!keep[id]) continue;
    const uint32_t d = new_id[id];
    if (lane < cap) pts2[(size_t)d * cap + lane] = pts[(size_t)id * cap + lane];
    if (lane == 0) {
      info2[d] = info[id];
      count2[d] = count[id];
    }
  }
}

// ---- search mirror ------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread3(uint32_t v) {  // 21 bits -> every third bit
  uint64_t x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
constexpr int kBlockBias = 1 << 18;  // block coordinates are within +-2^18 (voxel coordinates within +-2^20)

// key = (Morton code of the block << 6) | cell
__global__ void k_mirror_keys(const int4* __restrict__ info, uint32_t n_vox, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ ids) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_vox) return;
  const int4 c = info[id];
  const uint32_t bx = (uint32_t)((c.x >> kBlockShift) + kBlockBias), by = (uint32_t)((c.y >> kBlockShift) + kBlockBias),
                 bz = (uint32_t)((c.z >> kBlockShift) + kBlockBias);
  const uint64_t morton = spread3(bx) | (spread3(by) << 1) | (spread3(bz) << 2);
  keys[id] = (morton << 6) | cell_of(c.x, c.y, c.z);
  ids[id] = id;
}

// One warp per mirror position: move the bucket, write meta; block heads are flagged for the scan.
__global__ void k_mirror_gather(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids, uint32_t n_vox,
                                int cap, const float4* __restrict__ pts, const int32_t* __restrict__ count,
                                float4* __restrict__ r_pts, uint32_t* __restrict__ r_meta, uint32_t* __restrict__ head) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_vox; s += warps) {
    const uint32_t id = ids[s];
    const int c = count[id];
    if (lane < cap) {
      float4 p = lane < c ? pts[(size_t)id * cap + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
      // the meta word (id << 5 | count) rides in the first point: no separate metadata load in the search
      if (lane == 0) p.w = __int_as_float((int)((id << kCountBits) | (uint32_t)c));
      r_pts[(size_t)s * cap + lane] = p;
    }
    if (lane == 0) {
      r_meta[s] = (id << kCountBits) | (uint32_t)c;
      head[s] = (s == 0 || (keys[s] >> 6) != (keys[s - 1] >> 6)) ? 1u : 0u;
    }
  }
}

// Every mirror position ORs its cell bit into its block's mask; block heads record base and coordinates.
__global__ void k_mirror_blocks(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                                const int4* __restrict__ info, const uint32_t* __restrict__ head,
                                const uint32_t* __restrict__ block_of, uint32_t n_vox, int4* __restrict__ blk_hdr,
                                unsigned long long* __restrict__ blk_mask) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_vox) return;
  const uint32_t b = block_of[s];  // exclusive scan of head: a head sees its own block index, others index + 1
  const uint32_t blk = head[s] ? b : b - 1;
  atomicOr(blk_mask + blk, 1ull << (keys[s] & 63ull));
  if (head[s]) {
    const int4 c = info[ids[s]];
    blk_hdr[blk] = make_int4(c.x >> kBlockShift, c.y >> kBlockShift, c.z >> kBlockShift, (int)s);
  }
}

__global__ void k_mirror_table(const int4* __restrict__ blk_hdr, const unsigned long long* __restrict__ blk_mask,
                               uint32_t n_blocks, int4* __restrict__ btab, uint32_t bmask) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const int4 hd = blk_hdr[b];
  uint32_t h = hash_coord(hd.x, hd.y, hd.z) & bmask;
  while (true) {
    const unsigned old = atomicCAS((unsigned*)&btab[2 * (size_t)h].w, kEmpty, (unsigned)hd.w);
    if (old == kEmpty) break;
    h = (h + 1) & bmask;
  }
  btab[2 * (size_t)h].x = hd.x;
  btab[2 * (size_t)h].y = hd.y;
  btab[2 * (size_t)h].z = hd.z;
  const unsigned long long m = blk_mask[b];
  btab[2 * (size_t)h + 1] = make_int4((int)(uint32_t)(m & 0xffffffffull), (int)(uint32_t)(m >> 32), 0, 0);
}

// ---- Geometric::downsample (mimosa/src/lidar/geometric.cpp:55-126) ----------------------------------
__global__ void k_ds_mark(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ run_start, uint32_t n_runs,
                          uint32_t* __restrict__ first_flag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_runs) first_flag[vals[run_start[r]]] = 1u;
}

// One warp per run (= per voxel of the one-shot grid).  FlatContainerMinimal::add (lidar/utils.hpp:260-278):
// full -> reject, any kept point with (kept - p).squaredNorm() < min_sq -> reject, else keep.
__global__ void __launch_bounds__(kRunWarps * 32)
    k_ds_runs(const unsigned char* __restrict__ raw, size_t stride, const uint32_t* __restrict__ vals,
              const uint32_t* __restrict__ run_start, uint32_t n_runs, const uint32_t* __restrict__ first_rank, int cap,
              double min_sq, uint32_t* __restrict__ kept, uint32_t* __restrict__ kept_count) {
  __shared__ KeptSet s_kept[kRunWarps];
  const int lane = threadIdx.x & 31;
  KeptSet& ks = s_kept[threadIdx.x >> 5];
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_runs; r += warps) {
    const uint32_t t0 = run_start[r], t1 = run_start[r + 1];
    const uint32_t vox = first_rank[vals[t0]];
    int c = 0;
    for (uint32_t t = t0; t < t1 && c < cap; t += 32) {
      const bool have = t + lane < t1;
      uint32_t i = 0;
      double px = 0, py = 0, pz = 0;
      if (have) {
        i = vals[t + lane];
        const float* f = (const float*)(raw + (size_t)i * stride);
        px = (double)f[0], py = (double)f[1], pz = (double)f[2];
      }
      bool alive = have;
      for (int j = 0; j < c; ++j)  // sub3 / sqnorm3 order of operations, as the one-by-one rule had it
        alive = alive && !(run_sqdist<false>(ks.x[j], ks.y[j], ks.z[j], px, py, pz) < min_sq);
      const int c0 = c;
      const unsigned acc = resolve_chunk<false>(ks, c, cap, min_sq, alive, px, py, pz, lane);
      if ((acc >> lane) & 1u) kept[(size_t)vox * cap + c0 + __popc(acc & ((1u << lane) - 1u))] = i;
      __syncwarp();
    }
    if (lane == 0) kept_count[vox] = (uint32_t)c;
    __syncwarp();
  }
}

__global__ void k_ds_scatter(const uint32_t* __restrict__ kept, const uint32_t* __restrict__ kept_count,
                             const uint32_t* __restrict__ offset, uint32_t n_vox, int cap, uint32_t* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t v = t / cap, j = t % cap;
  if (v >= n_vox || j >= kept_count[v]) return;
  out[offset[v] + j] = kept[(size_t)v * cap + j];
}

constexpr int kKnnThreads = 128;

// Standalone restricted k-NN: one query per thread (see knn_thread in mb_internal.cuh).
template <int K>
__global__ void __launch_bounds__(kKnnThreads, 7)
    k_knn(MapView mv, const double* __restrict__ q, size_t nq, int k, uint64_t* __restrict__ idx,
          double* __restrict__ d2, uint8_t* __restrict__ ok) {
  __shared__ uint16_t s_tab[kTabEntries];
  __shared__ uint32_t s_pk_all[kMaxNbr * kKnnThreads];
  __shared__ uint32_t s_blk_all[24 * kKnnThreads];
  fill_scan_table(mv, s_tab);
  __syncthreads();
  uint32_t* s_pk = s_pk_all + threadIdx.x;
  uint32_t* s_blk = s_blk_all + threadIdx.x;
  const size_t i = (size_t)blockIdx.x * kKnnThreads + threadIdx.x;
  const bool active = i < nq;
  double qx = 0, qy = 0, qz = 0;
  if (active) {
    qx = q[3 * i];
    qy = q[3 * i + 1];
    qz = q[3 * i + 2];
  }
  double bd[K];
  uint32_t bs[K];
  knn_thread<K>(mv, s_tab, s_pk, s_blk, kKnnThreads, qx, qy, qz, k, active, bd, bs);
  if (!active) return;
  uint64_t g[K];
  float4 pts_unused[K];
  const int found = knn_resolve_all<K, false>(mv, s_pk, kKnnThreads, bs, k, g, pts_unused);
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (j < k) {
      idx[i * k + j] = g[j];
      d2[i * k + j] = g[j] != ~0ull ? bd[j] : DBL_MAX;
    }
  }
  ok[i] = found == k;
#if defined(MB_KNN_TIMING)
  MB_KNN_T(6);
  if (blockIdx.x == MB_KNN_TIMING && threadIdx.x == 0)
    printf("knn timing block %d: probes %lld  existence %lld  own %lld  phase4 %lld  loop %lld  resolve+store %lld  (cycles)\n",
           (int)blockIdx.x, g_knn_t[1] - g_knn_t[0], g_knn_t[2] - g_knn_t[1], g_knn_t[3] - g_knn_t[2], g_knn_t[4] - g_knn_t[3],
           g_knn_t[5] - g_knn_t[4], g_knn_t[6] - g_knn_t[5]);
#endif
}

__global__ void k_gather_points(const float4* __restrict__ pts, int cap, const uint64_t* __restrict__ idx, size_t n,
                                double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t g = idx[i];
  const float4 p = pts[(size_t)(g >> 32) * cap + (size_t)(g & 0xffffffffull)];
  out[3 * i] = p.x;
  out[3 * i + 1] = p.y;
  out[3 * i + 2] = p.z;
}

__global__ void k_sum_counts(const int32_t* __restrict__ count, uint32_t n_vox, unsigned long long* __restrict__ out) {
  unsigned long long local = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vox; i += gridDim.x * blockDim.x) local += count[i];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(kFull, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// Bump allocator over the map's scratch buffer.
struct Bump {
  char* base;
  size_t off = 0, cap;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = (T*)(base + off);
    off += n * sizeof(T);
    return p;
  }
};

int ensure_scratch(mb_map* m, size_t bytes) {
  if (bytes <= m->scratch_bytes) return MB_OK;
  MB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  if (m->scratch) dev_free_t(m->ctx, m->scratch);
  m->scratch = nullptr;
  m->scratch_bytes = 0;
  MB_TRY(dev_alloc_t(m->ctx, &m->scratch, bytes));
  m->scratch_bytes = bytes;
  return MB_OK;
}

int rebuild_table(mb_map* m, size_t want_table) {
  cudaStream_t st = m->ctx->stream;
  if (want_table != m->table_cap) {
    MB_CUDA(cudaStreamSynchronize(st));
    if (m->table) dev_free_t(m->ctx, m->table);
    m->table = nullptr;
    MB_TRY(dev_alloc_t(m->ctx, &m->table, want_table * sizeof(int4)));
    m->table_cap = want_table;
  }
  MB_CUDA(cudaMemsetAsync(m->table, 0xff, m->table_cap * sizeof(int4), st));
  if (m->n_vox) {
    k_rehash<<<blocks_for(m->n_vox, 256), 256, 0, st>>>(m->info, m->count, (uint32_t)m->n_vox, m->table,
                                                        (uint32_t)(m->table_cap - 1), m->epos);
    ++m->ctx->launches;
    MB_CUDA(cudaGetLastError());
  }
  return MB_OK;
}

}  // namespace

int map_reserve(mb_map* m, size_t want_vox) {
  cudaStream_t st = m->ctx->stream;
  if (want_vox >= (1u << (32 - kCountBits)) - 1) {
    set_error("map_reserve: %zu voxels exceed the %d-bit voxel id space", want_vox, 32 - kCountBits);
    return MB_ERR_CAPACITY;
  }
  if (want_vox > m->cap_vox) {
    // Grow geometrically while the map is small; round to 64 Ki voxels so that successive snapshots of a
    // slowly growing map ask the block pool for identical sizes.
    size_t new_cap = std::max<size_t>(want_vox, m->cap_vox + m->cap_vox / 2);
    new_cap = std::max<size_t>(new_cap, 4096);
    if (new_cap > 65536) new_cap = (new_cap + 65535) & ~(size_t)65535;
    float4* pts2 = nullptr;
    int4* info2 = nullptr;
    int32_t* count2 = nullptr;
    uint32_t* epos2 = nullptr;
    MB_TRY(dev_alloc_t(m->ctx, &pts2, new_cap * m->cap * sizeof(float4)));
    MB_TRY(dev_alloc_t(m->ctx, &info2, new_cap * sizeof(int4)));
    MB_TRY(dev_alloc_t(m->ctx, &count2, new_cap * sizeof(int32_t)));
    MB_TRY(dev_alloc_t(m->ctx, &epos2, new_cap * sizeof(uint32_t)));
    if (m->n_vox) {
      MB_CUDA(cudaMemcpyAsync(pts2, m->pts, m->n_vox * m->cap * sizeof(float4), cudaMemcpyDeviceToDevice, st));
      MB_CUDA(cudaMemcpyAsync(info2, m->info, m->n_vox * sizeof(int4), cudaMemcpyDeviceToDevice, st));
      MB_CUDA(cudaMemcpyAsync(count2, m->count, m->n_vox * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
      MB_CUDA(cudaMemcpyAsync(epos2, m->epos, m->n_vox * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    MB_CUDA(cudaStreamSynchronize(st));
    if (m->pts) dev_free_t(m->ctx, m->pts);
    if (m->info) dev_free_t(m->ctx, m->info);
    if (m->count) dev_free_t(m->ctx, m->count);
    if (m->epos) dev_free_t(m->ctx, m->epos);
    m->pts = pts2;
    m->info = info2;
    m->count = count2;
    m->epos = epos2;
    m->cap_vox = new_cap;
  }
  size_t want_table = 1024;
  while (want_table < 2 * m->cap_vox) want_table <<= 1;
  if (want_table > m->table_cap) MB_TRY(rebuild_table(m, want_table));
  return MB_OK;
}

int ensure_mirror(mb_map* m) {
  if (m->r_fresh) return MB_OK;
  mb_ctx* c = m->ctx;
  cudaStream_t st = c->stream;
  const size_t nv = m->n_vox;
  if (m->r_cap_vox < std::max<size_t>(nv, 1)) {
    MB_CUDA(cudaStreamSynchronize(st));
    dev_free_t(m->ctx, m->r_pts);
    dev_free_t(m->ctx, m->r_meta);
    m->r_pts = nullptr;
    m->r_meta = nullptr;
    m->r_cap_vox = 0;
    size_t cap_vox = std::max<size_t>(nv + nv / 8, 4096);
    if (cap_vox > 65536) cap_vox = (cap_vox + 65535) & ~(size_t)65535;
    MB_TRY(dev_alloc_t(m->ctx, &m->r_pts, cap_vox * m->cap * sizeof(float4)));
    MB_TRY(dev_alloc_t(m->ctx, &m->r_meta, cap_vox * sizeof(uint32_t)));
    m->r_cap_vox = cap_vox;
  }
  size_t want_b = 1024;
  while (want_b < 2 * std::max<size_t>(nv, 1)) want_b <<= 1;  // >= 2 x blocks for any block count <= n_vox ...
  // ... but blocks are usually ~10x fewer than voxels: size from the real count below once it is known
  size_t sort_temp = 0, scan_temp = 0;
  if (nv) {
    cub::DeviceRadixSort::SortPairs(nullptr, sort_temp, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)nv, 0, 63, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)nv, st);
  }
  const size_t temp_bytes = std::max(sort_temp, scan_temp);
  MB_TRY(ensure_scratch(m, nv * (16 + 8 + 8 + 16 + 8) + temp_bytes + 256 * 12 + 1024));
  uint32_t n_blocks = 0;
  int4* blk_hdr = nullptr;
  unsigned long long* blk_mask = nullptr;
  if (nv) {
    Bump b{(char*)m->scratch, 0, m->scratch_bytes};
    uint64_t* keys = b.take<uint64_t>(nv);
    uint64_t* keys_s = b.take<uint64_t>(nv);
    uint32_t* ids = b.take<uint32_t>(nv);
    uint32_t* ids_s = b.take<uint32_t>(nv);
    uint32_t* head = b.take<uint32_t>(nv);
    uint32_t* block_of = b.take<uint32_t>(nv);
    blk_hdr = b.take<int4>(nv);
    blk_mask = b.take<unsigned long long>(nv);
    uint32_t* counters = b.take<uint32_t>(4);
    void* temp = b.take<unsigned char>(temp_bytes);
    k_mirror_keys<<<blocks_for(nv, 256), 256, 0, st>>>(m->info, (uint32_t)nv, keys, ids);
    size_t tb = temp_bytes;
    MB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys_s, ids, ids_s, (int)nv, 0, 63, st));
    const unsigned grid = (unsigned)std::min<size_t>((nv + 7) / 8, (size_t)c->sm_count * 32);
    k_mirror_gather<<<grid, 256, 0, st>>>(keys_s, ids_s, (uint32_t)nv, m->cap, m->pts, m->count, m->r_pts, m->r_meta, head);
    tb = temp_bytes;
    MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, head, block_of, (int)nv, st));
    k_count2<<<1, 1, 0, st>>>(head, block_of, nullptr, nullptr, nv, counters);
    MB_CUDA(cudaMemsetAsync(blk_mask, 0, nv * sizeof(unsigned long long), st));
    k_mirror_blocks<<<blocks_for(nv, 256), 256, 0, st>>>(keys_s, ids_s, m->info, head, block_of, (uint32_t)nv, blk_hdr, blk_mask);
    c->launches += 4 + 6;
    MB_CUDA(cudaMemcpyAsync(&n_blocks, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
  }
  want_b = 1024;
  while (want_b < 2 * (size_t)n_blocks) want_b <<= 1;
  if (want_b > m->r_bcap) {
    MB_CUDA(cudaStreamSynchronize(st));
    dev_free_t(m->ctx, m->r_btab);
    m->r_btab = nullptr;
    MB_TRY(dev_alloc_t(m->ctx, &m->r_btab, 2 * want_b * sizeof(int4)));
    m->r_bcap = want_b;
  }
  MB_CUDA(cudaMemsetAsync(m->r_btab, 0xff, 2 * m->r_bcap * sizeof(int4), st));
  if (n_blocks) {
    k_mirror_table<<<blocks_for(n_blocks, 256), 256, 0, st>>>(blk_hdr, blk_mask, n_blocks, m->r_btab, (uint32_t)(m->r_bcap - 1));
    ++c->launches;
  }
  MB_CUDA(cudaGetLastError());
  m->r_nblocks = n_blocks;
  m->r_fresh = true;
  return MB_OK;
}

int launch_knn(mb_map* m, const double* d_q, size_t nq, int k, uint64_t* d_idx, double* d_d2, uint8_t* d_ok) {
  if (nq == 0) return MB_OK;
  MB_TRY(ensure_mirror(m));
  cudaStream_t st = m->ctx->stream;
  const unsigned grid = blocks_for(nq, kKnnThreads);
  if (k == 5)
    k_knn<5><<<grid, kKnnThreads, 0, st>>>(m->view(), d_q, nq, k, d_idx, d_d2, d_ok);
  else
    k_knn<MB_MAX_K><<<grid, kKnnThreads, 0, st>>>(m->view(), d_q, nq, k, d_idx, d_d2, d_ok);
  ++m->ctx->launches;
  MB_CUDA(cudaGetLastError());
  return MB_OK;
}

}  // namespace mb

using namespace mb;

// ---- C ABI ------------------------------------------------------------------------------------------
extern "C" {

int mb_map_create(mb_ctx* ctx, float leaf, float min_dist, int cap, int nbr_mode, uint64_t lru_horizon,
                  mb_map** out) {
  MB_REQUIRE(ctx && out, "null argument");
  MB_REQUIRE(leaf > 0.f && min_dist >= 0.f, "leaf must be > 0 and min_dist >= 0");
  MB_REQUIRE(nbr_mode == 1 || nbr_mode == 7 || nbr_mode == 19 || nbr_mode == 27,
             "neighbor_voxel_mode must be 1, 7, 19 or 27 (geometric_config.cpp:84-89)");
  if (cap < 1 || cap >= (1 << kCountBits)) {
    set_error("mb_map_create: cap %d outside [1, %d]", cap, (1 << kCountBits) - 1);
    return MB_ERR_UNSUPPORTED;
  }
  MB_CUDA(cudaSetDevice(ctx->device));
  mb_map* m = new mb_map;
  m->ctx = ctx;
  m->leaf = (double)leaf;
  m->inv_leaf = 1.0 / (double)leaf;
  m->min_sq_dist = (double)min_dist * (double)min_dist;
  m->cap = cap;
  m->nbr_mode = nbr_mode;
  m->lru_horizon = lru_horizon;
  const int n = neighbor_offsets(nbr_mode, m->off);
  m->n_off = n;
  if (const char* e = getenv("MB_KNN_PREF")) m->pref_frac = atof(e);  // development knob
  int s = map_reserve(m, 4096);
  if (s != MB_OK) {
    delete m;
    return s;
  }
  *out = m;
  return MB_OK;
}

int mb_map_release(mb_map* m) {
  if (!m) return MB_OK;
  if (m->refs.fetch_sub(1) > 1) return MB_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  dev_free_t(m->ctx, m->pts);
  dev_free_t(m->ctx, m->info);
  dev_free_t(m->ctx, m->count);
  dev_free_t(m->ctx, m->epos);
  dev_free_t(m->ctx, m->table);
  dev_free_t(m->ctx, m->scratch);
  dev_free_t(m->ctx, m->r_pts);
  dev_free_t(m->ctx, m->r_meta);
  dev_free_t(m->ctx, m->r_btab);
  dev_free_t(m->ctx, m->q_dev);
  dev_free_t(m->ctx, m->q_idx);
  dev_free_t(m->ctx, m->q_d2);
  dev_free_t(m->ctx, m->q_ok);
  delete m;
  return MB_OK;
}

int mb_map_snapshot(mb_map* m, mb_map** out) {
  MB_REQUIRE(m && out, "null argument");
  MB_CUDA(cudaSetDevice(m->ctx->device));
  mb_map* s = new mb_map;
  s->ctx = m->ctx;
  s->leaf = m->leaf;
  s->inv_leaf = m->inv_leaf;
  s->min_sq_dist = m->min_sq_dist;
  s->cap = m->cap;
  s->nbr_mode = m->nbr_mode;
  s->n_off = m->n_off;
  std::memcpy(s->off, m->off, sizeof(m->off));
  s->pref_frac = m->pref_frac;
  s->lru_horizon = m->lru_horizon;
  s->lru_counter = m->lru_counter;
  // headroom so that the insert which usually follows a snapshot (geometric.cpp:494-495) does not regrow it
  int st = map_reserve(s, std::max<size_t>(m->n_vox + std::max<size_t>(m->n_vox / 32, 65536), 4096));
  if (st != MB_OK) {
    mb_map_release(s);
    return st;
  }
  cudaStream_t stream = m->ctx->stream;
  if (m->n_vox) {
    MB_CUDA(cudaMemcpyAsync(s->pts, m->pts, m->n_vox * m->cap * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
    MB_CUDA(cudaMemcpyAsync(s->info, m->info, m->n_vox * sizeof(int4), cudaMemcpyDeviceToDevice, stream));
    MB_CUDA(cudaMemcpyAsync(s->count, m->count, m->n_vox * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
  }
  s->n_vox = m->n_vox;
  st = rebuild_table(s, s->table_cap);
  if (st != MB_OK) {
    mb_map_release(s);
    return st;
  }
  *out = s;
  return MB_OK;
}

int mb_map_size(mb_map* m, size_t* n_voxels, size_t* n_points, uint64_t* lru_counter) {
  MB_REQUIRE(m, "null map");
  MB_CUDA(cudaSetDevice(m->ctx->device));
  if (n_voxels) *n_voxels = m->n_vox;
  if (lru_counter) *lru_counter = m->lru_counter;
  if (n_points) {
    *n_points = 0;
    if (m->n_vox) {
      cudaStream_t st = m->ctx->stream;
      MB_TRY(ensure_scratch(m, 1024));
      unsigned long long* d_sum = (unsigned long long*)m->scratch;
      MB_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(unsigned long long), st));
      k_sum_counts<<<std::min<unsigned>(blocks_for(m->n_vox, 256), 1024u), 256, 0, st>>>(m->count, (uint32_t)m->n_vox,
                                                                                         d_sum);
      ++m->ctx->launches;
      unsigned long long h = 0;
      MB_CUDA(cudaMemcpyAsync(&h, d_sum, sizeof(h), cudaMemcpyDeviceToHost, st));
      MB_CUDA(cudaStreamSynchronize(st));
      *n_points = (size_t)h;
    }
  }
  return MB_OK;
}

int mb_map_insert(mb_map* m, const float* xyz, size_t n, size_t stride_bytes) {
  return mb::map_insert_impl(m, xyz, cudaMemcpyHostToDevice, n, stride_bytes);
}

}  // extern "C"

int mb::map_insert_impl(mb_map* m, const void* xyz, cudaMemcpyKind kind, size_t n, size_t stride_bytes) {
  MB_REQUIRE(m, "null map");
  MB_REQUIRE(n == 0 || xyz, "null points");
  MB_REQUIRE(stride_bytes >= 12 && stride_bytes % 4 == 0, "stride must be >= 12 and a multiple of 4");
  MB_REQUIRE(n < 0xffffffffull, "too many points in one insert");
  MB_REQUIRE(m->refs.load() == 1, "map is referenced by a factor: insert into a snapshot (geometric.cpp:494)");
  MB_CUDA(cudaSetDevice(m->ctx->device));
  m->r_fresh = false;
  cudaStream_t st = m->ctx->stream;
  mb_ctx* ctx = m->ctx;
  if (n > 0) {
    size_t sort_temp = 0, scan_temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_temp, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 63, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, st);
    const size_t temp_bytes = std::max(sort_temp, scan_temp);
    const size_t need = n * stride_bytes + n * (16 + 8 + 4 * 7) + temp_bytes + 256 * 16 + 1024;
    MB_TRY(ensure_scratch(m, need));
    Bump b{(char*)m->scratch, 0, m->scratch_bytes};
    unsigned char* raw = b.take<unsigned char>(n * stride_bytes);
    uint64_t* keys = b.take<uint64_t>(n);
    uint64_t* keys_s = b.take<uint64_t>(n);
    uint32_t* vals = b.take<uint32_t>(n);
    uint32_t* vals_s = b.take<uint32_t>(n);
    uint32_t* head = b.take<uint32_t>(n);
    uint32_t* run_id = b.take<uint32_t>(n);
    uint32_t* run_start = b.take<uint32_t>(n + 1);
    uint32_t* run_packed = b.take<uint32_t>(n);
    uint32_t* run_slot = b.take<uint32_t>(n);
    uint32_t* new_flag = b.take<uint32_t>(n);
    uint32_t* new_rank = b.take<uint32_t>(n);
    uint32_t* counters = b.take<uint32_t>(4);
    int* err = b.take<int>(1);
    void* temp = b.take<unsigned char>(temp_bytes);

    MB_CUDA(cudaMemcpyAsync(raw, xyz, n * stride_bytes, kind, st));
    MB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    MB_CUDA(cudaMemsetAsync(new_flag, 0, n * sizeof(uint32_t), st));
    k_make_keys<<<blocks_for(n, 256), 256, 0, st>>>(raw, n, stride_bytes, m->inv_leaf, keys, vals, err);
    size_t tb = temp_bytes;
    MB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys_s, vals, vals_s, (int)n, 0, 63, st));
    k_mark_heads<<<blocks_for(n, 256), 256, 0, st>>>(keys_s, n, head);
    tb = temp_bytes;
    MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, head, run_id, (int)n, st));
    k_run_starts<<<blocks_for(n, 256), 256, 0, st>>>(head, run_id, n, run_start);
    k_count2<<<1, 1, 0, st>>>(head, run_id, nullptr, nullptr, n, counters);
    ctx->launches += 4 + 6;  // own kernels + (approximate) CUB passes
    uint32_t h_counters[2] = {0, 0};
    int h_err = 0;
    MB_CUDA(cudaMemcpyAsync(h_counters, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (h_err) {
      set_error("mb_map_insert: non-finite point or voxel coordinate outside +-2^20");
      return MB_ERR_UNSUPPORTED;
    }
    const uint32_t n_runs = h_counters[0];
    k_lookup_runs<<<blocks_for(n_runs, 256), 256, 0, st>>>(keys_s, vals_s, run_start, n_runs, m->table,
                                                           (uint32_t)(m->table_cap - 1), run_packed, new_flag);
    tb = temp_bytes;
    MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, new_flag, new_rank, (int)n, st));
    k_count2<<<1, 1, 0, st>>>(nullptr, nullptr, new_flag, new_rank, n, counters);
    ctx->launches += 2 + 2;
    MB_CUDA(cudaMemcpyAsync(h_counters, counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    const uint32_t n_new = h_counters[1];
    MB_TRY(map_reserve(m, m->n_vox + n_new));  // may rebuild the table; existing ids are unchanged
    k_create_voxels<<<blocks_for(n_runs, 256), 256, 0, st>>>(keys_s, vals_s, run_start, n_runs, run_packed, new_rank,
                                                             (uint32_t)m->n_vox, (int)m->lru_counter, m->table,
                                                             (uint32_t)(m->table_cap - 1), m->info, m->count, m->epos,
                                                             run_slot);
    m->n_vox += n_new;
    const unsigned grid = (unsigned)std::min<size_t>(((size_t)n_runs + 7) / 8, (size_t)ctx->sm_count * 16);
    k_insert_runs<<<grid, 256, 0, st>>>(raw, stride_bytes, vals_s, run_start, n_runs, run_slot, m->cap,
                                        m->min_sq_dist, (int)m->lru_counter, m->pts, m->info, m->count, m->epos,
                                        m->table);
    ctx->launches += 2;
    MB_CUDA(cudaGetLastError());
  }
  // LRU bookkeeping (every lru_clear_cycle = 10 inserts): drop voxels with lru + horizon < counter.
  ++m->lru_counter;
  if (m->lru_counter % 10 == 0 && m->n_vox > 0) {
    size_t scan_temp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)m->n_vox, st);
    // scratch is free again here (the insert kernels above are ordered before us on the stream), but
    // ensure_scratch may reallocate, which synchronises first.
    MB_TRY(ensure_scratch(m, m->n_vox * 8 + scan_temp + 4096));
    Bump b{(char*)m->scratch, 0, m->scratch_bytes};
    uint32_t* keep = b.take<uint32_t>(m->n_vox);
    uint32_t* new_id = b.take<uint32_t>(m->n_vox);
    uint32_t* counters = b.take<uint32_t>(4);
    void* temp = b.take<unsigned char>(scan_temp);
    k_flag_keep<<<blocks_for(m->n_vox, 256), 256, 0, st>>>(m->info, (uint32_t)m->n_vox, (long long)m->lru_horizon,
                                                           (long long)m->lru_counter, keep);
    MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, scan_temp, keep, new_id, (int)m->n_vox, st));
    k_count2<<<1, 1, 0, st>>>(keep, new_id, nullptr, nullptr, m->n_vox, counters);
    ctx->launches += 4;
    uint32_t n_keep = 0;
    MB_CUDA(cudaMemcpyAsync(&n_keep, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    if (n_keep < m->n_vox) {
      float4* pts2 = nullptr;
      int4* info2 = nullptr;
      int32_t* count2 = nullptr;
      MB_TRY(dev_alloc_t(m->ctx, &pts2, m->cap_vox * m->cap * sizeof(float4)));
      MB_TRY(dev_alloc_t(m->ctx, &info2, m->cap_vox * sizeof(int4)));
      MB_TRY(dev_alloc_t(m->ctx, &count2, m->cap_vox * sizeof(int32_t)));
      const unsigned grid = (unsigned)std::min<size_t>((m->n_vox + 7) / 8, (size_t)ctx->sm_count * 16);
      k_compact<<<grid, 256, 0, st>>>(keep, new_id, (uint32_t)m->n_vox, m->cap, m->pts, m->info, m->count, pts2, info2,
                                      count2);
      ++ctx->launches;
      MB_CUDA(cudaStreamSynchronize(st));
      dev_free_t(m->ctx, m->pts);
      dev_free_t(m->ctx, m->info);
      dev_free_t(m->ctx, m->count);
      m->pts = pts2;
      m->info = info2;
      m->count = count2;
      m->n_vox = n_keep;
      MB_TRY(rebuild_table(m, m->table_cap));
    }
  }
  MB_CUDA(cudaStreamSynchronize(st));
  return MB_OK;
}

extern "C" {

int mb_map_download(mb_map* m, int32_t* coords, int32_t* counts, uint32_t* lru, float* pts) {
  MB_REQUIRE(m, "null map");
  MB_CUDA(cudaSetDevice(m->ctx->device));
  cudaStream_t st = m->ctx->stream;
  const size_t nv = m->n_vox;
  if (nv == 0) return MB_OK;
  std::vector<int4> info(nv);
  MB_CUDA(cudaMemcpyAsync(info.data(), m->info, nv * sizeof(int4), cudaMemcpyDeviceToHost, st));
  if (counts) MB_CUDA(cudaMemcpyAsync(counts, m->count, nv * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  std::vector<float4> hp;
  if (pts) {
    hp.resize(nv * m->cap);
    MB_CUDA(cudaMemcpyAsync(hp.data(), m->pts, nv * m->cap * sizeof(float4), cudaMemcpyDeviceToHost, st));
  }
  std::vector<int32_t> hc;
  if (pts && !counts) {
    hc.resize(nv);
    MB_CUDA(cudaMemcpyAsync(hc.data(), m->count, nv * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  }
  MB_CUDA(cudaStreamSynchronize(st));
  const int32_t* cnt = counts ? counts : hc.data();
  for (size_t v = 0; v < nv; ++v) {
    if (coords) {
      coords[3 * v] = info[v].x;
      coords[3 * v + 1] = info[v].y;
      coords[3 * v + 2] = info[v].z;
    }
    if (lru) lru[v] = (uint32_t)info[v].w;
    if (pts)
      for (int j = 0; j < m->cap; ++j) {
        float* f = pts + (v * m->cap + j) * 3;
        if (j < cnt[v]) {
          const float4 p = hp[v * m->cap + j];
          f[0] = p.x;
          f[1] = p.y;
          f[2] = p.z;
        } else {
          f[0] = f[1] = f[2] = 0.f;
        }
      }
  }
  return MB_OK;
}

int mb_map_upload(mb_map* m, const int32_t* coords, const int32_t* counts, const uint32_t* lru, const float* pts,
                  size_t n_vox, uint64_t lru_counter) {
  MB_REQUIRE(m && (n_vox == 0 || (coords && counts && pts)), "null argument");
  MB_REQUIRE(m->refs.load() == 1, "map is referenced by a factor");
  MB_CUDA(cudaSetDevice(m->ctx->device));
  m->r_fresh = false;
  cudaStream_t st = m->ctx->stream;
  m->n_vox = 0;
  MB_TRY(map_reserve(m, std::max<size_t>(n_vox, 4096)));
  std::vector<int4> info(n_vox);
  std::vector<float4> hp(n_vox * m->cap, make_float4(0.f, 0.f, 0.f, 0.f));
  for (size_t v = 0; v < n_vox; ++v) {
    MB_REQUIRE(counts[v] >= 0 && counts[v] <= m->cap, "voxel count outside [0, cap]");
    info[v] = make_int4(coords[3 * v], coords[3 * v + 1], coords[3 * v + 2], lru ? (int)lru[v] : 0);
    for (int j = 0; j < counts[v]; ++j) {
      const float* f = pts + (v * m->cap + j) * 3;
      hp[v * m->cap + j] = make_float4(f[0], f[1], f[2], 0.f);
    }
  }
  if (n_vox) {
    MB_CUDA(cudaMemcpyAsync(m->info, info.data(), n_vox * sizeof(int4), cudaMemcpyHostToDevice, st));
    MB_CUDA(cudaMemcpyAsync(m->count, counts, n_vox * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    MB_CUDA(cudaMemcpyAsync(m->pts, hp.data(), n_vox * m->cap * sizeof(float4), cudaMemcpyHostToDevice, st));
  }
  m->n_vox = n_vox;
  m->lru_counter = lru_counter;
  MB_TRY(rebuild_table(m, m->table_cap));
  MB_CUDA(cudaStreamSynchronize(st));
  return MB_OK;
}

static int stage_buffers(mb_map* m, size_t nq, int k) {
  if (nq > m->q_cap || k != m->q_k) {
    MB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    dev_free_t(m->ctx, m->q_dev);
    dev_free_t(m->ctx, m->q_idx);
    dev_free_t(m->ctx, m->q_d2);
    dev_free_t(m->ctx, m->q_ok);
    m->q_dev = nullptr;
    m->q_idx = nullptr;
    m->q_d2 = nullptr;
    m->q_ok = nullptr;
    m->q_cap = 0;
    MB_TRY(dev_alloc_t(m->ctx, &m->q_dev, nq * 3 * sizeof(double)));
    MB_TRY(dev_alloc_t(m->ctx, &m->q_idx, nq * k * sizeof(uint64_t)));
    MB_TRY(dev_alloc_t(m->ctx, &m->q_d2, nq * k * sizeof(double)));
    MB_TRY(dev_alloc_t(m->ctx, &m->q_ok, nq));
    m->q_cap = nq;
    m->q_k = k;
  }
  m->q_n = nq;
  return MB_OK;
}

int mb_map_knn_stage(mb_map* m, const double* q, size_t nq, int k) {
  MB_REQUIRE(m && q && nq > 0, "null argument");
  if (k < 1 || k > MB_MAX_K) {
    set_error("mb_map_knn: k=%d outside [1, %d]", k, MB_MAX_K);
    return MB_ERR_UNSUPPORTED;
  }
  MB_CUDA(cudaSetDevice(m->ctx->device));
  MB_TRY(stage_buffers(m, nq, k));
  MB_CUDA(cudaMemcpyAsync(m->q_dev, q, nq * 3 * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
  MB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return MB_OK;
}

int mb_map_knn_staged_run(mb_map* m) {
  MB_REQUIRE(m && m->q_n > 0, "nothing staged");
  MB_CUDA(cudaSetDevice(m->ctx->device));
  return launch_knn(m, m->q_dev, m->q_n, m->q_k, m->q_idx, m->q_d2, m->q_ok);
}

int mb_map_knn_staged_run_prefix(mb_map* m, size_t nq) {
  MB_REQUIRE(m && m->q_n > 0, "nothing staged");
  MB_CUDA(cudaSetDevice(m->ctx->device));
  return launch_knn(m, m->q_dev, std::min(nq, m->q_n), m->q_k, m->q_idx, m->q_d2, m->q_ok);
}

int mb_map_knn_staged_fetch(mb_map* m, uint64_t* idx, double* d2, uint8_t* ok) {
  MB_REQUIRE(m && m->q_n > 0, "nothing staged");
  MB_CUDA(cudaSetDevice(m->ctx->device));
  cudaStream_t st = m->ctx->stream;
  if (idx) MB_CUDA(cudaMemcpyAsync(idx, m->q_idx, m->q_n * m->q_k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (d2) MB_CUDA(cudaMemcpyAsync(d2, m->q_d2, m->q_n * m->q_k * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (ok) MB_CUDA(cudaMemcpyAsync(ok, m->q_ok, m->q_n, cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  return MB_OK;
}

int mb_map_knn(mb_map* m, const double* q, size_t nq, int k, uint64_t* idx, double* d2, uint8_t* ok) {
  if (nq == 0) return MB_OK;
  MB_TRY(mb_map_knn_stage(m, q, nq, k));
  MB_TRY(mb_map_knn_staged_run(m));
  return mb_map_knn_staged_fetch(m, idx, d2, ok);
}

int mb_map_points(mb_map* m, const uint64_t* idx, size_t n, double* xyz) {
  MB_REQUIRE(m && (n == 0 || (idx && xyz)), "null argument");
  if (n == 0) return MB_OK;
  MB_CUDA(cudaSetDevice(m->ctx->device));
  cudaStream_t st = m->ctx->stream;
  // validate on the host against a copy of the counts (indices come from the caller)
  std::vector<int32_t> cnt(m->n_vox);
  if (m->n_vox) MB_CUDA(cudaMemcpyAsync(cnt.data(), m->count, m->n_vox * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  for (size_t i = 0; i < n; ++i) {
    const uint64_t v = idx[i] >> 32, j = idx[i] & 0xffffffffull;
    MB_REQUIRE(v < m->n_vox && (int64_t)j < cnt[v], "point index out of range");
  }
  MB_TRY(ensure_scratch(m, n * (sizeof(uint64_t) + 3 * sizeof(double)) + 1024));
  Bump b{(char*)m->scratch, 0, m->scratch_bytes};
  uint64_t* d_idx = b.take<uint64_t>(n);
  double* d_out = b.take<double>(3 * n);
  MB_CUDA(cudaMemcpyAsync(d_idx, idx, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  k_gather_points<<<blocks_for(n, 256), 256, 0, st>>>(m->pts, m->cap, d_idx, n, d_out);
  ++m->ctx->launches;
  MB_CUDA(cudaMemcpyAsync(xyz, d_out, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  return MB_OK;
}

// Scratch for mb_downsample lives in the context-free path: allocate per call (scan-sized, a few MB).
int mb_downsample(mb_ctx* ctx, const float* xyz, size_t n, size_t stride_bytes, float leaf, size_t cap,
                  float min_dist, uint32_t* out_idx, size_t* n_out) {
  MB_REQUIRE(n == 0 || out_idx, "null argument");
  return mb::downsample_impl(ctx, xyz, cudaMemcpyHostToDevice, n, stride_bytes, leaf, cap, min_dist, out_idx, nullptr, n_out);
}

}  // extern "C"

int mb::downsample_impl(mb_ctx* ctx, const void* xyz, cudaMemcpyKind kind, size_t n, size_t stride_bytes, float leaf,
                        size_t cap, float min_dist, uint32_t* out_idx, uint32_t* d_out_idx, size_t* n_out) {
  MB_REQUIRE(ctx && n_out && (n == 0 || xyz), "null argument");
  MB_REQUIRE(stride_bytes >= 12 && stride_bytes % 4 == 0, "stride must be >= 12 and a multiple of 4");
  MB_REQUIRE(leaf > 0.f && min_dist >= 0.f, "leaf must be > 0 and min_dist >= 0");
  MB_REQUIRE(n < 0x7fffffffull, "too many points");
  if (cap < 1 || cap > 32) {
    set_error("mb_downsample: cap %zu outside [1, 32]", cap);
    return MB_ERR_UNSUPPORTED;
  }
  *n_out = 0;
  if (n == 0) return MB_OK;
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const double inv_leaf = 1.0 / (double)leaf;
  const double min_sq = (double)min_dist * (double)min_dist;
  size_t sort_temp = 0, scan_temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_temp, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)n, 0, 63, st);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, st);
  const size_t temp_bytes = std::max(sort_temp, scan_temp);
  const size_t bytes = n * stride_bytes + n * (16 + 8 + 4 * 8) + n * cap * 4 + temp_bytes + 256 * 20;
  void* scratch = nullptr;
  MB_TRY(dev_alloc_t(ctx, &scratch, bytes));
  struct Free {
    mb_ctx* c;
    void* p;
    cudaStream_t s;
    ~Free() {
      cudaStreamSynchronize(s);
      dev_free_t(c, p);
    }
  } guard{ctx, scratch, st};
  Bump b{(char*)scratch, 0, bytes};
  unsigned char* raw = b.take<unsigned char>(n * stride_bytes);
  uint64_t* keys = b.take<uint64_t>(n);
  uint64_t* keys_s = b.take<uint64_t>(n);
  uint32_t* vals = b.take<uint32_t>(n);
  uint32_t* vals_s = b.take<uint32_t>(n);
  uint32_t* head = b.take<uint32_t>(n);
  uint32_t* run_id = b.take<uint32_t>(n);
  uint32_t* run_start = b.take<uint32_t>(n + 1);
  uint32_t* first_flag = b.take<uint32_t>(n);
  uint32_t* first_rank = b.take<uint32_t>(n);
  uint32_t* kept_count = b.take<uint32_t>(n);
  uint32_t* offset = b.take<uint32_t>(n);
  uint32_t* out_d = b.take<uint32_t>(n);
  uint32_t* kept = b.take<uint32_t>(n * cap);
  uint32_t* counters = b.take<uint32_t>(4);
  int* err = b.take<int>(1);
  void* temp = b.take<unsigned char>(temp_bytes);

  MB_CUDA(cudaMemcpyAsync(raw, xyz, n * stride_bytes, kind, st));
  MB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  MB_CUDA(cudaMemsetAsync(first_flag, 0, n * sizeof(uint32_t), st));
  k_make_keys<<<blocks_for(n, 256), 256, 0, st>>>(raw, n, stride_bytes, inv_leaf, keys, vals, err);
  size_t tb = temp_bytes;
  MB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys_s, vals, vals_s, (int)n, 0, 63, st));
  k_mark_heads<<<blocks_for(n, 256), 256, 0, st>>>(keys_s, n, head);
  tb = temp_bytes;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, head, run_id, (int)n, st));
  k_run_starts<<<blocks_for(n, 256), 256, 0, st>>>(head, run_id, n, run_start);
  k_count2<<<1, 1, 0, st>>>(head, run_id, nullptr, nullptr, n, counters);
  uint32_t n_runs = 0;
  int h_err = 0;
  MB_CUDA(cudaMemcpyAsync(&n_runs, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(&h_err, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  if (h_err) {
    set_error("mb_downsample: non-finite point or voxel coordinate outside +-2^20");
    return MB_ERR_UNSUPPORTED;
  }
  k_ds_mark<<<blocks_for(n_runs, 256), 256, 0, st>>>(vals_s, run_start, n_runs, first_flag);
  tb = temp_bytes;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, first_flag, first_rank, (int)n, st));
  const unsigned grid = (unsigned)std::min<size_t>(((size_t)n_runs + 7) / 8, (size_t)ctx->sm_count * 16);
  k_ds_runs<<<grid, 256, 0, st>>>(raw, stride_bytes, vals_s, run_start, n_runs, first_rank, (int)cap, min_sq, kept,
                                  kept_count);
  tb = temp_bytes;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, kept_count, offset, (int)n_runs, st));
  k_count2<<<1, 1, 0, st>>>(kept_count, offset, nullptr, nullptr, n_runs, counters);
  k_ds_scatter<<<blocks_for((size_t)n_runs * cap, 256), 256, 0, st>>>(kept, kept_count, offset, n_runs, (int)cap, out_d);
  ctx->launches += 9 + 10;
  uint32_t total = 0;
  MB_CUDA(cudaMemcpyAsync(&total, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  if (out_idx) MB_CUDA(cudaMemcpyAsync(out_idx, out_d, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (d_out_idx) MB_CUDA(cudaMemcpyAsync(d_out_idx, out_d, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  MB_CUDA(cudaStreamSynchronize(st));
  *n_out = total;
  MB_CUDA(cudaGetLastError());
  return MB_OK;
}
