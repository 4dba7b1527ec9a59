// Device-resident incremental hashed voxel map (internal header).
//
// Replaces gtsam_points::iVox behind mimosa::lidar::IncrementalVoxelMapPCL
// (mimosa/include/mimosa/lidar/incremental_voxel_map.hpp:22-51).  Layout in HBM, all indexed by the
// voxel id (= creation order, compacted on LRU eviction, exactly the reference's flat-vector index so
// that (voxel_id << 32) | point_id reproduces the reference's global point index):
//   pts    float4[cap_vox * cap]   320-byte buckets (cap = 20), xyz stored as f32 — lossless, the reference
//                                  ingests V3F (mimosa/src/lidar/geometric.cpp:487-492)
//   info   int4  [cap_vox]         {cx, cy, cz, lru}
//   count  int32 [cap_vox]
//   epos   uint32[cap_vox]         position of the voxel's entry in `table`
//   table  int4  [table_cap]       open-addressing hash {cx, cy, cz, (id << 5) | count}, load <= 0.5
// Those arrays are what insert / LRU / snapshot / download work on (the authoritative state).
//
// SEARCH MIRROR (rebuilt lazily after the map changed, before the next search; mb::ensure_mirror):
//   r_pts   float4[n_vox * cap]    the same buckets re-ordered by (Morton code of the voxel's 4x4x4 block, cell
//                                  index inside the block): spatial neighbours are memory neighbours
//   r_meta  uint32[n_vox]          (voxel id << 5) | count of the bucket at that position
//   r_btab  int4[2 * r_bcap]       hashed block table {bx, by, bz, base} {mask_lo, mask_hi, 0, 0}: a 64-bit
//                                  occupancy mask + first bucket index per block, so the 19 neighbour lookups of a
//                                  query become <= 8 (typically 1-4) L2-resident block probes plus popcounts
#pragma once
#include "mb_internal.cuh"

struct mb_map {
  mb_ctx* ctx = nullptr;
  std::atomic<int> refs{1};
  // parameters
  double leaf = 1.0, inv_leaf = 1.0, min_sq_dist = 0.0;
  int cap = 20, nbr_mode = 7, n_off = 7;
  int8_t off[mb::kMaxNbr * 3] = {0};  // neighbour offsets in the reference's visiting order
  // early-prefetch radius of the search, in voxels (MapView::pref2).  Measured neutral on time from 0 to 0.6 and
  // costing DRAM traffic beyond the compulsory bytes (profiles/r1_experiments.md), hence off by default.
  double pref_frac = 0.0;
  uint64_t lru_horizon = 100, lru_counter = 0;
  // storage
  size_t cap_vox = 0, n_vox = 0, table_cap = 0;
  float4* pts = nullptr;
  int4* info = nullptr;
  int32_t* count = nullptr;
  uint32_t* epos = nullptr;
  int4* table = nullptr;
  unsigned long long* d_npts = nullptr;  // device counter of stored points
  // search mirror
  float4* r_pts = nullptr;
  uint32_t* r_meta = nullptr;
  int4* r_btab = nullptr;
  size_t r_cap_vox = 0, r_bcap = 0, r_nblocks = 0;
  bool r_fresh = false;
  // staged k-NN (roofline timing)
  double* q_dev = nullptr;
  size_t q_n = 0, q_cap = 0;
  int q_k = 0;
  uint64_t* q_idx = nullptr;
  double* q_d2 = nullptr;
  uint8_t* q_ok = nullptr;
  // scratch reused across inserts
  void* scratch = nullptr;
  size_t scratch_bytes = 0;

  mb::MapView view() const {  // valid only while r_fresh (mb::ensure_mirror)
    mb::MapView v;
    v.btab = r_btab;
    v.bmask = (uint32_t)(r_bcap - 1);
    v.pts = r_pts;
    v.meta = r_meta;
    v.cap = cap;
    v.n_off = n_off;
    v.inv_leaf = inv_leaf;
    v.pref2 = pref_frac * pref_frac * leaf * leaf;
    mb::fill_view_tables(v, off, n_off);
    return v;
  }
};

namespace mb {
int map_reserve(mb_map* m, size_t want_vox);
// (Re)build the search mirror if the map changed since it was last built.  Enqueued on the context stream.
int ensure_mirror(mb_map* m);
// Insert n records whose first three floats are xyz from host or device memory (`kind`).
int map_insert_impl(mb_map* m, const void* xyz, cudaMemcpyKind kind, size_t n, size_t stride_bytes);
// Geometric::downsample on host or device input; kept indices go to out_idx (host) and/or d_out_idx (device).
int downsample_impl(mb_ctx* ctx, const void* xyz, cudaMemcpyKind kind, size_t n, size_t stride_bytes, float leaf,
                    size_t cap, float min_dist, uint32_t* out_idx, uint32_t* d_out_idx, size_t* n_out);
// Launch the standalone search kernel over device-resident queries (nq x 3 doubles).
int launch_knn(mb_map* m, const double* d_q, size_t nq, int k, uint64_t* d_idx, double* d_d2, uint8_t* d_ok);
}  // namespace mb
