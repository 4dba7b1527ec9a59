// Internal definitions shared by the translation units of libmimosa_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/mimosa_b200.h"
#include "mb_math.cuh"

namespace mb {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define MB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::mb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MB_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)
#define MB_NCCL(expr)                                                                              \
  do {                                                                                             \
    ncclResult_t _e = (expr);                                                                      \
    if (_e != ncclSuccess) {                                                                       \
      ::mb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_e)); \
      return MB_ERR_NCCL;                                                                          \
    }                                                                                              \
  } while (0)
#define MB_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s != MB_OK) return _s; \
  } while (0)
#define MB_REQUIRE(cond, msg)                 \
  do {                                        \
    if (!(cond)) {                            \
      ::mb::set_error("%s: %s", __func__, msg); \
      return MB_ERR_INVALID_ARG;              \
    }                                         \
  } while (0)

}  // namespace mb

// ---- handles ----------------------------------------------------------------------------------------
struct mb_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint64_t launches = 0;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  void* pinned = nullptr;  // page-locked staging for host <-> device copies of scans
  size_t pinned_bytes = 0;
};

namespace mb {
// Grow-only page-locked staging buffer of the context (synchronises the stream when it has to grow).
int pinned_reserve(mb_ctx* c, size_t bytes);
}

namespace mb {

constexpr int kMaxNbr = 27;
constexpr uint32_t kEmpty = 0xffffffffu;
constexpr int kCountBits = 5;  // cap <= 31 points per voxel (reference: 20)

// What the search kernels need of a map; passed by value.
struct MapView {
  const int4* table;    // open addressing; {cx, cy, cz, (slot << 5) | count}; w == kEmpty -> free
  uint32_t table_mask;  // capacity - 1 (power of two)
  const float4* pts;    // [slot * cap + j], xyz are the stored (f32-exact) coordinates
  int cap;
  int n_off;
  double inv_leaf;
  int8_t off[kMaxNbr * 3];  // neighbour offsets in the reference's visiting order
};

__host__ __device__ __forceinline__ uint32_t hash_coord(int x, int y, int z) {
  uint32_t h = (uint32_t)x * 73856093u ^ (uint32_t)y * 19349669u ^ (uint32_t)z * 83492791u;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

#if defined(__CUDACC__)
// Returns the packed (slot << 5 | count) word of the voxel at (x,y,z), or kEmpty.
__device__ __forceinline__ uint32_t table_find(const int4* __restrict__ table, uint32_t mask, int x, int y, int z) {
  uint32_t h = hash_coord(x, y, z) & mask;
  while (true) {
    const int4 e = __ldg(table + h);
    if ((uint32_t)e.w == kEmpty) return kEmpty;
    if (e.x == x && e.y == y && e.z == z) return (uint32_t)e.w;
    h = (h + 1) & mask;
  }
}

constexpr int kKnnRounds = 8;                  // candidate slots per lane and chunk
constexpr int kKnnChunk = kKnnRounds * 32;     // 256 candidate slots per chunk
constexpr unsigned kFull = 0xffffffffu;

struct KnnOut {     // lane j (< k) holds the j-th nearest neighbour
  double d2;        // +inf when fewer than j+1 candidates exist
  uint32_t seq;     // ordinal * cap + j  (ordinal = rank of the voxel among the found ones); ~0 if none
  int found;        // number of neighbours found (same value in every lane)
};

// Warp-cooperative restricted k-NN for ONE query (all 32 lanes must call, converged).
//   1. lane l < n_off probes neighbour voxel l of the query's voxel in the hash table;
//   2. found voxels are compacted (visiting order kept) into s_vox[] = packed slot/count words;
//   3. candidate slot c = ordinal * cap + j is evaluated by lane c % 32 in round c / 32: one coalesced
//      float4 load per lane and round, fp64 distance, kept in registers;
//   4. the k smallest under the total order (d2, visiting sequence) are extracted with k rounds of three
//      32-bit warp min-reductions (high word, low word, sequence) — equal distances therefore resolve to
//      the earlier visitor exactly like the reference's strict-'<' insertion sort
//      (gtsam_points KnnResult::push; restated in oracle/ivox_ref.hpp).
// s_vox: per-warp shared scratch of 32 words; still valid (for knn_fetch) until the next call.
__device__ __forceinline__ void knn_warp(const MapView& mv, const int8_t* __restrict__ s_off, uint32_t* s_vox,
                                         double qx, double qy, double qz, int k, int lane, KnnOut& out) {
  const int cx = fast_floor(qx * mv.inv_leaf), cy = fast_floor(qy * mv.inv_leaf), cz = fast_floor(qz * mv.inv_leaf);
  uint32_t packed = kEmpty;
  if (lane < mv.n_off)
    packed = table_find(mv.table, mv.table_mask, cx + s_off[3 * lane], cy + s_off[3 * lane + 1], cz + s_off[3 * lane + 2]);
  const bool hit = packed != kEmpty && (packed & ((1u << kCountBits) - 1)) != 0;
  const unsigned hits = __ballot_sync(kFull, hit);
  const int n_found_vox = __popc(hits);
  __syncwarp();
  if (hit) s_vox[__popc(hits & ((1u << lane) - 1))] = packed;
  __syncwarp();

  const int cap = mv.cap;
  const int n_slots = n_found_vox * cap;
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);

  // best-so-far: lane j (< k) holds the j-th best; re-offered as the "carry" candidate when a further
  // chunk of candidates is merged in.
  double best_d2 = kInf;
  uint32_t best_seq = 0xffffffffu;

  for (int base = 0; base < n_slots; base += kKnnChunk) {
    double cd[kKnnRounds];
#pragma unroll
    for (int r = 0; r < kKnnRounds; ++r) {
      cd[r] = kInf;
      const int c = base + r * 32 + lane;
      if (base + r * 32 < n_slots) {  // warp-uniform
        if (c < n_slots) {
          const int ord = c / cap;
          const int j = c - ord * cap;
          const uint32_t pk = s_vox[ord];
          if (j < (int)(pk & ((1u << kCountBits) - 1))) {
            const float4 p = __ldg(mv.pts + (size_t)(pk >> kCountBits) * cap + j);
            cd[r] = sqdist4((double)p.x, (double)p.y, (double)p.z, qx, qy, qz);
          }
        }
      }
    }
    double carry_d2 = best_d2;
    uint32_t carry_seq = best_seq;
    best_d2 = kInf;
    best_seq = 0xffffffffu;
    for (int sel = 0; sel < k; ++sel) {
      // lane-local minimum: strict '<' scanning in round order keeps the earliest sequence on ties; the
      // carry comes from an earlier chunk (smaller sequence), so it is the initial value and wins ties.
      double md = carry_d2;
      uint32_t ms = carry_seq;
#pragma unroll
      for (int r = 0; r < kKnnRounds; ++r) {
        if (cd[r] < md) {
          md = cd[r];
          ms = (uint32_t)(base + r * 32 + lane);
        }
      }
      const uint32_t hi = (uint32_t)__double2hiint(md), lo = (uint32_t)__double2loint(md);
      const uint32_t mhi = __reduce_min_sync(kFull, hi);
      if (mhi >= 0x7ff00000u) break;  // nothing left (warp-uniform)
      const uint32_t mlo = __reduce_min_sync(kFull, hi == mhi ? lo : 0xffffffffu);
      const bool tie = hi == mhi && lo == mlo;
      const uint32_t mseq = __reduce_min_sync(kFull, tie ? ms : 0xffffffffu);
      if (lane == sel) {
        best_d2 = __hiloint2double((int)mhi, (int)mlo);
        best_seq = mseq;
      }
      if (tie && ms == mseq) {  // the winning lane retires that candidate
        if (ms == carry_seq) {
          carry_d2 = kInf;
          carry_seq = 0xffffffffu;
        } else {
#pragma unroll
          for (int r = 0; r < kKnnRounds; ++r)
            if ((uint32_t)(base + r * 32 + lane) == ms) cd[r] = kInf;
        }
      }
    }
  }
  out.d2 = best_d2;
  out.seq = best_seq;
  out.found = __popc(__ballot_sync(kFull, best_seq != 0xffffffffu));
}

// Translate a winner's sequence number into the reference's global index and the stored point.
__device__ __forceinline__ uint64_t knn_fetch(const MapView& mv, const uint32_t* s_vox, uint32_t seq, float4& p) {
  const int ord = seq / mv.cap;
  const int j = seq - ord * mv.cap;
  const uint32_t slot = s_vox[ord] >> kCountBits;
  p = __ldg(mv.pts + (size_t)slot * mv.cap + j);
  return ((uint64_t)slot << 32) | (uint64_t)j;
}
#endif  // __CUDACC__

}  // namespace mb
