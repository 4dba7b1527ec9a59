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
#include <unordered_map>
#include <vector>

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
  void* pin_small = nullptr;  // 4 KiB page-locked block for poses in / normal equations out
  // Size-keyed cache of released device blocks: a factor is created and destroyed for every scan with the
  // same sizes, and cudaMalloc/cudaFree would otherwise dominate the per-scan host cost.
  struct Block {
    void* p;
    size_t bytes;
  };
  std::vector<Block> pool;
  size_t pool_bytes = 0;
  std::unordered_map<void*, size_t> tracked;  // sizes of blocks handed out by dev_alloc_t
};

namespace mb {
// Grow-only page-locked staging buffer of the context (synchronises the stream when it has to grow).
int pinned_reserve(mb_ctx* c, size_t bytes);
// Pooled device allocations (exact-size reuse).  dev_free never synchronises: the caller guarantees that no
// work touching the block is still pending on the context's stream.
int dev_alloc(mb_ctx* c, void** p, size_t bytes);
void dev_free(mb_ctx* c, void* p, size_t bytes);
void dev_pool_release(mb_ctx* c);
// Same, with the size remembered per pointer (for owners that do not keep it).
template <typename T>
inline int dev_alloc_t(mb_ctx* c, T** p, size_t bytes) {
  void* v = nullptr;
  const int rc = dev_alloc(c, &v, bytes);
  *p = (T*)v;
  if (rc == MB_OK) c->tracked[v] = bytes;
  return rc;
}
inline void dev_free_t(mb_ctx* c, void* p) {
  if (!p) return;
  auto it = c->tracked.find(p);
  if (it == c->tracked.end()) {
    cudaFree(p);
    return;
  }
  const size_t bytes = it->second;
  c->tracked.erase(it);
  dev_free(c, p, bytes);
}
}

namespace mb {

constexpr int kMaxNbr = 27;
constexpr uint32_t kEmpty = 0xffffffffu;
constexpr int kCountBits = 5;  // cap <= 31 points per voxel (reference: 20)

// What the search kernels need of a map; passed by value.  Searches run on the map's read-optimised mirror
// (mb_map.cuh, "search mirror"): voxels grouped into 4x4x4 blocks, blocks hashed, every block entry carrying a
// 64-bit occupancy mask and the index of its first bucket in a Morton-ordered bucket array.
struct MapView {
  const int4* btab;     // block table, 2 x int4 per entry: {bx, by, bz, base} {mask_lo, mask_hi, -, -}; base == kEmpty -> free
  uint32_t bmask;       // entries - 1 (power of two)
  const float4* pts;    // [slot * cap + j], Morton/block order; xyz are the stored (f32-exact) coordinates
  const uint32_t* meta; // [slot] = (voxel id << 5) | count
  int cap;
  int n_off;
  double inv_leaf;
  int8_t off[kMaxNbr * 3];  // neighbour offsets in the reference's visiting order
};
constexpr int kBlockShift = 2;  // 4 x 4 x 4 voxels per block
__host__ __device__ __forceinline__ uint32_t cell_of(int x, int y, int z) {
  return (uint32_t)(x & 3) | ((uint32_t)(y & 3) << 2) | ((uint32_t)(z & 3) << 4);
}

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

constexpr unsigned kFull = 0xffffffffu;
constexpr int kOffBytes = 96 + 32;  // neighbour offsets (3 x 32 int8) + processing order (32 int8)

// Fill the per-block shared copy of the neighbour table: s_off[3 o .. 3 o + 2] = offset o, s_off[96 + p] = the
// offset index scanned at position p (ordered by |dx| + |dy| + |dz|, ties by visiting order).  Call with all
// threads of the block, then __syncthreads().
__device__ __forceinline__ void fill_offset_table(const MapView& mv, int8_t* s_off) {
  if (threadIdx.x < kMaxNbr * 3) s_off[threadIdx.x] = mv.off[threadIdx.x];
  if (threadIdx.x == 0) {
    int p = 0;
    for (int cls = 0; cls <= 3; ++cls)
      for (int o = 0; o < mv.n_off; ++o) {
        const int a = abs((int)mv.off[3 * o]) + abs((int)mv.off[3 * o + 1]) + abs((int)mv.off[3 * o + 2]);
        if (a == cls) s_off[96 + p++] = (int8_t)o;
      }
  }
}

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may start while its predecessor
// on the stream is still running; pdl_wait() blocks until the predecessor has completed and its writes are
// visible, pdl_launch_dependents() lets the successor's blocks be scheduled early.  Both are no-ops for a
// plain launch.  This hides the ~3.6 us launch-to-launch gap between the small dependent kernels of one ICP
// iteration.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr int kSeqShift = 5;  // sequence number = (offset index << 5) | point index  (cap <= 31)

// Index of the (0,0,0) offset in the reference's visiting order for each neighbourhood mode.
__host__ __device__ __forceinline__ int centre_offset_index(int n_off) { return n_off == 19 ? 9 : n_off == 27 ? 13 : 0; }

// Restricted k-NN, ONE QUERY PER THREAD (every lane of the warp must call; `active` = false idles a lane).
//
// The reference (gtsam_points KnnResult::push over the neighbour voxels, restated in oracle/ivox_ref.hpp) scans
// the stored points of the 1/7/19/27 voxels around the query's voxel in a fixed visiting order and keeps the k
// smallest squared distances with a strict-'<' insertion sort, so equal distances resolve to the earlier
// visitor.  Here every thread runs that scan for its own query with these changes, none of which alters the
// result:
//   * candidates carry their visiting sequence number ((offset index << 5) | point index) and the list is
//     ordered by (d2, sequence), which makes the outcome independent of the order voxels are processed in;
//   * neighbour voxels are located through the block grid of the map's search mirror: <= 8 block probes
//     (L2-resident table) give occupancy masks, bucket indices follow by popcount — no per-voxel hash probe;
//   * the query's own voxel is processed first, after which a neighbour voxel is skipped when the squared
//     distance from the query to that voxel's box (shrunk by 1e-6 voxel to stay conservative under rounding)
//     already exceeds the current k-th best — none of its points could enter the list; surviving buckets are
//     prefetched together, candidates are taken four at a time;
//   * all control flow is warp-converged (uniform trip counts, per-lane predicates).
// K is the compile-time list length (5 = the reference's num_corres_points, 8 = generic: the k nearest are the
// first k of the 8 nearest).  s_pk / s_blk are this thread's columns of shared [n_off][pk_stride] /
// [24][pk_stride] arrays; s_pk receives the bucket index of every scanned neighbour and stays valid for
// knn_resolve().
template <int K>
__device__ __forceinline__ void knn_thread(const MapView& mv, const int8_t* __restrict__ s_off, uint32_t* s_pk,
                                           uint32_t* s_blk, int pk_stride, double qx, double qy, double qz, int k,
                                           bool active,
                                           double (&bd)[K], uint32_t (&bs)[K]) {
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);
#pragma unroll
  for (int i = 0; i < K; ++i) {
    bd[i] = kInf;
    bs[i] = 0xffffffffu;
  }
  const int n_off = mv.n_off;
  const double ux = qx * mv.inv_leaf, uy = qy * mv.inv_leaf, uz = qz * mv.inv_leaf;
  const int cx = fast_floor(ux), cy = fast_floor(uy), cz = fast_floor(uz);

  // ---- locate the neighbourhood's blocks -------------------------------------------------------------
  // The 3x3x3 neighbourhood touches at most 2 blocks per axis.  s_blk[combo] (combo = ix | iy << 1 | iz << 2)
  // receives {mask_lo, mask_hi, base} of block (ix ? hi : lo) per axis.  Duplicate combos (hi == lo) are neither
  // probed nor read: slot_of() can only form a combo bit when the two blocks differ.  Both halves of an entry
  // are fetched together, four entries in flight.
  const int lx = (cx - 1) >> kBlockShift, hx = (cx + 1) >> kBlockShift;
  const int ly = (cy - 1) >> kBlockShift, hy = (cy + 1) >> kBlockShift;
  const int lz = (cz - 1) >> kBlockShift, hz = (cz + 1) >> kBlockShift;
  const unsigned dup_bits = (hx == lx ? 1u : 0u) | (hy == ly ? 2u : 0u) | (hz == lz ? 4u : 0u);
#pragma unroll
  for (int c0 = 0; c0 < 8; c0 += 4) {
    uint32_t h[4];
    int4 e[4], m[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int combo = c0 + u;
      h[u] = hash_coord((combo & 1) ? hx : lx, (combo & 2) ? hy : ly, (combo & 4) ? hz : lz) & mv.bmask;
      e[u] = make_int4(0, 0, 0, (int)kEmpty);
      m[u] = make_int4(0, 0, 0, 0);
      if (active && (combo & dup_bits) == 0) {
        e[u] = __ldg(mv.btab + 2 * (size_t)h[u]);
        m[u] = __ldg(mv.btab + 2 * (size_t)h[u] + 1);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int combo = c0 + u;
      if ((combo & dup_bits) == 0) {
        const int bx = (combo & 1) ? hx : lx, by = (combo & 2) ? hy : ly, bz = (combo & 4) ? hz : lz;
        while ((uint32_t)e[u].w != kEmpty && !(e[u].x == bx && e[u].y == by && e[u].z == bz)) {
          h[u] = (h[u] + 1) & mv.bmask;
          e[u] = __ldg(mv.btab + 2 * (size_t)h[u]);
          m[u] = __ldg(mv.btab + 2 * (size_t)h[u] + 1);
        }
        const bool hit = (uint32_t)e[u].w != kEmpty;
        s_blk[(combo * 3) * pk_stride] = hit ? (uint32_t)m[u].x : 0u;
        s_blk[(combo * 3 + 1) * pk_stride] = hit ? (uint32_t)m[u].y : 0u;
        s_blk[(combo * 3 + 2) * pk_stride] = hit ? (uint32_t)e[u].w : 0u;
      }
    }
  }
  // bucket index of neighbour o, or kEmpty when that voxel does not exist
  auto slot_of = [&](int o) -> uint32_t {
    const int x = cx + s_off[3 * o], y = cy + s_off[3 * o + 1], z = cz + s_off[3 * o + 2];
    const int combo = ((x >> kBlockShift) != lx ? 1 : 0) | ((y >> kBlockShift) != ly ? 2 : 0) | ((z >> kBlockShift) != lz ? 4 : 0);
    const uint32_t m_lo = s_blk[(combo * 3) * pk_stride], m_hi = s_blk[(combo * 3 + 1) * pk_stride];
    const unsigned long long mk = ((unsigned long long)m_hi << 32) | m_lo;
    const uint32_t cell = cell_of(x, y, z);
    if (((mk >> cell) & 1ull) == 0) return kEmpty;
    return s_blk[(combo * 3 + 2) * pk_stride] + (uint32_t)__popcll(mk & ((1ull << cell) - 1ull));
  };

  // ---- scan the candidates --------------------------------------------------------------------------
  // All control flow below is warp-converged (uniform trip counts, per-lane predicates).  Candidates are taken
  // four at a time: four independent loads in flight per lane, then four ordered insertion tests.  A bucket's
  // fill count travels in the .w of its first point, so no separate metadata load precedes the first chunk.
  const int centre = centre_offset_index(n_off);
  const int cap = mv.cap;
  const uint32_t kCntMask = (1u << kCountBits) - 1;

  // k-th best so far (the pruning radius); for K == 5 the kernel is only launched with k == 5.
  auto worst_of = [&]() {
    double w = bd[K - 1];
    if (K != 5) {
#pragma unroll
      for (int i = 0; i < K; ++i)
        if (i == k - 1) w = bd[i];
    }
    return w;
  };
  auto offer = [&](double d, uint32_t s) {
    if ((d < bd[K - 1]) | ((d == bd[K - 1]) & (s < bs[K - 1]))) {
      bool lt[K];
#pragma unroll
      for (int i = 0; i < K; ++i) lt[i] = (d < bd[i]) | ((d == bd[i]) & (s < bs[i]));
#pragma unroll
      for (int i = K - 1; i > 0; --i) {
        bd[i] = lt[i - 1] ? bd[i - 1] : (lt[i] ? d : bd[i]);
        bs[i] = lt[i - 1] ? bs[i - 1] : (lt[i] ? s : bs[i]);
      }
      bd[0] = lt[0] ? d : bd[0];
      bs[0] = lt[0] ? s : bs[0];
    }
  };
  // candidates j .. j+3 of `bucket` (those below its count), in order; returns the bucket's count, which is read
  // from the first point when j == 0 (`cnt` is only a lower bound >= 1 then)
  auto offer4 = [&](const float4* bucket, int o, int j, int cnt) -> int {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) p[u] = __ldg(bucket + min(j + u, cap - 1));
    if (j == 0) {
      cnt = (int)((uint32_t)__float_as_int(p[0].w) & kCntMask);
      if (cnt > 8) prefetch_l2(bucket + 8);
      if (cnt > 16) prefetch_l2(bucket + 16);
    }
    double d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) d[u] = sqdist4((double)p[u].x, (double)p[u].y, (double)p[u].z, qx, qy, qz);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (j + u < cnt) offer(d[u], ((uint32_t)o << kSeqShift) | (uint32_t)(j + u));
    return cnt;
  };

  // (1) the query's own voxel
  {
    const uint32_t slot = active ? slot_of(centre) : kEmpty;
    s_pk[centre * pk_stride] = slot;
    const float4* bucket = mv.pts + (size_t)(slot == kEmpty ? 0u : slot) * cap;
    int cnt = 0;
    if (slot != kEmpty) cnt = offer4(bucket, centre, 0, 1);
    const int max_cnt = __reduce_max_sync(kFull, cnt);
    for (int j = 4; j < max_cnt; j += 4)
      if (j < cnt) offer4(bucket, centre, j, cnt);
  }

  // (2) which neighbours can still contribute: bit o set when the voxel exists and the squared distance from
  //     the query to its box does not exceed the current k-th best; their first cache lines are prefetched
  const double fx = ux - (double)cx, fy = uy - (double)cy, fz = uz - (double)cz;  // position inside the voxel
  const double kMargin = 1e-6;
  const double leaf = 1.0 / mv.inv_leaf;
  const double lb_scale = (leaf * leaf) * (1.0 - 1e-9);
  const double glo_x = fmax(0.0, fx - kMargin), ghi_x = fmax(0.0, (1.0 - fx) - kMargin);
  const double glo_y = fmax(0.0, fy - kMargin), ghi_y = fmax(0.0, (1.0 - fy) - kMargin);
  const double glo_z = fmax(0.0, fz - kMargin), ghi_z = fmax(0.0, (1.0 - fz) - kMargin);
  auto box_lb = [&](int o) {
    const int ox = s_off[3 * o], oy = s_off[3 * o + 1], oz = s_off[3 * o + 2];
    const double gx = ox < 0 ? glo_x : (ox > 0 ? ghi_x : 0.0);
    const double gy = oy < 0 ? glo_y : (oy > 0 ? ghi_y : 0.0);
    const double gz = oz < 0 ? glo_z : (oz > 0 ? ghi_z : 0.0);
    return ((gx * gx + gy * gy) + gz * gz) * lb_scale;
  };
  // bit p of `todo` stands for offset s_off[96 + p]: positions are ordered faces, then edges, then corners, so
  // the nearer boxes are scanned first and the radius has tightened by the time the farther ones are re-checked
  uint32_t todo = 0;
  {
    const double worst = worst_of();
    for (int p = 0; p < n_off; ++p) {
      const int o = s_off[96 + p];
      uint32_t slot = kEmpty;
      if (active & (o != centre)) slot = slot_of(o);
      if (slot != kEmpty && !(box_lb(o) > worst)) {
        s_pk[o * pk_stride] = slot;
        prefetch_l2(mv.pts + (size_t)slot * cap);
        todo |= 1u << p;
      }
    }
  }

  // (3) the surviving neighbours, four candidates per lane and iteration; a lane moves to its next voxel
  //     (lowest set bit = nearest class of box) with a handful of predicated instructions, re-checking the bound
  //     against the radius as it stands then
  int o = 0, j = 0, cnt = 0;
  const float4* bucket = mv.pts;
  while (__any_sync(kFull, (todo != 0) | (j < cnt))) {
    if (j >= cnt && todo != 0) {
      o = s_off[96 + __ffs(todo) - 1];
      todo &= todo - 1;
      cnt = box_lb(o) > worst_of() ? 0 : 1;  // real count arrives with the first chunk
      bucket = mv.pts + (size_t)s_pk[o * pk_stride] * cap;
      j = 0;
    }
    if (j < cnt) {
      cnt = offer4(bucket, o, j, cnt);
      j += 4;
    }
  }
}

// Translate a winner's sequence number into the reference's global index and the stored point.
__device__ __forceinline__ uint64_t knn_resolve(const MapView& mv, const uint32_t* s_pk, int pk_stride, uint32_t seq,
                                                float4& p) {
  const uint32_t o = seq >> kSeqShift, j = seq & ((1u << kSeqShift) - 1);
  const uint32_t slot = s_pk[o * pk_stride];
  p = __ldg(mv.pts + (size_t)slot * mv.cap + j);
  const uint32_t id = __ldg(mv.meta + slot) >> kCountBits;
  return ((uint64_t)id << 32) | (uint64_t)j;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif  // __CUDACC__

}  // namespace mb
