// Restricted k-NN over the map's search mirror: the map view, the neighbourhood tables and the per-thread search
// routine shared by k_knn (mb_map.cu) and k_linearize (mb_factor.cu).
//
// This header also compiles as plain C++ (one emulated lane per call; tests/host_shim/shim.cpp supplies the
// handful of intrinsics) so that the CPU test-suite can run the very search code the kernels run against the
// oracle.  That host build is test infrastructure only — nothing in the library uses it.
#pragma once
#include <cstdint>

#include "mb_math.cuh"

#if defined(__CUDACC__)
#define MB_DEV __device__ __forceinline__
#define MB_HDC __host__ __device__
#else
#define MB_DEV inline
#define MB_HDC
#endif

namespace mb {

constexpr int kMaxNbr = 27;
constexpr uint32_t kEmpty = 0xffffffffu;
constexpr int kCountBits = 5;  // cap <= 31 points per voxel (reference: 20)

// The 3x3x3 cube around a query's voxel: cell c = (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1), centre = 13.  Every
// neighbourhood mode of the reference (1 / 7 / 19 / 27 voxels) is a subset of the cube; a mode is described by the
// visiting rank of each cell (MapView::rank).  Non-centre cells also have a SCAN POSITION, 0..25: faces first, then
// edges, then corners (ties by cell index), the order in which surviving neighbours are examined so that the
// nearer boxes tighten the radius before the farther ones are re-checked.
constexpr int kCube = 27, kCentre = 13, kScan = 26;
MB_HDC constexpr int cube_class(int c) { return ((c / 9) != 1) + (((c / 3) % 3) != 1) + ((c % 3) != 1); }
MB_HDC constexpr int cube_pos(int c) {
  int p = 0;
  for (int d = 0; d < kCube; ++d) {
    const int dc = cube_class(d), cc = cube_class(c);
    if (dc != 0 && (dc < cc || (dc == cc && d < c))) ++p;
  }
  return p;
}
// packed scan-table entry of a cube cell: cell | ix << 5 | iy << 7 | iz << 9 | rank << 11 (ix = dx + 1, ...)
MB_HDC constexpr uint16_t scan_entry(int c, int rank) {
  return (uint16_t)(c | ((c / 9) << 5) | (((c / 3) % 3) << 7) | ((c % 3) << 9) | (rank << 11));
}

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
  double pref2;             // neighbours whose box is within this squared distance are prefetched before the radius is known
  uint8_t rank[kCube];      // visiting rank of cube cell c in the reference's neighbour order, 0xff = not in the mode
  uint16_t scan[kScan];     // scan_entry() of the cell at every scan position
};
// Neighbour offsets of a neighbor_voxel_mode in the reference's visiting order (gtsam_points iVox; restated in
// oracle/ivox_ref.hpp::neighbor_offsets); returns how many.  Host side.
inline int neighbor_offsets(int mode, int8_t* off) {
  int n = 0;
  auto push = [&](int i, int j, int k) {
    off[3 * n] = (int8_t)i;
    off[3 * n + 1] = (int8_t)j;
    off[3 * n + 2] = (int8_t)k;
    ++n;
  };
  if (mode == 1) {
    push(0, 0, 0);
  } else if (mode == 7) {
    push(0, 0, 0), push(1, 0, 0), push(-1, 0, 0), push(0, 1, 0), push(0, -1, 0), push(0, 0, 1), push(0, 0, -1);
  } else {
    for (int i = -1; i <= 1; ++i)
      for (int j = -1; j <= 1; ++j)
        for (int k = -1; k <= 1; ++k)
          if (mode != 19 || i == 0 || j == 0 || k == 0) push(i, j, k);
  }
  return n;
}
// Fill MapView::rank / scan from the offset list.  Host side.
inline void fill_view_tables(MapView& v, const int8_t* off, int n_off) {
  for (int c = 0; c < kCube; ++c) v.rank[c] = 0xff;
  for (int o = 0; o < n_off; ++o) v.rank[(off[3 * o] + 1) * 9 + (off[3 * o + 1] + 1) * 3 + (off[3 * o + 2] + 1)] = (uint8_t)o;
  for (int c = 0; c < kCube; ++c)
    if (c != kCentre) v.scan[cube_pos(c)] = scan_entry(c, v.rank[c] == 0xff ? 31 : v.rank[c]);
}

constexpr int kBlockShift = 2;  // 4 x 4 x 4 voxels per block
MB_HD uint32_t cell_of(int x, int y, int z) {
  return (uint32_t)(x & 3) | ((uint32_t)(y & 3) << 2) | ((uint32_t)(z & 3) << 4);
}

MB_HD uint32_t hash_coord(int x, int y, int z) {
  uint32_t h = (uint32_t)x * 73856093u ^ (uint32_t)y * 19349669u ^ (uint32_t)z * 83492791u;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

constexpr unsigned kFull = 0xffffffffu;
constexpr int kTabEntries = 32;  // shared copy of MapView::scan (uint16 per scan position)
constexpr int kSeqShift = 5;  // sequence number = (visiting rank << 5) | point index  (cap <= 31)

#if defined(__CUDACC__)
__device__ __forceinline__ int warp_lane() { return (int)(threadIdx.x & 31u); }
#else
inline int warp_lane() { return lane_id(); }
#endif
#if defined(__CUDACC__)
// Copy the scan table into shared memory.  Call with all threads of the block, then __syncthreads().
__device__ __forceinline__ void fill_scan_table(const MapView& mv, uint16_t* s_tab) {
  if (threadIdx.x < kScan) s_tab[threadIdx.x] = mv.scan[threadIdx.x];
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// Restricted k-NN, ONE QUERY PER THREAD (every lane of the warp must call; `active` = false idles a lane).
//
// The reference (gtsam_points KnnResult::push over the neighbour voxels, restated in oracle/ivox_ref.hpp) scans
// the stored points of the 1/7/19/27 voxels around the query's voxel in a fixed visiting order and keeps the k
// smallest squared distances with a strict-'<' insertion sort, so equal distances resolve to the earlier
// visitor.  Here every thread runs that scan for its own query with these changes, none of which alters the
// result:
//   * candidates carry their visiting sequence number ((visiting rank << 5) | point index) and the list is
//     ordered by (d2, sequence), which makes the outcome independent of the order voxels are processed in;
//   * neighbour voxels are located through the block grid of the map's search mirror: <= 8 block probes
//     (own block + its three face-adjacent blocks first, then the four edge / corner blocks) give occupancy
//     masks, bucket indices follow by popcount — no per-voxel hash probe.  The 26 cube cells are enumerated
//     at compile time (offsets, block selection and cell bits are per-axis constants), the mode only decides
//     which cells take part (MapView::rank, a uniform constant-bank read);
//   * memory-level parallelism: the first chunk of the query's own bucket is requested as soon as the own
//     block entry has arrived and stays in flight while the remaining block entries are resolved and the
//     existing neighbours whose box lies within MapView::pref2 of the query are prefetched into L2 — before
//     the radius is known; neighbours that qualify only later are prefetched when they do;
//   * the query's own voxel is processed first, after which a neighbour voxel is skipped when the squared
//     distance from the query to that voxel's box (shrunk by 1e-6 voxel to stay conservative under rounding)
//     already exceeds the current k-th best — none of its points could enter the list; candidates are taken
//     four at a time;
//   * all control flow is warp-converged (uniform trip counts, per-lane predicates).
// K is the compile-time list length (5 = the reference's num_corres_points, 8 = generic: the k nearest are the
// first k of the 8 nearest).  s_pk / s_blk are this thread's columns of shared [n_off][pk_stride] /
// [24][pk_stride] arrays; s_pk receives the bucket index of every existing neighbour (by visiting rank) and
// stays valid for knn_resolve_all().
#if defined(MB_KNN_TIMING) && defined(__CUDACC__)
__device__ long long g_knn_t[16];
#define MB_KNN_T(i) do { if (blockIdx.x == MB_KNN_TIMING && threadIdx.x == 0) g_knn_t[i] = clock64(); } while (0)
#else
#define MB_KNN_T(i) do { } while (0)
#endif

template <int K>
MB_DEV void knn_thread(const MapView& mv, const uint16_t* __restrict__ s_tab, uint32_t* s_pk, uint32_t* s_blk,
                       int pk_stride, double qx, double qy, double qz, int k, bool active, double (&bd)[K],
                       uint32_t (&bs)[K]) {
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);
MB_UNROLL
  for (int i = 0; i < K; ++i) {
    bd[i] = kInf;
    bs[i] = 0xffffffffu;
  }
  const double ux = qx * mv.inv_leaf, uy = qy * mv.inv_leaf, uz = qz * mv.inv_leaf;
  const int cx = fast_floor(ux), cy = fast_floor(uy), cz = fast_floor(uz);
  const int cap = mv.cap;
  const uint32_t kCntMask = (1u << kCountBits) - 1;

  // k-th best so far (the pruning radius); for K == 5 the kernel is only launched with k == 5.
  auto worst_of = [&]() {
    double w = bd[K - 1];
    if (K != 5) {
MB_UNROLL
      for (int i = 0; i < K; ++i)
        if (i == k - 1) w = bd[i];
    }
    return w;
  };
  auto offer = [&](double d, uint32_t s) {
    if (d <= bd[K - 1]) {  // cheap gate; the exact (d2, sequence) order is applied inside
      bool lt[K];
MB_UNROLL
      for (int i = 0; i < K; ++i) lt[i] = (d < bd[i]) | ((d == bd[i]) & (s < bs[i]));
MB_UNROLL
      for (int i = K - 1; i > 0; --i) {
        bd[i] = lt[i - 1] ? bd[i - 1] : (lt[i] ? d : bd[i]);
        bs[i] = lt[i - 1] ? bs[i - 1] : (lt[i] ? s : bs[i]);
      }
      bd[0] = lt[0] ? d : bd[0];
      bs[0] = lt[0] ? s : bs[0];
    }
  };
  // Four loaded candidates j .. j+3 of a bucket (those below its count), in order.  A bucket's fill count
  // travels in the .w of its first point, so no separate metadata load precedes the first chunk: for j == 0
  // the count is decoded here (and the bucket's later cache lines are requested); returns the count.
  auto take4 = [&](const float4 (&p)[4], const float4* bucket, uint32_t rk, int j, int cnt) -> int {
    if (j == 0) {
      cnt = (int)((uint32_t)__float_as_int(p[0].w) & kCntMask);
      if (cnt > 8) prefetch_l2(bucket + 8);
      if (cnt > 16) prefetch_l2(bucket + 16);
    }
    double d[4];
MB_UNROLL
    for (int u = 0; u < 4; ++u) d[u] = sqdist4((double)p[u].x, (double)p[u].y, (double)p[u].z, qx, qy, qz);
MB_UNROLL
    for (int u = 0; u < 4; ++u)
      if (j + u < cnt) offer(d[u], (rk << kSeqShift) | (uint32_t)(j + u));
    return cnt;
  };
  auto offer4 = [&](const float4* bucket, uint32_t rk, int j, int cnt) -> int {
    float4 p[4];
MB_UNROLL
    for (int u = 0; u < 4; ++u) p[u] = __ldg(bucket + min(j + u, cap - 1));
    return take4(p, bucket, rk, j, cnt);
  };

  MB_KNN_T(0);
  // ---- (1) locate the neighbourhood's blocks; start the own bucket's loads ---------------------------------
  // Per axis the cube touches the own block ob and, when the voxel sits on a block face, one other block nb.
  // s_blk[combo] (combo bit a set = the other block on axis a) receives {mask_lo, mask_hi, base}.  Combos that
  // cannot be formed (no other block on an axis) are neither probed nor read.  Both halves of an entry are
  // fetched together, four entries in flight.
  const int ax = cx & 3, ay = cy & 3, az = cz & 3;
  const int obx = cx >> kBlockShift, oby = cy >> kBlockShift, obz = cz >> kBlockShift;
  const int nbx = obx + (ax == 0 ? -1 : 1), nby = oby + (ay == 0 ? -1 : 1), nbz = obz + (az == 0 ? -1 : 1);
  const unsigned dup_bits = ((ax == 1) | (ax == 2) ? 1u : 0u) | ((ay == 1) | (ay == 2) ? 2u : 0u) | ((az == 1) | (az == 2) ? 4u : 0u);
  uint32_t own_slot = kEmpty;
  const float4* own_bucket = mv.pts;
  float4 p0[4];
MB_UNROLL
  for (int u = 0; u < 4; ++u) p0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
MB_UNROLL
  for (int round = 0; round < 2; ++round) {
    uint32_t h[4];
    int4 e[4], m[4];
MB_UNROLL
    for (int u = 0; u < 4; ++u) {
      const int combo = round == 0 ? (u == 0 ? 0 : 1 << (u - 1)) : (u == 0 ? 3 : u == 1 ? 5 : u == 2 ? 6 : 7);
      h[u] = hash_coord((combo & 1) ? nbx : obx, (combo & 2) ? nby : oby, (combo & 4) ? nbz : obz) & mv.bmask;
      e[u] = make_int4(0, 0, 0, (int)kEmpty);
      m[u] = make_int4(0, 0, 0, 0);
      if (active && (combo & dup_bits) == 0) {
        e[u] = __ldg(mv.btab + 2 * (size_t)h[u]);
        m[u] = __ldg(mv.btab + 2 * (size_t)h[u] + 1);
      }
    }
MB_UNROLL
    for (int u = 0; u < 4; ++u) {
      const int combo = round == 0 ? (u == 0 ? 0 : 1 << (u - 1)) : (u == 0 ? 3 : u == 1 ? 5 : u == 2 ? 6 : 7);
      if ((combo & dup_bits) == 0) {
        const int bx = (combo & 1) ? nbx : obx, by = (combo & 2) ? nby : oby, bz = (combo & 4) ? nbz : obz;
        while ((uint32_t)e[u].w != kEmpty && !(e[u].x == bx && e[u].y == by && e[u].z == bz)) {
          h[u] = (h[u] + 1) & mv.bmask;
          e[u] = __ldg(mv.btab + 2 * (size_t)h[u]);
          m[u] = __ldg(mv.btab + 2 * (size_t)h[u] + 1);
        }
        const bool hit = (uint32_t)e[u].w != kEmpty;
        const uint32_t m_lo = hit ? (uint32_t)m[u].x : 0u, m_hi = hit ? (uint32_t)m[u].y : 0u;
        s_blk[(combo * 3) * pk_stride] = m_lo;
        s_blk[(combo * 3 + 1) * pk_stride] = m_hi;
        s_blk[(combo * 3 + 2) * pk_stride] = hit ? (uint32_t)e[u].w : 0u;
        if (combo == 0) {  // own block resolved: request the own bucket's first chunk now
          const uint32_t cell = (uint32_t)ax | ((uint32_t)ay << 2) | ((uint32_t)az << 4);
          const unsigned long long mk = ((unsigned long long)m_hi << 32) | m_lo;
          if ((mk >> cell) & 1ull) {
            own_slot = (uint32_t)e[u].w + (uint32_t)__popcll(mk & ((1ull << cell) - 1ull));
            own_bucket = mv.pts + (size_t)own_slot * cap;
MB_UNROLL
            for (int v = 0; v < 4; ++v) p0[v] = __ldg(own_bucket + min(v, cap - 1));
            prefetch_l2(own_bucket + 4);
          }
        }
      }
    }
  }

  MB_KNN_T(1);
  // ---- (2) which neighbours exist; prefetch the near ones ---------------------------------------------------
  // Lower bounds of the squared distance from the query to a neighbour's box, in FLOAT and rounded towards zero at
  // every step (the gaps are shrunk by 1e-6 voxel first): cheap to combine, never above the true bound.
  float g2x[3], g2y[3], g2z[3];
  {
    const double fx = ux - (double)cx, fy = uy - (double)cy, fz = uz - (double)cz;  // position inside the voxel
    const double kMargin = 1e-6;
    const double leaf = 1.0 / mv.inv_leaf;
    const double lb_scale = (leaf * leaf) * (1.0 - 1e-9);
    auto gap2 = [&](double g) {
      g = fmax(0.0, g - kMargin);
      return __double2float_rz((g * g) * lb_scale);
    };
    g2x[0] = gap2(fx), g2x[1] = 0.f, g2x[2] = gap2(1.0 - fx);
    g2y[0] = gap2(fy), g2y[1] = 0.f, g2y[2] = gap2(1.0 - fy);
    g2z[0] = gap2(fz), g2z[1] = 0.f, g2z[2] = gap2(1.0 - fz);
  }
  auto box_lb = [&](int ix, int iy, int iz) {  // compile-time indices: zero terms vanish
    float lb = 0.f;
    if (ix != 1) lb = g2x[ix];
    if (iy != 1) lb = (ix != 1) ? __fadd_rz(lb, g2y[iy]) : g2y[iy];
    if (iz != 1) lb = (ix != 1 || iy != 1) ? __fadd_rz(lb, g2z[iz]) : g2z[iz];
    return lb;
  };
  uint32_t exist = 0, near = 0;
  {
    const float pref2 = (float)mv.pref2;
    // per-axis block selection bit and cell bits of the three offsets
    const uint32_t bsel_x[3] = {ax == 0 ? 1u : 0u, 0u, ax == 3 ? 1u : 0u};
    const uint32_t bsel_y[3] = {ay == 0 ? 2u : 0u, 0u, ay == 3 ? 2u : 0u};
    const uint32_t bsel_z[3] = {az == 0 ? 4u : 0u, 0u, az == 3 ? 4u : 0u};
    const uint32_t cell_x[3] = {(uint32_t)(ax - 1) & 3u, (uint32_t)ax, (uint32_t)(ax + 1) & 3u};
    const uint32_t cell_y[3] = {((uint32_t)(ay - 1) & 3u) << 2, (uint32_t)ay << 2, ((uint32_t)(ay + 1) & 3u) << 2};
    const uint32_t cell_z[3] = {((uint32_t)(az - 1) & 3u) << 4, (uint32_t)az << 4, ((uint32_t)(az + 1) & 3u) << 4};
MB_UNROLL
    for (int c = 0; c < kCube; ++c) {
      if (c == kCentre) continue;
      const int ix = c / 9, iy = (c / 3) % 3, iz = c % 3;
      const uint32_t rk = mv.rank[c];
      if (rk != 0xffu) {  // uniform: the mode includes this cell
        const uint32_t combo = bsel_x[ix] | bsel_y[iy] | bsel_z[iz];
        const uint32_t cell = cell_x[ix] | cell_y[iy] | cell_z[iz];
        const uint32_t m_lo = s_blk[(combo * 3) * pk_stride], m_hi = s_blk[(combo * 3 + 1) * pk_stride];
        const uint32_t word = (cell & 32u) ? m_hi : m_lo;
        const uint32_t bit = 1u << (cell & 31u);
        uint32_t slot = kEmpty;
        if (active && (word & bit) != 0u)
          slot = s_blk[(combo * 3 + 2) * pk_stride] + (uint32_t)__popc(word & (bit - 1u)) + ((cell & 32u) ? (uint32_t)__popc(m_lo) : 0u);
        s_pk[rk * pk_stride] = slot;
        if (slot != kEmpty) {
          exist |= 1u << cube_pos(c);
          if (box_lb(ix, iy, iz) <= pref2) {
            near |= 1u << cube_pos(c);
            prefetch_l2(mv.pts + (size_t)slot * cap);
          }
        }
      }
    }
  }

  MB_KNN_T(2);
  // ---- (3) the query's own voxel ------------------------------------------------------------------------------
  const uint32_t rk_own = mv.rank[kCentre];
  s_pk[rk_own * pk_stride] = own_slot;
  {
    int cnt = 0;
    if (own_slot != kEmpty) cnt = take4(p0, own_bucket, rk_own, 0, 1);
    const int max_cnt = __reduce_max_sync(kFull, cnt);
    for (int j = 4; j < max_cnt; j += 4)
      if (j < cnt) offer4(own_bucket, rk_own, j, cnt);
  }

  MB_KNN_T(3);
  // ---- (4) which neighbours can still contribute ---------------------------------------------------------------
  // bit p of `todo` = the cell at scan position p exists and the lower bound of its box does not exceed the
  // current k-th best (rounded up to float); those not prefetched yet are prefetched now
  double wq = worst_of();                // the radius the candidate gate works with (refreshed at every drain)
  float wq_f = __double2float_ru(wq);
  uint32_t todo = 0;
  {
MB_UNROLL
    for (int c = 0; c < kCube; ++c) {
      if (c == kCentre) continue;
      if (mv.rank[c] != 0xffu && !(box_lb(c / 9, (c / 3) % 3, c % 3) > wq_f)) todo |= 1u << cube_pos(c);
    }
    todo &= exist;
#if defined(MB_KNN_SKIP_NB)  // experiment only (wrong results): cost of everything but the neighbour loop
    todo = 0;
#endif
    uint32_t late = todo & ~near;
    while (__any_sync(kFull, late != 0)) {
      if (late != 0) {
        const uint32_t e = s_tab[__ffs(late) - 1];
        late &= late - 1;
        prefetch_l2(mv.pts + (size_t)s_pk[(e >> 11) * pk_stride] * cap);
      }
    }
  }

  MB_KNN_T(4);
  // ---- (5) the surviving neighbours, four candidates per lane and iteration ---------------------------------
  // A lane moves to its next voxel (lowest set bit = nearest class of box) with a handful of predicated
  // instructions, re-checking the bound against the radius.  Candidates within the radius are only PUSHED onto
  // a small per-thread stack (s_blk is free by now: 8 entries of {d2 lo, d2 hi, sequence}); the ordered
  // insertion runs when a stack could overflow and at the end ("drain"), so its ~50 instructions are paid per
  // accepted candidate and not per candidate slot of the warp.  The list's (d2, sequence) order makes the result
  // independent of when a candidate is inserted; a stale radius only admits more candidates.
  int j = 0, cnt = 0, n_st = 0;
  uint32_t rk = 0;
  const float4* bucket = mv.pts;
  bool more = __any_sync(kFull, todo != 0);
  while (more) {
    if (j >= cnt && todo != 0) {
      const uint32_t e = s_tab[__ffs(todo) - 1];
      todo &= todo - 1;
      rk = e >> 11;
#if defined(MB_KNN_RECHECK)
      // re-check the bound against the radius as it stands now (only refreshed at drains, so rarely tighter than
      // when `todo` was formed: measured not worth its ~45 predicated instructions per iteration)
      const int ix = (e >> 5) & 3, iy = (e >> 7) & 3, iz = (e >> 9) & 3;
      const float lx = ix == 0 ? g2x[0] : (ix == 2 ? g2x[2] : 0.f);
      const float ly = iy == 0 ? g2y[0] : (iy == 2 ? g2y[2] : 0.f);
      const float lz = iz == 0 ? g2z[0] : (iz == 2 ? g2z[2] : 0.f);
      cnt = __fadd_rz(__fadd_rz(lx, ly), lz) > wq_f ? 0 : 1;  // real count arrives with the first chunk
#else
      cnt = 1;  // real count arrives with the first chunk
#endif
      bucket = mv.pts + (size_t)s_pk[rk * pk_stride] * cap;
#if defined(MB_KNN_FAKE_L1)  // experiment only (wrong results): what if every neighbour load hit L1?
      bucket = own_bucket;
#endif
      j = 0;
    }
    if (j < cnt) {
      float4 p[4];
MB_UNROLL
      for (int u = 0; u < 4; ++u) p[u] = __ldg(bucket + min(j + u, cap - 1));
      if (j == 0) {
        cnt = (int)((uint32_t)__float_as_int(p[0].w) & kCntMask);
        if (cnt > 8) prefetch_l2(bucket + 8);
        if (cnt > 16) prefetch_l2(bucket + 16);
      }
      double d[4];
MB_UNROLL
      for (int u = 0; u < 4; ++u) d[u] = sqdist4((double)p[u].x, (double)p[u].y, (double)p[u].z, qx, qy, qz);
MB_UNROLL
      for (int u = 0; u < 4; ++u)
        if (j + u < cnt && d[u] <= wq) {
#if defined(MB_KNN_NODEFER)
          offer(d[u], (rk << kSeqShift) | (uint32_t)(j + u));
          continue;
#endif
          s_blk[(3 * n_st) * pk_stride] = (uint32_t)__double2loint(d[u]);
          s_blk[(3 * n_st + 1) * pk_stride] = (uint32_t)__double2hiint(d[u]);
          s_blk[(3 * n_st + 2) * pk_stride] = (rk << kSeqShift) | (uint32_t)(j + u);
          ++n_st;
        }
      j += 4;
    }
    more = __any_sync(kFull, (todo != 0) | (j < cnt));
    if (!more || __any_sync(kFull, n_st > 4)) {  // drain the stacks
      while (__any_sync(kFull, n_st > 0)) {
        if (n_st > 0) {
          --n_st;
          // volatile: the three words are read together, before the gate.  (With plain loads ptxas 12.9 sinks the
          // sequence-word load below the gate and addresses it with the already decremented counter as if it
          // were the old one — it then reads the PREVIOUS entry's word; seen in k_linearize<5>, cuobjdump -sass.)
          const volatile uint32_t* ent = s_blk + (3 * n_st) * pk_stride;
          const uint32_t lo = ent[0], hi = ent[pk_stride], sq = ent[2 * pk_stride];
          offer(__hiloint2double((int)hi, (int)lo), sq);
        }
      }
      wq = worst_of();
      wq_f = __double2float_ru(wq);
    }
  }
  MB_KNN_T(5);
}

// Translate a winner's sequence number into the reference's global index and the stored point.
// Translate the winners' sequence numbers into the reference's global indices ((voxel id << 32) | point id) and,
// when kWantPts, fetch the stored points.  All loads are issued before the first use: a winner's bucket is usually
// still in L1 / L2 from the scan, but the loads of different winners must not wait for one another.  The bucket's
// meta word (voxel id << 5 | count) rides in the .w of its first point.  Returns how many of the first k exist.
template <int K, bool kWantPts>
MB_DEV int knn_resolve_all(const MapView& mv, const uint32_t* s_pk, int pk_stride, const uint32_t (&bs)[K], int k,
                           uint64_t (&g)[K], float4 (&p)[K]) {
  float w[K];
  MB_UNROLL
  for (int j = 0; j < K; ++j) {
    w[j] = 0.f;
    p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < k && bs[j] != 0xffffffffu) {
      const float4* bucket = mv.pts + (size_t)s_pk[(bs[j] >> kSeqShift) * pk_stride] * mv.cap;
      w[j] = __ldg(&bucket->w);
      if (kWantPts) p[j] = __ldg(bucket + (bs[j] & ((1u << kSeqShift) - 1)));
    }
  }
  int found = 0;
  MB_UNROLL
  for (int j = 0; j < K; ++j) {
    g[j] = ~0ull;
    if (j < k && bs[j] != 0xffffffffu) {
      g[j] = ((uint64_t)((uint32_t)__float_as_int(w[j]) >> kCountBits) << 32) | (uint64_t)(bs[j] & ((1u << kSeqShift) - 1));
      ++found;
    }
  }
  return found;
}

}  // namespace mb
