// Restricted k-NN for a WARP OF VOXEL-SORTED QUERIES: the search k_linearize runs (mb_factor.cu).
//
// A factor's scan is sorted by the map voxel of the transformed point (mb_factor.cu), so the 32 queries of a warp
// fall into a few voxels (benchmark scan: 1.9 on average, median 1).  Queries of one voxel share their whole
// neighbourhood (incremental_voxel_map.cpp:26-32 -> gtsam_points iVox::knn_search visits the same 1/7/19/27 voxels
// for all of them), so everything that depends on the voxel only is done ONCE PER VOXEL GROUP, by the whole warp:
//   1. groups: __match_any_sync on the packed voxel coordinates; the lowest lane of a group is its leader;
//   2. resolution: the (group, neighbour) pairs are dealt out to the lanes — one block-table probe and a popcount
//      each (mb_search.cuh describes the mirror) — instead of every query walking 8 block probes and 26 cube cells;
//   3. staging: when the existing neighbour buckets of all groups of the warp fit its shared-memory pool they are
//      fetched by the bulk-copy engine (cp.async.bulk.shared.global, one 16 * cap byte copy per bucket, completion on
//      an mbarrier transaction count) and every query of the group scans the shared copy; otherwise (many small
//      groups: sparse far-field voxels) the lanes read the buckets through L1 as the per-thread search does;
//   4. scan: one query per lane with the per-thread search's rules (mb_search.cuh::knn_thread): own voxel first, the
//      neighbours whose box can still contribute afterwards, four candidates per step, candidates within the radius
//      pushed on a per-lane stack and inserted in (d2, visiting sequence) order at drains.  The "can still
//      contribute" mask is OR-ed over the group so that its lanes walk the same buckets in lock step (one broadcast
//      shared-memory / L1 access per step instead of 32).
// The result per query is exactly knn_thread's: the list order (d2, then visiting rank << 5 | point index) does not
// depend on the order buckets are scanned in, and pruning only ever skips buckets that cannot hold a list member.
//
// Compiles as plain C++ for the 32-lane host emulation (tests/host_shim/search_shim.cpp): test infrastructure only.
#pragma once
#include "mb_search.cuh"

namespace mb {

constexpr int kGsStack = 8;  // candidate stack entries per lane ({d2 lo, d2 hi, sequence})

// Per-warp scratch.  ROWS = 19 (neighbourhood modes 1 / 7 / 19) or 27.
template <int ROWS>
struct GroupScratch {
  uint32_t slot[32 * ROWS];             // [group * ROWS + rank]: bucket index in the mirror, kEmpty = no such voxel
  int4 gc[32];                          // leader's voxel coordinates per group
  uint32_t exist[32];                   // per group: bit r = the neighbour with visiting rank r exists
  uint32_t todo[32];                    // per group: OR over its queries of "rank r can still contribute"
  uint32_t base[32];                    // per group: first pool bucket (staged passes)
  uint32_t stack[3 * kGsStack * 32];    // [(3 * entry + word) * 32 + lane]
  unsigned long long mbar;              // transaction barrier of the bulk copies
};

// rank -> packed cube cell: cell | ix << 5 | iy << 7 | iz << 9 (ix = dx + 1, ...); 0xffff for ranks beyond n_off
MB_HDC constexpr uint16_t rank_entry(int c) { return (uint16_t)(c | ((c / 9) << 5) | (((c / 3) % 3) << 7) | ((c % 3) << 9)); }

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// One bucket, global -> shared, by the bulk-copy engine; completes `bytes` on the barrier's transaction count.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#else
// host emulation: the copy is a memcpy, the barrier a no-op (callers __syncwarp() between issue and use)
inline void mbar_init(unsigned long long*, int) {}
inline void mbar_expect_tx(unsigned long long*, uint32_t) {}
inline void mbar_wait(unsigned long long*, uint32_t) {}
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long*) { std::memcpy(dst, src, bytes); }
#endif

// What a lane needs after the search to reach its group's buckets again (winner resolution).
struct GroupLane {
  int g;            // group ordinal inside the warp
  uint32_t exist;   // the group's existing neighbours, by rank
  int pool_base;    // first pool bucket of the group, -1 = buckets are read from global memory
  int n_groups;     // voxel groups in the warp (telemetry)
};

template <int ROWS>
MB_DEV const float4* group_bucket(const MapView& mv, const GroupScratch<ROWS>& S, const float4* pool, const GroupLane& gl,
                                  uint32_t r) {
  if (gl.pool_base >= 0) return pool + (size_t)(gl.pool_base + __popc(gl.exist & ((1u << r) - 1u))) * mv.cap;
  return mv.pts + (size_t)S.slot[gl.g * ROWS + r] * mv.cap;
}

// Every lane of the warp must call; `active` = false idles a lane.  s_rank[r] = rank_entry(cube cell of visiting rank
// r).  pool = the warp's staging pool of pool_buckets * cap float4 (pool_buckets = 0: never stage).  mbar_parity =
// the warp's running phase parity of S.mbar (0 after mbar_init).  bd / bs: the (d2, sequence) list, as knn_thread.
template <int K, int ROWS>
MB_DEV GroupLane knn_warp_groups(const MapView& mv, const uint16_t* __restrict__ s_rank, GroupScratch<ROWS>& S, float4* pool,
                                 int pool_buckets, uint32_t& mbar_parity, double qx, double qy, double qz, int k, bool active,
                                 double (&bd)[K], uint32_t (&bs)[K]) {
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);
MB_UNROLL
  for (int i = 0; i < K; ++i) {
    bd[i] = kInf;
    bs[i] = 0xffffffffu;
  }
  const int lane = warp_lane();
  const int cap = mv.cap, n_off = mv.n_off;
  const uint32_t kCntMask = (1u << kCountBits) - 1;
  const double ux = qx * mv.inv_leaf, uy = qy * mv.inv_leaf, uz = qz * mv.inv_leaf;
  const int cx = fast_floor(ux), cy = fast_floor(uy), cz = fast_floor(uz);

  // ---- (1) voxel groups ---------------------------------------------------------------------------------------
  const unsigned long long key =
      active ? ((unsigned long long)((uint32_t)cx & 0x1fffffu) | ((unsigned long long)((uint32_t)cy & 0x1fffffu) << 21) |
                ((unsigned long long)((uint32_t)cz & 0x1fffffu) << 42))
             : (~0ull - (unsigned long long)lane);
  const unsigned peers = __match_any_sync(kFull, key);
  const int leader = __ffs(peers) - 1;
  const bool is_leader = active && lane == leader;
  const unsigned leaders = __ballot_sync(kFull, is_leader);
  const int n_groups = __popc(leaders);
  GroupLane gl;
  gl.g = __popc(leaders & ((1u << leader) - 1u));
  gl.exist = 0u;
  gl.pool_base = -1;
  gl.n_groups = n_groups;
  S.exist[lane] = 0u;
  S.todo[lane] = 0u;
  if (is_leader) S.gc[gl.g] = make_int4(cx, cy, cz, 0);
  __syncwarp();

  // ---- (2) resolution: one (group, neighbour rank) pair per lane and round --------------------------------------
  const int n_tasks = n_groups * n_off;
  for (int t0 = 0; t0 < n_tasks; t0 += 32) {
    const int t = t0 + lane;
    if (t < n_tasks) {
      const int tg = t / n_off, r = t - tg * n_off;
      const int4 c = S.gc[tg];
      const uint32_t e = s_rank[r];
      const int vx = c.x + (int)((e >> 5) & 3u) - 1, vy = c.y + (int)((e >> 7) & 3u) - 1, vz = c.z + (int)((e >> 9) & 3u) - 1;
      const int bx = vx >> kBlockShift, by = vy >> kBlockShift, bz = vz >> kBlockShift;
      uint32_t h = hash_coord(bx, by, bz) & mv.bmask;
      int4 be = __ldg(mv.btab + 2 * (size_t)h), bm = __ldg(mv.btab + 2 * (size_t)h + 1);
      while ((uint32_t)be.w != kEmpty && !(be.x == bx && be.y == by && be.z == bz)) {
        h = (h + 1) & mv.bmask;
        be = __ldg(mv.btab + 2 * (size_t)h);
        bm = __ldg(mv.btab + 2 * (size_t)h + 1);
      }
      uint32_t slot = kEmpty;
      if ((uint32_t)be.w != kEmpty) {
        const uint32_t cell = cell_of(vx, vy, vz);
        const uint32_t word = (cell & 32u) ? (uint32_t)bm.y : (uint32_t)bm.x;
        const uint32_t bit = 1u << (cell & 31u);
        if (word & bit) slot = (uint32_t)be.w + (uint32_t)__popc(word & (bit - 1u)) + ((cell & 32u) ? (uint32_t)__popc((uint32_t)bm.x) : 0u);
      }
      S.slot[tg * ROWS + r] = slot;
      if (slot != kEmpty) atomicOr(&S.exist[tg], 1u << r);
    }
  }
  __syncwarp();

  // ---- (3) staging: all existing neighbour buckets of the warp's groups, if they fit the pool --------------------
  bool staged = false;
  {
    const uint32_t ex = lane < n_groups ? S.exist[lane] : 0u;
    int incl = __popc(ex);
MB_UNROLL
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(kFull, incl, 31);
    staged = total > 0 && total <= pool_buckets;
    if (staged) {
      S.base[lane] = (uint32_t)(incl - __popc(ex));
      if (lane == 0) mbar_expect_tx(&S.mbar, (uint32_t)total * (uint32_t)cap * 16u);
      __syncwarp();
      for (int t0 = 0; t0 < n_tasks; t0 += 32) {
        const int t = t0 + lane;
        if (t < n_tasks) {
          const int tg = t / n_off, r = t - tg * n_off;
          const uint32_t slot = S.slot[tg * ROWS + r];
          if (slot != kEmpty) {
            const uint32_t at = S.base[tg] + (uint32_t)__popc(S.exist[tg] & ((1u << r) - 1u));
            bulk_g2s(pool + (size_t)at * cap, mv.pts + (size_t)slot * cap, (uint32_t)cap * 16u, &S.mbar);
          }
        }
      }
    }
  }
  if (active) {
    gl.exist = S.exist[gl.g];
    if (staged) gl.pool_base = (int)S.base[gl.g];
  }

  // ---- per-query set-up that overlaps the copies: pruning bounds -----------------------------------------------
  // Lower bounds of the squared distance from the query to a neighbour's box, in FLOAT and rounded towards zero at
  // every step (the gaps are shrunk by 1e-6 voxel first): never above the true bound (as knn_thread).
  float g2x0, g2x2, g2y0, g2y2, g2z0, g2z2;
  {
    const double fx = ux - (double)cx, fy = uy - (double)cy, fz = uz - (double)cz;
    const double kMargin = 1e-6;
    const double leaf = 1.0 / mv.inv_leaf;
    const double lb_scale = (leaf * leaf) * (1.0 - 1e-9);
    auto gap2 = [&](double g) {
      g = fmax(0.0, g - kMargin);
      return __double2float_rz((g * g) * lb_scale);
    };
    g2x0 = gap2(fx), g2x2 = gap2(1.0 - fx);
    g2y0 = gap2(fy), g2y2 = gap2(1.0 - fy);
    g2z0 = gap2(fz), g2z2 = gap2(1.0 - fz);
  }
  auto box_lb = [&](uint32_t e) {  // e = s_rank entry
    const uint32_t ix = (e >> 5) & 3u, iy = (e >> 7) & 3u, iz = (e >> 9) & 3u;
    const float lx = ix == 0u ? g2x0 : (ix == 2u ? g2x2 : 0.f);
    const float ly = iy == 0u ? g2y0 : (iy == 2u ? g2y2 : 0.f);
    const float lz = iz == 0u ? g2z0 : (iz == 2u ? g2z2 : 0.f);
    return __fadd_rz(__fadd_rz(lx, ly), lz);
  };
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

  if (staged) {
    __syncwarp();  // (host emulation: the copies above are plain memcpys of other lanes)
    mbar_wait(&S.mbar, mbar_parity);
    mbar_parity ^= 1u;
  }

  // ---- (4) the query's own voxel --------------------------------------------------------------------------------
  const uint32_t rk_own = mv.rank[kCentre];
  {
    int cnt = 0;
    const float4* bucket = mv.pts;
    if (active && ((gl.exist >> rk_own) & 1u)) {
      bucket = group_bucket<ROWS>(mv, S, pool, gl, rk_own);
      cnt = 1;  // the real count arrives with the first chunk
    }
    int max_cnt = __reduce_max_sync(kFull, cnt);
    for (int j = 0; j < max_cnt; j += 4) {
      if (j < cnt) {
        float4 p[4];
MB_UNROLL
        for (int u = 0; u < 4; ++u) p[u] = bucket[min(j + u, cap - 1)];
        if (j == 0) cnt = (int)((uint32_t)__float_as_int(p[0].w) & kCntMask);
        double d[4];
MB_UNROLL
        for (int u = 0; u < 4; ++u) d[u] = sqdist4((double)p[u].x, (double)p[u].y, (double)p[u].z, qx, qy, qz);
MB_UNROLL
        for (int u = 0; u < 4; ++u)
          if (j + u < cnt) offer(d[u], (rk_own << kSeqShift) | (uint32_t)(j + u));
      }
      if (j == 0) max_cnt = __reduce_max_sync(kFull, cnt);
    }
  }

  // ---- (5) which neighbours can still contribute: per query, then OR-ed over the group ---------------------------
  double wq = worst_of();
  float wq_f = __double2float_ru(wq);
  {
    uint32_t pass = 0u;
    for (int r = 0; r < n_off; ++r) {
      if (((gl.exist >> r) & 1u) && (uint32_t)r != rk_own && !(box_lb(s_rank[r]) > wq_f)) pass |= 1u << r;
    }
    if (active && pass) atomicOr(&S.todo[gl.g], pass);
    __syncwarp();
  }
  uint32_t todo = active ? S.todo[gl.g] : 0u;

  // ---- (6) the surviving neighbours, four candidates per lane and step (knn_thread's loop) ------------------------
  uint32_t* const st = S.stack + lane;
  int j = 0, cnt = 0, n_st = 0;
  uint32_t rk = 0;
  const float4* bucket = mv.pts;
  bool more = __any_sync(kFull, todo != 0u);
  while (more) {
    if (j >= cnt && todo != 0u) {
      rk = (uint32_t)__ffs(todo) - 1u;
      todo &= todo - 1u;
      cnt = 1;  // the real count arrives with the first chunk
      bucket = group_bucket<ROWS>(mv, S, pool, gl, rk);
      j = 0;
    }
    if (j < cnt) {
      float4 p[4];
MB_UNROLL
      for (int u = 0; u < 4; ++u) p[u] = bucket[min(j + u, cap - 1)];
      if (j == 0) cnt = (int)((uint32_t)__float_as_int(p[0].w) & kCntMask);
      double d[4];
MB_UNROLL
      for (int u = 0; u < 4; ++u) d[u] = sqdist4((double)p[u].x, (double)p[u].y, (double)p[u].z, qx, qy, qz);
MB_UNROLL
      for (int u = 0; u < 4; ++u)
        if (j + u < cnt && d[u] <= wq) {
          st[(3 * n_st) * 32] = (uint32_t)__double2loint(d[u]);
          st[(3 * n_st + 1) * 32] = (uint32_t)__double2hiint(d[u]);
          st[(3 * n_st + 2) * 32] = (rk << kSeqShift) | (uint32_t)(j + u);
          ++n_st;
        }
      j += 4;
    }
    more = __any_sync(kFull, (todo != 0u) | (j < cnt));
    if (!more || __any_sync(kFull, n_st > kGsStack - 4)) {  // drain the stacks
      while (__any_sync(kFull, n_st > 0)) {
        if (n_st > 0) {
          --n_st;
          // volatile: the three words are read together, before the gate (see the note in knn_thread)
          const volatile uint32_t* ent = st + (3 * n_st) * 32;
          const uint32_t lo = ent[0], hi = ent[32], sq = ent[64];
          offer(__hiloint2double((int)hi, (int)lo), sq);
        }
      }
      wq = worst_of();
      wq_f = __double2float_ru(wq);
    }
  }
  return gl;
}

// The winners' global indices ((voxel id << 32) | point id) and stored points; returns how many of the first k exist.
// The bucket's meta word (voxel id << 5 | count) rides in the .w of its first point.
template <int K, int ROWS>
MB_DEV int group_resolve_all(const MapView& mv, const GroupScratch<ROWS>& S, const float4* pool, const GroupLane& gl,
                             const uint32_t (&bs)[K], int k, uint64_t (&g)[K], float4 (&p)[K]) {
  float w[K];
MB_UNROLL
  for (int j = 0; j < K; ++j) {
    w[j] = 0.f;
    p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < k && bs[j] != 0xffffffffu) {
      const float4* bucket = group_bucket<ROWS>(mv, S, pool, gl, bs[j] >> kSeqShift);
      w[j] = bucket->w;
      p[j] = bucket[bs[j] & ((1u << kSeqShift) - 1)];
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
