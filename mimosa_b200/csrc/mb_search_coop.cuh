// Restricted k-NN, G LANES PER QUERY (G = 4 or 8): the cooperative variant of mb_search.cuh::knn_thread.
//
// EXPERIMENTAL — selected with MB_KNN_VARIANT=coop4 / coop8 / coop4p / ... (mb_map.cu::launch_knn); the default search
// is still knn_thread.  Checked on the CPU against the oracle through the 32-lane warp emulation
// (tests/host_shim/search_shim.cpp, tests/test_search_host.py::test_coop_*) and on a B200 against the default kernel
// (tools/knn_variants.py: bit-exact on 3 x 131 072 queries).  MEASURED (profiles/r1_experiments.md, session 4): 2.2x
// faster than knn_thread on a launch of 8 192 queries (28 vs 60 us: the dependent chain is that much shorter), 8-55 %
// slower on the full 131 072 (86-99 us at G = 4, 117-126 us at G = 8, against 80 us): the per-query fixed work is
// spread over G lanes instead of one, ~330 / ~600 warp-instructions per query against 207, and the full launch is
// issue-bound.  Kept as the starting point for the next design (fewer lanes per query or a warp-wide bucket queue).
//
// The idea: with one query per thread the launch is a single wave of warps and lasts as long as the slowest warp's
// serial chain (~35 dependent memory steps, ~12 k dependent instructions).  Here the lanes of a group work on the SAME
// query so that its dependent chain is short and the launch runs as several waves of short-lived warps whose phases
// interleave:
//   1  block probes     the <= 8 blocks around the query's voxel, 8 / G per lane, ONE round of loads (one 256-bit
//                       load per entry); masks, bases and the per-axis box gaps are left in shared memory;
//   2  own voxel        every lane locates the query's own bucket (group lane 0 holds the own block's entry) and
//                       requests a different four-point chunk of it — speculatively up to the cap, the fill count
//                       arrives with the first chunk — so that the loads fly while step 3 computes;
//   3  cube cells       the 26 neighbour cells in scan-position order (faces, edges, corners), ceil(26 / G) per
//                       lane: occupancy bit, bucket index, lower bound of the box distance (byte-parallel arithmetic
//                       on a packed cell table);
//   4  merge            every lane keeps a private (d2, sequence)-ordered K-list; K rounds of "group minimum of the
//                       list heads" leave the merged list in EVERY lane (replicated), whose k-th entry is the radius;
//   5  neighbours       the surviving cells (box bound <= radius) are compacted into a per-group queue in scan order
//                       and — MODE 0 — dealt out one bucket per lane and round, STEP points per step, or — MODE 1 —
//                       visited two at a time by the whole group, lane gl reading points gl, gl + G, ... (every load
//                       instruction covers G consecutive points); candidates within the lane's radius are pushed on
//                       a per-thread stack and inserted at the drains (as in knn_thread);
//   6  merge            as 4; entries replicated in step 4 sit at the head of every list that holds them when they
//                       are the minimum and are popped together, so the result has no duplicates.
// The (d2, sequence) order is the reference's visiting order (KnnResult::push, oracle/ivox_ref.hpp), so the outcome
// does not depend on which lane saw which candidate.  A lane's radius is the k-th entry of its own list: never below
// the final k-th distance, so pruning with it cannot drop a winner.
// All control flow is warp-converged (shuffles use the full mask).  Compiles as plain C++ for the host emulation.
#pragma once
#include "mb_search.cuh"

namespace mb {

#if defined(__CUDACC__)
__device__ __forceinline__ int coop_lane() { return (int)(threadIdx.x & 31u); }
#else
inline int coop_lane() { return lane_id(); }
#endif

// Both halves of a block-table entry (32 bytes, 32-byte aligned) with ONE load instruction (LDG.E.256 on sm_100a):
// half the load instructions of the probe round.  (Measured neutral on the launch time, like the coalesced MODE 1
// loads: the L1 wavefront queue is not what bounds these kernels; profiles/r1_experiments.md, session 4.)
#if defined(__CUDACC__)
__device__ __forceinline__ void ld_block_entry(const int4* p, int4& e, int4& m) {
  asm volatile("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w), "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w)
               : "l"(p));
}
#else
inline void ld_block_entry(const int4* p, int4& e, int4& m) {
  e = p[0];
  m = p[1];
}
#endif

constexpr int kCoopQueue = 26;   // survivors per query
constexpr int kCoopStack = 8;    // candidates a lane can push per step
constexpr int kCoopBlk = 44;     // {mask_lo, mask_hi, base, base + popc(mask_lo)} of the 8 blocks around a query + 3 x 4 gap words

// Cell-table entry of scan position p for knn_group (see step 3), from MapView::scan[p].
MB_HDC inline uint32_t coop_tab_entry(uint16_t scan_e) {
  const uint32_t ix = (scan_e >> 5) & 3u, iy = (scan_e >> 7) & 3u, iz = (scan_e >> 9) & 3u, rk = (uint32_t)scan_e >> 11;
  return (ix + 3u) | ((iy + 3u) << 8) | ((iz + 3u) << 16) | (rk << 24);
}

// s_ctab: [kScan] coop_tab_entry() values (block-shared); s_pk: this GROUP's [kCube] bucket index by visiting rank
// (valid afterwards for the winners' resolution); s_blk: this group's [kCoopBlk] block / gap words; s_q: this group's
// [kCoopQueue]; s_st: this THREAD's column of a [3 * kCoopStack][st_stride] array.  Every lane of the warp must call;
// all lanes of a group pass the same query.  On return every lane of the group holds the same (bd, bs): the K best in
// (d2, sequence) order, +inf / 0xffffffff where fewer exist.
template <int K, int G, int STEP = 8, int MODE = 0>
MB_DEV void knn_group(const MapView& mv, const uint32_t* __restrict__ s_ctab, uint32_t* s_pk, uint32_t* s_blk, uint32_t* s_q,
                      uint32_t* s_st, int st_stride, double qx, double qy, double qz, int k, bool active,
                      double (&bd)[K], uint32_t (&bs)[K]) {
  static_assert(G == 4 || G == 8, "4 or 8 lanes per query");
  static_assert(STEP == 4 || STEP == 8, "4 or 8 points per neighbour step");
  static_assert(MODE == 0 || MODE == 1, "0: a bucket per lane, 1: a point per lane");
  constexpr int NOWN = MODE == 0 ? 4 : (20 + G - 1) / G;  // own-bucket loads per lane in the first round
  constexpr int PPL = 8 / G;                               // MODE 1: loads per lane that cover a bucket's first 8 points
  constexpr int NP = 8 / G;                // block probes per lane
  constexpr int T = (kScan + G - 1) / G;   // neighbour cells per lane
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);
  const int lane = coop_lane(), gl = lane & (G - 1), gbase = lane & ~(G - 1);
  const uint32_t kCntMask = (1u << kCountBits) - 1;
  const int cap = mv.cap;
MB_UNROLL
  for (int i = 0; i < K; ++i) {
    bd[i] = kInf;
    bs[i] = 0xffffffffu;
  }
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
    if (d <= bd[K - 1]) {
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
  // K rounds of "minimum of the group's list heads"; equal heads (replicated entries) are popped together.
  auto merge = [&]() {
    double nd[K];
    uint32_t ns[K];
MB_UNROLL
    for (int r = 0; r < K; ++r) {
      double md = bd[0];
      uint32_t ms = bs[0];
MB_UNROLL
      for (int off = 1; off < G; off <<= 1) {
        const double od = __shfl_sync(kFull, md, lane ^ off);
        const uint32_t os = __shfl_sync(kFull, ms, lane ^ off);
        const bool take = (od < md) | ((od == md) & (os < ms));
        md = take ? od : md;
        ms = take ? os : ms;
      }
      nd[r] = md;
      ns[r] = ms;
      const bool pop = (bd[0] == md) & (bs[0] == ms);
MB_UNROLL
      for (int i = 0; i + 1 < K; ++i) {
        bd[i] = pop ? bd[i + 1] : bd[i];
        bs[i] = pop ? bs[i + 1] : bs[i];
      }
      bd[K - 1] = pop ? kInf : bd[K - 1];
      bs[K - 1] = pop ? 0xffffffffu : bs[K - 1];
    }
MB_UNROLL
    for (int r = 0; r < K; ++r) {
      bd[r] = nd[r];
      bs[r] = ns[r];
    }
  };

  const double ux = qx * mv.inv_leaf, uy = qy * mv.inv_leaf, uz = qz * mv.inv_leaf;
  const int cx = fast_floor(ux), cy = fast_floor(uy), cz = fast_floor(uz);
  const int ax = cx & 3, ay = cy & 3, az = cz & 3;

  // ---- 1: block probes (combo bit a set = the other block on axis a; lane gl takes combos gl, gl + G) -------------
  // Every combo's {mask_lo, mask_hi, base, base + popc(mask_lo)} goes to s_blk[combo * 4 ..] (zeros when the block
  // is absent or cannot be formed), where step 3 picks it up.
  {
    const int obx = cx >> kBlockShift, oby = cy >> kBlockShift, obz = cz >> kBlockShift;
    const int nbx = obx + (ax == 0 ? -1 : 1), nby = oby + (ay == 0 ? -1 : 1), nbz = obz + (az == 0 ? -1 : 1);
    const unsigned dup_bits = ((ax == 1) | (ax == 2) ? 1u : 0u) | ((ay == 1) | (ay == 2) ? 2u : 0u) | ((az == 1) | (az == 2) ? 4u : 0u);
    uint32_t h[NP];
    int4 e[NP], m[NP];
    int bx[NP], by[NP], bz[NP];
    bool miss[NP];
    bool any_miss = false;
MB_UNROLL
    for (int u = 0; u < NP; ++u) {
      const unsigned combo = (unsigned)(gl + G * u);
      bx[u] = (combo & 1u) ? nbx : obx, by[u] = (combo & 2u) ? nby : oby, bz[u] = (combo & 4u) ? nbz : obz;
      h[u] = hash_coord(bx[u], by[u], bz[u]) & mv.bmask;
      e[u] = make_int4(0, 0, 0, (int)kEmpty);
      m[u] = make_int4(0, 0, 0, 0);
      if (active && (combo & dup_bits) == 0u) ld_block_entry(mv.btab + 2 * (size_t)h[u], e[u], m[u]);
    }
    // while the entries are in flight: group lane a < 3 leaves the squared gaps from the query to the lower / upper
    // neighbour slab of axis a (float, rounded towards zero at every step, shrunk by 1e-6 voxel: never above the true
    // bound) in s_blk[32 + 4 a + {0, 2}], zero between them — step 3 sums three of these per cell
    if (gl < 3) {
      const double f = gl == 0 ? ux - (double)cx : (gl == 1 ? uy - (double)cy : uz - (double)cz);  // position inside the voxel
      const double kMargin = 1e-6;
      const double leaf = 1.0 / mv.inv_leaf;
      const double lb_scale = (leaf * leaf) * (1.0 - 1e-9);
      auto gap2 = [&](double g) {
        g = fmax(0.0, g - kMargin);
        return (uint32_t)__float_as_int(__double2float_rz((g * g) * lb_scale));
      };
      uint32_t* dst = s_blk + 32 + 4 * gl;
      dst[0] = gap2(f);
      dst[1] = 0u;
      dst[2] = gap2(1.0 - f);
    }
MB_UNROLL
    for (int u = 0; u < NP; ++u) {
      miss[u] = (uint32_t)e[u].w != kEmpty && !(e[u].x == bx[u] && e[u].y == by[u] && e[u].z == bz[u]);
      any_miss |= miss[u];
    }
    while (__any_sync(kFull, any_miss)) {  // linear probing past a colliding entry (rare)
      any_miss = false;
MB_UNROLL
      for (int u = 0; u < NP; ++u) {
        if (miss[u]) {
          h[u] = (h[u] + 1) & mv.bmask;
          ld_block_entry(mv.btab + 2 * (size_t)h[u], e[u], m[u]);
          miss[u] = (uint32_t)e[u].w != kEmpty && !(e[u].x == bx[u] && e[u].y == by[u] && e[u].z == bz[u]);
        }
        any_miss |= miss[u];
      }
    }
MB_UNROLL
    for (int u = 0; u < NP; ++u) {
      const bool hit = (uint32_t)e[u].w != kEmpty;
      const uint32_t m_lo = hit ? (uint32_t)m[u].x : 0u, m_hi = hit ? (uint32_t)m[u].y : 0u, base = hit ? (uint32_t)e[u].w : 0u;
      uint32_t* dst = s_blk + 4 * (gl + G * u);
      dst[0] = m_lo;
      dst[1] = m_hi;
      dst[2] = base;
      dst[3] = base + (uint32_t)__popc(m_lo);
    }
  }
  __syncwarp();

  // ---- 2: the query's own bucket: chunk c goes to group lane c % G; first round requested now ---------------------------
  const uint32_t rk_own = mv.rank[kCentre];
  uint32_t own_slot = kEmpty;
  {
    const uint32_t cell = (uint32_t)ax | ((uint32_t)ay << 2) | ((uint32_t)az << 4);
    const uint32_t word = s_blk[cell >> 5], bit = 1u << (cell & 31u);
    if (active && (word & bit) != 0u) own_slot = s_blk[2 + (cell >> 5)] + (uint32_t)__popc(word & (bit - 1u));
    if (gl == 0) s_pk[rk_own] = own_slot;
  }
  const float4* own_bucket = mv.pts + (size_t)(own_slot != kEmpty ? own_slot : 0u) * cap;
  const int n_chunks = (cap + 3) >> 2;
  float4 p0[NOWN];
MB_UNROLL
  for (int u = 0; u < NOWN; ++u) p0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (MODE == 0) {  // lane gl: the four points of chunk gl
    if (own_slot != kEmpty && gl < n_chunks) {
MB_UNROLL
      for (int u = 0; u < 4; ++u) p0[u] = __ldg(own_bucket + min(4 * gl + u, cap - 1));
    }
  } else {  // lane gl: points gl, gl + G, ... — every load instruction reads G consecutive points per group
MB_UNROLL
    for (int u = 0; u < NOWN; ++u)
      if (own_slot != kEmpty && G * u + gl < cap) p0[u] = __ldg(own_bucket + G * u + gl);
  }

  // ---- 3: neighbour cells: lane gl takes scan positions gl, gl + G, ... ---------------------------------------------
  // s_ctab[p] = (ix + 3) | (iy + 3) << 8 | (iz + 3) << 16 | rank << 24 with ix = dx + 1 ...: adding the query's in-block
  // position per byte gives a + i + 3 in [3, 8] per axis, whose low two bits are the neighbour's in-block coordinate
  // and whose bit 2 says "same block".
  const float pref2 = (float)mv.pref2;
  const uint32_t a_pack = (uint32_t)ax | ((uint32_t)ay << 8) | ((uint32_t)az << 16);
  uint32_t cslot[T], crk[T];
  float clb[T];
  bool cnear[T];
MB_UNROLL
  for (int t = 0; t < T; ++t) {
    const int p = gl + G * t;
    const bool valid = p < kScan;
    const uint32_t e = s_ctab[valid ? p : 0];
    const uint32_t rk = e >> 24;  // 31: the cell is not part of the neighbourhood mode
    const uint32_t w = a_pack + (e & 0xffffffu);
    const uint32_t cb = w & 0x030303u, ob = (~w >> 2) & 0x010101u;
    const uint32_t cell = (cb | (cb >> 6) | (cb >> 12)) & 63u;
    const uint32_t combo = (ob | (ob >> 7) | (ob >> 14)) & 7u;
    const uint32_t* blk = s_blk + 4 * combo + (cell >> 5);
    const uint32_t word = blk[0], bit = 1u << (cell & 31u);
    uint32_t slot = kEmpty;
    if (active && valid && rk != 31u && (word & bit) != 0u) slot = blk[2] + (uint32_t)__popc(word & (bit - 1u));
    // per-axis gap: byte value 3 / 4 / 5 = lower neighbour / same slab / upper neighbour
    const float lx = __int_as_float((int)s_blk[32 - 3 + (e & 0xffu)]);
    const float ly = __int_as_float((int)s_blk[36 - 3 + ((e >> 8) & 0xffu)]);
    const float lz = __int_as_float((int)s_blk[40 - 3 + ((e >> 16) & 0xffu)]);
    cslot[t] = slot;
    crk[t] = rk;
    clb[t] = __fadd_rz(__fadd_rz(lx, ly), lz);
    cnear[t] = false;
    if (valid && rk != 31u) s_pk[rk] = slot;
    if (slot != kEmpty && clb[t] <= pref2) {  // near enough to be wanted whatever the radius turns out to be
      cnear[t] = true;
      prefetch_l2(mv.pts + (size_t)slot * cap);
    }
  }
  __syncwarp();

  // the own voxel's candidates: the first round's NOWN per lane go through a sorting network straight into the
  // (empty) list; further rounds only happen for buckets larger than the first round covers
  {
    int cnt = 0;
    {  // the fill count rides in the .w of the first point: group lane 0 has it
      const uint32_t w0 = __shfl_sync(kFull, (uint32_t)__float_as_int(p0[0].w), gbase);
      if (own_slot != kEmpty) cnt = (int)(w0 & kCntMask);
    }
    double d[NOWN];
    uint32_t sq[NOWN];
MB_UNROLL
    for (int u = 0; u < NOWN; ++u) {
      const int j = MODE == 0 ? 4 * gl + u : G * u + gl;
      const bool ok = j < cnt;
      d[u] = ok ? sqdist4((double)p0[u].x, (double)p0[u].y, (double)p0[u].z, qx, qy, qz) : kInf;
      sq[u] = ok ? (rk_own << kSeqShift) | (uint32_t)j : 0xffffffffu;
    }
    auto cswap = [&](int x, int y) {  // x < y: afterwards entry x <= entry y in (d2, sequence) order
      const bool sw = (d[y] < d[x]) | ((d[y] == d[x]) & (sq[y] < sq[x]));
      const double dx = d[x], dy = d[y];
      const uint32_t sx = sq[x], sy = sq[y];
      d[x] = sw ? dy : dx, d[y] = sw ? dx : dy;
      sq[x] = sw ? sy : sx, sq[y] = sw ? sx : sy;
    };
    if constexpr (NOWN == 3) {
      cswap(0, 1), cswap(1, 2), cswap(0, 1);
    } else if constexpr (NOWN == 4) {
      cswap(0, 1), cswap(2, 3), cswap(0, 2), cswap(1, 3), cswap(1, 2);
    } else {
      static_assert(NOWN == 5, "sorting networks for 3, 4 and 5 entries");
      cswap(0, 1), cswap(3, 4), cswap(2, 4), cswap(2, 3), cswap(1, 4), cswap(0, 3), cswap(0, 2), cswap(1, 3), cswap(1, 2);
    }
    static_assert(NOWN <= K, "the first round fits the list");
MB_UNROLL
    for (int u = 0; u < NOWN; ++u) bd[u] = d[u], bs[u] = sq[u];
    if constexpr (MODE == 0) {
      for (int c0 = G; __any_sync(kFull, 4 * c0 < cnt); c0 += G) {
        const int j = 4 * (c0 + gl);
        if (c0 + gl < n_chunks && j < cnt) {
MB_UNROLL
          for (int u = 0; u < 4; ++u) p0[u] = __ldg(own_bucket + min(j + u, cap - 1));
MB_UNROLL
          for (int u = 0; u < 4; ++u)
            if (j + u < cnt) offer(sqdist4((double)p0[u].x, (double)p0[u].y, (double)p0[u].z, qx, qy, qz), (rk_own << kSeqShift) | (uint32_t)(j + u));
        }
      }
    } else {
      for (int j0 = NOWN * G; __any_sync(kFull, j0 < cnt); j0 += G) {
        const int j = j0 + gl;
        if (j < cnt) {
          const float4 p = __ldg(own_bucket + j);
          offer(sqdist4((double)p.x, (double)p.y, (double)p.z, qx, qy, qz), (rk_own << kSeqShift) | (uint32_t)j);
        }
      }
    }
  }
  // ---- 4: merged list in every lane; its k-th entry is the radius ---------------------------------------------------
  merge();
  double wq = worst_of();
  float wq_f = __double2float_ru(wq);

  // ---- 5: surviving neighbours, compacted in scan order (t-major, lane-minor), one bucket per lane and round ----------
  int n_surv = 0;
MB_UNROLL
  for (int t = 0; t < T; ++t) {
    const bool surv = cslot[t] != kEmpty && !(clb[t] > wq_f);
    const uint32_t gb = (__ballot_sync(kFull, surv) >> gbase) & ((1u << G) - 1u);
    if (surv) {
      // queue word: the bound with its low five mantissa bits cleared (still a lower bound) | visiting rank
      s_q[n_surv + __popc(gb & ((1u << gl) - 1u))] = ((uint32_t)__float_as_int(clb[t]) & ~31u) | crk[t];
      if (!cnear[t]) prefetch_l2(mv.pts + (size_t)cslot[t] * cap);
    }
    n_surv += __popc(gb);
  }
  __syncwarp();
  const bool any_surv = __any_sync(kFull, n_surv > 0);
  int n_st = 0;
  auto push = [&](double d, uint32_t seq) {
    s_st[(3 * n_st) * st_stride] = (uint32_t)__double2loint(d);
    s_st[(3 * n_st + 1) * st_stride] = (uint32_t)__double2hiint(d);
    s_st[(3 * n_st + 2) * st_stride] = seq;
    ++n_st;
  };
  auto drain = [&]() {
    while (__any_sync(kFull, n_st > 0)) {
      if (n_st > 0) {
        --n_st;
        // volatile: the three words are read together, before offer()'s gate (see the note in knn_thread)
        const volatile uint32_t* ent = s_st + (3 * n_st) * st_stride;
        const uint32_t lo = ent[0], hi = ent[st_stride], sq = ent[2 * st_stride];
        offer(__hiloint2double((int)hi, (int)lo), sq);
      }
    }
    wq = worst_of();  // this lane's radius: replicated list + its own candidates
    wq_f = __double2float_ru(wq);
  };
  if constexpr (MODE == 1) {
    // The whole group works on the same two buckets per round: lane gl reads points gl, gl + G, ... so that every load
    // instruction covers G consecutive points (one line) per group; a lane keeps the candidates it computed.
    for (int b0 = 0; __any_sync(kFull, b0 < n_surv); b0 += 2) {
      const uint32_t wa = b0 < n_surv ? s_q[b0] : 0u, wb = b0 + 1 < n_surv ? s_q[b0 + 1] : 0u;
      const uint32_t rka = wa & 31u, rkb = wb & 31u;
      // (the radius may have tightened since the queue was formed; lanes of a group may disagree, which is fine:
      // a lane only skips work its own radius rules out)
      const bool ha = b0 < n_surv && !(__int_as_float((int)(wa & ~31u)) > wq_f);
      const bool hb = b0 + 1 < n_surv && !(__int_as_float((int)(wb & ~31u)) > wq_f);
      const float4* ba = mv.pts + (size_t)(b0 < n_surv ? s_pk[rka] : 0u) * cap;
      const float4* bb = mv.pts + (size_t)(b0 + 1 < n_surv ? s_pk[rkb] : 0u) * cap;
      float4 pa[PPL], pb[PPL];
MB_UNROLL
      for (int u = 0; u < PPL; ++u) {
        pa[u] = pb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        // group lane 0 always fetches the first point of a bucket the group visits: it carries the fill count
        if ((ha || (gl == 0 && u == 0 && b0 < n_surv)) && G * u + gl < cap) pa[u] = __ldg(ba + G * u + gl);
        if ((hb || (gl == 0 && u == 0 && b0 + 1 < n_surv)) && G * u + gl < cap) pb[u] = __ldg(bb + G * u + gl);
      }
      const uint32_t ca = __shfl_sync(kFull, (uint32_t)__float_as_int(pa[0].w), gbase) & kCntMask;
      const uint32_t cb = __shfl_sync(kFull, (uint32_t)__float_as_int(pb[0].w), gbase) & kCntMask;
      const int cnta = ha ? (int)ca : 0, cntb = hb ? (int)cb : 0;
MB_UNROLL
      for (int u = 0; u < PPL; ++u) {
        const int j = G * u + gl;
        const double da = sqdist4((double)pa[u].x, (double)pa[u].y, (double)pa[u].z, qx, qy, qz);
        if (j < cnta && da <= wq) push(da, (rka << kSeqShift) | (uint32_t)j);
        const double db = sqdist4((double)pb[u].x, (double)pb[u].y, (double)pb[u].z, qx, qy, qz);
        if (j < cntb && db <= wq) push(db, (rkb << kSeqShift) | (uint32_t)j);
      }
      drain();
      for (int j0 = 8; __any_sync(kFull, (j0 < cnta) | (j0 < cntb)); j0 += G) {  // buckets with more than 8 points
        const int j = j0 + gl;
        if (j < cnta) {
          const float4 p = __ldg(ba + j);
          const double d = sqdist4((double)p.x, (double)p.y, (double)p.z, qx, qy, qz);
          if (d <= wq) push(d, (rka << kSeqShift) | (uint32_t)j);
        }
        if (j < cntb) {
          const float4 p = __ldg(bb + j);
          const double d = sqdist4((double)p.x, (double)p.y, (double)p.z, qx, qy, qz);
          if (d <= wq) push(d, (rkb << kSeqShift) | (uint32_t)j);
        }
        drain();
      }
    }
  } else {
    for (int r0 = 0; __any_sync(kFull, r0 < n_surv); r0 += G) {
      const int item = r0 + gl;
      bool has = item < n_surv;
      const uint32_t w = has ? s_q[item] : 0u;
      const uint32_t rk = w & 31u;
      if (has && r0 > 0 && __int_as_float((int)(w & ~31u)) > wq_f) has = false;  // the radius has tightened since
      const float4* bucket = mv.pts + (size_t)(has ? s_pk[rk] : 0u) * cap;
      int cnt = has ? 1 : 0;  // the real count arrives with the first chunk
      for (int j = 0; __any_sync(kFull, j < cnt); j += STEP) {
        if (j < cnt) {
          float4 p[STEP];
  MB_UNROLL
          for (int u = 0; u < STEP; ++u) p[u] = __ldg(bucket + min(j + u, cap - 1));
          if (j == 0) cnt = (int)((uint32_t)__float_as_int(p[0].w) & kCntMask);
  MB_UNROLL
          for (int u = 0; u < STEP; ++u) {
            const double d = sqdist4((double)p[u].x, (double)p[u].y, (double)p[u].z, qx, qy, qz);
            if (j + u < cnt && d <= wq) {
              s_st[(3 * n_st) * st_stride] = (uint32_t)__double2loint(d);
              s_st[(3 * n_st + 1) * st_stride] = (uint32_t)__double2hiint(d);
              s_st[(3 * n_st + 2) * st_stride] = (rk << kSeqShift) | (uint32_t)(j + u);
              ++n_st;
            }
          }
        }
        while (__any_sync(kFull, n_st > 0)) {  // drain
          if (n_st > 0) {
            --n_st;
            // volatile: the three words are read together, before offer()'s gate (see the note in knn_thread)
            const volatile uint32_t* ent = s_st + (3 * n_st) * st_stride;
            const uint32_t lo = ent[0], hi = ent[st_stride], sq = ent[2 * st_stride];
            offer(__hiloint2double((int)hi, (int)lo), sq);
          }
        }
        wq = worst_of();  // this lane's radius: replicated list + its own candidates
        wq_f = __double2float_ru(wq);
      }
    }
  }
  // ---- 6: final merge (nothing to do when no group of the warp had a surviving neighbour) ---------------------------
  if (any_surv) merge();
}

}  // namespace mb
