// Test-only HOST build of the search code the kernels run (mimosa_b200/csrc/mb_search.cuh).  Two emulations:
//   * one lane per query: the warp intrinsics collapse to their single-lane meaning;
//   * a 32-lane warp: 32 host threads run the same code in lock step, every warp intrinsic (__any_sync,
//     __ballot_sync, __shfl_sync, __reduce_max_sync, __syncwarp) is an exchange through a std::barrier — legal
//     because the search code is warp-converged by construction (every lane reaches the same intrinsics in the
//     same order), which is exactly the property this emulation checks along with max-over-lanes loop logic.
// Loads are plain reads, prefetches vanish.
// The search mirror (block-ordered buckets + hashed block table with occupancy masks) is rebuilt here on the
// host from a voxel dump with the same rules as k_mirror_* in mb_map.cu.  Lets the CPU test-suite check the
// search logic (block probes, cell masks, pruning bounds, deferred insertion, tie order) against the oracle
// without a GPU.  Not a product path.
#include <algorithm>
#include <barrier>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct float4 {
  float x, y, z, w;
};
struct int4 {
  int x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
// ---- warp emulation -------------------------------------------------------------------------------------
struct WarpCtx {
  std::barrier<> bar{32};
  unsigned long long slot[32];
};
static thread_local WarpCtx* t_warp = nullptr;  // nullptr: single-lane emulation
static thread_local int t_lane = 0;
template <typename F>
static inline auto warp_exchange(unsigned long long mine, F combine) {
  WarpCtx& w = *t_warp;
  w.slot[t_lane] = mine;
  w.bar.arrive_and_wait();
  const auto r = combine(w.slot);
  w.bar.arrive_and_wait();
  return r;
}
static inline unsigned __ballot_sync(unsigned, bool p) {
  if (!t_warp) return p ? 1u : 0u;
  return warp_exchange(p ? 1ull : 0ull, [](const unsigned long long* s) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)s[l] << l;
    return m;
  });
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline int __reduce_max_sync(unsigned, int v) {
  if (!t_warp) return v;
  return warp_exchange((unsigned long long)(long long)v, [](const unsigned long long* s) {
    int m = (int)(long long)s[0];
    for (int l = 1; l < 32; ++l) m = std::max(m, (int)(long long)s[l]);
    return m;
  });
}
static inline unsigned long long shfl_bits(unsigned long long v, int src) {
  if (!t_warp) return v;
  return warp_exchange(v, [src](const unsigned long long* s) { return s[src & 31]; });
}
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return (unsigned)shfl_bits(v, src); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)(unsigned)shfl_bits((unsigned)v, src); }
static inline double __shfl_sync(unsigned, double v, int src) {
  unsigned long long u;
  std::memcpy(&u, &v, 8);
  u = shfl_bits(u, src);
  std::memcpy(&v, &u, 8);
  return v;
}
static inline int __shfl_up_sync(unsigned, int v, int delta) {
  if (!t_warp) return v;
  const int lane = t_lane;
  return warp_exchange((unsigned long long)(unsigned)v, [lane, delta](const unsigned long long* s) {
    return (int)(unsigned)s[lane >= delta ? lane - delta : lane];
  });
}
static inline unsigned __match_any_sync(unsigned, unsigned long long v) {
  if (!t_warp) return 1u;
  return warp_exchange(v, [v](const unsigned long long* s) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l)
      if (s[l] == v) m |= 1u << l;
    return m;
  });
}
// shared-memory atomics of the emulated warp's lanes (host threads between two barriers really do race)
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  if (t_warp) {
    t_warp->bar.arrive_and_wait();
  }
}
namespace mb {
static inline int lane_id() { return t_lane; }
}  // namespace mb
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {  // offset-th set bit at or above `base` (offset >= 1)
  for (unsigned b = base; b < 32; ++b)
    if ((mask >> b) & 1u)
      if (--offset == 0) return b;
  return 0xffffffffu;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline int __double2loint(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (int)(uint32_t)u; }
static inline int __double2hiint(double d) { uint64_t u; std::memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  std::memcpy(&d, &u, 8);
  return d;
}
static inline float rounded(double v, int mode) {
  const int old = fegetround();
  fesetround(mode);
  volatile double in = v;
  volatile float out = (float)in;
  fesetround(old);
  return out;
}
static inline float __double2float_rz(double v) { return rounded(v, FE_TOWARDZERO); }
static inline float __double2float_ru(double v) { return rounded(v, FE_UPWARD); }
static inline float __fadd_rz(float a, float b) {
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  volatile float x = a, y = b;
  volatile float out = x + y;
  fesetround(old);
  return out;
}
static inline void prefetch_l2(const void*) {}
using std::min;

#include "../../mimosa_b200/csrc/mb_search.cuh"

namespace {
struct HostMirror {
  std::vector<float4> pts;
  std::vector<uint32_t> meta;
  std::vector<int4> btab;
  mb::MapView view{};
};

// coords int32[n_vox][3], counts int32[n_vox], xyz float[n_vox][cap][3]; voxel id = index (creation order)
void build_mirror(HostMirror& M, const int32_t* coords, const int32_t* counts, const float* xyz, uint32_t n_vox, int cap,
                  int nbr_mode, double leaf, double pref_frac) {
  std::vector<uint32_t> order(n_vox);
  for (uint32_t i = 0; i < n_vox; ++i) order[i] = i;
  auto key = [&](uint32_t id) {  // (block coordinates, cell): any block-major order works for the search
    const int32_t* c = coords + 3 * (size_t)id;
    return std::make_tuple(c[0] >> mb::kBlockShift, c[1] >> mb::kBlockShift, c[2] >> mb::kBlockShift, mb::cell_of(c[0], c[1], c[2]));
  };
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
  M.pts.assign((size_t)n_vox * cap, make_float4(0, 0, 0, 0));
  M.meta.resize(n_vox);
  struct Blk {
    int x, y, z;
    uint32_t base;
    unsigned long long mask;
  };
  std::vector<Blk> blocks;
  for (uint32_t s = 0; s < n_vox; ++s) {
    const uint32_t id = order[s];
    const int32_t* c = coords + 3 * (size_t)id;
    for (int j = 0; j < counts[id]; ++j) {
      const float* p = xyz + ((size_t)id * cap + j) * 3;
      M.pts[(size_t)s * cap + j] = make_float4(p[0], p[1], p[2], 0.f);
    }
    M.meta[s] = (id << mb::kCountBits) | (uint32_t)counts[id];
    M.pts[(size_t)s * cap].w = __int_as_float((int)M.meta[s]);
    const int bx = c[0] >> mb::kBlockShift, by = c[1] >> mb::kBlockShift, bz = c[2] >> mb::kBlockShift;
    if (blocks.empty() || blocks.back().x != bx || blocks.back().y != by || blocks.back().z != bz)
      blocks.push_back(Blk{bx, by, bz, s, 0ull});
    blocks.back().mask |= 1ull << mb::cell_of(c[0], c[1], c[2]);
  }
  size_t bcap = 16;
  while (bcap < 2 * blocks.size()) bcap <<= 1;
  M.btab.assign(2 * bcap, make_int4(0, 0, 0, (int)mb::kEmpty));
  for (const Blk& b : blocks) {
    uint32_t h = mb::hash_coord(b.x, b.y, b.z) & (uint32_t)(bcap - 1);
    while ((uint32_t)M.btab[2 * (size_t)h].w != mb::kEmpty) h = (h + 1) & (uint32_t)(bcap - 1);
    M.btab[2 * (size_t)h] = make_int4(b.x, b.y, b.z, (int)b.base);
    M.btab[2 * (size_t)h + 1] = make_int4((int)(uint32_t)(b.mask & 0xffffffffull), (int)(uint32_t)(b.mask >> 32), 0, 0);
  }
  mb::MapView& v = M.view;
  v.btab = M.btab.data();
  v.bmask = (uint32_t)(bcap - 1);
  v.pts = M.pts.data();
  v.meta = M.meta.data();
  v.cap = cap;
  int8_t off[mb::kMaxNbr * 3];
  v.n_off = mb::neighbor_offsets(nbr_mode, off);
  v.inv_leaf = 1.0 / leaf;
  v.pref2 = pref_frac * pref_frac * leaf * leaf;
  mb::fill_view_tables(v, off, v.n_off);
}

template <int K>
void run(const HostMirror& M, const double* q, size_t nq, int k, uint64_t* idx, double* d2, uint8_t* ok) {
  std::vector<uint32_t> s_pk(mb::kMaxNbr), s_blk(24);
  uint16_t s_tab[mb::kTabEntries] = {0};
  for (int p = 0; p < mb::kScan; ++p) s_tab[p] = M.view.scan[p];
  for (size_t i = 0; i < nq; ++i) {
    double bd[K];
    uint32_t bs[K];
    std::fill(s_pk.begin(), s_pk.end(), 0xdeadbeefu);  // stale contents must never be used
    std::fill(s_blk.begin(), s_blk.end(), 0xdeadbeefu);
    mb::knn_thread<K>(M.view, s_tab, s_pk.data(), s_blk.data(), 1, q[3 * i], q[3 * i + 1], q[3 * i + 2], k, true, bd, bs);
    uint64_t g[K];
    float4 pts[K];
    const int found = mb::knn_resolve_all<K, true>(M.view, s_pk.data(), 1, bs, k, g, pts);
    for (int j = 0; j < k; ++j) {
      idx[i * k + j] = g[j];
      d2[i * k + j] = g[j] != ~0ull ? bd[j] : 1.7976931348623157e308;
    }
    ok[i] = found == k;
  }
}
// 32 queries per emulated warp, one host thread per lane, shared-memory columns laid out as on the device
template <int K>
void run_warps(const HostMirror& M, const double* q, size_t nq, int k, uint64_t* idx, double* d2, uint8_t* ok) {
  uint16_t s_tab[mb::kTabEntries] = {0};
  for (int p = 0; p < mb::kScan; ++p) s_tab[p] = M.view.scan[p];
  for (size_t w0 = 0; w0 < nq; w0 += 32) {
    WarpCtx ctx;
    std::vector<uint32_t> s_pk(mb::kMaxNbr * 32, 0xdeadbeefu), s_blk(24 * 32, 0xdeadbeefu);
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; ++lane)
      lanes.emplace_back([&, lane] {
        t_warp = &ctx;
        t_lane = lane;
        const size_t i = w0 + lane;
        const bool active = i < nq;
        const size_t qi = active ? i : 0;
        double bd[K];
        uint32_t bs[K];
        mb::knn_thread<K>(M.view, s_tab, s_pk.data() + lane, s_blk.data() + lane, 32, q[3 * qi], q[3 * qi + 1], q[3 * qi + 2], k, active, bd, bs);
        if (active) {
          uint64_t g[K];
          float4 pts[K];
          const int found = mb::knn_resolve_all<K, true>(M.view, s_pk.data() + lane, 32, bs, k, g, pts);
          for (int j = 0; j < k; ++j) {
            idx[i * k + j] = g[j];
            d2[i * k + j] = g[j] != ~0ull ? bd[j] : 1.7976931348623157e308;
          }
          ok[i] = found == k;
        }
        t_warp = nullptr;
      });
    for (auto& t : lanes) t.join();
  }
}
}  // namespace

extern "C" int shim_knn(const int32_t* coords, const int32_t* counts, const float* xyz, uint32_t n_vox, int cap, int nbr_mode,
                        double leaf, double pref_frac, const double* q, size_t nq, int k, uint64_t* idx, double* d2, uint8_t* ok) {
  if (k < 1 || k > 8) return 1;
  HostMirror M;
  build_mirror(M, coords, counts, xyz, n_vox, cap, nbr_mode, leaf, pref_frac);
  if (k == 5)
    run<5>(M, q, nq, k, idx, d2, ok);
  else
    run<8>(M, q, nq, k, idx, d2, ok);
  return 0;
}

// same, 32-lane warp emulation
extern "C" int shim_knn_warp(const int32_t* coords, const int32_t* counts, const float* xyz, uint32_t n_vox, int cap,
                             int nbr_mode, double leaf, double pref_frac, const double* q, size_t nq, int k, uint64_t* idx,
                             double* d2, uint8_t* ok) {
  if (k < 1 || k > 8) return 1;
  HostMirror M;
  build_mirror(M, coords, counts, xyz, n_vox, cap, nbr_mode, leaf, pref_frac);
  if (k == 5)
    run_warps<5>(M, q, nq, k, idx, d2, ok);
  else
    run_warps<8>(M, q, nq, k, idx, d2, ok);
  return 0;
}

