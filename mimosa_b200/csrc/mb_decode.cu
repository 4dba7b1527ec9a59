// PointCloud2 decode + input filters on the device: lidar::Manager::prepareInput
// (mimosa/src/lidar/manager.cpp:149-383) for every vendor layout of mimosa/include/mimosa/lidar/point.hpp:41-178,
// described by field offsets/types instead of one struct per vendor.
//
// The reference loop is sequential (points_full_ grows by emplace_back); here the per-point decisions are taken in
// parallel and the order is restored with exclusive scans, so points_full, geometric_point_idxs and the
// timestamp grouping come out exactly as the sequential loop produces them:
//   candidates i = 0, s, 2s, ...      s = create_full_res ? 1 : point_skip_divisor              (:247-250)
//   NaN / Livox-tag / intensity / range filters, time decode, t_ns > ns_max                     (:253-308)
//   points_full <- (x, y, z + z_offset, intensity, t_ns, i, sqrt(range_sq))                      (:312-313)
//   geometric idx <- new index when i % point_skip == 0 and ring % ring_skip == 0               (:317-335)
//   unique_ns = sorted distinct t_ns; every kept point is mapped to its timestamp group          (:341-371)
// transpose_pointcloud (:179-198) and organize_pointcloud_by_ring (:204-243) re-order the message before that loop;
// here they are an index map in front of the record fetch (mb_scan_from_cloud_ordered): candidate i of the re-ordered
// cloud reads record src(i) = T(perm[i]) of the message as received — T the transposition of the width x height grid,
// perm the stable ordering by ring number (a radix sort of ring keys, which is what the reference's counting sort
// computes) — so the message itself is never copied.
#include <cub/cub.cuh>

#include "mb_map.cuh"
#include "mb_scan.cuh"

namespace mb {
namespace {

struct OutRec {  // mimosa::lidar::Point, point.hpp:18-39
  float x, y, z, pad;
  float intensity;
  uint32_t t, idx;
  float range;
};
static_assert(sizeof(OutRec) == 32, "lidar::Point is 32 bytes");

template <typename T>
__device__ __forceinline__ T load_unaligned(const unsigned char* p) {
  T v;
  memcpy(&v, p, sizeof(T));
  return v;
}

// Record of the message that sits at index i of the re-ordered cloud: perm (nullable) = ring ordering over the
// (possibly transposed) cloud; t_h != 0: the cloud was transposed from t_w columns x t_h rows, so index j of the
// transposed cloud (t_h columns) is old row j % t_h, old column j / t_h (manager.cpp:192-198).
__device__ __forceinline__ size_t src_record(size_t i, const uint32_t* __restrict__ perm, uint32_t t_w, uint32_t t_h) {
  size_t j = perm ? (size_t)perm[i] : i;
  if (t_h) j = (j % t_h) * (size_t)t_w + j / t_h;
  return j;
}

// the ring number of a record: uint16 / uint8 / float32 (`++ring_counts[point.ring]`, manager.cpp:218, converts
// PointVelodyneAnybotics' float ring to an index by truncation)
__device__ __forceinline__ uint32_t load_ring(const unsigned char* p, const mb_cloud_layout& lay) {
  if (lay.ring_type == 0) return (uint32_t)load_unaligned<uint16_t>(p + lay.off_ring);
  if (lay.ring_type == 1) return (uint32_t)p[lay.off_ring];
  return (uint32_t)(int)load_unaligned<float>(p + lay.off_ring);
}

// The reference assigns a double expression to `uint32_t t_ns` (manager.cpp:285-304).  For values outside [0, 2^32)
// that conversion is what x86-64 compilers emit for it: truncate to a 64-bit integer, keep the low 32 bits — a point
// stamped BEFORE the header (negative offset) wraps to ~4.29e9 ns and is then dropped by the ns_max test (:306).
// Reproduced here (a saturating conversion would keep such points with t_ns = 0).
__device__ __forceinline__ uint32_t wrap_u32(double v) { return (uint32_t)(unsigned long long)__double2ll_rz(v); }

// ring number of every record of the (possibly transposed) cloud, as sort keys, and the identity permutation
__global__ void k_ring_keys(const unsigned char* __restrict__ data, size_t n, mb_cloud_layout lay, uint32_t t_w, uint32_t t_h,
                            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* p = data + src_record(i, nullptr, t_w, t_h) * lay.point_step;
  keys[i] = load_ring(p, lay);
  vals[i] = (uint32_t)i;
}

// Candidate c = index i of the (re-ordered) cloud, its record at p: the filters, the time decode and the output row.
__device__ __forceinline__ void decode_record(const unsigned char* p, size_t i, size_t c, const mb_cloud_layout lay,
                                              const mb_input_filter fl, float range_min_sq, float range_max_sq,
                                              OutRec* __restrict__ tmp, uint32_t* __restrict__ keep, uint32_t* __restrict__ geo) {
  uint32_t k = 0, g = 0;
  OutRec r;
  r.x = load_unaligned<float>(p + lay.off_x);
  r.y = load_unaligned<float>(p + lay.off_y);
  r.z = load_unaligned<float>(p + lay.off_z);
  r.pad = 1.f;
  r.intensity = 0.f;
  r.t = 0;
  r.idx = (uint32_t)i;
  r.range = 0.f;
  bool ok = !(isnan(r.x) || isnan(r.y) || isnan(r.z));
  if (ok && lay.off_tag >= 0) {
    const uint8_t tag = p[lay.off_tag];
    ok = (tag & 0x30) == 0x10 || (tag & 0x30) == 0x00;
  }
  if (ok) {
    float inten;
    if (lay.intensity_type == 0) {
      inten = load_unaligned<float>(p + lay.off_intensity);
      ok = !(isnan(inten) || inten < fl.intensity_min || inten > fl.intensity_max);
    } else {
      const uint16_t refl = load_unaligned<uint16_t>(p + lay.off_intensity);
      ok = !((float)refl < fl.intensity_min || (float)refl > fl.intensity_max);
      inten = (float)refl;
    }
    r.intensity = inten;
  }
  float range_sq = 0.f;
  if (ok) {
    range_sq = r.x * r.x + r.y * r.y + r.z * r.z;
    ok = !(range_sq < range_min_sq || range_sq > range_max_sq);
  }
  if (ok) {
    uint32_t t_ns;
    if (lay.time_type == 0) {
      t_ns = load_unaligned<uint32_t>(p + lay.off_time);
    } else if (lay.time_type == 1) {
      t_ns = wrap_u32((double)load_unaligned<float>(p + lay.off_time) * 1e9);
    } else if (lay.time_type == 2) {
      t_ns = wrap_u32((load_unaligned<double>(p + lay.off_time) - fl.header_ts) * 1e9);
    } else {
      t_ns = wrap_u32(load_unaligned<double>(p + lay.off_time) - fl.header_ts * 1e9);
    }
    ok = !((float)t_ns > fl.ns_max);
    r.t = t_ns;
  }
  if (ok) {
    k = 1;
    r.z = r.z + fl.z_offset;
    r.range = sqrtf(range_sq);
    bool gk = i % (size_t)fl.point_skip_divisor == 0;
    if (gk && lay.ring_filter) gk = load_ring(p, lay) % (uint32_t)fl.ring_skip_divisor == 0;
    g = gk ? 1u : 0u;
  }
  tmp[c] = r;
  keep[c] = k;
  geo[c] = g;
}

// the message as received (mb_scan_from_cloud)
__global__ void k_decode(const unsigned char* __restrict__ data, size_t n_cand, uint32_t stride_pts, mb_cloud_layout lay,
                         mb_input_filter fl, float range_min_sq, float range_max_sq, OutRec* __restrict__ tmp,
                         uint32_t* __restrict__ keep, uint32_t* __restrict__ geo) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const size_t i = c * stride_pts;
  decode_record(data + i * lay.point_step, i, c, lay, fl, range_min_sq, range_max_sq, tmp, keep, geo);
}

// the re-ordered message (mb_scan_from_cloud_ordered): index i of the re-ordered cloud reads record src_record(i)
__global__ void k_decode_ordered(const unsigned char* __restrict__ data, size_t n_cand, uint32_t stride_pts, mb_cloud_layout lay,
                                 mb_input_filter fl, float range_min_sq, float range_max_sq, const uint32_t* __restrict__ perm,
                                 uint32_t t_w, uint32_t t_h, OutRec* __restrict__ tmp, uint32_t* __restrict__ keep,
                                 uint32_t* __restrict__ geo) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const size_t i = c * stride_pts;
  decode_record(data + src_record(i, perm, t_w, t_h) * lay.point_step, i, c, lay, fl, range_min_sq, range_max_sq, tmp, keep, geo);
}

__global__ void k_compact(const OutRec* __restrict__ tmp, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ keep_pos,
                          const uint32_t* __restrict__ geo, const uint32_t* __restrict__ geo_pos, size_t n_cand,
                          OutRec* __restrict__ out, uint32_t* __restrict__ geo_idx, uint32_t* __restrict__ t_keys,
                          uint32_t* __restrict__ t_vals) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand || !keep[c]) return;
  const uint32_t pos = keep_pos[c];
  out[pos] = tmp[c];
  t_keys[pos] = tmp[c].t;
  t_vals[pos] = pos;
  if (geo[c]) geo_idx[geo_pos[c]] = pos;
}

__global__ void k_heads32(const uint32_t* __restrict__ keys, size_t n, uint32_t* __restrict__ head) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// group index of every kept point + the list of distinct timestamps
__global__ void k_groups(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ head,
                         const uint32_t* __restrict__ head_scan, size_t n, uint32_t* __restrict__ pose_index,
                         uint32_t* __restrict__ unique_ns) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t grp = head[i] ? head_scan[i] : head_scan[i] - 1;
  pose_index[vals[i]] = grp;
  if (head[i]) unique_ns[grp] = keys[i];
}

__global__ void k_last3(const uint32_t* a_flag, const uint32_t* a_scan, const uint32_t* b_flag, const uint32_t* b_scan, size_t n,
                        uint32_t* counters) {
  counters[0] = a_flag[n - 1] + a_scan[n - 1];
  counters[1] = b_flag[n - 1] + b_scan[n - 1];
}
__global__ void k_last1(const uint32_t* a_flag, const uint32_t* a_scan, const uint32_t* sorted_keys, size_t n, uint32_t* counters) {
  counters[2] = a_flag[n - 1] + a_scan[n - 1];
  counters[3] = sorted_keys[n - 1];  // largest timestamp = last_point_ns
}

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_scan_from_cloud(mb_ctx* ctx, const void* data, size_t n_points, const mb_cloud_layout* layout,
                                  const mb_input_filter* filter, mb_scan** points_full, uint32_t* geometric_idx,
                                  size_t* n_geometric, uint32_t* pose_index, uint32_t* unique_ns, size_t* n_unique,
                                  uint32_t* last_point_ns) {
  return mb_scan_from_cloud_ordered(ctx, data, n_points, layout, filter, nullptr, points_full, geometric_idx, n_geometric, pose_index,
                                    unique_ns, n_unique, last_point_ns);
}

extern "C" int mb_scan_from_cloud_ordered(mb_ctx* ctx, const void* data, size_t n_points, const mb_cloud_layout* layout,
                                          const mb_input_filter* filter, const mb_cloud_order* order, mb_scan** points_full,
                                          uint32_t* geometric_idx, size_t* n_geometric, uint32_t* pose_index,
                                          uint32_t* unique_ns, size_t* n_unique, uint32_t* last_point_ns) {
  MB_REQUIRE(ctx && layout && filter && points_full && n_geometric && n_unique, "null argument");
  uint32_t t_w = 0, t_h = 0;  // transposition: the message's width / height
  bool by_ring = false;
  if (order) {
    MB_REQUIRE((size_t)order->width * order->height == n_points, "width * height must equal the number of points");
    uint32_t height_now = order->height;  // height of the cloud the ring re-ordering would see
    if (order->transpose_pointcloud && n_points) {
      t_w = order->width;
      t_h = order->height;
      height_now = order->width;
    }
    // manager.cpp:210: only an unorganised cloud (height == 1) is re-ordered
    by_ring = order->organize_pointcloud_by_ring && height_now == 1 && n_points > 0;
    MB_REQUIRE(!by_ring || layout->off_ring >= 0, "organize_pointcloud_by_ring needs a ring field");
  }
  MB_REQUIRE(n_points == 0 || (data && geometric_idx && pose_index && unique_ns), "null buffer");
  MB_REQUIRE(layout->point_step >= 12 && filter->point_skip_divisor >= 1 && filter->ring_skip_divisor >= 1, "bad layout/filter");
  MB_REQUIRE(layout->intensity_type >= 0 && layout->intensity_type <= 1 && layout->time_type >= 0 && layout->time_type <= 3 &&
                 layout->ring_type >= 0 && layout->ring_type <= 2,
             "unknown field type");
  MB_REQUIRE(!layout->ring_filter || (layout->off_ring >= 0 && layout->ring_type != 2), "ring_filter needs an integer ring field");
  {
    // every field must lie inside a record
    auto inside = [&](int32_t off, size_t size) { return off >= 0 && (size_t)off + size <= (size_t)layout->point_step; };
    const size_t time_size = layout->time_type <= 1 ? 4 : 8, ring_size = layout->ring_type == 0 ? 2 : layout->ring_type == 1 ? 1 : 4;
    MB_REQUIRE(inside(layout->off_x, 4) && inside(layout->off_y, 4) && inside(layout->off_z, 4) &&
                   inside(layout->off_intensity, layout->intensity_type == 0 ? 4 : 2) && inside(layout->off_time, time_size) &&
                   (layout->off_ring < 0 || inside(layout->off_ring, ring_size)) && (layout->off_tag < 0 || inside(layout->off_tag, 1)),
               "a field of the layout lies outside the record (offset + size > point_step)");
  }
  MB_REQUIRE(n_points < 0x7fffffffull, "too many points");
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  *n_geometric = 0;
  *n_unique = 0;
  if (last_point_ns) *last_point_ns = 0;
  const uint32_t stride_pts = filter->create_full_res_pointcloud ? 1u : (uint32_t)filter->point_skip_divisor;
  const size_t n_cand = (n_points + stride_pts - 1) / stride_pts;
  if (n_cand == 0) {
    mb_scan* s = new mb_scan;
    s->ctx = ctx;
    s->stride = sizeof(OutRec);
    s->bytes = 256;
    MB_TRY(dev_alloc(ctx, (void**)&s->data, s->bytes));
    *points_full = s;
    return MB_OK;
  }
  size_t sort_temp = 0, scan_temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (int)n_cand, 0, 32, st);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_cand, st);
  const size_t temp_bytes = std::max(sort_temp, scan_temp);
  const size_t raw_bytes = n_points * layout->point_step;
  size_t ring_temp = 0;
  if (by_ring)
    cub::DeviceRadixSort::SortPairs(nullptr, ring_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (int)n_points, 0, 16, st);
  const size_t bytes = raw_bytes + n_cand * (sizeof(OutRec) + 4 * 12) + temp_bytes + (by_ring ? n_points * 16 + ring_temp : 0) + 256 * 26;
  unsigned char* scratch = nullptr;
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
  size_t off = 0;
  auto take = [&](size_t b) {
    off = (off + 255) & ~(size_t)255;
    unsigned char* p = scratch + off;
    off += b;
    return p;
  };
  unsigned char* raw = take(raw_bytes);
  OutRec* tmp = (OutRec*)take(n_cand * sizeof(OutRec));
  uint32_t* keep = (uint32_t*)take(n_cand * 4);
  uint32_t* keep_pos = (uint32_t*)take(n_cand * 4);
  uint32_t* geo = (uint32_t*)take(n_cand * 4);
  uint32_t* geo_pos = (uint32_t*)take(n_cand * 4);
  uint32_t* geo_idx = (uint32_t*)take(n_cand * 4);
  uint32_t* t_keys = (uint32_t*)take(n_cand * 4);
  uint32_t* t_vals = (uint32_t*)take(n_cand * 4);
  uint32_t* t_keys_s = (uint32_t*)take(n_cand * 4);
  uint32_t* t_vals_s = (uint32_t*)take(n_cand * 4);
  uint32_t* head = (uint32_t*)take(n_cand * 4);
  uint32_t* head_scan = (uint32_t*)take(n_cand * 4);
  uint32_t* d_pose_index = (uint32_t*)take(n_cand * 4);
  uint32_t* counters = (uint32_t*)take(16);
  void* temp = take(temp_bytes);
  uint32_t *ring_keys = nullptr, *ring_vals = nullptr, *ring_keys_s = nullptr, *perm = nullptr;
  void* ring_tmp = nullptr;
  if (by_ring) {
    ring_keys = (uint32_t*)take(n_points * 4);
    ring_vals = (uint32_t*)take(n_points * 4);
    ring_keys_s = (uint32_t*)take(n_points * 4);
    perm = (uint32_t*)take(n_points * 4);
    ring_tmp = take(ring_temp);
  }
  // keep_pos .. d_pose_index reuse: unique_ns goes into t_keys after the sort (no longer needed)

  MB_TRY(pinned_reserve(ctx, raw_bytes));
  std::memcpy(ctx->pinned, data, raw_bytes);
  MB_CUDA(cudaMemcpyAsync(raw, ctx->pinned, raw_bytes, cudaMemcpyHostToDevice, st));
  const float rmin2 = filter->range_min * filter->range_min, rmax2 = filter->range_max * filter->range_max;
  if (by_ring) {
    // stable order by ring number = the reference's counting sort (ring counts -> offsets -> in-order placement,
    // manager.cpp:216-239); 16 key bits cover both ring types
    k_ring_keys<<<blocks_for(n_points, 256), 256, 0, st>>>(raw, n_points, *layout, t_w, t_h, ring_keys, ring_vals);
    size_t rb = ring_temp;
    MB_CUDA(cub::DeviceRadixSort::SortPairs(ring_tmp, rb, ring_keys, ring_keys_s, ring_vals, perm, (int)n_points, 0, 16, st));
    ctx->launches += 4;
  }
  if (perm || t_h)
    k_decode_ordered<<<blocks_for(n_cand, 256), 256, 0, st>>>(raw, n_cand, stride_pts, *layout, *filter, rmin2, rmax2, perm, t_w, t_h,
                                                              tmp, keep, geo);
  else
    k_decode<<<blocks_for(n_cand, 256), 256, 0, st>>>(raw, n_cand, stride_pts, *layout, *filter, rmin2, rmax2, tmp, keep, geo);
  size_t tb = temp_bytes;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, keep, keep_pos, (int)n_cand, st));
  tb = temp_bytes;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, geo, geo_pos, (int)n_cand, st));
  k_last3<<<1, 1, 0, st>>>(keep, keep_pos, geo, geo_pos, n_cand, counters);
  uint32_t h[4] = {0, 0, 0, 0};
  MB_CUDA(cudaMemcpyAsync(h, counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  const size_t n_keep = h[0], n_geo = h[1];
  mb_scan* s = new mb_scan;
  s->ctx = ctx;
  s->n = n_keep;
  s->stride = sizeof(OutRec);
  s->bytes = std::max<size_t>(n_keep * sizeof(OutRec), 256);
  int rc = dev_alloc(ctx, (void**)&s->data, s->bytes);
  if (rc != MB_OK) {
    delete s;
    return rc;
  }
  ctx->launches += 6;
  if (n_keep) {
    k_compact<<<blocks_for(n_cand, 256), 256, 0, st>>>(tmp, keep, keep_pos, geo, geo_pos, n_cand, (OutRec*)s->data, geo_idx, t_keys,
                                                       t_vals);
    tb = temp_bytes;
    MB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, t_keys, t_keys_s, t_vals, t_vals_s, (int)n_keep, 0, 32, st));
    k_heads32<<<blocks_for(n_keep, 256), 256, 0, st>>>(t_keys_s, n_keep, head);
    tb = temp_bytes;
    MB_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, head, head_scan, (int)n_keep, st));
    k_groups<<<blocks_for(n_keep, 256), 256, 0, st>>>(t_keys_s, t_vals_s, head, head_scan, n_keep, d_pose_index, t_keys);
    k_last1<<<1, 1, 0, st>>>(head, head_scan, t_keys_s, n_keep, counters);
    ctx->launches += 8;
    MB_CUDA(cudaMemcpyAsync(h + 2, counters + 2, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaMemcpyAsync(pose_index, d_pose_index, n_keep * 4, cudaMemcpyDeviceToHost, st));
    if (n_geo) MB_CUDA(cudaMemcpyAsync(geometric_idx, geo_idx, n_geo * 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    MB_CUDA(cudaMemcpyAsync(unique_ns, t_keys, (size_t)h[2] * 4, cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    *n_unique = h[2];
    if (last_point_ns) *last_point_ns = h[3];
  }
  *n_geometric = n_geo;
  MB_CUDA(cudaGetLastError());
  *points_full = s;
  return MB_OK;
}
