// Scan preparation on the device: the steps that sit immediately before and after the ICP factor in
// mimosa's LiDAR callback.
//   mb_scan_deskew      lidar::Manager::deskewPoints, per-point part  (mimosa/src/lidar/manager.cpp:494-509):
//                       p <- R_Le_Lt * p + t_Le_Lt in FLOAT, one pose per unique timestamp; the IMU propagation
//                       that produces the pose table (:459-492) stays on the host
//   mb_scan_transform   Geometric::preprocess, T_B_L in float        (mimosa/src/lidar/geometric.cpp:153-160)
//   mb_scan_downsample  Geometric::downsample                         (geometric.cpp:55-126)
//   mb_map_insert_scan  Geometric::updateMap: float world transform of the FULL body-frame cloud, then insert
//                       into the (freshly snapshotted) map             (geometric.cpp:483-495)
#include "mb_map.cuh"
#include "mb_scan.cuh"

namespace mb {
namespace {

// Float 3x3 * 3 + 3 with the reference's reduction order (fixed-size Eigen: a0 + (a1 + a2)); -fmad=false keeps
// every operation a separate IEEE binary32 op like the CPU build.
__device__ __forceinline__ void xform_f32(const float* __restrict__ R, const float* __restrict__ t, float& x, float& y, float& z) {
  const float px = x, py = y, pz = z;
  x = (R[0] * px + (R[1] * py + R[2] * pz)) + t[0];
  y = (R[3] * px + (R[4] * py + R[5] * pz)) + t[1];
  z = (R[6] * px + (R[7] * py + R[8] * pz)) + t[2];
}

// poses: n_poses x 12 floats (R row-major, t); pose_index == nullptr -> pose 0 for every point.
__global__ void k_transform_f32(unsigned char* __restrict__ data, size_t n, size_t stride,
                                const uint32_t* __restrict__ pose_index, const float* __restrict__ poses) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* f = (float*)(data + i * stride);
  const float* P = poses + (pose_index ? (size_t)pose_index[i] * 12 : 0);
  float x = f[0], y = f[1], z = f[2];
  xform_f32(P, P + 9, x, y, z);
  f[0] = x;
  f[1] = y;
  f[2] = z;
}

__global__ void k_gather_records(const unsigned char* __restrict__ src, size_t stride, const uint32_t* __restrict__ idx,
                                 size_t n_out, unsigned char* __restrict__ dst) {
  const size_t words = stride / 4;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out * words) return;
  const size_t r = t / words, w = t % words;
  ((uint32_t*)dst)[r * words + w] = ((const uint32_t*)src)[(size_t)idx[r] * words + w];
}

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

int scan_alloc(mb_ctx* ctx, size_t n, size_t stride, mb_scan** out) {
  mb_scan* s = new mb_scan;
  s->ctx = ctx;
  s->n = n;
  s->stride = stride;
  s->bytes = std::max<size_t>(n * stride, 256);
  const int rc = dev_alloc(ctx, (void**)&s->data, s->bytes);
  if (rc != MB_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return MB_OK;
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" {

int mb_scan_upload(mb_ctx* ctx, const void* pts, size_t n, size_t stride_bytes, mb_scan** out) {
  MB_REQUIRE(ctx && out && (n == 0 || pts), "null argument");
  MB_REQUIRE(stride_bytes >= 12 && stride_bytes % 4 == 0, "stride must be >= 12 and a multiple of 4");
  MB_CUDA(cudaSetDevice(ctx->device));
  mb_scan* s = nullptr;
  MB_TRY(scan_alloc(ctx, n, stride_bytes, &s));
  if (n && host_is_page_locked(pts)) {
    // page-locked caller buffer (mb_host_register): DMA straight from it, no CPU staging pass
    MB_CUDA(cudaMemcpyAsync(s->data, pts, n * stride_bytes, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));  // the caller's buffer has been consumed when this returns
  } else if (n) {
    int rc = pinned_reserve(ctx, n * stride_bytes);
    if (rc != MB_OK) {
      mb_scan_release(s);
      return rc;
    }
    std::memcpy(ctx->pinned, pts, n * stride_bytes);
    MB_CUDA(cudaMemcpyAsync(s->data, ctx->pinned, n * stride_bytes, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  *out = s;
  return MB_OK;
}

int mb_scan_release(mb_scan* s) {
  if (!s) return MB_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  dev_free(s->ctx, s->data, s->bytes);
  delete s;
  return MB_OK;
}

int mb_scan_size(mb_scan* s, size_t* n, size_t* stride_bytes) {
  MB_REQUIRE(s, "null scan");
  if (n) *n = s->n;
  if (stride_bytes) *stride_bytes = s->stride;
  return MB_OK;
}

int mb_scan_download(mb_scan* s, void* pts) {
  MB_REQUIRE(s && (s->n == 0 || pts), "null argument");
  MB_CUDA(cudaSetDevice(s->ctx->device));
  if (s->n) {
    MB_CUDA(cudaMemcpyAsync(pts, s->data, s->n * s->stride, cudaMemcpyDeviceToHost, s->ctx->stream));
    MB_CUDA(cudaStreamSynchronize(s->ctx->stream));
  }
  return MB_OK;
}

static int transform_common(mb_scan* s, const uint32_t* pose_index, const float* poses, size_t n_poses) {
  mb_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  if (s->n == 0) return MB_OK;
  const size_t idx_bytes = pose_index ? s->n * sizeof(uint32_t) : 0;
  const size_t pose_bytes = n_poses * 12 * sizeof(float);
  void* d_tmp = nullptr;
  const size_t tmp_bytes = ((idx_bytes + 255) & ~(size_t)255) + pose_bytes;
  MB_TRY(dev_alloc(c, &d_tmp, tmp_bytes));
  uint32_t* d_idx = pose_index ? (uint32_t*)d_tmp : nullptr;
  float* d_poses = (float*)((char*)d_tmp + ((idx_bytes + 255) & ~(size_t)255));
  int rc = pinned_reserve(c, tmp_bytes);
  if (rc == MB_OK) {
    char* h = (char*)c->pinned;
    if (pose_index) std::memcpy(h, pose_index, idx_bytes);
    std::memcpy(h + ((idx_bytes + 255) & ~(size_t)255), poses, pose_bytes);
    cudaError_t e = cudaMemcpyAsync(d_tmp, h, tmp_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      k_transform_f32<<<blocks_for(s->n, 256), 256, 0, st>>>(s->data, s->n, s->stride, d_idx, d_poses);
      ++c->launches;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("scan transform failed: %s", cudaGetErrorString(e));
      rc = MB_ERR_CUDA;
    }
  }
  dev_free(c, d_tmp, tmp_bytes);
  return rc;
}

int mb_scan_deskew(mb_scan* s, const uint32_t* pose_index, const float* poses, size_t n_poses) {
  MB_REQUIRE(s && poses && n_poses > 0 && (s->n == 0 || pose_index), "null argument");
  for (size_t i = 0; i < s->n; ++i) MB_REQUIRE(pose_index[i] < n_poses, "pose index out of range");
  MB_CUDA(cudaSetDevice(s->ctx->device));
  return transform_common(s, pose_index, poses, n_poses);
}

int mb_scan_transform(mb_scan* s, const float R[9], const float t[3]) {
  MB_REQUIRE(s && R && t, "null argument");
  MB_CUDA(cudaSetDevice(s->ctx->device));
  float P[12];
  std::memcpy(P, R, 9 * sizeof(float));
  std::memcpy(P + 9, t, 3 * sizeof(float));
  return transform_common(s, nullptr, P, 1);
}

int mb_scan_downsample(mb_scan* s, float leaf, size_t cap, float min_dist, mb_scan** out) {
  MB_REQUIRE(s && out, "null argument");
  mb_ctx* c = s->ctx;
  MB_CUDA(cudaSetDevice(c->device));
  uint32_t* d_idx = nullptr;
  const size_t idx_bytes = std::max<size_t>(s->n, 1) * sizeof(uint32_t);
  MB_TRY(dev_alloc(c, (void**)&d_idx, idx_bytes));
  size_t kept = 0;
  int rc = downsample_impl(c, s->data, cudaMemcpyDeviceToDevice, s->n, s->stride, leaf, cap, min_dist, nullptr, d_idx, &kept);
  mb_scan* o = nullptr;
  if (rc == MB_OK) rc = scan_alloc(c, kept, s->stride, &o);
  if (rc == MB_OK && kept) {
    k_gather_records<<<blocks_for(kept * (s->stride / 4), 256), 256, 0, c->stream>>>(s->data, s->stride, d_idx, kept, o->data);
    ++c->launches;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
      set_error("scan gather failed: %s", cudaGetErrorString(e));
      rc = MB_ERR_CUDA;
    }
  }
  dev_free(c, d_idx, idx_bytes);
  if (rc != MB_OK) {
    mb_scan_release(o);
    return rc;
  }
  *out = o;
  return MB_OK;
}

int mb_scan_gather(mb_scan* s, const uint32_t* idx, size_t n, mb_scan** out) {
  MB_REQUIRE(s && out && (n == 0 || idx), "null argument");
  for (size_t i = 0; i < n; ++i) MB_REQUIRE(idx[i] < s->n, "index out of range");
  mb_ctx* c = s->ctx;
  MB_CUDA(cudaSetDevice(c->device));
  mb_scan* o = nullptr;
  MB_TRY(scan_alloc(c, n, s->stride, &o));
  if (n) {
    uint32_t* d_idx = nullptr;
    int rc = dev_alloc(c, (void**)&d_idx, n * sizeof(uint32_t));
    if (rc == MB_OK) rc = pinned_reserve(c, n * sizeof(uint32_t));
    if (rc == MB_OK) {
      std::memcpy(c->pinned, idx, n * sizeof(uint32_t));
      cudaError_t e = cudaMemcpyAsync(d_idx, c->pinned, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) {
        k_gather_records<<<blocks_for(n * (s->stride / 4), 256), 256, 0, c->stream>>>(s->data, s->stride, d_idx, n, o->data);
        ++c->launches;
        e = cudaGetLastError();
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) {
        set_error("scan gather failed: %s", cudaGetErrorString(e));
        rc = MB_ERR_CUDA;
      }
    }
    dev_free(c, d_idx, n * sizeof(uint32_t));
    if (rc != MB_OK) {
      mb_scan_release(o);
      return rc;
    }
  }
  *out = o;
  return MB_OK;
}

int mb_map_insert_scan(mb_map* map, mb_scan* s, const float R[9], const float t[3]) {
  MB_REQUIRE(map && s && R && t, "null argument");
  MB_REQUIRE(map->ctx == s->ctx, "map and scan belong to different contexts");
  mb_ctx* c = s->ctx;
  MB_CUDA(cudaSetDevice(c->device));
  // W_points = R_W_Be(f32) * p + t_W_Be(f32) on a temporary copy (the body-frame scan is kept)
  mb_scan* w = nullptr;
  MB_TRY(scan_alloc(c, s->n, s->stride, &w));
  int rc = MB_OK;
  if (s->n) {
    cudaError_t e = cudaMemcpyAsync(w->data, s->data, s->n * s->stride, cudaMemcpyDeviceToDevice, c->stream);
    if (e != cudaSuccess) {
      set_error("scan copy failed: %s", cudaGetErrorString(e));
      rc = MB_ERR_CUDA;
    }
    if (rc == MB_OK) rc = mb_scan_transform(w, R, t);
  }
  if (rc == MB_OK) rc = map_insert_impl(map, w->data, cudaMemcpyDeviceToDevice, w->n, w->stride);
  mb_scan_release(w);
  return rc;
}

}  // extern "C"
