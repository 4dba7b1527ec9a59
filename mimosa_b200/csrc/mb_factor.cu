// Point-to-plane scan-to-map ICP factor on the device: the body of mimosa::lidar::ICPFactor::linearize
// (mimosa/include/mimosa/lidar/geometric_factor.hpp:231-562) and estimatePlane (:176-229).
//
//   k_linearize   transform + data-association gate (:276-287), warp-cooperative restricted k-NN
//                 (:294, via mb_internal.cuh::knn_warp), gates (:296-302), plane fit (:176-229), residual,
//                 s-check, Huber, Jacobian (:319-355), per-point localizability vectors (:351-352) and the
//                 J^T J / J^T e / e^2 / status-count reduction (:364-366, 396-403) — ONE kernel, the
//                 neighbour points never leave the SM.
//   k_finalize    6x6 assembly, localizability eigen-decompositions, Schur complements, 4-DoF projection
//                 (:405-428, 464-475) and — for the stand-alone Gauss-Newton harness — the LDL^T solve and
//                 SE(3) retract that ISAM2 performs in the reference (mimosa/src/graph/manager.cpp:585-588).
//   k_loc_comp    the component-localizability second pass over valid points (:434-457).
// With more than one rank the 40-double packet between k_linearize and k_finalize (and the 6 doubles after
// k_loc_comp) are all-reduced with NCCL; every rank then computes the identical step.
//
// Per-point state (status, DA anchor, plane mean/normal, localizability vectors) lives in HBM inside the
// factor handle as structure-of-arrays, mirroring the `mutable` vectors at geometric_factor.hpp:79-106.
#include <algorithm>
#include <vector>

#include "mb_map.cuh"

namespace mb {
namespace {

constexpr int kLinWarps = 4;
constexpr int kPack = 40;  // 21 H upper-tri, 6 J^T e, 1 f, 9 status counts, 1 n_searched, 2 pad
constexpr int kPackH = 0, kPackB = 21, kPackF = 27, kPackCnt = 28, kPackSearched = 37;
static_assert(kPackF == 27 && kPackB == 21, "packing matches the lane ownership in k_linearize");

struct FactorView {
  const float4* src;
  uint8_t* status;
  double *p_da, *mean, *normal, *loc_rot, *loc_trans;  // SoA: component c of point i at [c * ld + i]
  uint64_t* knn_idx;                                   // [i * k + j]
  size_t n, ld;
  int k, use_huber;
  uint32_t flags;
  double da_gate, max_corr_sq, sigma, kh, pvd;
  double* partials;  // [grid][kPack]
  unsigned* ticket;
  double* packed;    // [kPack]
  double* partials2; // [grid2][8]
  unsigned* ticket2;
  double* loc_out;   // [8]: trans comp (3), rot comp (3)
};

struct DevState {          // small device-resident block per factor
  double pose[12];         // R row-major (9), t (3)
  double gravity[3];
  double lambda;
  mb_linearization lin;    // result of the most recent linearisation (loc_*_comp filled on the host)
};

__device__ __forceinline__ d3 ld3(const double* base, size_t ld, size_t i) {
  return mk3(base[i], base[ld + i], base[2 * ld + i]);
}
__device__ __forceinline__ void st3(double* base, size_t ld, size_t i, d3 v) {
  base[i] = v.x;
  base[ld + i] = v.y;
  base[2 * ld + i] = v.z;
}

// Plane through the k neighbours (geometric_factor.hpp:176-229).  Returns the new status
// (MB_UNPROCESSED = all gates passed); mean is always written, normal once the eigen gates pass.
__device__ __forceinline__ uint8_t fit_plane(const float4* nb, int k, d3 origin, double pvd, d3& mean, d3& normal,
                                             bool& normal_set) {
  d3 s = mk3(0, 0, 0);
  for (int j = 0; j < k; ++j) s = add3(s, mk3((double)nb[j].x, (double)nb[j].y, (double)nb[j].z));
  mean = div3(s, (double)k);
  double c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
  for (int j = 0; j < k; ++j) {
    const d3 c = sub3(mk3((double)nb[j].x, (double)nb[j].y, (double)nb[j].z), mean);
    c00 += c.x * c.x;
    c10 += c.y * c.x;
    c11 += c.y * c.y;
    c20 += c.z * c.x;
    c21 += c.z * c.y;
    c22 += c.z * c.z;
  }
  const double dn = (double)(k - 1);
  m33 cov;
  cov.m[0] = c00 / dn;
  cov.m[3] = c10 / dn;
  cov.m[4] = c11 / dn;
  cov.m[6] = c20 / dn;
  cov.m[7] = c21 / dn;
  cov.m[8] = c22 / dn;
  cov.m[1] = cov.m[3];
  cov.m[2] = cov.m[6];
  cov.m[5] = cov.m[7];
  double lam[3];
  m33 V;
  normal_set = false;
  if (!eigh33(cov, lam, V)) return MB_EIGEN_SOLVER_FAIL;
  if (lam[0] < 1e-6) return MB_MIN_EIGEN_VALUE_LOW;
  if (lam[2] > 3 * lam[1]) return MB_LINE;
  d3 nrm = mk3(V.m[0], V.m[3], V.m[6]);
  if (dot3(nrm, sub3(origin, mean)) < 0) nrm = mk3(-nrm.x, -nrm.y, -nrm.z);
  normal = nrm;
  normal_set = true;
  bool invalid = false;
  for (int j = 0; j < k; ++j) {
    const d3 c = sub3(mk3((double)nb[j].x, (double)nb[j].y, (double)nb[j].z), mean);
    if (fabs(dot3(c, nrm)) > pvd) invalid = true;
  }
  return invalid ? MB_CORRES_PLANE_INVALID : MB_UNPROCESSED;
}

__global__ void __launch_bounds__(kLinWarps * 32)
    k_linearize(MapView mv, FactorView fv, const double* __restrict__ pose) {
  __shared__ int8_t s_off[32 * 3];
  __shared__ uint32_t s_vox_all[kLinWarps][32];
  __shared__ float4 s_nb_all[kLinWarps][32][MB_MAX_K];
  __shared__ double s_row_all[kLinWarps][32][7];  // per point of the tile: whitened [J (6), e]
  __shared__ double s_red[kLinWarps][kPack];
  __shared__ bool s_last;
  if (threadIdx.x < kMaxNbr * 3) s_off[threadIdx.x] = mv.off[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* s_vox = s_vox_all[warp];
  float4(*s_nb)[MB_MAX_K] = s_nb_all[warp];
  double(*s_row)[7] = s_row_all[warp];
  // Lane a < 28 owns entry a of the packed upper triangle of [J e]^T [J e] (7x7): the 21 entries of
  // J^T J first (row-major, c >= r), then J^T e (6), then e^2.
  int pr = 0, pc = 0;
  {
    int u = 0;
    for (int r = 0; r < 6; ++r)
      for (int c = r; c < 6; ++c) {
        if (u == lane) {
          pr = r;
          pc = c;
        }
        ++u;
      }
    if (lane >= 21 && lane < 27) {
      pr = lane - 21;
      pc = 6;
    }
    if (lane == 27) pr = pc = 6;
  }

  m33 R;
#pragma unroll
  for (int a = 0; a < 9; ++a) R.m[a] = pose[a];
  const d3 T = mk3(pose[9], pose[10], pose[11]);
  const int k = fv.k;
  const bool forced = (fv.flags & 1u) != 0;

  double acc = 0.0;
  int cnt = 0;  // lane s < 9: points with status s; lane 9: points searched

  const size_t n_tiles = (fv.n + 31) / 32;
  const size_t warps_total = (size_t)gridDim.x * kLinWarps;
  for (size_t tile = (size_t)blockIdx.x * kLinWarps + warp; tile < n_tiles; tile += warps_total) {
    const size_t i = tile * 32 + lane;
    const bool act = i < fv.n;
    d3 ps = mk3(0, 0, 0), pt = mk3(0, 0, 0);
    uint8_t st = MB_UNPROCESSED;
    bool need = false;
    if (act) {
      const float4 s = __ldg(fv.src + i);
      ps = mk3((double)s.x, (double)s.y, (double)s.z);
      pt = add3(mul33v(R, ps), T);
      st = fv.status[i];
      const d3 da = ld3(fv.p_da, fv.ld, i);
      need = forced || sqrt(sqnorm3(sub3(pt, da))) > fv.da_gate;
    }
    const unsigned mask = __ballot_sync(kFull, need);
    int my_found = 0;
    double my_dk = 0.0;
    for (unsigned m = mask; m; m &= m - 1) {
      const int src_lane = __ffs(m) - 1;
      const double qx = __shfl_sync(kFull, pt.x, src_lane), qy = __shfl_sync(kFull, pt.y, src_lane),
                   qz = __shfl_sync(kFull, pt.z, src_lane);
      KnnOut o;
      knn_warp(mv, s_off, s_vox, qx, qy, qz, k, lane, o);
      if (lane < k) {
        uint64_t g = ~0ull;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o.seq != 0xffffffffu) g = knn_fetch(mv, s_vox, o.seq, p);
        s_nb[src_lane][lane] = p;
        if (fv.knn_idx) fv.knn_idx[(tile * 32 + src_lane) * k + lane] = o.found == k ? g : ~0ull;
      }
      const double dk = __shfl_sync(kFull, o.d2, k - 1);
      if (lane == src_lane) {
        my_found = o.found;
        my_dk = dk;
      }
      __syncwarp();
    }

    double row[7] = {0, 0, 0, 0, 0, 0, 0};
    if (act) {
      bool proceed = false;
      d3 mean = mk3(0, 0, 0), normal = mk3(0, 0, 0);
      if (need) {
        st3(fv.p_da, fv.ld, i, pt);
        st = MB_UNPROCESSED;
        if (my_found != k) {
          st = MB_INSUFFICIENT_CORRES_POINTS;
        } else if (my_dk > fv.max_corr_sq) {
          st = MB_CORRES_MAX_DIST;
        } else {
          bool normal_set;
          st = fit_plane(s_nb[lane], k, T, fv.pvd, mean, normal, normal_set);
          st3(fv.mean, fv.ld, i, mean);
          if (normal_set) st3(fv.normal, fv.ld, i, normal);
          proceed = st == MB_UNPROCESSED;
        }
      } else if (st > MB_CORRES_PLANE_INVALID) {
        mean = ld3(fv.mean, fv.ld, i);
        normal = ld3(fv.normal, fv.ld, i);
        proceed = true;
      }
      if (proceed) {
        double e = dot3(normal, sub3(mean, pt));
        const double s_chk = 1 - 0.9 * fabs(e) / sqrt(sqrt(sqnorm3(ps)));
        if (s_chk < 0.9) {
          st = MB_MAX_ERROR;
        } else {
          double sqrt_w = 1.0;
          if (fv.use_huber) {
            const double we = e / fv.sigma;
            if (fabs(we) > fv.kh) sqrt_w = sqrt(fv.kh / fabs(we));
          }
          const double scale = sqrt_w / fv.sigma;
          e *= scale;
          const d3 ns = mul33Tv(R, normal);
          const d3 jr = cross3(ns, ps);
          const double z = sqnorm3(jr);
          st3(fv.loc_rot, fv.ld, i, z > 0 ? div3(jr, sqrt(z)) : jr);
          st3(fv.loc_trans, fv.ld, i, mk3(-ns.x, -ns.y, -ns.z));
          row[0] = jr.x * scale;
          row[1] = jr.y * scale;
          row[2] = jr.z * scale;
          row[3] = -ns.x * scale;
          row[4] = -ns.y * scale;
          row[5] = -ns.z * scale;
          row[6] = e;
          st = MB_VALID;
        }
      }
      fv.status[i] = st;
    }
    // [J e]^T [J e] over the tile: lane a sums its entry over the 32 rows in point order.
    const unsigned any_valid = __ballot_sync(kFull, act && st == MB_VALID);
    if (any_valid) {
#pragma unroll
      for (int a = 0; a < 7; ++a) s_row[lane][a] = row[a];
      __syncwarp();
#pragma unroll 8
      for (int p = 0; p < 32; ++p) acc += s_row[p][pr] * s_row[p][pc];
      __syncwarp();
    }
#pragma unroll
    for (int s = 0; s < 9; ++s) {
      const int c = __popc(__ballot_sync(kFull, act && st == s));
      if (lane == s) cnt += c;
    }
    if (lane == 9) cnt += __popc(mask);
  }

  // warp -> block -> grid reduction (fixed order, deterministic for a given launch shape)
  if (lane < 28) s_red[warp][lane] = acc;
  if (lane >= 30) s_red[warp][lane + 8] = 0.0;  // pad entries 38, 39
  if (lane < 10) s_red[warp][kPackCnt + lane] = (double)cnt;
  __syncthreads();
  if (threadIdx.x < kPack) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kLinWarps; ++w) v += s_red[w][threadIdx.x];
    fv.partials[(size_t)blockIdx.x * kPack + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(fv.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x < kPack) {
      double v = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(fv.partials + (size_t)b * kPack + threadIdx.x);
      fv.packed[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *fv.ticket = 0u;
  }
}

__device__ __forceinline__ void localizability(const m33& JtJ, double loc[3], m33& V) {
  double lam[3];
  eigh33(JtJ, lam, V);
  for (int a = 0; a < 3; ++a) loc[a] = sqrt(lam[a]);
}

// Single thread: everything after the per-point loop of ICPFactor::linearize, plus the harness GN step.
__global__ void k_finalize(const double* __restrict__ packed, DevState* ds, int reg_4_dof, int linearize_count,
                           int do_step, int iter, mb_icp_trace* trace) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double H[36], b[6];
  {
    int u = 0;
    for (int a = 0; a < 6; ++a)
      for (int c = a; c < 6; ++c) {
        H[6 * a + c] = packed[kPackH + u];
        H[6 * c + a] = packed[kPackH + u];
        ++u;
      }
  }
  for (int a = 0; a < 6; ++a) b[a] = packed[kPackB + a];
  const double f = packed[kPackF];
  mb_linearization& L = ds->lin;
  m33 Hrr, Hrt, Htr, Htt;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      Hrr.m[3 * r + c] = H[6 * r + c];
      Hrt.m[3 * r + c] = H[6 * r + 3 + c];
      Htr.m[3 * r + c] = H[6 * (r + 3) + c];
      Htt.m[3 * r + c] = H[6 * (r + 3) + 3 + c];
    }
  m33 Vr, Vt, Dr, Dt;
  localizability(Hrr, L.loc_rot_final, Vr);
  localizability(Htt, L.loc_trans_final, Vt);
  const m33 Srr = inv33(sub33(Hrr, mul33(mul33(Hrt, inv33(Htt)), Htr)));
  const m33 Stt = inv33(sub33(Htt, mul33(mul33(Htr, inv33(Hrr)), Hrt)));
  localizability(Srr, L.degen_rot, Dr);
  localizability(Stt, L.degen_trans, Dt);
  for (int a = 0; a < 3; ++a) L.degen_rot[a] = L.degen_rot[a] * 57.29578;

  m33 R;
  for (int a = 0; a < 9; ++a) R.m[a] = ds->pose[a];
  d3 T = mk3(ds->pose[9], ds->pose[10], ds->pose[11]);
  if (reg_4_dof) {
    const d3 gz = mk3(-ds->gravity[0], -ds->gravity[1], -ds->gravity[2]);
    const d3 lz = mul33Tv(R, gz);
    const double l[3] = {lz.x, lz.y, lz.z};
    m33 P;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) P.m[3 * r + c] = l[r] * l[c];
    const m33 nrr = mul33(mul33(P, Hrr), P);
    const m33 nrt = mul33(P, Hrt);
    const m33 ntr = mul33(Htr, P);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        H[6 * r + c] = nrr.m[3 * r + c];
        H[6 * r + 3 + c] = nrt.m[3 * r + c];
        H[6 * (r + 3) + c] = ntr.m[3 * r + c];
      }
    const d3 pb = mul33v(P, mk3(b[0], b[1], b[2]));
    b[0] = pb.x;
    b[1] = pb.y;
    b[2] = pb.z;
  }
  for (int a = 0; a < 36; ++a) L.H[a] = H[a];
  for (int a = 0; a < 6; ++a) L.g[a] = -b[a];
  L.f = f;
  for (int a = 0; a < 9; ++a) L.counts[a] = (int64_t)packed[kPackCnt + a];
  for (int a = 0; a < 9; ++a) {
    L.eigvec_rot[a] = Vr.m[a];
    L.eigvec_trans[a] = Vt.m[a];
    L.degen_eigvec_rot[a] = Dr.m[a];
    L.degen_eigvec_trans[a] = Dt.m[a];
  }
  L.linearize_count = linearize_count;
  L.n_searched = (int32_t)packed[kPackSearched];

  if (do_step) {
    double delta[6] = {0, 0, 0, 0, 0, 0};
    const bool ok = solve6_ldlt(L.H, ds->lambda, L.g, delta);
    if (ok) {
      se3_retract(R, T, delta);
      for (int a = 0; a < 9; ++a) ds->pose[a] = R.m[a];
      ds->pose[9] = T.x;
      ds->pose[10] = T.y;
      ds->pose[11] = T.z;
    }
    if (trace) {
      mb_icp_trace& tr = trace[iter];
      for (int a = 0; a < 36; ++a) tr.H[a] = L.H[a];
      for (int a = 0; a < 6; ++a) {
        tr.g[a] = L.g[a];
        tr.delta[a] = delta[a];
      }
      tr.f = L.f;
      for (int a = 0; a < 9; ++a) {
        tr.R[a] = R.m[a];
        tr.counts[a] = L.counts[a];
      }
      tr.t[0] = T.x;
      tr.t[1] = T.y;
      tr.t[2] = T.z;
      tr.n_searched = L.n_searched;
      tr.solve_ok = ok ? 1 : 0;
    }
  }
}

// Component localizabilities (geometric_factor.hpp:434-457): sum over Valid points of |loc_i^T V| with
// entries below 0.5 zeroed.
__global__ void __launch_bounds__(256) k_loc_comp(FactorView fv, const DevState* __restrict__ ds) {
  __shared__ double s_red[8][6];
  __shared__ bool s_last;
  m33 Vr, Vt;
#pragma unroll
  for (int a = 0; a < 9; ++a) {
    Vr.m[a] = ds->lin.eigvec_rot[a];
    Vt.m[a] = ds->lin.eigvec_trans[a];
  }
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < fv.n; i += stride) {
    if (fv.status[i] != MB_VALID) continue;
    const d3 tc = mul33Tv(Vt, ld3(fv.loc_trans, fv.ld, i));
    const d3 rc = mul33Tv(Vr, ld3(fv.loc_rot, fv.ld, i));
    const double v[6] = {fabs(tc.x), fabs(tc.y), fabs(tc.z), fabs(rc.x), fabs(rc.y), fabs(rc.z)};
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[a] += v[a] >= 0.5 ? v[a] : 0.0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    double v = acc[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0) s_red[warp][a] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
    fv.partials2[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(fv.ticket2, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (threadIdx.x < 6) {
      double v = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(fv.partials2 + (size_t)b * 8 + threadIdx.x);
      fv.loc_out[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *fv.ticket2 = 0u;
  }
}

}  // namespace
}  // namespace mb

using namespace mb;

struct mb_factor {
  mb_ctx* ctx = nullptr;
  mb_map* map = nullptr;
  mb_icp_config cfg{};
  size_t n = 0, ld = 0, n_total = 0, begin = 0;
  float4* src = nullptr;
  uint8_t* status = nullptr;
  double* vecs = nullptr;  // 5 SoA blocks of 3*ld doubles: p_da, mean, normal, loc_rot, loc_trans
  uint64_t* knn_idx = nullptr;
  double* partials = nullptr;
  double* partials2 = nullptr;
  double* packed = nullptr;   // kPack + 8 (loc_out)
  unsigned* tickets = nullptr;
  DevState* ds = nullptr;
  mb_icp_trace* d_trace = nullptr;
  int trace_cap = 0;
  int grid = 0, grid2 = 0;
  int linearize_count = 0;
  uint32_t flags = 0;
  // cached CUDA graph of an mb_icp_run sequence
  cudaGraphExec_t graph = nullptr;
  int graph_iters = 0, graph_count0 = -1;
  bool graph_trace = false;

  FactorView view() const {
    FactorView v;
    v.src = src;
    v.status = status;
    v.p_da = vecs;
    v.mean = vecs + 3 * ld;
    v.normal = vecs + 6 * ld;
    v.loc_rot = vecs + 9 * ld;
    v.loc_trans = vecs + 12 * ld;
    v.knn_idx = knn_idx;
    v.n = n;
    v.ld = ld;
    v.k = (int)cfg.num_corres_points;
    v.use_huber = cfg.use_huber;
    v.flags = flags;
    const float da_gate_f = cfg.target_ivox_map_min_dist_in_voxel / 4;          // geometric_factor.hpp:283
    v.da_gate = (double)da_gate_f;
    const float max_corr_f = cfg.max_corres_distance * cfg.max_corres_distance;  // :299
    v.max_corr_sq = (double)max_corr_f;
    v.sigma = (double)cfg.lidar_point_noise_std_dev;
    v.kh = (double)cfg.huber_threshold;
    v.pvd = (double)cfg.plane_validity_distance;
    v.partials = partials;
    v.ticket = tickets;
    v.packed = packed;
    v.partials2 = partials2;
    v.ticket2 = tickets + 1;
    v.loc_out = packed + kPack;
    return v;
  }
};

namespace {

int reset_state(mb_factor* f) {
  cudaStream_t st = f->ctx->stream;
  if (f->n) {
    MB_CUDA(cudaMemsetAsync(f->status, 0, f->n, st));
    MB_CUDA(cudaMemsetAsync(f->vecs, 0, 15 * f->ld * sizeof(double), st));
    MB_CUDA(cudaMemsetAsync(f->knn_idx, 0xff, f->n * f->cfg.num_corres_points * sizeof(uint64_t), st));
  }
  f->linearize_count = 0;
  return MB_OK;
}

// Enqueue one linearisation (+ optional GN step) on the context stream.
int enqueue_linearize(mb_factor* f, int do_step, int iter, mb_icp_trace* d_trace, int linearize_count) {
  mb_ctx* c = f->ctx;
  cudaStream_t st = c->stream;
  const FactorView fv = f->view();
  k_linearize<<<f->grid, kLinWarps * 32, 0, st>>>(f->map->view(), fv, f->ds->pose);
  if (c->world > 1) MB_NCCL(ncclAllReduce(f->packed, f->packed, kPack, ncclDouble, ncclSum, c->comm, st));
  k_finalize<<<1, 32, 0, st>>>(f->packed, f->ds, f->cfg.reg_4_dof, linearize_count, do_step, iter, d_trace);
  k_loc_comp<<<f->grid2, 256, 0, st>>>(fv, f->ds);
  if (c->world > 1) MB_NCCL(ncclAllReduce(f->packed + kPack, f->packed + kPack, 6, ncclDouble, ncclSum, c->comm, st));
  c->launches += 3;
  MB_CUDA(cudaGetLastError());
  return MB_OK;
}

void drop_graph(mb_factor* f) {
  if (f->graph) cudaGraphExecDestroy(f->graph);
  f->graph = nullptr;
  f->graph_iters = 0;
  f->graph_count0 = -1;
}

}  // namespace

extern "C" {

int mb_factor_create(mb_ctx* ctx, mb_map* map, const void* pts, size_t n, size_t stride_bytes,
                     const mb_icp_config* cfg, size_t shard_begin, size_t shard_end, mb_factor** out) {
  MB_REQUIRE(ctx && map && cfg && out, "null argument");
  MB_REQUIRE(n == 0 || pts, "null scan");
  MB_REQUIRE(map->ctx == ctx, "map belongs to another context");
  MB_REQUIRE(stride_bytes >= 12 && stride_bytes % 4 == 0, "stride must be >= 12 and a multiple of 4");
  MB_REQUIRE(shard_begin <= shard_end && shard_end <= n, "bad shard range");
  if (cfg->project_on_degneneracy) {
    set_error("mb_factor_create: project_on_degneneracy=true is not supported (the reference branch at "
              "geometric_factor.hpp:477-557 re-sums arrays that are never written)");
    return MB_ERR_UNSUPPORTED;
  }
  if (cfg->num_corres_points < 3 || cfg->num_corres_points > MB_MAX_K) {
    set_error("mb_factor_create: num_corres_points=%llu outside [3, %d] (geometric_config.cpp:32-48)",
              (unsigned long long)cfg->num_corres_points, MB_MAX_K);
    return MB_ERR_UNSUPPORTED;
  }
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  mb_factor* f = new mb_factor;
  f->ctx = ctx;
  f->map = map;
  map->refs.fetch_add(1);
  f->cfg = *cfg;
  f->n_total = n;
  f->begin = shard_begin;
  f->n = shard_end - shard_begin;
  f->ld = std::max<size_t>((f->n + 31) & ~(size_t)31, 32);
  const size_t k = cfg->num_corres_points;
  const size_t n_tiles = (f->n + 31) / 32;
  f->grid = (int)std::max<size_t>(1, std::min<size_t>((n_tiles + kLinWarps - 1) / kLinWarps, (size_t)ctx->sm_count * 8));
  f->grid2 = (int)std::max<size_t>(1, std::min<size_t>((f->n + 255) / 256, (size_t)ctx->sm_count * 2));
  int rc = MB_OK;
  auto alloc = [&](void** p, size_t bytes) {
    if (rc != MB_OK) return;
    cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 256));
    if (e != cudaSuccess) {
      set_error("mb_factor_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      rc = MB_ERR_CUDA;
    }
  };
  alloc((void**)&f->src, f->ld * sizeof(float4));
  alloc((void**)&f->status, f->ld);
  alloc((void**)&f->vecs, 15 * f->ld * sizeof(double));
  alloc((void**)&f->knn_idx, f->ld * k * sizeof(uint64_t));
  alloc((void**)&f->partials, (size_t)f->grid * kPack * sizeof(double));
  alloc((void**)&f->partials2, (size_t)f->grid2 * 8 * sizeof(double));
  alloc((void**)&f->packed, (kPack + 8) * sizeof(double));
  alloc((void**)&f->tickets, 4 * sizeof(unsigned));
  alloc((void**)&f->ds, sizeof(DevState));
  if (rc != MB_OK) {
    mb_factor_release(f);
    return rc;
  }
  // scan: host AoS with arbitrary stride -> device float4 (only xyz is read by the factor,
  // geometric_factor.hpp:277,323,346)
  {
    int prc = pinned_reserve(ctx, f->ld * sizeof(float4));
    if (prc != MB_OK) {
      mb_factor_release(f);
      return prc;
    }
  }
  float4* h = (float4*)ctx->pinned;
  for (size_t i = 0; i < f->n; ++i) {
    const float* p = (const float*)((const char*)pts + (shard_begin + i) * stride_bytes);
    h[i] = make_float4(p[0], p[1], p[2], 0.f);
  }
  for (size_t i = f->n; i < f->ld; ++i) h[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  MB_CUDA(cudaMemcpyAsync(f->src, h, f->ld * sizeof(float4), cudaMemcpyHostToDevice, st));
  MB_CUDA(cudaMemsetAsync(f->tickets, 0, 4 * sizeof(unsigned), st));
  MB_CUDA(cudaMemsetAsync(f->packed, 0, (kPack + 8) * sizeof(double), st));
  MB_CUDA(cudaMemsetAsync(f->ds, 0, sizeof(DevState), st));
  MB_TRY(reset_state(f));
  MB_CUDA(cudaStreamSynchronize(st));
  *out = f;
  return MB_OK;
}

int mb_factor_release(mb_factor* f) {
  if (!f) return MB_OK;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  drop_graph(f);
  cudaFree(f->src);
  cudaFree(f->status);
  cudaFree(f->vecs);
  cudaFree(f->knn_idx);
  cudaFree(f->partials);
  cudaFree(f->partials2);
  cudaFree(f->packed);
  cudaFree(f->tickets);
  cudaFree(f->ds);
  cudaFree(f->d_trace);
  mb_map_release(f->map);
  delete f;
  return MB_OK;
}

int mb_factor_reset(mb_factor* f) {
  MB_REQUIRE(f, "null factor");
  MB_CUDA(cudaSetDevice(f->ctx->device));
  return reset_state(f);
}

int mb_factor_set_flags(mb_factor* f, uint32_t flags) {
  MB_REQUIRE(f, "null factor");
  if (flags != f->flags) drop_graph(f);
  f->flags = flags;
  return MB_OK;
}

int mb_factor_linearize(mb_factor* f, const double R[9], const double t[3], const double gravity_unit[3],
                        mb_linearization* out) {
  MB_REQUIRE(f && R && t && gravity_unit && out, "null argument");
  MB_CUDA(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  double h[15];
  std::memcpy(h, R, 9 * sizeof(double));
  std::memcpy(h + 9, t, 3 * sizeof(double));
  std::memcpy(h + 12, gravity_unit, 3 * sizeof(double));
  MB_CUDA(cudaMemcpyAsync(f->ds->pose, h, sizeof(h), cudaMemcpyHostToDevice, st));
  ++f->linearize_count;
  MB_TRY(enqueue_linearize(f, 0, 0, nullptr, f->linearize_count));
  double loc[6];
  MB_CUDA(cudaMemcpyAsync(out, &f->ds->lin, sizeof(mb_linearization), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaMemcpyAsync(loc, f->packed + kPack, sizeof(loc), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  for (int a = 0; a < 3; ++a) {
    out->loc_trans_comp[a] = loc[a];
    out->loc_rot_comp[a] = loc[3 + a];
  }
  return MB_OK;
}

int mb_factor_download_state(mb_factor* f, uint8_t* status, double* p_da, double* mean, double* normal,
                             double* loc_rot, double* loc_trans, uint64_t* knn_idx) {
  MB_REQUIRE(f, "null factor");
  MB_CUDA(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  const size_t n = f->n;
  if (n == 0) return MB_OK;
  if (status) MB_CUDA(cudaMemcpyAsync(status, f->status, n, cudaMemcpyDeviceToHost, st));
  if (knn_idx)
    MB_CUDA(cudaMemcpyAsync(knn_idx, f->knn_idx, n * f->cfg.num_corres_points * sizeof(uint64_t),
                            cudaMemcpyDeviceToHost, st));
  double* outs[5] = {p_da, mean, normal, loc_rot, loc_trans};
  std::vector<double> soa;
  for (int v = 0; v < 5; ++v) {
    if (!outs[v]) continue;
    soa.resize(3 * f->ld);
    MB_CUDA(cudaMemcpyAsync(soa.data(), f->vecs + (size_t)v * 3 * f->ld, 3 * f->ld * sizeof(double),
                            cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    for (size_t i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) outs[v][3 * i + c] = soa[(size_t)c * f->ld + i];
  }
  MB_CUDA(cudaStreamSynchronize(st));
  return MB_OK;
}

int mb_icp_run(mb_factor* f, double R[9], double t[3], int iters, double lambda, mb_icp_trace* trace) {
  MB_REQUIRE(f && R && t, "null argument");
  MB_REQUIRE(iters >= 0 && iters <= 4096, "iters outside [0, 4096]");
  MB_CUDA(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  if (iters > f->trace_cap) {
    MB_CUDA(cudaStreamSynchronize(st));
    drop_graph(f);
    cudaFree(f->d_trace);
    f->d_trace = nullptr;
    MB_CUDA(cudaMalloc(&f->d_trace, (size_t)iters * sizeof(mb_icp_trace)));
    f->trace_cap = iters;
  }
  double h[16];
  std::memcpy(h, R, 9 * sizeof(double));
  std::memcpy(h + 9, t, 3 * sizeof(double));
  h[12] = 0.0;
  h[13] = 0.0;
  h[14] = -1.0;
  h[15] = lambda;
  MB_CUDA(cudaMemcpyAsync(f->ds->pose, h, sizeof(h), cudaMemcpyHostToDevice, st));
  const bool use_graph = (f->flags & 2u) != 0;
  if (use_graph) {
    // The captured sequence bakes the iteration index and linearize_count into k_finalize's arguments.
    if (!f->graph || f->graph_iters != iters || f->graph_count0 != f->linearize_count) {
      drop_graph(f);
      cudaGraph_t g = nullptr;
      MB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int rc = MB_OK;
      for (int it = 0; it < iters && rc == MB_OK; ++it)
        rc = enqueue_linearize(f, 1, it, f->d_trace, f->linearize_count + it + 1);
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (rc != MB_OK) {
        if (g) cudaGraphDestroy(g);
        return rc;
      }
      MB_CUDA(e);
      e = cudaGraphInstantiate(&f->graph, g, 0);
      cudaGraphDestroy(g);
      MB_CUDA(e);
      f->graph_iters = iters;
      f->graph_count0 = f->linearize_count;
    } else {
      f->ctx->launches += 3ull * iters;
    }
    MB_CUDA(cudaGraphLaunch(f->graph, st));
    f->linearize_count += iters;
  } else {
    for (int it = 0; it < iters; ++it) {
      ++f->linearize_count;
      MB_TRY(enqueue_linearize(f, 1, it, f->d_trace, f->linearize_count));
    }
  }
  double ho[12];
  MB_CUDA(cudaMemcpyAsync(ho, f->ds->pose, sizeof(ho), cudaMemcpyDeviceToHost, st));
  if (trace && iters)
    MB_CUDA(cudaMemcpyAsync(trace, f->d_trace, (size_t)iters * sizeof(mb_icp_trace), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  std::memcpy(R, ho, 9 * sizeof(double));
  std::memcpy(t, ho + 9, 3 * sizeof(double));
  return MB_OK;
}

}  // extern "C"
