// Point-to-plane scan-to-map ICP factor on the device: the body of mimosa::lidar::ICPFactor::linearize
// (mimosa/include/mimosa/lidar/geometric_factor.hpp:231-562) and estimatePlane (:176-229).
//
//   k_linearize   per tile of 128 scan points, in one kernel:
//                   A  transform + data-association gate (:276-287), block-compaction of the points that
//                      must re-associate;
//                   B  restricted k-NN (:294; one query per thread, mb_internal.cuh::knn_thread), the two
//                      distance gates (:296-302) and the plane fit (:176-229) for the compacted points;
//                   C  residual, s-check, Huber, Jacobian, per-point localizability vectors (:319-355) and the
//                      [J e]^T [J e] / status-count reduction (:364-366, 396-403).
//                 The neighbour points never leave the SM; per-point state is read and written once.
//   k_finalize    6x6 assembly, localizability eigen-decompositions, Schur complements, 4-DoF projection
//                 (:405-428, 464-475) and — for the stand-alone Gauss-Newton harness — the LDL^T solve and
//                 SE(3) retract that ISAM2 performs in the reference (mimosa/src/graph/manager.cpp:585-588).
//                 Five independent roles run on five warps.
//   k_loc_comp    the component-localizability second pass over valid points (:434-457); in the device-resident
//                 loop that pass is folded into the next k_linearize and this kernel only runs after the last
//                 iteration.  Its last block hands the finished linearisation to the polling host.
// With more than one rank the 48-double packet between k_linearize and k_finalize (and the 6 doubles after
// k_loc_comp) travel through peer-memory mailboxes (mb_internal.cuh) or, as the fallback, ncclAllReduce; every
// rank then computes the identical step.
//
// Per-point state (status, DA anchor, plane mean/normal, localizability vectors) lives in HBM inside the
// factor handle as structure-of-arrays, mirroring the `mutable` vectors at geometric_factor.hpp:79-106.
// Every reduction has a fixed order, so results are bitwise repeatable for a given launch shape.
#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include <type_traits>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "mb_map.cuh"
#include "mb_scan.cuh"

namespace mb {
namespace {

constexpr int kLinThreads = 128;
constexpr int kLinWarps = kLinThreads / 32;
// 21 H upper-tri, 6 J^T e, 1 f, 9 status counts, 1 n_searched, 2 pad, then (device-resident loop only) the six
// component localizabilities of the PREVIOUS linearisation, folded into this pass (see k_linearize), 2 pad
constexpr int kPack = 48;
constexpr int kPackB = 21, kPackF = 27, kPackCnt = 28, kPackSearched = 37, kPackLoc = 40;
constexpr int kGroup = 32;  // blocks per first-level reduction group
constexpr int kLocThreads = 256;
// packed upper triangle of [J e]^T [J e] (7x7), entry a = (kTriRow[a], kTriCol[a]): the 21 entries of J^T J row-major
// (c >= r), then J^T e (6), then e^2; entries 28..31 unused
__constant__ int8_t kTriRow[32] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5, 0, 1, 2, 3, 4, 5, 6, 0, 0, 0, 0};
__constant__ int8_t kTriCol[32] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5, 6, 6, 6, 6, 6, 6, 6, 0, 0, 0, 0};

struct FactorView {
  const float4* src;
  uint8_t* status;
  double *p_da, *mean, *normal, *loc_rot, *loc_trans;  // SoA: component c of point i at [c * ld + i]
  uint64_t* knn_idx;                                   // [i * k + j]
  size_t n, ld;
  int k, use_huber;
  int tile;  // points per block tile of k_linearize: 128, 64 or 32
  uint32_t flags;
  int fold_loc;  // 1: also sum the component localizabilities of the previous linearisation (device-resident loop)
  double da_gate_sq, max_corr_sq, sigma, kh, pvd;  // da_gate_sq: smallest d2 with sqrt(d2) > the gate (host: da_gate_sq_min)
  const double* rroot;  // sqrt(||p_src||) per point (geometric_factor.hpp:323), constant per scan
  double* partials;    // [grid][kPack]
  double* gpartials;   // [n_groups][kPack]
  unsigned* gtickets;  // [n_groups]
  unsigned* ticket;    // final
  double* packed;      // [kPack]
  double* partials2;   // [grid2][8]
  unsigned* ticket2;
  double* loc_out;     // [8]: trans comp (3), rot comp (3)
  // persistent loop (k_icp_loop)
  struct LoopCtl* ctl;
  uint32_t* queue;       // [n] points that re-associate in the current linearisation
  uint32_t* tile_stamp;  // [n_tiles] linearize_count of the last linearisation that deferred the tile to its second pass
};

// Pose handed to k_linearize BY VALUE (kernel parameter space) on the host-facing single-call path: no H2D copy
// is enqueued for 15 doubles.
struct PoseArg {
  double v[16];  // R (9), t (3), gravity (3), lambda
};

struct DevState {        // small device-resident block per factor
  double pose[12];       // R row-major (9), t (3)
  double gravity[3];
  double lambda;
  mb_linearization lin;  // result of the most recent linearisation (loc_*_comp filled on the host)
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
template <int K>
__device__ __forceinline__ uint8_t fit_plane(const float4 (&nb)[K], int k, d3 origin, double pvd, d3& mean, d3& normal,
                                             bool& normal_set) {
  d3 s = mk3(0, 0, 0);
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (j < k) s = add3(s, mk3((double)nb[j].x, (double)nb[j].y, (double)nb[j].z));
  mean = div3(s, (double)k);
  double c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (j < k) {
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
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (j < k) {
      const d3 c = sub3(mk3((double)nb[j].x, (double)nb[j].y, (double)nb[j].z), mean);
      if (fabs(dot3(c, nrm)) > pvd) invalid = true;
    }
  return invalid ? MB_CORRES_PLANE_INVALID : MB_UNPROCESSED;
}

// Sum `count` rows of `width`-double records (row stride = width) in ascending row order into out[0..width),
// using all `threads` threads of the block: thread t owns value t % width and every (threads / width)-th row.
// Loads are issued eight at a time.  s_tmp holds (threads / width) * width doubles.
__device__ __forceinline__ void block_sum_rows(const double* rows, int count, int width, double* s_tmp, double* out,
                                               int threads) {
  const int parts = threads / width;
  const int a = threadIdx.x % width, part = threadIdx.x / width;
  if (part < parts) {
    double v = 0.0;
    int b = part;
    for (; b + 15 * parts < count; b += 16 * parts) {  // 16 rows in flight: one L2 round trip for up to 16 * parts rows
      double x[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) x[u] = __ldcg(rows + (size_t)(b + u * parts) * width + a);
#pragma unroll
      for (int u = 0; u < 16; ++u) v += x[u];
    }
    {
      double x[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) x[u] = b + u * parts < count ? __ldcg(rows + (size_t)(b + u * parts) * width + a) : 0.0;
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (b + u * parts < count) v += x[u];
    }
    s_tmp[part * width + a] = v;
  }
  __syncthreads();
  if (threadIdx.x < width) {
    double v = 0.0;
    for (int p = 0; p < parts; ++p) v += s_tmp[p * width + threadIdx.x];
    out[threadIdx.x] = v;
  }
}

// computeLocalizability (mimosa/include/mimosa/utils.hpp:308-313).  The closed-form solver (checked a posteriori,
// iterative fallback) replaces the reference's iterative one here: its input, the reduced 6x6, already differs
// from the CPU's in the last bits, so bit parity is not at stake, and it is 5x shorter as a single-thread chain.
__device__ __forceinline__ void localizability(const m33& JtJ, double loc[3], m33& V) {
  double lam[3];
  eigh33_direct(JtJ, lam, V);
#pragma unroll
  for (int a = 0; a < 3; ++a) loc[a] = sqrt(lam[a]);
}

#ifndef MB_LIN_BLOCKS
#define MB_LIN_BLOCKS 4  // measured: 4 x 128 threads at 128 registers beat 5 x 96 and 3 x 156
#endif
#if defined(MB_LOOP_TIMING)  // development: SM clock of block 0 at the phase boundaries of every linearisation
__device__ long long g_loop_t[64][12];
#define MB_LOOP_T(it, slot) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (it) < 64) g_loop_t[it][slot] = clock64(); } while (0)
__device__ long long g_loop_f[64][12];
#define MB_LOOP_F(it, slot) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (it) < 64) g_loop_f[it][slot] = clock64(); } while (0)
#define MB_FIN_STORE(it, slot) g_loop_f[it][slot] = clock64()
#else
#define MB_LOOP_F(it, slot) do { } while (0)
#define MB_FIN_STORE(it, slot) (void)0
#define MB_LOOP_T(it, slot) do { } while (0)
#endif
#define MB_FIN_T(it, slot) do { if (blockIdx.x == 0 && (it) < 64) MB_FIN_STORE(it, slot); } while (0)
// What follows the reduction of one linearisation (arguments of the finalize roles).
struct FinArgs {
  int reg_4_dof, linearize_count, do_step, iter;
  mb_icp_trace* trace;
};

// Everything after the per-point loop of ICPFactor::linearize (geometric_factor.hpp:405-428, 464-475, 559-560), plus
// the harness GN step that ISAM2 performs in the reference (mimosa/src/graph/manager.cpp:585-588).  One THREAD per
// role: 0/1 localizability of the rotational / translational block, 2/3 Schur-complement degeneracy info,
// 4 projection + packing + solve + retract.  `packed` = the reduced 48-double packet (all ranks summed).
__device__ __noinline__ void finalize_role(const double* packed, const DevState* in, DevState* out, const FinArgs& fa, int role) {
  double H[36];
  {
    int u = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int c = a; c < 6; ++c) {
        H[6 * a + c] = packed[u];
        H[6 * c + a] = packed[u];
        ++u;
      }
  }
  mb_linearization& L = out->lin;
  m33 Hrr, Hrt, Htr, Htt;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Hrr.m[3 * r + c] = H[6 * r + c];
      Hrt.m[3 * r + c] = H[6 * r + 3 + c];
      Htr.m[3 * r + c] = H[6 * (r + 3) + c];
      Htt.m[3 * r + c] = H[6 * (r + 3) + 3 + c];
    }
  if (role == 4) MB_FIN_T(fa.iter, 8);
  if (role < 4) {
    // The four eigen roles run as four LANES of one warp: the same instruction stream for all of them (one trip
    // through the instruction cache instead of four), selected inputs and outputs.  Lanes 2 / 3 first form their Schur
    // complement  A - B C^-1 D  with (A, B, C, D) = (Hrr, Hrt, Htt, Htr) / (Htt, Htr, Hrr, Hrt), inverted (:413-422).
    m33 M = role == 0 ? Hrr : Htt;
    if (role >= 2) {
      const m33& A = role == 2 ? Hrr : Htt;
      const m33& B = role == 2 ? Hrt : Htr;
      const m33& C = role == 2 ? Htt : Hrr;
      const m33& D = role == 2 ? Htr : Hrt;
      M = inv33(sub33(A, mul33(mul33(B, inv33(C)), D)));
    }
    m33 V;
    double loc[3];
    localizability(M, loc, V);
    double* const loc_out = role == 0 ? L.loc_rot_final : role == 1 ? L.loc_trans_final : role == 2 ? L.degen_rot : L.degen_trans;
    double* const vec_out = role == 0 ? L.eigvec_rot : role == 1 ? L.eigvec_trans : role == 2 ? L.degen_eigvec_rot : L.degen_eigvec_trans;
    for (int a = 0; a < 3; ++a) loc_out[a] = role == 2 ? loc[a] * 57.29578 : loc[a];  // RAD2DEG (PCL's macro), :428
    for (int a = 0; a < 9; ++a) vec_out[a] = V.m[a];
    if (role == 0) MB_FIN_T(fa.iter, 11);
  } else {
    double b[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) b[a] = packed[kPackB + a];
    const double f = packed[kPackF];
    m33 R;
#pragma unroll
    for (int a = 0; a < 9; ++a) R.m[a] = in->pose[a];
    d3 T = mk3(in->pose[9], in->pose[10], in->pose[11]);
    if (fa.reg_4_dof) {
      const d3 gz = mk3(-in->gravity[0], -in->gravity[1], -in->gravity[2]);
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
    double g[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) g[a] = -b[a];
#pragma unroll
    for (int a = 0; a < 36; ++a) L.H[a] = H[a];
#pragma unroll
    for (int a = 0; a < 6; ++a) L.g[a] = g[a];
    L.f = f;
    for (int a = 0; a < 9; ++a) L.counts[a] = (int64_t)packed[kPackCnt + a];
    L.linearize_count = fa.linearize_count;
    L.n_searched = (int32_t)packed[kPackSearched];
    MB_FIN_T(fa.iter, 9);
    if (fa.do_step) {
      double delta[6] = {0, 0, 0, 0, 0, 0};
      const bool ok = solve6_ldlt(H, in->lambda, g, delta);
      MB_FIN_T(fa.iter, 10);
      if (ok) {
        se3_retract(R, T, delta);
#pragma unroll
        for (int a = 0; a < 9; ++a) out->pose[a] = R.m[a];
        out->pose[9] = T.x;
        out->pose[10] = T.y;
        out->pose[11] = T.z;
      }
      if (fa.trace) {
        mb_icp_trace& tr = fa.trace[fa.iter];
#pragma unroll
        for (int a = 0; a < 36; ++a) tr.H[a] = H[a];
        for (int a = 0; a < 6; ++a) {
          tr.g[a] = g[a];
          tr.delta[a] = delta[a];
        }
        tr.f = f;
        for (int a = 0; a < 9; ++a) {
          tr.R[a] = R.m[a];
          tr.counts[a] = (int64_t)packed[kPackCnt + a];
        }
        tr.t[0] = T.x;
        tr.t[1] = T.y;
        tr.t[2] = T.z;
        tr.n_searched = (int32_t)packed[kPackSearched];
        tr.solve_ok = ok ? 1 : 0;
        // the packet also carries the component localizabilities of the PREVIOUS linearisation (folded pass)
        if (fa.iter > 0) {
          for (int a = 0; a < 3; ++a) {
            fa.trace[fa.iter - 1].loc_trans_comp[a] = packed[kPackLoc + a];
            fa.trace[fa.iter - 1].loc_rot_comp[a] = packed[kPackLoc + 3 + a];
          }
        }
      }
    }
  }
}

// What follows the reduction, in its own small kernel (two warps: the solve, and the four eigen roles as four lanes of
// the other).  Running the roles inside k_linearize's last block was built and measured (profiles/r2_experiments.md):
// under k_linearize's register cap the same chain takes 10-13 us instead of 4-6.
__global__ void __launch_bounds__(64) k_finalize(const double* packed, DevState* ds, FinArgs fa, unsigned role_mask) {
  pdl_launch_dependents();
  pdl_wait();
  // warp 0, lane 0: projection + packing + solve + retract; warp 1, lanes 0..3: the four eigen roles in lock step
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = warp == 0 ? (lane == 0 ? 4 : -1) : (lane < 4 ? lane : -1);
  if (role < 0 || ((role_mask >> role) & 1u) == 0) return;  // role_mask != 31 only in the timing diagnostic
  finalize_role(packed, ds, ds, fa, role);
}

// One linearisation: block tiles of 128 points of the voxel-ordered scan, block-level compaction of the points that
// re-associate, one query per thread (mb_search.cuh::knn_thread).  Two other shapes of this kernel were built and
// measured in round 2 and lost (profiles/r2_experiments.md): one warp per 32-point tile with the neighbourhood
// resolved once per voxel group and staged through the bulk-copy engine, and a three-phase form with grid barriers
// and pulled, compacted search passes.
template <int K, int ROWS>
__global__ void __launch_bounds__(kLinThreads, MB_LIN_BLOCKS) k_linearize(MapView mv, FactorView fv, DevState* ds) {
  double* pose_dev = ds->pose;
  __shared__ uint16_t s_tab[kTabEntries];
  // s_pk (phase B: probed neighbour words, [n_off][thread]; cooperative search: the per-thread candidate stacks,
  // [3 * kCoopStack][thread]) is re-used as s_row (phase C: whitened [J (6), e] per point, [warp][32][7] doubles =
  // 7168 B <= 19 * 128 * 4 B).
  __shared__ __align__(16) uint32_t s_pk_all[ROWS * kLinThreads];
  // per-thread {mask_lo, mask_hi, base} of the <= 8 blocks around a query; cooperative search: kGroupWords per group
  __shared__ uint32_t s_blk_all[24 * kLinThreads];
  static_assert(sizeof(s_pk_all) >= sizeof(double) * 7 * kLinThreads, "s_row fits");
  __shared__ double s_pt[3][kLinThreads];          // transformed point of each tile member
  __shared__ uint8_t s_status[kLinThreads];
  __shared__ uint16_t s_queue[kLinThreads];
  __shared__ int s_warp_need[kLinWarps];
  __shared__ double s_red[kLinWarps][kPack];
  __shared__ double s_tmp[(kLinThreads / kPack) * kPack];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_launch_dependents();
  fill_scan_table(mv, s_tab);
  double(*s_row)[7] = reinterpret_cast<double(*)[7]>(s_pk_all) + warp * 32;
  // Lane a < 28 owns entry a of the packed upper triangle of [J e]^T [J e] (7x7): the 21 entries of
  // J^T J first (row-major, c >= r), then J^T e (6), then e^2.
  const int pr = kTriRow[lane], pc = kTriCol[lane];

  // everything above is independent of earlier kernels; from here on we read the pose the previous iteration's
  // k_finalize wrote and overwrite per-point state the previous k_loc_comp may still be reading
  pdl_wait();
  m33 R;
#pragma unroll
  for (int a = 0; a < 9; ++a) R.m[a] = pose_dev[a];
  const d3 T = mk3(pose_dev[9], pose_dev[10], pose_dev[11]);
  const int k = fv.k;
  const bool forced = (fv.flags & 1u) != 0;
  const double inv_sigma = 1.0 / fv.sigma;  // = sqrt_w / sigma for the points Huber leaves alone (sqrt_w == 1)

  double acc = 0.0, lacc = 0.0;  // lacc: lane a < 6 sums component a of the previous linearisation's localizability pass
  int cnt = 0;  // lane s < 9: points with status s; lane 9: points searched
  // Folded localizability pass (device-resident loop): before a point's status and localizability vectors are
  // overwritten, its contribution |loc^T V| (entries below 0.5 dropped, geometric_factor.hpp:434-457) to the
  // component localizabilities of the PREVIOUS linearisation is taken with that linearisation's eigenvectors, which
  // the previous k_finalize left in ds->lin.  It saves a pass over the points and a kernel launch per iteration.
  __shared__ double s_V[18];
  if (fv.fold_loc && tid < 18) s_V[tid] = tid < 9 ? ds->lin.eigvec_trans[tid] : ds->lin.eigvec_rot[tid - 9];
  __syncthreads();

  // Points per tile: 128 when the scan fills the device, 64 or 32 for shards that would otherwise leave SMs without a
  // block (mb_factor: tile_points).  The threads beyond a tile's points idle in phases A and C.
  const int tile_pts = fv.tile;
  const size_t n_tiles = (fv.n + tile_pts - 1) / tile_pts;
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t i = tile * tile_pts + tid;
    const bool act = tid < tile_pts && i < fv.n;
    // ---- A: transform, gate, compaction ------------------------------------------------------------
    d3 ps = mk3(0, 0, 0), pt = mk3(0, 0, 0);
    uint8_t st = MB_UNPROCESSED;
    bool need = false;
    if (act) {
      const float4 s = __ldg(fv.src + i);
      ps = mk3((double)s.x, (double)s.y, (double)s.z);
      pt = add3(mul33v(R, ps), T);
      st = fv.status[i];
      const d3 da = ld3(fv.p_da, fv.ld, i);
      need = forced || sqnorm3(sub3(pt, da)) >= fv.da_gate_sq;
    }
    s_pt[0][tid] = pt.x;
    s_pt[1][tid] = pt.y;
    s_pt[2][tid] = pt.z;
    const unsigned mask = __ballot_sync(kFull, need);
    if (lane == 0) s_warp_need[warp] = __popc(mask);
    __syncthreads();
    int base = 0, n_need = 0;
#pragma unroll
    for (int w = 0; w < kLinWarps; ++w) {
      if (w < warp) base += s_warp_need[w];
      n_need += s_warp_need[w];
    }
    if (need) s_queue[base + __popc(mask & ((1u << lane) - 1))] = (uint16_t)tid;
    __syncthreads();
    // ---- B: search + plane fit for the compacted points -------------------------------------------
      // Queries per warp and pass: a tile's re-associations are dealt out evenly over its four warps.  A warp's search
      // time grows with the number of different queries it carries (every lane walks its own buckets and the warp
      // follows the union: measured 7 us for two queries, 50 us for thirty-two), so a tile with 17 points to
      // re-associate runs four warps of 5 instead of one warp of 17 — and a 32-point tile four warps of 8.
      const int per_warp = min(32, max(1, (n_need + kLinWarps - 1) / kLinWarps));
      for (int q0 = warp * per_warp; q0 < n_need; q0 += kLinWarps * per_warp) {
        const int qi = q0 + lane;
        const bool on = lane < per_warp && qi < n_need;
        const int li = on ? (int)s_queue[qi] : 0;
        const double qx = s_pt[0][li], qy = s_pt[1][li], qz = s_pt[2][li];
        double bd[K];
        uint32_t bs[K];
        uint32_t* s_pk = s_pk_all + tid;
        knn_thread<K>(mv, s_tab, s_pk, s_blk_all + tid, kLinThreads, qx, qy, qz, k, on, bd, bs);
        if (on) {
          const size_t gi = tile * tile_pts + li;
          float4 nb[K];
          uint64_t g[K];
          const int found = knn_resolve_all<K, true>(mv, s_pk, kLinThreads, bs, k, g, nb);
          double dk = 0.0;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            if (j < k) {
              if (j == k - 1) dk = bd[j];
              // indices are only meaningful when all k exist (the reference discards partial results)
              if (fv.knn_idx) fv.knn_idx[gi * k + j] = g[j];
            }
          }
          uint8_t rs = MB_UNPROCESSED;
          d3 mean = mk3(0, 0, 0), normal = mk3(0, 0, 0);
          if (found != k) {
            rs = MB_INSUFFICIENT_CORRES_POINTS;
            if (fv.knn_idx)
              for (int j = 0; j < k; ++j) fv.knn_idx[gi * k + j] = ~0ull;
          } else if (dk > fv.max_corr_sq) {
            rs = MB_CORRES_MAX_DIST;
          } else {
            bool normal_set;
            rs = fit_plane<K>(nb, k, T, fv.pvd, mean, normal, normal_set);
            st3(fv.mean, fv.ld, gi, mean);
            if (normal_set) st3(fv.normal, fv.ld, gi, normal);
          }
          s_status[li] = rs;  // a fitted plane itself reaches its owner through fv.mean / fv.normal (same block, barrier below)
        }
      }
    __syncthreads();
    // ---- C: residual, Jacobian, accumulation (own point) --------------------------------------------
    double row[7] = {0, 0, 0, 0, 0, 0, 0};
    if (fv.fold_loc) {
      double v[6] = {0, 0, 0, 0, 0, 0};
      if (act && st == MB_VALID) {  // st is still the status the previous linearisation left
        const d3 lt = ld3(fv.loc_trans, fv.ld, i), lr = ld3(fv.loc_rot, fv.ld, i);
#pragma unroll
        for (int a = 0; a < 3; ++a) {  // column a of V: (V^T loc)_a, in mul33Tv's order of operations
          const double tc = fabs(s_V[a] * lt.x + (s_V[3 + a] * lt.y + s_V[6 + a] * lt.z));
          const double rc = fabs(s_V[9 + a] * lr.x + (s_V[12 + a] * lr.y + s_V[15 + a] * lr.z));
          v[a] = tc >= 0.5 ? tc : 0.0;
          v[3 + a] = rc >= 0.5 ? rc : 0.0;
        }
      }
      if (__ballot_sync(kFull, act && st == MB_VALID)) {
#pragma unroll
        for (int a = 0; a < 6; ++a) s_row[lane][a] = v[a];
        __syncwarp();
        if (lane < 6) {
#pragma unroll 8
          for (int p = 0; p < 32; ++p) lacc += s_row[p][lane];
        }
        __syncwarp();
      }
    }
    if (act) {
      bool proceed = false;
      d3 mean = mk3(0, 0, 0), normal = mk3(0, 0, 0);
      if (need) {
        st3(fv.p_da, fv.ld, i, pt);
        st = s_status[tid];
        if (st == MB_UNPROCESSED) {
          mean = ld3(fv.mean, fv.ld, i);
          normal = ld3(fv.normal, fv.ld, i);
          proceed = true;
        }
      } else if (st > MB_CORRES_PLANE_INVALID) {
        mean = ld3(fv.mean, fv.ld, i);
        normal = ld3(fv.normal, fv.ld, i);
        proceed = true;
      }
      if (proceed) {
        double e = dot3(normal, sub3(mean, pt));
        const double s_chk = 1 - 0.9 * fabs(e) / fv.rroot[i];
        if (s_chk < 0.9) {
          st = MB_MAX_ERROR;
        } else {
          double sqrt_w = 1.0;
          if (fv.use_huber) {
            const double we = e / fv.sigma;
            if (fabs(we) > fv.kh) sqrt_w = sqrt(fv.kh / fabs(we));
          }
          const double scale = sqrt_w == 1.0 ? inv_sigma : sqrt_w / fv.sigma;
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
    // [J e]^T [J e] over the warp's 32 points: lane a sums its entry over the rows in point order.
    const unsigned any_valid = __ballot_sync(kFull, act && st == MB_VALID);
    if (any_valid) {
#pragma unroll
      for (int a = 0; a < 7; ++a) s_row[lane][a] = row[a];
      __syncwarp();
#pragma unroll 8
      for (int p = 0; p < 32; ++p) acc += s_row[p][pr] * s_row[p][pc];
    }
#pragma unroll
    for (int s = 0; s < 9; ++s) {
      const int c = __popc(__ballot_sync(kFull, act && st == s));
      if (lane == s) cnt += c;
    }
    if (lane == 9) cnt += __popc(mask);
    __syncthreads();  // s_row (= s_pk), s_queue, s_status ... are rewritten by the next tile
  }

  // ---- block partial -> group partial -> packet (two ticketed levels, fixed order) ------------------------
  if (lane < 28) s_red[warp][lane] = acc;
  if (lane >= 30) s_red[warp][lane + 8] = s_red[warp][lane + 16] = 0.0;  // pad entries 38, 39, 46, 47
  if (lane < 10) s_red[warp][kPackCnt + lane] = (double)cnt;
  if (lane < 6) s_red[warp][kPackLoc + lane] = lacc;
  __syncthreads();
  if (tid < kPack) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kLinWarps; ++w) v += s_red[w][tid];
    fv.partials[(size_t)blockIdx.x * kPack + tid] = v;
  }
  const int g = blockIdx.x / kGroup;
  const int n_groups = (gridDim.x + kGroup - 1) / kGroup;
  const int g_size = min(kGroup, (int)gridDim.x - g * kGroup);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(fv.gtickets + g, 1u) == (unsigned)g_size - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  block_sum_rows(fv.partials + (size_t)g * kGroup * kPack, g_size, kPack, s_tmp, fv.gpartials + (size_t)g * kPack,
                 kLinThreads);
  if (tid == 0) fv.gtickets[g] = 0u;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(fv.ticket, 1u) == (unsigned)n_groups - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  block_sum_rows(fv.gpartials, n_groups, kPack, s_tmp, fv.packed, kLinThreads);
  if (tid == 0) *fv.ticket = 0u;
}

// Component localizabilities (geometric_factor.hpp:434-457): sum over Valid points of |loc_i^T V| with
// entries below 0.5 zeroed.
__global__ void __launch_bounds__(kLocThreads) k_loc_comp(FactorView fv, const DevState* __restrict__ ds) {
  __shared__ double s_red[kLocThreads / 32][8];
  __shared__ double s_tmp[(kLocThreads / 8) * 8];
  __shared__ bool s_last;
  pdl_launch_dependents();
  pdl_wait();
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
  if (lane == 0) s_red[warp][6] = s_red[warp][7] = 0.0;
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0.0;
    for (int w = 0; w < kLocThreads / 32; ++w) v += s_red[w][threadIdx.x];
    fv.partials2[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(fv.ticket2, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  block_sum_rows(fv.partials2, (int)gridDim.x, 8, s_tmp, fv.loc_out, kLocThreads);
  if (threadIdx.x == 0) *fv.ticket2 = 0u;
}

// ================================================================================================================
// The persistent loop: ALL linearisations of an mb_icp_run (or the one of an mb_factor_linearize) in ONE cooperative
// launch, one 512-thread block per SM, grid-wide barriers instead of kernel boundaries.  A block is four GROUPS of
// 128 threads; a group works on a tile of the voxel-ordered scan exactly like a block of k_linearize (named barriers
// instead of __syncthreads).  One linearisation:
//   A   per tile: transform + data-association gate (:276-287) and, in the device-resident loop, the folded
//       localizability pass of the previous linearisation.  A tile in which no point re-associates is finished on
//       the spot (C).  Otherwise its re-associating points are appended to a global queue (one reservation per tile,
//       so the queue keeps the voxel order tile by tile) and the tile is deferred.
//       -> block partial -> grid barrier #1 (it also publishes the queue length)
//   B   only when the queue is not empty: the queue is dealt out EVENLY over all warps of the device — restricted
//       k-NN (:294), distance gates (:296-302), plane fit (:176-229); results go to the points' state.  An iteration
//       that re-associates 13 % of the points, all of them in a few far-away tiles, keeps every SM busy with a few
//       queries per warp instead of a few blocks with full warps (a warp's search time grows with the number of
//       different queries it carries: 7 us for two, 50 us for thirty-two).  -> grid barrier #2
//   C'  the deferred tiles: residual, s-check, Huber, Jacobian, localizability vectors (:319-355), reduction
//       (:364-366, 396-403).  -> block partial -> grid barrier #3
//   D   EVERY block sums the block partials in block order (bitwise identical everywhere), exchanges the packet with
//       the other ranks through the peer mailboxes (block 0 stores, every block reads its own rank's mailbox), and
//       runs the finalize roles on its own copy of the factor's DevState in shared memory: the next linearisation
//       starts without another barrier.  Block 0 writes the trace.
// A fully cached linearisation costs ONE grid barrier.  After the last linearisation: the component-localizability
// pass (:434-457) over the block's own tiles, one more barrier, and block 0 hands the result to the device-side
// DevState, the mailboxes and the polling host.
struct LoopCtl {
  unsigned long long bar;    // grid barrier arrivals: monotonic, a multiple of the grid size between launches
  unsigned long long xflag;  // several ranks: number of the last exchange whose summed packet block 0 has published
  unsigned long long sflag;  // resident kernel: 2 * request number (+ 1: leave) block 0 has published to the other blocks
  unsigned long long epass;  // final passes (component localizabilities) this factor has completed
  unsigned q_count[2];       // queue length by linearisation parity
  unsigned q_next[2];        // second and later rounds of phase B: tasks handed out so far
};

struct LoopArgs {
  int iters, do_step, reg_4_dof, linearize_count0, has_pose;
  // serve (host-facing call): after the result of request req0 has been handed over the kernel stays
  // resident for window_ns, taking the poses of requests req0 + 1, req0 + 2, ... from the context's mapped block
  // (mb_internal.cuh: kSrv*) — one linearisation each — until the window passes without a request or the host asks
  // it to leave.
  int serve;
  unsigned long long req0;
  unsigned long long window_ns;
  char* mapped;  // the context's mapped page-locked block (nullptr: the launch hands nothing to a polling host)
  PoseArg pose;  // has_pose: pose / gravity of the host-facing call
  mb_icp_trace* trace;
};

#ifndef MB_LOOP_THREADS
#define MB_LOOP_THREADS 512
#endif
constexpr int kLoopThreads = MB_LOOP_THREADS, kLoopWarps = kLoopThreads / 32, kLoopGroups = kLoopThreads / kLinThreads;

constexpr uint8_t kFresh = 0x80;  // status bit: written by phase B, consumed by phase C'
constexpr int kPpt = 2;            // points per thread and tile in phases A / C (tiles of up to kPpt * 128 points)

template <int ROWS>
struct LoopShared {
  uint32_t pk[kLoopGroups][ROWS * kLinThreads];  // phase B: knn_thread's s_pk; phases A / C: [warp][32][7] doubles
  uint32_t blk[kLoopGroups][24 * kLinThreads];
  double rows[kLoopWarps][kPpt][32][8];  // phases A / C: the whitened [J (6), e, 0] rows of a warp's points, both slots
  double cmat[kLoopWarps][8][8];         // a warp's accumulated [J e]^T [J e] (tensor-core fragment -> packed entries)
  double red[kLoopWarps][kPack];
  double tmp[kLoopThreads];
  double packed[kXchgDoubles];
  DevState ds;
  unsigned long long bar_next;
  int warp_need[kLoopGroups][kLinWarps];
  unsigned qbase[kLoopGroups];
  uint16_t tab[kTabEntries];
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Flag-in-data words of the peer mailbox (mb_internal.cuh): one 8-byte store / load each, system scope, no ordering
// needed — a word carries its own exchange number.
__device__ __forceinline__ void ll_store(unsigned long long* p, uint32_t data, uint32_t flag) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(((unsigned long long)flag << 32) | data) : "memory");
}
__device__ __forceinline__ uint32_t ll_wait(const unsigned long long* p, uint32_t flag) {
  unsigned long long v;
  do {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  } while ((uint32_t)(v >> 32) != flag);
  return (uint32_t)v;
}
// This rank's `count` doubles src[0..count) into entries first.. of every rank's mailbox slot (all threads of the block).
__device__ __forceinline__ void ll_send(const PeerTable* peer, unsigned long long seq, const double* src, int first, int count) {
  const int world = peer->world;
  const size_t slot = (size_t)((seq & 1ull) * kMaxRanks + (unsigned)peer->rank) * (2 * kXchgDoubles);
  for (int x = threadIdx.x; x < world * count * 2; x += blockDim.x) {
    const int dst = x / (count * 2), w = x - dst * (count * 2);
    const unsigned long long bits = (unsigned long long)__double_as_longlong(src[w >> 1]);
    ll_store(peer->ll[dst] + slot + 2 * first + w, (uint32_t)(bits >> (32 * (w & 1))), (uint32_t)seq);
  }
}
// Wait for every rank's `count` doubles of exchange `seq` in this rank's mailbox and add them in rank order into
// dst[0..count) (shared memory; s_words: world * count * 2 words of scratch).  All threads of the block.
__device__ __forceinline__ void ll_gather(const PeerTable* peer, unsigned long long seq, int first, int count, uint32_t* s_words, double* dst) {
  const int world = peer->world;
  const unsigned long long* base = peer->ll[peer->rank] + (size_t)((seq & 1ull) * kMaxRanks) * (2 * kXchgDoubles);
  for (int x = threadIdx.x; x < world * count * 2; x += blockDim.x) {
    const int r = x / (count * 2), w = x - r * (count * 2);
    s_words[x] = ll_wait(base + (size_t)r * (2 * kXchgDoubles) + 2 * first + w, (uint32_t)seq);
  }
  __syncthreads();
  if ((int)threadIdx.x < count) {
    double v = 0.0;
    for (int r = 0; r < world; ++r) {
      const uint32_t lo = s_words[(r * count + threadIdx.x) * 2], hi = s_words[(r * count + threadIdx.x) * 2 + 1];
      v += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    }
    dst[threadIdx.x] = v;
  }
  __syncthreads();
}
// Device-side barrier of all ranks: one exchange of a single word (entry kXchgDoubles - 1 of the mailbox slot).
__global__ void __launch_bounds__(64) k_rank_barrier(const PeerTable* __restrict__ peer) {
  __shared__ double s_one[1];
  __shared__ uint32_t s_words[2 * kMaxRanks];
  const unsigned long long seq = *peer->xseq + 1ull;
  if (threadIdx.x == 0) s_one[0] = 1.0;
  __syncthreads();
  ll_send(peer, seq, s_one, kXchgDoubles - 1, 1);
  ll_gather(peer, seq, kXchgDoubles - 1, 1, s_words, s_one);
  if (threadIdx.x == 0) *peer->xseq = seq;
}
__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kLinThreads) : "memory");
}
__device__ __forceinline__ d3 ld3cg(const double* base, size_t ld, size_t i) {
  return mk3(__ldcg(base + i), __ldcg(base + ld + i), __ldcg(base + 2 * ld + i));
}
// All blocks of the (cooperative, fully resident) grid.  `next` = this block's count of arrivals the barrier has to
// reach, kept in shared memory; one arrival and one polling thread per block.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long* next, unsigned n_blocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long target = *next + n_blocks;
    *next = target;
    __threadfence();
    atomicAdd(bar, 1ull);
    while (ld_acquire_gpu(bar) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

template <int K, int ROWS>
__global__ void __launch_bounds__(kLoopThreads, 1)
    k_icp_loop(MapView mv, FactorView fv, DevState* ds_g, LoopArgs la, const PeerTable* __restrict__ peer) {
  extern __shared__ __align__(16) unsigned char s_loop_raw[];
  LoopShared<ROWS>& S = *reinterpret_cast<LoopShared<ROWS>*>(s_loop_raw);
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5, grp = tid >> 7, gt = tid & (kLinThreads - 1), gw = wib & (kLinWarps - 1);
  const unsigned n_blocks = gridDim.x;
  LoopCtl* const ctl = fv.ctl;
#if defined(MB_LOOP_TIMING)
  if (blockIdx.x == 0 && tid == 0) {
    unsigned long long gt0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
    g_loop_t[63][0] = (long long)gt0;
  }
#endif
  if (tid == 0) S.bar_next = (ld_acquire_gpu(&ctl->bar) / n_blocks) * n_blocks;  // nobody passes barrier #1 before this block arrives
  fill_scan_table(mv, S.tab);
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(ds_g);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&S.ds);
    for (int w = tid; w < (int)(sizeof(DevState) / 8); w += kLoopThreads) dst[w] = src[w];
  }
  __syncthreads();
  if (la.has_pose && tid < 16) reinterpret_cast<double*>(&S.ds)[tid] = la.pose.v[tid];  // pose (12), gravity (3), lambda lead DevState
  const unsigned long long xseq0 = peer ? *peer->xseq : 0ull;
  const unsigned long long epass0 = __ldcg(&ctl->epass);  // (block 0 advances it only after every block's first final pass)
  double(*s_row)[8] = S.rows[wib][0];
  const int pr = kTriRow[lane], pc = kTriCol[lane];
  const int k = fv.k;
  const bool forced = (fv.flags & 1u) != 0;
  const double inv_sigma = 1.0 / fv.sigma;
  const double kh_sigma_lo = fv.kh * fv.sigma * 0.9999999;  // |e| at or below this: |e / sigma| <= k for certain
  const int tile_pts = fv.tile;
  const size_t n_tiles = (fv.n + tile_pts - 1) / tile_pts;
  const size_t vg = (size_t)blockIdx.x * kLoopGroups + grp, n_vg = (size_t)n_blocks * kLoopGroups;
  __syncthreads();
  // A thread's points of the most recent linearisation: localizability vectors and "Valid", kept in SHARED memory
  // for the component-localizability sums (:434-457) that need that linearisation's eigenvectors — taken in the
  // next linearisation's phase A (device-resident loop) or after the last one.  They live in the group's block-probe
  // scratch, which only phase B uses (written in C / C', read before the next B), and in the head of S.pk.  Only
  // when a group owns a single tile; otherwise those passes read the points' state back from memory.
  // (Keeping them in registers was built and measured: 416 B of spills and every phase 15-30 % slower.)
  const bool one_tile = n_tiles <= n_vg;
  double* const k_loc = reinterpret_cast<double*>(S.blk[grp]) + gt;  // [(c * kPpt + u) * 128], c < 3 rot, c >= 3 trans
  uint8_t* const k_valid = reinterpret_cast<uint8_t*>(S.pk[grp]) + gt;  // [u * 128]
  static_assert(sizeof(uint32_t) * 24 * kLinThreads >= sizeof(double) * 6 * kPpt * kLinThreads, "kept vectors fit S.blk");
  static_assert(sizeof(uint32_t) * ROWS * kLinThreads >= kPpt * kLinThreads, "kept flags fit S.pk");
  bool kept = false;  // uniform over the group: the scratch holds its tile's vectors (phase B overwrites the scratch)

  for (int it0 = 0;; it0 += la.iters) {  // one pass; serve mode: one pass per request
  for (int it = it0; it < it0 + la.iters; ++it) {
    const int par = it & 1;
    MB_LOOP_T(it, 0);
    const uint32_t stamp = (uint32_t)(la.linearize_count0 + it + 1);
    if (blockIdx.x == 0 && tid == 0) ctl->q_count[par ^ 1] = ctl->q_next[par ^ 1] = 0u;  // the next linearisation's counters: idle during this one
    m33 R;
#pragma unroll
    for (int a = 0; a < 9; ++a) R.m[a] = S.ds.pose[a];
    const d3 T = mk3(S.ds.pose[9], S.ds.pose[10], S.ds.pose[11]);
    const double* const Vt = S.ds.lin.eigvec_trans;  // of the previous linearisation (folded pass)
    const double* const Vr = S.ds.lin.eigvec_rot;
    // [J e]^T [J e] accumulates on the FP64 tensor core: with M = the 8 x 32 matrix of a warp's rows (row 7 zero) the
    // sum is M M^T, eight mma.m8n8k4 per 32 points, and because the product is symmetric the A and B fragments of a
    // lane are the SAME element — M[lane >> 2][4 kk + (lane & 3)] — so a lane needs one shared-memory load per four
    // points.  (Lane a summing entry a over the rows in a scalar loop pulled 64 doubles per lane and slot through
    // shared memory: 1.8 us of a 6.5 us phase, bandwidth bound.)  c0 / c1 = C[lane >> 2][2 (lane & 3) + {0, 1}].
    double c0 = 0.0, c1 = 0.0, lacc = 0.0;
    int cnt = 0;

    // C: residual, Jacobian, accumulation of a thread's kPpt points (independent chains, interleaved by the compiler).
    // st carries kFresh when phase B has just fitted the point's plane.
    auto finish = [&](const size_t (&i)[kPpt], const bool (&act)[kPpt], const d3 (&ps)[kPpt], const d3 (&pt)[kPpt], uint8_t (&st)[kPpt]) {
      double row[kPpt][7];
      bool go[kPpt];
      d3 mean[kPpt], normal[kPpt];
      double rr[kPpt];
#pragma unroll
      for (int u = 0; u < kPpt; ++u) {
#pragma unroll
        for (int a = 0; a < 7; ++a) row[u][a] = 0.0;
        const bool fresh = (st[u] & kFresh) != 0;
        st[u] &= (uint8_t)~kFresh;
        go[u] = act[u] && (fresh ? st[u] == MB_UNPROCESSED : st[u] > MB_CORRES_PLANE_INVALID);
        mean[u] = normal[u] = mk3(0, 0, 0);
        rr[u] = 1.0;
        if (go[u]) {
          mean[u] = ld3cg(fv.mean, fv.ld, i[u]);
          normal[u] = ld3cg(fv.normal, fv.ld, i[u]);
          rr[u] = __ldg(fv.rroot + i[u]);
        }
      }
      MB_LOOP_F(it, 4);
      // branch-free per point (selects instead of branches), so that the two points' fp64 chains interleave
      d3 lrot[kPpt], ltrans[kPpt];
      bool valid[kPpt];
#pragma unroll
      for (int u = 0; u < kPpt; ++u) {
        // (a skipped point computes on harmless non-zero stand-ins: a zero numerator would send the whole warp through
        // the division's slow path)
        double e = go[u] ? dot3(normal[u], sub3(mean[u], pt[u])) : 1.0;
        // s-check (:323-326): 1 - 0.9 |e| / rr < 0.9.  With x = 0.9 |e| the outcome is certain without the division
        // unless x / rr lies within 1e-7 of 0.1 (the quotient's rounding moves it by 1e-17): only then is it formed.
        const double x9 = 0.9 * fabs(e);
        bool max_err = x9 > rr[u] * 0.1000001;
        if (!max_err && !(x9 < rr[u] * 0.0999999)) max_err = 1 - x9 / rr[u] < 0.9;
        valid[u] = go[u] && !max_err;
        // Huber (:330-336): |e / sigma| > k decides; the quotient itself is only needed beyond the threshold.
        double scale = inv_sigma;
        if (fv.use_huber) {  // uniform
          const double ae = fabs(e);
          if (ae > kh_sigma_lo) {  // rare: at or beyond the threshold (exact test inside)
            const double we = e / fv.sigma;
            if (fabs(we) > fv.kh) scale = sqrt(fv.kh / fabs(we)) / fv.sigma;
          }
        }
        e *= scale;
        const d3 ns = mul33Tv(R, normal[u]);
        const d3 jr = go[u] ? cross3(ns, ps[u]) : mk3(1.0, 1.0, 1.0);
        const double z = sqnorm3(jr);
        const double zr = z > 0 ? sqrt(z) : 1.0;
        const d3 jn = div3(jr, zr);
        lrot[u] = z > 0 ? jn : jr;
        ltrans[u] = mk3(-ns.x, -ns.y, -ns.z);
        row[u][0] = valid[u] ? jr.x * scale : 0.0;
        row[u][1] = valid[u] ? jr.y * scale : 0.0;
        row[u][2] = valid[u] ? jr.z * scale : 0.0;
        row[u][3] = valid[u] ? -ns.x * scale : 0.0;
        row[u][4] = valid[u] ? -ns.y * scale : 0.0;
        row[u][5] = valid[u] ? -ns.z * scale : 0.0;
        row[u][6] = valid[u] ? e : 0.0;
        if (go[u]) st[u] = valid[u] ? MB_VALID : MB_MAX_ERROR;
      }
#pragma unroll
      for (int u = 0; u < kPpt; ++u) {
        if (valid[u]) {
          st3(fv.loc_rot, fv.ld, i[u], lrot[u]);
          st3(fv.loc_trans, fv.ld, i[u], ltrans[u]);
        }
        if (act[u]) fv.status[i[u]] = st[u];
        if (one_tile) {
          kept = true;
          k_loc[(0 * kPpt + u) * kLinThreads] = lrot[u].x;
          k_loc[(1 * kPpt + u) * kLinThreads] = lrot[u].y;
          k_loc[(2 * kPpt + u) * kLinThreads] = lrot[u].z;
          k_loc[(3 * kPpt + u) * kLinThreads] = ltrans[u].x;
          k_loc[(4 * kPpt + u) * kLinThreads] = ltrans[u].y;
          k_loc[(5 * kPpt + u) * kLinThreads] = ltrans[u].z;
          k_valid[u * kLinThreads] = valid[u] ? 1 : 0;
        }
      }
      MB_LOOP_F(it, 5);
      // rows -> shared memory -> tensor-core fragments
#pragma unroll
      for (int u = 0; u < kPpt; ++u) {
        if (u == 1) MB_LOOP_F(it, 6);
        if (__ballot_sync(kFull, act[u] && st[u] == MB_VALID)) {  // (rows of points that are not Valid are zero)
          double(*r)[8] = S.rows[wib][u];
#pragma unroll
          for (int a = 0; a < 7; ++a) r[lane][a] = row[u][a];
          r[lane][7] = 0.0;
          __syncwarp();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const double m = r[4 * kk + (lane & 3)][lane >> 2];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(m), "d"(m));
          }
          __syncwarp();
        }
#pragma unroll
        for (int s = 0; s < 9; ++s) {
          const int c = __popc(__ballot_sync(kFull, act[u] && st[u] == s));
          if (lane == s) cnt += c;
        }
      }
    };
    auto write_partials = [&]() {
      S.cmat[wib][lane >> 2][2 * (lane & 3)] = c0;
      S.cmat[wib][lane >> 2][2 * (lane & 3) + 1] = c1;
      __syncwarp();
      if (lane < 28) S.red[wib][lane] = S.cmat[wib][pr][pc];
      if (lane >= 30) S.red[wib][lane + 8] = S.red[wib][lane + 16] = 0.0;  // pad entries 38, 39, 46, 47
      if (lane < 10) S.red[wib][kPackCnt + lane] = (double)cnt;
      if (lane < 6) S.red[wib][kPackLoc + lane] = lacc;
      __syncthreads();
      if (tid < kPack) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kLoopWarps; ++w) v += S.red[w][tid];
        __stcg(fv.partials + (size_t)blockIdx.x * kPack + tid, v);
      }
    };

    // ---- A (+ C for the tiles that keep every association) -----------------------------------------------------
    // Thread gt of the group owns points gt and gt + 128 of the tile (the second only in 256-point tiles).
    for (size_t tile = vg; tile < n_tiles; tile += n_vg) {
      MB_LOOP_F(it, 0);
      size_t i[kPpt];
      bool act[kPpt], need[kPpt];
      d3 ps[kPpt], pt[kPpt];
      uint8_t st[kPpt];
#pragma unroll
      for (int u = 0; u < kPpt; ++u) {
        i[u] = tile * tile_pts + u * kLinThreads + gt;
        act[u] = u * kLinThreads + gt < tile_pts && i[u] < fv.n;
        ps[u] = pt[u] = mk3(0, 0, 0);
        st[u] = MB_UNPROCESSED;
        need[u] = false;
        if (act[u]) {
          const float4 s = __ldg(fv.src + i[u]);
          ps[u] = mk3((double)s.x, (double)s.y, (double)s.z);
          pt[u] = add3(mul33v(R, ps[u]), T);
          st[u] = __ldcg(fv.status + i[u]);
          const d3 da = ld3cg(fv.p_da, fv.ld, i[u]);
          need[u] = forced || sqnorm3(sub3(pt[u], da)) >= fv.da_gate_sq;
        }
      }
      MB_LOOP_F(it, 1);
      // The points' share of the PREVIOUS linearisation's component localizabilities (:434-457).  (The first
      // linearisation of a launch has no predecessor whose sums anyone reads.)
      if (fv.fold_loc && it > 0) {
        double v[kPpt][6];
        bool was_valid[kPpt];
#pragma unroll
        for (int u = 0; u < kPpt; ++u) {
#pragma unroll
          for (int a = 0; a < 6; ++a) v[u][a] = 0.0;
          was_valid[u] = kept ? k_valid[u * kLinThreads] != 0 : (act[u] && st[u] == MB_VALID);
          if (was_valid[u]) {
            d3 lt, lr;
            if (kept) {
              lr = mk3(k_loc[(0 * kPpt + u) * kLinThreads], k_loc[(1 * kPpt + u) * kLinThreads], k_loc[(2 * kPpt + u) * kLinThreads]);
              lt = mk3(k_loc[(3 * kPpt + u) * kLinThreads], k_loc[(4 * kPpt + u) * kLinThreads], k_loc[(5 * kPpt + u) * kLinThreads]);
            } else {
              lt = ld3cg(fv.loc_trans, fv.ld, i[u]);
              lr = ld3cg(fv.loc_rot, fv.ld, i[u]);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {  // column a of V: (V^T loc)_a, in mul33Tv's order of operations
              const double tc = fabs(Vt[a] * lt.x + (Vt[3 + a] * lt.y + Vt[6 + a] * lt.z));
              const double rc = fabs(Vr[a] * lr.x + (Vr[3 + a] * lr.y + Vr[6 + a] * lr.z));
              v[u][a] = tc >= 0.5 ? tc : 0.0;
              v[u][3 + a] = rc >= 0.5 ? rc : 0.0;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kPpt; ++u) {
          if (__ballot_sync(kFull, was_valid[u])) {
#pragma unroll
            for (int a = 0; a < 6; ++a) s_row[lane][a] = v[u][a];
            __syncwarp();
            if (lane < 6) {
#pragma unroll 8
              for (int p = 0; p < 32; ++p) lacc += s_row[p][lane];
            }
            __syncwarp();
          }
        }
      }
      MB_LOOP_F(it, 2);
      unsigned mask[kPpt];
      int warp_need = 0;
#pragma unroll
      for (int u = 0; u < kPpt; ++u) {
        mask[u] = __ballot_sync(kFull, need[u]);
        warp_need += __popc(mask[u]);
      }
      if (lane == 0) S.warp_need[grp][gw] = warp_need;
      group_sync(grp);
      int base = 0, n_need = 0;
#pragma unroll
      for (int w = 0; w < kLinWarps; ++w) {
        if (w < gw) base += S.warp_need[grp][w];
        n_need += S.warp_need[grp][w];
      }
      MB_LOOP_F(it, 3);
      if (n_need == 0) {
        finish(i, act, ps, pt, st);
        MB_LOOP_F(it, 7);
      } else {
        if (gt == 0) {
          S.qbase[grp] = atomicAdd(&ctl->q_count[par], (unsigned)n_need);
          fv.tile_stamp[tile] = stamp;
        }
        group_sync(grp);
#pragma unroll
        for (int u = 0; u < kPpt; ++u) {
          if (need[u]) {
            fv.queue[S.qbase[grp] + (unsigned)(base + __popc(mask[u] & ((1u << lane) - 1)))] = (uint32_t)i[u];
            st3(fv.p_da, fv.ld, i[u], pt[u]);
          }
          base += __popc(mask[u]);
        }
        if (lane == 9) cnt += warp_need;
      }
      group_sync(grp);  // warp_need / qbase are rewritten by the next tile
    }
    MB_LOOP_T(it, 1);
    write_partials();
    MB_LOOP_T(it, 2);
    grid_barrier(&ctl->bar, &S.bar_next, n_blocks);
    MB_LOOP_T(it, 3);
    const unsigned q = __ldcg(&ctl->q_count[par]);
    if (q) {
      // ---- B: the queue, dealt out evenly over every warp of the device ----------------------------------------
      kept = false;  // the search scratch is about to be overwritten
      // Tasks of `per` consecutive queue entries, dealt out over the SMs first (task t -> warp t / n_blocks of block
      // t % n_blocks).  A short queue is spread thin — one query per warp runs 27 us, thirty-two 42 us — and once the
      // queue exceeds a round of the device the tasks are full warps: a pass executes almost the same instructions for
      // 14 queries as for 28 (measured: tasks of half the length made a full phase B 40 % slower).
      const unsigned W = n_blocks * kLoopWarps;
      const unsigned per = min(32u, max(1u, (q + W - 1u) / W));
      const unsigned n_tasks = (q + per - 1u) / per;
      // A warp's first task is its own number; the tasks beyond the first round are PULLED (a warp whose first task was
      // cheap takes more of them: a task's cost varies 2x with the voxels its queries fall into).
      for (unsigned task = (unsigned)wib * n_blocks + blockIdx.x; task < n_tasks;) {
        const unsigned qi = task * per + lane;
        const bool on = (unsigned)lane < per && qi < q;
        const size_t i = on ? (size_t)__ldcg(fv.queue + qi) : 0;
        const float4 s = __ldg(fv.src + i);
        const d3 pt = add3(mul33v(R, mk3((double)s.x, (double)s.y, (double)s.z)), T);
        double bd[K];
        uint32_t bs[K];
        uint32_t* s_pk = S.pk[grp] + gt;
        knn_thread<K>(mv, S.tab, s_pk, S.blk[grp] + gt, kLinThreads, pt.x, pt.y, pt.z, k, on, bd, bs);
        if (on) {
          float4 nb[K];
          uint64_t g[K];
          const int found = knn_resolve_all<K, true>(mv, s_pk, kLinThreads, bs, k, g, nb);
          double dk = 0.0;
#pragma unroll
          for (int j = 0; j < K; ++j) {
            if (j < k) {
              if (j == k - 1) dk = bd[j];
              // indices are only meaningful when all k exist (the reference discards partial results)
              if (fv.knn_idx) fv.knn_idx[i * k + j] = found == k ? g[j] : ~0ull;
            }
          }
          uint8_t rs = MB_UNPROCESSED;
          if (found != k) {
            rs = MB_INSUFFICIENT_CORRES_POINTS;
          } else if (dk > fv.max_corr_sq) {
            rs = MB_CORRES_MAX_DIST;
          } else {
            bool normal_set;
            d3 mean = mk3(0, 0, 0), normal = mk3(0, 0, 0);
            rs = fit_plane<K>(nb, k, T, fv.pvd, mean, normal, normal_set);
            st3(fv.mean, fv.ld, i, mean);
            if (normal_set) st3(fv.normal, fv.ld, i, normal);
          }
          fv.status[i] = (uint8_t)(rs | kFresh);
        }
        __syncwarp();
        unsigned nxt = 0;
        if (lane == 0) nxt = W + atomicAdd(&ctl->q_next[par], 1u);
        task = __shfl_sync(kFull, nxt, 0);
      }
      MB_LOOP_T(it, 4);
      grid_barrier(&ctl->bar, &S.bar_next, n_blocks);
      MB_LOOP_T(it, 5);
      // ---- C': the deferred tiles --------------------------------------------------------------------------------
      for (size_t tile = vg; tile < n_tiles; tile += n_vg) {
        if (__ldcg(fv.tile_stamp + tile) != stamp) continue;  // uniform over the group
        size_t i[kPpt];
        bool act[kPpt];
        d3 ps[kPpt], pt[kPpt];
        uint8_t st[kPpt];
#pragma unroll
        for (int u = 0; u < kPpt; ++u) {
          i[u] = tile * tile_pts + u * kLinThreads + gt;
          act[u] = u * kLinThreads + gt < tile_pts && i[u] < fv.n;
          ps[u] = pt[u] = mk3(0, 0, 0);
          st[u] = MB_UNPROCESSED;
          if (act[u]) {
            const float4 s = __ldg(fv.src + i[u]);
            ps[u] = mk3((double)s.x, (double)s.y, (double)s.z);
            pt[u] = add3(mul33v(R, ps[u]), T);
            st[u] = __ldcg(fv.status + i[u]);
          }
        }
        finish(i, act, ps, pt, st);
      }
      __syncthreads();  // S.red: the first write_partials' readers are long done, but keep the phases apart
      MB_LOOP_T(it, 6);
      write_partials();
      grid_barrier(&ctl->bar, &S.bar_next, n_blocks);
      MB_LOOP_T(it, 7);
    }
    // ---- D: packet, exchange, finalize — in every block ---------------------------------------------------------
    if (!peer) {
      block_sum_rows(fv.partials, (int)n_blocks, kPack, S.tmp, S.packed, kLoopThreads);
      __syncthreads();
      MB_LOOP_T(it, 8);
    } else {
      // Several ranks: block 0 alone talks to the peers — its packet goes straight into every rank's mailbox (peer
      // stores over NVLink, flag-in-data words: no fence, no separate flag), it waits for the other ranks' words in
      // this rank's mailbox, adds the packets in rank order (bit-identical on every rank) and publishes the sum to the
      // other blocks of this GPU through fv.packed and a gpu-scope flag.  Measured per exchange at N = 2: every block
      // polling mailbox flags at system scope 9-14 us, block 0 alone with fence + flags 7-8 us.
      const unsigned long long seq = xseq0 + 1ull + (unsigned long long)it;
      if (blockIdx.x == 0) {
        block_sum_rows(fv.partials, (int)n_blocks, kPack, S.tmp, S.packed, kLoopThreads);
        __syncthreads();
        MB_LOOP_T(it, 8);
        ll_send(peer, seq, S.packed, 0, kPack);
        ll_gather(peer, seq, 0, kPack, reinterpret_cast<uint32_t*>(S.tmp), S.packed);  // (includes this rank's own packet)
        if (tid < kPack) __stcg(fv.packed + tid, S.packed[tid]);
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_gpu(&ctl->xflag, seq);
      } else {
        if (tid == 0) {
          while (ld_acquire_gpu(&ctl->xflag) < seq) {
          }
        }
        __syncthreads();
        if (tid < kPack) S.packed[tid] = __ldcg(fv.packed + tid);
        __syncthreads();
      }
    }
    MB_LOOP_T(it, 10);
    {
      FinArgs fa;
      fa.reg_4_dof = la.reg_4_dof;
      fa.linearize_count = la.linearize_count0 + it + 1;
      fa.do_step = la.do_step;
      fa.iter = it;
      fa.trace = blockIdx.x == 0 ? la.trace : nullptr;
      // warp 0, lane 0: projection + packing + solve + retract; warp 1, lanes 0..3: the four eigen roles in lock step
      const int role = wib == 0 ? (lane == 0 ? 4 : -1) : (wib == 1 && lane < 4 ? lane : -1);
      if (role >= 0) finalize_role(S.packed, &S.ds, &S.ds, fa, role);
    }
    __syncthreads();
    MB_LOOP_T(it, 9);
  }

  // ---- after the last linearisation: its component localizabilities (:434-457), results out ------------------------
  {
    double a6[6] = {0, 0, 0, 0, 0, 0};
    const double* const Vt = S.ds.lin.eigvec_trans;
    const double* const Vr = S.ds.lin.eigvec_rot;
    if (kept) {  // the thread's own points, from the group's scratch
#pragma unroll
      for (int u = 0; u < kPpt; ++u) {
        if (k_valid[u * kLinThreads]) {
          const d3 lr = mk3(k_loc[(0 * kPpt + u) * kLinThreads], k_loc[(1 * kPpt + u) * kLinThreads], k_loc[(2 * kPpt + u) * kLinThreads]);
          const d3 lt = mk3(k_loc[(3 * kPpt + u) * kLinThreads], k_loc[(4 * kPpt + u) * kLinThreads], k_loc[(5 * kPpt + u) * kLinThreads]);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const double tc = fabs(Vt[a] * lt.x + (Vt[3 + a] * lt.y + Vt[6 + a] * lt.z));
            const double rc = fabs(Vr[a] * lr.x + (Vr[3 + a] * lr.y + Vr[6 + a] * lr.z));
            a6[a] += tc >= 0.5 ? tc : 0.0;
            a6[3 + a] += rc >= 0.5 ? rc : 0.0;
          }
        }
      }
    } else {
      for (size_t tile = vg; tile < n_tiles; tile += n_vg) {
#pragma unroll
        for (int u = 0; u < kPpt; ++u) {
          const size_t i = tile * tile_pts + u * kLinThreads + gt;
          if (u * kLinThreads + gt < tile_pts && i < fv.n && __ldcg(fv.status + i) == MB_VALID) {
            const d3 lt = ld3cg(fv.loc_trans, fv.ld, i), lr = ld3cg(fv.loc_rot, fv.ld, i);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const double tc = fabs(Vt[a] * lt.x + (Vt[3 + a] * lt.y + Vt[6 + a] * lt.z));
              const double rc = fabs(Vr[a] * lr.x + (Vr[3 + a] * lr.y + Vr[6 + a] * lr.z));
              a6[a] += tc >= 0.5 ? tc : 0.0;
              a6[3 + a] += rc >= 0.5 ? rc : 0.0;
            }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double v = a6[a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
      if (lane == 0) S.red[wib][a] = v;
    }
    __syncthreads();
    // Block sums -> block 0, WITHOUT a grid barrier: every block leaves its six sums as flag-in-data words (two 8-byte
    // words {32 data bits, pass number} per double, as in the peer mailbox) and goes on — to the next request, or home;
    // block 0 waits for all blocks' words and adds them in block order.
    const unsigned long long pass = epass0 + (unsigned long long)(it0 / la.iters) + 1ull;
    unsigned long long* const ll2 = reinterpret_cast<unsigned long long*>(fv.partials2);
    if (tid < 12) {
      double v = 0.0;
      for (int w = 0; w < kLoopWarps; ++w) v += S.red[w][tid >> 1];
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
      const unsigned long long word = ((pass & 0xffffffffull) << 32) | ((bits >> (32 * (tid & 1))) & 0xffffffffull);
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ll2 + (size_t)blockIdx.x * 16 + tid), "l"(word) : "memory");
    }
    if (blockIdx.x == 0) {
    __syncthreads();  // S.red is re-used for the gathered words
    uint32_t* const s_words = reinterpret_cast<uint32_t*>(&S.red[0][0]);  // n_blocks * 12 words (S.red and S.tmp are adjacent)
    static_assert(sizeof(S.red) + sizeof(S.tmp) >= 160 * 12 * sizeof(uint32_t), "gathered words fit");
    for (int x = tid; x < (int)n_blocks * 12; x += kLoopThreads) {
      const unsigned long long* src = ll2 + (size_t)(x / 12) * 16 + (x % 12);
      unsigned long long w;
      do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
      } while ((uint32_t)(w >> 32) != (uint32_t)pass);
      s_words[x] = (uint32_t)w;
    }
    __syncthreads();
    double* const loc = S.packed + kPack;  // [8]
    {
      // component e: eight chains over consecutive blocks, then the chains in order (a fixed order: bitwise repeatable)
      __shared__ double s_chain[6][8];
      const int per_chain = ((int)n_blocks + 7) / 8;
      if (tid < 48) {
        const int e = tid >> 3, c = tid & 7;
        double v = 0.0;
        for (int b = c * per_chain; b < min((c + 1) * per_chain, (int)n_blocks); ++b) {
          const uint32_t lo = s_words[b * 12 + 2 * e], hi = s_words[b * 12 + 2 * e + 1];
          v += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
        }
        s_chain[e][c] = v;
      }
      __syncthreads();
      if (tid < 8) {
        double v = 0.0;
        if (tid < 6)
          for (int c = 0; c < 8; ++c) v += s_chain[tid][c];
        loc[tid] = v;
      }
      if (tid == 0) ctl->epass = pass;
      if (tid < 2) ctl->q_count[tid] = ctl->q_next[tid] = 0u;  // every block has finished its last linearisation
    }
    __syncthreads();
    if (peer) {
      // the six sums of every rank, through the same mailbox slot as the last packet (entries 48..53)
      const unsigned long long seq = xseq0 + (unsigned long long)(it0 + la.iters);
      ll_send(peer, seq, loc, kPack, 6);
      ll_gather(peer, seq, kPack, 6, reinterpret_cast<uint32_t*>(S.tmp), loc);
      if (tid == 0) *peer->xseq = seq;
      __syncthreads();
    }
    if (la.mapped) {
      // Hand the finished linearisation to the polling host: flag-in-data words (mb_internal.cuh: kSrvOut), one store
      // per word, no fence, no separate flag (with a fence and a flag the hand-over alone was ~3 us longer).
      unsigned long long* const out = reinterpret_cast<unsigned long long*>(la.mapped + kSrvOut);
      constexpr int kWords = (int)(sizeof(mb_linearization) / 8);
      const unsigned long long* lin = reinterpret_cast<const unsigned long long*>(&S.ds.lin);
      const uint32_t req = (uint32_t)(la.req0 + (unsigned long long)(it0 / la.iters));
      for (int x = tid; x < 2 * (int)kSrvOutDoubles; x += kLoopThreads) {
        const int d = x >> 1;
        const unsigned long long bits = d < kWords       ? lin[d]
                                        : d < kWords + 6 ? (unsigned long long)__double_as_longlong(loc[d - kWords])
                                                         : (unsigned long long)__double_as_longlong(S.ds.pose[d - kWords - 6]);
        ll_store(out + x, (uint32_t)(bits >> (32 * (x & 1))), req);
      }
#if defined(MB_LOOP_TIMING)
      if (tid == 0) {
        unsigned long long g0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        g_loop_t[62][4] = g_loop_t[62][0];  // previous response
        g_loop_t[62][0] = (long long)g0;
      }
#endif
    }
    // device-side copies (after the host has its answer): DevState (pose, last linearisation), packet, component
    // localizabilities
    {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(ds_g);
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&S.ds);
      for (int w = tid; w < (int)(sizeof(DevState) / 8); w += kLoopThreads) dst[w] = src[w];
      if (tid < kPack) fv.packed[tid] = S.packed[tid];
      if (tid < 8) fv.loc_out[tid] = tid < 6 ? loc[tid] : 0.0;
    }
    __syncthreads();  // (resident: warp 0 writes the next request's pose into DevState)
#if defined(MB_LOOP_TIMING)
    if (tid == 0) {
      unsigned long long gt1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
      g_loop_t[63][1] = (long long)gt1;
    }
#endif
    }  // block 0
  }
  if (!la.serve) return;
  // ---- resident: wait for the next request's pose (block 0 asks the host, the other blocks ask block 0) -----------------
  {
    const unsigned long long next = la.req0 + (unsigned long long)(it0 / la.iters) + 1ull;
    __shared__ int s_go;
    if (blockIdx.x == 0) {
      // Warp 0 polls the request record with ONE coalesced 256-byte read per round (one PCIe round trip fetches the
      // request number, the pose and the stop word together; thread 0 reading them one after the other was measured:
      // 12.8 us for the 16 doubles of the pose alone).  The record is four 64-byte lines, each 7 payload words + the
      // request number in its last word, written payload first: a line that shows the number is complete.
      if (wib == 0) {
        const volatile unsigned long long* rec = reinterpret_cast<const volatile unsigned long long*>(la.mapped + kSrvRec);
        unsigned long long t0, t1, w;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        int go = 0;
        for (;;) {
          w = rec[lane];
          const bool line_ok = __shfl_sync(kFull, w, 7) == next && __shfl_sync(kFull, w, 15) == next && __shfl_sync(kFull, w, 23) == next;
          if (line_ok) {
            go = 1;
            break;
          }
          if (__shfl_sync(kFull, w, 24) >= la.req0) break;  // stop word
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (__shfl_sync(kFull, t1, 0) - t0 > la.window_ns) break;
        }
#if defined(MB_LOOP_TIMING)
        if (lane == 0) {
          unsigned long long g1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
          g_loop_t[62][1] = (long long)g1;
        }
#endif
        if (go) {
          // payload word of lane l = double l - l / 8 of [pose (12), gravity (3), lambda], which lead DevState
          const int idx = lane - (lane >> 3);
          if ((lane & 7) != 7 && idx < 16) reinterpret_cast<unsigned long long*>(ds_g)[idx] = w;
        } else if (lane == 0) {
          *reinterpret_cast<volatile unsigned long long*>(la.mapped + kSrvExit) = next;  // request `next` will not be served
          __threadfence_system();
        }
        __syncwarp();
#if defined(MB_LOOP_TIMING)
        if (lane == 0) {
          unsigned long long g2;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g2));
          g_loop_t[62][2] = (long long)g2;
        }
#endif
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          st_release_gpu(&ctl->sflag, 2ull * next + (go ? 0ull : 1ull));
          s_go = go;
        }
      }
    } else if (tid == 0) {
      unsigned long long v;
      while ((v = ld_acquire_gpu(&ctl->sflag)) < 2ull * next) {
      }
      s_go = (v & 1ull) == 0ull;
    }
    __syncthreads();
    if (!s_go) return;
    if (tid < 16) reinterpret_cast<double*>(&S.ds)[tid] = __ldcg(reinterpret_cast<const double*>(ds_g) + tid);
    __syncthreads();
#if defined(MB_LOOP_TIMING)
    if (blockIdx.x == 0 && tid == 0) {
      unsigned long long g3;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g3));
      g_loop_t[62][3] = (long long)g3;
    }
#endif
  }
  }  // requests
}

// device scan records -> float4 source points of the factor (xyz only), zero padded to ld
__global__ void k_pack_src(const unsigned char* __restrict__ data, size_t stride, size_t begin, size_t n, size_t ld,
                           float4* __restrict__ src) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n) {
    const float* f = (const float*)(data + (begin + i) * stride);
    v = make_float4(f[0], f[1], f[2], 0.f);
  }
  src[i] = v;
}

// ---- voxel order of the scan --------------------------------------------------------------------------------
// k_linearize wants the 32 points of a warp in as few map voxels as possible (mb_search_group.cuh).  Before a
// factor's FIRST linearisation (and after a reset) its points are grouped by the map voxel they fall into under that
// call's pose; later poses differ by centimetres, so the grouping stays good.  Grouping, not ordering, is what
// matters, and LiDAR scans arrive spatially coherent (consecutive records are neighbouring beams), so this is a
// STABLE COUNTING SORT INSIDE CHUNKS of 4096 consecutive points on an 11-bit hash of the voxel coordinates: one
// kernel, one block per chunk, no global passes (the benchmark scan: 2.1 voxels per warp against 1.9 for a global
// sort and 11.8 unsorted; a 63-bit radix sort of the scan costs 130 us on a B200, this kernel a few).  Voxels whose
// hashes collide merely interleave.  All per-point state lives in sorted order; `perm` (sorted position -> index in
// the caller's scan) brings it back to reference order in mb_factor_download_state (geometric_factor.hpp:79-84: the
// reference's arrays are indexed like the source cloud).  The result is a deterministic function of scan and pose.
constexpr int kGsThreads = 1024, kGsWarps = kGsThreads / 32;
constexpr int kGsRows = 4;                          // rows of 32 points per warp
constexpr int kGsChunk = kGsThreads * kGsRows;      // points per block
constexpr int kGsBins = 2048;
constexpr size_t kGsSmem = (size_t)kGsWarps * kGsBins * sizeof(uint16_t) + kGsBins * sizeof(uint32_t) + 64 * sizeof(uint32_t);
__global__ void __launch_bounds__(kGsThreads, 1)
    k_group_sort(const float4* __restrict__ src_raw, size_t n, const DevState* __restrict__ ds, PoseArg pa, int has_pose, double inv_leaf,
                 float4* __restrict__ src, uint32_t* __restrict__ perm, double* __restrict__ rroot) {
  extern __shared__ __align__(16) unsigned char s_gs[];
  uint16_t* hist = reinterpret_cast<uint16_t*>(s_gs);                                   // [warp][bin]
  uint32_t* total = reinterpret_cast<uint32_t*>(s_gs + (size_t)kGsWarps * kGsBins * 2);  // [bin], then the exclusive scan
  uint32_t* wsum = total + kGsBins;                                                     // [32] scan scratch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int x = tid; x < kGsWarps * kGsBins / 2; x += kGsThreads) reinterpret_cast<uint32_t*>(hist)[x] = 0u;
  m33 R;
#pragma unroll
  for (int a = 0; a < 9; ++a) R.m[a] = has_pose ? pa.v[a] : ds->pose[a];
  const d3 T = has_pose ? mk3(pa.v[9], pa.v[10], pa.v[11]) : mk3(ds->pose[9], ds->pose[10], ds->pose[11]);
  __syncthreads();
  const size_t chunk0 = (size_t)blockIdx.x * kGsChunk;
  uint16_t* const my_hist = hist + (size_t)warp * kGsBins;
  // 1. keys and the rank of every point among the earlier points of its warp with the same key
  float4 pt[kGsRows];
  uint32_t key[kGsRows], rank[kGsRows];
#pragma unroll
  for (int r = 0; r < kGsRows; ++r) {
    const size_t i = chunk0 + (size_t)warp * (32 * kGsRows) + r * 32 + lane;
    const bool on = i < n;
    pt[r] = on ? src_raw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const d3 w = add3(mul33v(R, mk3((double)pt[r].x, (double)pt[r].y, (double)pt[r].z)), T);
    key[r] = on ? (hash_coord(fast_floor(w.x * inv_leaf), fast_floor(w.y * inv_leaf), fast_floor(w.z * inv_leaf)) & (kGsBins - 1)) : 0xffffffffu;
    const unsigned peers = __match_any_sync(kFull, key[r]);
    const int leader = __ffs(peers) - 1;
    uint32_t before = 0u;
    if (on && lane == leader) {
      before = my_hist[key[r]];
      my_hist[key[r]] = (uint16_t)(before + (uint32_t)__popc(peers));
    }
    rank[r] = __shfl_sync(kFull, before, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // 2. per bin: exclusive prefix over the warps, total
  for (int b = tid; b < kGsBins; b += kGsThreads) {
    uint32_t run = 0u;
#pragma unroll 8
    for (int w = 0; w < kGsWarps; ++w) {
      const uint32_t c = hist[(size_t)w * kGsBins + b];
      hist[(size_t)w * kGsBins + b] = (uint16_t)run;
      run += c;
    }
    total[b] = run;
  }
  __syncthreads();
  // 3. exclusive scan of the totals (two bins per thread)
  {
    const uint32_t a0 = total[2 * tid], a1 = total[2 * tid + 1];
    uint32_t v = a0 + a1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(kFull, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
      uint32_t x = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(kFull, x, o);
        if (lane >= o) x += u;
      }
      wsum[lane] = x - wsum[lane];  // exclusive
    }
    __syncthreads();
    const uint32_t excl = wsum[warp] + v - (a0 + a1);
    total[2 * tid] = excl;
    total[2 * tid + 1] = excl + a0;
  }
  __syncthreads();
  // 4. scatter
#pragma unroll
  for (int r = 0; r < kGsRows; ++r) {
    const size_t i = chunk0 + (size_t)warp * (32 * kGsRows) + r * 32 + lane;
    if (i < n) {
      const size_t pos = chunk0 + total[key[r]] + my_hist[key[r]] + rank[r];
      src[pos] = pt[r];
      perm[pos] = (uint32_t)i;
      rroot[pos] = sqrt(sqrt(sqnorm3(mk3((double)pt[r].x, (double)pt[r].y, (double)pt[r].z))));
    }
  }
}

}  // namespace
}  // namespace mb

using namespace mb;

// The reference gates on sqrt(d2) > g (geometric_factor.hpp:281-283).  sqrt is monotone and correctly rounded on both
// sides, so there is a smallest double d2* with sqrt(d2*) > g and the test equals d2 >= d2*: found here once.
static double da_gate_sq_min(double g) {
  if (!(g >= 0.0)) return 0.0;  // sqrt(x) >= 0 > g for every x >= 0 (a NaN gate never fires either way: d2 >= NaN is false)
  double x = g * g;
  while (x > 0.0 && std::sqrt(x) > g) x = std::nextafter(x, 0.0);
  while (!(std::sqrt(x) > g)) x = std::nextafter(x, INFINITY);
  return x;
}

struct mb_factor {
  mb_ctx* ctx = nullptr;
  mb_map* map = nullptr;
  mb_icp_config cfg{};
  size_t n = 0, ld = 0, n_total = 0, begin = 0;
  // one pooled device block, carved into the arrays below
  void* block = nullptr;
  size_t block_bytes = 0;
  float4* src = nullptr;
  uint8_t* status = nullptr;
  double* vecs = nullptr;  // 5 SoA blocks of 3*ld doubles: p_da, mean, normal, loc_rot, loc_trans
  uint64_t* knn_idx = nullptr;
  double* partials = nullptr;
  double* gpartials = nullptr;
  double* partials2 = nullptr;
  double* packed = nullptr;  // kPack + 8 (loc_out)
  unsigned* tickets = nullptr;  // [0] final, [1] loc, [2..] groups
  DevState* ds = nullptr;
  // persistent loop (k_icp_loop): control words, re-association queue, per-tile stamps
  LoopCtl* ctl = nullptr;
  uint32_t* queue = nullptr;
  uint32_t* tile_stamp = nullptr;
  size_t n_tiles = 0;
  int loop_grid = 0;
  mb_icp_trace* d_trace = nullptr;
  int trace_cap = 0;
  int grid = 0, grid2 = 0, n_groups = 0;
  int tile_points = 128;  // points per block tile of k_linearize
  int loop_tile = 256;    // points per group tile of k_icp_loop
  // k_linearize launch shape: kernel variant (k == 5 and <= 19 neighbour voxels, or generic), staging pool per warp
  bool lin_small = true;
  // voxel order (see k_group_sort): the caller's points, sorted position -> caller's index
  float4* src_raw = nullptr;
  uint32_t* perm = nullptr;
  double* rroot = nullptr;  // sqrt(||p||) per sorted point
  bool sorted = false;
  void* raw = nullptr;  // the caller's records on the device while a factor is being built from page-locked memory
  size_t raw_bytes = 0;
  int linearize_count = 0;
  uint32_t flags = 0;
  // cached CUDA graph of an mb_icp_run sequence
  cudaGraphExec_t graph = nullptr;
  int graph_iters = 0, graph_count0 = -1;

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
    v.tile = tile_points;
    v.flags = flags;
    v.fold_loc = 0;
    const float da_gate_f = cfg.target_ivox_map_min_dist_in_voxel / 4;          // geometric_factor.hpp:283
    v.da_gate_sq = da_gate_sq_min((double)da_gate_f);
    v.rroot = rroot;
    const float max_corr_f = cfg.max_corres_distance * cfg.max_corres_distance;  // :299
    v.max_corr_sq = (double)max_corr_f;
    v.sigma = (double)cfg.lidar_point_noise_std_dev;
    v.kh = (double)cfg.huber_threshold;
    v.pvd = (double)cfg.plane_validity_distance;
    v.partials = partials;
    v.gpartials = gpartials;
    v.gtickets = tickets + 2;
    v.ticket = tickets;
    v.packed = packed;
    v.partials2 = partials2;
    v.ticket2 = tickets + 1;
    v.loc_out = packed + kPack;
    v.ctl = ctl;
    v.queue = queue;
    v.tile_stamp = tile_stamp;
    return v;
  }
};

namespace {

int reset_state(mb_factor* f) {
  cudaStream_t st = f->ctx->stream;
  if (f->n) {
    // status, the five vector blocks and the index block are contiguous: one memset for the zeros
    MB_CUDA(cudaMemsetAsync(f->vecs, 0, 15 * f->ld * sizeof(double) + f->ld, st));
    MB_CUDA(cudaMemsetAsync(f->knn_idx, 0xff, f->n * f->cfg.num_corres_points * sizeof(uint64_t), st));
    MB_CUDA(cudaMemsetAsync(f->tile_stamp, 0, f->n_tiles * sizeof(uint32_t), st));  // stamps count from 1 again
  }
  f->linearize_count = 0;
  f->sorted = false;  // the next first linearisation re-sorts under its own pose (all state is zero again)
  return MB_OK;
}

// Before the first linearisation: group the scan by map voxel under that call's pose (k_group_sort).
int enqueue_sort(mb_factor* f, const PoseArg* pose_arg) {
  if (f->n == 0) return MB_OK;
  mb_ctx* c = f->ctx;
  cudaStream_t st = c->stream;
  static bool opted_in = false;
  if (!opted_in) {
    MB_CUDA(cudaFuncSetAttribute(k_group_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGsSmem));
    opted_in = true;
  }
  const unsigned blocks = (unsigned)((f->n + kGsChunk - 1) / kGsChunk);
  PoseArg pa;
  std::memset(&pa, 0, sizeof(pa));
  if (pose_arg) pa = *pose_arg;
  k_group_sort<<<blocks, kGsThreads, kGsSmem, st>>>(f->src_raw, f->n, f->ds, pa, pose_arg ? 1 : 0, f->map->inv_leaf, f->src, f->perm, f->rroot);
  ++c->launches;
  f->sorted = true;
  MB_CUDA(cudaGetLastError());
  return MB_OK;
}

// Several ranks WITHOUT the peer-memory exchange: one linearisation (+ optional GN step) as separate kernels on the
// context stream, the 48-double packet all-reduced with NCCL between them (north_star's "single NCCL allreduce of the
// 6x6 normal equations per ICP iteration").
int enqueue_linearize(mb_factor* f, int do_step, int iter, mb_icp_trace* d_trace, int linearize_count) {
  mb_ctx* c = f->ctx;
  cudaStream_t st = c->stream;
  if (linearize_count == 1) MB_TRY(enqueue_sort(f, nullptr));  // first linearisation since construction / reset
  FactorView fv = f->view();
  // Device-resident loop: the localizability pass of this linearisation is folded into the NEXT k_linearize (and
  // mb_icp_run launches k_loc_comp once, after the last iteration); the host-facing single call runs it right away.
  fv.fold_loc = do_step ? 1 : 0;
  FinArgs fa;
  fa.reg_4_dof = (int)f->cfg.reg_4_dof;
  fa.linearize_count = linearize_count;
  fa.do_step = do_step;
  fa.iter = iter;
  fa.trace = d_trace;
  const dim3 grid(f->grid), block(kLinThreads);
  if (f->lin_small)  // the specialised kernel: k = 5 and at most 19 neighbour voxels
    MB_CUDA(launch_pdl(k_linearize<5, 19>, grid, block, st, f->map->view(), fv, f->ds));
  else
    MB_CUDA(launch_pdl(k_linearize<MB_MAX_K, 27>, grid, block, st, f->map->view(), fv, f->ds));
  MB_NCCL(ncclAllReduce(f->packed, f->packed, kPack, ncclDouble, ncclSum, c->comm, st));
  MB_CUDA(launch_pdl(k_finalize, dim3(1), dim3(64), st, (const double*)f->packed, f->ds, fa, 31u));
  c->launches += 2;
  if (!do_step) {
    MB_CUDA(launch_pdl(k_loc_comp, dim3(f->grid2), dim3(kLocThreads), st, fv, (const DevState*)f->ds));
    ++c->launches;
    MB_NCCL(ncclAllReduce(f->packed + kPack, f->packed + kPack, 6, ncclDouble, ncclSum, c->comm, st));
  }
  MB_CUDA(cudaGetLastError());
  return MB_OK;
}

// The localizability pass of the loop's LAST linearisation (the earlier ones were folded into their successors).
int enqueue_last_loc_comp(mb_factor* f) {
  mb_ctx* c = f->ctx;
  MB_CUDA(launch_pdl(k_loc_comp, dim3(f->grid2), dim3(kLocThreads), c->stream, f->view(), (const DevState*)f->ds));
  MB_NCCL(ncclAllReduce(f->packed + kPack, f->packed + kPack, 6, ncclDouble, ncclSum, c->comm, c->stream));
  ++c->launches;
  return MB_OK;
}

// The persistent loop serves a single GPU and several GPUs with the peer-memory exchange set up; several ranks WITHOUT
// it (ncclAllReduce between the kernels) run the per-linearisation kernels.
bool use_loop(const mb_ctx* c) { return c->world == 1 || c->d_peer != nullptr; }

// One cooperative launch of the persistent loop: `iters` linearisations (each followed by the harness GN step when
// do_step), starting from the pose in DevState or, host-facing call (request number `req`), in `pose_arg`; the
// host-facing launch hands its result to the polling host and, when `serve`, stays resident for the context's window.
int enqueue_loop(mb_factor* f, int iters, int do_step, mb_icp_trace* d_trace, const PoseArg* pose_arg, unsigned long long req, bool serve) {
  mb_ctx* c = f->ctx;
  cudaStream_t st = c->stream;
  if (iters <= 0) return MB_OK;
  if (f->linearize_count == 0) MB_TRY(enqueue_sort(f, pose_arg));  // first linearisation since construction / reset
  MapView mv = f->map->view();
  FactorView fv = f->view();
  // The component localizabilities of the linearisations BEFORE the last one only ever reach the caller through the
  // trace: without a trace that pass is not run (the last linearisation's come out of the final pass either way).
  fv.fold_loc = do_step && d_trace ? 1 : 0;
  fv.tile = f->loop_tile;
  LoopArgs la;
  std::memset(&la, 0, sizeof(la));
  la.iters = iters;
  la.do_step = do_step;
  la.reg_4_dof = (int)f->cfg.reg_4_dof;
  la.linearize_count0 = f->linearize_count;
  la.has_pose = pose_arg ? 1 : 0;
  if (pose_arg) {
    la.pose = *pose_arg;
    la.mapped = (char*)c->pin_small;
    la.req0 = req;
    la.serve = serve ? 1 : 0;
    la.window_ns = (unsigned long long)c->srv_window_us * 1000ull;
  }
  la.trace = d_trace;
  const PeerTable* peer = c->world > 1 ? c->d_peer : nullptr;
  DevState* ds = f->ds;
  void* args[] = {&mv, &fv, &ds, &la, &peer};
  const void* kern;
  size_t smem;
  if (f->lin_small) {
    kern = (const void*)k_icp_loop<5, 19>;
    smem = sizeof(LoopShared<19>);
  } else {
    kern = (const void*)k_icp_loop<MB_MAX_K, 27>;
    smem = sizeof(LoopShared<27>);
  }
  static bool opted_in = false;
  if (!opted_in) {
    MB_CUDA(cudaFuncSetAttribute((const void*)k_icp_loop<5, 19>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LoopShared<19>)));
    MB_CUDA(cudaFuncSetAttribute((const void*)k_icp_loop<MB_MAX_K, 27>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LoopShared<27>)));
    opted_in = true;
  }
  MB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(f->loop_grid), dim3(kLoopThreads), args, smem, st));  // (a plain launch was measured: no faster)
  ++c->launches;
  return MB_OK;
}

void drop_graph(mb_factor* f) {
  if (f->graph) cudaGraphExecDestroy(f->graph);
  f->graph = nullptr;
  f->graph_iters = 0;
  f->graph_count0 = -1;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// Everything that can fail once the factor holds its map reference; on failure the caller releases the factor
// (device block, raw-scan block, map reference) in one place.
static int factor_init(mb_factor* f, mb_ctx* ctx, mb_map* map, const void* pts, const mb_scan* dscan, size_t n, size_t stride_bytes,
                       const mb_icp_config* cfg, size_t shard_begin, size_t shard_end) {
  cudaStream_t st = ctx->stream;
  f->map = map;
  map->refs.fetch_add(1);
  f->cfg = *cfg;
  f->n_total = n;
  f->begin = shard_begin;
  f->n = shard_end - shard_begin;
  f->ld = std::max<size_t>(align_up(f->n, 32), 32);
  const size_t k = cfg->num_corres_points;
  // k_linearize: blocks of 128 threads, one 128-point tile at a time, as many blocks as the device holds
  f->lin_small = k == 5 && map->n_off <= 19;
  int per_sm = MB_LIN_BLOCKS;
  if (f->lin_small)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_linearize<5, 19>, kLinThreads, 0);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_linearize<MB_MAX_K, 27>, kLinThreads, 0);
  per_sm = std::max(per_sm, 1);
  // Points per tile: the smallest of 32 / 64 / 128 whose tiles still fit the device in one round.  A full scan on one
  // GPU needs 128 (and two rounds); a shard of a scan on 4 or 8 GPUs gets 64 or 32, which spreads its points over all
  // SMs with fewer queries per warp — the only lever on the search's latency once every SM has a block.
  const size_t capacity = (size_t)ctx->sm_count * per_sm;
  f->tile_points = 128;
  for (int tp : {32, 64})
    if ((f->n + tp - 1) / tp <= capacity) {
      f->tile_points = tp;
      break;
    }
  const size_t n_tiles = (f->n + f->tile_points - 1) / f->tile_points;
  // k_icp_loop: groups of 128 threads, two points per thread in a full tile; again the smallest tile that still covers
  // the shard in one round of the device's groups
  {
    const size_t groups = (size_t)ctx->sm_count * kLoopGroups;
    const size_t per_group = (f->n + groups - 1) / groups;
    f->loop_tile = (int)std::min<size_t>(kPpt * kLinThreads, std::max<size_t>(32, (per_group + 31) / 32 * 32));
  }
  f->n_tiles = std::max(n_tiles, (f->n + f->loop_tile - 1) / f->loop_tile);
  f->loop_grid = ctx->sm_count;
  f->grid = (int)std::max<size_t>(1, std::min<size_t>(n_tiles, capacity));
  f->n_groups = (f->grid + kGroup - 1) / kGroup;
  f->grid2 = (int)std::max<size_t>(1, std::min<size_t>((f->n + kLocThreads - 1) / kLocThreads, (size_t)ctx->sm_count * 4));

  // carve one block: [src | src_raw | perm | vecs (15 ld doubles) | status (ld bytes) | knn_idx | partials |
  //                   gpartials | partials2 | packed | tickets | DevState]
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_src = take(f->ld * sizeof(float4));
  const size_t o_raw = take(f->ld * sizeof(float4));
  const size_t o_perm = take(f->ld * sizeof(uint32_t));
  const size_t o_rroot = take(f->ld * sizeof(double));
  const size_t o_vecs = take(15 * f->ld * sizeof(double) + f->ld);  // status follows the vectors directly
  const size_t o_idx = take(f->ld * k * sizeof(uint64_t));
  const size_t o_queue = take(f->ld * sizeof(uint32_t));
  const size_t o_par = take((size_t)std::max(f->grid, f->loop_grid) * kPack * sizeof(double));
  const size_t o_gpar = take((size_t)f->n_groups * kPack * sizeof(double));
  const size_t o_par2 = take((size_t)std::max(f->grid2, f->loop_grid) * 16 * sizeof(double));  // (k_icp_loop: 12 flag-in-data words per block)
  const size_t o_packed = take((kPack + 8) * sizeof(double));
  const size_t o_tick = take((2 + (size_t)f->n_groups) * sizeof(unsigned));  // final, loc, groups...
  const size_t o_ds = take(sizeof(DevState));
  const size_t o_ctl = take(sizeof(LoopCtl));
  const size_t o_stamp = take(std::max<size_t>(n_tiles, 1) * sizeof(uint32_t));
  f->block_bytes = off;
  int rc = dev_alloc(ctx, &f->block, f->block_bytes);
  if (rc != MB_OK) {
    return rc;
  }
  char* base = (char*)f->block;
  f->src = (float4*)(base + o_src);
  f->src_raw = (float4*)(base + o_raw);
  f->perm = (uint32_t*)(base + o_perm);
  f->rroot = (double*)(base + o_rroot);
  f->vecs = (double*)(base + o_vecs);
  f->status = (uint8_t*)(base + o_vecs + 15 * f->ld * sizeof(double));
  f->knn_idx = (uint64_t*)(base + o_idx);
  f->partials = (double*)(base + o_par);
  f->gpartials = (double*)(base + o_gpar);
  f->partials2 = (double*)(base + o_par2);
  f->packed = (double*)(base + o_packed);
  f->tickets = (unsigned*)(base + o_tick);
  f->ds = (DevState*)(base + o_ds);
  f->ctl = (LoopCtl*)(base + o_ctl);
  f->queue = (uint32_t*)(base + o_queue);
  f->tile_stamp = (uint32_t*)(base + o_stamp);

  bool staged = true;  // the context's pinned staging buffer was used: wait for the copy before returning
  bool copied_from_caller = false;
  if (dscan) {
    staged = false;
    // device-resident scan: pack xyz straight into the factor's float4 array
    k_pack_src<<<(unsigned)((f->ld + 255) / 256), 256, 0, st>>>(dscan->data, dscan->stride, shard_begin, f->n, f->ld, f->src_raw);
    ++ctx->launches;
  } else if (f->n > 0 && host_is_page_locked(pts)) {
    // The caller's records are page-locked (cudaHostAlloc / mb_host_register): the copy engine takes this
    // rank's block of them as it is and a kernel extracts xyz — no CPU pass over the scan, and nothing here
    // has to wait for the copy (the first linearisation follows it in stream order).
    const size_t raw_bytes = f->n * stride_bytes;
    rc = dev_alloc(ctx, &f->raw, raw_bytes);
    if (rc != MB_OK) {
      return rc;
    }
    f->raw_bytes = raw_bytes;
    unsigned char* raw = (unsigned char*)f->raw;
    MB_CUDA(cudaMemcpyAsync(raw, (const char*)pts + shard_begin * stride_bytes, raw_bytes, cudaMemcpyHostToDevice, st));
    MB_CUDA(cudaEventRecord(ctx->ev1, st));  // the caller's buffer is free again once this copy has run (below)
    k_pack_src<<<(unsigned)((f->ld + 255) / 256), 256, 0, st>>>(raw, stride_bytes, 0, f->n, f->ld, f->src_raw);
    ++ctx->launches;
    dev_free(ctx, raw, raw_bytes);  // pooled: only ever handed out again to work on this same stream
    f->raw = nullptr;
    staged = false;
    copied_from_caller = true;
  } else {
    // host AoS with arbitrary stride -> page-locked float4 staging -> device (only xyz is read by the factor,
    // geometric_factor.hpp:277,323,346)
    rc = pinned_reserve(ctx, f->ld * sizeof(float4));
    if (rc != MB_OK) {
      return rc;
    }
    float4* h = (float4*)ctx->pinned;
    const char* rec = (const char*)pts + shard_begin * stride_bytes;
    size_t i = 0;
#if defined(__SSE2__)
    // records of >= 16 bytes: one unaligned 16-byte load per point, w masked to zero (the last record is done
    // scalar so that nothing is read past the end of the caller's buffer)
    if (stride_bytes >= 16 && f->n > 1) {
      const __m128 keep_xyz = _mm_castsi128_ps(_mm_set_epi32(0, -1, -1, -1));
      for (; i + 1 < f->n; ++i)
        _mm_stream_ps((float*)(h + i), _mm_and_ps(_mm_loadu_ps((const float*)(rec + i * stride_bytes)), keep_xyz));
      _mm_sfence();
    }
#endif
    for (; i < f->n; ++i) {
      const float* p = (const float*)(rec + i * stride_bytes);
      h[i] = make_float4(p[0], p[1], p[2], 0.f);
    }
    for (size_t z = f->n; z < f->ld; ++z) h[z] = make_float4(0.f, 0.f, 0.f, 0.f);
    MB_CUDA(cudaMemcpyAsync(f->src_raw, h, f->ld * sizeof(float4), cudaMemcpyHostToDevice, st));
  }
  // everything from `partials2` to the end of the block (flag-in-data block sums, packet, tickets, DevState, loop control
  // words, tile stamps) starts at zero
  MB_CUDA(cudaMemsetAsync(base + o_par2, 0, f->block_bytes - o_par2, st));
  MB_TRY(reset_state(f));
  if (staged) MB_CUDA(cudaStreamSynchronize(st));  // the pinned staging buffer is free again
  // Like the reference's constructor (geometric_factor.hpp:125 copies the cloud), the caller's buffer has been
  // consumed when this returns; the pack kernel, the memsets and whatever is enqueued next are not waited for.
  if (copied_from_caller) {
    cudaError_t q;
    while ((q = cudaEventQuery(ctx->ev1)) == cudaErrorNotReady) {
    }
    MB_CUDA(q);
  }
  return MB_OK;
}

static int factor_create_impl(mb_ctx* ctx, mb_map* map, const void* pts, const mb_scan* dscan, size_t n,
                              size_t stride_bytes, const mb_icp_config* cfg, size_t shard_begin, size_t shard_end,
                              mb_factor** out) {
  MB_REQUIRE(ctx && map && cfg && out, "null argument");
  MB_REQUIRE(n == 0 || pts || dscan, "null scan");
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
  server_stop(ctx);
  MB_CUDA(cudaSetDevice(ctx->device));
  mb_factor* f = new mb_factor;
  f->ctx = ctx;
  {
    const int mrc = ensure_mirror(map);  // the map is immutable from here on (refs > 1): build its search mirror once
    if (mrc != MB_OK) {
      delete f;
      return mrc;
    }
  }
  const int rc = factor_init(f, ctx, map, pts, dscan, n, stride_bytes, cfg, shard_begin, shard_end);
  if (rc != MB_OK) {
    mb_factor_release(f);  // also gives the map reference back
    return rc;
  }
  *out = f;
  return MB_OK;
}

extern "C" {

int mb_comm_barrier(mb_ctx* c) {
  MB_REQUIRE(c, "null ctx");
  if (c->world <= 1) return MB_OK;
  MB_REQUIRE(c->d_peer, "mb_comm_barrier needs the peer-memory exchange (mb_comm_ipc_open)");
  server_stop(c);
  MB_CUDA(cudaSetDevice(c->device));
  k_rank_barrier<<<1, 64, 0, c->stream>>>(c->d_peer);
  ++c->launches;
  MB_CUDA(cudaGetLastError());
  return MB_OK;
}

int mb_factor_create(mb_ctx* ctx, mb_map* map, const void* pts, size_t n, size_t stride_bytes,
                     const mb_icp_config* cfg, size_t shard_begin, size_t shard_end, mb_factor** out) {
  return factor_create_impl(ctx, map, pts, nullptr, n, stride_bytes, cfg, shard_begin, shard_end, out);
}

int mb_factor_create_from_scan(mb_ctx* ctx, mb_map* map, mb_scan* scan, const mb_icp_config* cfg, size_t shard_begin,
                               size_t shard_end, mb_factor** out) {
  MB_REQUIRE(scan, "null scan");
  MB_REQUIRE(scan->ctx == ctx, "scan belongs to another context");
  return factor_create_impl(ctx, map, nullptr, scan, scan->n, scan->stride, cfg, shard_begin, shard_end, out);
}

// Timing diagnostic (not part of the documented ABI): average device time of k_finalize restricted to the roles
// in `role_mask` (bit 7: with the harness GN step in role 4), `reps` back-to-back launches on the factor's last packet.
MB_API int mb_debug_time_finalize(mb_factor* f, unsigned role_mask, int reps, float* us_per_launch) {
  MB_REQUIRE(f && us_per_launch && reps > 0, "bad argument");
  server_stop(f->ctx);
  MB_CUDA(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  FinArgs fa;
  fa.reg_4_dof = (int)f->cfg.reg_4_dof, fa.linearize_count = 0, fa.do_step = (role_mask & 128u) ? 1 : 0, fa.iter = 0, fa.trace = nullptr;
  role_mask &= 127u;
  auto one = [&]() { k_finalize<<<1, 64, 0, st>>>(f->packed, f->ds, fa, role_mask); };
  for (int w = 0; w < 3; ++w) one();
  MB_CUDA(cudaEventRecord(f->ctx->ev0, st));
  for (int r = 0; r < reps; ++r) one();
  MB_CUDA(cudaEventRecord(f->ctx->ev1, st));
  MB_CUDA(cudaEventSynchronize(f->ctx->ev1));
  float ms = 0.f;
  MB_CUDA(cudaEventElapsedTime(&ms, f->ctx->ev0, f->ctx->ev1));
  *us_per_launch = ms * 1e3f / reps;
  return MB_OK;
}

#if defined(MB_LOOP_TIMING)
MB_API int mb_debug_loop_times(long long* out /* 64 x 12 */) {
  return cudaMemcpyFromSymbol(out, g_loop_t, sizeof(long long) * 64 * 12) == cudaSuccess ? MB_OK : MB_ERR_CUDA;
}
MB_API int mb_debug_loop_fine(long long* out /* 64 x 12 */) {
  return cudaMemcpyFromSymbol(out, g_loop_f, sizeof(long long) * 64 * 12) == cudaSuccess ? MB_OK : MB_ERR_CUDA;
}
#endif

int mb_factor_release(mb_factor* f) {
  if (!f) return MB_OK;
  server_stop(f->ctx);
  if (f->ctx->srv_factor == f) f->ctx->srv_factor = nullptr;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  drop_graph(f);
  dev_free(f->ctx, f->block, f->block_bytes);
  dev_free(f->ctx, f->raw, f->raw_bytes);
  if (f->d_trace) dev_free(f->ctx, f->d_trace, (size_t)f->trace_cap * sizeof(mb_icp_trace));
  mb_map_release(f->map);
  delete f;
  return MB_OK;
}

int mb_factor_reset(mb_factor* f) {
  MB_REQUIRE(f, "null factor");
  server_stop(f->ctx);
  MB_CUDA(cudaSetDevice(f->ctx->device));
  return reset_state(f);
}

int mb_factor_set_flags(mb_factor* f, uint32_t flags) {
  MB_REQUIRE(f, "null factor");
  server_stop(f->ctx);  // the flags travel in the launch parameters
  if (flags != f->flags) drop_graph(f);
  f->flags = flags;
  return MB_OK;
}

int mb_factor_linearize(mb_factor* f, const double R[9], const double t[3], const double gravity_unit[3],
                        mb_linearization* out) {
  MB_REQUIRE(f && R && t && gravity_unit && out, "null argument");
  mb_ctx* c = f->ctx;
  MB_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  // page-locked, device-mapped block: [pose 12 | gravity 3 | lambda] in at 0, [mb_linearization | loc 6] out at 256,
  // request / response numbers at kSrv* (mb_internal.cuh)
  double* hin = (double*)c->pin_small;
  char* hout = (char*)c->pin_small + 256;  // (NCCL mode: plain copies land here)
  char local_out[sizeof(mb_linearization) + 6 * sizeof(double)];
  const char* result = hout;
  static_assert(sizeof(mb_linearization) % 8 == 0, "mb_linearization is copied as 8-byte words");
  if (use_loop(c)) {
    // Single GPU, or several with the peer-memory exchange set up.  No copy operation and no stream synchronisation on
    // this path: the pose travels in the kernel parameters, block 0 writes the result straight into mapped host memory
    // as flag-in-data words the host polls.  The kernel stays resident for a short
    // window afterwards (srv_window_us) and a call that arrives inside it only POSTS its pose to the mapped block — no
    // launch (measured: 41 us per cached call with a launch each, of which the device works 26).
    const unsigned long long n = ++c->srv_req;
    auto* gone = (const unsigned long long*)((const char*)c->pin_small + kSrvExit);
    unsigned long long res[kSrvOutDoubles];
    // Poll for request n's result: busy for the first ~100 us (a call normally ends well inside that), then yielding
    // the core between polls; the stream is queried now and then so that a failed launch ends the wait with its error,
    // and a kernel that never answers ends it after kPollTimeoutS instead of hanging the caller (mimosa's node has
    // three threads that may sit here, mimosa_node.cpp:28-43).  1: served, 0: the resident kernel left without
    // serving it, < 0: error.
    auto wait = [&](bool posted) -> int {
      constexpr double kPollTimeoutS = 30.0;
      const auto t_begin = std::chrono::steady_clock::now();
      unsigned spins = 0;
      bool yielding = false;
      for (;;) {
        if (read_result(c, n, res)) return 1;
        if (posted && __atomic_load_n(gone, __ATOMIC_ACQUIRE) >= n) return read_result(c, n, res) ? 1 : 0;
        if ((++spins & 0x3ffu) == 0) {
          const cudaError_t q = cudaStreamQuery(st);
          if (q == cudaSuccess) return read_result(c, n, res) ? 1 : 0;
          if (q != cudaErrorNotReady) {
            set_error("mb_factor_linearize: %s", cudaGetErrorString(q));
            return -1;
          }
          const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
          yielding = waited > 100e-6;
          if (waited > kPollTimeoutS) {
            set_error("mb_factor_linearize: no result after %.0f s (kernel lost or device hung)", kPollTimeoutS);
            return -1;
          }
        }
        if (yielding) {
          std::this_thread::yield();
        } else {
#if defined(__x86_64__)
          __builtin_ia32_pause();
#endif
        }
      }
    };
    int served = 0;
    if (c->srv_live && c->srv_factor == f) {
      double v[16];
      std::memcpy(v, R, 9 * sizeof(double));
      std::memcpy(v + 9, t, 3 * sizeof(double));
      std::memcpy(v + 12, gravity_unit, 3 * sizeof(double));
      v[15] = 0.0;
      unsigned long long* rec = (unsigned long long*)((char*)c->pin_small + kSrvRec);
      for (int i = 0; i < 16; ++i) std::memcpy(rec + i + i / 7, v + i, 8);  // payload first ...
      for (int line = 0; line < 3; ++line) __atomic_store_n(rec + 8 * line + 7, n, __ATOMIC_RELEASE);  // ... then the number
      served = wait(true);
      if (served < 0) return MB_ERR_CUDA;
      if (!served) c->srv_live = false;  // its window had closed: launch
    }
    if (!served) {
      server_stop(c);  // (a kernel resident for another factor)
      PoseArg pa;
      std::memcpy(pa.v, R, 9 * sizeof(double));
      std::memcpy(pa.v + 9, t, 3 * sizeof(double));
      std::memcpy(pa.v + 12, gravity_unit, 3 * sizeof(double));
      pa.v[15] = 0.0;
      const bool serve = c->srv_window_us > 0;  // (several ranks: every rank's caller posts to its own rank's kernel)
      MB_TRY(enqueue_loop(f, 1, 0, nullptr, &pa, n, serve));
      c->srv_live = serve;
      c->srv_factor = f;
      if (wait(false) != 1) {
        c->srv_live = false;
        return MB_ERR_CUDA;
      }
    }
    ++f->linearize_count;
    std::memcpy(local_out, res, sizeof(local_out));
    result = local_out;
  } else {
    ++f->linearize_count;
    std::memcpy(hin, R, 9 * sizeof(double));
    std::memcpy(hin + 9, t, 3 * sizeof(double));
    std::memcpy(hin + 12, gravity_unit, 3 * sizeof(double));
    MB_CUDA(cudaMemcpyAsync(f->ds->pose, hin, 15 * sizeof(double), cudaMemcpyHostToDevice, st));
    MB_TRY(enqueue_linearize(f, 0, 0, nullptr, f->linearize_count));
    MB_CUDA(cudaMemcpyAsync(hout, &f->ds->lin, sizeof(mb_linearization), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaMemcpyAsync(hout + sizeof(mb_linearization), f->packed + kPack, 6 * sizeof(double),
                            cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
  }
  std::memcpy(out, result, sizeof(mb_linearization));
  const double* loc = (const double*)(result + sizeof(mb_linearization));
  for (int a = 0; a < 3; ++a) {
    out->loc_trans_comp[a] = loc[a];
    out->loc_rot_comp[a] = loc[3 + a];
  }
  return MB_OK;
}

int mb_factor_download_state(mb_factor* f, uint8_t* status, double* p_da, double* mean, double* normal,
                             double* loc_rot, double* loc_trans, uint64_t* knn_idx) {
  MB_REQUIRE(f, "null factor");
  server_stop(f->ctx);
  MB_CUDA(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  const size_t n = f->n, k = f->cfg.num_corres_points;
  if (n == 0) return MB_OK;
  // the device arrays are in voxel order (k_group_sort); perm[s] = index of sorted position s in the caller's scan
  std::vector<uint32_t> perm(n);
  if (f->sorted) {
    MB_CUDA(cudaMemcpyAsync(perm.data(), f->perm, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  } else {
    for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
  }
  std::vector<uint8_t> st_s;
  std::vector<uint64_t> idx_s;
  if (status) {
    st_s.resize(n);
    MB_CUDA(cudaMemcpyAsync(st_s.data(), f->status, n, cudaMemcpyDeviceToHost, st));
  }
  if (knn_idx) {
    idx_s.resize(n * k);
    MB_CUDA(cudaMemcpyAsync(idx_s.data(), f->knn_idx, n * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  }
  MB_CUDA(cudaStreamSynchronize(st));
  if (status)
    for (size_t s = 0; s < n; ++s) status[perm[s]] = st_s[s];
  if (knn_idx)
    for (size_t s = 0; s < n; ++s) std::memcpy(knn_idx + (size_t)perm[s] * k, idx_s.data() + s * k, k * sizeof(uint64_t));
  double* outs[5] = {p_da, mean, normal, loc_rot, loc_trans};
  std::vector<double> soa;
  for (int v = 0; v < 5; ++v) {
    if (!outs[v]) continue;
    soa.resize(3 * f->ld);
    MB_CUDA(cudaMemcpyAsync(soa.data(), f->vecs + (size_t)v * 3 * f->ld, 3 * f->ld * sizeof(double),
                            cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    for (size_t s = 0; s < n; ++s)
      for (int c = 0; c < 3; ++c) outs[v][3 * (size_t)perm[s] + c] = soa[(size_t)c * f->ld + s];
  }
  return MB_OK;
}

int mb_icp_run(mb_factor* f, double R[9], double t[3], int iters, double lambda, mb_icp_trace* trace) {
  MB_REQUIRE(f && R && t, "null argument");
  MB_REQUIRE(iters >= 0 && iters <= 4096, "iters outside [0, 4096]");
  server_stop(f->ctx);
  MB_CUDA(cudaSetDevice(f->ctx->device));
  cudaStream_t st = f->ctx->stream;
  if (iters > f->trace_cap) {
    MB_CUDA(cudaStreamSynchronize(st));
    drop_graph(f);
    if (f->d_trace) dev_free(f->ctx, f->d_trace, (size_t)f->trace_cap * sizeof(mb_icp_trace));
    f->d_trace = nullptr;
    f->trace_cap = 0;
    MB_TRY(dev_alloc(f->ctx, (void**)&f->d_trace, (size_t)iters * sizeof(mb_icp_trace)));
    f->trace_cap = iters;
  }
  if (use_loop(f->ctx)) {
    // One launch, no copy operations: the start pose travels in the kernel parameters and block 0 leaves the final
    // pose and the last linearisation's component localizabilities in the context's mapped block.
    if (iters == 0) return MB_OK;
    PoseArg pa;
    std::memcpy(pa.v, R, 9 * sizeof(double));
    std::memcpy(pa.v + 9, t, 3 * sizeof(double));
    pa.v[12] = 0.0, pa.v[13] = 0.0, pa.v[14] = -1.0, pa.v[15] = lambda;
    const unsigned long long n = ++f->ctx->srv_req;
    MB_TRY(enqueue_loop(f, iters, 1, trace ? f->d_trace : nullptr, &pa, n, false));
    f->linearize_count += iters;
    if (trace) MB_CUDA(cudaMemcpyAsync(trace, f->d_trace, (size_t)iters * sizeof(mb_icp_trace), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaStreamSynchronize(st));
    unsigned long long words[kSrvOutDoubles];
    if (!read_result(f->ctx, n, words)) {
      set_error("mb_icp_run: the kernel ended without handing over its result");
      return MB_ERR_CUDA;
    }
    const double* res = (const double*)(words + sizeof(mb_linearization) / 8);  // loc (6), pose (12)
    if (trace) {
      for (int a = 0; a < 3; ++a) {
        trace[iters - 1].loc_trans_comp[a] = res[a];
        trace[iters - 1].loc_rot_comp[a] = res[3 + a];
      }
    }
    std::memcpy(R, res + 6, 9 * sizeof(double));
    std::memcpy(t, res + 15, 3 * sizeof(double));
    return MB_OK;
  }
  double* hin = (double*)f->ctx->pin_small;
  std::memcpy(hin, R, 9 * sizeof(double));
  std::memcpy(hin + 9, t, 3 * sizeof(double));
  hin[12] = 0.0;
  hin[13] = 0.0;
  hin[14] = -1.0;
  hin[15] = lambda;
  MB_CUDA(cudaMemcpyAsync(f->ds->pose, hin, 16 * sizeof(double), cudaMemcpyHostToDevice, st));
  const bool use_graph = (f->flags & 2u) != 0 && iters > 0;
  if (use_graph) {
    // The captured sequence bakes the iteration index and linearize_count into k_finalize's arguments.
    if (!f->graph || f->graph_iters != iters || f->graph_count0 != f->linearize_count) {
      drop_graph(f);
      cudaGraph_t g = nullptr;
      MB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int rc = MB_OK;
      for (int it = 0; it < iters && rc == MB_OK; ++it)
        rc = enqueue_linearize(f, 1, it, f->d_trace, f->linearize_count + it + 1);
      if (rc == MB_OK) rc = enqueue_last_loc_comp(f);
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
      f->ctx->launches += 2ull * iters + 1ull + (f->graph_count0 == 0 && f->n ? 1ull : 0ull);
    }
    MB_CUDA(cudaGraphLaunch(f->graph, st));
    f->linearize_count += iters;
  } else {
    for (int it = 0; it < iters; ++it) {
      ++f->linearize_count;
      MB_TRY(enqueue_linearize(f, 1, it, f->d_trace, f->linearize_count));
    }
    if (iters) MB_TRY(enqueue_last_loc_comp(f));
  }
  double* hout = (double*)((char*)f->ctx->pin_small + 256);
  MB_CUDA(cudaMemcpyAsync(hout, f->ds->pose, 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
  double* hloc = hout + 16;  // the last iteration's component localizabilities, from the trailing k_loc_comp
  if (trace && iters) {
    MB_CUDA(cudaMemcpyAsync(trace, f->d_trace, (size_t)iters * sizeof(mb_icp_trace), cudaMemcpyDeviceToHost, st));
    MB_CUDA(cudaMemcpyAsync(hloc, f->packed + kPack, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  MB_CUDA(cudaStreamSynchronize(st));
  if (trace && iters) {
    for (int a = 0; a < 3; ++a) {
      trace[iters - 1].loc_trans_comp[a] = hloc[a];
      trace[iters - 1].loc_rot_comp[a] = hloc[3 + a];
    }
  }
  std::memcpy(R, hout, 9 * sizeof(double));
  std::memcpy(t, hout + 9, 3 * sizeof(double));
  return MB_OK;
}

}  // extern "C"
