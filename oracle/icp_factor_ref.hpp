// ORACLE — TEST INFRASTRUCTURE ONLY (see linalg_ref.hpp for the rules).  PARITY UNPINNED: the
// reference has no tests or golden vectors for this path; the restatement is pinned only by the
// analytic known-answer tests in tests/test_oracle_kat.py.
//
// CPU restatement of mimosa::lidar::ICPFactor (unary form, the only one mimosa constructs:
// mimosa/src/lidar/geometric.cpp:194):
//   constructor state         mimosa/include/mimosa/lidar/geometric_factor.hpp:119-156
//   estimatePlane             geometric_factor.hpp:176-229
//   linearize                 geometric_factor.hpp:231-562
//   computeLocalizability     mimosa/include/mimosa/utils.hpp:308-313
//   getProjectionMatrix       mimosa/include/mimosa/lidar/utils.hpp:191-213
//   RegistrationConfig        mimosa/include/mimosa/lidar/geometric_config.hpp:17-33 (reals are float and
//                             are promoted to double exactly where the reference promotes them)
// `abs(e)` at geometric_factor.hpp:323,334,335 is taken as the floating-point overload (x86-64 behaviour
// with Eigen's headers in scope).  RAD2DEG/DEG2RAD are PCL's macros (x*57.29578,
// x*0.017453293).
#pragma once
#include <cstdint>
#include <vector>

#include "ivox_ref.hpp"

namespace mimosa_oracle {

struct RegistrationConfigRef {  // geometric_config.hpp:17-33, field order kept
  float source_voxel_grid_filter_leaf_size = 0.5f;
  float source_voxel_grid_min_dist_in_voxel = 0.1f;
  float target_ivox_map_leaf_size = 0.5f;
  float target_ivox_map_min_dist_in_voxel = 0.1f;
  uint64_t num_corres_points = 5;
  float max_corres_distance = 2.24f;
  float plane_validity_distance = 0.04f;
  float lidar_point_noise_std_dev = 0.02f;
  int32_t use_huber = 1;
  float huber_threshold = 1.345f;
  int32_t reg_4_dof = 0;
  int32_t project_on_degneneracy = 1;
  float degen_thresh_rot = 10.f;
  float degen_thresh_trans = 15.f;
};

enum RejectStatusRef : uint8_t {  // geometric_factor.hpp:35-46
  kUnprocessed = 0,
  kInsufficientCorresPoints,
  kCorresMaxDist,
  kEigenSolverFail,
  kMinEigenValueLow,
  kLine,
  kCorresPlaneInvalid,
  kMaxError,
  kValid
};

struct LinearizationRef {
  double H[36];  // J^T J, row-major
  double g[6];   // -J^T e  (the HessianFactor's linear term, geometric_factor.hpp:559-560)
  double f;      // sum e^2
  int64_t counts[9];
  double loc_trans_comp[3], loc_rot_comp[3], loc_trans_final[3], loc_rot_final[3];
  double eigvec_trans[9], eigvec_rot[9];  // row-major; column j belongs to eigenvalue j (ascending)
  double degen_rot[3], degen_trans[3], degen_eigvec_rot[9], degen_eigvec_trans[9];
  int32_t linearize_count;
  int32_t n_searched;  // points that redid data association in this call (not a reference output)
};

inline void compute_localizability(const M3& JtJ, double loc[3], M3& V) {  // utils.hpp:308-313
  double lam[3];
  eigh3(JtJ, lam, V);
  for (int i = 0; i < 3; ++i) loc[i] = std::sqrt(lam[i]);
}

// lidar/utils.hpp:191-213
inline bool get_projection_matrix(const double loc[3], double thresh, const M3& V, M3& P) {
  if (loc[0] > thresh && loc[1] > thresh && loc[2] > thresh) {
    P = m3_identity();
    return false;
  }
  P = m3_zero();
  for (int i = 0; i < 3; ++i)
    if (loc[i] > thresh)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) P(r, c) += V(r, i) * V(c, i);
  return true;
}

class IcpFactorRef {
 public:
  IcpFactorRef(std::shared_ptr<const IVoxRef> map, const void* pts, size_t n, size_t stride_bytes,
               const RegistrationConfigRef& cfg)
      : map_(std::move(map)), cfg_(cfg) {
    src_.resize(n);
    for (size_t i = 0; i < n; ++i) {
      const float* f = (const float*)((const char*)pts + i * stride_bytes);
      src_[i] = V3{(double)f[0], (double)f[1], (double)f[2]};
    }
    const V3 z{0, 0, 0};
    p_t_.assign(n, z);
    p_da_.assign(n, z);
    mean_.assign(n, z);
    normal_.assign(n, z);
    loc_rot_.assign(n, z);
    loc_trans_.assign(n, z);
    status_.assign(n, kUnprocessed);
    knn_idx_.assign(n * cfg.num_corres_points, ~0ull);
  }

  size_t size() const { return src_.size(); }
  const std::vector<uint8_t>& status() const { return status_; }
  const std::vector<V3>& p_da() const { return p_da_; }
  const std::vector<V3>& mean() const { return mean_; }
  const std::vector<V3>& normal() const { return normal_; }
  const std::vector<V3>& loc_rot() const { return loc_rot_; }
  const std::vector<V3>& loc_trans() const { return loc_trans_; }
  // Correspondence indices of the most recent data association of each point (~0 when the search
  // found fewer than k); kept by the oracle for parity checks only.
  const std::vector<uint64_t>& knn_idx() const { return knn_idx_; }
  int linearize_count() const { return count_; }

  void reset_state() {
    const V3 z{0, 0, 0};
    std::fill(p_da_.begin(), p_da_.end(), z);
    std::fill(mean_.begin(), mean_.end(), z);
    std::fill(normal_.begin(), normal_.end(), z);
    std::fill(status_.begin(), status_.end(), (uint8_t)kUnprocessed);
    count_ = 0;
  }

  // n_threads = 4 reproduces the reference (geometric_factor.hpp:261,273); chunks are libgomp's
  // static schedule.  `parallel` only decides whether the chunks really run concurrently.
  void linearize(const Pose& T, const V3& gravity_unit, LinearizationRef& out, int n_threads = 4,
                 bool parallel = true) {
    ++count_;
    const size_t n = src_.size();
    const int k = (int)cfg_.num_corres_points;
    const V3 origin = T.t;  // delta_pose * 0, geometric_factor.hpp:253
    const float da_gate_f = cfg_.target_ivox_map_min_dist_in_voxel / 4;          // :283 float quotient
    const double da_gate = (double)da_gate_f;
    const float max_corr_f = cfg_.max_corres_distance * cfg_.max_corres_distance;  // :299 float product
    const double max_corr_sq = (double)max_corr_f;
    const double sigma = (double)cfg_.lidar_point_noise_std_dev;
    const double kh = (double)cfg_.huber_threshold;
    const double pvd = (double)cfg_.plane_validity_distance;

    struct Acc {
      double H[36];
      double b[6];
      double f;
      int searched;
    };
    std::vector<Acc> acc(n_threads);
    for (auto& a : acc) std::memset(&a, 0, sizeof(Acc));

#pragma omp parallel for num_threads(n_threads) schedule(static, 1) if (parallel)
    for (int tid = 0; tid < n_threads; ++tid) {
      const size_t q = n / n_threads, r = n % n_threads;
      const size_t begin = tid * q + ((size_t)tid < r ? tid : r);
      const size_t end = begin + q + ((size_t)tid < r ? 1 : 0);
      Acc& A = acc[tid];
      std::vector<uint64_t> idx(k);
      std::vector<double> d2(k);
      std::vector<V3> P(k), C(k);
      for (size_t i = begin; i < end; ++i) {
        const V3 ps = src_[i];
        const V3 pt = mul(T.R, ps) + T.t;  // :276-277
        p_t_[i] = pt;
        if (norm(pt - p_da_[i]) > da_gate) {  // :281-287
          p_da_[i] = pt;
          status_[i] = kUnprocessed;  // :290
          ++A.searched;
          const int found = map_->knn(pt, k, idx.data(), d2.data());
          for (int j = 0; j < k; ++j) knn_idx_[i * k + j] = found == k ? idx[j] : ~0ull;
          if (found != k) {  // :294-298
            status_[i] = kInsufficientCorresPoints;
            continue;
          }
          if (d2[k - 1] > max_corr_sq) {  // :299-302
            status_[i] = kCorresMaxDist;
            continue;
          }
          // estimatePlane, :176-229
          for (int j = 0; j < k; ++j) P[j] = map_->point(idx[j]);
          V3 s{0, 0, 0};
          for (int j = 0; j < k; ++j) s = s + P[j];
          const V3 m = s / (double)k;
          mean_[i] = m;  // :191 — written before the gates
          for (int j = 0; j < k; ++j) C[j] = P[j] - m;
          M3 cov = m3_zero();
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
              double acc_ab = 0.0;
              for (int j = 0; j < k; ++j) acc_ab += (&C[j].x)[a] * (&C[j].x)[b];
              cov(a, b) = acc_ab / (double)(k - 1);
            }
          double lam[3];
          M3 V;
          if (!eigh3(cov, lam, V)) {  // :197-200
            status_[i] = kEigenSolverFail;
            continue;
          }
          if (lam[0] < 1e-6) {  // :202-206
            status_[i] = kMinEigenValueLow;
            continue;
          }
          if (lam[2] > 3 * lam[1]) {  // :209-213
            status_[i] = kLine;
            continue;
          }
          V3 nrm{V(0, 0), V(1, 0), V(2, 0)};                 // :215
          if (dot(nrm, origin - m) < 0) nrm = nrm * -1.0;   // :218-220
          normal_[i] = nrm;
          bool invalid = false;
          for (int j = 0; j < k; ++j)
            if (std::fabs(dot(C[j], nrm)) > pvd) invalid = true;  // :222-226
          if (invalid) {
            status_[i] = kCorresPlaneInvalid;
            continue;
          }
        } else if (status_[i] <= kCorresPlaneInvalid) {  // :314-316
          continue;
        }

        double e = dot(normal_[i], mean_[i] - pt);                             // :319
        const double s_chk = 1 - 0.9 * std::fabs(e) / std::sqrt(norm(ps));      // :322-323
        if (s_chk < 0.9) {                                                      // :325-328
          status_[i] = kMaxError;
          continue;
        }
        double sqrt_w = 1.0;
        if (cfg_.use_huber) {  // :331-337
          const double we = e / sigma;
          if (std::fabs(we) > kh) sqrt_w = std::sqrt(kh / std::fabs(we));
        }
        const double scale = sqrt_w / sigma;
        e *= scale;  // :339
        const V3 ns = mulT(T.R, normal_[i]);  // :343
        const V3 jr = cross(ns, ps);          // :345-347
        double J[6] = {jr.x, jr.y, jr.z, -ns.x, -ns.y, -ns.z};
        {  // :351-352, before whitening.  Eigen normalized(): v / sqrt(z) when z > 0.
          const double z = sqnorm(jr);
          loc_rot_[i] = z > 0 ? jr / std::sqrt(z) : jr;
          loc_trans_[i] = V3{J[3], J[4], J[5]};
        }
        for (int a = 0; a < 6; ++a) J[a] *= scale;  // :355
        for (int a = 0; a < 6; ++a) {               // :364-366
          for (int b = 0; b < 6; ++b) A.H[6 * a + b] += J[a] * J[b];
          A.b[a] += J[a] * e;
        }
        A.f += e * e;
        status_[i] = kValid;  // :385
      }
    }

    double H[36] = {0}, b[6] = {0}, f = 0;  // :389-403
    int searched = 0;
    for (int t = 0; t < n_threads; ++t) {
      for (int a = 0; a < 36; ++a) H[a] += acc[t].H[a];
      for (int a = 0; a < 6; ++a) b[a] += acc[t].b[a];
      f += acc[t].f;
      searched += acc[t].searched;
    }

    M3 Hrr, Hrt, Htr, Htt;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        Hrr(r, c) = H[6 * r + c];
        Hrt(r, c) = H[6 * r + 3 + c];
        Htr(r, c) = H[6 * (r + 3) + c];
        Htt(r, c) = H[6 * (r + 3) + 3 + c];
      }
    M3 Vr, Vt;
    compute_localizability(Hrr, out.loc_rot_final, Vr);    // :406-408
    compute_localizability(Htt, out.loc_trans_final, Vt);  // :409-411
    const M3 Srr = inv3(sub(Hrr, mul(mul(Hrt, inv3(Htt)), Htr)));  // :413-417
    const M3 Stt = inv3(sub(Htt, mul(mul(Htr, inv3(Hrr)), Hrt)));  // :418-422
    M3 Dr, Dt;
    compute_localizability(Srr, out.degen_rot, Dr);    // :425
    compute_localizability(Stt, out.degen_trans, Dt);  // :426
    for (int a = 0; a < 3; ++a) out.degen_rot[a] = out.degen_rot[a] * 57.29578;  // :428 (PCL RAD2DEG)

    double ltc[3] = {0, 0, 0}, lrc[3] = {0, 0, 0};  // :434-457 (serial, point order)
    for (size_t i = 0; i < n; ++i) {
      if (status_[i] != kValid) continue;
      const V3 tc = mulT(Vt, loc_trans_[i]);
      const V3 rc = mulT(Vr, loc_rot_[i]);
      const double ta[3] = {std::fabs(tc.x), std::fabs(tc.y), std::fabs(tc.z)};
      const double ra[3] = {std::fabs(rc.x), std::fabs(rc.y), std::fabs(rc.z)};
      for (int a = 0; a < 3; ++a) {
        ltc[a] += ta[a] >= 0.5 ? ta[a] : 0.0;
        lrc[a] += ra[a] >= 0.5 ? ra[a] : 0.0;
      }
    }

    if (cfg_.reg_4_dof) {  // :257-259, 464-475
      const V3 global_z = neg(gravity_unit);
      const V3 lz = mulT(T.R, global_z);
      M3 Pm;
      const double l[3] = {lz.x, lz.y, lz.z};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Pm(r, c) = l[r] * l[c];
      const M3 nrr = mul(mul(Pm, Hrr), Pm);
      const M3 nrt = mul(Pm, Hrt);
      const M3 ntr = mul(Htr, Pm);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          H[6 * r + c] = nrr(r, c);
          H[6 * r + 3 + c] = nrt(r, c);
          H[6 * (r + 3) + c] = ntr(r, c);
        }
      const V3 pb = mul(Pm, V3{b[0], b[1], b[2]});
      b[0] = pb.x;
      b[1] = pb.y;
      b[2] = pb.z;
    }

    if (cfg_.project_on_degneneracy) {  // :477-557, mirrored literally: the per-point arrays it re-sums
      M3 Prot, Ptrans;                  // (:270-271) are never written, so the result is all-zero.
      const bool rd = get_projection_matrix(out.loc_rot_final, (double)cfg_.degen_thresh_rot, Vr, Prot);
      const bool td = get_projection_matrix(out.loc_trans_final, (double)cfg_.degen_thresh_trans, Vt, Ptrans);
      if (rd || td) {
        for (int a = 0; a < 36; ++a) H[a] = 0.0;
        for (int a = 0; a < 6; ++a) b[a] = 0.0;
        compute_localizability(m3_zero(), out.loc_rot_final, Vr);
        compute_localizability(m3_zero(), out.loc_trans_final, Vt);
      }
    }

    for (int a = 0; a < 36; ++a) out.H[a] = H[a];
    for (int a = 0; a < 6; ++a) out.g[a] = -b[a];
    out.f = f;
    for (int a = 0; a < 9; ++a) out.counts[a] = 0;
    for (size_t i = 0; i < n; ++i) ++out.counts[status_[i]];
    for (int a = 0; a < 3; ++a) {
      out.loc_trans_comp[a] = ltc[a];
      out.loc_rot_comp[a] = lrc[a];
    }
    for (int a = 0; a < 9; ++a) {
      out.eigvec_rot[a] = Vr.m[a];
      out.eigvec_trans[a] = Vt.m[a];
      out.degen_eigvec_rot[a] = Dr.m[a];
      out.degen_eigvec_trans[a] = Dt.m[a];
    }
    out.linearize_count = count_;
    out.n_searched = searched;
  }

  // Per-point residual rows of the most recent linearisation (whitened e, whitened J) recomputed
  // from the cached state; for tests only.
 private:
  std::shared_ptr<const IVoxRef> map_;
  RegistrationConfigRef cfg_;
  std::vector<V3> src_, p_t_, p_da_, mean_, normal_, loc_rot_, loc_trans_;
  std::vector<uint8_t> status_;
  std::vector<uint64_t> knn_idx_;
  int count_ = 0;
};

struct IcpTraceRef {  // one Gauss-Newton harness iteration
  double H[36], g[6], f, delta[6];
  double R[9], t[3];  // pose AFTER the retract of this iteration
  int64_t counts[9];
  int32_t n_searched;
  int32_t solve_ok;
  double loc_trans_comp[3], loc_rot_comp[3];  // of this iteration's linearisation (geometric_factor.hpp:434-457)
};

// delta = (H + lambda I)^-1 g with g = -J^T e, T <- T * Exp(delta).
inline void icp_run_ref(IcpFactorRef& f, Pose& T, int iters, double lambda, IcpTraceRef* trace, int n_threads,
                        bool parallel) {
  const V3 g_unit{0, 0, -1};
  for (int it = 0; it < iters; ++it) {
    LinearizationRef L;
    f.linearize(T, g_unit, L, n_threads, parallel);
    double delta[6] = {0, 0, 0, 0, 0, 0};
    const bool ok = solve6_ldlt(L.H, lambda, L.g, delta);
    if (ok) T = se3_retract(T, delta);
    if (trace) {
      IcpTraceRef& tr = trace[it];
      for (int a = 0; a < 36; ++a) tr.H[a] = L.H[a];
      for (int a = 0; a < 6; ++a) {
        tr.g[a] = L.g[a];
        tr.delta[a] = delta[a];
      }
      tr.f = L.f;
      for (int a = 0; a < 9; ++a) {
        tr.R[a] = T.R.m[a];
        tr.counts[a] = L.counts[a];
      }
      tr.t[0] = T.t.x;
      tr.t[1] = T.t.y;
      tr.t[2] = T.t.z;
      tr.n_searched = L.n_searched;
      tr.solve_ok = ok ? 1 : 0;
      for (int a = 0; a < 3; ++a) {
        tr.loc_trans_comp[a] = L.loc_trans_comp[a];
        tr.loc_rot_comp[a] = L.loc_rot_comp[a];
      }
    }
  }
}

}  // namespace mimosa_oracle
