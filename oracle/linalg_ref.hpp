// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the product
// path (mimosa_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.
//
// PARITY UNPINNED: the reference (ntnu-arl/mimosa) has no tests, golden vectors or fixtures for
// this path, and none of its dependencies (Eigen 3.3.7 from Ubuntu focal, ntnu-arl/gtsam branch
// feature/imu_factor_with_gravity, ntnu-arl/gtsam_points branch minimal_updated — no commit pinned,
// .github/docker/ci-base.Dockerfile:7,39-41) exist in this environment.  What follows is a plain
// C++ restatement of the *published algorithms* those call sites rely on.
//
// Small dense linear algebra used by the restated ICP factor:
//  * eigh3(): symmetric 3x3 eigen-decomposition restating Eigen 3.3.7's
//    SelfAdjointEigenSolver<Matrix3d>::compute() (scale to [-1,1], closed-form 3x3 Householder
//    tridiagonalisation, implicit symmetric QR with Wilkinson shift, ascending selection sort) as
//    called at mimosa/include/mimosa/lidar/geometric_factor.hpp:196 and mimosa/include/mimosa/utils.hpp:310.
//  * inv3(): cofactor inverse as Eigen's fixed-size 3x3 .inverse() (geometric_factor.hpp:413-422).
//  * Expmap / retract restating gtsam::Pose3::Expmap (full SE(3) exponential, GTSAM_POSE3_EXPMAP=ON,
//    README.md:54) used by the stand-alone Gauss-Newton harness.
//
// Build with -ffp-contract=off: every +,-,*,/ and sqrt below is one IEEE-754 binary64 operation
// in source order, which is what the CUDA path reproduces with __dadd_rn/__dmul_rn.
#pragma once
#include <cmath>
#include <cfloat>
#include <limits>

namespace mimosa_oracle {

struct V3 {
  double x, y, z;
};
inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(const V3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 neg(const V3& a) { return {-a.x, -a.y, -a.z}; }
// Evaluation order x*x' + (y*y' + z*z'): Eigen 3.3.7 reduces a fixed-size 3-vector with
// redux_novec_unroller, which splits [0,3) into [0,1) and [1,3) (Eigen/src/Core/Redux.h).  This
// cannot be verified here (no Eigen in the image) — documented assumption, ulp-level effect only.
inline double dot(const V3& a, const V3& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline double sqnorm(const V3& a) { return dot(a, a); }
inline double norm(const V3& a) { return std::sqrt(sqnorm(a)); }
inline V3 cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Row-major 3x3.
struct M3 {
  double m[9];
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
};
inline M3 m3_zero() { return M3{{0, 0, 0, 0, 0, 0, 0, 0, 0}}; }
inline M3 m3_identity() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
// Coefficient-based fixed-size products: each output coefficient is a 3-term reduction in the same
// a0 + (a1 + a2) order as dot().
inline V3 mul(const M3& A, const V3& v) {
  return {A.m[0] * v.x + (A.m[1] * v.y + A.m[2] * v.z), A.m[3] * v.x + (A.m[4] * v.y + A.m[5] * v.z),
          A.m[6] * v.x + (A.m[7] * v.y + A.m[8] * v.z)};
}
inline V3 mulT(const M3& A, const V3& v) {  // A^T v
  return {A.m[0] * v.x + (A.m[3] * v.y + A.m[6] * v.z), A.m[1] * v.x + (A.m[4] * v.y + A.m[7] * v.z),
          A.m[2] * v.x + (A.m[5] * v.y + A.m[8] * v.z)};
}
inline M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C(r, c) = A(r, 0) * B(0, c) + (A(r, 1) * B(1, c) + A(r, 2) * B(2, c));
  return C;
}
inline M3 transpose(const M3& A) {
  M3 T;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T(r, c) = A(c, r);
  return T;
}
inline M3 sub(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] - B.m[i];
  return C;
}

// Eigen's compute_inverse_size3: cofactors scaled by 1/det.
inline M3 inv3(const M3& A) {
  const double c00 = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1);
  const double c10 = A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2);  // cofactor feeding result(0,1)
  const double c20 = A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1);
  const double det = A(0, 0) * c00 + A(1, 0) * c10 + A(2, 0) * c20;
  const double id = 1.0 / det;
  M3 R;
  R(0, 0) = c00 * id;
  R(0, 1) = c10 * id;
  R(0, 2) = c20 * id;
  R(1, 0) = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * id;
  R(1, 1) = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * id;
  R(1, 2) = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * id;
  R(2, 0) = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * id;
  R(2, 1) = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * id;
  R(2, 2) = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * id;
  return R;
}

// ---- Eigen 3.3.7 SelfAdjointEigenSolver<Matrix3d> restated --------------------------------------
namespace detail {
struct Givens {
  double c, s;
};
// JacobiRotation<double>::makeGivens(p, q) (real branch).
inline Givens make_givens(double p, double q) {
  Givens g;
  if (q == 0.0) {
    g.c = p < 0.0 ? -1.0 : 1.0;
    g.s = 0.0;
  } else if (p == 0.0) {
    g.c = 0.0;
    g.s = q < 0.0 ? 1.0 : -1.0;
  } else if (std::fabs(p) > std::fabs(q)) {
    const double t = q / p;
    double u = std::sqrt(1.0 + t * t);
    if (p < 0.0) u = -u;
    g.c = 1.0 / u;
    g.s = -t * g.c;
  } else {
    const double t = p / q;
    double u = std::sqrt(1.0 + t * t);
    if (q < 0.0) u = -u;
    g.s = -1.0 / u;
    g.c = -t * g.s;
  }
  return g;
}
// numext::hypot of Eigen 3.3.7 (its own formula, not libm's).
inline double eigen_hypot(double x, double y) {
  const double ax = std::fabs(x), ay = std::fabs(y);
  double p, qp;
  if (ax > ay) {
    p = ax;
    qp = ay / p;
  } else {
    p = ay;
    qp = ax / p;
  }
  if (p == 0.0) return 0.0;
  return p * std::sqrt(1.0 + qp * qp);
}
// internal::tridiagonal_qr_step for n = 3; Q is row-major 3x3 and receives Q <- Q * G.
inline void tridiagonal_qr_step(double* diag, double* subdiag, int start, int end, double* Q) {
  const double td = (diag[end - 1] - diag[end]) * 0.5;
  const double e = subdiag[end - 1];
  double mu = diag[end];
  if (td == 0.0) {
    mu -= std::fabs(e);
  } else {
    const double e2 = e * e;
    const double h = eigen_hypot(td, e);
    if (e2 == 0.0)
      mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
    else
      mu -= e2 / (td + (td > 0.0 ? h : -h));
  }
  double x = diag[start] - mu;
  double z = subdiag[start];
  for (int k = start; k < end; ++k) {
    const Givens rot = make_givens(x, z);
    const double sdk = rot.s * diag[k] + rot.c * subdiag[k];
    const double dkp1 = rot.s * subdiag[k] + rot.c * diag[k + 1];
    diag[k] = rot.c * (rot.c * diag[k] - rot.s * subdiag[k]) -
              rot.s * (rot.c * subdiag[k] - rot.s * diag[k + 1]);
    diag[k + 1] = rot.s * sdk + rot.c * dkp1;
    subdiag[k] = rot.c * sdk - rot.s * dkp1;
    if (k > start) subdiag[k - 1] = rot.c * subdiag[k - 1] - rot.s * z;
    x = subdiag[k];
    if (k < end - 1) {
      z = -rot.s * subdiag[k + 1];
      subdiag[k + 1] = rot.c * subdiag[k + 1];
    }
    // q.applyOnTheRight(k, k+1, rot): columns k, k+1 of Q.
    for (int r = 0; r < 3; ++r) {
      const double xi = Q[3 * r + k], yi = Q[3 * r + k + 1];
      Q[3 * r + k] = rot.c * xi - rot.s * yi;
      Q[3 * r + k + 1] = rot.s * xi + rot.c * yi;
    }
  }
}
}  // namespace detail

// Returns false on NoConvergence (RejectStatus::EigenSolverFail, geometric_factor.hpp:197-200).
// lam ascending; column j of V (V(r,j)) is the eigenvector of lam[j].  Only the lower triangle
// of A is read, as Eigen does.
inline bool eigh3(const M3& A, double lam[3], M3& V) {
  double a00 = A(0, 0), a10 = A(1, 0), a11 = A(1, 1), a20 = A(2, 0), a21 = A(2, 1), a22 = A(2, 2);
  double scale = std::fabs(a00);
  const double cand[5] = {a10, a11, a20, a21, a22};
  for (double c : cand)
    if (std::fabs(c) > scale) scale = std::fabs(c);
  if (scale == 0.0) scale = 1.0;
  a00 /= scale;
  a10 /= scale;
  a11 /= scale;
  a20 /= scale;
  a21 /= scale;
  a22 /= scale;

  double diag[3], subdiag[2];
  double Q[9];
  const double tol = std::numeric_limits<double>::min();
  diag[0] = a00;
  const double v1norm2 = a20 * a20;
  if (v1norm2 <= tol) {
    diag[1] = a11;
    diag[2] = a22;
    subdiag[0] = a10;
    subdiag[1] = a21;
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i) Q[i] = I[i];
  } else {
    const double beta = std::sqrt(a10 * a10 + v1norm2);
    const double invBeta = 1.0 / beta;
    const double m01 = a10 * invBeta;
    const double m02 = a20 * invBeta;
    const double q = 2.0 * m01 * a21 + m02 * (a22 - a11);
    diag[1] = a11 + m02 * q;
    diag[2] = a22 - m02 * q;
    subdiag[0] = beta;
    subdiag[1] = a21 - m01 * q;
    const double Qi[9] = {1, 0, 0, 0, m01, m02, 0, m02, -m01};
    for (int i = 0; i < 9; ++i) Q[i] = Qi[i];
  }

  const int n = 3, max_iter = 30;
  int end = n - 1, start = 0, iter = 0;
  const double consider_zero = std::numeric_limits<double>::min();
  const double precision = 2.0 * std::numeric_limits<double>::epsilon();
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (std::fabs(subdiag[i]) <= (std::fabs(diag[i]) + std::fabs(diag[i + 1])) * precision ||
          std::fabs(subdiag[i]) <= consider_zero)
        subdiag[i] = 0.0;
    while (end > 0 && subdiag[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > max_iter * n) break;
    start = end - 1;
    while (start > 0 && subdiag[start - 1] != 0.0) start--;
    detail::tridiagonal_qr_step(diag, subdiag, start, end, Q);
  }
  const bool ok = iter <= max_iter * n;
  if (ok) {
    for (int i = 0; i < n - 1; ++i) {
      int k = 0;
      double best = diag[i];
      for (int j = 1; j < n - i; ++j)
        if (diag[i + j] < best) {
          best = diag[i + j];
          k = j;
        }
      if (k > 0) {
        const double t = diag[i];
        diag[i] = diag[k + i];
        diag[k + i] = t;
        for (int r = 0; r < 3; ++r) {
          const double tq = Q[3 * r + i];
          Q[3 * r + i] = Q[3 * r + k + i];
          Q[3 * r + k + i] = tq;
        }
      }
    }
  }
  for (int i = 0; i < 3; ++i) lam[i] = diag[i] * scale;
  for (int i = 0; i < 9; ++i) V.m[i] = Q[i];
  return ok;
}

// ---- SE(3) -----------------------------------------------------------------------------------------
struct Pose {
  M3 R;
  V3 t;
};

// gtsam::SO3::Expmap (so3::ExpmapFunctor) in matrix form.
inline M3 so3_expmap(const V3& w) {
  const double theta2 = dot(w, w);
  const M3 W{{0, -w.z, w.y, w.z, 0, -w.x, -w.y, w.x, 0}};
  M3 R = m3_identity();
  if (theta2 <= std::numeric_limits<double>::epsilon()) {
    for (int i = 0; i < 9; ++i) R.m[i] += W.m[i];
    return R;
  }
  const double theta = std::sqrt(theta2);
  const double sin_theta = std::sin(theta);
  const double s2 = std::sin(theta / 2.0);
  const double one_minus_cos = 2.0 * s2 * s2;
  M3 K;
  for (int i = 0; i < 9; ++i) K.m[i] = W.m[i] / theta;
  const M3 KK = mul(K, K);
  for (int i = 0; i < 9; ++i) R.m[i] += sin_theta * K.m[i] + one_minus_cos * KK.m[i];
  return R;
}

// gtsam::Pose3::Expmap, xi = [omega; v].
inline Pose se3_expmap(const double xi[6]) {
  const V3 w{xi[0], xi[1], xi[2]}, v{xi[3], xi[4], xi[5]};
  Pose P;
  P.R = so3_expmap(w);
  const double theta2 = dot(w, w);
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const V3 t_parallel = w * dot(w, v);
    const V3 w_cross_v = cross(w, v);
    P.t = (w_cross_v - mul(P.R, w_cross_v) + t_parallel) / theta2;
  } else {
    P.t = v;
  }
  return P;
}

// T <- T * Exp(xi)  (gtsam::Pose3::retract with the full exponential map).
inline Pose se3_retract(const Pose& T, const double xi[6]) {
  const Pose D = se3_expmap(xi);
  Pose O;
  O.R = mul(T.R, D.R);
  O.t = mul(T.R, D.t) + T.t;
  return O;
}

// Solve (H + lambda I) x = rhs for the symmetric 6x6 H (row-major) with an unpivoted LDL^T.
// Returns false when a pivot is not strictly positive.  Harness construct standing in for ISAM2
//  — not reference arithmetic.
inline bool solve6_ldlt(const double H[36], double lambda, const double rhs[6], double x[6]) {
  double L[36], D[6];
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    double d = H[6 * j + j] + lambda;
    for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k] * D[k];
    if (!(d > 0.0)) return false;
    D[j] = d;
    L[6 * j + j] = 1.0;
    for (int i = j + 1; i < 6; ++i) {
      double s = H[6 * i + j];
      for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k] * D[k];
      L[6 * i + j] = s / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = rhs[i];
    for (int k = 0; k < i; ++k) s -= L[6 * i + k] * y[k];
    y[i] = s;
  }
  for (int i = 0; i < 6; ++i) y[i] = y[i] / D[i];
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * x[k];
    x[i] = s;
  }
  return true;
}

}  // namespace mimosa_oracle
