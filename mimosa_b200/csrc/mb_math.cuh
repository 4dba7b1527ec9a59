// Small fixed-size double-precision linear algebra shared by the device kernels (and compilable on the
// host for unit tests).  Operation ORDER matters here: correspondence indices, reject statuses and plane
// normals must agree bit-for-bit with the CPU restatement of the reference, so every reduction is written
// in the order the reference's Eigen expressions evaluate them (3-term fixed-size sums as a0 + (a1 + a2),
// homogeneous 4-vector squared norms as (x^2 + z^2) + (y^2 + w^2)) and this file is compiled with
// -fmad=false so no multiply-add is contracted.
//
// Reference call sites: mimosa/include/mimosa/lidar/geometric_factor.hpp:176-229 (estimatePlane),
// :406-428 (localizability / Schur complements), mimosa/include/mimosa/utils.hpp:308-313.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#define MB_UNROLL _Pragma("unroll")
#else
#define MB_HD inline
#define MB_UNROLL
#endif

namespace mb {

struct d3 {
  double x, y, z;
};
MB_HD d3 mk3(double x, double y, double z) {
  d3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
MB_HD d3 add3(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
MB_HD d3 sub3(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
MB_HD d3 scale3(d3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
MB_HD d3 div3(d3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
MB_HD double dot3(d3 a, d3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
MB_HD double sqnorm3(d3 a) { return dot3(a, a); }
MB_HD d3 cross3(d3 a, d3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// Squared distance as the voxel map evaluates it (homogeneous coordinates, w difference = 0).
MB_HD double sqdist4(double ax, double ay, double az, double bx, double by, double bz) {
  const double dx = ax - bx, dy = ay - by, dz = az - bz;
  return (dx * dx + dz * dz) + dy * dy;
}

// Row-major 3x3.
struct m33 {
  double m[9];
};
MB_HD d3 mul33v(const m33& A, d3 v) {
  return mk3(A.m[0] * v.x + (A.m[1] * v.y + A.m[2] * v.z), A.m[3] * v.x + (A.m[4] * v.y + A.m[5] * v.z),
             A.m[6] * v.x + (A.m[7] * v.y + A.m[8] * v.z));
}
MB_HD d3 mul33Tv(const m33& A, d3 v) {
  return mk3(A.m[0] * v.x + (A.m[3] * v.y + A.m[6] * v.z), A.m[1] * v.x + (A.m[4] * v.y + A.m[7] * v.z),
             A.m[2] * v.x + (A.m[5] * v.y + A.m[8] * v.z));
}
MB_HD m33 mul33(const m33& A, const m33& B) {
  m33 C;
  MB_UNROLL
  for (int r = 0; r < 3; ++r)
    MB_UNROLL
    for (int c = 0; c < 3; ++c)
      C.m[3 * r + c] = A.m[3 * r] * B.m[c] + (A.m[3 * r + 1] * B.m[3 + c] + A.m[3 * r + 2] * B.m[6 + c]);
  return C;
}
MB_HD m33 sub33(const m33& A, const m33& B) {
  m33 C;
  MB_UNROLL
  for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] - B.m[i];
  return C;
}
// Cofactor inverse scaled by 1/det (fixed-size 3x3 inverse).
MB_HD m33 inv33(const m33& A) {
  const double* a = A.m;
  const double c00 = a[4] * a[8] - a[5] * a[7];
  const double c10 = a[2] * a[7] - a[1] * a[8];
  const double c20 = a[1] * a[5] - a[2] * a[4];
  const double det = a[0] * c00 + a[3] * c10 + a[6] * c20;
  const double id = 1.0 / det;
  m33 R;
  R.m[0] = c00 * id;
  R.m[1] = c10 * id;
  R.m[2] = c20 * id;
  R.m[3] = (a[5] * a[6] - a[3] * a[8]) * id;
  R.m[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  R.m[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  R.m[6] = (a[3] * a[7] - a[4] * a[6]) * id;
  R.m[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  R.m[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return R;
}

// ---- symmetric 3x3 eigen-decomposition --------------------------------------------------------------
// Same published algorithm the reference gets from Eigen::SelfAdjointEigenSolver<Matrix3d>
// (geometric_factor.hpp:196, utils.hpp:310): scale by the largest |coefficient| of the lower triangle,
// one Householder step to tridiagonal form, implicit-shift (Wilkinson) QR sweeps with deflation at
// 2*eps, ascending selection sort.  diag/sub/Q are kept in scalars so nothing is dynamically indexed
// (registers, not local memory, on the device).
struct rot2 {
  double c, s;
};
MB_HD rot2 givens(double p, double q) {
  rot2 g;
  if (q == 0.0) {
    g.c = p < 0.0 ? -1.0 : 1.0;
    g.s = 0.0;
  } else if (p == 0.0) {
    g.c = 0.0;
    g.s = q < 0.0 ? 1.0 : -1.0;
  } else if (fabs(p) > fabs(q)) {
    const double t = q / p;
    double u = sqrt(1.0 + t * t);
    if (p < 0.0) u = -u;
    g.c = 1.0 / u;
    g.s = -t * g.c;
  } else {
    const double t = p / q;
    double u = sqrt(1.0 + t * t);
    if (q < 0.0) u = -u;
    g.s = -1.0 / u;
    g.c = -t * g.s;
  }
  return g;
}
MB_HD double scaled_hypot(double x, double y) {
  const double ax = fabs(x), ay = fabs(y);
  const double p = ax > ay ? ax : ay;
  if (p == 0.0) return 0.0;
  const double qp = (ax > ay ? ay : ax) / p;
  return p * sqrt(1.0 + qp * qp);
}
MB_HD double wilkinson_shift(double d_hi, double d_lo, double e) {
  // d_hi = diag[end-1], d_lo = diag[end], e = sub[end-1]
  const double td = (d_hi - d_lo) * 0.5;
  double mu = d_lo;
  if (td == 0.0) {
    mu -= fabs(e);
  } else {
    const double e2 = e * e;
    const double h = scaled_hypot(td, e);
    if (e2 == 0.0)
      mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
    else
      mu -= e2 / (td + (td > 0.0 ? h : -h));
  }
  return mu;
}
// One Givens similarity step on the (k, k+1) pair of a tridiagonal matrix: updates dk, dk1, sk and
// returns the rotation so the caller can apply it to Q.
MB_HD void qr_pair(rot2 g, double& dk, double& dk1, double& sk) {
  const double sdk = g.s * dk + g.c * sk;
  const double dkp1 = g.s * sk + g.c * dk1;
  dk = g.c * (g.c * dk - g.s * sk) - g.s * (g.c * sk - g.s * dk1);
  dk1 = g.s * sdk + g.c * dkp1;
  sk = g.c * sdk - g.s * dkp1;
}
MB_HD void rot_cols(rot2 g, double& a, double& b) {  // (a, b) <- (c a - s b, s a + c b)
  const double xa = a, yb = b;
  a = g.c * xa - g.s * yb;
  b = g.s * xa + g.c * yb;
}

// Returns false when the QR iteration does not converge (-> EigenSolverFail).  lam ascending; column j of
// V is the eigenvector of lam[j].  Reads the lower triangle of A only.
MB_HD bool eigh33(const m33& A, double lam[3], m33& V) {
  double a00 = A.m[0], a10 = A.m[3], a11 = A.m[4], a20 = A.m[6], a21 = A.m[7], a22 = A.m[8];
  double scale = fabs(a00);
  if (fabs(a10) > scale) scale = fabs(a10);
  if (fabs(a11) > scale) scale = fabs(a11);
  if (fabs(a20) > scale) scale = fabs(a20);
  if (fabs(a21) > scale) scale = fabs(a21);
  if (fabs(a22) > scale) scale = fabs(a22);
  if (scale == 0.0) scale = 1.0;
  a00 /= scale;
  a10 /= scale;
  a11 /= scale;
  a20 /= scale;
  a21 /= scale;
  a22 /= scale;

  double d0 = a00, d1, d2, s0, s1;
  double q00 = 1, q01 = 0, q02 = 0, q10 = 0, q11, q12, q20 = 0, q21, q22;
  const double v1n2 = a20 * a20;
  if (v1n2 <= DBL_MIN) {
    d1 = a11;
    d2 = a22;
    s0 = a10;
    s1 = a21;
    q11 = 1;
    q12 = 0;
    q21 = 0;
    q22 = 1;
  } else {
    const double beta = sqrt(a10 * a10 + v1n2);
    const double inv_beta = 1.0 / beta;
    const double m01 = a10 * inv_beta;
    const double m02 = a20 * inv_beta;
    const double q = 2.0 * m01 * a21 + m02 * (a22 - a11);
    d1 = a11 + m02 * q;
    d2 = a22 - m02 * q;
    s0 = beta;
    s1 = a21 - m01 * q;
    q11 = m01;
    q12 = m02;
    q21 = m02;
    q22 = -m01;
  }

  const double prec = 2.0 * DBL_EPSILON;
  int end = 2, iter = 0;
  const int max_iter = 90;  // 30 * n
  while (end > 0) {
    // deflation tests on the active block [start, end]; for n = 3 testing both sub-diagonals is exact
    // because an already-zero entry stays zero.
    if (fabs(s0) <= (fabs(d0) + fabs(d1)) * prec || fabs(s0) <= DBL_MIN) s0 = 0.0;
    if (end == 2 && (fabs(s1) <= (fabs(d1) + fabs(d2)) * prec || fabs(s1) <= DBL_MIN)) s1 = 0.0;
    while (end > 0 && (end == 2 ? s1 : s0) == 0.0) --end;
    if (end <= 0) break;
    ++iter;
    if (iter > max_iter) break;
    int start = end - 1;
    if (start > 0 && s0 != 0.0) start = 0;
    if (end == 2 && start == 0) {
      // full 3x3 sweep: rotations on (0,1) then (1,2) with bulge chasing
      const double mu = wilkinson_shift(d1, d2, s1);
      double x = d0 - mu, z = s0;
      rot2 g = givens(x, z);
      qr_pair(g, d0, d1, s0);
      x = s0;
      z = -g.s * s1;
      s1 = g.c * s1;
      rot_cols(g, q00, q01);
      rot_cols(g, q10, q11);
      rot_cols(g, q20, q21);
      g = givens(x, z);
      qr_pair(g, d1, d2, s1);
      s0 = g.c * s0 - g.s * z;
      rot_cols(g, q01, q02);
      rot_cols(g, q11, q12);
      rot_cols(g, q21, q22);
    } else if (end == 2) {
      // 2x2 block (1,2)
      const double mu = wilkinson_shift(d1, d2, s1);
      const rot2 g = givens(d1 - mu, s1);
      qr_pair(g, d1, d2, s1);
      rot_cols(g, q01, q02);
      rot_cols(g, q11, q12);
      rot_cols(g, q21, q22);
    } else {
      // 2x2 block (0,1)
      const double mu = wilkinson_shift(d0, d1, s0);
      const rot2 g = givens(d0 - mu, s0);
      qr_pair(g, d0, d1, s0);
      rot_cols(g, q00, q01);
      rot_cols(g, q10, q11);
      rot_cols(g, q20, q21);
    }
  }
  const bool ok = iter <= max_iter;
  if (ok) {
    // selection sort, ascending: position 0 takes the minimum of (d0,d1,d2), first index on ties;
    // then position 1 the minimum of the remaining two.
    int k = 0;
    double best = d0;
    if (d1 < best) {
      best = d1;
      k = 1;
    }
    if (d2 < best) k = 2;
    if (k == 1) {
      double t = d0; d0 = d1; d1 = t;
      t = q00; q00 = q01; q01 = t;
      t = q10; q10 = q11; q11 = t;
      t = q20; q20 = q21; q21 = t;
    } else if (k == 2) {
      double t = d0; d0 = d2; d2 = t;
      t = q00; q00 = q02; q02 = t;
      t = q10; q10 = q12; q12 = t;
      t = q20; q20 = q22; q22 = t;
    }
    if (d2 < d1) {
      double t = d1; d1 = d2; d2 = t;
      t = q01; q01 = q02; q02 = t;
      t = q11; q11 = q12; q12 = t;
      t = q21; q21 = q22; q22 = t;
    }
  }
  lam[0] = d0 * scale;
  lam[1] = d1 * scale;
  lam[2] = d2 * scale;
  V.m[0] = q00; V.m[1] = q01; V.m[2] = q02;
  V.m[3] = q10; V.m[4] = q11; V.m[5] = q12;
  V.m[6] = q20; V.m[7] = q21; V.m[8] = q22;
  return ok;
}

// ---- closed-form symmetric 3x3 eigen-decomposition ----------------------------------------------------
// Used only where bit-parity with the reference's iterative solver is NOT required (the 6x6-level
// localizability / degeneracy outputs, whose input already differs from the CPU's by summation order): the
// iterative QR above is ~1600 dependent fp64 instructions for one thread, this is ~300.  Trigonometric roots of
// the characteristic polynomial of the shifted, scaled matrix; eigenvectors from cross products of rows of
// (A - lambda I) (the published "direct" 3x3 method, cf. Eigen's computeDirect); eigenvalues are then
// re-evaluated as Rayleigh quotients, which squares their error.  lam ascending, column j of V for lam[j].
MB_HD d3 kernel_vec(const m33& M, d3& rep_row) {
  // M is singular (rank <= 2): returns a unit vector of its null space; rep_row = the row with the largest
  // diagonal magnitude, normalised (used to build the next eigenvector).
  const double a0 = fabs(M.m[0]), a1 = fabs(M.m[4]), a2 = fabs(M.m[8]);
  const int i0 = a0 >= a1 ? (a0 >= a2 ? 0 : 2) : (a1 >= a2 ? 1 : 2);
  const d3 r0 = mk3(M.m[0], M.m[1], M.m[2]), r1 = mk3(M.m[3], M.m[4], M.m[5]), r2 = mk3(M.m[6], M.m[7], M.m[8]);
  const d3 rep = i0 == 0 ? r0 : (i0 == 1 ? r1 : r2);
  const d3 oa = i0 == 0 ? r1 : r0, ob = i0 == 2 ? r1 : r2;
  const d3 c0 = cross3(rep, oa), c1 = cross3(rep, ob);
  const double n0 = sqnorm3(c0), n1 = sqnorm3(c1);
  const d3 c = n0 > n1 ? c0 : c1;
  const double n = n0 > n1 ? n0 : n1;
  rep_row = div3(rep, sqrt(sqnorm3(rep)));
  return div3(c, sqrt(n));
}
MB_HD bool eigh33_direct_raw(const m33& A, double lam[3], m33& V);
// Closed-form solve with an a-posteriori check: the off-diagonal residue b_ij = v_i^T A v_j bounds the error of
// each Rayleigh quotient by sum_j b_ij^2 / |lam_i - lam_j|; when that exceeds 1e-10 relative (near-repeated
// eigenvalues far below the spectral radius — where the trigonometric roots lose digits) the iterative solver
// is used instead, so the result is always as good as eigh33's to ~1e-10.
MB_HD void eigh33_direct(const m33& A, double lam[3], m33& V) {
  if (!eigh33_direct_raw(A, lam, V)) eigh33(A, lam, V);
}
MB_HD bool eigh33_direct_raw(const m33& A, double lam[3], m33& V) {
  const double shift = (A.m[0] + A.m[4] + A.m[8]) / 3.0;
  double m00 = A.m[0] - shift, m11 = A.m[4] - shift, m22 = A.m[8] - shift;
  double m10 = A.m[3], m20 = A.m[6], m21 = A.m[7];
  double scale = fabs(m00);
  scale = fmax(scale, fabs(m11));
  scale = fmax(scale, fabs(m22));
  scale = fmax(scale, fabs(m10));
  scale = fmax(scale, fabs(m20));
  scale = fmax(scale, fabs(m21));
  if (!(scale > 0.0)) {  // multiple of the identity (or NaN input, which then propagates through lam)
    lam[0] = lam[1] = lam[2] = shift + scale * 0.0;
    for (int i = 0; i < 9; ++i) V.m[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return scale == 0.0;  // NaN input -> let the iterative solver produce the reference's answer
  }
  const double inv = 1.0 / scale;
  m00 *= inv; m11 *= inv; m22 *= inv; m10 *= inv; m20 *= inv; m21 *= inv;
  // roots of x^3 - c2 x^2 + c1 x - c0 with c2 = trace = 0 after the shift
  const double c0 = m00 * m11 * m22 + 2.0 * m10 * m20 * m21 - m00 * m21 * m21 - m11 * m20 * m20 - m22 * m10 * m10;
  const double c1 = m00 * m11 - m10 * m10 + m00 * m22 - m20 * m20 + m11 * m22 - m21 * m21;
  const double c2 = m00 + m11 + m22;
  const double c2_3 = c2 / 3.0;
  double a_3 = (c2 * c2_3 - c1) / 3.0;
  if (a_3 < 0.0) a_3 = 0.0;
  const double half_b = 0.5 * (c0 + c2_3 * (2.0 * c2_3 * c2_3 - c1));
  double q = a_3 * a_3 * a_3 - half_b * half_b;
  if (q < 0.0) q = 0.0;
  const double rho = sqrt(a_3);
  const double theta = atan2(sqrt(q), half_b) / 3.0;
  const double ct = cos(theta), st = sin(theta);
  const double s3 = 1.7320508075688772;
  double r0 = c2_3 - rho * (ct + s3 * st), r1 = c2_3 - rho * (ct - s3 * st), r2 = c2_3 + 2.0 * rho * ct;
  d3 v0, v1, v2;
  if (r2 - r0 <= DBL_EPSILON) {
    v0 = mk3(1, 0, 0);
    v1 = mk3(0, 1, 0);
    v2 = mk3(0, 0, 1);
  } else {
    m33 M;
    M.m[1] = M.m[3] = m10;
    M.m[2] = M.m[6] = m20;
    M.m[5] = M.m[7] = m21;
    // start with the eigenvalue that is best separated from the other two
    const double d_hi = r2 - r1, d_lo = r1 - r0;
    const bool top = d_hi > d_lo;  // true: k = 2 (largest) first, then l = 0; false: k = 0 first, then l = 2
    const double rk = top ? r2 : r0, rl = top ? r0 : r2;
    const double d_small = top ? d_lo : d_hi;  // separation of the remaining pair
    M.m[0] = m00 - rk;
    M.m[4] = m11 - rk;
    M.m[8] = m22 - rk;
    d3 rep;
    const d3 vk = kernel_vec(M, rep);
    d3 vl;
    if (d_small <= 2.0 * DBL_EPSILON * (top ? d_hi : d_lo)) {
      // the other two eigenvalues coincide: any orthonormal completion will do
      const d3 tmp = sub3(rep, scale3(vk, dot3(vk, rep)));
      vl = div3(tmp, sqrt(sqnorm3(tmp)));
    } else {
      M.m[0] = m00 - rl;
      M.m[4] = m11 - rl;
      M.m[8] = m22 - rl;
      d3 rep2;
      vl = kernel_vec(M, rep2);
      // re-orthogonalise against vk
      const d3 tmp = sub3(vl, scale3(vk, dot3(vk, vl)));
      vl = div3(tmp, sqrt(sqnorm3(tmp)));
    }
    const d3 vm = cross3(top ? vk : vl, top ? vl : vk);  // middle = v2 x v0 (sign is free)
    v0 = top ? vl : vk;
    v2 = top ? vk : vl;
    v1 = vm;
  }
  // Rayleigh quotients on the ORIGINAL matrix
  const m33 As = A;
  auto rq = [&](d3 v) { return dot3(v, mul33v(As, v)); };
  // A may only have its lower triangle meaningful for callers of eigh33; here callers pass full symmetric blocks
  lam[0] = rq(v0);
  lam[1] = rq(v1);
  lam[2] = rq(v2);
  V.m[0] = v0.x; V.m[3] = v0.y; V.m[6] = v0.z;
  V.m[1] = v1.x; V.m[4] = v1.y; V.m[7] = v1.z;
  V.m[2] = v2.x; V.m[5] = v2.y; V.m[8] = v2.z;
  (void)r1;
  const d3 Av1 = mul33v(As, v1), Av2 = mul33v(As, v2);
  const double b01 = dot3(v0, Av1), b02 = dot3(v0, Av2), b12 = dot3(v1, Av2);
  const double g01 = fabs(lam[1] - lam[0]), g02 = fabs(lam[2] - lam[0]), g12 = fabs(lam[2] - lam[1]);
  const double rad = fmax(fabs(lam[0]), fabs(lam[2]));
  const double floor_abs = 1e-15 * rad;
  const double e0 = b01 * b01 / g01 + b02 * b02 / g02;
  const double e1 = b01 * b01 / g01 + b12 * b12 / g12;
  const double e2 = b02 * b02 / g02 + b12 * b12 / g12;
  // (0/0 = NaN and x/0 = inf both fail the comparisons below, as they should)
  return e0 <= 1e-10 * fabs(lam[0]) + floor_abs && e1 <= 1e-10 * fabs(lam[1]) + floor_abs &&
         e2 <= 1e-10 * fabs(lam[2]) + floor_abs;
}

// ---- SE(3) exponential (gtsam::Pose3::Expmap with the full exponential map, xi = [omega; v]) -------
MB_HD m33 so3_exp(d3 w) {
  const double th2 = dot3(w, w);
  m33 W;
  W.m[0] = 0; W.m[1] = -w.z; W.m[2] = w.y;
  W.m[3] = w.z; W.m[4] = 0; W.m[5] = -w.x;
  W.m[6] = -w.y; W.m[7] = w.x; W.m[8] = 0;
  m33 R;
  MB_UNROLL
  for (int i = 0; i < 9; ++i) R.m[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (th2 <= DBL_EPSILON) {
    MB_UNROLL
    for (int i = 0; i < 9; ++i) R.m[i] += W.m[i];
    return R;
  }
  const double th = sqrt(th2);
  const double sn = sin(th);
  const double s2 = sin(th / 2.0);
  const double omc = 2.0 * s2 * s2;
  m33 K;
  MB_UNROLL
  for (int i = 0; i < 9; ++i) K.m[i] = W.m[i] / th;
  const m33 KK = mul33(K, K);
  MB_UNROLL
  for (int i = 0; i < 9; ++i) R.m[i] += sn * K.m[i] + omc * KK.m[i];
  return R;
}
MB_HD void se3_retract(m33& R, d3& t, const double xi[6]) {
  const d3 w = mk3(xi[0], xi[1], xi[2]), v = mk3(xi[3], xi[4], xi[5]);
  const m33 dR = so3_exp(w);
  const double th2 = dot3(w, w);
  d3 dt;
  if (th2 > DBL_EPSILON) {
    const d3 tpar = scale3(w, dot3(w, v));
    const d3 wxv = cross3(w, v);
    dt = div3(add3(sub3(wxv, mul33v(dR, wxv)), tpar), th2);
  } else {
    dt = v;
  }
  t = add3(mul33v(R, dt), t);
  R = mul33(R, dR);
}

// (H + lambda I) x = rhs, unpivoted LDL^T; false when a pivot is not strictly positive.
MB_HD bool solve6_ldlt(const double H[36], double lambda, const double rhs[6], double x[6]) {
  // Every loop has constant bounds and is fully unrolled on the device, so L, D, y live in registers.
  double L[36], D[6], y[6];
  bool ok = true;
  MB_UNROLL
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  MB_UNROLL
  for (int j = 0; j < 6; ++j) {
    double d = H[6 * j + j] + lambda;
    MB_UNROLL
    for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k] * D[k];
    ok = ok && (d > 0.0);
    D[j] = d;
    L[6 * j + j] = 1.0;
    MB_UNROLL
    for (int i = j + 1; i < 6; ++i) {
      double s = H[6 * i + j];
      MB_UNROLL
      for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k] * D[k];
      L[6 * i + j] = s / d;
    }
  }
  MB_UNROLL
  for (int i = 0; i < 6; ++i) {
    double s = rhs[i];
    MB_UNROLL
    for (int k = 0; k < i; ++k) s -= L[6 * i + k] * y[k];
    y[i] = s;
  }
  MB_UNROLL
  for (int i = 0; i < 6; ++i) y[i] = y[i] / D[i];
  MB_UNROLL
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    MB_UNROLL
    for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * x[k];
    x[i] = s;
  }
  return ok;
}

// int(x) - (x < int(x)), mimosa/include/mimosa/lidar/utils.hpp:218-222.
MB_HD int fast_floor(double x) {
  const int i = (int)x;
  return i - (x < (double)i ? 1 : 0);
}

}  // namespace mb
