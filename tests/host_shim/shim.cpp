// Test-only host build of the device math header (mimosa_b200/csrc/mb_math.cuh compiles as plain C++).
// Lets the CPU test-suite check, bit for bit, that the arithmetic the kernels run (eigen-decomposition,
// SE(3) retract, LDL^T, fast_floor, distance order) equals the oracle's.  Not a product path.
#include "../../mimosa_b200/csrc/mb_math.cuh"

extern "C" {
int shim_eigh33(const double* A, double* lam, double* V) {
  mb::m33 a, v;
  for (int i = 0; i < 9; ++i) a.m[i] = A[i];
  const bool ok = mb::eigh33(a, lam, v);
  for (int i = 0; i < 9; ++i) V[i] = v.m[i];
  return ok ? 1 : 0;
}
int shim_eigh33_direct(const double* A, double* lam, double* V) {
  mb::m33 a, v;
  for (int i = 0; i < 9; ++i) a.m[i] = A[i];
  const bool fast = mb::eigh33_direct_raw(a, lam, v);
  if (!fast) mb::eigh33(a, lam, v);
  for (int i = 0; i < 9; ++i) V[i] = v.m[i];
  return fast ? 1 : 0;
}
void shim_se3_retract(double* R, double* t, const double* xi) {
  mb::m33 r;
  for (int i = 0; i < 9; ++i) r.m[i] = R[i];
  mb::d3 tt = mb::mk3(t[0], t[1], t[2]);
  mb::se3_retract(r, tt, xi);
  for (int i = 0; i < 9; ++i) R[i] = r.m[i];
  t[0] = tt.x;
  t[1] = tt.y;
  t[2] = tt.z;
}
int shim_solve6(const double* H, double lambda, const double* rhs, double* x) {
  return mb::solve6_ldlt(H, lambda, rhs, x) ? 1 : 0;
}
int shim_fast_floor(double x) { return mb::fast_floor(x); }
void shim_inv33(const double* A, double* out) {
  mb::m33 a;
  for (int i = 0; i < 9; ++i) a.m[i] = A[i];
  const mb::m33 r = mb::inv33(a);
  for (int i = 0; i < 9; ++i) out[i] = r.m[i];
}
}
