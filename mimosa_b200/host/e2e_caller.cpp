// The reference-facing call sequence of one scan, written the way a mimosa (C++) translation unit drives it and
// timed on the host: ICPFactor construction from a HOST scan, then `iters` x { linearize(pose) -> H, g, f on the
// host ; 6x6 solve + SE(3) retract on the host (the stand-in for ISAM2's update, graph/manager.cpp:585-588) }.
// Only the public C ABI (include/mimosa_b200.h) is used.  bench.py calls mb_e2e_scan() once per timed step for
// its `e2e` number so that the measured loop carries no Python interpreter overhead; nothing in the library
// depends on this file.
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>

#include "../../include/mimosa_b200.h"

extern "C" __attribute__((visibility("default"))) int mb_e2e_scan(
    mb_ctx* ctx, mb_map* map, const void* scan, size_t n, size_t stride_bytes, const mb_icp_config* cfg, size_t shard_begin,
    size_t shard_end, const double R0[9], const double t0[3], int iters, double lambda, double R_out[9], double t_out[3],
    double* seconds) {
  const double gravity[3] = {0.0, 0.0, -1.0};
  const auto a = std::chrono::steady_clock::now();
  mb_factor* f = nullptr;
  int rc = mb_factor_create(ctx, map, scan, n, stride_bytes, cfg, shard_begin, shard_end, &f);  // H2D: the scan
  if (rc != MB_OK) return rc;
  const auto a1 = std::chrono::steady_clock::now();
  double R[9], t[3], delta[6];
  for (int i = 0; i < 9; ++i) R[i] = R0[i];
  for (int i = 0; i < 3; ++i) t[i] = t0[i];
  mb_linearization L;
  for (int it = 0; it < iters && rc == MB_OK; ++it) {
    rc = mb_factor_linearize(f, R, t, gravity, &L);  // H2D pose, D2H normal equations
    int ok = 0;
    if (rc == MB_OK) rc = mb_gn_step(L.H, L.g, lambda, R, t, delta, &ok);
  }
  if (rc == MB_OK) rc = mb_sync(ctx);
  const auto b = std::chrono::steady_clock::now();
  *seconds = std::chrono::duration<double>(b - a).count();
  if (std::getenv("MB_E2E_VERBOSE"))
    std::fprintf(stderr, "[e2e] factor create %.1f us, %d x (linearize + gn_step) %.1f us\n",
                 1e6 * std::chrono::duration<double>(a1 - a).count(), iters, 1e6 * std::chrono::duration<double>(b - a1).count());
  for (int i = 0; i < 9; ++i) R_out[i] = R[i];
  for (int i = 0; i < 3; ++i) t_out[i] = t[i];
  const int rc2 = mb_factor_release(f);
  return rc != MB_OK ? rc : rc2;
}
