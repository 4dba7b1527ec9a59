"""Generates tests/golden/c1_golden.npz: config C1 (8 192-pt scan vs ~100 k-pt ground-plane map, 5 GN
iterations, lambda = 1) — inputs plus the CPU oracle's per-iteration outputs.

The reference itself cannot be built or run in this environment (ROS / PCL / GTSAM / gtsam_points / Eigen are
absent) and ships no golden vectors, so these vectors come from the oracle
(oracle/icp_factor_ref.hpp), which is pinned by the analytic tests in tests/test_oracle_kat.py.
PARITY UNPINNED against the real reference.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import oracle_py as orc  # noqa: E402
import synth  # noqa: E402
from mimosa_b200.host import HORNBILL_MAP, hornbill_config  # noqa: E402

ITERS, LAMBDA = 5, 1.0


def c1_inputs():
    rng = synth.rng_for(1)
    m = orc.IVoxRef(**HORNBILL_MAP)
    while m.size()[1] < 100_000:
        m.insert(synth.sample_ground(20_000, 50.0, rng))
    scan = synth.plane_scan(8192, 30.0, -synth.GROUND_Z, rng)
    R0, t0 = synth.perturbed_start(np.eye(3), np.zeros(3))
    return m, scan, R0, t0


def trace_arrays(trace):
    f = lambda name, shape: np.array([np.array(getattr(tr, name)).reshape(shape) for tr in trace])
    return dict(H=f("H", (6, 6)), g=f("g", (6,)), f=np.array([tr.f for tr in trace]), delta=f("delta", (6,)),
                R=f("R", (3, 3)), t=f("t", (3,)), counts=f("counts", (9,)),
                n_searched=np.array([tr.n_searched for tr in trace]), solve_ok=np.array([tr.solve_ok for tr in trace]))


def main():
    orc.build()
    m, scan, R0, t0 = c1_inputs()
    coords, counts, lru, pts, lru_counter = m.download()
    f = orc.IcpFactorRef(m, scan, hornbill_config())
    R, t, trace, _ = f.icp_run(R0, t0, ITERS, LAMBDA, n_threads=0)
    st = f.download_state()
    out = {"in_coords": coords, "in_counts": counts, "in_lru": lru, "in_lru_counter": np.int64(lru_counter),
           "in_pts": pts[np.arange(pts.shape[1])[None, :] < counts[:, None]], "in_scan": scan[:, :3].copy(),
           "in_R0": R0, "in_t0": t0, "iters": np.int64(ITERS), "lam": np.float64(LAMBDA),
           "out_R": R, "out_t": t}
    out.update({"tr_" + k: v for k, v in trace_arrays(trace).items()})
    out.update({"st_" + k: v for k, v in st.items()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e6, "MB;", "map", m.size(), "counts", trace[-1].counts[:], "t", t)


if __name__ == "__main__":
    main()
