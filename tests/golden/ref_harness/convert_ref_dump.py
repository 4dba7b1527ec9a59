"""Turns the ref_*.bin files make_golden_ref.cpp wrote (the REAL reference's results on C1) into an .npz and compares it
with the committed oracle-generated golden file: statuses, correspondence indices and per-point planes must be equal,
H / g / f within 1e-9.  Exit code 0 = the oracle's eight iVox assumptions and its ICPFactor restatement hold: parity PINNED.
usage: python tests/golden/ref_harness/convert_ref_dump.py <dir> [out.npz]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(d, out):
    g = np.load(os.path.join(os.path.dirname(HERE), "c1_golden.npz"))
    n, iters = g["in_scan"].shape[0], int(g["iters"])
    rd = lambda name, dt: np.fromfile(os.path.join(d, name), dt)
    ref = {"map_points": rd("ref_map_points_xyz.bin", "<f4").reshape(-1, 3), "knn_idx": rd("ref_knn_idx.bin", "<u8").reshape(n, 5),
           "knn_d2": rd("ref_knn_d2.bin", "<f8").reshape(n, 5), "knn_ok": rd("ref_knn_ok.bin", "u1").astype(bool),
           "tr_H": rd("ref_tr_H.bin", "<f8").reshape(iters, 6, 6), "tr_g": rd("ref_tr_g.bin", "<f8").reshape(iters, 6),
           "tr_f": rd("ref_tr_f.bin", "<f8"), "tr_pose": rd("ref_tr_pose.bin", "<f8").reshape(iters, 12),
           "status": rd("ref_status.bin", "u1").reshape(iters, n), "mean": rd("ref_mean.bin", "<f8").reshape(iters, n, 3),
           "normal": rd("ref_normal.bin", "<f8").reshape(iters, n, 3)}
    np.savez_compressed(out, **ref)
    bad = []
    if not np.array_equal(ref["map_points"], g["in_pts"]):
        bad.append("map contents / point order (iVox insert: assumptions 1-3, 6, 7)")
    rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    for it in range(iters):
        if rel(ref["tr_H"][it], g["tr_H"][it]) > 1e-9 or abs(ref["tr_f"][it] - g["tr_f"][it]) > 1e-9 * abs(g["tr_f"][it]):
            bad.append(f"H / f of iteration {it}")
    if not np.array_equal(ref["status"][-1], g["st_status"]):
        bad.append("final statuses")
    if not (np.array_equal(ref["mean"][-1], g["st_mean"]) and np.array_equal(ref["normal"][-1], g["st_normal"])):
        bad.append("final plane means / normals")
    if np.abs(ref["tr_pose"][-1][9:] - g["out_t"]).max() > 1e-8:
        bad.append("final pose")
    print("PARITY PINNED: the reference reproduces the oracle's C1 golden vectors" if not bad else "MISMATCH: " + "; ".join(bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "c1_reference.npz"))
