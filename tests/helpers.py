"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_golden.npz")


def load_golden():
    g = dict(np.load(GOLDEN))
    counts = g["in_counts"]
    cap = 20
    pts = np.zeros((counts.shape[0], cap, 3), np.float32)
    pts[np.arange(cap)[None, :] < counts[:, None]] = g["in_pts"]
    g["in_pts_padded"] = pts
    return g


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.linalg.norm(a - b)
    n = max(np.linalg.norm(b), 1e-300)
    return d / n


def g_err(g, g_ref, H_ref, f_ref):
    """Error of the linear term relative to its natural scale sqrt(trace(H) f) (see assert_linearization_close)."""
    scale = max(np.sqrt(np.trace(np.asarray(H_ref, np.float64).reshape(6, 6)) * f_ref), 1e-300)
    return np.linalg.norm(np.asarray(g, np.float64) - np.asarray(g_ref, np.float64)) / scale


def lin_fields(L):
    return {k: np.array(getattr(L, k)) for k in ("H", "g", "counts", "loc_trans_comp", "loc_rot_comp", "loc_trans_final",
                                                 "loc_rot_final", "eigvec_trans", "eigvec_rot", "degen_rot", "degen_trans",
                                                 "degen_eigvec_rot", "degen_eigvec_trans")} | {
        "f": L.f, "linearize_count": L.linearize_count, "n_searched": L.n_searched}


def assert_linearization_close(Lg, Lo, tol=1e-9, eig_tol=1e-6):
    """H, g, f differ only by summation order (tolerance `tol` relative, the contract is 1e-5);
    counts exact; eigen outputs compared up to sign and conditioning."""
    a, b = lin_fields(Lg), lin_fields(Lo)
    assert a["counts"].tolist() == b["counts"].tolist()
    assert a["n_searched"] == b["n_searched"] and a["linearize_count"] == b["linearize_count"]
    assert rel_err(a["H"], b["H"]) <= tol, rel_err(a["H"], b["H"])
    # g = -J^T e is a sum of signed terms that cancels near convergence: its rounding error scales with
    # sum |J_i| |e_i| <= sqrt(trace(H) f), so that is the scale the tolerance is relative to.
    g_scale = max(np.sqrt(np.trace(b["H"].reshape(6, 6)) * b["f"]), 1e-300)
    assert np.linalg.norm(a["g"] - b["g"]) <= tol * g_scale, np.linalg.norm(a["g"] - b["g"]) / g_scale
    assert abs(a["f"] - b["f"]) <= tol * max(abs(b["f"]), 1e-300)
    for k in ("loc_trans_final", "loc_rot_final", "loc_trans_comp", "loc_rot_comp"):
        assert np.allclose(a[k], b[k], rtol=eig_tol, atol=eig_tol * max(1.0, np.abs(b[k]).max())), (k, a[k], b[k])
    for k in ("degen_rot", "degen_trans"):
        fin = np.isfinite(b[k])
        assert np.array_equal(fin, np.isfinite(a[k]))
        assert np.allclose(a[k][fin], b[k][fin], rtol=1e-4), (k, a[k], b[k])


def assert_state_equal(sg, so, float_tol=0.0):
    """Per-point state: statuses and correspondence indices bit-exact; vectors exact by default."""
    assert np.array_equal(sg["status"], so["status"]), np.flatnonzero(sg["status"] != so["status"])[:10]
    assert np.array_equal(sg["knn_idx"], so["knn_idx"])
    for k in ("p_da", "mean", "normal", "loc_rot", "loc_trans"):
        if float_tol == 0.0:
            assert np.array_equal(sg[k], so[k]), (k, np.abs(sg[k] - so[k]).max())
        else:
            assert np.allclose(sg[k], so[k], rtol=0, atol=float_tol), (k, np.abs(sg[k] - so[k]).max())
