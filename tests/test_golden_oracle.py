"""CPU: the oracle must reproduce the committed C1 golden vectors (guards against oracle drift), and the
golden inputs must regenerate from the seed."""
import numpy as np

from helpers import load_golden
from mimosa_b200.host import HORNBILL_MAP, hornbill_config


def test_oracle_reproduces_golden(oracle):
    g = load_golden()
    m = oracle.IVoxRef(**HORNBILL_MAP)
    m.load_raw(g["in_coords"], g["in_counts"], g["in_lru"], g["in_pts_padded"], int(g["in_lru_counter"]))
    f = oracle.IcpFactorRef(m, g["in_scan"], hornbill_config())
    R, t, trace, _ = f.icp_run(g["in_R0"], g["in_t0"], int(g["iters"]), float(g["lam"]), n_threads=0)
    assert np.array_equal(R, g["out_R"]) and np.array_equal(t, g["out_t"])
    for it, tr in enumerate(trace):
        assert np.array_equal(np.array(tr.H).reshape(6, 6), g["tr_H"][it])
        assert np.array_equal(np.array(tr.g), g["tr_g"][it])
        assert tr.f == g["tr_f"][it]
        assert list(tr.counts) == g["tr_counts"][it].tolist()
        assert tr.n_searched == g["tr_n_searched"][it]
    st = f.download_state()
    for k, v in st.items():
        assert np.array_equal(v, g["st_" + k]), k


def test_golden_inputs_regenerate_from_seed(oracle):
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    m, scan, R0, t0 = mg.c1_inputs()
    g = load_golden()
    coords, counts, _, pts, _ = m.download()
    assert np.array_equal(coords, g["in_coords"]) and np.array_equal(counts, g["in_counts"])
    assert np.array_equal(pts, g["in_pts_padded"]) and np.array_equal(scan[:, :3], g["in_scan"])


def test_golden_sanity():
    g = load_golden()
    # the DA cache is live: every point searches in iteration 1, fewer afterwards
    assert g["tr_n_searched"][0] == 8192 and g["tr_n_searched"][-1] < 8192
    assert np.all(g["tr_solve_ok"] == 1)
    assert g["tr_counts"].sum(axis=1).tolist() == [8192] * int(g["iters"])
    assert abs(g["out_t"][2]) < 5e-3  # z is observable on a ground plane and converges
