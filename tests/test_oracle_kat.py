"""Analytic known-answer tests that pin the CPU oracle.  The reference ships no tests,
golden vectors or fixtures for this path (parity unpinned), so the oracle is pinned by closed-form cases."""
import numpy as np
import pytest

import synth
from mimosa_b200.host import RegistrationConfig, hornbill_config


def test_fast_floor(oracle):
    L = oracle.lib()
    for x, want in [(0.0, 0), (0.5, 0), (-0.5, -1), (-1.0, -1), (1.0, 1), (-1e-12, -1), (2.999999, 2), (-3.000001, -4)]:
        assert L.orc_fast_floor(x) == want


def test_eigh3_matches_numpy(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        A = rng.normal(size=(3, 3))
        A = A @ A.T * 10 ** rng.uniform(-6, 3)
        ok, lam, V = oracle.eigh3(A)
        assert ok
        w = np.linalg.eigvalsh(A)
        assert np.allclose(lam, w, rtol=1e-10, atol=1e-14 * max(1, abs(w).max()))
        assert np.all(np.diff(lam) >= 0)
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-12)
        assert np.allclose(A @ V, V * lam, atol=1e-9 * max(1e-300, abs(w).max()))


def test_eigh3_diagonal_and_degenerate(oracle):
    ok, lam, V = oracle.eigh3(np.diag([3.0, 1.0, 2.0]))
    assert ok and np.allclose(lam, [1, 2, 3])
    assert np.allclose(np.abs(V), [[0, 0, 1], [1, 0, 0], [0, 1, 0]])
    ok, lam, V = oracle.eigh3(np.zeros((3, 3)))
    assert ok and np.all(lam == 0) and np.allclose(V, np.eye(3))


def test_se3_expmap_vs_matrix_exponential(oracle):
    from scipy.linalg import expm

    rng = np.random.default_rng(2)
    for scale in (1e-9, 1e-3, 0.3, 2.5):
        xi = rng.normal(size=6) * scale
        R, t = oracle.se3_expmap(xi)
        w, v = xi[:3], xi[3:]
        X = np.zeros((4, 4))
        X[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        X[:3, 3] = v
        E = expm(X)
        assert np.allclose(R, E[:3, :3], atol=1e-12)
        assert np.allclose(t, E[:3, 3], atol=1e-12)


def test_solve6(oracle):
    rng = np.random.default_rng(3)
    A = rng.normal(size=(6, 6))
    H = A @ A.T + np.eye(6)
    b = rng.normal(size=6)
    ok, x = oracle.solve6(H, 0.5, b)
    assert ok and np.allclose((H + 0.5 * np.eye(6)) @ x, b, atol=1e-10)
    ok, _ = oracle.solve6(np.zeros((6, 6)), 0.0, b)
    assert not ok


# ---- voxel map semantics ------------------------------------------------------------------------------
def test_insert_cap_and_min_dist(oracle):
    m = oracle.IVoxRef(1.0, 0.2, 20, 19, 1000)
    # (v) 25 well separated points in one voxel: only the first 20 are kept, even though the rest are far
    g = np.stack(np.meshgrid(np.arange(5), np.arange(5)), -1).reshape(-1, 2) * 0.19 + 0.02
    pts = np.concatenate([g, np.full((25, 1), 0.5)], 1).astype(np.float32)
    # 0.19 spacing < 0.2 -> neighbours along a row are too close: use 0.21 spacing instead, 4 x 0.21 < 1
    pts[:, :2] = (np.stack(np.meshgrid(np.arange(5), np.arange(5)), -1).reshape(-1, 2) * 0.21 + 0.02).astype(np.float32)
    m.insert(pts)
    nv, npts, lru = m.size()
    assert (nv, npts, lru) == (1, 20, 1)
    coords, counts, _, stored, _ = m.download()
    assert np.array_equal(stored[0, :20], pts[:20])
    # min-dist: strict '<' on the squared distance, first come wins
    m2 = oracle.IVoxRef(1.0, 0.2, 20, 19, 1000)
    m2.insert(np.array([[0.5, 0.5, 0.5], [0.5, 0.5, 0.6], [0.5, 0.5, 0.75]], np.float32))
    assert m2.size()[1] == 2
    assert np.allclose(m2.download()[3][0, 1], [0.5, 0.5, 0.75])


def test_insert_order_dependence(oracle):
    # (iv) the same points in two orders give different stored sets, each obeying the sequential rule
    a = np.array([[0.10, 0.5, 0.5], [0.25, 0.5, 0.5], [0.40, 0.5, 0.5]], np.float32)
    m1 = oracle.IVoxRef(1.0, 0.2, 20, 19, 1000)
    m1.insert(a)
    m2 = oracle.IVoxRef(1.0, 0.2, 20, 19, 1000)
    m2.insert(a[[1, 0, 2]])
    s1 = m1.download()[3][0, : m1.download()[1][0]]
    s2 = m2.download()[3][0, : m2.download()[1][0]]
    assert np.allclose(s1, a[[0, 2]]) and np.allclose(s2, a[[1]])


def test_voxel_ids_are_creation_order_and_negative_coords(oracle):
    m = oracle.IVoxRef(1.0, 0.0, 20, 19, 1000)
    m.insert(np.array([[2.5, 0.5, 0.5], [-0.5, 0.5, 0.5], [2.6, 0.5, 0.5], [-1.0, -0.0, 0.5]], np.float32))
    coords, counts, _, _, _ = m.download()
    assert coords.tolist() == [[2, 0, 0], [-1, 0, 0]]
    assert counts.tolist() == [2, 2]
    idx, d2, ok = m.knn_search(np.array([[2.52, 0.5, 0.5]]), 2)
    assert ok[0] and idx[0].tolist() == [(0 << 32) | 0, (0 << 32) | 1]


def test_knn_neighbourhood_modes(oracle):
    # (ii) a point in a (+1,+1,+1) corner voxel nearer than everything else: mode 19 must not see it
    pts = np.array([[1.05, 1.05, 1.05], [0.1, 0.1, 0.1], [0.2, 0.8, 0.3]], np.float32)
    q = np.array([[0.95, 0.95, 0.95]])
    m19 = oracle.IVoxRef(1.0, 0.0, 20, 19, 1000)
    m19.insert(pts)
    m27 = oracle.IVoxRef(1.0, 0.0, 20, 27, 1000)
    m27.insert(pts)
    i19, d19, ok19 = m19.knn_search(q, 2)
    i27, d27, ok27 = m27.knn_search(q, 2)
    assert ok19[0] and ok27[0]
    assert (i27[0, 0] >> np.uint64(32)) == 0 and np.isclose(d27[0, 0], 3 * 0.1**2, atol=1e-6)
    assert set((i19[0] >> np.uint64(32)).tolist()) == {1}
    # mode 7 sees only face neighbours; mode 1 only the centre voxel
    m7 = oracle.IVoxRef(1.0, 0.0, 20, 7, 1000)
    m7.insert(np.array([[0.5, 0.5, 0.5], [1.5, 0.5, 0.5], [1.5, 1.5, 0.5]], np.float32))
    idx, _, ok = m7.knn_search(np.array([[0.9, 0.9, 0.5]]), 3)
    assert not ok[0]  # the (1,1,0) edge voxel is invisible in mode 7 -> only 2 found
    idx, _, ok = m7.knn_search(np.array([[0.9, 0.9, 0.5]]), 2)
    assert ok[0]


def test_knn_ties_keep_visiting_order(oracle):
    # (iii) equal distances: the earlier visited voxel wins.  Mode 7 order: centre, +x, -x, +y, -y, +z, -z.
    m = oracle.IVoxRef(1.0, 0.0, 20, 7, 1000)
    m.insert(np.array([[-0.25, 0.5, 0.5], [1.25, 0.5, 0.5]], np.float32))  # voxel ids 0 (-x), 1 (+x)
    idx, d2, ok = m.knn_search(np.array([[0.5, 0.5, 0.5]]), 1)
    assert d2[0, 0] == 0.75**2 and (idx[0, 0] >> np.uint64(32)) == 1  # +x is visited before -x
    # mode 19 order is nested i,j,k ascending: (-1,0,0) comes before (1,0,0)
    m = oracle.IVoxRef(1.0, 0.0, 20, 19, 1000)
    m.insert(np.array([[1.25, 0.5, 0.5], [-0.25, 0.5, 0.5]], np.float32))  # ids 0 (+x), 1 (-x)
    idx, d2, ok = m.knn_search(np.array([[0.5, 0.5, 0.5]]), 1)
    assert (idx[0, 0] >> np.uint64(32)) == 1
    # within a voxel: stored order
    m = oracle.IVoxRef(1.0, 0.0, 20, 1, 1000)
    m.insert(np.array([[0.25, 0.5, 0.5], [0.75, 0.5, 0.5]], np.float32))
    idx, d2, ok = m.knn_search(np.array([[0.5, 0.5, 0.5]]), 2)
    assert idx[0].tolist() == [0, 1] and d2[0, 0] == d2[0, 1]


def test_knn_vs_bruteforce(oracle):
    # (vii) restricted k-NN equals a brute-force search over exactly the neighbourhood voxels
    rng = np.random.default_rng(5)
    pts = rng.uniform(-3, 3, (4000, 3)).astype(np.float32)
    m = oracle.IVoxRef(1.0, 0.1, 20, 19, 1000)
    m.insert(pts)
    coords, counts, _, stored, _ = m.download()
    q = rng.uniform(-3.5, 3.5, (300, 3))
    idx, d2, ok = m.knn_search(q, 5)
    offs = [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) if not (i and j and k)]
    lut = {tuple(c): v for v, c in enumerate(coords.tolist())}
    for qi in range(q.shape[0]):
        c = np.floor(q[qi]).astype(int)
        cand = []
        for o in offs:
            v = lut.get((c[0] + o[0], c[1] + o[1], c[2] + o[2]))
            if v is None:
                continue
            for j in range(counts[v]):
                d = stored[v, j].astype(np.float64) - q[qi]
                cand.append(((d[0] * d[0] + d[2] * d[2]) + d[1] * d[1], (v << 32) | j))
        cand.sort(key=lambda x: x[0])
        assert ok[qi] == (len(cand) >= 5)
        n = min(5, len(cand))
        assert d2[qi, :n].tolist() == [c_[0] for c_ in cand[:n]]
        assert idx[qi, :n].tolist() == [c_[1] for c_ in cand[:n]]


def test_lru_eviction_compacts_ids(oracle):
    m = oracle.IVoxRef(1.0, 0.0, 20, 19, 3)  # horizon 3, clear cycle 10
    m.insert(np.array([[0.5, 0.5, 0.5]], np.float32))  # voxel A, lru 0
    for i in range(8):
        m.insert(np.array([[10.5 + i, 0.5, 0.5]], np.float32))  # lru 1..8
    assert m.size()[0] == 9
    m.insert(np.array([[30.5, 0.5, 0.5]], np.float32))  # 10th insert: lru 9; counter -> 10; evict lru+3 < 10
    coords, counts, lru, _, counter = m.download()
    assert counter == 10
    # removed iff lru + horizon < counter: lru 7 stays (7 + 3 < 10 is false)
    assert coords[:, 0].tolist() == [16, 17, 30] and lru.tolist() == [7, 8, 9]
    # ids were compacted: the survivor that was voxel 8 is now voxel 1
    idx, _, ok = m.knn_search(np.array([[17.5, 0.5, 0.5]]), 1)
    assert ok[0] and idx[0, 0] == (1 << 32)


def test_snapshot_is_deep(oracle):
    m = oracle.IVoxRef(1.0, 0.0, 20, 19, 1000)
    m.insert(np.array([[0.5, 0.5, 0.5]], np.float32))
    s = m.snapshot()
    m.insert(np.array([[0.6, 0.5, 0.5], [5.5, 0.5, 0.5]], np.float32))
    assert s.size()[:2] == (1, 1) and m.size()[:2] == (2, 3)


def test_downsample_order(oracle):
    pts = np.array([[0.1, 0.1, 0.1], [5.1, 0.1, 0.1], [0.15, 0.1, 0.1], [0.9, 0.9, 0.9], [5.5, 0.5, 0.5]], np.float32)
    keep = oracle.downsample(pts, 1.0, 20, 0.2)
    assert keep.tolist() == [0, 3, 1, 4]  # voxel creation order, then in-voxel order; point 2 is too close to 0
    keep = oracle.downsample(pts, 1.0, 1, 0.2)
    assert keep.tolist() == [0, 1]


# ---- ICP factor -------------------------------------------------------------------------------------------
def _plane_case(oracle, n_map=60000, n_scan=2000, seed=7, cfg=None):
    rng = np.random.default_rng(seed)
    cfg = cfg or hornbill_config()
    m = oracle.IVoxRef(1.0, 0.2, 20, 19, 1000)
    m.insert(synth.sample_ground(n_map, 20.0, rng))
    scan = synth.plane_scan(n_scan, 15.0, -synth.GROUND_Z, rng)
    return m, scan, cfg


def test_linearize_plane_closed_form(oracle):
    # (i) noisy plane z = c, scan offset by known dz, roll, pitch: e ~ n.(m - p), J = [(n_s x p)^T, -n_s^T] / sigma
    m, scan, cfg = _plane_case(oracle)
    cfg.use_huber = False
    f = oracle.IcpFactorRef(m, scan, cfg)
    dz = 0.03
    R = synth.rot_from_rpy(0.004, -0.003, 0.0)
    t = np.array([0.0, 0.0, dz])
    L = f.linearize(R, t)
    st = f.download_state()
    valid = st["status"] == 8
    assert valid.sum() > 0.5 * scan.shape[0]  # the Line gate (l2 > 3 l1) rejects ~30% of 5-point patches
    assert np.array(L.counts).sum() == scan.shape[0] and L.counts[8] == valid.sum() and L.n_searched == scan.shape[0]
    ps = scan[:, :3].astype(np.float64)
    pt = ps @ R.T + t
    n, mu = st["normal"], st["mean"]
    assert np.all(n[valid, 2] > 0.99)  # oriented towards the sensor (above the plane)
    sigma = float(np.float32(cfg.lidar_point_noise_std_dev))
    e = np.einsum("ij,ij->i", n, mu - pt) / sigma
    ns = n @ R  # R^T n, row-wise
    J = np.concatenate([np.cross(ns, ps), -ns], 1) / sigma
    H = (J[valid, :, None] * J[valid, None, :]).sum(0)
    b = (J[valid] * e[valid, None]).sum(0)
    Ho = np.array(L.H).reshape(6, 6)
    assert np.allclose(Ho, H, rtol=1e-9, atol=1e-6)
    assert np.allclose(Ho, Ho.T)
    assert np.allclose(np.array(L.g), -b, rtol=1e-9, atol=1e-6)
    assert np.isclose(L.f, (e[valid] ** 2).sum(), rtol=1e-9)
    # residual sign/magnitude: the plane is dz below the transformed scan on average
    assert abs(np.median(e[valid]) * sigma + dz) < 0.01
    # a single plane observes z, roll, pitch only: H has (near) zero rows for x, y, yaw
    assert Ho[3, 3] < 1e-2 * Ho[5, 5] and Ho[4, 4] < 1e-2 * Ho[5, 5] and Ho[2, 2] < 1e-2 * Ho[0, 0]
    # localizability outputs
    lam_t = np.linalg.eigvalsh(Ho[3:, 3:])
    assert np.allclose(np.array(L.loc_trans_final) ** 2, lam_t, rtol=1e-6, atol=1e-6)
    assert L.loc_trans_comp[2] > 0.9 * valid.sum()  # every valid normal projects onto the z eigenvector
    assert L.linearize_count == 1


def test_da_cache_and_status_semantics(oracle):
    m, scan, cfg = _plane_case(oracle, seed=8)
    f = oracle.IcpFactorRef(m, scan, cfg)
    R, t = np.eye(3), np.zeros(3)
    L1 = f.linearize(R, t)
    s1 = f.download_state()
    assert L1.n_searched == scan.shape[0]
    # move by less than min_dist/4 = 0.05: nobody re-associates, planes are reused, residuals change
    L2 = f.linearize(R, t + np.array([0, 0, 0.02]))
    s2 = f.download_state()
    assert L2.n_searched == 0 and L2.linearize_count == 2
    assert np.array_equal(s1["p_da"], s2["p_da"]) and np.array_equal(s1["mean"], s2["mean"])
    assert np.array_equal(s1["knn_idx"], s2["knn_idx"])
    assert not np.isclose(L1.f, L2.f)
    # statuses <= CorresPlaneInvalid are sticky while cached
    rejected = s1["status"] <= 6
    assert np.array_equal(s1["status"][rejected], s2["status"][rejected])
    # move by more than the gate: everybody re-associates
    L3 = f.linearize(R, t + np.array([0.2, 0, 0.0]))
    assert L3.n_searched == scan.shape[0]


def test_gates(oracle):
    cfg = hornbill_config()
    # (vi) noise-free plane -> MinEigenValueLow
    g = np.stack(np.meshgrid(np.arange(-20, 20) * 0.25, np.arange(-20, 20) * 0.25), -1).reshape(-1, 2)
    flat = np.concatenate([g, np.zeros((g.shape[0], 1))], 1).astype(np.float32)
    m = oracle.IVoxRef(1.0, 0.2, 20, 19, 1000)
    m.insert(flat)
    scan = np.array([[0.3, 0.2, 0.4], [1.1, -2.2, 0.3]], np.float32)
    f = oracle.IcpFactorRef(m, scan, cfg)
    f.linearize(np.eye(3), np.zeros(3))
    assert f.download_state()["status"].tolist() == [4, 4]
    # too few neighbours -> InsufficientCorresPoints; far neighbours -> CorresMaxDist
    m = oracle.IVoxRef(1.0, 0.2, 20, 19, 1000)
    m.insert(np.array([[0.1, 0.1, 0.1], [0.5, 0.5, 0.5], [0.9, 0.1, 0.5], [5.5, 5.5, 5.5], [5.2, 5.5, 5.5],
                       [5.8, 5.5, 5.5], [5.5, 5.2, 5.5], [5.5, 5.8, 5.5], [5.5, 5.5, 6.5]], np.float32))
    scan = np.array([[0.5, 0.5, 0.4], [5.5, 5.5, 5.6]], np.float32)
    f = oracle.IcpFactorRef(m, scan, cfg)
    f.linearize(np.eye(3), np.zeros(3))
    st = f.download_state()
    assert st["status"][0] == 1
    assert st["status"][1] in (5, 6, 4, 2)  # 5 found within reach; which gate fires depends on geometry
    # a line of points -> Line
    line = np.stack([np.arange(0, 40) * 0.21, np.zeros(40), np.zeros(40)], 1)
    line[:, 1:] += np.random.default_rng(0).normal(0, 0.004, (40, 2))
    m = oracle.IVoxRef(1.0, 0.2, 20, 19, 1000)
    m.insert(line.astype(np.float32))
    f = oracle.IcpFactorRef(m, np.array([[4.0, 0.05, 0.02]], np.float32), cfg)
    f.linearize(np.eye(3), np.zeros(3))
    assert f.download_state()["status"][0] == 5


def test_huber_and_max_error(oracle):
    m, scan, cfg = _plane_case(oracle, seed=9)
    f = oracle.IcpFactorRef(m, scan, cfg)
    # 0.5 m above the plane: |e| = 0.5; s = 1 - 0.9*0.5/sqrt(range) < 0.9 whenever sqrt(range) < 4.5 -> MaxError
    L = f.linearize(np.eye(3), np.array([0, 0, 0.5]))
    st = f.download_state()["status"]
    rng_ = np.linalg.norm(scan[:, :3].astype(np.float64), axis=1)
    searched_ok = st >= 7
    assert np.all(st[searched_ok & (np.sqrt(rng_) < 4.4)] == 7)
    # Huber: 0.15 m offset -> |e/sigma| ~ 2.1 > 1.345 -> weight sqrt(kh/|e/sigma|)
    f2 = oracle.IcpFactorRef(m, scan, cfg)
    Lh = f2.linearize(np.eye(3), np.array([0, 0, 0.15]))
    cfg2 = hornbill_config()
    cfg2.use_huber = False
    f3 = oracle.IcpFactorRef(m, scan, cfg2)
    Ln = f3.linearize(np.eye(3), np.array([0, 0, 0.15]))
    assert Lh.f < Ln.f and Lh.counts[8] == Ln.counts[8]


def test_reg_4_dof_projection(oracle):
    m, scan, cfg = _plane_case(oracle, seed=10)
    cfg.reg_4_dof = True
    f = oracle.IcpFactorRef(m, scan, cfg)
    R = synth.rot_from_rpy(0.01, 0.02, 0.3)
    L = f.linearize(R, np.array([0, 0, 0.02]), gravity_unit=(0, 0, -1.0))
    H = np.array(L.H).reshape(6, 6)
    z = R.T @ np.array([0, 0, 1.0])
    # the rotational block only acts along local z: vectors orthogonal to z are in its null space
    u = np.cross(z, [1, 0, 0])
    assert np.allclose(H[:3, :3] @ u, 0, atol=1e-6 * abs(H).max())
    assert np.allclose(H[:3, 3:].T @ u, 0, atol=1e-6 * abs(H).max())
    assert np.allclose(np.cross(np.array(L.g)[:3], z), 0, atol=1e-6 * abs(np.array(L.g)).max())


def test_project_on_degeneracy_mirrors_reference_bug(oracle):
    # geometric_factor.hpp:477-557 re-sums arrays that are never written: G = 0, g = 0 when triggered
    m, scan, cfg = _plane_case(oracle, seed=11)
    cfg.project_on_degneneracy = True
    cfg.degen_thresh_trans = 40.0
    f = oracle.IcpFactorRef(m, scan, cfg)
    L = f.linearize(np.eye(3), np.zeros(3))
    assert not np.any(np.array(L.H)) and not np.any(np.array(L.g)) and L.f > 0


def test_icp_converges_on_plane(oracle):
    m, scan, cfg = _plane_case(oracle, seed=12, n_scan=4000)
    f = oracle.IcpFactorRef(m, scan, cfg)
    R0, t0 = synth.perturbed_start(np.eye(3), np.zeros(3), (0.010, -0.008, 0.0, 0.0, 0.0, 0.03))
    R, t, trace, _ = f.icp_run(R0, t0, 8, lam=1.0)
    assert all(tr.solve_ok for tr in trace)
    assert abs(t[2]) < 0.005 and abs(R[2, 0]) < 1e-3 and abs(R[2, 1]) < 1e-3
    assert trace[-1].f < trace[0].f


def test_parallel_equals_serial(oracle):
    m, scan, cfg = _plane_case(oracle, seed=13)
    fa = oracle.IcpFactorRef(m, scan, cfg)
    fb = oracle.IcpFactorRef(m, scan, cfg)
    La = fa.linearize(np.eye(3), np.array([0, 0, 0.01]), n_threads=0)
    Lb = fb.linearize(np.eye(3), np.array([0, 0, 0.01]), n_threads=4)
    assert bytes(La) == bytes(Lb)
