"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libmimosa_b200.so), against the CPU
oracle on identical seeded inputs.  Bar: map contents, correspondence indices, squared distances, statuses
and per-point state bit-exact; H / g / f within 1e-9 relative (contract: 1e-5; they differ only by the order
of the fp64 summation); poses within 1e-8."""
import numpy as np
import pytest

import synth
from helpers import assert_linearization_close, assert_state_equal, g_err, load_golden, rel_err
from mimosa_b200 import HORNBILL_MAP, ICPFactor, IncrementalVoxelMap, Scan, gn_step, hornbill_config
from mimosa_b200.capi import MB_ERR_INVALID_ARG, MB_ERR_UNSUPPORTED, MimosaError

pytestmark = pytest.mark.gpu

H_TOL = 1e-9  # relative; north_star contract is 1e-5
POSE_TOL = 1e-8


def both_maps(ctx, oracle, **kw):
    kw = {**HORNBILL_MAP, **kw}
    return IncrementalVoxelMap(ctx, **kw), oracle.IVoxRef(**kw)


def assert_maps_equal(mg, mo):
    cg, ng, lg, pg, kg = mg.download()
    co, no, lo, po, ko = mo.download()
    assert mg.size() == mo.size()
    assert np.array_equal(cg, co) and np.array_equal(ng, no) and np.array_equal(lg, lo) and kg == ko
    assert np.array_equal(pg, po)


# ---- map ----------------------------------------------------------------------------------------------
def test_insert_matches_oracle_multi_batch(ctx, oracle):
    rng = np.random.default_rng(100)
    mg, mo = both_maps(ctx, oracle)
    for b in range(6):
        pts = synth.sample_world(30000, 60.0, rng)
        if b == 2:  # dense cluster: thousands of candidates for a handful of voxels (long runs)
            pts = (rng.uniform(-1.5, 1.5, (30000, 3)) + [3.0, -2.0, 0.5]).astype(np.float32)
        if b == 4:  # exact duplicates of earlier points and negative coordinates on voxel faces
            pts = np.concatenate([pts[:1000], pts[:1000], np.array([[-1.0, -2.0, -3.0], [-1.0, -2.0, -3.0], [0.0, 0.0, 0.0]], np.float32)])
        mg.insert(pts)
        mo.insert(pts)
        assert_maps_equal(mg, mo)
    mg.release()


def test_insert_stride_and_empty(ctx, oracle):
    rng = np.random.default_rng(101)
    mg, mo = both_maps(ctx, oracle)
    rec = np.zeros((5000, 8), np.float32)  # lidar::Point rows, stride 32
    rec[:, :3] = synth.sample_ground(5000, 10.0, rng)
    rec[:, 3:] = 7.0
    mg.insert(rec)
    mo.insert(rec)
    mg.insert(np.zeros((0, 3), np.float32))  # an empty insert still advances the LRU clock
    mo.insert(np.zeros((0, 3), np.float32))
    assert_maps_equal(mg, mo)
    assert np.array_equal(mg.get_cloud(), mo.download()[3][np.arange(20)[None, :] < mo.download()[1][:, None]])
    mg.release()


def test_lru_eviction_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(102)
    mg, mo = both_maps(ctx, oracle, lru_horizon=7)
    for b in range(45):
        centre = np.array([b * 3.0, 0.0, 0.0])
        pts = (rng.uniform(-4, 4, (800, 3)) * [1, 1, 0.1] + centre).astype(np.float32)
        mg.insert(pts)
        mo.insert(pts)
        if b % 5 == 4:
            assert_maps_equal(mg, mo)
    assert mo.size()[0] < 45 * 30  # something was evicted
    q = rng.uniform(-4, 4, (500, 3)) * [1, 1, 0.1] + [44 * 3.0, 0, 0]
    ig, dg, og = mg.knn_search(q, 5)
    io, do, oo = mo.knn_search(q, 5)
    assert np.array_equal(ig, io) and np.array_equal(dg, do) and np.array_equal(og, oo)
    mg.release()


def test_snapshot_is_deep_and_guarded(ctx, oracle):
    rng = np.random.default_rng(103)
    mg, mo = both_maps(ctx, oracle)
    a = synth.sample_ground(20000, 15.0, rng)
    mg.insert(a)
    mo.insert(a)
    sg, so = mg.snapshot(), mo.snapshot()
    b = synth.sample_ground(20000, 25.0, rng)
    mg.insert(b)
    mo.insert(b)
    assert_maps_equal(sg, so)
    assert_maps_equal(mg, mo)
    # a map referenced by a factor is immutable: insert must fail, snapshot + insert must work
    f = ICPFactor(ctx, mg, a[:100], hornbill_config())
    with pytest.raises(MimosaError) as e:
        mg.insert(b)
    assert e.value.code == MB_ERR_INVALID_ARG
    s2 = mg.snapshot()
    s2.insert(b)
    f.release()
    for m in (sg, s2, mg):
        m.release()


def test_upload_download_roundtrip(ctx, oracle):
    rng = np.random.default_rng(104)
    mg, mo = both_maps(ctx, oracle)
    mo.insert(synth.sample_world(50000, 40.0, rng))
    co, no, lo, po, ko = mo.download()
    mg.upload(co, no, lo, po, ko)
    assert_maps_equal(mg, mo)
    more = synth.sample_world(20000, 40.0, rng)
    mg.insert(more)
    mo.insert(more)
    assert_maps_equal(mg, mo)
    mg.release()


@pytest.mark.parametrize("mode", [1, 7, 19, 27])
def test_knn_matches_oracle(ctx, oracle, mode):
    rng = np.random.default_rng(110 + mode)
    mg, mo = both_maps(ctx, oracle, nbr_mode=mode, min_dist=0.05)  # dense buckets: many voxels at cap 20
    pts = rng.uniform(-4, 4, (60000, 3)).astype(np.float32)
    mg.insert(pts)
    mo.insert(pts)
    q = rng.uniform(-5, 5, (4000, 3))
    q[:50] = pts[:50].astype(np.float64)  # exact hits: d2 == 0
    for k in (1, 3, 5, 8):
        ig, dg, og = mg.knn_search(q, k)
        io, do, oo = mo.knn_search(q, k)
        assert np.array_equal(og, oo)
        assert np.array_equal(ig, io)
        assert np.array_equal(dg, do)
    got = mg.points(ig[og].ravel())
    cloud_idx = ig[og].ravel()
    co, no, lo, po, ko = mo.download()
    assert np.array_equal(got, po[(cloud_idx >> np.uint64(32)).astype(np.int64), (cloud_idx & np.uint64(0xFFFFFFFF)).astype(np.int64)].astype(np.float64))
    mg.release()


def test_knn_ties_and_corners(ctx, oracle):
    # lattice points: huge numbers of exactly equal squared distances across and inside voxels
    g = np.stack(np.meshgrid(*[np.arange(-8, 8) * 0.25 + 0.125] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    for mode in (7, 19, 27):
        mg, mo = both_maps(ctx, oracle, nbr_mode=mode, min_dist=0.0, cap=20)
        mg.insert(g)
        mo.insert(g)
        q = np.stack(np.meshgrid(*[np.arange(-6, 6) * 0.25] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
        for k in (1, 5, 8):
            ig, dg, og = mg.knn_search(q, k)
            io, do, oo = mo.knn_search(q, k)
            assert np.array_equal(ig, io) and np.array_equal(dg, do) and np.array_equal(og, oo)
        mg.release()


def test_knn_sparse_and_empty(ctx, oracle):
    mg, mo = both_maps(ctx, oracle)
    q = np.array([[0.5, 0.5, 0.5], [100.0, 100.0, 100.0]])
    ig, dg, og = mg.knn_search(q, 5)  # empty map
    io, do, oo = mo.knn_search(q, 5)
    assert np.array_equal(ig, io) and np.array_equal(dg, do) and not og.any()
    pts = np.array([[0.4, 0.5, 0.5], [0.9, 0.5, 0.5], [1.2, 0.5, 0.5]], np.float32)
    mg.insert(pts)
    mo.insert(pts)
    ig, dg, og = mg.knn_search(q, 5)  # fewer than k: partial lists padded with ~0 / DBL_MAX
    io, do, oo = mo.knn_search(q, 5)
    assert np.array_equal(ig, io) and np.array_equal(dg, do) and np.array_equal(og, oo)
    mg.release()


# ---- factor -------------------------------------------------------------------------------------------
def _golden_pair(ctx, oracle, cfg=None):
    g = load_golden()
    cfg = cfg or hornbill_config()
    mg, mo = both_maps(ctx, oracle)
    mo.load_raw(g["in_coords"], g["in_counts"], g["in_lru"], g["in_pts_padded"], int(g["in_lru_counter"]))
    mg.upload(g["in_coords"], g["in_counts"], g["in_lru"], g["in_pts_padded"], int(g["in_lru_counter"]))
    return g, mg, mo, ICPFactor(ctx, mg, g["in_scan"], cfg), oracle.IcpFactorRef(mo, g["in_scan"], cfg)


def test_c1_linearize_sequence_matches_oracle(ctx, oracle):
    """Config C1, driven call by call through mb_factor_linearize with the oracle's poses (so both sides
    see identical inputs every iteration): state exact, normal equations to 1e-9."""
    g, mg, mo, fg, fo = _golden_pair(ctx, oracle)
    R, t = g["in_R0"], g["in_t0"]
    for it in range(int(g["iters"])):
        Lg = fg.linearize(R, t)
        Lo = fo.linearize(R, t)
        assert_linearization_close(Lg, Lo, H_TOL)
        assert_state_equal(fg.download_state(), fo.download_state())
        R, t = g["tr_R"][it], g["tr_t"][it]
    assert fg.get_linearize_count() == int(g["iters"])
    fg.release()
    mg.release()


def test_c1_icp_run_matches_golden(ctx):
    """Device-resident GN loop against the committed golden trace (no oracle involved at run time)."""
    g = load_golden()
    mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
    mg.upload(g["in_coords"], g["in_counts"], g["in_lru"], g["in_pts_padded"], int(g["in_lru_counter"]))
    for graph in (False, True):
        fg = ICPFactor(ctx, mg, g["in_scan"], hornbill_config())
        fg.set_flags(cuda_graph=graph)
        R, t, trace = fg.icp_run(g["in_R0"], g["in_t0"], int(g["iters"]), float(g["lam"]))
        for it, tr in enumerate(trace):
            assert list(tr.counts) == g["tr_counts"][it].tolist(), it
            assert tr.n_searched == g["tr_n_searched"][it] and tr.solve_ok == 1
            assert rel_err(np.array(tr.H).reshape(6, 6), g["tr_H"][it]) <= 1e-7
            assert g_err(tr.g, g["tr_g"][it], g["tr_H"][it], g["tr_f"][it]) <= 1e-7
            assert abs(tr.f - g["tr_f"][it]) <= 1e-7 * g["tr_f"][it]
            assert np.abs(np.array(tr.R).reshape(3, 3) - g["tr_R"][it]).max() <= POSE_TOL
            assert np.abs(np.array(tr.t) - g["tr_t"][it]).max() <= POSE_TOL
        assert np.abs(R - g["out_R"]).max() <= POSE_TOL and np.abs(t - g["out_t"]).max() <= POSE_TOL
        st = fg.download_state()
        assert np.array_equal(st["status"], g["st_status"])
        assert np.array_equal(st["knn_idx"], g["st_knn_idx"])
        assert np.abs(st["mean"] - g["st_mean"]).max() == 0.0
        assert np.abs(st["normal"] - g["st_normal"]).max() == 0.0
        # run it again from a reset factor: bitwise repeatable (fixed reduction order)
        fg.reset()
        R2, t2, trace2 = fg.icp_run(g["in_R0"], g["in_t0"], int(g["iters"]), float(g["lam"]))
        assert np.array_equal(R2, R) and np.array_equal(t2, t)
        assert bytes(trace2[-1]) == bytes(trace[-1])
        fg.release()
    mg.release()


def _world_case(ctx, oracle, n_map, n_scan, seed, half=60.0, pattern="os0", **mapkw):
    rng = np.random.default_rng(seed)
    mg, mo = both_maps(ctx, oracle, **mapkw)
    for _ in range(max(1, n_map // 100000)):
        pts = synth.sample_world(min(n_map, 100000), half, rng)
        mg.insert(pts)
        mo.insert(pts)
    R_true = synth.rot_from_rpy(0.01, -0.02, 0.4)
    t_true = np.array([3.0, -2.0, 0.3])
    scan = synth.make_scan(R_true, t_true, n_scan, rng, pattern=pattern, max_range=half * 0.9)
    R0, t0 = synth.perturbed_start(R_true, t_true)
    return mg, mo, scan, R0, t0, R_true, t_true


def test_world_icp_matches_oracle(ctx, oracle):
    """Ground + boxes, ray-cast OS0-128-style scan (stride-32 lidar::Point rows), 10 GN iterations, lambda 0:
    every iteration's trace and the final per-point state against the oracle."""
    mg, mo, scan, R0, t0, R_true, t_true = _world_case(ctx, oracle, 600000, 30000, 120)
    assert_maps_equal(mg, mo)
    cfg = hornbill_config()
    fg, fo = ICPFactor(ctx, mg, scan, cfg), oracle.IcpFactorRef(mo, scan, cfg)
    Rg, tg, trg = fg.icp_run(R0, t0, 10, 0.0)
    Ro, to, tro, _ = fo.icp_run(R0, t0, 10, 0.0, n_threads=4)
    for it, (a, b) in enumerate(zip(trg, tro)):
        assert list(a.counts) == list(b.counts), (it, list(a.counts), list(b.counts))
        assert a.n_searched == b.n_searched and a.solve_ok == b.solve_ok == 1
        # component localizabilities: iterations before the last come out of the pass folded into the next
        # k_linearize, the last one out of k_loc_comp
        assert np.allclose(np.array(a.loc_trans_comp), np.array(b.loc_trans_comp), rtol=1e-4)  # same tolerance as helpers
        assert np.allclose(np.array(a.loc_rot_comp), np.array(b.loc_rot_comp), rtol=1e-4)
        assert rel_err(a.H, b.H) <= 1e-7 and g_err(a.g, b.g, b.H, b.f) <= 1e-7
        assert np.abs(np.array(a.R) - np.array(b.R)).max() <= POSE_TOL
        assert np.abs(np.array(a.t) - np.array(b.t)).max() <= POSE_TOL
    assert_state_equal(fg.download_state(), fo.download_state(), float_tol=1e-9)
    # and ICP actually registers the scan
    assert np.abs(tg - t_true).max() < 0.02 and np.abs(Rg - R_true).max() < 2e-3
    assert trg[1].n_searched > 0 and trg[-1].n_searched < trg[0].n_searched
    fg.release()
    mg.release()


@pytest.mark.parametrize("variant", ["enwide", "k8_mode27", "reg4dof_nohuber", "mode7_k3"])
def test_config_variants(ctx, oracle, variant):
    cfg = hornbill_config()
    mapkw = {}
    if variant == "enwide":  # mimosa/config/enwide/params.yaml:87-90
        cfg.target_ivox_map_leaf_size = cfg.source_voxel_grid_filter_leaf_size = 0.5
        cfg.target_ivox_map_min_dist_in_voxel = cfg.source_voxel_grid_min_dist_in_voxel = 0.15
        mapkw = dict(leaf=0.5, min_dist=0.15)
    elif variant == "k8_mode27":
        cfg.num_corres_points = 8
        mapkw = dict(nbr_mode=27)
    elif variant == "reg4dof_nohuber":
        cfg.reg_4_dof, cfg.use_huber = True, False
    elif variant == "mode7_k3":
        cfg.num_corres_points = 3
        mapkw = dict(nbr_mode=7)
    mg, mo, scan, R0, t0, _, _ = _world_case(ctx, oracle, 200000, 6000, 130, half=40.0, **mapkw)
    fg, fo = ICPFactor(ctx, mg, scan, cfg), oracle.IcpFactorRef(mo, scan, cfg)
    g_unit = np.array([0.02, -0.01, -1.0])
    g_unit /= np.linalg.norm(g_unit)
    R, t = R0, t0
    for it in range(3):
        Lg, Lo = fg.linearize(R, t, g_unit), fo.linearize(R, t, g_unit)
        assert_linearization_close(Lg, Lo, H_TOL)
        assert_state_equal(fg.download_state(), fo.download_state())
        ok, d = oracle.solve6(np.array(Lo.H), 1e-3, np.array(Lo.g))
        R, t = oracle.se3_retract(R, t, d if ok else np.zeros(6))
    fg.release()
    mg.release()


def test_forced_search_flag_and_reset(ctx, oracle):
    mg, mo, scan, R0, t0, _, _ = _world_case(ctx, oracle, 200000, 5000, 140, half=40.0)
    cfg = hornbill_config()
    fg, fo = ICPFactor(ctx, mg, scan, cfg), oracle.IcpFactorRef(mo, scan, cfg)
    L1 = fg.linearize(R0, t0)
    L2 = fg.linearize(R0, t0)  # same pose: fully cached
    assert L1.n_searched == scan.shape[0] and L2.n_searched == 0
    assert rel_err(L2.H, L1.H) <= 1e-14 and list(L1.counts) == list(L2.counts)
    fg.set_flags(forced_search=True)
    L3 = fg.linearize(R0, t0)
    assert L3.n_searched == scan.shape[0] and rel_err(L3.H, L1.H) <= 1e-14
    fg.set_flags()
    fg.reset()
    fo.linearize(R0, t0)
    L4 = fg.linearize(R0, t0)
    assert L4.linearize_count == 1 and bytes(L4.H) == bytes(L1.H)
    assert_state_equal(fg.download_state(), fo.download_state())
    fg.release()
    mg.release()


def test_ragged_and_tiny_scans(ctx, oracle):
    mg, mo, scan, R0, t0, _, _ = _world_case(ctx, oracle, 100000, 1000, 150, half=30.0)
    cfg = hornbill_config()
    for n in (1, 31, 32, 33, 257, 999):
        fg, fo = ICPFactor(ctx, mg, scan[:n], cfg), oracle.IcpFactorRef(mo, scan[:n], cfg)
        assert_linearization_close(fg.linearize(R0, t0), fo.linearize(R0, t0), H_TOL)
        assert_state_equal(fg.download_state(), fo.download_state())
        fg.release()
    # scan entirely outside the map: everything InsufficientCorresPoints, H = 0
    far = scan[:64].copy()
    far[:, :3] += 500.0
    fg = ICPFactor(ctx, mg, far, cfg)
    L = fg.linearize(np.eye(3), np.zeros(3))
    assert L.counts[1] == 64 and not np.any(np.array(L.H)) and L.f == 0.0
    fg.release()
    mg.release()


def test_page_locked_scan_buffers_take_the_dma_path(ctx, oracle):
    """A scan handed over from page-locked memory (mb_host_register) is fetched by DMA and unpacked on the device;
    results must be identical to the pageable path (CPU staging), for the factor and for mb_scan_upload, and a
    sharded factor must pick its own block of the registered buffer."""
    mg, mo, scan, R0, t0, _, _ = _world_case(ctx, oracle, 100000, 3000, 151, half=30.0)
    cfg = hornbill_config()
    pageable = np.ascontiguousarray(scan, dtype=np.float32)
    locked = pageable.copy()
    ctx.host_register(locked)
    try:
        for shard in (None, (100, 2077)):
            fa, fb = ICPFactor(ctx, mg, pageable, cfg, shard), ICPFactor(ctx, mg, locked, cfg, shard)
            locked_backup = locked.copy()
            La, Lb = fa.linearize(R0, t0), fb.linearize(R0, t0)
            assert np.array_equal(np.array(La.H), np.array(Lb.H)) and list(La.counts) == list(Lb.counts)
            sa, sb = fa.download_state(), fb.download_state()
            for k in sa:
                assert np.array_equal(sa[k], sb[k]), k
            assert np.array_equal(locked, locked_backup)
            fa.release()
            fb.release()
        fo = oracle.IcpFactorRef(mo, scan, cfg)
        fg = ICPFactor(ctx, mg, locked, cfg)
        assert_linearization_close(fg.linearize(R0, t0), fo.linearize(R0, t0), H_TOL)
        fg.release()
        sa, sb = Scan(ctx, pageable), Scan(ctx, locked)
        assert np.array_equal(sa.download(), sb.download())
        sa.release()
        sb.release()
    finally:
        ctx.host_unregister(locked)
    mg.release()


def test_unsupported_and_invalid(ctx):
    mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
    cfg = hornbill_config()
    cfg.project_on_degneneracy = True
    with pytest.raises(MimosaError) as e:
        ICPFactor(ctx, mg, np.zeros((4, 3), np.float32), cfg)
    assert e.value.code == MB_ERR_UNSUPPORTED
    cfg = hornbill_config()
    cfg.num_corres_points = 9
    with pytest.raises(MimosaError) as e:
        ICPFactor(ctx, mg, np.zeros((4, 3), np.float32), cfg)
    assert e.value.code == MB_ERR_UNSUPPORTED
    with pytest.raises(MimosaError):
        IncrementalVoxelMap(ctx, 1.0, 0.2, 20, 13, 10)  # neighbor_voxel_mode must be 1/7/19/27
    with pytest.raises(MimosaError):
        mg.insert(np.array([[np.nan, 0, 0]], np.float32))
    mg.release()


def test_downsample_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(160)
    R = synth.rot_from_rpy(0.0, 0.0, 0.2)
    scan = synth.make_scan(R, np.array([1.0, 2.0, 0.0]), 40000, rng, max_range=60.0)
    for leaf, cap, md in ((1.0, 20, 0.2), (0.5, 20, 0.15), (2.0, 3, 0.0)):
        got = ctx.downsample(scan, leaf, cap, md)
        want = oracle.downsample(scan, leaf, cap, md)
        assert np.array_equal(got, want)
    assert ctx.downsample(np.zeros((0, 3), np.float32), 1.0, 20, 0.2).size == 0


# ---- full-size properties (BASELINE.json sizes; the oracle is only sampled) ---------------------------------
def test_full_size_properties(ctx, oracle):
    """131 072-pt scan vs a ~2 M-pt map (config C2 sizes): size-independent properties —
    insert idempotence, sorted distances, k-NN == oracle on a random sample, cached == forced linearisation,
    graph == stream launch, and the ICP converges to the true pose."""
    rng = synth.rng_for(2)
    mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
    fed = synth.build_map(mg.insert, 2_000_000, 250.0, rng, chunk=1 << 20, size_fn=lambda: mg.size()[1])
    nv, npts, _ = mg.size()
    assert npts >= 2_000_000
    snap = mg.snapshot()
    snap.insert(fed[0])  # re-inserting already-inserted points changes nothing (min-dist / cap reject them)
    assert snap.size()[:2] == (nv, npts)
    snap.release()
    R_true, t_true = synth.rot_from_rpy(0.0, 0.0, 0.3), np.array([1.0, -1.0, 0.2])
    scan = synth.make_scan(R_true, t_true, 131072, rng)
    assert scan.shape[0] == 131072
    R0, t0 = synth.perturbed_start(R_true, t_true)
    q = scan[:, :3].astype(np.float64) @ R0.T + t0
    idx, d2, ok = mg.knn_search(q, 5)
    assert np.all(np.diff(d2[ok], axis=1) >= 0)
    pts = mg.points(idx[ok].ravel()).reshape(-1, 5, 3)
    d = pts - q[ok][:, None, :]
    assert np.array_equal((d[..., 0] ** 2 + d[..., 2] ** 2) + d[..., 1] ** 2, d2[ok])
    # oracle on the same map (loaded from the device dump) for a random sample of queries
    mo = oracle.IVoxRef(**HORNBILL_MAP)
    mo.load_raw(*mg.download())
    sel = rng.choice(q.shape[0], 4000, replace=False)
    io, do, oo = mo.knn_search(q[sel], 5, n_threads=4)
    assert np.array_equal(idx[sel], io) and np.array_equal(d2[sel], do) and np.array_equal(ok[sel], oo)
    cfg = hornbill_config()
    fg = ICPFactor(ctx, mg, scan, cfg)
    R, t, tr = fg.icp_run(R0, t0, 20, 0.0)
    assert np.abs(t - t_true).max() < 0.02 and np.abs(R - R_true).max() < 2e-3
    fg.reset()
    fg.set_flags(cuda_graph=True)
    R2, t2, tr2 = fg.icp_run(R0, t0, 20, 0.0)
    assert np.array_equal(R, R2) and np.array_equal(t, t2) and bytes(tr[-1]) == bytes(tr2[-1])
    fg.reset()
    fg.set_flags(forced_search=True)
    R3, t3, tr3 = fg.icp_run(R0, t0, 20, 0.0)
    assert all(x.n_searched == 131072 for x in tr3)
    # forced search re-associates every iteration, so the trajectories differ slightly but agree at convergence
    assert np.abs(t3 - t).max() < 5e-3
    # one full oracle linearisation at full size (a few seconds of CPU)
    fo = oracle.IcpFactorRef(mo, scan, cfg)
    fg.reset()
    fg.set_flags()
    assert_linearization_close(fg.linearize(R0, t0), fo.linearize(R0, t0, n_threads=4), H_TOL)
    assert_state_equal(fg.download_state(), fo.download_state())
    fg.release()
    mg.release()


def test_c4_size_parity(ctx, oracle):
    """BASELINE.json's headline configuration at FULL size (131 072-point scan, ~10 M-point / 1.27 M-voxel map, bench.py's
    inputs): restricted k-NN bit-exact against the oracle on 4 000 sampled queries, then two full linearisations (the second
    after a Gauss-Newton step: data-association cache in play) with statuses, correspondence indices and per-point planes
    bit-exact and H, g, f within 1e-9."""
    import bench

    rng, scan, R0, t0, R_true, t_true = bench.make_inputs()
    mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
    synth.build_map(mg.insert, 10_000_000, 500.0, rng, size_fn=lambda: mg.size()[1])
    nv, npts, _ = mg.size()
    assert npts >= 10_000_000 and nv > 1_000_000
    mo = oracle.IVoxRef(**HORNBILL_MAP)
    mo.load_raw(*mg.download())
    q = scan[:, :3].astype(np.float64) @ R0.T + t0
    sel = synth.rng_for(77).choice(q.shape[0], 4000, replace=False)
    ig, dg, og = mg.knn_search(q[sel], 5)
    io, do, oo = mo.knn_search(q[sel], 5, n_threads=8)
    assert np.array_equal(og, oo) and np.array_equal(ig, io) and np.array_equal(dg, do) and oo.mean() > 0.9
    cfg = hornbill_config()
    fg, fo = ICPFactor(ctx, mg, scan, cfg), oracle.IcpFactorRef(mo, scan, cfg)
    Lg, Lo = fg.linearize(R0, t0), fo.linearize(R0, t0, n_threads=8)
    assert_linearization_close(Lg, Lo, H_TOL)
    assert_state_equal(fg.download_state(), fo.download_state())
    R1, t1, _, ok = gn_step(Lg, R0, t0, 0.0)
    assert ok
    Lg, Lo = fg.linearize(R1, t1), fo.linearize(R1, t1, n_threads=8)
    assert_linearization_close(Lg, Lo, H_TOL)
    assert_state_equal(fg.download_state(), fo.download_state())
    assert Lg.counts[8] > 60000
    fg.release()
    mg.release()


def test_sharded_factors_sum_to_full(ctx, oracle):
    """Scan-block sharding (the multi-GPU decomposition) on one GPU: per-shard normal equations add up to the
    unsharded ones, per-point state of each shard equals the matching slice."""
    from mimosa_b200 import shard_range

    mg, mo, scan, R0, t0, _, _ = _world_case(ctx, oracle, 200000, 9001, 170, half=40.0)
    cfg = hornbill_config()
    full = ICPFactor(ctx, mg, scan, cfg)
    Lf = full.linearize(R0, t0)
    sf = full.download_state()
    H = np.zeros(36)
    g = np.zeros(6)
    f = 0.0
    counts = np.zeros(9, dtype=np.int64)
    for world in (3,):
        for r in range(world):
            b, e = shard_range(scan.shape[0], r, world)
            part = ICPFactor(ctx, mg, scan, cfg, shard=(b, e))
            Lp = part.linearize(R0, t0)
            sp = part.download_state()
            for k in ("status", "knn_idx", "mean", "normal", "p_da"):
                assert np.array_equal(sp[k], sf[k][b:e]), k
            H += np.array(Lp.H)
            g += np.array(Lp.g)
            f += Lp.f
            counts += np.array(Lp.counts)
            part.release()
    assert rel_err(H, Lf.H) <= 1e-12 and rel_err(g, Lf.g) <= 1e-9 and abs(f - Lf.f) <= 1e-12 * Lf.f
    assert counts.tolist() == list(Lf.counts)
    full.release()
    mg.release()


@pytest.mark.parametrize("pattern,n_scans,n_pts", [("os0", 12, 8000), ("airy", 3, 65280)])
def test_streaming_pipeline_matches_oracle(ctx, oracle, pattern, n_scans, n_pts):
    """Configs C3 / C5 in miniature, device-resident end to end: upload -> deskew (per-timestamp float poses,
    lidar/manager.cpp:494-509) -> T_B_L (geometric.cpp:153-160) -> voxel downsample (geometric.cpp:55-126) ->
    ICP factor + GN iterations -> snapshot + float world transform + insert (geometric.cpp:483-495), scan after
    scan with LRU eviction live, against the oracle fed the identical inputs."""
    from mimosa_b200 import Scan

    rng = np.random.default_rng(180)
    mg, mo = both_maps(ctx, oracle, lru_horizon=6)
    for _ in range(2):
        pts = synth.sample_world(120000, 60.0, rng)
        mg.insert(pts)
        mo.insert(pts)
    cfg = hornbill_config()
    R_B_L = synth.rot_from_rpy(0.01, -0.02, 0.03).astype(np.float32)
    t_B_L = np.array([0.1, -0.05, 0.2], np.float32)
    T_BL = np.concatenate([R_B_L.reshape(9), t_B_L])[None, :]
    R_BL64, t_BL64 = R_B_L.astype(np.float64), t_B_L.astype(np.float64)
    for s in range(n_scans):
        R_true = synth.rot_from_rpy(0.0, 0.0, 0.05 * s)
        t_true = np.array([0.6 * s, 0.2 * s, 0.3])
        # sensor pose = body pose * T_B_L
        R_s, t_s = R_true @ R_BL64, R_true @ t_BL64 + t_true
        rec = synth.make_scan(R_s, t_s, n_pts, rng, pattern=pattern, max_range=50.0)
        n = rec.shape[0]
        n_poses = 40
        pose_index = ((np.arange(n, dtype=np.int64) * n_poses) // n).astype(np.uint32)
        poses = np.zeros((n_poses, 12), np.float32)
        for p in range(n_poses):
            a = 1e-3 * (n_poses - 1 - p)
            poses[p, :9] = synth.rot_from_rpy(0.2 * a, -0.1 * a, a).reshape(9)
            poses[p, 9:] = [0.5 * a, -0.3 * a, 0.1 * a]
        # --- GPU
        sc = Scan(ctx, rec)
        sc.deskew(pose_index, poses)
        sc.transform(R_B_L, t_B_L)
        ds = sc.downsample(cfg.source_voxel_grid_filter_leaf_size, 20, cfg.source_voxel_grid_min_dist_in_voxel)
        # --- oracle
        rec_o = rec.copy()
        oracle.transform_f32(rec_o, poses, pose_index)
        oracle.transform_f32(rec_o, T_BL)
        keep = oracle.downsample(rec_o, cfg.source_voxel_grid_filter_leaf_size, 20, cfg.source_voxel_grid_min_dist_in_voxel)
        ds_o = np.ascontiguousarray(rec_o[keep])
        assert np.array_equal(sc.download().view(np.uint32), rec_o.view(np.uint32))
        assert np.array_equal(ds.download().view(np.uint32), ds_o.view(np.uint32))
        # --- factor + GN
        R0, t0 = synth.perturbed_start(R_true, t_true, (0.004, -0.003, 0.005, 0.03, -0.02, 0.02))
        fg, fo = ICPFactor(ctx, mg, ds, cfg), oracle.IcpFactorRef(mo, ds_o, cfg)
        Rg, tg, trg = fg.icp_run(R0, t0, 6, 1e-6)
        Ro, to, tro, _ = fo.icp_run(R0, t0, 6, 1e-6, n_threads=4)
        for a, b in zip(trg, tro):
            assert list(a.counts) == list(b.counts) and a.n_searched == b.n_searched
            assert rel_err(a.H, b.H) <= 1e-7
        assert np.abs(Rg - Ro).max() <= POSE_TOL and np.abs(tg - to).max() <= POSE_TOL
        assert_state_equal(fg.download_state(), fo.download_state(), float_tol=1e-9)
        fg.release()
        # --- keyframe: snapshot, world transform (float), insert (both sides use the oracle's pose estimate)
        T_W = np.concatenate([Ro.astype(np.float32).reshape(9), to.astype(np.float32)])[None, :]
        new_g, new_o = mg.snapshot(), mo.snapshot()
        new_g.insert_scan(sc, T_W[0, :9], T_W[0, 9:])
        W = rec_o.copy()
        oracle.transform_f32(W, T_W)
        new_o.insert(W)
        mg.release()
        mg, mo = new_g, new_o
        assert_maps_equal(mg, mo)
        sc.release()
        ds.release()
    if n_scans >= 10:
        assert mo.size()[2] >= 10  # the LRU clock crossed a clear cycle
    mg.release()


@pytest.mark.parametrize("name", ["ouster", "ouster_odyssey", "ouster_r8", "velodyne", "velodyne_anybotics", "hesai", "rslidar",
                                  "livox", "livox_custom2"])
def test_pointcloud2_decode_matches_oracle(ctx, name):
    """lidar::Manager::prepareInput (manager.cpp:149-383) on the device vs the numpy restatement, for the nine vendor
    layouts of point.hpp:41-178, full-resolution and skipped (some points stamped before the header: dropped through the
    reference's uint32 wrap); then the decoded scan flows into deskew + gather like in the callback."""
    import decode_ref
    from cloud_layouts import default_filter, make_cloud
    from mimosa_b200 import Scan

    rng = np.random.default_rng(200)
    data, lay = make_cloud(name, 60000, rng, early_frac=0.01)
    for full, skip, ring_skip in ((1, 4, 2), (0, 4, 1), (1, 1, 1)):
        f = default_filter(create_full_res_pointcloud=full, point_skip_divisor=skip, ring_skip_divisor=ring_skip)
        want = decode_ref.prepare_input(data, lay, f)
        sc, geo, pose_index, unique_ns, last = Scan.from_cloud(ctx, data, lay, f)
        got = sc.download()
        assert got.shape == want[0].shape and got.shape[0] > 100
        assert np.array_equal(got.view(np.uint32), want[0].view(np.uint32))
        assert np.array_equal(geo, want[1]) and np.array_equal(pose_index, want[2])
        assert np.array_equal(unique_ns, want[3]) and last == want[4]
        # one pose per timestamp -> deskew -> the geometric subset, as lidar::Manager::callback chains them
        poses = np.tile(np.concatenate([np.eye(3).reshape(9), [0.0, 0.0, 0.0]]).astype(np.float32), (unique_ns.size, 1))
        poses[:, 9] = np.linspace(0.0, 0.2, unique_ns.size, dtype=np.float32)
        sc.deskew(pose_index, poses)
        sub = sc.gather(geo)
        expect = want[0].copy()
        expect[:, 0] = expect[:, 0] + poses[pose_index, 9]
        assert np.array_equal(sub.download().view(np.uint32), expect[geo].view(np.uint32))
        sub.release()
        sc.release()
