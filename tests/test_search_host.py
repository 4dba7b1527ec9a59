"""The search code the kernels run (mimosa_b200/csrc/mb_search.cuh: block probes, occupancy masks, pruning
bounds, deferred ordered insertion, tie order) compiled for the HOST, one emulated lane per query
(tests/host_shim/search_shim.cpp), against the oracle's iVox k-NN: indices and squared distances bit-exact for
every neighbourhood mode, k, leaf size and early-prefetch radius.  CPU-only coverage of the hot path's logic;
the GPU parity tests check the same kernels through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim():
    src = os.path.join(HERE, "host_shim", "search_shim.cpp")
    out = os.path.join(HERE, "host_shim", "libsearch_shim.so")
    hdrs = [os.path.join(HERE, "..", "mimosa_b200", "csrc", h) for h in ("mb_search.cuh", "mb_math.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.run(["g++", "-O2", "-std=c++20", "-ffp-contract=off", "-frounding-math", "-fPIC", "-shared", "-pthread", src,
                        "-o", out], check=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def shim_knn(shim, m, q, k, nbr_mode, leaf, pref_frac=0.4, warp=False):
    coords, counts, _, pts, _ = m.download()
    q = np.ascontiguousarray(q, np.float64).reshape(-1, 3)
    nq = q.shape[0]
    idx, d2, ok = np.empty((nq, k), np.uint64), np.empty((nq, k), np.float64), np.empty(nq, np.uint8)
    leaf = float(np.float32(leaf))  # the map's leaf size is a float (mb_map_create, IVoxRef): 1 / (double)leaf_f32
    rc = (shim.shim_knn_warp if warp else shim.shim_knn)(_p(coords), _p(counts), _p(pts), C.c_uint32(coords.shape[0]), C.c_int(m.cap), C.c_int(nbr_mode),
                       C.c_double(leaf), C.c_double(pref_frac), _p(q), C.c_size_t(nq), C.c_int(k), _p(idx), _p(d2), _p(ok))
    assert rc == 0
    return idx, d2, ok.astype(bool)


def check(shim, oracle, pts, q, k, mode, leaf, min_dist, pref_frac=0.4, cap=20, warp=False):
    m = oracle.IVoxRef(leaf, min_dist, cap, mode, 1000)
    m.insert(pts)
    io, do, oo = m.knn_search(q, k)
    ih, dh, oh = shim_knn(shim, m, q, k, mode, leaf, pref_frac, warp)
    assert np.array_equal(oo, oh)
    assert np.array_equal(io[oo], ih[oo]), f"indices differ (mode {mode}, k {k})"
    assert np.array_equal(do[oo], dh[oo]), "squared distances differ"
    return int(oo.sum())


@pytest.mark.parametrize("mode", [1, 7, 19, 27])
@pytest.mark.parametrize("k", [1, 3, 5, 8])
def test_search_matches_oracle_world(shim, oracle, mode, k):
    import synth

    rng = synth.rng_for(300 + mode + k)
    pts = synth.sample_world(60000, 30.0, rng)
    q = pts[rng.integers(0, pts.shape[0], 3000), :3].astype(np.float64) + rng.normal(0, 0.15, (3000, 3))
    q = np.concatenate([q, rng.uniform(-35, 35, (500, 3))])  # some far from everything
    n_ok = check(shim, oracle, pts, q, k, mode, 1.0, 0.2)
    assert n_ok > 500


@pytest.mark.parametrize("leaf,min_dist", [(0.5, 0.15), (2.0, 0.0), (0.25, 0.05)])
def test_search_matches_oracle_leaf_sizes(shim, oracle, leaf, min_dist):
    rng = np.random.default_rng(7)
    pts = rng.uniform(-6, 6, (40000, 3)).astype(np.float32)  # volumetric: every neighbour voxel occupied
    q = rng.uniform(-6.5, 6.5, (3000, 3))
    for pref in (0.0, 0.4, 2.0):  # the early-prefetch radius must not change results
        assert check(shim, oracle, pts, q, 5, 19, leaf, min_dist, pref) > 1000


def test_search_ties_and_negative_coordinates(shim, oracle):
    # lattice points: many exactly equal distances, resolved by visiting order; queries on voxel faces / corners
    g = np.arange(-8, 8) * 0.5
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(3)
    pts = pts[rng.permutation(pts.shape[0])]
    q = np.concatenate([pts[:1500].astype(np.float64) + 0.25, pts[:1500].astype(np.float64),
                        np.round(rng.uniform(-4, 4, (1000, 3)))])
    for mode in (7, 19, 27):
        for k in (5, 8):
            check(shim, oracle, pts, q, k, mode, 1.0, 0.0)


def test_search_sparse_voxels_and_small_cap(shim, oracle):
    rng = np.random.default_rng(11)
    pts = rng.uniform(-20, 20, (3000, 3)).astype(np.float32)  # ~0.05 points per voxel: own voxel mostly empty
    q = rng.uniform(-20, 20, (4000, 3))
    check(shim, oracle, pts, q, 5, 27, 1.0, 0.0)
    check(shim, oracle, pts, q, 2, 19, 1.0, 0.0)
    dense = rng.uniform(-2, 2, (20000, 3)).astype(np.float32)
    check(shim, oracle, dense, rng.uniform(-2, 2, (2000, 3)), 5, 19, 1.0, 0.0, cap=7)  # cap not a multiple of 4
    check(shim, oracle, dense, rng.uniform(-2, 2, (2000, 3)), 5, 19, 1.0, 0.0, cap=31)


@pytest.mark.parametrize("mode,k", [(19, 5), (27, 8), (7, 3)])
def test_search_32_lane_warp_emulation(shim, oracle, mode, k):
    """The same code run as a 32-lane warp (32 host threads in lock step, warp intrinsics exchanged through a
    barrier): every lane must reach the same intrinsics in the same order (or the emulation dead-locks and the
    test times out) and the max-over-lanes loops, deferred insertion and drains must still give every lane its own
    exact result.  Lanes of a warp are deliberately heterogeneous: dense and sparse regions, far queries, a ragged
    last warp."""
    import synth

    rng = synth.rng_for(900 + mode + k)
    pts = np.concatenate([synth.sample_world(40000, 25.0, rng), rng.uniform(-25, 25, (3000, 3)).astype(np.float32)])
    q = np.concatenate([pts[rng.integers(0, pts.shape[0], 500), :3].astype(np.float64) + rng.normal(0, 0.2, (500, 3)),
                        rng.uniform(-30, 30, (141, 3))])
    q = q[rng.permutation(q.shape[0])]  # 641 queries: 20 full warps + one lane
    assert check(shim, oracle, pts, q, k, mode, 1.0, 0.2, warp=True) > 300
