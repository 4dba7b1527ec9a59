"""The EXPERIMENTAL k-NN kernels (lanes per query: mimosa_b200/csrc/mb_search_coop.cuh, MB_KNN_VARIANT=coop4 / coop8 / ...;
warp-wide chunk queue: knn_thread<K, true> in mb_search.cuh, MB_KNN_VARIANT=threadq — the latter has not run on a GPU yet)
through the C ABI against the oracle's iVox: indices, squared distances and found flags bit-exact, like the default
kernel's tests in test_gpu_parity.py.  The variants are not the product path (measured slower than the default at full
load, profiles/r1_experiments.md session 4), so this file only runs when MB_TEST_EXPERIMENTAL=1; their logic is covered
on the CPU by tests/test_search_host.py::test_coop_* and their k = 5 kernels were compared bit for bit with the default
kernel on a B200 by tools/knn_variants.py."""
import os

import numpy as np
import pytest

from mimosa_b200 import IncrementalVoxelMap

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MB_TEST_EXPERIMENTAL") != "1", reason="experimental k-NN variants: set MB_TEST_EXPERIMENTAL=1")]


@pytest.fixture(params=["coop4", "coop8", "coop4p", "coop8p", "threadq"])
def variant(request):
    old = os.environ.get("MB_KNN_VARIANT")
    os.environ["MB_KNN_VARIANT"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("MB_KNN_VARIANT", None)
    else:
        os.environ["MB_KNN_VARIANT"] = old


def _maps(ctx, oracle, **kw):
    args = dict(leaf=1.0, min_dist=0.2, cap=20, nbr_mode=19, lru_horizon=1000)
    args.update(kw)
    return IncrementalVoxelMap(ctx, **args), oracle.IVoxRef(args["leaf"], args["min_dist"], args["cap"], args["nbr_mode"], args["lru_horizon"])


def _same(mg, mo, q, k):
    ig, dg, og = mg.knn_search(q, k)
    io, do, oo = mo.knn_search(q, k)
    assert np.array_equal(og, oo)
    assert np.array_equal(ig, io)
    assert np.array_equal(dg, do)


@pytest.mark.parametrize("mode", [1, 7, 19, 27])
def test_coop_knn_matches_oracle(ctx, oracle, variant, mode):
    rng = np.random.default_rng(210 + mode)
    mg, mo = _maps(ctx, oracle, nbr_mode=mode, min_dist=0.05)  # dense buckets: many voxels at cap 20
    pts = rng.uniform(-4, 4, (60000, 3)).astype(np.float32)
    mg.insert(pts)
    mo.insert(pts)
    q = rng.uniform(-5, 5, (4001, 3))  # ragged last block
    q[:50] = pts[:50].astype(np.float64)
    for k in (1, 3, 5, 8):
        _same(mg, mo, q, k)
    mg.release()


def test_coop_knn_ties_sparse_and_world(ctx, oracle, variant):
    import synth

    g = np.stack(np.meshgrid(*[np.arange(-8, 8) * 0.25 + 0.125] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    q = np.stack(np.meshgrid(*[np.arange(-6, 6) * 0.25] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    for mode in (7, 19, 27):
        mg, mo = _maps(ctx, oracle, nbr_mode=mode, min_dist=0.0)
        mg.insert(g)
        mo.insert(g)
        for k in (5, 8):
            _same(mg, mo, q, k)
        mg.release()
    mg, mo = _maps(ctx, oracle)
    _same(mg, mo, np.array([[0.5, 0.5, 0.5], [100.0, 100.0, 100.0]]), 5)  # empty map
    pts = np.array([[0.4, 0.5, 0.5], [0.9, 0.5, 0.5], [1.2, 0.5, 0.5]], np.float32)
    mg.insert(pts)
    mo.insert(pts)
    _same(mg, mo, np.array([[0.5, 0.5, 0.5], [100.0, 100.0, 100.0]]), 5)  # fewer than k
    mg.release()
    rng = synth.rng_for(77)
    mg, mo = _maps(ctx, oracle)
    world = synth.sample_world(200000, 40.0, rng)
    mg.insert(world)
    mo.insert(world)
    q = world[rng.integers(0, world.shape[0], 20000), :3].astype(np.float64) + rng.normal(0, 0.1, (20000, 3))
    _same(mg, mo, q, 5)
    mg.release()
