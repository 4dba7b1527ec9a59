"""CPU known-answer tests pinning the numpy restatement of lidar::Manager::prepareInput (oracle/decode_ref.py)."""
import numpy as np

import decode_ref
from cloud_layouts import LAYOUTS, OUSTER, default_filter, make_cloud


def _ouster(rows):
    c = np.zeros(len(rows), OUSTER)
    for i, r in enumerate(rows):
        c[i] = r
    return c.view(np.uint8).reshape(len(rows), 32)


def test_filters_hand_checked():
    lay = LAYOUTS["ouster"][1]
    #        x     y    z    inten  t_ns   refl ring
    rows = [(3.0, 0.0, 0.0, 10.0, 500, 0, 0),       # kept, geometric (i=0)
            (3.0, 4.0, 0.0, 10.0, 100, 0, 1),       # kept (full res), not geometric (i=1, skip 2)
            (np.nan, 0.0, 0.0, 10.0, 0, 0, 0),      # NaN
            (1.0, 0.0, 0.0, 10.0, 0, 0, 0),         # range < 2
            (3.0, 0.0, 0.0, 300.0, 0, 0, 0),        # intensity > max
            (3.0, 0.0, 0.0, 10.0, 20_000_000, 0, 0),  # t_ns > ns_max
            (0.0, 0.0, 5.0, 10.0, 100, 0, 3),       # kept, i=6 geometric by skip but ring 3 % 2 != 0
            (0.0, 70.0, 0.0, 10.0, 100, 0, 0),      # range > 60
            (0.0, 6.0, 8.0, 20.0, 500, 0, 2)]       # kept, geometric (i=8, ring 2)
    f = default_filter(create_full_res_pointcloud=1, point_skip_divisor=2, ring_skip_divisor=2)
    pts, geo, pose_index, unique_ns, last = decode_ref.prepare_input(_ouster(rows), lay, f)
    assert pts[:, 6].view(np.uint32).tolist() == [0, 1, 6, 8]
    assert pts[:, 5].view(np.uint32).tolist() == [500, 100, 100, 500]
    assert np.allclose(pts[:, 7], [3.0, 5.0, 5.0, 10.0]) and np.allclose(pts[:, 2], [0.036, 0.036, 5.036, 8.036])
    assert geo.tolist() == [0, 3]
    assert unique_ns.tolist() == [100, 500] and pose_index.tolist() == [1, 0, 0, 1] and last == 500
    # without full resolution only every point_skip-th input point is even looked at
    f2 = default_filter(create_full_res_pointcloud=0, point_skip_divisor=2, ring_skip_divisor=1)
    pts2, geo2, *_ = decode_ref.prepare_input(_ouster(rows), lay, f2)
    assert pts2[:, 6].view(np.uint32).tolist() == [0, 6, 8] and geo2.tolist() == [0, 1, 2]


def test_time_encodings_agree():
    rng = np.random.default_rng(3)
    ref = None
    for name in ("ouster", "velodyne", "hesai"):
        data, lay = make_cloud(name, 4096, np.random.default_rng(3))
        pts, geo, pi, uns, last = decode_ref.prepare_input(data, lay, default_filter())
        t = pts[:, 5].view(np.uint32)
        if ref is None:
            ref = t
        else:  # float32 seconds / float64 absolute seconds decode to the same nanoseconds within rounding
            assert t.shape == ref.shape and np.abs(t.astype(np.int64) - ref.astype(np.int64)).max() <= 1000
        assert np.array_equal(uns[pi], t) and np.all(np.diff(uns.astype(np.int64)) > 0)


def test_livox_tag_and_reflectivity_layouts():
    data, lay = make_cloud("livox", 2000, np.random.default_rng(4))
    pts, geo, *_ = decode_ref.prepare_input(data, lay, default_filter(create_full_res_pointcloud=1))
    tags = data[pts[:, 6].view(np.uint32), 20]
    assert np.all(((tags & 0x30) == 0x10) | ((tags & 0x30) == 0x00)) and 0 < pts.shape[0] < 2000
    data, lay = make_cloud("ouster_odyssey", 2000, np.random.default_rng(5))
    pts, geo, *_ = decode_ref.prepare_input(data, lay, default_filter(create_full_res_pointcloud=1))
    assert np.all(pts[:, 4] <= 250.0) and pts.shape[0] > 0


def _shuffled_ring_cloud(name, n, rng, n_rings=32):
    """A cloud whose records are NOT in row-major ring order (the Hesai JT128 case the ring re-ordering exists for)."""
    data, lay = make_cloud(name, n, rng, n_rings=n_rings)
    return data[rng.permutation(n)], lay


def test_reorder_transposition_and_ring_order_hand_checked():
    lay = LAYOUTS["ouster"][1]
    # 2 rows x 3 columns, x = 10 * row + column, ring = column
    rows = [(10.0 * r + c, 0.0, 0.0, 1.0, 0, 0, c) for r in range(2) for c in range(3)]
    data = _ouster(rows)
    t, w, h = decode_ref.reorder_cloud(data, 3, 2, True, False, lay.off_ring, lay.ring_type)
    assert (w, h) == (2, 3)
    assert t[:, :4].copy().view(np.float32).ravel().tolist() == [0.0, 10.0, 1.0, 11.0, 2.0, 12.0]  # column-major walk
    # unorganised (height 1): stable by ring — equal rings keep the message order
    flat, w, h = decode_ref.reorder_cloud(data, 6, 1, False, True, lay.off_ring, lay.ring_type)
    assert flat[:, :4].copy().view(np.float32).ravel().tolist() == [0.0, 10.0, 1.0, 11.0, 2.0, 12.0]
    # an organised cloud is left alone by the ring re-ordering (manager.cpp:210)
    same, *_ = decode_ref.reorder_cloud(data, 3, 2, False, True, lay.off_ring, lay.ring_type)
    assert np.array_equal(same, data)
    # transposing an N x 1 cloud makes it unorganised, so both apply in sequence
    both, w, h = decode_ref.reorder_cloud(data, 1, 6, True, True, lay.off_ring, lay.ring_type)
    assert (w, h) == (6, 1) and np.array_equal(both, flat)


def test_reorder_vectorised_equals_reference_loops():
    rng = np.random.default_rng(8)
    for name, w, h in (("ouster", 37, 16), ("hesai", 600, 1), ("velodyne", 1, 480), ("ouster", 1, 1)):
        data, lay = _shuffled_ring_cloud(name, w * h, rng)
        for tr in (False, True):
            for org in (False, True):
                a = decode_ref.reorder_cloud(data, w, h, tr, org, lay.off_ring, lay.ring_type)
                b = decode_ref.reorder_cloud_loops(data, w, h, tr, org, lay.off_ring, lay.ring_type)
                assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]


def test_points_stamped_before_the_header_are_dropped_by_the_uint32_wrap():
    """manager.cpp:291-306: `uint32_t t_ns = (timestamp - header) * 1e9` of a NEGATIVE offset wraps (x86-64 conversion) to
    ~4.29e9 ns, which the ns_max test then rejects; a saturating conversion would keep the point with t_ns = 0."""
    n = 4000
    for name in ("hesai", "rslidar", "livox"):
        clean, lay = make_cloud(name, n, np.random.default_rng(9), nan_frac=0.0)
        dirty, _ = make_cloud(name, n, np.random.default_rng(9), nan_frac=0.0, early_frac=0.05)
        f = default_filter(create_full_res_pointcloud=1)
        a, *_ = decode_ref.prepare_input(clean, lay, f)
        b, *_ = decode_ref.prepare_input(dirty, lay, f)
        changed = np.flatnonzero((clean != dirty).any(1))
        assert 100 < changed.size < 400
        kept_a, kept_b = set(a[:, 6].view(np.uint32).tolist()), set(b[:, 6].view(np.uint32).tolist())
        assert kept_b == kept_a - set(changed.tolist())  # exactly the early points disappear
    assert decode_ref._wrap_u32(np.array([-5.0, 0.0, 7.9, 4294967296.0 + 3, -4294967296.0 * 3 - 2.5])).tolist() == [4294967291, 0, 7, 3, 4294967294]


def test_anybotics_float_ring_orders_but_does_not_filter():
    rng = np.random.default_rng(10)
    data, lay = _shuffled_ring_cloud("velodyne_anybotics", 480, rng)
    assert lay.ring_type == 2 and lay.ring_filter == 0
    a = decode_ref.reorder_cloud(data, 480, 1, False, True, lay.off_ring, lay.ring_type)
    b = decode_ref.reorder_cloud_loops(data, 480, 1, False, True, lay.off_ring, lay.ring_type)
    assert np.array_equal(a[0], b[0])
    rings = a[0][:, 20:24].copy().view(np.float32).ravel()
    assert np.all(np.diff(rings) >= 0)
    f1, f2 = default_filter(ring_skip_divisor=1), default_filter(ring_skip_divisor=4)
    assert np.array_equal(decode_ref.prepare_input(a[0], lay, f1)[1], decode_ref.prepare_input(a[0], lay, f2)[1])
    # the same cloud read as a plain Velodyne cloud (integer ring, filtered) does depend on the ring divisor
    v, vlay = _shuffled_ring_cloud("velodyne", 480, np.random.default_rng(10))
    assert decode_ref.prepare_input(v, vlay, f1)[1].size > decode_ref.prepare_input(v, vlay, f2)[1].size
