"""mb_scan_from_cloud_ordered: the message re-orderings of lidar::Manager::prepareInput (transpose_pointcloud,
organize_pointcloud_by_ring; manager.cpp:179-243) as an index map in front of the device decode, against the oracle
(oracle/decode_ref.py: re-order the message, then decode).  First run on a B200 in round 2 (profiles/r2_experiments.md)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("name,width,height,transpose,by_ring", [
    ("ouster", 500, 120, True, False),     # RSAiry-style transposition of an organised cloud
    ("hesai", 60000, 1, False, True),      # JT128-style unorganised cloud, records in arbitrary ring order
    ("velodyne", 1, 60000, True, True),    # transposition yields height 1, so the ring re-ordering applies too
    ("ouster", 500, 120, False, True),     # organised: the ring flag must change nothing
    ("rslidar", 96, 680, True, False),     # RSAiry: the sensor the transposition exists for (manager.cpp:179-198)
    ("velodyne_anybotics", 1, 60000, True, True),  # float ring numbers: ordered by ring, NOT filtered by ring (:318-330)
    ("ouster_r8", 60000, 1, False, True),  # uint8 ring numbers
])
def test_ordered_decode_matches_oracle(ctx, name, width, height, transpose, by_ring):
    import decode_ref
    from cloud_layouts import default_filter, make_cloud
    from mimosa_b200 import Scan

    rng = np.random.default_rng(300)
    n = width * height
    data, lay = make_cloud(name, n, rng, n_rings=64)
    data = data[rng.permutation(n)]
    for full, skip, ring_skip in ((1, 4, 2), (0, 4, 1)):
        f = default_filter(create_full_res_pointcloud=full, point_skip_divisor=skip, ring_skip_divisor=ring_skip)
        ordered, *_ = decode_ref.reorder_cloud(data, width, height, transpose, by_ring, lay.off_ring, lay.ring_type)
        want = decode_ref.prepare_input(ordered, lay, f)
        sc, geo, pose_index, unique_ns, last = Scan.from_cloud(ctx, data, lay, f, width=width, height=height,
                                                               transpose_pointcloud=transpose, organize_pointcloud_by_ring=by_ring)
        got = sc.download()
        assert got.shape == want[0].shape and got.shape[0] > 100
        assert np.array_equal(got.view(np.uint32), want[0].view(np.uint32))
        assert np.array_equal(geo, want[1]) and np.array_equal(pose_index, want[2])
        assert np.array_equal(unique_ns, want[3]) and last == want[4]
        sc.release()
