"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (the reference has no tests for this path).

numpy restatement of lidar::Manager::prepareInput (mimosa/src/lidar/manager.cpp:149-383) for a PointCloud2
payload described by field offsets (the nine vendor structs of mimosa/include/mimosa/lidar/point.hpp:41-178 differ
only in where and how intensity / time / ring / tag are stored).  Follows the reference's sequential loop:
candidates i = 0, s, 2s, ... with s = create_full_res ? 1 : point_skip_divisor (:247-250); NaN (:255), Livox tag
(:258-264), intensity (:267-278), range (:281-282) filters; time decode (:285-304); (float)t_ns > ns_max (:306);
points_full row = (x, y, z + z_offset, intensity, t_ns, i, sqrt(range_sq)) (:312-313); geometric idx when
i % point_skip == 0 and ring % ring_skip == 0 (:317-335); timestamps sorted and grouped (:341-371).
The two message re-orderings in front of that loop — transpose_pointcloud (:179-198) and
organize_pointcloud_by_ring (:204-243) — are restated by reorder_cloud() (vectorised) and reorder_cloud_loops()
(the reference's loops, literally, for small inputs); prepare_input() runs on the re-ordered message.
"""
import numpy as np


def reorder_cloud_loops(data, width, height, transpose_pointcloud, organize_pointcloud_by_ring, off_ring, ring_type):
    """manager.cpp:179-243 loop by loop (pure Python: small inputs only).  data: (n, point_step) uint8 with
    n == width * height.  Returns (re-ordered data, width, height)."""
    data = np.ascontiguousarray(data, np.uint8)
    n = data.shape[0]
    assert n == width * height
    if transpose_pointcloud:  # :186-199
        out = np.empty_like(data)
        t_width, t_height = height, width
        for i in range(n):
            current_row, current_col = i // width, i % width
            new_row, new_col = current_col, current_row
            out[new_row * t_width + new_col] = data[i]
        data, width, height = out, t_width, t_height
    if organize_pointcloud_by_ring and height == 1:  # :210-241
        num_rings = 128
        if ring_type == 2:  # float ring (PointVelodyneAnybotics): `ring_counts[point.ring]` truncates to an index
            ring = [int(np.frombuffer(bytes(data[i, off_ring:off_ring + 4]), "<f4")[0]) for i in range(n)]
        else:
            size = 2 if ring_type == 0 else 1
            ring = [int.from_bytes(bytes(data[i, off_ring:off_ring + size]), "little") for i in range(n)]
        ring_counts = [0] * num_rings
        for r in ring:
            ring_counts[r] += 1
        ring_offsets, offset = [0] * num_rings, 0
        for r in range(num_rings):
            ring_offsets[r] = offset
            offset += ring_counts[r]
        out = np.empty_like(data)
        cursors = list(ring_offsets)
        for i in range(n):
            out[cursors[ring[i]]] = data[i]
            cursors[ring[i]] += 1
        data = out
    return data, width, height


def reorder_cloud(data, width, height, transpose_pointcloud, organize_pointcloud_by_ring, off_ring, ring_type):
    """Same as reorder_cloud_loops, vectorised: the transposition is a swap of the grid's axes, the counting sort a
    stable sort by ring number."""
    data = np.ascontiguousarray(data, np.uint8)
    n, step = data.shape
    assert n == width * height
    if transpose_pointcloud and n:
        data = np.ascontiguousarray(data.reshape(height, width, step).swapaxes(0, 1)).reshape(n, step)
        width, height = height, width
    if organize_pointcloud_by_ring and height == 1 and n:
        ring = _ring(data, off_ring, ring_type)
        data = data[np.argsort(ring, kind="stable")]
    return data, width, height


def _field(data, off, dtype):
    n = data.shape[0]
    size = np.dtype(dtype).itemsize
    return np.ascontiguousarray(data[:, off:off + size]).view(dtype).reshape(n)


def _ring(data, off_ring, ring_type):
    if ring_type == 2:
        return _field(data, off_ring, np.float32).astype(np.int64)
    return _field(data, off_ring, np.uint16 if ring_type == 0 else np.uint8).astype(np.int64)


def _wrap_u32(td):
    """`uint32_t t_ns = <double>` as x86-64 compilers emit it (manager.cpp:285-304): truncate to int64, keep the low 32
    bits — negative offsets wrap to ~4.29e9 and are then dropped by the ns_max test."""
    return (np.trunc(td).astype(np.int64) & np.int64(0xFFFFFFFF)).astype(np.uint32)


def prepare_input(data, layout, filt):
    """data: (n, point_step) uint8.  layout / filt: objects with the fields of mb_cloud_layout / mb_input_filter.
    Returns (points_full (m, 8) float32 lidar::Point rows, geometric_idx, pose_index, unique_ns, last_point_ns)."""
    data = np.ascontiguousarray(data, np.uint8).reshape(-1, layout.point_step)
    s = 1 if filt.create_full_res_pointcloud else filt.point_skip_divisor
    idx = np.arange(0, data.shape[0], s)
    d = data[idx]
    x, y, z = (_field(d, o, np.float32) for o in (layout.off_x, layout.off_y, layout.off_z))
    ok = ~(np.isnan(x) | np.isnan(y) | np.isnan(z))
    if layout.off_tag >= 0:
        tag = d[:, layout.off_tag]
        ok &= ((tag & 0x30) == 0x10) | ((tag & 0x30) == 0x00)
    imin, imax = np.float32(filt.intensity_min), np.float32(filt.intensity_max)
    if layout.intensity_type == 0:
        inten = _field(d, layout.off_intensity, np.float32)
        with np.errstate(invalid="ignore"):
            ok &= ~(np.isnan(inten) | (inten < imin) | (inten > imax))
    else:
        inten = _field(d, layout.off_intensity, np.uint16).astype(np.float32)
        ok &= ~((inten < imin) | (inten > imax))
    with np.errstate(invalid="ignore", over="ignore"):
        range_sq = (x * x + y * y) + z * z  # float32, left to right
        rmin2 = np.float32(filt.range_min) * np.float32(filt.range_min)
        rmax2 = np.float32(filt.range_max) * np.float32(filt.range_max)
        ok &= ~((range_sq < rmin2) | (range_sq > rmax2))
    if layout.time_type == 0:
        t_ns = _field(d, layout.off_time, np.uint32).astype(np.uint32)
    else:
        with np.errstate(invalid="ignore"):
            if layout.time_type == 1:
                td = _field(d, layout.off_time, np.float32).astype(np.float64) * 1e9
            elif layout.time_type == 2:
                td = (_field(d, layout.off_time, np.float64) - filt.header_ts) * 1e9
            else:
                td = _field(d, layout.off_time, np.float64) - filt.header_ts * 1e9
            td = np.where(ok, td, 0.0)
        t_ns = _wrap_u32(td)
    ok &= ~(t_ns.astype(np.float32) > np.float32(filt.ns_max))
    kept = np.flatnonzero(ok)
    m = kept.size
    out = np.zeros((m, 8), np.float32)
    out[:, 0], out[:, 1] = x[kept], y[kept]
    out[:, 2] = z[kept] + np.float32(filt.z_offset)
    out[:, 3] = 1.0
    out[:, 4] = inten[kept]
    out[:, 5] = t_ns[kept].view(np.float32)
    out[:, 6] = idx[kept].astype(np.uint32).view(np.float32)
    out[:, 7] = np.sqrt(range_sq[kept])
    gk = idx[kept] % filt.point_skip_divisor == 0
    if layout.ring_filter:  # :318-330; PointVelodyneAnybotics carries a ring but is not filtered by it
        ring = _ring(d, layout.off_ring, layout.ring_type)[kept]
        gk &= ring % filt.ring_skip_divisor == 0
    geometric_idx = np.flatnonzero(gk).astype(np.uint32)
    unique_ns, pose_index = np.unique(t_ns[kept], return_inverse=True)
    last = int(unique_ns[-1]) if m else 0
    return out, geometric_idx, pose_index.astype(np.uint32), unique_ns.astype(np.uint32), last
