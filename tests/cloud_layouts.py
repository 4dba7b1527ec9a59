"""Synthetic sensor_msgs/PointCloud2 payloads in the vendor layouts of mimosa/include/mimosa/lidar/point.hpp:41-178."""
import numpy as np

from mimosa_b200.capi import CloudLayout, InputFilter

# numpy dtypes with the serialized field offsets PCL's registration macros produce for the padded structs
OUSTER = np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ring"],
                   "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "<u2"],
                   "offsets": [0, 4, 8, 16, 20, 24, 26], "itemsize": 32})
OUSTER_ODYSSEY = np.dtype({"names": ["x", "y", "z", "t", "reflectivity", "near_ir"],
                           "formats": ["<f4", "<f4", "<f4", "<u4", "<u2", "<u2"], "offsets": [0, 4, 8, 16, 20, 22], "itemsize": 32})
VELODYNE = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                     "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"], "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
HESAI = np.dtype({"names": ["x", "y", "z", "intensity", "timestamp", "ring"],
                  "formats": ["<f4", "<f4", "<f4", "<f4", "<f8", "<u2"], "offsets": [0, 4, 8, 16, 24, 32], "itemsize": 48})
LIVOX = np.dtype({"names": ["x", "y", "z", "intensity", "tag", "line", "timestamp"],
                  "formats": ["<f4", "<f4", "<f4", "<f4", "u1", "u1", "<f8"], "offsets": [0, 4, 8, 16, 20, 21, 24], "itemsize": 32})

OUSTER_R8 = np.dtype({"names": ["x", "y", "z", "intensity", "t", "reflectivity", "ring"],
                      "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u2", "u1"],
                      "offsets": [0, 4, 8, 16, 20, 24, 26], "itemsize": 32})
LIVOX_CUSTOM2 = np.dtype({"names": ["x", "y", "z", "t", "intensity", "tag", "line"],
                          "formats": ["<f4", "<f4", "<f4", "<u4", "<f4", "u1", "u1"], "offsets": [0, 4, 8, 12, 16, 20, 21], "itemsize": 32})
VELODYNE_ANYBOTICS = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                               "formats": ["<f4", "<f4", "<f4", "<f4", "<f4", "<f4"], "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})
RSLIDAR = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "timestamp"],
                    "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f8"], "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})

# CloudLayout(point_step, off_x, off_y, off_z, off_intensity, intensity_type, off_time, time_type, off_ring, ring_type, off_tag,
#             ring_filter) for the nine vendor structs
LAYOUTS = {
    "ouster": (OUSTER, CloudLayout(32, 0, 4, 8, 16, 0, 20, 0, 26, 0, -1, 1)),
    "ouster_odyssey": (OUSTER_ODYSSEY, CloudLayout(32, 0, 4, 8, 20, 1, 16, 0, -1, 0, -1, 0)),
    "ouster_r8": (OUSTER_R8, CloudLayout(32, 0, 4, 8, 16, 0, 20, 0, 26, 1, -1, 1)),
    "velodyne": (VELODYNE, CloudLayout(32, 0, 4, 8, 16, 0, 24, 1, 20, 0, -1, 1)),
    "velodyne_anybotics": (VELODYNE_ANYBOTICS, CloudLayout(32, 0, 4, 8, 16, 0, 24, 1, 20, 2, -1, 0)),
    "hesai": (HESAI, CloudLayout(48, 0, 4, 8, 16, 0, 24, 2, 32, 0, -1, 1)),
    "rslidar": (RSLIDAR, CloudLayout(32, 0, 4, 8, 16, 0, 24, 2, 20, 0, -1, 1)),
    "livox": (LIVOX, CloudLayout(32, 0, 4, 8, 16, 0, 24, 3, -1, 0, 20, 0)),
    "livox_custom2": (LIVOX_CUSTOM2, CloudLayout(32, 0, 4, 8, 16, 0, 12, 0, -1, 0, 20, 0)),
}


def make_cloud(name, n, rng, header_ts=1.7e9, n_rings=128, nan_frac=0.01, early_frac=0.0):
    dt, layout = LAYOUTS[name]
    c = np.zeros(n, dt)
    xyz = rng.normal(0, 20, (n, 3)).astype(np.float32)
    xyz[rng.random(n) < nan_frac, rng.integers(0, 3)] = np.nan
    c["x"], c["y"], c["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    ring = (np.arange(n) // max(n // n_rings, 1)) % n_rings
    col = np.arange(n) % max(n // n_rings, 1)
    t_ns = (col * 97).astype(np.uint32) * np.uint32(1000)  # many points share a timestamp (one per column)
    if "intensity" in dt.names:
        inten = rng.uniform(-5, 300, n).astype(np.float32)
        inten[rng.random(n) < nan_frac] = np.nan
        c["intensity"] = inten
    if "reflectivity" in dt.names:
        c["reflectivity"] = rng.integers(0, 4000, n).astype(np.uint16)
    if "ring" in dt.names:
        c["ring"] = ring.astype(dt["ring"])
    if "tag" in dt.names:
        c["tag"] = rng.integers(0, 256, n).astype(np.uint8)
    if "t" in dt.names:
        c["t"] = t_ns
    if "time" in dt.names:
        c["time"] = (t_ns.astype(np.float64) * 1e-9).astype(np.float32)
    if "timestamp" in dt.names:
        if name == "livox":
            c["timestamp"] = header_ts * 1e9 + t_ns.astype(np.float64)
        else:
            c["timestamp"] = header_ts + t_ns.astype(np.float64) * 1e-9
        early = rng.random(n) < early_frac  # points stamped BEFORE the header: the reference's uint32 conversion wraps
        c["timestamp"][early] -= (5e6 if name == "livox" else 5e-3)
    return c.view(np.uint8).reshape(n, dt.itemsize), layout


def default_filter(header_ts=1.7e9, **kw):
    f = InputFilter(intensity_min=0.0, intensity_max=250.0, range_min=2.0, range_max=60.0, ns_max=9.0e6,
                    point_skip_divisor=4, ring_skip_divisor=1, create_full_res_pointcloud=0, z_offset=0.036,
                    header_ts=header_ts)
    for k, v in kw.items():
        setattr(f, k, v)
    return f
