"""Deterministic synthetic worlds, maps and LiDAR scans for the tests and bench.py.

World: ground plane z = GROUND_Z plus a 20 x 20 grid of axis-aligned boxes (10 x 10 x 8 m, 40 m pitch).
Surfaces are sampled uniformly per area with N(0, 0.01^2) noise along the normal — the reference rejects
noise-free planes (lambda_0 < 1e-6, geometric_factor.hpp:202).  Scans are ray-cast analytically.

Everything is numpy with a seeded PCG64 generator, so the CUDA path and the CPU oracle are fed byte-identical
float32 inputs.  Nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import numpy as np

GROUND_Z = -1.37
BOX_PITCH = 40.0
BOX_HALF = 5.0
BOX_HEIGHT = 8.0
BOX_GRID = 20
SEED0 = 20260101


def rng_for(config_id: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(SEED0 + config_id))


def box_centers(grid: int = BOX_GRID) -> np.ndarray:
    c = (np.arange(grid) - (grid - 1) / 2.0) * BOX_PITCH
    gx, gy = np.meshgrid(c, c, indexing="ij")
    return np.stack([gx.ravel(), gy.ravel()], axis=1)


# ---- surface sampling -------------------------------------------------------------------------------
def sample_ground(n: int, half_extent: float, rng, noise: float = 0.01) -> np.ndarray:
    p = np.empty((n, 3), dtype=np.float64)
    p[:, 0] = rng.uniform(-half_extent, half_extent, n)
    p[:, 1] = rng.uniform(-half_extent, half_extent, n)
    p[:, 2] = GROUND_Z + rng.normal(0.0, noise, n)
    return p.astype(np.float32)


def boxes_within(half_extent: float) -> np.ndarray:
    """Centres of the world's boxes that lie entirely inside |x|,|y| <= half_extent."""
    c = box_centers()
    return c[(np.abs(c[:, 0]) + BOX_HALF <= half_extent) & (np.abs(c[:, 1]) + BOX_HALF <= half_extent)]


def sample_boxes(n: int, rng, noise: float = 0.01, centers=None) -> np.ndarray:
    """Uniform per area over the 4 walls and the top of every box."""
    centers = box_centers() if centers is None else centers
    side = 2 * BOX_HALF
    areas = np.array([side * BOX_HEIGHT] * 4 + [side * side])
    face = rng.choice(5, size=n, p=areas / areas.sum())
    b = rng.integers(0, centers.shape[0], n)
    u = rng.uniform(-BOX_HALF, BOX_HALF, n)
    v = rng.uniform(0.0, 1.0, n)
    d = rng.normal(0.0, noise, n)
    p = np.empty((n, 3), dtype=np.float64)
    cx, cy = centers[b, 0], centers[b, 1]
    z_wall = GROUND_Z + v * BOX_HEIGHT
    # faces 0/1: x = cx -/+ half; 2/3: y = cy -/+ half; 4: top
    for f, (ax, sgn) in enumerate([(0, -1), (0, 1), (1, -1), (1, 1)]):
        m = face == f
        if ax == 0:
            p[m, 0] = cx[m] + sgn * (BOX_HALF + d[m])
            p[m, 1] = cy[m] + u[m]
        else:
            p[m, 0] = cx[m] + u[m]
            p[m, 1] = cy[m] + sgn * (BOX_HALF + d[m])
        p[m, 2] = z_wall[m]
    m = face == 4
    p[m, 0] = cx[m] + u[m]
    p[m, 1] = cy[m] + (v[m] * 2 - 1) * BOX_HALF
    p[m, 2] = GROUND_Z + BOX_HEIGHT + d[m]
    return p.astype(np.float32)


def sample_world(n: int, half_extent: float, rng, noise: float = 0.01) -> np.ndarray:
    """Ground + boxes inside |x|,|y| <= half_extent, area-proportional, shuffled."""
    centers = boxes_within(half_extent)
    ground_area = (2 * half_extent) ** 2
    box_area = centers.shape[0] * (4 * 2 * BOX_HALF * BOX_HEIGHT + (2 * BOX_HALF) ** 2)
    n_box = int(round(n * box_area / (ground_area + box_area))) if centers.shape[0] else 0
    parts = [sample_ground(n - n_box, half_extent, rng, noise)]
    if n_box:
        parts.append(sample_boxes(n_box, rng, noise, centers))
    p = np.concatenate(parts, axis=0)
    return p[rng.permutation(p.shape[0])]


def build_map(inserter, target_points: int, half_extent: float, rng, chunk: int = 1 << 20, size_fn=None,
              max_chunks: int = 200, noise: float = 0.01):
    """Feed world samples to `inserter(xyz_f32)` chunk by chunk until `size_fn()` >= target_points.
    Returns the list of chunks fed (so a second implementation can be fed the identical sequence)."""
    fed = []
    for _ in range(max_chunks):
        pts = sample_world(chunk, half_extent, rng, noise)
        inserter(pts)
        fed.append(pts)
        if size_fn is not None and size_fn() >= target_points:
            break
    return fed


# ---- ray casting ------------------------------------------------------------------------------------
def os0_128_dirs(n_az: int = 1024, n_beams: int = 128, elev_deg=(-45.0, 45.0)) -> np.ndarray:
    """Ouster OS0-128 style pattern, azimuth-major firing order (column by column)."""
    el = np.deg2rad(np.linspace(elev_deg[0], elev_deg[1], n_beams))
    az = np.linspace(0.0, 2 * np.pi, n_az, endpoint=False)
    A, E = np.meshgrid(az, el, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1)
    return d.reshape(-1, 3)


def airy_dirs(n_az: int = 680, n_beams: int = 96) -> np.ndarray:
    """Robosense Airy style hemispherical pattern: 96 beams in [0, 90] deg elevation."""
    return os0_128_dirs(n_az, n_beams, (0.0, 90.0))


def raycast(origin: np.ndarray, dirs: np.ndarray, max_range: float = 100.0) -> np.ndarray:
    """Distance to the first hit of the analytic world along each ray (inf if none within max_range)."""
    o = np.asarray(origin, dtype=np.float64)
    d = np.asarray(dirs, dtype=np.float64)
    n = d.shape[0]
    best = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (GROUND_Z - o[2]) / d[:, 2]
    tg = np.where((tg > 0) & np.isfinite(tg), tg, np.inf)
    best = np.minimum(best, tg)
    centers = box_centers()
    near = centers[np.hypot(centers[:, 0] - o[0], centers[:, 1] - o[1]) < max_range + 2 * BOX_HALF]
    lo_z, hi_z = GROUND_Z, GROUND_Z + BOX_HEIGHT
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
    for c in near:
        lo = np.array([c[0] - BOX_HALF, c[1] - BOX_HALF, lo_z])
        hi = np.array([c[0] + BOX_HALF, c[1] + BOX_HALF, hi_z])
        with np.errstate(invalid="ignore"):
            t1 = (lo - o) * inv
            t2 = (hi - o) * inv
        tmin = np.nanmax(np.minimum(t1, t2), axis=1)
        tmax = np.nanmin(np.maximum(t1, t2), axis=1)
        hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0)
        best = np.where(hit & (tmin < best), tmin, best)
    best[best > max_range] = np.inf
    return best


def rot_from_rpy(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def make_scan(R: np.ndarray, t: np.ndarray, n_points: int, rng, pattern: str = "os0", max_range: float = 100.0,
              min_range: float = 1.0, range_noise: float = 0.02) -> np.ndarray:
    """Ray-cast scan from pose (R, t), returned as (n_points, 8) float32 rows laid out like
    mimosa::lidar::Point (x, y, z, pad, intensity, t, idx, range — point.hpp:18-39), body frame."""
    out = []
    n_az = 1024 if pattern == "os0" else 680
    total = 0
    az_mult = 1
    while total < n_points:
        dirs_b = os0_128_dirs(n_az * az_mult) if pattern == "os0" else airy_dirs(n_az * az_mult)
        dirs_w = dirs_b @ R.T
        rng_d = raycast(t, dirs_w, max_range)
        ok = np.isfinite(rng_d) & (rng_d >= min_range)
        r = rng_d[ok] + rng.normal(0.0, range_noise, int(ok.sum()))
        pts = dirs_b[ok] * r[:, None]
        out = [pts]
        total = pts.shape[0]
        az_mult *= 2
        if az_mult > 64:
            break
    pts = out[0][:n_points]
    n = pts.shape[0]
    rec = np.zeros((n, 8), dtype=np.float32)
    rec[:, :3] = pts.astype(np.float32)
    rec[:, 3] = 1.0
    rec[:, 4] = 100.0
    rec[:, 5] = np.arange(n, dtype=np.uint32).view(np.float32)  # t: ns since scan start (bit pattern)
    rec[:, 6] = np.arange(n, dtype=np.uint32).view(np.float32)
    rec[:, 7] = np.linalg.norm(rec[:, :3].astype(np.float64), axis=1).astype(np.float32)
    return rec


def plane_scan(n_points: int, radius: float, sensor_height: float, rng, range_noise: float = 0.02) -> np.ndarray:
    """C1: surface samples of the ground plane within `radius` of a sensor `sensor_height` above it, body frame
    (identity attitude), with range noise along the ray."""
    r = radius * np.sqrt(rng.uniform(0.04, 1.0, n_points))
    a = rng.uniform(0, 2 * np.pi, n_points)
    p = np.stack([r * np.cos(a), r * np.sin(a), np.full(n_points, -sensor_height)], axis=1)
    rn = np.linalg.norm(p, axis=1)
    p = p * ((rn + rng.normal(0.0, range_noise, n_points)) / rn)[:, None]
    rec = np.zeros((n_points, 8), dtype=np.float32)
    rec[:, :3] = p.astype(np.float32)
    rec[:, 3] = 1.0
    rec[:, 7] = np.linalg.norm(rec[:, :3].astype(np.float64), axis=1).astype(np.float32)
    return rec


def expmap_se3(xi):
    """numpy SE(3) exponential, xi = [omega; v] (for generating perturbed start poses only)."""
    w, v = np.asarray(xi[:3], float), np.asarray(xi[3:], float)
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + W, v.copy()
    A, B, Cc = np.sin(th) / th, (1 - np.cos(th)) / th**2, (th - np.sin(th)) / th**3
    R = np.eye(3) + A * W + B * W @ W
    V = np.eye(3) + B * W + Cc * W @ W
    return R, V @ v


def perturbed_start(R_true, t_true, xi=(0.010, -0.008, 0.012, 0.05, -0.04, 0.03)):
    """T0 = T* . Exp(xi)  (SURVEY.md §8d)."""
    dR, dt = expmap_se3(np.asarray(xi))
    return R_true @ dR, R_true @ dt + t_true


def spread_queries(map_cloud: np.ndarray, n: int, rng, jitter: float = 0.05) -> np.ndarray:
    """The HBM-bound k-NN regime: queries drawn uniformly from the stored map points plus N(0, jitter^2),
    sorted by voxel Morton order (the order a spatially sorted scan would present them in)."""
    sel = rng.integers(0, map_cloud.shape[0], n)
    q = map_cloud[sel].astype(np.float64) + rng.normal(0.0, jitter, (n, 3))
    return q[morton_order(q)]


def _part1by2(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_order(p: np.ndarray, leaf: float = 1.0) -> np.ndarray:
    c = np.floor(np.asarray(p, dtype=np.float64) / leaf).astype(np.int64) + (1 << 20)
    key = _part1by2(c[:, 0]) | (_part1by2(c[:, 1]) << np.uint64(1)) | (_part1by2(c[:, 2]) << np.uint64(2))
    return np.argsort(key, kind="stable")
