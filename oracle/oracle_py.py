"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so (the CPU restatement of the
reference's ICP factor path).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under mimosa_b200/ does.  PARITY UNPINNED
(see oracle/icp_factor_ref.hpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "icp_factor_ref.hpp", "ivox_ref.hpp", "linalg_ref.hpp")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, env=env, capture_output=True)
    return LIB_PATH


class IcpConfig(C.Structure):  # same layout as mimosa_b200.capi.IcpConfig / RegistrationConfigRef
    _fields_ = [
        ("source_voxel_grid_filter_leaf_size", C.c_float),
        ("source_voxel_grid_min_dist_in_voxel", C.c_float),
        ("target_ivox_map_leaf_size", C.c_float),
        ("target_ivox_map_min_dist_in_voxel", C.c_float),
        ("num_corres_points", C.c_uint64),
        ("max_corres_distance", C.c_float),
        ("plane_validity_distance", C.c_float),
        ("lidar_point_noise_std_dev", C.c_float),
        ("use_huber", C.c_int32),
        ("huber_threshold", C.c_float),
        ("reg_4_dof", C.c_int32),
        ("project_on_degneneracy", C.c_int32),
        ("degen_thresh_rot", C.c_float),
        ("degen_thresh_trans", C.c_float),
    ]


class Linearization(C.Structure):
    _fields_ = [
        ("H", C.c_double * 36), ("g", C.c_double * 6), ("f", C.c_double), ("counts", C.c_int64 * 9),
        ("loc_trans_comp", C.c_double * 3), ("loc_rot_comp", C.c_double * 3),
        ("loc_trans_final", C.c_double * 3), ("loc_rot_final", C.c_double * 3),
        ("eigvec_trans", C.c_double * 9), ("eigvec_rot", C.c_double * 9),
        ("degen_rot", C.c_double * 3), ("degen_trans", C.c_double * 3),
        ("degen_eigvec_rot", C.c_double * 9), ("degen_eigvec_trans", C.c_double * 9),
        ("linearize_count", C.c_int32), ("n_searched", C.c_int32),
    ]


class IcpTrace(C.Structure):
    _fields_ = [
        ("H", C.c_double * 36), ("g", C.c_double * 6), ("f", C.c_double), ("delta", C.c_double * 6),
        ("R", C.c_double * 9), ("t", C.c_double * 3), ("counts", C.c_int64 * 9),
        ("n_searched", C.c_int32), ("solve_ok", C.c_int32),
        ("loc_trans_comp", C.c_double * 3), ("loc_rot_comp", C.c_double * 3),
    ]


_lib = None
_P, _SZ = C.c_void_p, C.c_size_t


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        sig = {
            "orc_map_create": (_P, [C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint64]),
            "orc_map_release": (None, [_P]),
            "orc_map_snapshot": (_P, [_P]),
            "orc_map_insert": (None, [_P, _P, _SZ, _SZ]),
            "orc_map_num_voxels": (_SZ, [_P]),
            "orc_map_num_points": (_SZ, [_P]),
            "orc_map_lru_counter": (C.c_uint64, [_P]),
            "orc_map_download": (None, [_P, _P, _P, _P, _P]),
            "orc_map_load_raw": (None, [_P, _P, _P, _P, _P, _SZ, C.c_uint64]),
            "orc_map_knn": (None, [_P, _P, _SZ, C.c_int, _P, _P, _P, C.c_int]),
            "orc_factor_create": (_P, [_P, _P, _SZ, _SZ, C.POINTER(IcpConfig)]),
            "orc_factor_release": (None, [_P]),
            "orc_factor_reset": (None, [_P]),
            "orc_factor_linearize": (None, [_P, _P, _P, _P, C.POINTER(Linearization), C.c_int]),
            "orc_factor_download_state": (None, [_P, _P, _P, _P, _P, _P, _P, _P]),
            "orc_icp_run": (C.c_double, [_P, _P, _P, C.c_int, C.c_double, _P, C.c_int]),
            "orc_downsample": (_SZ, [_P, _SZ, _SZ, C.c_float, _SZ, C.c_float, _P]),
            "orc_transform_f32": (None, [_P, _SZ, _SZ, _P, _P]),
            "orc_eigh3": (C.c_int, [_P, _P, _P]),
            "orc_se3_expmap": (None, [_P, _P, _P]),
            "orc_se3_retract": (None, [_P, _P, _P]),
            "orc_solve6": (C.c_int, [_P, C.c_double, _P, _P]),
            "orc_fast_floor": (C.c_int, [C.c_double]),
            "orc_max_threads": (C.c_int, []),
            "orc_sizeof": (_SZ, [C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        for which, struct in enumerate((IcpConfig, Linearization, IcpTrace)):
            assert L.orc_sizeof(which) == C.sizeof(struct), (struct.__name__, L.orc_sizeof(which), C.sizeof(struct))
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def config_to_c(cfg) -> IcpConfig:
    """Accepts a mimosa_b200.RegistrationConfig-like object (attribute access)."""
    return IcpConfig(
        cfg.source_voxel_grid_filter_leaf_size, cfg.source_voxel_grid_min_dist_in_voxel,
        cfg.target_ivox_map_leaf_size, cfg.target_ivox_map_min_dist_in_voxel, int(cfg.num_corres_points),
        cfg.max_corres_distance, cfg.plane_validity_distance, cfg.lidar_point_noise_std_dev, int(cfg.use_huber),
        cfg.huber_threshold, int(cfg.reg_4_dof), int(cfg.project_on_degneneracy), cfg.degen_thresh_rot,
        cfg.degen_thresh_trans)


class IVoxRef:
    def __init__(self, leaf, min_dist=0.1, cap=20, nbr_mode=7, lru_horizon=100, _h=None):
        self.L = lib()
        self.cap = cap
        self.h = _h if _h is not None else self.L.orc_map_create(leaf, min_dist, cap, nbr_mode, lru_horizon)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_map_release(self.h)
            self.h = None

    def insert(self, xyz):
        pts = np.ascontiguousarray(xyz, dtype=np.float32)
        self.L.orc_map_insert(self.h, _ptr(pts), pts.shape[0], pts.strides[0] if pts.shape[0] else 12)

    def snapshot(self):
        return IVoxRef(0, cap=self.cap, _h=self.L.orc_map_snapshot(self.h))

    def size(self):
        return (int(self.L.orc_map_num_voxels(self.h)), int(self.L.orc_map_num_points(self.h)),
                int(self.L.orc_map_lru_counter(self.h)))

    def download(self):
        nv = int(self.L.orc_map_num_voxels(self.h))
        coords = np.zeros((nv, 3), np.int32)
        counts = np.zeros(nv, np.int32)
        lru = np.zeros(nv, np.uint32)
        pts = np.zeros((nv, self.cap, 3), np.float32)
        self.L.orc_map_download(self.h, _ptr(coords), _ptr(counts), _ptr(lru), _ptr(pts))
        return coords, counts, lru, pts, int(self.L.orc_map_lru_counter(self.h))

    def load_raw(self, coords, counts, lru, pts, lru_counter=0):
        coords = np.ascontiguousarray(coords, np.int32)
        counts = np.ascontiguousarray(counts, np.int32)
        lru = None if lru is None else np.ascontiguousarray(lru, np.uint32)
        pts = np.ascontiguousarray(pts, np.float32)
        self.L.orc_map_load_raw(self.h, _ptr(coords), _ptr(counts), _ptr(lru), _ptr(pts), coords.shape[0], lru_counter)

    def knn_search(self, q, k, n_threads=1):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 3)
        nq = q.shape[0]
        idx = np.empty((nq, k), np.uint64)
        d2 = np.empty((nq, k), np.float64)
        ok = np.empty(nq, np.uint8)
        self.L.orc_map_knn(self.h, _ptr(q), nq, k, _ptr(idx), _ptr(d2), _ptr(ok), n_threads)
        return idx, d2, ok.astype(bool)


class IcpFactorRef:
    def __init__(self, target: IVoxRef, scan, cfg):
        self.L = lib()
        pts = np.ascontiguousarray(scan, dtype=np.float32)
        self.n = pts.shape[0]
        self.k = int(cfg.num_corres_points)
        c = config_to_c(cfg)
        self.target = target
        self.h = self.L.orc_factor_create(target.h, _ptr(pts), self.n, pts.strides[0] if self.n else 12, C.byref(c))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_factor_release(self.h)
            self.h = None

    def reset(self):
        self.L.orc_factor_reset(self.h)

    def linearize(self, R, t, gravity_unit=(0.0, 0.0, -1.0), n_threads=0) -> Linearization:
        """n_threads = 0: serial with the reference's four static chunks; > 0: that many OpenMP threads."""
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        t = np.ascontiguousarray(t, np.float64).reshape(3)
        g = np.ascontiguousarray(gravity_unit, np.float64).reshape(3)
        out = Linearization()
        self.L.orc_factor_linearize(self.h, _ptr(R), _ptr(t), _ptr(g), C.byref(out), n_threads)
        return out

    def download_state(self):
        n = self.n
        st = np.empty(n, np.uint8)
        arrs = [np.empty((n, 3), np.float64) for _ in range(5)]
        idx = np.empty((n, self.k), np.uint64)
        self.L.orc_factor_download_state(self.h, _ptr(st), *[_ptr(a) for a in arrs], _ptr(idx))
        return dict(status=st, p_da=arrs[0], mean=arrs[1], normal=arrs[2], loc_rot=arrs[3], loc_trans=arrs[4], knn_idx=idx)

    def icp_run(self, R, t, iters, lam=0.0, n_threads=0):
        R = np.array(R, np.float64).reshape(9).copy()
        t = np.array(t, np.float64).reshape(3).copy()
        trace = (IcpTrace * max(iters, 1))()
        secs = self.L.orc_icp_run(self.h, _ptr(R), _ptr(t), iters, lam, trace, n_threads)
        return R.reshape(3, 3), t, list(trace)[:iters], float(secs)


def downsample(xyz, leaf, cap, min_dist):
    pts = np.ascontiguousarray(xyz, np.float32)
    out = np.empty(pts.shape[0], np.uint32)
    n = lib().orc_downsample(_ptr(pts), pts.shape[0], pts.strides[0] if pts.shape[0] else 12, leaf, cap, min_dist, _ptr(out))
    return out[:n].copy()


def transform_f32(pts, poses, pose_index=None):
    """In-place float transform of the xyz columns of `pts` ((n, >=3) float32 rows): deskew / T_B_L / T_W_Be."""
    assert pts.dtype == np.float32 and pts.flags.c_contiguous
    poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
    pi = None if pose_index is None else np.ascontiguousarray(pose_index, np.uint32)
    lib().orc_transform_f32(_ptr(pts), pts.shape[0], pts.strides[0] if pts.shape[0] else 12, _ptr(pi), _ptr(poses))
    return pts


def eigh3(A):
    A = np.ascontiguousarray(A, np.float64).reshape(9)
    lam = np.empty(3)
    V = np.empty(9)
    ok = lib().orc_eigh3(_ptr(A), _ptr(lam), _ptr(V))
    return bool(ok), lam, V.reshape(3, 3)


def se3_expmap(xi):
    xi = np.ascontiguousarray(xi, np.float64)
    R = np.empty(9)
    t = np.empty(3)
    lib().orc_se3_expmap(_ptr(xi), _ptr(R), _ptr(t))
    return R.reshape(3, 3), t


def se3_retract(R, t, xi):
    R = np.array(R, np.float64).reshape(9).copy()
    t = np.array(t, np.float64).reshape(3).copy()
    xi = np.ascontiguousarray(xi, np.float64)
    lib().orc_se3_retract(_ptr(R), _ptr(t), _ptr(xi))
    return R.reshape(3, 3), t


def solve6(H, lam, rhs):
    H = np.ascontiguousarray(H, np.float64).reshape(36)
    rhs = np.ascontiguousarray(rhs, np.float64).reshape(6)
    x = np.zeros(6)
    ok = lib().orc_solve6(_ptr(H), lam, _ptr(rhs), _ptr(x))
    return bool(ok), x


def max_threads() -> int:
    return int(lib().orc_max_threads())
