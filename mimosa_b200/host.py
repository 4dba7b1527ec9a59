"""Python mirror of the reference's C++ interface for the LiDAR geometric-factor path, over the C ABI.

Names follow the reference so parity tests read like tests of mimosa itself:

* ``RegistrationConfig``     mimosa/include/mimosa/lidar/geometric_config.hpp:17-33
* ``IncrementalVoxelMap``    mimosa::lidar::IncrementalVoxelMapPCL, incremental_voxel_map.hpp:22-51
* ``ICPFactor``              mimosa::lidar::ICPFactor (unary), geometric_factor.hpp:25-563

The production host language is C++ (see mimosa_b200/host/ for the header-only C++ mirror); this module exists
because the test and benchmark harness of this repository is Python.  All compute happens in
libmimosa_b200.so on the GPU; numpy is used for argument marshalling only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import CloudLayout, CloudOrder, IcpConfig, IcpTrace, InputFilter, Linearization, check


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class RegistrationConfig:
    """Defaults are the struct defaults of geometric_config.hpp:17-33."""

    source_voxel_grid_filter_leaf_size: float = 0.5
    source_voxel_grid_min_dist_in_voxel: float = 0.1
    target_ivox_map_leaf_size: float = 0.5
    target_ivox_map_min_dist_in_voxel: float = 0.1
    num_corres_points: int = 5
    max_corres_distance: float = 2.24
    plane_validity_distance: float = 0.04
    lidar_point_noise_std_dev: float = 0.02
    use_huber: bool = True
    huber_threshold: float = 1.345
    reg_4_dof: bool = False
    project_on_degneneracy: bool = True
    degen_thresh_rot: float = 10.0
    degen_thresh_trans: float = 15.0

    def to_c(self) -> IcpConfig:
        return IcpConfig(
            self.source_voxel_grid_filter_leaf_size,
            self.source_voxel_grid_min_dist_in_voxel,
            self.target_ivox_map_leaf_size,
            self.target_ivox_map_min_dist_in_voxel,
            int(self.num_corres_points),
            self.max_corres_distance,
            self.plane_validity_distance,
            self.lidar_point_noise_std_dev,
            int(self.use_huber),
            self.huber_threshold,
            int(self.reg_4_dof),
            int(self.project_on_degneneracy),
            self.degen_thresh_rot,
            self.degen_thresh_trans,
        )


def hornbill_config() -> RegistrationConfig:
    """mimosa/config/hornbill/params.yaml:86-102 (also magpie / lapwing / parrot / euroc)."""
    return RegistrationConfig(
        source_voxel_grid_filter_leaf_size=1.0,
        source_voxel_grid_min_dist_in_voxel=0.2,
        target_ivox_map_leaf_size=1.0,
        target_ivox_map_min_dist_in_voxel=0.2,
        num_corres_points=5,
        max_corres_distance=1.0,
        plane_validity_distance=0.07,
        lidar_point_noise_std_dev=0.07,
        use_huber=True,
        huber_threshold=1.345,
        reg_4_dof=False,
        project_on_degneneracy=False,
        degen_thresh_rot=0.0,
        degen_thresh_trans=40.0,
    )


HORNBILL_MAP = dict(leaf=1.0, min_dist=0.2, cap=20, nbr_mode=19, lru_horizon=1000)


class Context:
    """One per process and GPU (mb_init)."""

    def __init__(self, device: int = 0):
        self.lib = capi.load()
        h = C.c_void_p()
        check(self.lib.mb_init(device, C.byref(h)))
        self.h = h
        self.device = device
        self.rank, self.world = 0, 1

    def close(self):
        if self.h:
            self.lib.mb_shutdown(self.h)
            self.h = None

    def sync(self):
        check(self.lib.mb_sync(self.h))

    def timer_begin(self):
        check(self.lib.mb_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        check(self.lib.mb_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def set_resident_window(self, microseconds: int):
        """How long the linearisation kernel stays resident after a host-facing call (0: every call launches)."""
        check(self.lib.mb_set_resident_window(self.h, int(microseconds)))

    def launch_count(self) -> int:
        v = C.c_uint64()
        check(self.lib.mb_launch_count(self.h, C.byref(v)))
        return int(v.value)

    def flush_l2(self, nbytes: int = 256 << 20, by_reading: bool = False):
        check((self.lib.mb_flush_l2_read if by_reading else self.lib.mb_flush_l2)(self.h, nbytes))

    def host_register(self, a: np.ndarray):
        """Page-lock a C-contiguous numpy array in place (mb_host_register); scans taken from it skip the CPU staging pass."""
        assert a.flags["C_CONTIGUOUS"]
        check(self.lib.mb_host_register(_ptr(a), a.nbytes))

    def host_unregister(self, a: np.ndarray):
        check(self.lib.mb_host_unregister(_ptr(a)))

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(self.lib.mb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        check(self.lib.mb_comm_init(self.h, rank, world, buf))
        self.rank, self.world = rank, world

    def comm_ipc_handle(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's packet mailbox (peer-memory exchange, single node)."""
        buf = C.create_string_buffer(64)
        check(self.lib.mb_comm_ipc_handle(self.h, buf))
        return buf.raw

    def comm_ipc_open(self, handles: bytes):
        """`handles` = the world x 64 bytes of every rank's mb_comm_ipc_handle, in rank order."""
        buf = C.create_string_buffer(handles, len(handles))
        check(self.lib.mb_comm_ipc_open(self.h, buf))

    def comm_barrier(self):
        """Device-side barrier of all ranks on the context's stream (peer mailboxes)."""
        check(self.lib.mb_comm_barrier(self.h))

    def downsample(self, xyz: np.ndarray, leaf: float, cap: int, min_dist: float) -> np.ndarray:
        """Geometric::downsample (geometric.cpp:55-126): indices of the kept points, reference order."""
        pts = np.ascontiguousarray(xyz, dtype=np.float32)
        n, stride = pts.shape[0], (pts.strides[0] if pts.shape[0] else 12)
        out = np.empty(max(n, 1), dtype=np.uint32)
        n_out = C.c_size_t()
        check(self.lib.mb_downsample(self.h, _ptr(pts), n, stride, leaf, cap, min_dist, _ptr(out), C.byref(n_out)))
        return out[: n_out.value].copy()


class Scan:
    """Device-resident scan: (n, 8) float32 lidar::Point rows (or any (n, >=3) float32 rows)."""

    def __init__(self, ctx: Context, pts: np.ndarray = None, _handle=None, _cols=None):
        self.ctx, self.lib = ctx, ctx.lib
        if _handle is None:
            pts = np.ascontiguousarray(pts, dtype=np.float32)
            _cols = pts.shape[1]
            h = C.c_void_p()
            check(self.lib.mb_scan_upload(ctx.h, _ptr(pts), pts.shape[0], pts.strides[0] if pts.shape[0] else 4 * _cols, C.byref(h)))
            _handle = h
        self.h, self.cols = _handle, _cols

    @staticmethod
    def from_cloud(ctx: Context, data: np.ndarray, layout: CloudLayout, filt: InputFilter, width: int = 0, height: int = 0,
                   transpose_pointcloud: bool = False, organize_pointcloud_by_ring: bool = False):
        """lidar::Manager::prepareInput (manager.cpp:149-383).  `data`: (n, point_step) uint8 PointCloud2 payload;
        width x height + the two ManagerConfig flags select the message re-orderings of manager.cpp:179-243.
        Returns (points_full Scan, geometric_idx, pose_index, unique_ns, last_point_ns)."""
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, layout.point_step)
        n = data.shape[0]
        if transpose_pointcloud or organize_pointcloud_by_ring:
            order = CloudOrder(width, height, int(transpose_pointcloud), int(organize_pointcloud_by_ring))
            geo, pi, uns = (np.empty(max(n, 1), np.uint32) for _ in range(3))
            h, n_geo, n_un, last = C.c_void_p(), C.c_size_t(), C.c_size_t(), C.c_uint32()
            check(ctx.lib.mb_scan_from_cloud_ordered(ctx.h, _ptr(data), n, C.byref(layout), C.byref(filt), C.byref(order), C.byref(h),
                                                     _ptr(geo), C.byref(n_geo), _ptr(pi), _ptr(uns), C.byref(n_un), C.byref(last)))
            sc = Scan(ctx, _handle=h, _cols=8)
            return sc, geo[: n_geo.value].copy(), pi[: sc.n].copy(), uns[: n_un.value].copy(), int(last.value)
        geo = np.empty(max(n, 1), np.uint32)
        pi = np.empty(max(n, 1), np.uint32)
        uns = np.empty(max(n, 1), np.uint32)
        h, n_geo, n_un, last = C.c_void_p(), C.c_size_t(), C.c_size_t(), C.c_uint32()
        check(ctx.lib.mb_scan_from_cloud(ctx.h, _ptr(data), n, C.byref(layout), C.byref(filt), C.byref(h), _ptr(geo),
                                         C.byref(n_geo), _ptr(pi), _ptr(uns), C.byref(n_un), C.byref(last)))
        sc = Scan(ctx, _handle=h, _cols=8)
        return sc, geo[: n_geo.value].copy(), pi[: sc.n].copy(), uns[: n_un.value].copy(), int(last.value)

    def gather(self, idx: np.ndarray) -> "Scan":
        """points[idx] as a new scan (Geometric::preprocess's idxs, geometric.cpp:151-158)."""
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        h = C.c_void_p()
        check(self.lib.mb_scan_gather(self.h, _ptr(idx), idx.size, C.byref(h)))
        return Scan(self.ctx, _handle=h, _cols=self.cols)

    def release(self):
        if self.h:
            self.lib.mb_scan_release(self.h)
            self.h = None

    @property
    def n(self) -> int:
        n, st = C.c_size_t(), C.c_size_t()
        check(self.lib.mb_scan_size(self.h, C.byref(n), C.byref(st)))
        return int(n.value)

    def download(self) -> np.ndarray:
        out = np.empty((self.n, self.cols), dtype=np.float32)
        check(self.lib.mb_scan_download(self.h, _ptr(out)))
        return out

    def deskew(self, pose_index: np.ndarray, poses: np.ndarray):
        """lidar::Manager::deskewPoints (manager.cpp:494-509): poses (n_poses, 12) float32 = [R row-major | t]."""
        pose_index = np.ascontiguousarray(pose_index, dtype=np.uint32)
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 12)
        check(self.lib.mb_scan_deskew(self.h, _ptr(pose_index), _ptr(poses), poses.shape[0]))

    def transform(self, R, t):
        """Geometric::preprocess T_B_L (geometric.cpp:153-160), float arithmetic."""
        R = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
        t = np.ascontiguousarray(t, dtype=np.float32).reshape(3)
        check(self.lib.mb_scan_transform(self.h, _ptr(R), _ptr(t)))

    def downsample(self, leaf: float, cap: int, min_dist: float) -> "Scan":
        h = C.c_void_p()
        check(self.lib.mb_scan_downsample(self.h, leaf, cap, min_dist, C.byref(h)))
        return Scan(self.ctx, _handle=h, _cols=self.cols)


class IncrementalVoxelMap:
    """IncrementalVoxelMapPCL: insert / knn_search / getCloud / copy (snapshot)."""

    def __init__(self, ctx: Context, leaf: float, min_dist: float = 0.1, cap: int = 20, nbr_mode: int = 7,
                 lru_horizon: int = 100, _handle=None):
        self.ctx, self.lib = ctx, ctx.lib
        self.leaf, self.min_dist, self.cap, self.nbr_mode, self.lru_horizon = leaf, min_dist, cap, nbr_mode, lru_horizon
        if _handle is None:
            h = C.c_void_p()
            check(self.lib.mb_map_create(ctx.h, leaf, min_dist, cap, nbr_mode, lru_horizon, C.byref(h)))
            _handle = h
        self.h = _handle

    def release(self):
        if self.h:
            self.lib.mb_map_release(self.h)
            self.h = None

    def insert(self, xyz: np.ndarray):
        """xyz: (n, >=3) float32 rows; only the first three columns are read (stride honoured)."""
        pts = np.ascontiguousarray(xyz, dtype=np.float32)
        check(self.lib.mb_map_insert(self.h, _ptr(pts), pts.shape[0], pts.strides[0] if pts.shape[0] else 12))

    def insert_scan(self, scan: "Scan", R, t):
        """Geometric::updateMap (geometric.cpp:483-495): float world transform of the device scan, then insert."""
        R = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
        t = np.ascontiguousarray(t, dtype=np.float32).reshape(3)
        check(self.lib.mb_map_insert_scan(self.h, scan.h, _ptr(R), _ptr(t)))

    def snapshot(self) -> "IncrementalVoxelMap":
        """The deep copy mimosa makes before every insert (geometric.cpp:494)."""
        h = C.c_void_p()
        check(self.lib.mb_map_snapshot(self.h, C.byref(h)))
        return IncrementalVoxelMap(self.ctx, self.leaf, self.min_dist, self.cap, self.nbr_mode, self.lru_horizon, h)

    def size(self):
        nv, npts, lru = C.c_size_t(), C.c_size_t(), C.c_uint64()
        check(self.lib.mb_map_size(self.h, C.byref(nv), C.byref(npts), C.byref(lru)))
        return int(nv.value), int(npts.value), int(lru.value)

    def knn_search(self, q: np.ndarray, k: int):
        """Returns (idx (nq,k) uint64, d2 (nq,k) float64, ok (nq,) bool) — ok == (found == k)."""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 3)
        nq = q.shape[0]
        idx = np.empty((nq, k), dtype=np.uint64)
        d2 = np.empty((nq, k), dtype=np.float64)
        ok = np.empty(nq, dtype=np.uint8)
        check(self.lib.mb_map_knn(self.h, _ptr(q), nq, k, _ptr(idx), _ptr(d2), _ptr(ok)))
        return idx, d2, ok.astype(bool)

    def knn_stage(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 3)
        self._staged = (q.shape[0], k)
        check(self.lib.mb_map_knn_stage(self.h, _ptr(q), q.shape[0], k))

    def knn_staged_run(self, prefix: int = 0):
        if prefix:
            check(self.lib.mb_map_knn_staged_run_prefix(self.h, prefix))
        else:
            check(self.lib.mb_map_knn_staged_run(self.h))

    def knn_staged_fetch(self):
        nq, k = self._staged
        idx = np.empty((nq, k), dtype=np.uint64)
        d2 = np.empty((nq, k), dtype=np.float64)
        ok = np.empty(nq, dtype=np.uint8)
        check(self.lib.mb_map_knn_staged_fetch(self.h, _ptr(idx), _ptr(d2), _ptr(ok)))
        return idx, d2, ok.astype(bool)

    def points(self, idx: np.ndarray) -> np.ndarray:
        """iVox::point(i) for global indices (voxel_id << 32) | point_id."""
        idx = np.ascontiguousarray(idx, dtype=np.uint64).ravel()
        out = np.empty((idx.size, 3), dtype=np.float64)
        check(self.lib.mb_map_points(self.h, _ptr(idx), idx.size, _ptr(out)))
        return out

    def download(self):
        """(coords (nv,3) i32, counts (nv,) i32, lru (nv,) u32, pts (nv,cap,3) f32, lru_counter)."""
        nv, _, lru_counter = self.size()
        coords = np.zeros((nv, 3), dtype=np.int32)
        counts = np.zeros(nv, dtype=np.int32)
        lru = np.zeros(nv, dtype=np.uint32)
        pts = np.zeros((nv, self.cap, 3), dtype=np.float32)
        check(self.lib.mb_map_download(self.h, _ptr(coords), _ptr(counts), _ptr(lru), _ptr(pts)))
        return coords, counts, lru, pts, lru_counter

    def upload(self, coords, counts, lru, pts, lru_counter: int = 0):
        coords = np.ascontiguousarray(coords, dtype=np.int32)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        lru = None if lru is None else np.ascontiguousarray(lru, dtype=np.uint32)
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        check(self.lib.mb_map_upload(self.h, _ptr(coords), _ptr(counts), _ptr(lru), _ptr(pts), coords.shape[0], lru_counter))

    def get_cloud(self) -> np.ndarray:
        """IncrementalVoxelMapPCL::getCloud: all stored points, voxel order then in-voxel order."""
        _, counts, _, pts, _ = self.download()
        keep = np.arange(self.cap)[None, :] < counts[:, None]
        return pts[keep]


class ICPFactor:
    """Unary scan-to-map ICP factor.  `scan` is (n, >=3) float32 (xyz first), e.g. (n, 8) lidar::Point rows."""

    def __init__(self, ctx: Context, target: IncrementalVoxelMap, scan: np.ndarray, config: RegistrationConfig,
                 shard=None):
        self.ctx, self.lib, self.target, self.config = ctx, ctx.lib, target, config
        self.k = int(config.num_corres_points)
        h = C.c_void_p()
        cfg = config.to_c()
        if isinstance(scan, Scan):  # device-resident scan
            n = scan.n
            begin, end = (0, n) if shard is None else shard
            check(self.lib.mb_factor_create_from_scan(ctx.h, target.h, scan.h, C.byref(cfg), begin, end, C.byref(h)))
        else:
            pts = np.ascontiguousarray(scan, dtype=np.float32)
            n = pts.shape[0]
            begin, end = (0, n) if shard is None else shard
            check(self.lib.mb_factor_create(ctx.h, target.h, _ptr(pts), n, pts.strides[0] if n else 12, C.byref(cfg),
                                            begin, end, C.byref(h)))
        self.n_total, self.begin, self.end = n, begin, end
        self.h = h
        self.last = None

    @property
    def n(self):
        return self.end - self.begin

    def release(self):
        if self.h:
            self.lib.mb_factor_release(self.h)
            self.h = None

    def reset(self):
        check(self.lib.mb_factor_reset(self.h))

    def set_flags(self, forced_search: bool = False, cuda_graph: bool = False):
        check(self.lib.mb_factor_set_flags(self.h, (1 if forced_search else 0) | (2 if cuda_graph else 0)))

    def linearize(self, R, t, gravity_unit=(0.0, 0.0, -1.0)) -> Linearization:
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(3)
        g = np.ascontiguousarray(gravity_unit, dtype=np.float64).reshape(3)
        out = Linearization()
        check(self.lib.mb_factor_linearize(self.h, _ptr(R), _ptr(t), _ptr(g), C.byref(out)))
        self.last = out
        return out

    def download_state(self, knn: bool = True):
        n = self.n
        st = np.empty(n, dtype=np.uint8)
        arrs = [np.empty((n, 3), dtype=np.float64) for _ in range(5)]
        idx = np.empty((n, self.k), dtype=np.uint64) if knn else None
        check(self.lib.mb_factor_download_state(self.h, _ptr(st), *[_ptr(a) for a in arrs], _ptr(idx)))
        return dict(status=st, p_da=arrs[0], mean=arrs[1], normal=arrs[2], loc_rot=arrs[3], loc_trans=arrs[4], knn_idx=idx)

    # accessors named after geometric_factor.hpp:48-72
    def get_statuses(self):
        return self.download_state(knn=False)["status"]

    def get_corres_means_target(self):
        return self.download_state(knn=False)["mean"]

    def get_corres_normals_target(self):
        return self.download_state(knn=False)["normal"]

    def get_localizabilities(self):
        L = self.last
        a = lambda x, shape=None: np.array(x, dtype=np.float64).reshape(shape) if shape else np.array(x, dtype=np.float64)
        return (a(L.loc_trans_comp), a(L.loc_rot_comp), a(L.loc_trans_final), a(L.loc_rot_final),
                a(L.eigvec_trans, (3, 3)), a(L.eigvec_rot, (3, 3)))

    def get_linearize_count(self):
        return 0 if self.last is None else int(self.last.linearize_count)

    def icp_run(self, R, t, iters: int, lam: float = 0.0, want_trace: bool = True):
        """Device-resident Gauss-Newton loop; returns (R, t, trace list)."""
        R = np.array(R, dtype=np.float64).reshape(9).copy()
        t = np.array(t, dtype=np.float64).reshape(3).copy()
        trace = (IcpTrace * iters)() if (want_trace and iters) else None
        check(self.lib.mb_icp_run(self.h, _ptr(R), _ptr(t), iters, lam, trace))
        return R.reshape(3, 3), t, (list(trace) if trace is not None else [])


def gn_step(L: Linearization, R, t, lam: float = 0.0):
    """Host-side Gauss-Newton step (mb_gn_step): returns (R, t, delta, ok)."""
    lib = capi.load()
    R = np.array(R, dtype=np.float64).reshape(9).copy()
    t = np.array(t, dtype=np.float64).reshape(3).copy()
    delta = np.zeros(6)
    ok = C.c_int()
    check(lib.mb_gn_step(C.byref(L, Linearization.H.offset), C.byref(L, Linearization.g.offset), lam, _ptr(R), _ptr(t),
                         _ptr(delta), C.byref(ok)))
    return R.reshape(3, 3), t, delta, bool(ok.value)


def degeneracy_flags_from(loc_rot_comp, loc_trans_comp, eigvec_rot, eigvec_trans, degen_thresh_rot, degen_thresh_trans):
    """Geometric::getFactors' consumer of the localizabilities (geometric.cpp:218-228): six booleans "component
    localizability below its (float) threshold", rotations first, and blkdiag(V_rot, V_trans)."""
    ev = np.zeros((6, 6))
    ev[:3, :3] = np.asarray(eigvec_rot, np.float64).reshape(3, 3)
    ev[3:, 3:] = np.asarray(eigvec_trans, np.float64).reshape(3, 3)
    d = np.zeros(6, dtype=bool)
    d[:3] = np.asarray(loc_rot_comp, np.float64) < float(np.float32(degen_thresh_rot))
    d[3:] = np.asarray(loc_trans_comp, np.float64) < float(np.float32(degen_thresh_trans))
    return d, ev


def degeneracy_flags(L: Linearization, config: RegistrationConfig):
    """The same from a linearisation and the factor's configuration; returns (eigenvectors_block_matrix, degen_directions)."""
    d, ev = degeneracy_flags_from(L.loc_rot_comp, L.loc_trans_comp, L.eigvec_rot, L.eigvec_trans, config.degen_thresh_rot,
                                  config.degen_thresh_trans)
    return ev, d.astype(np.float64)


class KeyframeGate:
    """The keyframe rule of Geometric::updateMap (mimosa/src/lidar/geometric.cpp:437-478), as in the C++ mirror
    (mimosa_b200/host/mimosa_b200.hpp::KeyframeGate): nearest map pose by (float) translation distance, first minimum;
    update when that distance exceeds map_keyframe_trans_thresh or the largest |yaw|, |pitch|, |roll| of
    R_B_L^-1 (R_kf^-1 R_now) R_B_L exceeds DEG2RAD(map_keyframe_rot_thresh_deg); the first
    initial_clouds_to_force_map_update calls always update.  Host logic; no compute."""

    def __init__(self, trans_thresh, rot_thresh_deg, initial_clouds_to_force_map_update, R_B_L=None):
        self.trans_thresh = np.float32(trans_thresh)
        self.rot_thresh_deg = np.float32(rot_thresh_deg)
        self.forced_left = int(initial_clouds_to_force_map_update)
        self.R_B_L = np.eye(3) if R_B_L is None else np.asarray(R_B_L, np.float64).reshape(3, 3)
        self.R = np.zeros((0, 3, 3))
        self.t = np.zeros((0, 3))

    @staticmethod
    def _rq_abs_max(A):  # gtsam::RQ -> Rot3::ypr(), only max |angle| is consumed
        x = -np.arctan2(-A[2, 1], A[2, 2])
        cx, sx = np.cos(-x), np.sin(-x)
        B = A @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        y = -np.arctan2(B[2, 0], B[2, 2])
        cy, sy = np.cos(-y), np.sin(-y)
        Cm = B @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        z = -np.arctan2(-Cm[1, 0], Cm[1, 1])
        return max(abs(x), abs(y), abs(z))

    def should_update(self, R, t) -> bool:
        R, t = np.asarray(R, np.float64).reshape(3, 3), np.asarray(t, np.float64).reshape(3)
        update = True
        if self.t.shape[0]:
            d = np.sqrt(((self.t - t) ** 2).sum(1)).astype(np.float32)
            i = int(np.argmin(d))  # first minimum, like the reference's strict '<'
            rot_diff = self.R_B_L.T @ (self.R[i].T @ R) @ self.R_B_L
            if d[i] > self.trans_thresh:
                update = True
            elif self._rq_abs_max(rot_diff) > float(self.rot_thresh_deg) * 0.017453293:
                update = True
            else:
                update = False
        if self.forced_left > 0:
            update = True
            self.forced_left -= 1
        return update

    def add_keyframe(self, R, t):
        self.R = np.concatenate([self.R, np.asarray(R, np.float64).reshape(1, 3, 3)])
        self.t = np.concatenate([self.t, np.asarray(t, np.float64).reshape(1, 3)])


def shard_range(n: int, rank: int, world: int):
    """Contiguous block of the scan owned by `rank` (ceil(n/world) points per rank)."""
    per = (n + world - 1) // world
    b = min(n, rank * per)
    return b, min(n, b + per)
