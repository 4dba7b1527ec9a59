"""ctypes binding of libmimosa_b200.so (the C ABI declared in include/mimosa_b200.h).

This is plumbing for the Python tests / bench; C++ callers include the header directly.  There is no
fallback: if the library is missing, or the machine has no sm_100 GPU, loading / mb_init raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmimosa_b200.so")

MB_OK = 0
MB_ERR_INVALID_ARG = 1
MB_ERR_NO_DEVICE = 2
MB_ERR_CUDA = 3
MB_ERR_UNSUPPORTED = 4
MB_ERR_NCCL = 5
MB_ERR_CAPACITY = 6
MB_MAX_K = 8
MB_POINT_STRIDE = 32

STATUS_NAMES = [
    "Unprocessed",
    "InsufficientCorresPoints",
    "CorresMaxDist",
    "EigenSolverFail",
    "MinEigenValueLow",
    "Line",
    "CorresPlaneInvalid",
    "MaxError",
    "Valid",
]


class IcpConfig(C.Structure):
    """mb_icp_config == mimosa::lidar::RegistrationConfig (geometric_config.hpp:17-33)."""

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


class CloudLayout(C.Structure):
    _fields_ = [("point_step", C.c_uint32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_intensity", C.c_int32), ("intensity_type", C.c_int32), ("off_time", C.c_int32), ("time_type", C.c_int32),
                ("off_ring", C.c_int32), ("ring_type", C.c_int32), ("off_tag", C.c_int32), ("ring_filter", C.c_int32)]


class InputFilter(C.Structure):
    _fields_ = [("intensity_min", C.c_float), ("intensity_max", C.c_float), ("range_min", C.c_float), ("range_max", C.c_float),
                ("ns_max", C.c_float), ("point_skip_divisor", C.c_int32), ("ring_skip_divisor", C.c_int32),
                ("create_full_res_pointcloud", C.c_int32), ("z_offset", C.c_float), ("header_ts", C.c_double)]


class CloudOrder(C.Structure):
    """mb_cloud_order: the message re-orderings of lidar::Manager::prepareInput (manager.cpp:179-243)."""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("transpose_pointcloud", C.c_int32),
                ("organize_pointcloud_by_ring", C.c_int32)]


class Linearization(C.Structure):
    _fields_ = [
        ("H", C.c_double * 36),
        ("g", C.c_double * 6),
        ("f", C.c_double),
        ("counts", C.c_int64 * 9),
        ("loc_trans_comp", C.c_double * 3),
        ("loc_rot_comp", C.c_double * 3),
        ("loc_trans_final", C.c_double * 3),
        ("loc_rot_final", C.c_double * 3),
        ("eigvec_trans", C.c_double * 9),
        ("eigvec_rot", C.c_double * 9),
        ("degen_rot", C.c_double * 3),
        ("degen_trans", C.c_double * 3),
        ("degen_eigvec_rot", C.c_double * 9),
        ("degen_eigvec_trans", C.c_double * 9),
        ("linearize_count", C.c_int32),
        ("n_searched", C.c_int32),
    ]


class IcpTrace(C.Structure):
    _fields_ = [
        ("H", C.c_double * 36),
        ("g", C.c_double * 6),
        ("f", C.c_double),
        ("delta", C.c_double * 6),
        ("R", C.c_double * 9),
        ("t", C.c_double * 3),
        ("counts", C.c_int64 * 9),
        ("n_searched", C.c_int32),
        ("solve_ok", C.c_int32),
        ("loc_trans_comp", C.c_double * 3),
        ("loc_rot_comp", C.c_double * 3),
    ]


# name -> (restype, argtypes); every symbol include/mimosa_b200.h declares.
_P = C.c_void_p
_SZ = C.c_size_t
SIGNATURES = {
    "mb_init": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "mb_shutdown": (C.c_int, [_P]),
    "mb_last_error": (C.c_char_p, []),
    "mb_version": (C.c_int, []),
    "mb_sizeof": (C.c_size_t, [C.c_int]),
    "mb_sync": (C.c_int, [_P]),
    "mb_timer_begin": (C.c_int, [_P]),
    "mb_timer_end": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "mb_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "mb_set_resident_window": (C.c_int, [_P, C.c_uint]),
    "mb_flush_l2": (C.c_int, [_P, _SZ]),
    "mb_flush_l2_read": (C.c_int, [_P, _SZ]),
    "mb_host_register": (C.c_int, [_P, C.c_size_t]),
    "mb_host_unregister": (C.c_int, [_P]),
    "mb_comm_unique_id": (C.c_int, [_P]),
    "mb_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "mb_comm_ipc_handle": (C.c_int, [_P, _P]),
    "mb_comm_ipc_open": (C.c_int, [_P, _P]),
    "mb_comm_barrier": (C.c_int, [_P]),
    "mb_map_create": (C.c_int, [_P, C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint64, C.POINTER(_P)]),
    "mb_map_release": (C.c_int, [_P]),
    "mb_map_insert": (C.c_int, [_P, _P, _SZ, _SZ]),
    "mb_map_snapshot": (C.c_int, [_P, C.POINTER(_P)]),
    "mb_map_size": (C.c_int, [_P, C.POINTER(_SZ), C.POINTER(_SZ), C.POINTER(C.c_uint64)]),
    "mb_map_knn": (C.c_int, [_P, _P, _SZ, C.c_int, _P, _P, _P]),
    "mb_map_points": (C.c_int, [_P, _P, _SZ, _P]),
    "mb_map_download": (C.c_int, [_P, _P, _P, _P, _P]),
    "mb_map_upload": (C.c_int, [_P, _P, _P, _P, _P, _SZ, C.c_uint64]),
    "mb_map_knn_stage": (C.c_int, [_P, _P, _SZ, C.c_int]),
    "mb_map_knn_staged_run": (C.c_int, [_P]),
    "mb_map_knn_staged_run_prefix": (C.c_int, [_P, C.c_size_t]),
    "mb_map_knn_staged_fetch": (C.c_int, [_P, _P, _P, _P]),
    "mb_factor_create": (C.c_int, [_P, _P, _P, _SZ, _SZ, C.POINTER(IcpConfig), _SZ, _SZ, C.POINTER(_P)]),
    "mb_factor_release": (C.c_int, [_P]),
    "mb_factor_reset": (C.c_int, [_P]),
    "mb_factor_linearize": (C.c_int, [_P, _P, _P, _P, C.POINTER(Linearization)]),
    "mb_factor_download_state": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "mb_icp_run": (C.c_int, [_P, _P, _P, C.c_int, C.c_double, _P]),
    "mb_factor_set_flags": (C.c_int, [_P, C.c_uint32]),
    "mb_gn_step": (C.c_int, [_P, _P, C.c_double, _P, _P, _P, C.POINTER(C.c_int)]),
    "mb_downsample": (C.c_int, [_P, _P, _SZ, _SZ, C.c_float, _SZ, C.c_float, _P, C.POINTER(_SZ)]),
    "mb_scan_upload": (C.c_int, [_P, _P, _SZ, _SZ, C.POINTER(_P)]),
    "mb_scan_from_cloud": (C.c_int, [_P, _P, _SZ, C.POINTER(CloudLayout), C.POINTER(InputFilter), C.POINTER(_P), _P,
                                    C.POINTER(_SZ), _P, _P, C.POINTER(_SZ), C.POINTER(C.c_uint32)]),
    "mb_scan_from_cloud_ordered": (C.c_int, [_P, _P, _SZ, C.POINTER(CloudLayout), C.POINTER(InputFilter), C.POINTER(CloudOrder),
                                            C.POINTER(_P), _P, C.POINTER(_SZ), _P, _P, C.POINTER(_SZ), C.POINTER(C.c_uint32)]),
    "mb_scan_gather": (C.c_int, [_P, _P, _SZ, C.POINTER(_P)]),
    "mb_scan_release": (C.c_int, [_P]),
    "mb_scan_size": (C.c_int, [_P, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "mb_scan_download": (C.c_int, [_P, _P]),
    "mb_scan_deskew": (C.c_int, [_P, _P, _P, _SZ]),
    "mb_scan_transform": (C.c_int, [_P, _P, _P]),
    "mb_scan_downsample": (C.c_int, [_P, C.c_float, _SZ, C.c_float, C.POINTER(_P)]),
    "mb_factor_create_from_scan": (C.c_int, [_P, _P, _P, C.POINTER(IcpConfig), _SZ, _SZ, C.POINTER(_P)]),
    "mb_map_insert_scan": (C.c_int, [_P, _P, _P, _P]),
}


class MimosaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mimosa_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every symbol.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  mimosa_b200 has no CPU fallback."
        )
    try:
        # If torch is going to be used in this process (bench.py, tests) it must map ITS bundled libnccl.so.2
        # first; otherwise the system NCCL this library links against would shadow it by soname.
        import torch  # noqa: F401
    except Exception:  # pragma: no cover - torch is plumbing, not a requirement of the library
        pass
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    for which, struct in enumerate((IcpConfig, Linearization, IcpTrace)):
        if lib.mb_sizeof(which) != C.sizeof(struct):
            raise ImportError(f"ABI mismatch: {struct.__name__} is {C.sizeof(struct)} B here, {lib.mb_sizeof(which)} B in the library")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != MB_OK:
        raise MimosaError(code, load().mb_last_error().decode("utf-8", "replace"))
