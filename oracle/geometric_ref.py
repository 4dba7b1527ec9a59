"""ORACLE (test infrastructure only — never imported by the product): CPU restatement of the host-side rules of
mimosa::lidar::Geometric that surround the ICP factor.

  KeyframeGateRef   Geometric::updateMap's keyframe rule, mimosa/src/lidar/geometric.cpp:437-478
  degeneracy_info   the degeneracy flags / eigenvector block of Geometric::getFactors, geometric.cpp:208-228

PARITY UNPINNED for the Euler angles: `rot_diff.ypr()` is gtsam::Rot3::ypr() of the GTSAM 4.2 fork (ntnu-arl/gtsam,
branch feature/imu_factor_with_gravity; not under /root/reference).  Its published algorithm (gtsam/geometry/Rot3.cpp,
RQ(): x = -atan2(-A21, A22), B = A Rx(-x)... ; ypr = (z, y, x)) is restated in rq_xyz below and checked against
scipy's 'ZYX' Euler angles in tests/test_geometric_host.py.  Everything else follows lines that are in the tree."""
import numpy as np

DEG2RAD_PCL = 0.017453293  # PCL's DEG2RAD macro as used at geometric.cpp:467


def rot_x(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def rot_y(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rq_xyz(A):
    """gtsam::RQ: angles (x, y, z) such that A = Rz(z) Ry(y) Rx(x)."""
    A = np.asarray(A, np.float64).reshape(3, 3)
    x = -np.arctan2(-A[2, 1], A[2, 2])
    B = A @ rot_x(-x)
    y = -np.arctan2(B[2, 0], B[2, 2])
    C = B @ rot_y(-y)
    z = -np.arctan2(-C[1, 0], C[1, 1])
    return np.array([x, y, z])


class KeyframeGateRef:
    def __init__(self, trans_thresh, rot_thresh_deg, initial_clouds_to_force_map_update, R_B_L=np.eye(3)):
        self.trans_thresh = np.float32(trans_thresh)      # float map_keyframe_trans_thresh, geometric_config.hpp:50
        self.rot_thresh_deg = np.float32(rot_thresh_deg)  # float map_keyframe_rot_thresh_deg, :51
        self.forced_left = int(initial_clouds_to_force_map_update)  # the function-static counter, geometric.cpp:474
        self.R_B_L = np.asarray(R_B_L, np.float64).reshape(3, 3)
        self.map_poses = []  # (R, t)

    def should_update(self, R, t):
        R, t = np.asarray(R, np.float64).reshape(3, 3), np.asarray(t, np.float64).reshape(3)
        update = True
        if self.map_poses:  # :439
            min_diff, min_idx = np.float32(np.finfo(np.float32).max), 0
            for i, (_, tk) in enumerate(self.map_poses):  # :444-451, float distance, strict '<': first minimum wins
                d = np.float32(np.linalg.norm(tk - t))
                if d < min_diff:
                    min_diff, min_idx = d, i
            Rk = self.map_poses[min_idx][0]
            rot_diff = self.R_B_L.T @ (Rk.T @ R) @ self.R_B_L  # :453-455, Rot3::between(a, b) = a^-1 b
            ypr_abs_max = np.abs(rq_xyz(rot_diff)).max()        # :457
            if min_diff > self.trans_thresh:                     # :465
                update = True
            elif ypr_abs_max > float(self.rot_thresh_deg) * DEG2RAD_PCL:  # :467 (float * double)
                update = True
            else:
                update = False
        if self.forced_left > 0:  # :474-478
            update = True
            self.forced_left -= 1
        return update

    def add_keyframe(self, R, t):  # :499
        self.map_poses.append((np.asarray(R, np.float64).reshape(3, 3).copy(), np.asarray(t, np.float64).reshape(3).copy()))


def degeneracy_info(loc_rot_comp, loc_trans_comp, eigvec_rot, eigvec_trans, degen_thresh_rot, degen_thresh_trans):
    """geometric.cpp:218-228: blkdiag(V_rot, V_trans) and the six 'component localizability < threshold' flags (the
    thresholds are floats, geometric_config.hpp:31-32, promoted in the comparison)."""
    M = np.zeros((6, 6))
    M[:3, :3] = np.asarray(eigvec_rot, np.float64).reshape(3, 3)
    M[3:, 3:] = np.asarray(eigvec_trans, np.float64).reshape(3, 3)
    d = np.zeros(6)
    d[:3] = np.asarray(loc_rot_comp, np.float64) < float(np.float32(degen_thresh_rot))
    d[3:] = np.asarray(loc_trans_comp, np.float64) < float(np.float32(degen_thresh_trans))
    return M, d
