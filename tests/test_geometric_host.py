"""Host-side rules of lidar::Geometric around the factor: the keyframe rule of updateMap (geometric.cpp:437-478) and the
degeneracy flags of getFactors (geometric.cpp:208-228).  Three statements are compared: the oracle
(oracle/geometric_ref.py), the C++ mirror a maintainer compiles against (mimosa_b200/host/mimosa_b200.hpp, run through
tests/cpp/geometric_host_check.cpp) and the Python mirror bench.py's streaming config uses (mimosa_b200/host.py)."""
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import geometric_ref as gref
from mimosa_b200 import host as mbh

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def checker():
    src = os.path.join(HERE, "cpp", "geometric_host_check.cpp")
    out = os.path.join(HERE, "cpp", "geometric_host_check")
    hdr = os.path.join(ROOT, "mimosa_b200", "host", "mimosa_b200.hpp")
    lib = os.path.join(ROOT, "mimosa_b200", "lib")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", src, "-L", lib, "-lmimosa_b200", f"-Wl,-rpath,{lib}", "-o", out], check=True)
    return out


def test_rq_angles_are_gtsam_ypr():
    """RQ's (x, y, z) are the intrinsic z-y'-x'' (yaw, pitch, roll) angles: R = Rz(z) Ry(y) Rx(x)."""
    rng = np.random.default_rng(0)
    for _ in range(200):
        ypr = rng.uniform([-3.1, -1.5, -3.1], [3.1, 1.5, 3.1])
        R = Rotation.from_euler("ZYX", ypr).as_matrix()
        x, y, z = gref.rq_xyz(R)
        assert np.allclose([z, y, x], ypr, atol=1e-12)
        assert np.allclose(Rotation.from_euler("ZYX", [z, y, x]).as_matrix(), R, atol=1e-12)


def pose_walk(rng, n, step, rot_step):
    R, t, out = np.eye(3), np.zeros(3), []
    for _ in range(n):
        t = t + rng.normal(0, step, 3)
        R = R @ Rotation.from_rotvec(rng.normal(0, rot_step, 3)).as_matrix()
        out.append((R.copy(), t.copy()))
    return out


@pytest.mark.parametrize("seed,tt,rt,forced,step,rot_step", [
    (1, 1.0, 10.0, 10, 0.3, 0.02),   # hornbill: translation decides
    (2, 2.0, 30.0, 0, 0.05, 0.15),   # enwide thresholds, rotation decides, nothing forced
    (3, 0.1, 10.0, 3, 0.02, 0.01),   # struct defaults, revisits old keyframes
    (4, 1.0, 10.0, 1, 0.0, 0.0),     # standing still: only the first cloud updates
])
def test_keyframe_gate_three_ways(checker, seed, tt, rt, forced, step, rot_step):
    rng = np.random.default_rng(seed)
    R_B_L = Rotation.from_euler("ZYX", [0.3, -0.2, 3.0]).as_matrix() if seed % 2 else np.eye(3)
    poses = pose_walk(rng, 120, step, rot_step)
    if seed == 3:  # walk back over the same ground: the nearest keyframe is not the last one
        poses = poses + poses[::-1]
    ref = gref.KeyframeGateRef(tt, rt, forced, R_B_L)
    py = mbh.KeyframeGate(tt, rt, forced, R_B_L)
    want, got_py = [], []
    for R, t in poses:
        u = ref.should_update(R, t)
        if u:
            ref.add_keyframe(R, t)
        want.append(int(u))
        v = py.should_update(R, t)
        if v:
            py.add_keyframe(R, t)
        got_py.append(int(v))
    txt = [f"{tt} {rt} {forced}", " ".join(repr(float(v)) for v in R_B_L.ravel()), str(len(poses))]
    txt += [" ".join(repr(float(v)) for v in np.concatenate([R.ravel(), t])) for R, t in poses]
    txt += ["0"]
    out = subprocess.run([checker], input="\n".join(txt) + "\n", capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == want
    assert got_py == want
    assert 0 < sum(want) < len(want) or seed == 4
    if seed == 4:
        assert want == [1] + [0] * (len(want) - 1)


def test_degeneracy_flags_three_ways(checker):
    rng = np.random.default_rng(5)
    rows = []
    for _ in range(50):
        lr, lt = rng.uniform(0, 60, 3), rng.uniform(0, 60, 3)
        tr, tt = rng.choice([0.0, 10.0, 15.0, 40.0]), rng.choice([0.0, 15.0, 40.0])
        if rng.uniform() < 0.3:  # exactly on the threshold: strict '<' must say "not degenerate"
            lr[0], lt[1] = tr, tt
        rows.append((lr, lt, tr, tt))
    txt = ["1 10 0", "1 0 0 0 1 0 0 0 1", "0", str(len(rows))] + [" ".join(repr(float(v)) for v in [*lr, *lt, tr, tt]) for lr, lt, tr, tt in rows]
    out = subprocess.run([checker], input="\n".join(txt) + "\n", capture_output=True, text=True, check=True).stdout.strip().split("\n")
    Vr, Vt = np.arange(1, 10.0).reshape(3, 3), np.arange(11, 20.0).reshape(3, 3)
    for line, (lr, lt, tr, tt) in zip(out, rows):
        M, d = gref.degeneracy_info(lr, lt, Vr, Vt, tr, tt)
        vals = line.split()
        assert [int(v) for v in vals[:6]] == [int(v) for v in d]
        assert [float(v) for v in vals[6:]] == [M[0, 0], M[3, 3], M[0, 3]]
        flags, Mp = mbh.degeneracy_flags_from(lr, lt, Vr, Vt, tr, tt)
        assert list(flags) == [bool(v) for v in d] and np.array_equal(Mp, M)
