"""The header the CUDA kernels take their arithmetic from (mimosa_b200/csrc/mb_math.cuh) is plain C++ when
compiled by g++.  Build it on the host and require BIT-EXACT agreement with the oracle: this is what makes
statuses / normals / poses reproducible between the GPU path and the CPU restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim():
    src = os.path.join(HERE, "host_shim", "shim.cpp")
    out = os.path.join(HERE, "host_shim", "libshim.so")
    hdr = os.path.join(HERE, "..", "mimosa_b200", "csrc", "mb_math.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", src, "-o", out],
                       check=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_eigh33_bit_exact(shim, oracle):
    rng = np.random.default_rng(42)
    cases = []
    for _ in range(3000):
        A = rng.normal(size=(3, 3))
        cases.append(A @ A.T * 10 ** rng.uniform(-8, 2))
    # covariance-like inputs of 5 nearly coplanar points (the plane-fit regime), diagonal and degenerate inputs
    for _ in range(3000):
        P = rng.uniform(-0.5, 0.5, (5, 3)) * [1, 1, 0.01]
        Cc = P - P.mean(0)
        cases.append(Cc.T @ Cc / 4)
    cases += [np.diag([3.0, 1.0, 2.0]), np.zeros((3, 3)), np.eye(3), np.diag([1e-300, 1.0, 1e300]),
              np.array([[2.0, 1, 0], [1, 2, 0], [0, 0, 5]]), np.array([[1.0, 0, 1e-200], [0, 1, 0], [1e-200, 0, 1]])]
    for A in cases:
        A = np.ascontiguousarray(A)
        lam, V = np.empty(3), np.empty(9)
        ok = shim.shim_eigh33(_p(A), _p(lam), _p(V))
        ok_o, lam_o, V_o = oracle.eigh3(A)
        assert bool(ok) == ok_o
        assert lam.tobytes() == lam_o.tobytes()
        assert V.tobytes() == V_o.reshape(9).tobytes()


def test_retract_solve_floor_bit_exact(shim, oracle):
    rng = np.random.default_rng(43)
    shim.shim_fast_floor.argtypes = [C.c_double]
    for _ in range(500):
        xi = rng.normal(size=6) * 10 ** rng.uniform(-10, 0.5)
        R0, t0 = oracle.se3_expmap(rng.normal(size=6))
        R, t = R0.reshape(9).copy(), t0.copy()
        shim.shim_se3_retract(_p(R), _p(t), _p(xi))
        Ro, to = oracle.se3_retract(R0, t0, xi)
        assert R.tobytes() == Ro.reshape(9).tobytes() and t.tobytes() == to.tobytes()
        A = rng.normal(size=(6, 6))
        H = np.ascontiguousarray(A @ A.T)
        b = rng.normal(size=6)
        x = np.zeros(6)
        shim.shim_solve6.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        ok = shim.shim_solve6(_p(H), 0.1, _p(b), _p(x))
        ok_o, x_o = oracle.solve6(H, 0.1, b)
        assert bool(ok) == ok_o and x.tobytes() == x_o.tobytes()
        v = float(rng.normal() * 100)
        assert shim.shim_fast_floor(v) == oracle.lib().orc_fast_floor(v) == int(np.floor(v))


def test_eigh33_direct_accuracy(shim):
    """The closed-form solver used for the 6x6-level localizability outputs: eigenvalues to 1e-12 of the
    spectral radius, eigenvectors orthonormal and A V = V diag(lam) to 1e-9, on random, planar-scene-like
    (two small, one large eigenvalue) and repeated-eigenvalue inputs."""
    rng = np.random.default_rng(7)
    cases = []
    for _ in range(2000):
        Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        ev = 10 ** rng.uniform(-3, 6, 3)
        cases.append(Q @ np.diag(ev) @ Q.T)
    for _ in range(500):  # planar scene: (a, a(1+eps), huge)
        Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        a = 10 ** rng.uniform(0, 3)
        cases.append(Q @ np.diag([a, a * (1 + 10 ** rng.uniform(-4, -1)), a * 1e3]) @ Q.T)
    cases += [np.diag([3.0, 1.0, 2.0]), np.eye(3) * 5, np.diag([1.0, 1.0, 7.0]), np.diag([2.0, 9.0, 9.0])]
    n_fast = 0
    for A in cases:
        A = np.ascontiguousarray((A + A.T) / 2)
        lam, V = np.empty(3), np.empty(9)
        n_fast += shim.shim_eigh33_direct(_p(A), _p(lam), _p(V))
        V = V.reshape(3, 3)
        w = np.linalg.eigvalsh(A)
        rad = np.abs(w).max()
        assert np.all(np.diff(lam) >= -1e-12 * rad)
        assert np.allclose(lam, w, rtol=1e-9, atol=1e-14 * rad), (lam, w)
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-9)
        gap_ok = np.min(np.diff(w)) > 1e-6 * rad
        if gap_ok:
            assert np.allclose(A @ V, V * lam, atol=1e-8 * rad)
    assert n_fast > 0.8 * len(cases)  # the closed form answers the bulk; the rest falls back to the QR solver
