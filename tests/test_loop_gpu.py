"""GPU tests of the persistent loop kernel's own mechanics (mb_factor.cu: k_icp_loop) against the CPU oracle:
  * a scan large enough that a group of the kernel owns SEVERAL tiles (the folded localizability pass and the final
    pass then read the points' state back from memory instead of the group's shared scratch);
  * the resident window of the host-facing call (mb_set_resident_window): a sequence of mb_factor_linearize calls
    must give bit-identical results whether each call launches or posts its pose to a resident kernel, also with
    other calls interleaved (they close the window) and with two factors alternating on one context."""
import time

import numpy as np
import pytest

import synth
from helpers import assert_linearization_close, assert_state_equal, g_err, lin_fields, rel_err
from mimosa_b200 import HORNBILL_MAP, ICPFactor, IncrementalVoxelMap, gn_step, hornbill_config

pytestmark = pytest.mark.gpu


def _case(ctx, oracle, n_scan, seed, n_map=400000, half=50.0):
    rng = np.random.default_rng(seed)
    mg, mo = IncrementalVoxelMap(ctx, **HORNBILL_MAP), oracle.IVoxRef(**HORNBILL_MAP)
    for _ in range(n_map // 100000):
        pts = synth.sample_world(100000, half, rng)
        mg.insert(pts)
        mo.insert(pts)
    R_true, t_true = synth.rot_from_rpy(0.0, 0.01, -0.3), np.array([-2.0, 1.5, 0.2])
    scan = synth.make_scan(R_true, t_true, n_scan, rng, max_range=half * 0.9)
    R0, t0 = synth.perturbed_start(R_true, t_true)
    return mg, mo, scan, R0, t0


def test_several_tiles_per_group(ctx, oracle):
    """262 144 points: more than 148 SMs x 4 groups x 256 points, so groups walk two tiles per linearisation."""
    mg, mo, scan, R0, t0 = _case(ctx, oracle, 262144, 7)
    cfg = hornbill_config()
    fg, fo = ICPFactor(ctx, mg, scan, cfg), oracle.IcpFactorRef(mo, scan, cfg)
    Rg, tg, trg = fg.icp_run(R0, t0, 5, 0.0)
    Ro, to, tro, _ = fo.icp_run(R0, t0, 5, 0.0, n_threads=0)
    for it, (a, b) in enumerate(zip(trg, tro)):
        assert list(a.counts) == list(b.counts), it
        assert a.n_searched == b.n_searched
        assert rel_err(a.H, b.H) <= 1e-7 and g_err(a.g, b.g, b.H, b.f) <= 1e-7
        assert np.allclose(np.array(a.loc_trans_comp), np.array(b.loc_trans_comp), rtol=1e-4), it
        assert np.allclose(np.array(a.loc_rot_comp), np.array(b.loc_rot_comp), rtol=1e-4), it
    assert np.abs(Rg - Ro).max() <= 1e-8 and np.abs(tg - to).max() <= 1e-8
    assert_state_equal(fg.download_state(), fo.download_state(), float_tol=1e-9)
    # and the host-facing call on the same factor
    fg.reset()
    fo.reset()
    assert_linearization_close(fg.linearize(R0, t0), fo.linearize(R0, t0, n_threads=0))
    fg.release()
    mg.release()


def _lin_vec(L):
    d = lin_fields(L)
    return np.concatenate([np.ravel(np.asarray(d[k], np.float64)) for k in sorted(d)])


def test_resident_window_is_transparent(ctx, oracle):
    mg, mo, scan, R0, t0 = _case(ctx, oracle, 20000, 8)
    cfg = hornbill_config()
    fo = oracle.IcpFactorRef(mo, scan, cfg)

    def sequence(window_us, interleave):
        ctx.set_resident_window(window_us)
        f = ICPFactor(ctx, mg, scan, cfg)
        R, t = R0.copy(), t0.copy()
        out = []
        for it in range(8):
            L = f.linearize(R, t)
            out.append(_lin_vec(L))
            if interleave and it == 3:
                f.download_state(knn=False)  # closes the window; the next call launches again
            if interleave and it == 5:
                time.sleep(0.002)  # the window closes on its own; the next call finds the kernel gone
            R, t, _, _ = gn_step(L, R, t, 0.0)
        st = f.download_state()
        f.release()
        return np.array(out), st, (R, t)

    base, st0, pose0 = sequence(0, False)
    for window, inter in ((30, False), (30, True), (2000, True), (1, False)):
        got, st, pose = sequence(window, inter)
        assert np.array_equal(base, got), (window, inter)
        assert_state_equal(st, st0)
        assert np.array_equal(pose[0], pose0[0]) and np.array_equal(pose[1], pose0[1])
    # against the oracle, first linearisation
    fo.reset()
    ctx.set_resident_window(30)
    f = ICPFactor(ctx, mg, scan, cfg)
    assert_linearization_close(f.linearize(R0, t0), fo.linearize(R0, t0, n_threads=0))
    f.release()
    mg.release()


def test_two_factors_alternate_on_one_context(ctx, oracle):
    mg, mo, scan, R0, t0 = _case(ctx, oracle, 12000, 9)
    cfg = hornbill_config()
    ctx.set_resident_window(200)
    fa, fb = ICPFactor(ctx, mg, scan, cfg), ICPFactor(ctx, mg, scan[::2].copy(), cfg)
    ctx.set_resident_window(0)
    ra, rb = ICPFactor(ctx, mg, scan, cfg), ICPFactor(ctx, mg, scan[::2].copy(), cfg)
    R, t = R0.copy(), t0.copy()
    for it in range(4):
        ctx.set_resident_window(200)
        La, Lb = fa.linearize(R, t), fb.linearize(R, t)  # fb's launch asks fa's resident kernel to leave
        La2 = fa.linearize(R, t)
        ctx.set_resident_window(0)
        Ra, Rb = ra.linearize(R, t), rb.linearize(R, t)
        Ra2 = ra.linearize(R, t)
        assert np.array_equal(_lin_vec(La), _lin_vec(Ra)) and np.array_equal(_lin_vec(Lb), _lin_vec(Rb))
        assert np.array_equal(_lin_vec(La2), _lin_vec(Ra2))
        R, t, _, _ = gn_step(La2, R, t, 0.0)
    ctx.set_resident_window(30)
    for f in (fa, fb, ra, rb):
        f.release()
    mg.release()
