"""Development: phase times of the persistent loop (library built with MB_NVCC_EXTRA=-DMB_LOOP_TIMING): SM clock of
block 0 at the phase boundaries of every linearisation of one bench scan."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch  # noqa: F401

import bench
import synth
from mimosa_b200 import HORNBILL_MAP, Context, ICPFactor, IncrementalVoxelMap, capi, hornbill_config



def print_table(lib, n_it=20):
    buf = (ctypes.c_longlong * (64 * 12))()
    assert lib.mb_debug_loop_times(buf) == 0
    t = np.array(buf, dtype=np.int64).reshape(64, 12)[:n_it]
    mhz = 1965.0
    print("it   total |  A+C  part  bar1 |    B   bar2    C'  p+b3 |  sum  xchg   fin   (us, block 0)", file=sys.stderr)
    for it in range(n_it):
        r = t[it]
        searched = r[4] > r[3] and r[4] != 0 and r[7] > r[3]
        a = (r[1] - r[0]) / mhz
        pw = (r[2] - r[1]) / mhz
        b1 = (r[3] - r[2]) / mhz
        if searched:
            B = (r[4] - r[3]) / mhz
            b2 = (r[5] - r[4]) / mhz
            c2 = (r[6] - r[5]) / mhz
            b3 = (r[7] - r[6]) / mhz
            s0 = r[7]
        else:
            B = b2 = c2 = b3 = 0.0
            s0 = r[3]
        sm = (r[8] - s0) / mhz
        xc = (r[10] - r[8]) / mhz
        fin = (r[9] - r[10]) / mhz
        tot = (r[9] - r[0]) / mhz
        print(f"{it:2d} {tot:7.1f} | {a:5.1f} {pw:5.1f} {b1:5.1f} | {B:5.1f} {b2:5.1f} {c2:5.1f} {b3:5.1f} | {sm:5.1f} {xc:5.1f} {fin:5.1f}", file=sys.stderr)


def print_fine(lib, n_it=20):
    buf = (ctypes.c_longlong * (64 * 12))()
    assert lib.mb_debug_loop_fine(buf) == 0
    t = np.array(buf, dtype=np.int64).reshape(64, 12)[:n_it]
    mhz = 1965.0
    print("it | loads+gate  fold  ballot+sync  C:loads  C:math+stores  acc0  acc1+counts | fin: pack  solve  (eig role 0 done after)", file=sys.stderr)
    for it in range(n_it):
        r = t[it]
        d = [(r[k + 1] - r[k]) / mhz for k in range(7)]
        fin = [(r[9] - r[8]) / mhz, (r[10] - r[9]) / mhz, (r[11] - r[8]) / mhz]
        print(f"{it:2d} | " + "  ".join(f"{x:6.2f}" for x in d) + " | " + "  ".join(f"{x:6.2f}" for x in fin), file=sys.stderr)


if __name__ == "__main__":
    ctx = Context(0)
    rng, scan, R0, t0, _, _ = bench.make_inputs()
    mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
    synth.build_map(mg.insert, bench.MAP_POINTS, bench.MAP_HALF_EXTENT, rng, size_fn=lambda: mg.size()[1])
    div = int(sys.argv[sys.argv.index("--shard") + 1]) if "--shard" in sys.argv else 1
    f = ICPFactor(ctx, mg, scan, hornbill_config(), (0, scan.shape[0] // div))
    lib = capi.load()
    for rep in range(4):
        f.reset()
        ctx.flush_l2()
        ctx.sync()
        f.icp_run(R0, t0, 20, 0.0, want_trace=False)
    print_table(lib, 20)
    print_fine(lib, 20)
