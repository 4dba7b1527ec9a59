"""Development: host wall time of the host-facing mb_factor_linearize call on the bench inputs — one fully cached call
(same pose every time) — through ctypes (adds ~1 us of Python per call) and through the C++ e2e caller."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch  # noqa: F401

import bench
import synth
from mimosa_b200 import HORNBILL_MAP, Context, ICPFactor, IncrementalVoxelMap, hornbill_config

ctx = Context(0)
rng, scan, R0, t0, _, _ = bench.make_inputs()
mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
synth.build_map(mg.insert, bench.MAP_POINTS, bench.MAP_HALF_EXTENT, rng, size_fn=lambda: mg.size()[1])
div = int(sys.argv[sys.argv.index("--shard") + 1]) if "--shard" in sys.argv else 1
f = ICPFactor(ctx, mg, scan, hornbill_config(), (0, scan.shape[0] // div))
for _ in range(5):
    f.linearize(R0, t0)
for rep in range(3):
    n = 500
    a = time.perf_counter()
    for _ in range(n):
        f.linearize(R0, t0)
    b = time.perf_counter()
    print(f"cached host-facing call: {(b - a) / n * 1e6:.1f} us per call (wall, {n} calls)", flush=True)
# device time of the same launch: 200 launches back to back would need an async entry point; use the loop of one iteration
ms = []
for rep in range(5):
    ctx.sync()
    ctx.timer_begin()
    f.icp_run(R0, t0, 1, 0.0, want_trace=False)
    ms.append(ctx.timer_end())
print(f"device time of one cached 1-iteration loop launch (events, incl. pose copy in / out): {np.median(ms) * 1e3:.1f} us")

import ctypes
lib = ctx.lib
if hasattr(lib, "mb_debug_loop_times"):
    f.linearize(R0, t0)
    buf = (ctypes.c_longlong * (64 * 12))()
    lib.mb_debug_loop_times(buf)
    t = np.array(buf, dtype=np.int64).reshape(64, 12)
    print(f"device: kernel entry -> result handed to the host: {(t[63][1] - t[63][0]) / 1e3:.1f} us (globaltimer)")
    r = t[0]
    mhz = 1965.0
    print("phases (us): A+C %.1f part %.1f bar %.1f sum %.1f fin %.1f" % ((r[1]-r[0])/mhz, (r[2]-r[1])/mhz, (r[3]-r[2])/mhz, (r[8]-r[3])/mhz, (r[9]-r[10])/mhz))
    # resident path: 20 calls in a row, stamps of the last hand-over (globaltimer, ns)
    for _ in range(20):
        f.linearize(R0, t0)
    lib.mb_debug_loop_times(buf)
    t = np.array(buf, dtype=np.int64).reshape(64, 12)
    r = t[62]
    print("resident (us): response n-1 -> response n %.1f | response n-1 -> request seen %.1f | pose read %.1f | published + all blocks released %.1f | linearisation + hand-over %.1f"
          % ((r[0] - r[4]) / 1e3, (r[1] - r[4]) / 1e3, (r[2] - r[1]) / 1e3, (r[3] - r[2]) / 1e3, (r[0] - r[3]) / 1e3))
