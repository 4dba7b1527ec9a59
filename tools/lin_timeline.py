"""Development (library built with MB_NVCC_EXTRA=-DMB_LIN_TIMING): %globaltimer timeline of every k_linearize launch of
one 20-iteration device-resident loop on the bench inputs."""
import ctypes as C
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
from mimosa_b200 import HORNBILL_MAP, Context, ICPFactor, IncrementalVoxelMap, hornbill_config

ctx = Context(0)
rng, scan, R0, t0, _, _ = bench.make_inputs()
mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
synth.build_map(mg.insert, bench.MAP_POINTS, bench.MAP_HALF_EXTENT, rng, size_fn=lambda: mg.size()[1])
f = ICPFactor(ctx, mg, scan, hornbill_config())
f.set_flags(cuda_graph="--graph" in sys.argv)
for rep in range(3):
    f.reset()
    ctx.flush_l2()
    ctx.sync()
    f.icp_run(R0, t0, 20, 0.0, want_trace=False)
buf = (C.c_ulonglong * 512)()
assert ctx.lib.mb_debug_lin_timeline(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(64, 8)
print("iter: entry->wait  wait->tiles  tiles->grp  grp->ticket  ticket->packet  packet->solve  (eigen roles)  | total  gap to next entry (us)")
for it in range(1, 21):
    r = t[it]
    d = lambda a, b: (r[b] - r[a]) / 1e3
    nxt = (t[it + 1][0] - r[6]) / 1e3 if it < 20 else float("nan")
    print(f"{it:3d}: {d(0,1):7.1f} {d(1,2):9.1f} {d(2,3):9.1f} {d(3,4):9.1f} {d(4,5):9.1f} {d(5,6):9.1f}   ({d(5,7):6.1f})  | {d(0,6):7.1f}  {nxt:7.1f}")
