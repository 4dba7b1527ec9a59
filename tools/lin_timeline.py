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
buf = (C.c_ulonglong * (512 + 4096 + 4096))()
assert ctx.lib.mb_debug_lin_timeline(buf) == 0
allb = np.array(buf, dtype=np.int64)
t = allb[:512].reshape(64, 8)
print("iter: entry->wait   P1 (gate+cached)   P2 (search passes)   P3 (re-associated)   reduce->ticket   packet | kernel total (us)")
for it in range(1, 21):
    r = t[it]
    d = lambda a, b: (r[b] - r[a]) / 1e3 if r[a] and r[b] else float("nan")
    p2 = d(2, 3) if r[3] else 0.0
    p3 = d(3, 4) if r[3] else d(2, 4)
    print(f"{it:3d}: {d(0,1):7.1f} {d(1,2):12.1f} {p2:18.1f} {p3:18.1f} {d(4,5):16.1f} {d(5,6):10.1f} | {d(0,6):8.1f}")
