"""Development: inputs for ncu captures of k_icp_loop.  mode `cached`: six host-facing linearisations at one pose (the
launches after the first are fully cached: phases A + C, D and the localizability pass only); mode `loop`: three
20-iteration device-resident scans."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: F401

import bench
import synth
from mimosa_b200 import HORNBILL_MAP, Context, ICPFactor, IncrementalVoxelMap, hornbill_config

mode = sys.argv[1] if len(sys.argv) > 1 else "cached"
ctx = Context(0)
rng, scan, R0, t0, _, _ = bench.make_inputs()
mg = IncrementalVoxelMap(ctx, **HORNBILL_MAP)
synth.build_map(mg.insert, bench.MAP_POINTS, bench.MAP_HALF_EXTENT, rng, size_fn=lambda: mg.size()[1])
f = ICPFactor(ctx, mg, scan, hornbill_config())
if mode == "cached":
    for _ in range(6):
        f.linearize(R0, t0)
else:
    for _ in range(3):
        f.reset()
        ctx.sync()
        f.icp_run(R0, t0, 20, 0.0, want_trace=False)
print("done")
