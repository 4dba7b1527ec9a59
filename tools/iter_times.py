"""Development: device time of the first n iterations of the device-resident loop (bench inputs), n = 1, 2, ...:
differences give the cost of each iteration (sort + searches first, cached iterations later)."""
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
div = int(sys.argv[sys.argv.index("--shard") + 1]) if "--shard" in sys.argv else 1  # time one rank's share of the scan
f = ICPFactor(ctx, mg, scan, hornbill_config(), (0, scan.shape[0] // div))
print(f"points {scan.shape[0] // div}", flush=True)
flush = "--no-flush" not in sys.argv
prev = 0.0
for iters in (1, 2, 3, 4, 5, 6, 10, 20):
    ms = []
    for rep in range(6):
        f.reset()
        if flush:
            ctx.flush_l2()
        ctx.sync()
        ctx.timer_begin()
        f.icp_run(R0, t0, iters, 0.0, want_trace=False)
        ms.append(ctx.timer_end())
    m = float(np.median(ms[2:])) * 1e3
    print(f"iters {iters:2d}: {m:8.1f} us   (+{m - prev:7.1f} us over the previous row)", flush=True)
    prev = m
